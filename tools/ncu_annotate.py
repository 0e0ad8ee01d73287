#!/usr/bin/env python
"""Source lines of one kernel annotated with their share of executed warp instructions and stall samples.
    python tools/ncu_annotate.py REPORT KERNEL_REGEX CUBIN_SUBSTR MANGLED_SUBSTR FILE [min_pct]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre, cub, mang, fname = sys.argv[1:6]
minp = float(sys.argv[6]) if len(sys.argv) > 6 else 0.3
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kre, cub, mang, "1000"], capture_output=True, text=True).stdout
d = {}
other = {}
for l in out.splitlines():
    m = re.match(r"\('([^']+)', (\d+)\): inst ([\d.]+)%  samples ([\d.]+)%", l)
    if m:
        if m.group(1) == os.path.basename(fname):
            d[int(m.group(2))] = (float(m.group(3)), float(m.group(4)))
        else:
            other[(m.group(1), int(m.group(2)))] = (float(m.group(3)), float(m.group(4)))
    elif l.startswith(("kernel", "stall")):
        print(l.replace("stall_", ""))
src = open(fname).read().split("\n")
for i in sorted(d):
    if d[i][0] >= minp or d[i][1] >= minp:
        print(f"{i:5d} {d[i][0]:4.1f} {d[i][1]:4.1f} | {src[i-1].strip()[:120]}")
for k in sorted(other):
    if other[k][0] >= minp or other[k][1] >= minp:
        print(k, other[k])
