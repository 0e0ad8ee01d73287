#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN_SUBSTR MANGLED_SUBSTR [top]

Joins `ncu --page source --print-source sass --csv` (per-SASS-instruction samples) with `nvdisasm -g` of the cubin extracted
from square_wheel_b200/libsqw_cuda.so (the -lineinfo line table), so that the hot lines of OUR source can be read here
without the GUI.  Used for the profiles/ summaries."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kregex, cubin_sub, mangled_sub = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "square_wheel_b200", "libsqw_cuda.so")], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_sub)][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    cur_fn = cur_line = None
    addr2line = {}
    for l in sass.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur_fn and mangled_sub in cur_fn:
            addr2line[int(m.group(1), 16)] = cur_line
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kregex, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0

    base = None
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    agg = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[idx["Address"]].startswith("0x"):
            if base is not None and r and r[0] == "Kernel Name":
                break  # next launch of the same kernel
            continue
        a = int(r[idx["Address"]], 16)
        base = a if base is None else base
        p = per[addr2line.get(a - base)]
        p[0] += I(r[idx["Instructions Executed"]])
        p[1] += I(r[idx["# Samples"]])
        for s in stalls:
            v = I(r[idx[s]])
            if v:
                p[2][s] += v
                agg[s] += v
    tot = sum(p[0] for p in per.values()) or 1
    ts = sum(p[1] for p in per.values()) or 1
    print(f"kernel {kregex}: {tot} warp instructions, {ts} samples")
    print("stall reasons (% of samples):", ", ".join(f"{s[6:]} {100 * v / ts:.1f}" for s, v in agg.most_common(8)))
    for ln, p in sorted(per.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{ln}: inst {100 * p[0] / tot:.1f}%  samples {100 * p[1] / ts:.1f}%  {dict(p[2].most_common(3))}")


if __name__ == "__main__":
    main()
