#!/usr/bin/env python
"""Markdown table of the main `ncu --set full` numbers of every captured launch.   python tools/ncu_summary.py REPORT.ncu-rep"""
import csv
import io
import subprocess
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", 1e-3, "ns"),
    ("grid", "launch__grid_size", 1, None),
    ("regs", "launch__registers_per_thread", 1, None),
    ("dram read MB", "dram__bytes_read.sum", None, "bytes"),
    ("dram write MB", "dram__bytes_write.sum", None, "bytes"),
    ("sm %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, None),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1, None),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, None),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1, None),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1, None),
    ("warp insts M", "smsp__inst_executed.sum", 1e-6, None),
    ("sm active cycles min/avg/max k", None, None, None),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        v = float(r[col[m]].replace(",", ""))
        return v * UNIT.get(units[col[m]], 1.0)

    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1]
        out = []
        for title, m, scale, kind in COLS:
            if m is None:
                out.append("/".join(f"{val(r, 'sm__cycles_active.' + s) / 1e3:.0f}" for s in ("min", "avg", "max")))
            elif kind == "bytes":
                out.append(f"{val(r, m) / 1e6:.1f}")
            elif kind == "ns":
                out.append(f"{val(r, m) / 1e3:.1f}")
            else:
                v = val(r, m) * scale
                out.append(f"{v:.1f}" if v != int(v) else f"{int(v)}")
        print(f"| {short} | " + " | ".join(out) + " |")


if __name__ == "__main__":
    main()
