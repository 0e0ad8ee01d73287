#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an `ncu --set full` report -> profiles/ncu_traffic.json (read by bench.py for
`roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/prof_v7.ncu-rep profiles/ncu_traffic.json

traffic = dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of a kernel.  The capture runs
with ncu's default cache control (caches flushed before every replay), so these are cold-cache numbers."""
import collections
import csv
import io
import json
import subprocess
import sys

NAMES = {"surface_build_kernel": "surface_build", "raster_setup_kernel": "raster_setup", "raster_tile_kernel": "raster_tile", "raster_tile_opaque_kernel": "raster_tile", "tonemap_kernel": "tonemap"}
# issue / occupancy figures bench.py copies into `roofline` next to the traffic (the tile kernel is issue-bound, not DRAM-bound)
EXTRA = {
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers": "launch__registers_per_thread",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1:3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum," + ",".join(EXTRA.values())],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    # per (stage, kernel flavour): the tile stage of a frame is up to two launches, the opaque tile kernel (every frame) and the
    # general one (frames with triangles / blended polygons); their bytes, time and instructions ADD UP per frame, while the
    # occupancy figures quoted in `roofline` are the opaque kernel's (it is the dominant one)
    acc = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    extra = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = next((v for k, v in NAMES.items() if k in name), None)
        if key is None:
            continue
        if key == "raster_tile" and "1>" in name.replace(" ", "").split("_kernel")[1][:40]:
            key = "raster_tile_tonemap"
        flavour = "opaque" if "raster_tile_opaque_kernel" in name else ("general" if "raster_tile_kernel" in name else "")
        a = acc[(key, flavour)]
        for j, m in enumerate(["dram__bytes_read.sum", "dram__bytes_write.sum"]):
            a[j] += float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
        a[2] += float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
        a[3] += 1
        for k, m in EXTRA.items():
            if m in col:
                try:
                    extra[(key, flavour)][k] += float(r[col[m]].replace(",", ""))
                except ValueError:
                    pass
    res = {"source": rep, "note": "ncu --set full, default cache control (cold caches); per launch for single-kernel stages; for the tile stage per FRAME (opaque + general launch summed; issue / warps active / registers are the opaque kernel's), with both kernels also listed per launch", "kernels": {}}
    stages = sorted({k for k, _ in acc})
    for k in stages:
        parts = {f: acc[(k, f)] for f in ("", "opaque", "general") if (k, f) in acc}
        frames = parts["opaque"][3] if "opaque" in parts else max(p[3] for p in parts.values())
        rd, wr, us = (sum(p[j] for p in parts.values()) for j in range(3))
        e = {"launches": sum(p[3] for p in parts.values()), "frames": frames, "dram_read_bytes": rd / frames, "dram_write_bytes": wr / frames, "traffic_bytes": (rd + wr) / frames, "us_under_ncu": us / frames}
        main_f = "opaque" if "opaque" in parts else next(iter(parts))
        for kk, v in extra[(k, main_f)].items():
            e[kk] = v / parts[main_f][3]
        e["warp_instructions"] = sum(extra[(k, f)].get("warp_instructions", 0.0) for f in parts) / frames
        if len(parts) > 1 or "opaque" in parts:
            for f, p in parts.items():
                e[f + "_kernel"] = {"launches": p[3], "traffic_bytes": (p[0] + p[1]) / p[3], "us_under_ncu": p[2] / p[3], **{kk: v / p[3] for kk, v in extra[(k, f)].items()}}
        res["kernels"][k] = e
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(res["kernels"], indent=1))


if __name__ == "__main__":
    main()
