#!/usr/bin/env python
"""Re-derive the fields of a saved bench.py line that bench.py COPIES from profiles/ncu_traffic.json (roofline.traffic, issue /
occupancy figures, per-kernel warp instructions and issue-slot fractions) from the ncu capture of the same build.

bench.py runs before the capture of its own build is summarised (tools/measure_round.sh runs bench.py first, ncu afterwards), so the
line it prints carries those fields from the PREVIOUS capture.  Nothing measured by bench.py itself is touched.

    python tools/refresh_ncu_fields.py profiles/r02_v7_bench.json profiles/ncu_traffic.json
"""
import json
import sys


def main():
    line_path, traffic_path = sys.argv[1:3]
    d = json.load(open(line_path))
    tj = json.load(open(traffic_path))
    r = d["roofline"]
    dominant = r["kernel"]
    k = tj["kernels"][dominant]
    r["traffic"] = k["traffic_bytes"]
    r["traffic_source"] = "profiles/ncu_traffic.json (%s; %s) [refreshed by tools/refresh_ncu_fields.py from the capture of this build]" % (tj.get("source"), tj.get("note"))
    for f in ("issue_active_pct", "warps_active_pct", "warp_instructions", "registers"):
        if f in k:
            r[f] = k[f]
    sm_count, clock_hz = 148, 1e6 * float((d.get("clocks") or {}).get("sm_mhz") or 1965.0)
    frame_insts = 0.0
    for name, kk in r["kernels"].items():
        wi = tj["kernels"].get(name, {}).get("warp_instructions")
        if wi:
            kk["warp_instructions_ncu"] = wi
            kk["issue_slot_frac"] = wi / (sm_count * 4 * clock_hz * kk["ms_per_frame"] * 1e-3)
            frame_insts += wi
    if frame_insts and "issue" in r:
        per_frame_s = d["ms_per_step"] * 1e-3 / d["config"]["frames_per_step"]
        r["issue"]["frame_warp_instructions"] = frame_insts
        r["issue"]["frame_issue_slot_frac"] = frame_insts / (sm_count * 4 * clock_hz * per_frame_s)
    json.dump(d, open(line_path, "w"))
    print("traffic", r["traffic"], "warp instructions", r.get("warp_instructions"), "frame issue-slot frac", r.get("issue", {}).get("frame_issue_slot_frac"))


if __name__ == "__main__":
    main()
