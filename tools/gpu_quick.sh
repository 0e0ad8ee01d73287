#!/bin/bash
# Development iteration on the GPU box: parity tests, kernels-only bench of configs[1], the named configs.  usage: tools/gpu_quick.sh TAG [only]
TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/q_$TAG.json 2> gpurun_out/q_$TAG.err || tail -5 gpurun_out/q_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$TAG.json"))
k=d["roofline"]["kernels"]
print("configs[1] value",round(d["value"],1), " ".join(f"{a}={round(v['ms_per_frame']*1000,2)}us" for a,v in k.items()))
PY
timeout 900 python tools/bench_configs.py --out gpurun_out/qc_$TAG.json ${2:+--only "$2"} > gpurun_out/qc_$TAG.log 2>&1 || tail -5 gpurun_out/qc_$TAG.log
python - <<PY
import json
d=json.load(open("gpurun_out/qc_$TAG.json"))
for n,c in d["configs"].items():
    print(n[:60], round(c["frames_per_sec"],1), "fps", " ".join(f"{a}={round(v['us_per_frame'],1)}" for a,v in c["kernels"].items()))
    g=c["front_end_geometry"]; print("     geometry: candidates", round(g["candidates_per_frame"]), "host geometry us", round(g["host_mirror_us_per_frame"],1), "host front end us", round(g["host_mirror_whole_front_end_us_per_frame"],1), "gpu kernels us", round(g["gpu_kernels_us_per_frame"],1), "gpu call us", round(g["gpu_call_us_per_frame"],1), "bit-exact", g["bit_exact_vs_host_mirror"], "chain", {k: round(v, 1) for k, v in (g.get("chain") or {}).items() if k != "note"})
PY
