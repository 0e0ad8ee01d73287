#!/bin/bash
# Round-end measurement set on the GPU box (one GPU).  Usage: tools/measure_round.sh TAG
#   parity tests, smoke, bench line, reference-arm line, the other named configs, ncu launch list of the bench command,
#   ncu --set full capture of two frames.  Everything lands in gpurun_out/*_TAG.* ; summaries are copied to profiles/ afterwards.
TAG=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_$TAG.log
cat gpurun_out/t_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 900 python tools/bench_configs.py --out gpurun_out/configs_$TAG.json > gpurun_out/configs_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --kernels-only > gpurun_out/launches_$TAG.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"surface_build|raster_tile|raster_setup" --launch-skip 300 -c 6 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --kernels-only > gpurun_out/prof_$TAG.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"frac",round(d["roofline"]["frac"],3),"frame frac",round(d["roofline"]["frame_frac_total"],3), "d2h ceiling", d["e2e"]["d2h_ceiling"]["pictures_per_s"])
for k,v in d["roofline"]["kernels"].items(): print(k, round(v["ms_per_frame"]*1000,2),"us", round(v.get("frac_of_measured_peak",0),3))
print("fast surfaces", round(d["fast_surface_numerics"]["value"],1), "view batch", round(d["view_batch"]["value"],1), "cpu", round(d["cpu_baseline"]["value"],1), d["cpu_baseline"]["stage_ms_per_frame"])
r=json.load(open("gpurun_out/bench_ref_$TAG.json")); print("reference arm", round(r["value"],1), r["cpu_baseline"]["cores"], "threads")
PY
