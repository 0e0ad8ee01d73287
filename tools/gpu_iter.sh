#!/bin/bash
# One build -> measure iteration on the GPU box: parity tests, bench line, per-kernel summary.  Usage: tools/gpu_iter.sh TAG [ncu]
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_$TAG.log
python bench.py > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err
cat gpurun_out/t_$TAG.log
python - <<PY
import json
d=json.load(open("gpurun_out/b_$TAG.json"))
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1))
for k,v in d["roofline"]["kernels"].items(): print(k, round(v["ms_per_frame"]*1000,2),"us", round(v.get("frac_of_measured_peak",0),3))
PY
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on --launch-skip 150 -c 9 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 > gpurun_out/p_$TAG.log 2>&1
fi
