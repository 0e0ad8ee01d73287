#!/bin/bash
# On the GPU box: bench every build_variants/*.so (copied over the in-tree library) and print the per-kernel times.
for so in build_variants/*.so; do
  cp $so square_wheel_b200/libsqw_cuda.so
  n=$(basename $so .so)
  python -m pytest tests/test_surfaces_gpu.py tests/test_frames_gpu.py -x -q 2>&1 | tail -1
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bv_$n.json 2> gpurun_out/bv_$n.err || { echo "$n FAILED"; tail -3 gpurun_out/bv_$n.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bv_$n.json"))
k=d["roofline"]["kernels"]
print("$n", "value",round(d["value"],1), " ".join(f"{a}={round(v['ms_per_frame']*1000,2)}us" for a,v in k.items()))
PY
done
