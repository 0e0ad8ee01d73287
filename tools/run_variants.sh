#!/bin/bash
# On the GPU box: bench every build_variants/*.so (copied over the in-tree library) and print the per-kernel times.
# usage: [CONFIG_ONLY=substring] tools/run_variants.sh [test files...]   (default: the fused-path and raster tests as a parity gate per variant)
TESTS=${@:-tests/test_fused_gpu.py tests/test_raster_gpu.py}
cp square_wheel_b200/libsqw_cuda.so /tmp/libsqw_cuda.keep
for so in build_variants/*.so; do
  cp $so square_wheel_b200/libsqw_cuda.so
  n=$(basename $so .so)
  timeout 300 python -m pytest $TESTS -x -q 2>&1 | tail -1
  timeout 300 python bench.py --steps 6 --warmup 3 --kernels-only > gpurun_out/bv_$n.json 2> gpurun_out/bv_$n.err || { echo "$n FAILED"; tail -3 gpurun_out/bv_$n.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bv_$n.json"))
k=d["roofline"]["kernels"]
print("$n", "value",round(d["value"],1), " ".join(f"{a}={round(v['ms_per_frame']*1000,2)}us" for a,v in k.items()))
PY
  if [ -n "$CONFIG_ONLY" ]; then
    timeout 300 python tools/bench_configs.py --only "$CONFIG_ONLY" --frames 16 --repeats 2 --out gpurun_out/bvc_$n.json > /dev/null 2>&1 && python - <<PY
import json
d=json.load(open("gpurun_out/bvc_$n.json"))
for name,c in d["configs"].items():
    print("   ", name[:50], round(c["frames_per_sec"],1), "fps", " ".join(f"{a}={round(v['us_per_frame'],1)}" for a,v in c["kernels"].items()))
PY
  fi
done
cp /tmp/libsqw_cuda.keep square_wheel_b200/libsqw_cuda.so
