#!/bin/bash
# compute-sanitizer over the GPU parity tests of the kernels changed this round (memcheck + racecheck); logs in gpurun_out/sanitize_*.log
T1="tests/test_depth_gpu.py tests/test_raster_gpu.py tests/test_models_gpu.py"
T2="tests/test_polygons_gpu.py::test_polygons_prepare_matches_the_host_mirror_bit_for_bit tests/test_fused_gpu.py tests/test_surfaces_gpu.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T1 -x -q -k "not self_test" > gpurun_out/sanitize_${tool}_1.log 2>&1; echo "$tool set 1 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_1.log | tail -3
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T2 -x -q -k "not bench_scene and not 11000 and not 5000" > gpurun_out/sanitize_${tool}_2.log 2>&1; echo "$tool set 2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_2.log | tail -3
done
