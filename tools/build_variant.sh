#!/bin/bash
# Development helper: build a libsqw_cuda.so variant with extra -D flags (applied to every source) into build_variants/NAME.so
# usage: tools/build_variant.sh NAME [-DFOO=1 ...]
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p build_variants/obj/$NAME
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
OBJS=""
for s in sqw_api surface_kernels surface_kernels_fast post_kernels depth_kernels model_kernels polygon_kernels draw_setup_kernels readback_decode raster_kernels; do
  if [ $s = surface_kernels_fast ]; then M=-fmad=true; else M=-fmad=false; fi
  if [ $s = readback_decode ]; then M="$M -Xcompiler -mavx2"; fi
  nvcc $F $M "$@" -c square_wheel_b200/csrc/$s.cu -o build_variants/obj/$NAME/$s.o &
  OBJS="$OBJS build_variants/obj/$NAME/$s.o"
done
wait
nvcc -shared -o build_variants/$NAME.so $OBJS
