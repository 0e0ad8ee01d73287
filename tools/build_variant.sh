#!/bin/bash
# Development helper: build libsqw_cuda.so variants with extra -D flags for raster_kernels.cu into build_variants/NAME.so
# usage: tools/build_variant.sh NAME [-DFOO=1 ...]
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p build_variants/obj
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC"
for s in sqw_api surface_kernels post_kernels depth_kernels; do
  if [ ! -f build_variants/obj/$s.o ] || [ square_wheel_b200/csrc/$s.cu -nt build_variants/obj/$s.o ]; then
    nvcc $F -c square_wheel_b200/csrc/$s.cu -o build_variants/obj/$s.o &
  fi
done
nvcc $F "$@" -c square_wheel_b200/csrc/raster_kernels.cu -o build_variants/obj/raster_$NAME.o
wait
nvcc -shared -o build_variants/$NAME.so build_variants/obj/sqw_api.o build_variants/obj/surface_kernels.o build_variants/obj/post_kernels.o build_variants/obj/depth_kernels.o build_variants/obj/raster_$NAME.o
