// Stage 1 with the numerics class of the reference's release build (SQW_CTX_FAST_SURFACE_NUMERICS, include/sqw_cuda.h): the same
// source as surface_kernels.cu, with inv_fast / inv_sqrt_fast / f32_mul_add mapped to MUFU.RCP / MUFU.RSQ / FFMA (sqw_device.cuh).
// Compiled with -fmad=true (__graft_entry__.py); exports launch_surface_build_fast.
#define SQW_NUMERICS_FAST 1
#include "surface_kernels.cu"
