// Device-side pixel and fixed-point primitives for the sm_100a world-drawing kernels.
//
// Semantics follow the PORTABLE branch of the reference's src/square_wheel/fast_math.rs (:370-725),
// src/square_wheel/abstract_color.rs, src/common/color.rs and src/common/fixed_math.rs:
//   - inv_fast = IEEE 1/x, inv_sqrt_fast = IEEE 1/sqrt(x), f32_mul_add = x*y+z with two roundings
//     (this translation unit is compiled with -fmad=false so nvcc never contracts them),
//   - float -> int conversions truncate,
//   - integer colour math wraps like Rust release builds.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sqw
{

// ---- fixed_math.rs -----------------------------------------------------------------------------------------
typedef int32_t Fixed16;
#define SQW_FIXED16_ONE 65536
#define SQW_FIXED16_HALF 32768

__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
__device__ __forceinline__ int32_t wshl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }

__device__ __forceinline__ Fixed16 int_to_fixed16(int32_t x) { return wshl(x, 16); }                       // fixed_math.rs:6
__device__ __forceinline__ int32_t fixed16_round_to_int(Fixed16 x) { return wadd(x, SQW_FIXED16_HALF) >> 16; } // :11
__device__ __forceinline__ int32_t fixed16_floor_to_int(Fixed16 x) { return x >> 16; }                     // :16
__device__ __forceinline__ float fixed16_to_f32(Fixed16 x) { return (float)x / 65536.0f; }                 // :26
__device__ __forceinline__ Fixed16 fixed16_mul(Fixed16 x, Fixed16 y) { return (int32_t)(((int64_t)x * (int64_t)y) >> 16); } // :41
// fixed16_div (:56): ((x as i64) << 16) / (y as i64), truncated toward zero, `as i32`.  The plain 64-bit integer division
// (kept for sqw_self_test) costs well over a hundred instructions on the GPU and the triangle set-up needs two dozen of them
// per scan-line chunk, so the product path divides the magnitudes in double precision — |x| * 2^16 < 2^48 is exact in a
// double, the rounded quotient is off by less than 2^-4, so its truncation is within one of the true quotient — and
// repairs the last unit with the exact integer remainder.
__device__ __forceinline__ Fixed16 fixed16_div_reference(Fixed16 x, Fixed16 y)
{
	return (int32_t)((int64_t)((uint64_t)(int64_t)x << 16) / (int64_t)y);
}
// __noinline__: the set-up kernel calls it some forty times; inlined, its code no longer fits the instruction cache
static __device__ __noinline__ Fixed16 fixed16_div(Fixed16 x, Fixed16 y)
{
	if (y == 0)
		return fixed16_div_reference(x, y); // undefined in the reference (it panics); whatever the device division yields
	const uint64_t a = (uint64_t)(x < 0 ? -(int64_t)x : (int64_t)x) << 16;
	const uint64_t b = (uint64_t)(y < 0 ? -(int64_t)y : (int64_t)y);
	uint64_t q = (uint64_t)__double2ull_rz(__ddiv_rn(__ull2double_rn(a), __ull2double_rn(b)));
	const int64_t r = (int64_t)(a - q * b);
	if (r < 0)
		--q;
	else if ((uint64_t)r >= b)
		++q;
	return (int32_t)(((x < 0) != (y < 0)) ? -(int64_t)q : (int64_t)q);
}
__device__ __forceinline__ Fixed16 fixed16_mul_div(Fixed16 x, Fixed16 y, Fixed16 z) // :68
{
	return (int32_t)((int64_t)x * (int64_t)y / (int64_t)z);
}
// `to_int_unchecked` as the reference's x86 build executes it (cvttss2si: out of range -> INT_MIN).
__device__ __forceinline__ int32_t f32_to_i32_trunc(float x)
{
	if (!(x > -2147483904.0f && x < 2147483648.0f))
		return INT32_MIN;
	return __float2int_rz(x);
}
__device__ __forceinline__ int64_t f32_to_i64_trunc(float x) // rasterizer.rs:1552
{
	if (!(x > -9223373136366403584.0f && x < 9223372036854775808.0f))
		return INT64_MIN;
	return __float2ll_rz(x);
}
__device__ __forceinline__ Fixed16 f32_to_fixed16_unchecked(float x) { return f32_to_i32_trunc(x * 65536.0f); } // fixed_math.rs:36

// ---- fast_math.rs portable branch ---------------------------------------------------------------------------
#define SQW_MIN_POSITIVE_VALUE (1.0f / 1073741824.0f) /* fast_math.rs:744 */

#if defined(SQW_NUMERICS_FAST) && SQW_NUMERICS_FAST
// The numerics CLASS of the reference's release build (fast_math.rs:13-22 rcpss / rsqrtss, :729-733 fused multiply-add): MUFU.RCP /
// MUFU.RSQ (1 ulp; rcpss / rsqrtss promise 1.5 * 2^-12) and FFMA.  Only the translation unit of the fast surface kernels
// (surface_kernels_fast.cu, SQW_CTX_FAST_SURFACE_NUMERICS) is compiled this way.
__device__ __forceinline__ float inv_sqrt_fast(float x)
{
	float r;
	asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float inv_fast(float x)
{
	float r;
	asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float f32_mul_add(float x, float y, float z) { return __fmaf_rn(x, y, z); }
#else
// 1.0 / y correctly rounded IS the IEEE reciprocal: rcp.rn gives the same bits as div.rn(1, y) with a shorter sequence
__device__ __forceinline__ float inv_sqrt_fast(float x) { return __frcp_rn(__fsqrt_rn(x)); } // :376
__device__ __forceinline__ float inv_fast(float x) { return __frcp_rn(x); }                  // :381
__device__ __forceinline__ float f32_mul_add(float x, float y, float z) { return __fadd_rn(__fmul_rn(x, y), z); } // :737
#endif

struct Vec3
{
	float x, y, z;
};
__device__ __forceinline__ Vec3 v3(float x, float y, float z)
{
	Vec3 r;
	r.x = x;
	r.y = y;
	r.z = z;
	return r;
}
__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Vec3 operator*(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float vec3_dot(Vec3 a, Vec3 b) { return f32_mul_add(a.x, b.x, f32_mul_add(a.y, b.y, a.z * b.z)); } // :780
__device__ __forceinline__ Vec3 vec3_scalar_mul_add(Vec3 a, float s, Vec3 b) // :786
{
	return v3(f32_mul_add(a.x, s, b.x), f32_mul_add(a.y, s, b.y), f32_mul_add(a.z, s, b.z));
}
__device__ __forceinline__ float vec3_len2(Vec3 v) { return vec3_dot(v, v); }

// ColorVec lanes [B, G, R, A] (fast_math.rs:443).
struct ColorVec
{
	float b, g, r, a;
};
__device__ __forceinline__ ColorVec cv(float b, float g, float r, float a)
{
	ColorVec c;
	c.b = b;
	c.g = g;
	c.r = r;
	c.a = a;
	return c;
}
__device__ __forceinline__ ColorVec cv_from_color32(uint32_t c)
{
	return cv((float)(c & 0xFF), (float)((c >> 8) & 0xFF), (float)((c >> 16) & 0xFF), (float)(c >> 24));
}
__device__ __forceinline__ ColorVec cv_from_color64(uint64_t c)
{
	const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
	return cv((float)(lo & 0xFFFF), (float)(lo >> 16), (float)(hi & 0xFFFF), (float)(hi >> 16));
}
__device__ __forceinline__ ColorVec cv_from_f32x3(const float* c) { return cv(c[2], c[1], c[0], 0.0f); } // :505
__device__ __forceinline__ ColorVec cv_add(ColorVec x, ColorVec y) { return cv(x.b + y.b, x.g + y.g, x.r + y.r, x.a + y.a); }
__device__ __forceinline__ ColorVec cv_mul(ColorVec x, ColorVec y) { return cv(x.b * y.b, x.g * y.g, x.r * y.r, x.a * y.a); }
__device__ __forceinline__ ColorVec cv_scalar_mul(ColorVec x, float s) { return cv(x.b * s, x.g * s, x.r * s, x.a * s); }
__device__ __forceinline__ ColorVec cv_mul_scalar_add(ColorVec x, float s, ColorVec y) // :555
{
	return cv(f32_mul_add(x.b, s, y.b), f32_mul_add(x.g, s, y.g), f32_mul_add(x.r, s, y.r), f32_mul_add(x.a, s, y.a));
}

// into_color32 / into_color64 (:472-498): min(max) then truncate.  Negative input (UB in the reference's
// to_int_unchecked) is defined as 0 here and in the oracle; cvt.rzi.u32.f32 saturates exactly that way.
__device__ __forceinline__ uint32_t f32_to_u_clamped(float x, float maxv) { return __float2uint_rz(fminf(x, maxv)); }

// ColorVecI lanes [B, G, R, A] (fast_math.rs:594).
struct ColorVecI
{
	int32_t b, g, r, a;
};
__device__ __forceinline__ ColorVecI cvi(int32_t b, int32_t g, int32_t r, int32_t a)
{
	ColorVecI c;
	c.b = b;
	c.g = g;
	c.r = r;
	c.a = a;
	return c;
}
__device__ __forceinline__ ColorVecI cvi_from_color32(uint32_t c)
{
	return cvi((int32_t)(c & 0xFF), (int32_t)((c >> 8) & 0xFF), (int32_t)((c >> 16) & 0xFF), (int32_t)(c >> 24));
}
__device__ __forceinline__ ColorVecI cvi_from_color64(uint64_t c)
{
	const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
	return cvi((int32_t)(lo & 0xFFFF), (int32_t)(lo >> 16), (int32_t)(hi & 0xFFFF), (int32_t)(hi >> 16));
}
__device__ __forceinline__ ColorVecI cvi_add(ColorVecI x, ColorVecI y) { return cvi(wadd(x.b, y.b), wadd(x.g, y.g), wadd(x.r, y.r), wadd(x.a, y.a)); }
__device__ __forceinline__ ColorVecI cvi_mul(ColorVecI x, ColorVecI y) { return cvi(wmul(x.b, y.b), wmul(x.g, y.g), wmul(x.r, y.r), wmul(x.a, y.a)); }
__device__ __forceinline__ ColorVecI cvi_mul_scalar(ColorVecI x, int32_t s) { return cvi(wmul(x.b, s), wmul(x.g, s), wmul(x.r, s), wmul(x.a, s)); }
template <int N>
__device__ __forceinline__ ColorVecI cvi_shr(ColorVecI x)
{
	return cvi(x.b >> N, x.g >> N, x.r >> N, x.a >> N);
}
// `as i32` of an f32 (saturating, NaN -> 0): fast_math.rs:623
__device__ __forceinline__ int32_t f32_as_i32(float x) { return __float2int_rz(x); }

// ---- colour types (abstract_color.rs) --------------------------------------------------------------------------
struct Color32
{
	typedef uint32_t Raw;
	static __device__ __forceinline__ Raw from_vec(ColorVec c)
	{
		return f32_to_u_clamped(c.b, 255.0f) | (f32_to_u_clamped(c.g, 255.0f) << 8) | (f32_to_u_clamped(c.r, 255.0f) << 16) |
			(f32_to_u_clamped(c.a, 255.0f) << 24);
	}
	static __device__ __forceinline__ Raw from_veci(ColorVecI c) // (x as u32).min(255)
	{
		return min((uint32_t)c.b, 255u) | (min((uint32_t)c.g, 255u) << 8) | (min((uint32_t)c.r, 255u) << 16) | (min((uint32_t)c.a, 255u) << 24);
	}
	static __device__ __forceinline__ ColorVecI to_veci(Raw c) { return cvi_from_color32(c); }
	static __device__ __forceinline__ Raw average(Raw a, Raw b) { return (((a ^ b) & 0xFEFEFEFEu) >> 1) + (a & b); } // color.rs:46
	static __device__ __forceinline__ Raw saturated_sum(Raw a, Raw b)                                                 // fast_math.rs:386
	{
		Raw r = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			const int s = i * 8;
			r |= min(((a >> s) & 0xFF) + ((b >> s) & 0xFF), 0xFFu) << s;
		}
		return r;
	}
	static __device__ __forceinline__ int32_t get_alpha(Raw c) { return (int32_t)(c >> 24); }
	static __device__ __forceinline__ bool test_alpha(Raw c) { return (c & 0xFF000000u) > 0x7F000000u; }
};

struct Color64
{
	typedef uint64_t Raw;
	static __device__ __forceinline__ Raw from_vec(ColorVec c)
	{
		const uint32_t lo = f32_to_u_clamped(c.b, 65535.0f) | (f32_to_u_clamped(c.g, 65535.0f) << 16);
		const uint32_t hi = f32_to_u_clamped(c.r, 65535.0f) | (f32_to_u_clamped(c.a, 65535.0f) << 16);
		return (uint64_t)lo | ((uint64_t)hi << 32);
	}
	static __device__ __forceinline__ Raw from_veci(ColorVecI c) // (x as u64).min(65535): negatives saturate
	{
		const uint32_t lo = min((uint32_t)c.b, 65535u) | (min((uint32_t)c.g, 65535u) << 16);
		const uint32_t hi = min((uint32_t)c.r, 65535u) | (min((uint32_t)c.a, 65535u) << 16);
		return (uint64_t)lo | ((uint64_t)hi << 32);
	}
	static __device__ __forceinline__ ColorVecI to_veci(Raw c) { return cvi_from_color64(c); }
	static __device__ __forceinline__ Raw average(Raw a, Raw b) { return (((a ^ b) & 0xFFFEFFFEFFFEFFFEull) >> 1) + (a & b); } // color.rs:163
	static __device__ __forceinline__ Raw saturated_sum(Raw a, Raw b)                                                           // fast_math.rs:397
	{
		Raw r = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			const int s = i * 16;
			r |= (uint64_t)min((uint32_t)((a >> s) & 0xFFFF) + (uint32_t)((b >> s) & 0xFFFF), 0xFFFFu) << s;
		}
		return r;
	}
	static __device__ __forceinline__ int32_t get_alpha(Raw c) { return (int32_t)(c >> 48); }
	static __device__ __forceinline__ bool test_alpha(Raw c) { return (c & 0xFFFF000000000000ull) > 0x007F000000000000ull; }
};

template <class C>
__device__ __forceinline__ typename C::Raw alpha_blend(typename C::Raw dst, typename C::Raw src) // abstract_color.rs:30-40
{
	const int32_t alpha = C::get_alpha(src);
	return C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(C::to_veci(dst), 255 - alpha), cvi_mul_scalar(C::to_veci(src), alpha))));
}

// write_into_framebuffer (rasterizer.rs:1511-1536), blend mode uniform per command.
template <class C>
__device__ __forceinline__ typename C::Raw blend_pixel(int blend, typename C::Raw dst, typename C::Raw texel)
{
	switch (blend)
	{
	case 0: return texel;
	case 1: return C::average(dst, texel);
	case 2: return C::saturated_sum(dst, texel);
	case 3: return C::test_alpha(texel) ? texel : dst;
	default: return alpha_blend<C>(dst, texel);
	}
}

// TonemappingFunction::do_it (postprocessor.rs:1019-1023) for the three colour lanes; the alpha lane is 0/0 or x/0 in the
// reference and ends as 255 after `min(255.0)` in both of its implementations.
__device__ __forceinline__ uint32_t tonemap_pixel(ColorVec c, float inv_scale_b, float inv_scale_g, float inv_scale_r)
{
	const float inv_255 = 1.0f / 255.0f;
	const float b = __fdiv_rn(c.b, f32_mul_add(c.b, inv_255, inv_scale_b));
	const float g = __fdiv_rn(c.g, f32_mul_add(c.g, inv_255, inv_scale_g));
	const float r = __fdiv_rn(c.r, f32_mul_add(c.r, inv_255, inv_scale_r));
	return f32_to_u_clamped(b, 255.0f) | (f32_to_u_clamped(g, 255.0f) << 8) | (f32_to_u_clamped(r, 255.0f) << 16) | 0xFF000000u;
}


} // namespace sqw
