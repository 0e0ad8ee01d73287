// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17) of the Square-Wheel world-drawing hot path, used as the parity checker for the CUDA
// implementation and as the CPU baseline of bench.py.  Nothing under square_wheel_b200/ may include, link or call
// anything in this directory; only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) do.
//
// PARITY UNPINNED: the reference (Rust, no rustc in this image) cannot be built here and ships no tests, golden
// vectors or fixtures for this path.  This restatement follows the reference line by line —
// every function cites the file:line it restates — and its own known-answer vectors live in tests/golden/.
//
// Numerics: the PORTABLE branch of src/square_wheel/fast_math.rs (:370-725) is normative:
//   inv_fast = 1/x, inv_sqrt_fast = 1/sqrt(x), f32_mul_add = x*y+z (two roundings), float->int truncates.
// Build with -ffp-contract=off so the compiler never fuses what the reference keeps separate.
//
// -DORACLE_NUMERICS_SSE (liboracle_sse.so) restates the OTHER branch, fast_math.rs:6-368 — what the reference's shipping build
// executes (-Ctarget-cpu=haswell, src/build_release.sh:2): rcpss / rsqrtss approximations, cvtps2dq round-to-nearest-even
// conversions, fused multiply-add, pavgb / pavgw rounding-up averages.  It is the yardstick for north_star's tolerance
// (+-1 LSB per 8-bit channel, 1e-3 relative on HDR values) and the build timed as the CPU baseline.  rcpss / rsqrtss results are
// implementation-defined (they differ between CPU vendors), which is why this mode bounds a tolerance and is not bit-normative.
#pragma once
#ifdef ORACLE_NUMERICS_SSE
#include <immintrin.h>
#endif

#include "../include/sqw_cuda.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle
{

// ---- common/fixed_math.rs -------------------------------------------------------------------------------------
typedef int32_t Fixed16;
constexpr Fixed16 FIXED16_BASE = 16;
constexpr Fixed16 FIXED16_ONE = 1 << FIXED16_BASE;
constexpr Fixed16 FIXED16_HALF = FIXED16_ONE >> 1;

// Rust release builds wrap on integer overflow; do the arithmetic in unsigned to get the same bits without UB.
inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
inline int64_t wadd64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
inline int64_t wmul64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

inline Fixed16 int_to_fixed16(int32_t x) { return (Fixed16)((uint32_t)x << FIXED16_BASE); }      // fixed_math.rs:6
inline int32_t fixed16_round_to_int(Fixed16 x) { return wadd(x, FIXED16_HALF) >> FIXED16_BASE; } // :11
inline int32_t fixed16_floor_to_int(Fixed16 x) { return x >> FIXED16_BASE; }                     // :16
inline float fixed16_to_f32(Fixed16 x) { return (float)x / (float)FIXED16_ONE; }                 // :26
inline Fixed16 fixed16_mul(Fixed16 x, Fixed16 y) { return (int32_t)(((int64_t)x * (int64_t)y) >> FIXED16_BASE); } // :41
inline Fixed16 fixed16_div(Fixed16 x, Fixed16 y) // :56 (y != 0 guaranteed by callers)
{
	return (int32_t)((int64_t)((uint64_t)(int64_t)x << FIXED16_BASE) / (int64_t)y);
}
inline Fixed16 fixed16_mul_div(Fixed16 x, Fixed16 y, Fixed16 z) // :68
{
	return (int32_t)((int64_t)x * (int64_t)y / (int64_t)z);
}
// f32_to_fixed16_unchecked (:36): cvttss2si semantics; out-of-range gives INT_MIN like the hardware does.
inline int32_t f32_to_i32_trunc(float x)
{
	if (!(x > -2147483904.0f && x < 2147483648.0f))
		return INT32_MIN;
	return (int32_t)x;
}
inline int64_t f32_to_i64_trunc(float x) // rasterizer.rs:1552 unchecked_to_int64
{
	if (!(x > -9223373136366403584.0f && x < 9223372036854775808.0f))
		return INT64_MIN;
	return (int64_t)x;
}
inline Fixed16 f32_to_fixed16_unchecked(float x) { return f32_to_i32_trunc(x * (float)FIXED16_ONE); }

// ---- fast_math.rs portable branch ------------------------------------------------------------------------------
constexpr float MIN_POSITIVE_VALUE = 1.0f / (float)(1 << 30); // fast_math.rs:744

#ifdef ORACLE_NUMERICS_SSE
inline float inv_sqrt_fast(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set1_ps(x))); } // fast_math.rs:13
inline float inv_fast(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set1_ps(x))); }        // :19
inline float f32_mul_add(float x, float y, float z) { return _mm_cvtss_f32(_mm_fmadd_ss(_mm_set_ss(x), _mm_set_ss(y), _mm_set_ss(z))); } // :731 (target_feature = "fma")
// cvtps2dq of one lane: round to nearest even (MXCSR default), out of range -> INT_MIN
inline int32_t f32_to_i32_round(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }
#else
inline float inv_sqrt_fast(float x) { return 1.0f / std::sqrt(x); } // :376
inline float inv_fast(float x) { return 1.0f / x; }                 // :381
inline float f32_mul_add(float x, float y, float z) { return x * y + z; } // :737
#endif
// Rust f32::max / min: if one operand is NaN the other is returned.
inline float fmax_(float a, float b) { return std::fmax(a, b); }
inline float fmin_(float a, float b) { return std::fmin(a, b); }

struct Vec3
{
	float x, y, z;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float vec3_dot(Vec3 a, Vec3 b) { return f32_mul_add(a.x, b.x, f32_mul_add(a.y, b.y, a.z * b.z)); } // :780
inline Vec3 vec3_scalar_mul_add(Vec3 a, float s, Vec3 b)                                                  // :786
{
	return {f32_mul_add(a.x, s, b.x), f32_mul_add(a.y, s, b.y), f32_mul_add(a.z, s, b.z)};
}
inline float vec3_len2(Vec3 v) { return vec3_dot(v, v); } // :795

// ColorVec: 4 x f32, lanes [B, G, R, A] (fast_math.rs:443-591).
struct ColorVec
{
	float v[4];
};
inline ColorVec cv_zero() { return {{0, 0, 0, 0}}; }
inline ColorVec cv_from_color32(uint32_t c) // :452
{
	return {{(float)(c & 0xFF), (float)((c >> 8) & 0xFF), (float)((c >> 16) & 0xFF), (float)((c >> 24) & 0xFF)}};
}
inline ColorVec cv_from_color64(uint64_t c) // :462
{
	return {{(float)(c & 0xFFFF), (float)((c >> 16) & 0xFFFF), (float)((c >> 32) & 0xFFFF), (float)((c >> 48) & 0xFFFF)}};
}
inline ColorVec cv_from_f32x3(const float c[3]) { return {{c[2], c[1], c[0], 0.0f}}; }          // :505
inline ColorVec cv_from_f32x3_with_one(const float c[3]) { return {{c[2], c[1], c[0], 1.0f}}; } // :515
inline ColorVec cv_add(ColorVec a, ColorVec b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }
inline ColorVec cv_mul(ColorVec a, ColorVec b) { return {{a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]}}; }
inline ColorVec cv_scalar_mul(ColorVec a, float s) { return {{a.v[0] * s, a.v[1] * s, a.v[2] * s, a.v[3] * s}}; }
inline ColorVec cv_mul_scalar_add(ColorVec a, float s, ColorVec b) // :555
{
	return {{f32_mul_add(a.v[0], s, b.v[0]), f32_mul_add(a.v[1], s, b.v[1]), f32_mul_add(a.v[2], s, b.v[2]), f32_mul_add(a.v[3], s, b.v[3])}};
}
inline ColorVec cv_mul_add(ColorVec a, ColorVec b, ColorVec c) // :565
{
	return {{f32_mul_add(a.v[0], b.v[0], c.v[0]), f32_mul_add(a.v[1], b.v[1], c.v[1]), f32_mul_add(a.v[2], b.v[2], c.v[2]), f32_mul_add(a.v[3], b.v[3], c.v[3])}};
}
inline ColorVec cv_div(ColorVec a, ColorVec b) { return {{a.v[0] / b.v[0], a.v[1] / b.v[1], a.v[2] / b.v[2], a.v[3] / b.v[3]}}; }

// into_color32 / into_color64 (:472-498): min(max) then truncating conversion.  The reference conversion is
// `to_int_unchecked` (UB for negative input, which valid scenes never produce); the restatement and the CUDA
// kernels both clamp negatives to 0 so that the behaviour is defined and identical.
inline uint32_t f32_to_u_clamped(float x, float maxv)
{
	float m = fmin_(x, maxv);
#ifdef ORACLE_NUMERICS_SSE
	return m > 0.0f ? (uint32_t)f32_to_i32_round(m) : 0u; // _mm_min_ps + _mm_cvtps_epi32 (fast_math.rs:131-160)
#else
	return m > 0.0f ? (uint32_t)m : 0u;
#endif
}
inline uint32_t cv_into_color32(ColorVec c)
{
	return f32_to_u_clamped(c.v[0], 255.0f) | (f32_to_u_clamped(c.v[1], 255.0f) << 8) | (f32_to_u_clamped(c.v[2], 255.0f) << 16) |
		(f32_to_u_clamped(c.v[3], 255.0f) << 24);
}
inline uint64_t cv_into_color64(ColorVec c)
{
	return (uint64_t)f32_to_u_clamped(c.v[0], 65535.0f) | ((uint64_t)f32_to_u_clamped(c.v[1], 65535.0f) << 16) |
		((uint64_t)f32_to_u_clamped(c.v[2], 65535.0f) << 32) | ((uint64_t)f32_to_u_clamped(c.v[3], 65535.0f) << 48);
}

// ColorVecI: 4 x i32, lanes [B, G, R, A] (fast_math.rs:594-724).
struct ColorVecI
{
	int32_t v[4];
};
inline ColorVecI cvi_from_color32(uint32_t c)
{
	return {{(int32_t)(c & 0xFF), (int32_t)((c >> 8) & 0xFF), (int32_t)((c >> 16) & 0xFF), (int32_t)((c >> 24) & 0xFF)}};
}
inline ColorVecI cvi_from_color64(uint64_t c)
{
	return {{(int32_t)(c & 0xFFFF), (int32_t)((c >> 16) & 0xFFFF), (int32_t)((c >> 32) & 0xFFFF), (int32_t)((c >> 48) & 0xFFFF)}};
}
inline ColorVecI cvi_from_i32x3(const int32_t c[3]) { return {{c[2], c[1], c[0], 0}}; }
inline ColorVecI cvi_from_f32x3(const float c[3]) // :623  `as i32` (saturating truncation)
{
	auto cv = [](float x) -> int32_t {
#ifdef ORACLE_NUMERICS_SSE
		return f32_to_i32_round(x); // _mm_cvtps_epi32 (fast_math.rs:289)
#endif
		if (x != x)
			return 0;
		if (x >= 2147483648.0f)
			return INT32_MAX;
		if (x <= -2147483648.0f)
			return INT32_MIN;
		return (int32_t)x;
	};
	return {{cv(c[2]), cv(c[1]), cv(c[0]), 0}};
}
inline ColorVecI cvi_add(ColorVecI a, ColorVecI b) { return {{wadd(a.v[0], b.v[0]), wadd(a.v[1], b.v[1]), wadd(a.v[2], b.v[2]), wadd(a.v[3], b.v[3])}}; }
inline ColorVecI cvi_mul(ColorVecI a, ColorVecI b) { return {{wmul(a.v[0], b.v[0]), wmul(a.v[1], b.v[1]), wmul(a.v[2], b.v[2]), wmul(a.v[3], b.v[3])}}; }
inline ColorVecI cvi_mul_scalar(ColorVecI a, int32_t s) { return {{wmul(a.v[0], s), wmul(a.v[1], s), wmul(a.v[2], s), wmul(a.v[3], s)}}; }
template <int N>
inline ColorVecI cvi_shr(ColorVecI a)
{
	return {{a.v[0] >> N, a.v[1] >> N, a.v[2] >> N, a.v[3] >> N}};
}
// into_color32 (:647): (x as u32).min(255) — negative values wrap to huge and saturate.
inline uint32_t cvi_into_color32(ColorVecI c)
{
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i)
		r |= std::min<uint32_t>((uint32_t)c.v[i], 255u) << (i * 8);
	return r;
}
inline uint64_t cvi_into_color64(ColorVecI c) // :657  (x as u64).min(65535): i32 -> u64 sign-extends
{
	uint64_t r = 0;
	for (int i = 0; i < 4; ++i)
		r |= std::min<uint64_t>((uint64_t)(int64_t)c.v[i], 65535u) << (i * 16);
	return r;
}

// ---- abstract_color.rs / common/color.rs -------------------------------------------------------------------------
struct Color32
{
	typedef uint32_t Raw;
	static Raw from_vec(ColorVec c) { return cv_into_color32(c); }
	static Raw from_veci(ColorVecI c) { return cvi_into_color32(c); }
	static ColorVecI to_veci(Raw c) { return cvi_from_color32(c); }
#ifdef ORACLE_NUMERICS_SSE
	static Raw average(Raw a, Raw b) { return (Raw)_mm_cvtsi128_si32(_mm_avg_epu8(_mm_cvtsi32_si128((int)a), _mm_cvtsi32_si128((int)b))); } // fast_math.rs:43 (rounds up)
#else
	static Raw average(Raw a, Raw b) { return (((a ^ b) & 0xFEFEFEFEu) >> 1) + (a & b); } // color.rs:46
#endif
	static Raw saturated_sum(Raw a, Raw b) // fast_math.rs:386
	{
		Raw r = 0;
		for (int i = 0; i < 4; ++i)
		{
			int s = i * 8;
			r |= std::min<uint32_t>(((a >> s) & 0xFF) + ((b >> s) & 0xFF), 0xFFu) << s;
		}
		return r;
	}
	static int32_t get_alpha(Raw c) { return (int32_t)(c >> 24); }               // abstract_color.rs:25
	static bool test_alpha(Raw c) { return (c & 0xFF000000u) > 0x7F000000u; }      // :42
};
struct Color64
{
	typedef uint64_t Raw;
	static Raw from_vec(ColorVec c) { return cv_into_color64(c); }
	static Raw from_veci(ColorVecI c) { return cvi_into_color64(c); }
	static ColorVecI to_veci(Raw c) { return cvi_from_color64(c); }
#ifdef ORACLE_NUMERICS_SSE
	static Raw average(Raw a, Raw b) { return (Raw)_mm_cvtsi128_si64(_mm_avg_epu16(_mm_cvtsi64_si128((long long)a), _mm_cvtsi64_si128((long long)b))); } // fast_math.rs:54
#else
	static Raw average(Raw a, Raw b) { return (((a ^ b) & 0xFFFEFFFEFFFEFFFEull) >> 1) + (a & b); } // color.rs:163
#endif
	static Raw saturated_sum(Raw a, Raw b) // fast_math.rs:397
	{
		Raw r = 0;
		for (int i = 0; i < 4; ++i)
		{
			int s = i * 16;
			r |= std::min<uint64_t>(((a >> s) & 0xFFFF) + ((b >> s) & 0xFFFF), 0xFFFFull) << s;
		}
		return r;
	}
	static int32_t get_alpha(Raw c) { return (int32_t)(c >> 48); }                                    // abstract_color.rs:84
	static bool test_alpha(Raw c) { return (c & 0xFFFF000000000000ull) > 0x007F000000000000ull; }     // :101
};
template <class C>
inline typename C::Raw alpha_blend(typename C::Raw dst, typename C::Raw src) // abstract_color.rs:30-40 / :89-99
{
	int32_t alpha = C::get_alpha(src);
	return C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(C::to_veci(dst), 255 - alpha), cvi_mul_scalar(C::to_veci(src), alpha))));
}

enum : int
{
	BLEND_NONE = SQW_BLEND_NONE,
	BLEND_AVERAGE = SQW_BLEND_AVERAGE,
	BLEND_ADDITIVE = SQW_BLEND_ADDITIVE,
	BLEND_ALPHA_TEST = SQW_BLEND_ALPHA_TEST,
	BLEND_ALPHA_BLEND = SQW_BLEND_ALPHA_BLEND
};

// Image table entry shared by all oracle entry points (mirrors what sqw_upload_image stores on the device).
struct OracleImage
{
	const void* pixels;
	uint32_t width, height;
	int32_t kind; // SQW_IMAGE_*
	uint32_t pad;
};

// Debug-assert port: counts violated reference `debug_assert!`s instead of aborting so tests can assert zero.
extern "C" uint64_t oracle_assert_failures();
void oracle_note_assert_failure();
#define ORACLE_ASSERT(cond)             \
	do                                  \
	{                                   \
		if (!(cond))                    \
			oracle_note_assert_failure(); \
	} while (0)

} // namespace oracle
