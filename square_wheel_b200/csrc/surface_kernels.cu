// Stage 1 — surface building on sm_100a.
//
// Replaces build_surface_impl_6_static_params (src/square_wheel/surfaces.rs:433-852) and the emissive pass of
// mix_surface_with_texture (:854-971) as called from PartialRenderer::build_polygons_surfaces
// (partial_renderer.rs:1682-1819).
//
// Mapping to the GPU
//   * one CTA (256 threads) builds one BAND = up to 32 consecutive rows of one surface job holding about
//     SQW_SURF_BAND_TEXELS texels, so narrow (high-mip) and wide surfaces give every thread the same amount of work.
//     The band's texels are numbered row-major and strided over the CTA: consecutive lanes take consecutive texels of a
//     row, so global stores are contiguous 128 B (Color32) / 256 B (Color64) segments and texture loads are contiguous
//     8 B-per-lane segments that wrap at the texture border exactly like the reference's src_u counter.
//   * the reference's per-row `line_lightmap` scratch (:477-502) lives in shared memory, one line per band row; the
//     CTA first interpolates the two lightmap rows of every band row cooperatively, then every thread interpolates
//     along u for its texels.
//   * the six const-generic parameters of the reference become template parameters (colour width, lightmap kind,
//     dynamic lights, normal map, specular type); the lightmap scale stays a runtime value (it only feeds shifts).
//     Every band carries its variant and ONE grid per frame (two when some polygons have dynamic lights: those
//     variants need twice the registers) dispatches on it per CTA, so a frame's few hundred surfaces do not pay
//     one launch latency per variant.
//   * bound: HBM for the plain variants (8 B texel in, C bytes out), FP32/divide for specular + dynamic lights.
//
// Compiled with -fmad=false: the portable reference never fuses (fast_math.rs:737), and every 1/x, 1/sqrt is the
// IEEE-rounded one (fast_math.rs:376-384) so results are bit-identical to the CPU restatement.
#include "sqw_device.cuh"
#include "sqw_internal.h"
#include <algorithm>
#include <climits>

namespace sqw
{
namespace
{

constexpr int MAX_LINE_LIGHTMAP = 19; // lightmap::MAX_LIGHTMAP_SIZE + 2 (surfaces.rs:466)

constexpr float DIELECTRIC_ZERO_REFLECTIVITY = 0.04f;                                   // surfaces.rs:1152
constexpr float DIELECTRIC_AVERAGE_REFLECTIVITY = DIELECTRIC_ZERO_REFLECTIVITY * 3.0f; // :1153
constexpr float METAL_AVERAGE_SCHLICK_FACTOR = 0.5f;                                    // :1154
constexpr float NORMAL_UNPACK_SCALE = 1.0f / 127.0f;                                    // textures.rs:66

__device__ __forceinline__ float lm_mix(float a, float b, float ratio, float one_minus_ratio) // surfaces.rs:996 / :1022
{
	return f32_mul_add(a, ratio, b * one_minus_ratio);
}

__device__ __forceinline__ float get_specular_intensity(float angle_cos, float inv_roughness) // :1088 (new formula)
{
	const float x = (angle_cos - 1.0f) * inv_roughness;
	const float SQRT_7 = 2.645751311f;
	const float A = 2.0f * 3.14159265358979323846f / SQRT_7;
	const float B = 3.14159265358979323846f * SQRT_7 / 2.0f;
	return inv_fast(f32_mul_add(x * x, B, A)) * inv_roughness;
}
__device__ __forceinline__ float get_fresnel_factor_base(float angle_cos) // :1116
{
	const float omc = fmaxf(1.0f - angle_cos, 0.0f);
	const float omc2 = omc * omc;
	return omc2 * omc2 * omc;
}
__device__ __forceinline__ float get_specular_k_dielectric(float fresnel_factor_base, float roughness) // :1125
{
	const float fresnel_factor = f32_mul_add(fresnel_factor_base, 1.0f - DIELECTRIC_ZERO_REFLECTIVITY, DIELECTRIC_ZERO_REFLECTIVITY);
	return f32_mul_add(fresnel_factor, 1.0f - roughness, DIELECTRIC_AVERAGE_REFLECTIVITY * roughness);
}
__device__ __forceinline__ float get_specular_k_metal(float fresnel_factor_base, float roughness) // :1143
{
	return f32_mul_add(fresnel_factor_base, 1.0f - roughness, METAL_AVERAGE_SCHLICK_FACTOR * roughness);
}

// ---- light.rs:185-294 ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cube_side_fetch(const sqw_light& l, const float* __restrict__ sm, float vx, float vy, float vz, uint32_t side) // :231
{
	const float ONE_MINUS_EPS = 1.0f - 1.0f / 65536.0f;
	const float size_f = (float)l.shadow_map_size;
	const float depth = inv_fast(fmaxf(vz, SQW_MIN_POSITIVE_VALUE));
	const float half_depth = 0.5f * depth;
	const float u_f = fminf(fmaxf(f32_mul_add(vx, half_depth, 0.5f), 0.0f), ONE_MINUS_EPS) * size_f;
	const float v_f = fminf(fmaxf(f32_mul_add(vy, half_depth, 0.5f), 0.0f), ONE_MINUS_EPS) * size_f;
	const uint32_t u = __float2uint_rz(u_f);
	const uint32_t v = __float2uint_rz(v_f);
	const size_t side_size = (size_t)l.shadow_map_size * l.shadow_map_size;
	const float value = __ldg(sm + (size_t)l.shadow_map_offset + (size_t)side * side_size + (size_t)u + (size_t)v * l.shadow_map_size);
	return depth >= value ? 1.0f : 0.0f;
}
__device__ __forceinline__ float cube_shadow_map_fetch(const sqw_light& l, const float* __restrict__ sm, Vec3 vec) // :196
{
	const float ax = fabsf(vec.x), ay = fabsf(vec.y), az = fabsf(vec.z);
	if (ax >= ay && ax >= az)
		return vec.x >= 0.0f ? cube_side_fetch(l, sm, -vec.y, vec.z, ax, 1) : cube_side_fetch(l, sm, vec.y, vec.z, ax, 0);
	else if (ay >= ax && ay >= az)
		return vec.y >= 0.0f ? cube_side_fetch(l, sm, vec.x, vec.z, ay, 3) : cube_side_fetch(l, sm, -vec.x, vec.z, ay, 2);
	else
		return vec.z >= 0.0f ? cube_side_fetch(l, sm, -vec.x, vec.y, az, 5) : cube_side_fetch(l, sm, -vec.x, -vec.y, az, 4);
}
__device__ __forceinline__ float projector_shadow_map_fetch(const sqw_light& l, const float* __restrict__ sm, Vec3 vec) // :261
{
	const float z = vec3_dot(v3(l.basis_z[0], l.basis_z[1], l.basis_z[2]), vec);
	if (z <= 0.0f)
		return 0.0f;
	const float depth = inv_fast(z);
	const float half_depth = 0.5f * depth;
	const float size_f = (float)l.shadow_map_size;
	const float u_f = f32_mul_add(vec3_dot(v3(l.basis_x[0], l.basis_x[1], l.basis_x[2]), vec), half_depth, 0.5f) * size_f;
	const float v_f = f32_mul_add(vec3_dot(v3(l.basis_y[0], l.basis_y[1], l.basis_y[2]), vec), half_depth, 0.5f) * size_f;
	if (u_f < 0.0f || v_f < 0.0f || u_f >= size_f || v_f >= size_f)
		return 0.0f;
	const uint32_t u = __float2uint_rz(u_f);
	const uint32_t v = __float2uint_rz(v_f);
	const float value = __ldg(sm + (size_t)l.shadow_map_offset + (size_t)u + (size_t)v * l.shadow_map_size);
	return depth >= value ? 1.0f : 0.0f;
}
__device__ __forceinline__ float get_light_shadow_factor(const sqw_light& l, const float* __restrict__ sm, Vec3 vec) // :185
{
	if (l.shadow_type == SQW_SHADOW_CUBE)
		return cube_shadow_map_fetch(l, sm, vec);
	if (l.shadow_type == SQW_SHADOW_PROJECTOR)
		return projector_shadow_map_fetch(l, sm, vec);
	return 1.0f;
}

// Line-lightmap element in shared memory.  Simple lightmaps use the first 3 floats only.
struct LmEl
{
	float ambient[3];
	float dir[3];
	float deviation;
	float color[3];
};

// mix_surface_with_texture_impl for one texel (surfaces.rs:926-962) with the two blend modes the emissive pass uses.
template <class C>
__device__ __forceinline__ typename C::Raw emissive_mix(uint32_t texel, ColorVecI light_vec, bool additive, typename C::Raw dst)
{
	const ColorVecI m = cvi_shr<8>(cvi_mul(cvi_from_color32(texel), light_vec));
	if (additive)
		return C::from_veci(cvi_add(m, C::to_veci(dst)));
	return C::from_veci(m);
}

// rem_euclid (surfaces.rs:507-510) with the power-of-two shortcut (two's complement & mask == rem_euclid).
// The general case is an out-of-line call on purpose: inlined, the compiler if-converts the branch and every texel pays the
// ~30 instructions of the integer division even when all textures are powers of two (a tenth of the kernel's instructions).
static __device__ __noinline__ uint32_t wrap_coord_general(int32_t x, int32_t size)
{
	const int32_t m = x % size;
	return (uint32_t)(m < 0 ? m + size : m);
}
__device__ __forceinline__ uint32_t wrap_coord(int32_t x, int32_t size, uint32_t pow2_mask)
{
	if (pow2_mask)
		return (uint32_t)x & pow2_mask;
	return wrap_coord_general(x, size);
}
__device__ __forceinline__ uint32_t pow2_mask_of(uint32_t size) { return (size & (size - 1)) == 0 ? size - 1 : 0; }

typedef LmEl LineLightmap[MAX_LINE_LIGHTMAP + 1];

// One band of one job: build_surface_impl_6_static_params (+ the emissive pass) for rows [band.row0, band.row0 + band.num_rows).
// G = threads that share the band: the whole CTA (SQW_SURF_THREADS, bands of about SQW_SURF_BAND_TEXELS texels) or one warp (32:
// small surfaces, eight independent bands per CTA and no CTA barrier).  The line lightmap is `lm_pitch` elements per band row.
template <class C, bool DIRECTIONAL, bool USE_DYNAMIC_LIGHTS, bool USE_NORMAL_MAP, int SPECULAR_TYPE, bool EMISSIVE_ONLY, int G>
__device__ __forceinline__ void surface_band(
	const sqw_surface_job& j, const SurfaceBand band, LmEl* const line_lightmap, const ImageDev* __restrict__ images, const float* __restrict__ lightmaps,
	const sqw_directional_lightmap_element* __restrict__ directional_lightmaps, const sqw_light* __restrict__ lights, const float* __restrict__ shadow_maps,
	float cam_x, float cam_y, float cam_z, typename C::Raw* __restrict__ surfaces)
{
	typedef typename C::Raw Raw;
	const uint32_t tid = G == 32 ? (threadIdx.x & 31u) : threadIdx.x;
	const uint32_t size_u = j.surface_size[0];
	const uint32_t num_rows = band.num_rows;
	const uint32_t total = num_rows * size_u;
	Raw* const out = surfaces + j.surface_pixels_offset;
	// (row, u) of this thread's first texel and its advance per stride of SQW_SURF_THREADS texels
	const uint32_t step_rows = (uint32_t)G / size_u, step_u = (uint32_t)G % size_u;
	uint32_t r = tid / size_u, dst_u = tid % size_u;

	// Emissive layer (partial_renderer.rs:1801-1819): Additive over the diffuse layer, None when there is none.
	const bool emissive = (j.flags & SQW_SURF_EMISSIVE) != 0;
	ColorVecI emissive_light_vec = cvi(0, 0, 0, 0);
	const uint32_t* emissive_pixels = nullptr;
	uint32_t em_w = 1, em_h = 1, em_w_mask = 0, em_h_mask = 0;
	int32_t em_u0 = 0;
	if (emissive)
	{
		const ImageDev eimg = images[j.emissive_image];
		emissive_pixels = static_cast<const uint32_t*>(eimg.pixels);
		em_w = eimg.width;
		em_h = eimg.height;
		em_w_mask = pow2_mask_of(em_w);
		em_h_mask = pow2_mask_of(em_h);
		em_u0 = (int32_t)wrap_coord(j.tc_start[0], (int32_t)em_w, em_w_mask);
		// surfaces.rs:911-914  light * 256 -> ColorVecI (`as i32`), lanes [B, G, R, 0]
		emissive_light_vec = cvi(f32_as_i32(j.emissive_light[2] * 256.0f), f32_as_i32(j.emissive_light[1] * 256.0f), f32_as_i32(j.emissive_light[0] * 256.0f), 0);
	}

	if (EMISSIVE_ONLY)
	{
		for (uint32_t i = tid; i < total; i += G)
		{
			const uint32_t dst_v = band.row0 + r;
			const uint32_t ev = wrap_coord(j.tc_start[1] + (int32_t)dst_v, (int32_t)em_h, em_h_mask);
			const uint32_t eu = wrap_coord(em_u0 + (int32_t)dst_u, (int32_t)em_w, em_w_mask);
			out[(size_t)dst_v * size_u + dst_u] = emissive_mix<C>(__ldg(emissive_pixels + (size_t)ev * em_w + eu), emissive_light_vec, false, 0);
			r += step_rows;
			dst_u += step_u;
			if (dst_u >= size_u)
			{
				dst_u -= size_u;
				++r;
			}
		}
		return;
	}

	const ImageDev timg = images[j.texture_image];
	const uint2* const texture = static_cast<const uint2*>(timg.pixels);
	const int32_t tex_w = (int32_t)timg.width, tex_h = (int32_t)timg.height;
	const uint32_t tex_w_mask = pow2_mask_of(timg.width), tex_h_mask = pow2_mask_of(timg.height);

	const Vec3 u_vec = v3(j.basis_u[0], j.basis_u[1], j.basis_u[2]);
	const Vec3 v_vec = v3(j.basis_v[0], j.basis_v[1], j.basis_v[2]);
	const Vec3 basis_normal = v3(j.basis_normal[0], j.basis_normal[1], j.basis_normal[2]);
	const Vec3 cam_pos = v3(cam_x, cam_y, cam_z);
	// surfaces.rs:456  start + u*(tc_min.x + 0.5) + v*(tc_min.y + 0.5)   (cgmath Vec3 ops: plain mul and add)
	const Vec3 start_pos =
		v3(j.basis_start[0], j.basis_start[1], j.basis_start[2]) + u_vec * ((float)j.tc_start[0] + 0.5f) + v_vec * ((float)j.tc_start[1] + 0.5f);
	const Vec3 u_vec_normalized = u_vec * inv_sqrt_fast(fmaxf(vec3_len2(u_vec), SQW_MIN_POSITIVE_VALUE)); // :461
	const Vec3 v_vec_normalized = v_vec * inv_sqrt_fast(fmaxf(vec3_len2(v_vec), SQW_MIN_POSITIVE_VALUE)); // :462

	const bool has_lightmap = j.lightmap_offset != 0; // lightmap_data.is_empty() (:140)
	const uint32_t lm_scale_log2 = j.lightmap_scale_log2;
	const float lightmap_scale_f = (float)(1u << (lm_scale_log2 & 31));
	const float inv_lightmap_scale_f = 1.0f / lightmap_scale_f;
	const float k_shift = 0.5f * inv_lightmap_scale_f;
	const uint32_t lightmap_fetch_mask = (1u << (lm_scale_log2 & 31)) - 1;
	const uint32_t lm_w = j.lightmap_size[0];
	const uint32_t lm_cols = min(lm_w, (uint32_t)MAX_LINE_LIGHTMAP);
	// row pitch of the line lightmap: fixed for CTA bands, the band's own column count for warp bands (their rows share a small pool)
	const uint32_t lm_pitch = G == 32 ? lm_cols : (uint32_t)(MAX_LINE_LIGHTMAP + 1);

	const int32_t src_u0 = (int32_t)wrap_coord(j.tc_start[0], tex_w, tex_w_mask); // rem_euclid (:510)

	if (has_lightmap)
	{
		// :477-502 interpolate two lightmap rows into the line lightmap of every row of the band.
		const uint32_t entries = num_rows * lm_cols;
		// e / lm_cols by a 16-bit reciprocal: exact for e < 65536 / lm_cols (e < 32 rows x 19 columns here), instead of the
		// generic 32-bit division sequence in every iteration
		const uint32_t lm_cols_rcp = 65536u / lm_cols + 1u;
		for (uint32_t e = tid; e < entries; e += G)
		{
			const uint32_t rr = (e * lm_cols_rcp) >> 16, x = e - rr * lm_cols;
			const uint32_t lightmap_base_v = band.row0 + rr + j.lightmap_tc_shift[1];
			const uint32_t lightmap_v = lightmap_base_v >> lm_scale_log2;
			const float k = f32_mul_add((float)(lightmap_base_v & lightmap_fetch_mask), inv_lightmap_scale_f, k_shift);
			const float omk = 1.0f - k;
			const uint32_t base = lightmap_v * lm_w;
			float* d = reinterpret_cast<float*>(line_lightmap + rr * lm_pitch + x);
			if (DIRECTIONAL)
			{
				const float* l0 = reinterpret_cast<const float*>(directional_lightmaps + (size_t)j.lightmap_offset + base + x);
				const float* l1 = l0 + (size_t)lm_w * 10;
#pragma unroll
				for (int c = 0; c < 10; ++c)
					d[c] = lm_mix(__ldg(l1 + c), __ldg(l0 + c), k, omk);
			}
			else
			{
				const float* l0 = lightmaps + ((size_t)j.lightmap_offset + base + x) * 3;
				const float* l1 = l0 + (size_t)lm_w * 3;
#pragma unroll
				for (int c = 0; c < 3; ++c)
					d[c] = lm_mix(__ldg(l1 + c), __ldg(l0 + c), k, omk);
			}
		}
	}
	if (G == 32)
		__syncwarp();
	else
		__syncthreads();

	for (uint32_t i = tid; i < total; i += G)
	{
		const uint32_t dst_v = band.row0 + r;
		const uint32_t src_v = wrap_coord(j.tc_start[1] + (int32_t)dst_v, tex_h, tex_h_mask); // :507
		const Vec3 start_pos_v = vec3_scalar_mul_add(v_vec, (float)dst_v, start_pos);          // :512
		const Vec3 pos = vec3_scalar_mul_add(u_vec, (float)dst_u, start_pos_v);                // :515
		const uint32_t su = wrap_coord(src_u0 + (int32_t)dst_u, tex_w, tex_w_mask); // the reference's wrapping src_u counter (:843-847)
		const uint32_t* em_line = nullptr;
		if (emissive)
			em_line = emissive_pixels + (size_t)wrap_coord(j.tc_start[1] + (int32_t)dst_v, (int32_t)em_h, em_h_mask) * em_w;
		const uint2* const src_line = texture + (size_t)src_v * tex_w;
		const uint2 texel_value = __ldg(src_line + su);
		const uint32_t diffuse = texel_value.x;
		const int32_t packed = (int32_t)texel_value.y;

		Vec3 texel_normal = v3(0.0f, 0.0f, 1.0f);
		float texel_roughness = 0.0f;
		if (USE_NORMAL_MAP)
		{
			// PackedNormalRoughness::unpack (textures.rs:42-58)
			texel_normal.x = (float)((int32_t)((uint32_t)packed << 24) >> 24) * NORMAL_UNPACK_SCALE;
			texel_normal.y = (float)((int32_t)((uint32_t)packed << 16) >> 24) * NORMAL_UNPACK_SCALE;
			texel_normal.z = (float)((int32_t)((uint32_t)packed << 8) >> 24) * NORMAL_UNPACK_SCALE;
			texel_roughness = (float)(packed >> 24) * NORMAL_UNPACK_SCALE;
		}
		else if (SPECULAR_TYPE != 0)
			texel_roughness = (float)(packed >> 24) * NORMAL_UNPACK_SCALE; // unpack_roughness (textures.rs:60)

		ColorVec total_light_albedo_modulated = cv(0.0f, 0.0f, 0.0f, 0.0f);
		ColorVec total_light_direct = cv(0.0f, 0.0f, 0.0f, 0.0f);

		if (has_lightmap)
		{
			const uint32_t lightmap_base_u = dst_u + j.lightmap_tc_shift[0];
			const uint32_t lightmap_u = lightmap_base_u >> lm_scale_log2;
			const float k = f32_mul_add((float)(lightmap_base_u & lightmap_fetch_mask), inv_lightmap_scale_f, k_shift);
			const float omk = 1.0f - k;
			const LmEl& l0 = line_lightmap[r * lm_pitch + lightmap_u];
			const LmEl& l1 = line_lightmap[r * lm_pitch + lightmap_u + 1];
			float constant_c[3];
#pragma unroll
			for (int c = 0; c < 3; ++c)
				constant_c[c] = lm_mix(l1.ambient[c], l0.ambient[c], k, omk);
			Vec3 vs = v3(0.0f, 0.0f, 0.0f);
			float deviation = 0.0f;
			float dir_color_c[3] = {0.0f, 0.0f, 0.0f};
			if (DIRECTIONAL)
			{
				vs = v3(lm_mix(l1.dir[0], l0.dir[0], k, omk), lm_mix(l1.dir[1], l0.dir[1], k, omk), lm_mix(l1.dir[2], l0.dir[2], k, omk));
				deviation = lm_mix(l1.deviation, l0.deviation, k, omk);
#pragma unroll
				for (int c = 0; c < 3; ++c)
					dir_color_c[c] = lm_mix(l1.color[c], l0.color[c], k, omk);
			}
			const ColorVec dir_color = cv_from_f32x3(dir_color_c);

			if (SPECULAR_TYPE == 0)
			{
				total_light_albedo_modulated = cv(constant_c[2], constant_c[1], constant_c[0], 1.0f); // from_color_f32x3_with_one (:553)
				if (DIRECTIONAL)
				{
					const float dot = USE_NORMAL_MAP ? fmaxf(vec3_dot(vs, texel_normal), 0.0f) : vs.z;
					total_light_albedo_modulated = cv_mul_scalar_add(dir_color, dot, total_light_albedo_modulated);
				}
			}
			else
			{
				const Vec3 normal = texel_normal; // unit_z when there is no normal map (:575-582)
				const Vec3 vec_to_camera = cam_pos - pos;
				const Vec3 vtc_ts = v3(vec3_dot(vec_to_camera, u_vec_normalized), vec3_dot(vec_to_camera, v_vec_normalized), vec3_dot(vec_to_camera, basis_normal));
				const float vec_to_camera_normal_dot = vec3_dot(vtc_ts, normal);
				const Vec3 vec_to_camera_reflected = normal * (2.0f * vec_to_camera_normal_dot) - vtc_ts;
				const float vec_to_camera_len2 = fmaxf(vec3_len2(vec_to_camera_reflected), SQW_MIN_POSITIVE_VALUE);
				const float angle_cos = fmaxf(vec_to_camera_normal_dot * inv_sqrt_fast(vec_to_camera_len2), 0.0f);
				const float fresnel_factor_base = get_fresnel_factor_base(angle_cos);
				const float specular_k =
					SPECULAR_TYPE == 1 ? get_specular_k_dielectric(fresnel_factor_base, texel_roughness) : get_specular_k_metal(fresnel_factor_base, texel_roughness);
				const float one_minus_specular_k = 1.0f - specular_k;

				const ColorVec constant_component = cv_from_f32x3(constant_c);
				total_light_albedo_modulated = cv_scalar_mul(constant_component, one_minus_specular_k);
				total_light_direct = cv_scalar_mul(constant_component, specular_k);
				total_light_albedo_modulated.a = 1.0f; // :621

				if (DIRECTIONAL)
				{
					const float direction_vec_len2 = vec3_len2(vs);
					const float direction_vec_len = direction_vec_len2 * inv_sqrt_fast(direction_vec_len2);
					const float refl_light_angle_cos = vec3_dot(vec_to_camera_reflected, vs) * inv_sqrt_fast(vec_to_camera_len2 * direction_vec_len2);
					const float inv_roughness_corrected = fmaxf(inv_fast(texel_roughness + deviation), 0.75f); // :633
					const float specular_intensity = get_specular_intensity(refl_light_angle_cos, inv_roughness_corrected);
					if (SPECULAR_TYPE == 1)
					{
						const float diffuse_intensity = fmaxf(vec3_dot(vs, texel_normal), 0.0f);
						total_light_albedo_modulated = cv_mul_scalar_add(dir_color, diffuse_intensity * one_minus_specular_k, total_light_albedo_modulated);
						total_light_direct = cv_mul_scalar_add(dir_color, specular_intensity * specular_k * direction_vec_len, total_light_direct);
					}
					else
					{
						const float scale = specular_intensity * direction_vec_len;
						total_light_albedo_modulated = cv_mul_scalar_add(dir_color, one_minus_specular_k * scale, total_light_albedo_modulated);
						total_light_direct = cv_mul_scalar_add(dir_color, specular_k * scale, total_light_direct);
					}
				}
			}
		}

		if (USE_DYNAMIC_LIGHTS)
		{
			Vec3 normal;
			if (USE_NORMAL_MAP)
			{
				normal = v3( // :686-702
					f32_mul_add(texel_normal.x, u_vec_normalized.x, f32_mul_add(texel_normal.y, v_vec_normalized.x, texel_normal.z * basis_normal.x)),
					f32_mul_add(texel_normal.x, u_vec_normalized.y, f32_mul_add(texel_normal.y, v_vec_normalized.y, texel_normal.z * basis_normal.y)),
					f32_mul_add(texel_normal.x, u_vec_normalized.z, f32_mul_add(texel_normal.y, v_vec_normalized.z, texel_normal.z * basis_normal.z)));
			}
			else
				normal = basis_normal;

			Vec3 vec_to_camera_reflected = v3(0.0f, 0.0f, 0.0f);
			float vec_to_camera_len2 = SQW_MIN_POSITIVE_VALUE;
			float specular_k = 0.0f;
			float inv_roughness = 0.0f;
			if (SPECULAR_TYPE != 0)
			{
				const Vec3 vec_to_camera = cam_pos - pos;
				const float vec_to_camera_normal_dot = vec3_dot(vec_to_camera, normal);
				vec_to_camera_reflected = normal * (2.0f * vec_to_camera_normal_dot) - vec_to_camera;
				vec_to_camera_len2 = fmaxf(vec3_len2(vec_to_camera_reflected), SQW_MIN_POSITIVE_VALUE);
				const float angle_cos = fmaxf(vec_to_camera_normal_dot * inv_sqrt_fast(vec_to_camera_len2), 0.0f);
				const float fresnel_factor_base = get_fresnel_factor_base(angle_cos);
				specular_k =
					SPECULAR_TYPE == 1 ? get_specular_k_dielectric(fresnel_factor_base, texel_roughness) : get_specular_k_metal(fresnel_factor_base, texel_roughness);
				inv_roughness = inv_fast(texel_roughness);
			}

			const uint32_t num_lights = min(j.num_lights, (uint32_t)SQW_MAX_POLYGON_LIGHTS);
			for (uint32_t li = 0; li < num_lights; ++li)
			{
				const sqw_light& light = lights[j.lights[li]];
				const Vec3 vec_to_light = v3(light.position[0], light.position[1], light.position[2]) - pos;
				const float shadow_factor = get_light_shadow_factor(light, shadow_maps, vec_to_light);
				const float vec_to_light_len2 = fmaxf(vec3_len2(vec_to_light), SQW_MIN_POSITIVE_VALUE);
				const float distance_factor = fmaxf(inv_fast(vec_to_light_len2) - light.inv_square_radius, 0.0f);
				const float shadow_distance_factor = shadow_factor * distance_factor;
				const float diffuse_intensity = SPECULAR_TYPE == 2 ? 0.0f : fmaxf(vec3_dot(normal, vec_to_light) * inv_sqrt_fast(vec_to_light_len2), 0.0f);
				float specular_intensity = 0.0f;
				if (SPECULAR_TYPE != 0)
				{
					const float c = vec3_dot(vec_to_camera_reflected, vec_to_light) * inv_sqrt_fast(vec_to_camera_len2 * vec_to_light_len2);
					specular_intensity = get_specular_intensity(c, inv_roughness);
				}
				const ColorVec light_color = cv_from_f32x3(light.color);
				if (SPECULAR_TYPE == 0)
					total_light_albedo_modulated = cv_mul_scalar_add(light_color, diffuse_intensity * shadow_distance_factor, total_light_albedo_modulated);
				else if (SPECULAR_TYPE == 1)
				{
					total_light_albedo_modulated =
						cv_mul_scalar_add(light_color, diffuse_intensity * (1.0f - specular_k) * shadow_distance_factor, total_light_albedo_modulated);
					total_light_direct = cv_mul_scalar_add(light_color, specular_intensity * specular_k * shadow_distance_factor, total_light_direct);
				}
				else
				{
					const float sisdf = specular_intensity * shadow_distance_factor;
					total_light_albedo_modulated = cv_mul_scalar_add(light_color, (1.0f - specular_k) * sisdf, total_light_albedo_modulated);
					total_light_direct = cv_mul_scalar_add(light_color, specular_k * sisdf, total_light_direct);
				}
			}
		}

		ColorVec result_color = cv_mul(cv_from_color32(diffuse), total_light_albedo_modulated); // :833
		if (SPECULAR_TYPE != 0)
			result_color = cv_mul_scalar_add(total_light_direct, 255.0f, result_color);
		Raw result = C::from_vec(result_color);

		if (emissive)
		{
			const uint32_t eu = wrap_coord(em_u0 + (int32_t)dst_u, (int32_t)em_w, em_w_mask);
			result = emissive_mix<C>(__ldg(em_line + eu), emissive_light_vec, true, result);
		}
		out[(size_t)dst_v * size_u + dst_u] = result;

		r += step_rows;
		dst_u += step_u;
		if (dst_u >= size_u)
		{
			dst_u -= size_u;
			++r;
		}
	}
}

// Variant byte of a band: bit0 directional lightmap, bit1 normal map, bits2-3 specular type, bit4 emissive-only.
#define SQW_BAND_CASE(DIR, NM, SPEC)                                                                                                                       \
	case ((DIR) | ((NM) << 1) | ((SPEC) << 2)):                                                                                                            \
		surface_band<C, (DIR) != 0, DYN, (NM) != 0, SPEC, false, G>(                                                                                       \
			job, band, line_lightmap, images, lightmaps, directional_lightmaps, lights, shadow_maps, cam_x, cam_y, cam_z, surfaces);                       \
		break;
template <class C, bool DYN, int G>
__device__ __forceinline__ void surface_band_dispatch(
	const sqw_surface_job& job, const SurfaceBand band, LmEl* const line_lightmap, const ImageDev* __restrict__ images, const float* __restrict__ lightmaps,
	const sqw_directional_lightmap_element* __restrict__ directional_lightmaps, const sqw_light* __restrict__ lights, const float* __restrict__ shadow_maps,
	float cam_x, float cam_y, float cam_z, typename C::Raw* __restrict__ surfaces)
{
	switch (band.variant)
	{
		SQW_BAND_CASE(0, 0, 0)
		SQW_BAND_CASE(0, 0, 1)
		SQW_BAND_CASE(0, 0, 2)
		SQW_BAND_CASE(0, 1, 0)
		SQW_BAND_CASE(0, 1, 1)
		SQW_BAND_CASE(0, 1, 2)
		SQW_BAND_CASE(1, 0, 0)
		SQW_BAND_CASE(1, 0, 1)
		SQW_BAND_CASE(1, 0, 2)
		SQW_BAND_CASE(1, 1, 0)
		SQW_BAND_CASE(1, 1, 1)
		SQW_BAND_CASE(1, 1, 2)
	default:
		if (!DYN)
			surface_band<C, false, false, false, 0, true, G>(
				job, band, line_lightmap, images, lightmaps, directional_lightmaps, lights, shadow_maps, cam_x, cam_y, cam_z, surfaces);
		break;
	}
}
#undef SQW_BAND_CASE

// One grid per frame and dynamic-light class.  CTAs [0, num_cta_bands) build one CTA band each (up to SQW_SURF_MAX_BAND_ROWS rows,
// about SQW_SURF_BAND_TEXELS texels, 256 threads); the CTAs after them build eight WARP bands each — small surfaces (far or finely
// tessellated geometry: the stress map averages 76 texels per polygon), one warp per band and no CTA barrier: a CTA band would
// leave most of its threads idle and still pay the prologue.  A warp's line lightmap is a pool of SQW_SURF_WARP_LM_ELEMENTS
// elements shared by the band's rows (the host sizes the band so that rows x lightmap columns fits).  Both kinds take the same
// time per band (eight texels per thread), so they share one grid instead of two launches with a tail each.
template <class C, bool DYN>
__global__ void __launch_bounds__(SQW_SURF_THREADS, DYN ? 3 : 4) surface_build_kernel(
	const sqw_surface_job* __restrict__ jobs, const SurfaceBand* __restrict__ bands, uint32_t num_cta_bands, uint32_t num_bands, const ImageDev* __restrict__ images,
	const float* __restrict__ lightmaps, const sqw_directional_lightmap_element* __restrict__ directional_lightmaps,
	const sqw_light* __restrict__ lights, const float* __restrict__ shadow_maps, float cam_x, float cam_y, float cam_z,
	typename C::Raw* __restrict__ surfaces)
{
	constexpr int WARPS = SQW_SURF_THREADS / 32;
	static_assert(WARPS * SQW_SURF_WARP_LM_ELEMENTS <= SQW_SURF_MAX_BAND_ROWS * (MAX_LINE_LIGHTMAP + 1), "the warp pools fit the CTA band's line lightmap");
	__shared__ LmEl s_line_lightmap[SQW_SURF_MAX_BAND_ROWS * (MAX_LINE_LIGHTMAP + 1)];
	__shared__ sqw_surface_job s_job[WARPS];

	if (blockIdx.x < num_cta_bands) // (CTA-uniform)
	{
		const SurfaceBand band = bands[blockIdx.x];
		{
			// Cooperative 136-byte job load into shared memory (34 words).
			const uint32_t* src = reinterpret_cast<const uint32_t*>(jobs + band.job);
			uint32_t* dst = reinterpret_cast<uint32_t*>(&s_job[0]);
			if (threadIdx.x < sizeof(sqw_surface_job) / 4)
				dst[threadIdx.x] = src[threadIdx.x];
		}
		__syncthreads();
		surface_band_dispatch<C, DYN, SQW_SURF_THREADS>(
			s_job[0], band, s_line_lightmap, images, lightmaps, directional_lightmaps, lights, shadow_maps, cam_x, cam_y, cam_z, surfaces);
		return;
	}
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t band_index = num_cta_bands + (blockIdx.x - num_cta_bands) * WARPS + warp;
	if (band_index >= num_bands)
		return;
	const SurfaceBand band = bands[band_index];
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(jobs + band.job);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&s_job[warp]);
		for (uint32_t i = lane; i < sizeof(sqw_surface_job) / 4; i += 32)
			dst[i] = src[i];
	}
	__syncwarp();
	surface_band_dispatch<C, DYN, 32>(
		s_job[warp], band, s_line_lightmap + warp * SQW_SURF_WARP_LM_ELEMENTS, images, lightmaps, directional_lightmaps, lights, shadow_maps, cam_x, cam_y, cam_z, surfaces);
}

// mix_surface_with_texture_impl (surfaces.rs:903-971) with any of the five blending modes: the overlay of a portal /
// mirror glass texture on a rendered portal texture (partial_renderer.rs:2261-2271).  One thread per texel, consecutive
// threads on consecutive texels of a row (coalesced read-modify-write of the surface, wrapped reads of the overlay).
template <class C>
__global__ void __launch_bounds__(256) mix_texture_kernel(
	typename C::Raw* __restrict__ surface, uint32_t size_u, uint32_t size_v, int32_t tc_min_u, int32_t tc_min_v, const uint32_t* __restrict__ texture, uint32_t tex_w,
	uint32_t tex_h, int blending_mode, int32_t light_b, int32_t light_g, int32_t light_r)
{
	const uint32_t tex_w_mask = pow2_mask_of(tex_w), tex_h_mask = pow2_mask_of(tex_h);
	const ColorVecI light_vec = cvi(light_b, light_g, light_r, 0);
	const uint64_t total = (uint64_t)size_u * size_v;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t dst_v = (uint32_t)(i / size_u), dst_u = (uint32_t)(i - (uint64_t)dst_v * size_u);
		const uint32_t src_v = wrap_coord(tc_min_v + (int32_t)dst_v, (int32_t)tex_h, tex_h_mask);
		const uint32_t src_u = wrap_coord(tc_min_u + (int32_t)dst_u, (int32_t)tex_w, tex_w_mask);
		const uint32_t texel_value = __ldg(texture + (size_t)src_v * tex_w + src_u);
		const ColorVecI m = cvi_shr<8>(cvi_mul(cvi_from_color32(texel_value), light_vec));
		typename C::Raw& dst = surface[i];
		switch (blending_mode)
		{
		case SQW_BLEND_NONE: dst = C::from_veci(m); break;
		case SQW_BLEND_AVERAGE: dst = C::from_veci(cvi_shr<1>(cvi_add(m, C::to_veci(dst)))); break;
		case SQW_BLEND_ADDITIVE: dst = C::from_veci(cvi_add(m, C::to_veci(dst))); break;
		case SQW_BLEND_ALPHA_TEST:
			if (Color32::test_alpha(texel_value))
				dst = C::from_veci(m);
			break;
		default:
		{
			const int32_t alpha = Color32::get_alpha(texel_value);
			dst = C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(m, alpha), cvi_mul_scalar(C::to_veci(dst), 255 - alpha))));
			break;
		}
		}
	}
}

template <class C, bool DYN>
cudaError_t launch_one(const SurfaceLaunch& l, cudaStream_t stream)
{
	const uint32_t warp_ctas = (l.num_bands - l.num_cta_bands + SQW_SURF_THREADS / 32 - 1) / (SQW_SURF_THREADS / 32);
	surface_build_kernel<C, DYN><<<l.num_cta_bands + warp_ctas, SQW_SURF_THREADS, 0, stream>>>(
		l.jobs, l.bands, l.num_cta_bands, l.num_bands, l.images, l.lightmaps, l.directional_lightmaps, l.lights, l.shadow_maps, l.cam_pos[0], l.cam_pos[1], l.cam_pos[2],
		static_cast<typename C::Raw*>(l.surfaces));
	return cudaGetLastError();
}

} // namespace

// This file is compiled twice: as is (portable numerics, bit-exact) and through surface_kernels_fast.cu with SQW_NUMERICS_FAST
// (sqw_device.cuh), which exports the second entry point.
#if defined(SQW_NUMERICS_FAST) && SQW_NUMERICS_FAST
cudaError_t launch_surface_build_fast(bool color64, bool dynamic_lights, const SurfaceLaunch& l, cudaStream_t stream)
#else
cudaError_t launch_surface_build(bool color64, bool dynamic_lights, const SurfaceLaunch& l, cudaStream_t stream)
#endif
{
	if (l.num_bands == 0)
		return cudaSuccess;
	if (color64)
		return dynamic_lights ? launch_one<Color64, true>(l, stream) : launch_one<Color64, false>(l, stream);
	return dynamic_lights ? launch_one<Color32, true>(l, stream) : launch_one<Color32, false>(l, stream);
}

#if !(defined(SQW_NUMERICS_FAST) && SQW_NUMERICS_FAST)
cudaError_t launch_mix_texture(
	bool color64, void* surface, uint32_t size_u, uint32_t size_v, const int32_t tc_min[2], const void* texture, uint32_t tex_w, uint32_t tex_h, int blending_mode,
	const float light[3], cudaStream_t stream)
{
	const uint64_t total = (uint64_t)size_u * size_v;
	if (total == 0)
		return cudaSuccess;
	// surfaces.rs:911-914  light * 256 -> ColorVecI (`as i32`: saturating, NaN -> 0), lanes [B, G, R, 0]
	auto as_i32 = [](float x) -> int32_t {
		if (!(x == x))
			return 0;
		if (x >= 2147483648.0f)
			return INT32_MAX;
		if (x <= -2147483648.0f)
			return INT32_MIN;
		return (int32_t)x;
	};
	const int32_t lb = as_i32(light[2] * 256.0f), lg = as_i32(light[1] * 256.0f), lr = as_i32(light[0] * 256.0f);
	const uint32_t blocks = (uint32_t)std::min<uint64_t>((total + 255) / 256, 148ull * 8);
	if (color64)
		mix_texture_kernel<Color64><<<blocks, 256, 0, stream>>>(
			static_cast<uint64_t*>(surface), size_u, size_v, tc_min[0], tc_min[1], static_cast<const uint32_t*>(texture), tex_w, tex_h, blending_mode, lb, lg, lr);
	else
		mix_texture_kernel<Color32><<<blocks, 256, 0, stream>>>(
			static_cast<uint32_t*>(surface), size_u, size_v, tc_min[0], tc_min[1], static_cast<const uint32_t*>(texture), tex_w, tex_h, blending_mode, lb, lg, lr);
	return cudaGetLastError();
}
#endif

} // namespace sqw
