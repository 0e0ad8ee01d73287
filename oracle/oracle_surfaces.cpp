// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).  PARITY UNPINNED.
//
// Restatement of src/square_wheel/surfaces.rs (stage 1: surface building) and the light fetches of
// src/square_wheel/light.rs, portable numerics.
#include "oracle_common.h"
#include <atomic>
#include "oracle_pool.h"
#include <vector>

namespace oracle
{

static std::atomic<uint64_t> g_assert_failures{0};
void oracle_note_assert_failure() { g_assert_failures.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t oracle_assert_failures() { return g_assert_failures.load(); }
extern "C" void oracle_reset_assert_failures() { g_assert_failures.store(0); }

namespace
{

constexpr uint32_t SPECULAR_TYPE_NONE = 0;       // surfaces.rs:425
constexpr uint32_t SPECULAR_TYPE_DIELECTRIC = 1; // :426
constexpr uint32_t SPECULAR_TYPE_METAL = 2;      // :427
constexpr uint32_t NO_LIGHTMAP_SCALE = SQW_NO_LIGHTMAP_SCALE; // :429
constexpr uint32_t MAX_LIGHTMAP_SIZE = 17; // common/lightmap.rs:9

constexpr float DIELECTRIC_ZERO_REFLECTIVITY = 0.04f;                                // :1152
constexpr float DIELECTRIC_AVERAGE_REFLECTIVITY = DIELECTRIC_ZERO_REFLECTIVITY * 3.0f; // :1153
constexpr float METAL_AVERAGE_SCHLICK_FACTOR = 0.5f;                                 // :1154

struct TextureElement // textures.rs:20-27
{
	uint32_t diffuse;
	int32_t packed_normal_roughness;
};

// PackedNormalRoughness::unpack (textures.rs:42-58) via upack_bytes_into_f32x4 (fast_math.rs:430).
constexpr float NORMAL_UNPACK_SCALE = 1.0f / 127.0f;
inline void unpack_normal_roughness(int32_t b, Vec3& n, float& roughness)
{
	n.x = (float)((int32_t)((uint32_t)b << 24) >> 24) * NORMAL_UNPACK_SCALE;
	n.y = (float)((int32_t)((uint32_t)b << 16) >> 24) * NORMAL_UNPACK_SCALE;
	n.z = (float)((int32_t)((uint32_t)b << 8) >> 24) * NORMAL_UNPACK_SCALE;
	roughness = (float)(b >> 24) * NORMAL_UNPACK_SCALE;
}
inline float unpack_roughness(int32_t b) { return (float)(b >> 24) * NORMAL_UNPACK_SCALE; } // textures.rs:60

// ---- light.rs:185-294 -------------------------------------------------------------------------------------------
inline float cube_shadow_map_side_fetch(const sqw_light& l, const float* shadow_maps, Vec3 vec, uint32_t side) // :231
{
	const float ONE_MINUS_EPS = 1.0f - 1.0f / 65536.0f;
	float cubemap_size_f = (float)l.shadow_map_size;
	float depth = inv_fast(fmax_(vec.z, MIN_POSITIVE_VALUE));
	float half_depth = 0.5f * depth;
	float u_f = fmin_(fmax_(f32_mul_add(vec.x, half_depth, 0.5f), 0.0f), ONE_MINUS_EPS) * cubemap_size_f;
	float v_f = fmin_(fmax_(f32_mul_add(vec.y, half_depth, 0.5f), 0.0f), ONE_MINUS_EPS) * cubemap_size_f;
	uint32_t u = (uint32_t)u_f;
	uint32_t v = (uint32_t)v_f;
	ORACLE_ASSERT(u < l.shadow_map_size);
	ORACLE_ASSERT(v < l.shadow_map_size);
	size_t texel_address = (size_t)u + (size_t)v * l.shadow_map_size;
	float value = shadow_maps[(size_t)l.shadow_map_offset + (size_t)side * l.shadow_map_size * l.shadow_map_size + texel_address];
	return depth >= value ? 1.0f : 0.0f;
}

inline float cube_shadow_map_fetch(const sqw_light& l, const float* sm, Vec3 vec) // :196
{
	Vec3 a = {std::fabs(vec.x), std::fabs(vec.y), std::fabs(vec.z)};
	if (a.x >= a.y && a.x >= a.z)
	{
		if (vec.x >= 0.0f)
			return cube_shadow_map_side_fetch(l, sm, Vec3{-vec.y, vec.z, a.x}, 1);
		else
			return cube_shadow_map_side_fetch(l, sm, Vec3{vec.y, vec.z, a.x}, 0);
	}
	else if (a.y >= a.x && a.y >= a.z)
	{
		if (vec.y >= 0.0f)
			return cube_shadow_map_side_fetch(l, sm, Vec3{vec.x, vec.z, a.y}, 3);
		else
			return cube_shadow_map_side_fetch(l, sm, Vec3{-vec.x, vec.z, a.y}, 2);
	}
	else
	{
		if (vec.z >= 0.0f)
			return cube_shadow_map_side_fetch(l, sm, Vec3{-vec.x, vec.y, a.z}, 5);
		else
			return cube_shadow_map_side_fetch(l, sm, Vec3{-vec.x, -vec.y, a.z}, 4);
	}
}

inline float projector_shadow_map_fetch(const sqw_light& l, const float* sm, Vec3 vec) // :261
{
	Vec3 bx = {l.basis_x[0], l.basis_x[1], l.basis_x[2]};
	Vec3 by = {l.basis_y[0], l.basis_y[1], l.basis_y[2]};
	Vec3 bz = {l.basis_z[0], l.basis_z[1], l.basis_z[2]};
	float z = vec3_dot(bz, vec);
	if (z <= 0.0f)
		return 0.0f;
	float depth = inv_fast(z);
	float half_depth = 0.5f * depth;
	float size_f = (float)l.shadow_map_size;
	float u_f = f32_mul_add(vec3_dot(bx, vec), half_depth, 0.5f) * size_f;
	float v_f = f32_mul_add(vec3_dot(by, vec), half_depth, 0.5f) * size_f;
	if (u_f < 0.0f || v_f < 0.0f || u_f >= size_f || v_f >= size_f)
		return 0.0f;
	uint32_t u = (uint32_t)u_f;
	uint32_t v = (uint32_t)v_f;
	ORACLE_ASSERT(u < l.shadow_map_size);
	ORACLE_ASSERT(v < l.shadow_map_size);
	float value = sm[(size_t)l.shadow_map_offset + (size_t)u + (size_t)v * l.shadow_map_size];
	return depth >= value ? 1.0f : 0.0f;
}

inline float get_light_shadow_factor(const sqw_light& l, const float* sm, Vec3 vec) // :185
{
	switch (l.shadow_type)
	{
	case SQW_SHADOW_CUBE: return cube_shadow_map_fetch(l, sm, vec);
	case SQW_SHADOW_PROJECTOR: return projector_shadow_map_fetch(l, sm, vec);
	default: return 1.0f;
	}
}

// ---- surfaces.rs:973-1086 lightmap element ops -----------------------------------------------------------------
struct LmSimple
{
	typedef struct
	{
		float c[3];
	} El;
	static constexpr bool directional = false;
	static El mix(const El& a, const El& b, float ratio) // :996
	{
		float omr = 1.0f - ratio;
		return {{f32_mul_add(a.c[0], ratio, b.c[0] * omr), f32_mul_add(a.c[1], ratio, b.c[1] * omr), f32_mul_add(a.c[2], ratio, b.c[2] * omr)}};
	}
	static const float* constant(const El& e) { return e.c; }
};
struct LmDirectional
{
	typedef sqw_directional_lightmap_element El;
	static constexpr bool directional = true;
	static El mix(const El& a, const El& b, float ratio) // :1022
	{
		float omr = 1.0f - ratio;
		El r;
		for (int i = 0; i < 3; ++i)
			r.ambient_light[i] = f32_mul_add(a.ambient_light[i], ratio, b.ambient_light[i] * omr);
		for (int i = 0; i < 3; ++i)
			r.light_direction_vector_scaled[i] = f32_mul_add(a.light_direction_vector_scaled[i], ratio, b.light_direction_vector_scaled[i] * omr);
		r.directional_light_deviation = f32_mul_add(a.directional_light_deviation, ratio, b.directional_light_deviation * omr);
		for (int i = 0; i < 3; ++i)
			r.directional_light_color[i] = f32_mul_add(a.directional_light_color[i], ratio, b.directional_light_color[i] * omr);
		return r;
	}
	static const float* constant(const El& e) { return e.ambient_light; }
};

inline float get_specular_intensity(float angle_cos, float inv_roughness) // :1088 (new formula branch)
{
	float x = (angle_cos - 1.0f) * inv_roughness;
	const float SQRT_7 = 2.645751311f;
	const float A = 2.0f * 3.14159265358979323846f / SQRT_7;
	const float B = 3.14159265358979323846f * SQRT_7 / 2.0f;
	return inv_fast(f32_mul_add(x * x, B, A)) * inv_roughness;
}
inline float get_fresnel_factor_base(float angle_cos) // :1116
{
	float omc = fmax_(1.0f - angle_cos, 0.0f);
	float omc2 = omc * omc;
	return omc2 * omc2 * omc;
}
inline float get_specular_k_dielectric(float fresnel_factor_base, float roughness) // :1125
{
	float fresnel_factor = f32_mul_add(fresnel_factor_base, 1.0f - DIELECTRIC_ZERO_REFLECTIVITY, DIELECTRIC_ZERO_REFLECTIVITY);
	return f32_mul_add(fresnel_factor, 1.0f - roughness, DIELECTRIC_AVERAGE_REFLECTIVITY * roughness);
}
inline float get_specular_k_metal(float fresnel_factor_base, float roughness) // :1143
{
	return f32_mul_add(fresnel_factor_base, 1.0f - roughness, METAL_AVERAGE_SCHLICK_FACTOR * roughness);
}

struct SurfaceArgs
{
	Vec3 u, v, start, normal;
	uint32_t surface_size[2];
	int32_t surface_tc_min[2];
	const TextureElement* texture;
	uint32_t texture_size[2];
	uint32_t lightmap_size[2];
	uint32_t lightmap_scale_log2; // NO_LIGHTMAP_SCALE when the lightmap slice is empty (surfaces.rs:140)
	uint32_t lightmap_tc_shift[2];
	const void* lightmap_data;
	const sqw_light* const* dynamic_lights;
	uint32_t num_dynamic_lights;
	const float* shadow_maps;
	Vec3 cam_pos;
};

// surfaces.rs:433-852 build_surface_impl_6_static_params
template <class C, class Lm, bool USE_DYNAMIC_LIGHTS, bool USE_NORMAL_MAP, uint32_t SPECULAR_TYPE>
void build_surface_impl(const SurfaceArgs& a, typename C::Raw* out_surface_data)
{
	typedef typename Lm::El El;
	const uint32_t LIGHTAP_SCALE_LOG2 = a.lightmap_scale_log2;
	const Vec3 u_vec = a.u, v_vec = a.v;
	// :456 Add 0.5 to get position for pixel centers.
	const Vec3 start_pos = a.start + u_vec * ((float)a.surface_tc_min[0] + 0.5f) + v_vec * ((float)a.surface_tc_min[1] + 0.5f);

	const Vec3 u_vec_normalized = u_vec * inv_sqrt_fast(fmax_(vec3_len2(u_vec), MIN_POSITIVE_VALUE)); // :461
	const Vec3 v_vec_normalized = v_vec * inv_sqrt_fast(fmax_(vec3_len2(v_vec), MIN_POSITIVE_VALUE)); // :462

	El line_lightmap[MAX_LIGHTMAP_SIZE + 2];
	std::memset(line_lightmap, 0, sizeof(line_lightmap));
	const bool has_lightmap = LIGHTAP_SCALE_LOG2 != NO_LIGHTMAP_SCALE;
	const float lightmap_scale_f = (float)(1u << (LIGHTAP_SCALE_LOG2 & 31));
	const float inv_lightmap_scale_f = 1.0f / lightmap_scale_f;
	const float k_shift = 0.5f * inv_lightmap_scale_f;
	const uint32_t lightmap_fetch_mask = (1u << (LIGHTAP_SCALE_LOG2 & 31)) - 1;
	const El* lightmap_data = (const El*)a.lightmap_data;

	for (uint32_t dst_v = 0; dst_v < a.surface_size[1]; ++dst_v)
	{
		if (has_lightmap)
		{
			uint32_t lightmap_base_v = dst_v + a.lightmap_tc_shift[1];
			uint32_t lightmap_v = lightmap_base_v >> LIGHTAP_SCALE_LOG2;
			uint32_t lightmap_v_plus_one = lightmap_v + 1;
			ORACLE_ASSERT(lightmap_v_plus_one < a.lightmap_size[1]); // :482
			float k = f32_mul_add((float)(lightmap_base_v & lightmap_fetch_mask), inv_lightmap_scale_f, k_shift);
			uint32_t base_lightmap_address = lightmap_v * a.lightmap_size[0];
			uint32_t n = std::min<uint32_t>(a.lightmap_size[0], MAX_LIGHTMAP_SIZE + 2); // zip() stops at the shortest
			for (uint32_t i = 0; i < n; ++i)
				line_lightmap[i] = Lm::mix(lightmap_data[base_lightmap_address + a.lightmap_size[0] + i], lightmap_data[base_lightmap_address + i], k);
		}

		typename C::Raw* dst_line = out_surface_data + (size_t)dst_v * a.surface_size[0];

		int32_t ts1 = (int32_t)a.texture_size[1];
		int32_t src_v = (a.surface_tc_min[1] + (int32_t)dst_v) % ts1; // rem_euclid :507
		if (src_v < 0)
			src_v += ts1;
		const TextureElement* src_line = a.texture + (size_t)src_v * a.texture_size[0];
		int32_t ts0 = (int32_t)a.texture_size[0];
		int32_t src_u = a.surface_tc_min[0] % ts0;
		if (src_u < 0)
			src_u += ts0;
		const Vec3 start_pos_v = vec3_scalar_mul_add(v_vec, (float)dst_v, start_pos);
		for (uint32_t dst_u = 0; dst_u < a.surface_size[0]; ++dst_u)
		{
			const Vec3 pos = vec3_scalar_mul_add(u_vec, (float)dst_u, start_pos_v);

			const TextureElement texel_value = src_line[src_u];
			Vec3 texel_normal = {0.0f, 0.0f, 1.0f};
			float texel_roughness = 0.0f;
			if (USE_NORMAL_MAP)
				unpack_normal_roughness(texel_value.packed_normal_roughness, texel_normal, texel_roughness);
			else if (SPECULAR_TYPE != SPECULAR_TYPE_NONE)
				texel_roughness = unpack_roughness(texel_value.packed_normal_roughness);

			ColorVec total_light_albedo_modulated = cv_zero();
			ColorVec total_light_direct = cv_zero();
			if (has_lightmap)
			{
				uint32_t lightmap_base_u = dst_u + a.lightmap_tc_shift[0];
				uint32_t lightmap_u = lightmap_base_u >> LIGHTAP_SCALE_LOG2;
				uint32_t lightmap_u_plus_one = lightmap_u + 1;
				ORACLE_ASSERT(lightmap_u_plus_one < a.lightmap_size[0]); // :541
				const El l0 = line_lightmap[lightmap_u];
				const El l1 = line_lightmap[lightmap_u_plus_one];
				float k = f32_mul_add((float)(lightmap_base_u & lightmap_fetch_mask), inv_lightmap_scale_f, k_shift);
				const El l_mixed = Lm::mix(l1, l0, k);

				if (SPECULAR_TYPE == SPECULAR_TYPE_NONE)
				{
					total_light_albedo_modulated = cv_from_f32x3_with_one(Lm::constant(l_mixed));
					if constexpr (Lm::directional)
					{
						const Vec3 vs = {l_mixed.light_direction_vector_scaled[0], l_mixed.light_direction_vector_scaled[1], l_mixed.light_direction_vector_scaled[2]};
						float dot = USE_NORMAL_MAP ? fmax_(vec3_dot(vs, texel_normal), 0.0f) : vs.z;
						total_light_albedo_modulated = cv_mul_scalar_add(cv_from_f32x3(l_mixed.directional_light_color), dot, total_light_albedo_modulated);
					}
				}
				else
				{
					const Vec3 normal = USE_NORMAL_MAP ? texel_normal : Vec3{0.0f, 0.0f, 1.0f};

					const Vec3 vec_to_camera = a.cam_pos - pos;
					const Vec3 vec_to_camera_texture_space = {
						vec3_dot(vec_to_camera, u_vec_normalized), vec3_dot(vec_to_camera, v_vec_normalized), vec3_dot(vec_to_camera, a.normal)};

					float vec_to_camera_normal_dot = vec3_dot(vec_to_camera_texture_space, normal);
					const Vec3 vec_to_camera_reflected = normal * (2.0f * vec_to_camera_normal_dot) - vec_to_camera_texture_space;
					float vec_to_camera_len2 = fmax_(vec3_len2(vec_to_camera_reflected), MIN_POSITIVE_VALUE);

					float vec_to_camera_normal_angle_cos = fmax_(vec_to_camera_normal_dot * inv_sqrt_fast(vec_to_camera_len2), 0.0f);
					float fresnel_factor_base = get_fresnel_factor_base(vec_to_camera_normal_angle_cos);

					float specular_k = SPECULAR_TYPE == SPECULAR_TYPE_DIELECTRIC ? get_specular_k_dielectric(fresnel_factor_base, texel_roughness)
																			  : get_specular_k_metal(fresnel_factor_base, texel_roughness);
					float one_minus_specular_k = 1.0f - specular_k;

					const ColorVec constant_component = cv_from_f32x3(Lm::constant(l_mixed));
					total_light_albedo_modulated = cv_scalar_mul(constant_component, one_minus_specular_k);
					total_light_direct = cv_scalar_mul(constant_component, specular_k);
					total_light_albedo_modulated.v[3] = 1.0f; // :621

					if constexpr (Lm::directional)
					{
						const Vec3 vs = {l_mixed.light_direction_vector_scaled[0], l_mixed.light_direction_vector_scaled[1], l_mixed.light_direction_vector_scaled[2]};
						float direction_vec_len2 = vec3_len2(vs);
						float direction_vec_len = direction_vec_len2 * inv_sqrt_fast(direction_vec_len2);

						float refl_light_angle_cos = vec3_dot(vec_to_camera_reflected, vs) * inv_sqrt_fast(vec_to_camera_len2 * direction_vec_len2);

						float inv_roughness_corrected = fmax_(inv_fast(texel_roughness + l_mixed.directional_light_deviation), 0.75f);
						float specular_intensity = get_specular_intensity(refl_light_angle_cos, inv_roughness_corrected);

						const ColorVec dir_color = cv_from_f32x3(l_mixed.directional_light_color);
						if (SPECULAR_TYPE == SPECULAR_TYPE_DIELECTRIC)
						{
							float diffuse_intensity = fmax_(vec3_dot(vs, texel_normal), 0.0f);
							float light_intensity_diffuse = diffuse_intensity * one_minus_specular_k;
							total_light_albedo_modulated = cv_mul_scalar_add(dir_color, light_intensity_diffuse, total_light_albedo_modulated);
							float light_intensity_specular = specular_intensity * specular_k * direction_vec_len;
							total_light_direct = cv_mul_scalar_add(dir_color, light_intensity_specular, total_light_direct);
						}
						else
						{
							float scale = specular_intensity * direction_vec_len;
							float light_intensity_modulated = one_minus_specular_k * scale;
							total_light_albedo_modulated = cv_mul_scalar_add(dir_color, light_intensity_modulated, total_light_albedo_modulated);
							float light_intensity_direct = specular_k * scale;
							total_light_direct = cv_mul_scalar_add(dir_color, light_intensity_direct, total_light_direct);
						}
					}
				}
			}

			if (USE_DYNAMIC_LIGHTS)
			{
				Vec3 normal;
				if (USE_NORMAL_MAP)
				{
					normal = { // :686-702
						f32_mul_add(texel_normal.x, u_vec_normalized.x, f32_mul_add(texel_normal.y, v_vec_normalized.x, texel_normal.z * a.normal.x)),
						f32_mul_add(texel_normal.x, u_vec_normalized.y, f32_mul_add(texel_normal.y, v_vec_normalized.y, texel_normal.z * a.normal.y)),
						f32_mul_add(texel_normal.x, u_vec_normalized.z, f32_mul_add(texel_normal.y, v_vec_normalized.z, texel_normal.z * a.normal.z))};
				}
				else
					normal = a.normal;

				Vec3 vec_to_camera_reflected = {0, 0, 0};
				float vec_to_camera_len2 = MIN_POSITIVE_VALUE;
				float specular_k = 0.0f;
				float inv_roughness = 0.0f;
				if (SPECULAR_TYPE != SPECULAR_TYPE_NONE)
				{
					const Vec3 vec_to_camera = a.cam_pos - pos;
					float vec_to_camera_normal_dot = vec3_dot(vec_to_camera, normal);
					vec_to_camera_reflected = normal * (2.0f * vec_to_camera_normal_dot) - vec_to_camera;
					vec_to_camera_len2 = fmax_(vec3_len2(vec_to_camera_reflected), MIN_POSITIVE_VALUE);
					float angle_cos = fmax_(vec_to_camera_normal_dot * inv_sqrt_fast(vec_to_camera_len2), 0.0f);
					float fresnel_factor_base = get_fresnel_factor_base(angle_cos);
					specular_k = SPECULAR_TYPE == SPECULAR_TYPE_DIELECTRIC ? get_specular_k_dielectric(fresnel_factor_base, texel_roughness)
																		 : get_specular_k_metal(fresnel_factor_base, texel_roughness);
					inv_roughness = inv_fast(texel_roughness);
				}

				for (uint32_t li = 0; li < a.num_dynamic_lights; ++li)
				{
					const sqw_light& light = *a.dynamic_lights[li];
					const Vec3 vec_to_light = Vec3{light.position[0], light.position[1], light.position[2]} - pos;

					float shadow_factor = get_light_shadow_factor(light, a.shadow_maps, vec_to_light);
					float vec_to_light_len2 = fmax_(vec3_len2(vec_to_light), MIN_POSITIVE_VALUE);
					float distance_factor = fmax_(inv_fast(vec_to_light_len2) - light.inv_square_radius, 0.0f);
					float shadow_distance_factor = shadow_factor * distance_factor;

					float diffuse_intensity =
						SPECULAR_TYPE == SPECULAR_TYPE_METAL ? 0.0f : fmax_(vec3_dot(normal, vec_to_light) * inv_sqrt_fast(vec_to_light_len2), 0.0f);

					float specular_intensity = 0.0f;
					if (SPECULAR_TYPE != SPECULAR_TYPE_NONE)
					{
						float c = vec3_dot(vec_to_camera_reflected, vec_to_light) * inv_sqrt_fast(vec_to_camera_len2 * vec_to_light_len2);
						specular_intensity = get_specular_intensity(c, inv_roughness);
					}

					const ColorVec light_color = cv_from_f32x3(light.color);
					if (SPECULAR_TYPE == SPECULAR_TYPE_NONE)
					{
						total_light_albedo_modulated = cv_mul_scalar_add(light_color, diffuse_intensity * shadow_distance_factor, total_light_albedo_modulated);
					}
					else if (SPECULAR_TYPE == SPECULAR_TYPE_DIELECTRIC)
					{
						float light_intensity_diffuse = diffuse_intensity * (1.0f - specular_k) * shadow_distance_factor;
						total_light_albedo_modulated = cv_mul_scalar_add(light_color, light_intensity_diffuse, total_light_albedo_modulated);
						float light_intensity_specular = specular_intensity * specular_k * shadow_distance_factor;
						total_light_direct = cv_mul_scalar_add(light_color, light_intensity_specular, total_light_direct);
					}
					else
					{
						float sisdf = specular_intensity * shadow_distance_factor;
						float light_intensity_modulated = (1.0f - specular_k) * sisdf;
						total_light_albedo_modulated = cv_mul_scalar_add(light_color, light_intensity_modulated, total_light_albedo_modulated);
						float light_intensity_direct = specular_k * sisdf;
						total_light_direct = cv_mul_scalar_add(light_color, light_intensity_direct, total_light_direct);
					}
				}
			}

			ColorVec result_color = cv_mul(cv_from_color32(texel_value.diffuse), total_light_albedo_modulated); // :833
			if (SPECULAR_TYPE != SPECULAR_TYPE_NONE)
				result_color = cv_mul_scalar_add(total_light_direct, 255.0f, result_color);

			dst_line[dst_u] = C::from_vec(result_color);
			src_u += 1;
			if (src_u == ts0)
				src_u = 0;
		}
	}
}

// Dispatch chain surfaces.rs:124-423
template <class C, class Lm, bool DYN, bool NM>
void dispatch_specular(const SurfaceArgs& a, uint32_t flags, typename C::Raw* out)
{
	if (flags & SQW_SURF_HAS_NON_ONE_ROUGHNESS)
	{
		if (flags & SQW_SURF_IS_METAL)
			build_surface_impl<C, Lm, DYN, NM, SPECULAR_TYPE_METAL>(a, out);
		else
			build_surface_impl<C, Lm, DYN, NM, SPECULAR_TYPE_DIELECTRIC>(a, out);
	}
	else
		build_surface_impl<C, Lm, DYN, NM, SPECULAR_TYPE_NONE>(a, out);
}
template <class C, class Lm>
void dispatch_surface(const SurfaceArgs& a, uint32_t flags, typename C::Raw* out)
{
	const bool dyn = a.num_dynamic_lights != 0;
	const bool nm = (flags & SQW_SURF_HAS_NORMAL_MAP) != 0;
	if (dyn)
	{
		if (nm)
			dispatch_specular<C, Lm, true, true>(a, flags, out);
		else
			dispatch_specular<C, Lm, true, false>(a, flags, out);
	}
	else
	{
		if (nm)
			dispatch_specular<C, Lm, false, true>(a, flags, out);
		else
			dispatch_specular<C, Lm, false, false>(a, flags, out);
	}
}

// surfaces.rs:903-971 mix_surface_with_texture_impl
template <class C>
void mix_surface_with_texture(
	const uint32_t surface_size[2], const int32_t surface_tc_min[2], const uint32_t* texture, const uint32_t texture_size[2], int blending_mode,
	const float light[3], typename C::Raw* surface_data)
{
	const int LIGHT_SHIFT = 8;
	const float light_scale = (float)(1 << LIGHT_SHIFT);
	const float ls[3] = {light[0] * light_scale, light[1] * light_scale, light[2] * light_scale};
	const ColorVecI light_vec = cvi_from_f32x3(ls);

	for (uint32_t dst_v = 0; dst_v < surface_size[1]; ++dst_v)
	{
		typename C::Raw* dst_line = surface_data + (size_t)dst_v * surface_size[0];
		int32_t ts1 = (int32_t)texture_size[1], ts0 = (int32_t)texture_size[0];
		int32_t src_v = (surface_tc_min[1] + (int32_t)dst_v) % ts1;
		if (src_v < 0)
			src_v += ts1;
		const uint32_t* src_line = texture + (size_t)src_v * texture_size[0];
		int32_t src_u = surface_tc_min[0] % ts0;
		if (src_u < 0)
			src_u += ts0;
		for (uint32_t x = 0; x < surface_size[0]; ++x)
		{
			uint32_t texel_value = src_line[src_u];
			ColorVecI m = cvi_shr<LIGHT_SHIFT>(cvi_mul(cvi_from_color32(texel_value), light_vec));
			typename C::Raw& dst = dst_line[x];
			switch (blending_mode)
			{
			case BLEND_NONE: dst = C::from_veci(m); break;
			case BLEND_AVERAGE: dst = C::from_veci(cvi_shr<1>(cvi_add(m, C::to_veci(dst)))); break;
			case BLEND_ADDITIVE: dst = C::from_veci(cvi_add(m, C::to_veci(dst))); break;
			case BLEND_ALPHA_TEST:
				if (Color32::test_alpha(texel_value))
					dst = C::from_veci(m);
				break;
			case BLEND_ALPHA_BLEND:
			{
				int32_t alpha = Color32::get_alpha(texel_value);
				dst = C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(m, alpha), cvi_mul_scalar(C::to_veci(dst), 255 - alpha))));
				break;
			}
			}
			src_u += 1;
			if (src_u == ts0)
				src_u = 0;
		}
	}
}

struct SurfacesCall
{
	const sqw_surface_job* jobs;
	uint32_t num_jobs;
	const sqw_light* lights;
	uint32_t num_lights;
	const float* shadow_maps;
	Vec3 cam_pos;
	const OracleImage* images;
	uint32_t num_images;
	const float* lightmaps;
	const sqw_directional_lightmap_element* directional_lightmaps;
	void* out;
};

// One iteration of the closure in PartialRenderer::build_polygons_surfaces (partial_renderer.rs:1682-1820).
template <class C>
void run_job(const SurfacesCall& c, const sqw_surface_job& j)
{
	typename C::Raw* out = (typename C::Raw*)c.out + j.surface_pixels_offset;
	if (j.flags & SQW_SURF_APPLY_DIFFUSE)
	{
		const OracleImage& img = c.images[j.texture_image];
		SurfaceArgs a;
		a.u = {j.basis_u[0], j.basis_u[1], j.basis_u[2]};
		a.v = {j.basis_v[0], j.basis_v[1], j.basis_v[2]};
		a.start = {j.basis_start[0], j.basis_start[1], j.basis_start[2]};
		a.normal = {j.basis_normal[0], j.basis_normal[1], j.basis_normal[2]};
		a.surface_size[0] = j.surface_size[0];
		a.surface_size[1] = j.surface_size[1];
		a.surface_tc_min[0] = j.tc_start[0];
		a.surface_tc_min[1] = j.tc_start[1];
		a.texture = (const TextureElement*)img.pixels;
		a.texture_size[0] = img.width;
		a.texture_size[1] = img.height;
		a.lightmap_size[0] = j.lightmap_size[0];
		a.lightmap_size[1] = j.lightmap_size[1];
		a.lightmap_tc_shift[0] = j.lightmap_tc_shift[0];
		a.lightmap_tc_shift[1] = j.lightmap_tc_shift[1];
		const sqw_light* lights[SQW_MAX_POLYGON_LIGHTS];
		a.num_dynamic_lights = std::min<uint32_t>(j.num_lights, SQW_MAX_POLYGON_LIGHTS);
		for (uint32_t i = 0; i < a.num_dynamic_lights; ++i)
			lights[i] = &c.lights[j.lights[i]];
		a.dynamic_lights = lights;
		a.shadow_maps = c.shadow_maps;
		a.cam_pos = c.cam_pos;
		const bool directional = (j.flags & SQW_SURF_DIRECTIONAL_LIGHTMAP) != 0;
		const bool has_lm = j.lightmap_offset != 0;
		a.lightmap_scale_log2 = has_lm ? j.lightmap_scale_log2 : NO_LIGHTMAP_SCALE; // surfaces.rs:140
		if (directional)
		{
			a.lightmap_data = has_lm ? (const void*)(c.directional_lightmaps + j.lightmap_offset) : nullptr;
			dispatch_surface<C, LmDirectional>(a, j.flags, out);
		}
		else
		{
			a.lightmap_data = has_lm ? (const void*)(c.lightmaps + (size_t)j.lightmap_offset * 3) : nullptr;
			dispatch_surface<C, LmSimple>(a, j.flags, out);
		}
	}
	if (j.flags & SQW_SURF_EMISSIVE)
	{
		const OracleImage& img = c.images[j.emissive_image];
		const uint32_t ts[2] = {img.width, img.height};
		mix_surface_with_texture<C>(
			j.surface_size, j.tc_start, (const uint32_t*)img.pixels, ts, (j.flags & SQW_SURF_APPLY_DIFFUSE) ? BLEND_ADDITIVE : BLEND_NONE,
			j.emissive_light, out);
	}
}

template <class C>
void run_jobs(const SurfacesCall& c, int num_threads)
{
	if (num_threads <= 1)
	{
		for (uint32_t i = 0; i < c.num_jobs; ++i)
			run_job<C>(c, c.jobs[i]);
		return;
	}
	// rayon par_iter over visible polygons (partial_renderer.rs:1822-1831): dynamic work distribution over the persistent pool.
	const uint32_t chunk = 8;
	ThreadPool::instance().parallel_for((c.num_jobs + chunk - 1) / chunk, num_threads, [&](uint32_t t) {
		const uint32_t e = std::min((t + 1) * chunk, c.num_jobs);
		for (uint32_t i = t * chunk; i < e; ++i)
			run_job<C>(c, c.jobs[i]);
	});
}

} // namespace
} // namespace oracle

extern "C" int oracle_build_surfaces(
	int color_bits, const sqw_surface_job* jobs, uint32_t num_jobs, const sqw_light* lights, uint32_t num_lights, const float* shadow_maps,
	const float cam_pos[3], const oracle::OracleImage* images, uint32_t num_images, const float* lightmaps,
	const sqw_directional_lightmap_element* directional_lightmaps, void* out_surfaces, int num_threads)
{
	oracle::SurfacesCall c{jobs, num_jobs, lights, num_lights, shadow_maps, {cam_pos[0], cam_pos[1], cam_pos[2]}, images, num_images,
						   lightmaps, directional_lightmaps, out_surfaces};
	if (color_bits == 32)
		oracle::run_jobs<oracle::Color32>(c, num_threads);
	else if (color_bits == 64)
		oracle::run_jobs<oracle::Color64>(c, num_threads);
	else
		return -1;
	return 0;
}

// Standalone mix_surface_with_texture with an explicit blend mode (portal overlays, partial_renderer.rs:2261-2271).
extern "C" int oracle_mix_surface_with_texture(
	int color_bits, const uint32_t surface_size[2], const int32_t surface_tc_min[2], const uint32_t* texture, const uint32_t texture_size[2],
	int blending_mode, const float light[3], void* surface_data)
{
	if (color_bits == 32)
		oracle::mix_surface_with_texture<oracle::Color32>(surface_size, surface_tc_min, texture, texture_size, blending_mode, light, (uint32_t*)surface_data);
	else
		oracle::mix_surface_with_texture<oracle::Color64>(surface_size, surface_tc_min, texture, texture_size, blending_mode, light, (uint64_t*)surface_data);
	return 0;
}
