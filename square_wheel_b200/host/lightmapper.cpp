// Lightmapper: direct light of the map's light entities and its sun into simple + directional lightmaps (SURVEY.md §8 row f4:
// "a direct-light-only lightmapper emitting valid simple + directional lightmaps").  A C++ restatement of what the reference's
// lightmapper computes when run with `--num-passes 1 --no-emissive-surfaces-light --no-secondary-light` (paths relative to
// /root/reference/src/common/):
//   map_lights.rs             extract_map_lights, extract_sun_lights
//   light_trace.rs            opacity / sky tables, get_shadow_factor, get_sun_shadow_factor (BSP ray walk)
//   lightmap.rs:36            calculate_lightmap_basis
//   lightmaps_builder.rs      allocate_lightmaps :299, build_primary_lightmap :467, build_polygon_diretional_lightmap :969 (primary
//                             and sun lights), correct_sample_position :2122
//   light_hemisphere.rs       add_point_light, extract_ambient_light, calculate_light_direction
//   light_cube.rs             LightCube (add_light_sample, scale, convert_into_light_grid_sample)
// The PVS that limits the lights per leaf (pvs.rs) is restated too and used on request (LightmappingSettings::use_pvs, `--use-pvs`);
// by default a contribution cut-off (LightmappingSettings::light_cutoff) picks the lights of a polygon, which is what the demo map
// of the benchmark was lit with.  Not restated: secondary (bounced) and emissive-surface light.  The light grid (prepare_light_grid :1614, calculate_light_grid :1643,
// calculate_light_for_grid_point :1861 for primary and sun lights, correct_light_grid_sample_position :2005, compress_light_grid
// :1772, light_cube.rs) is built from the same direct light.  Host-only, offline.  Parity with the Rust tool: UNPINNED, as for the
// map compiler.
#include "map_compiler.hpp"
#include "renderer.hpp" // ClippingPolygon (common/clipping_polygon.rs)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <unordered_map>

namespace sqwh
{
namespace mapc
{
namespace
{
const float MIN_POSITIVE_VALUE = 1.0f / (float)(1 << 30);
const float TEXEL_NORMAL_SHIFT = 1.0f / 16.0f;
const uint32_t MAX_SAMPLE_GRID_SIZE = 8;
const float MAP_LIGHTS_SCALE = 32.0f; // map_lights.rs:248

inline float magnitude2(Vec3f v) { return dot(v, v); }

struct LightDirection
{
	Vec3f dir_normalized;
	float half_angle_cos;
};
struct PointLight // map_lights.rs:4-10
{
	Vec3f pos;
	float color[3];
	bool has_direction = false;
	LightDirection direction;
};
struct SunLight
{
	Vec3f dir;
	float color[3];
};

std::string map_string(const BSPMap& map, StringRef r) { return std::string((const char*)map.strings_data.data() + r.offset, r.size); }
bool starts_with(const std::string& s, const char* prefix) { return s.compare(0, std::strlen(prefix), prefix) == 0; }
float clamp01(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }

std::vector<PointLight> extract_map_lights(const BSPMap& map) // map_lights.rs:27-158
{
	std::vector<PointLight> result;
	std::unordered_map<std::string, Vec3f> named_targets;
	for (const Entity& e : map.entities)
	{
		bool has_origin = false, has_name = false;
		Vec3f origin{0, 0, 0};
		std::string targetname;
		for (uint32_t i = 0; i < e.num_key_value_pairs; ++i)
		{
			const KeyValuePair& kv = map.key_value_pairs[e.first_key_value_pair + i];
			const std::string key = map_string(map, kv.key), value = map_string(map, kv.value);
			if (key == "origin" && parse_vec3(value, origin))
				has_origin = true;
			if (key == "targetname")
			{
				targetname = value;
				has_name = true;
			}
		}
		if (has_origin && has_name)
			named_targets[targetname] = origin;
	}
	for (const Entity& e : map.entities)
	{
		bool is_light = false, has_origin = false, has_intensity = false, has_color = false, has_target = false, has_angle = false;
		Vec3f origin{0, 0, 0}, color{0, 0, 0};
		float intensity = 0, angle = 0;
		std::string target;
		for (uint32_t i = 0; i < e.num_key_value_pairs; ++i)
		{
			const KeyValuePair& kv = map.key_value_pairs[e.first_key_value_pair + i];
			const std::string key = map_string(map, kv.key), value = map_string(map, kv.value);
			if (key == "classname" && starts_with(value, "light"))
				is_light = true;
			if (key == "origin" && parse_vec3(value, origin))
				has_origin = true;
			if (starts_with(key, "light") || key == "_light")
			{
				float v;
				if (parse_key_value_number(value, v))
				{
					intensity = v;
					has_intensity = true;
				}
			}
			if (key == "color" && parse_vec3(value, color))
				has_color = true;
			if (key == "target")
			{
				target = value;
				has_target = true;
			}
			if (key == "angle" && parse_key_value_number(value, angle))
				has_angle = true;
		}
		if (!is_light || !has_origin)
			continue;
		const float scaled = std::fmax(has_intensity ? intensity : 300.0f, 0.0f) * MAP_LIGHTS_SCALE;
		PointLight l;
		l.pos = origin;
		l.color[0] = l.color[1] = l.color[2] = scaled;
		if (has_color)
		{
			l.color[0] *= clamp01(color.x / 255.0f);
			l.color[1] *= clamp01(color.y / 255.0f);
			l.color[2] *= clamp01(color.z / 255.0f);
		}
		if (has_target)
		{
			auto it = named_targets.find(target);
			if (it != named_targets.end())
			{
				const Vec3f dir = it->second - origin;
				const float len = magnitude(dir);
				if (len > 0.0f)
				{
					l.has_direction = true;
					l.direction.dir_normalized = dir / len;
					l.direction.half_angle_cos = std::cos((has_angle ? angle : 40.0f) * (0.5f * 3.14159265358979323846f / 180.0f));
				}
			}
		}
		if (l.color[0] > 0.0f || l.color[1] > 0.0f || l.color[2] > 0.0f)
			result.push_back(l);
	}
	return result;
}

std::vector<SunLight> extract_sun_lights(const BSPMap& map, const Vec3f& bbox_min, const Vec3f& bbox_max) // :160-246
{
	std::vector<SunLight> result;
	if (map.entities.empty())
		return result;
	const float bbox_max_dimension = magnitude(bbox_max - bbox_min);
	const Entity& world = map.entities[0];
	bool has_angles = false, has_intensity = false, has_color = false;
	Vec3f angles{0, 0, 0}, color{0, 0, 0};
	float intensity = 0;
	for (uint32_t i = 0; i < world.num_key_value_pairs; ++i)
	{
		const KeyValuePair& kv = map.key_value_pairs[world.first_key_value_pair + i];
		const std::string key = map_string(map, kv.key), value = map_string(map, kv.value);
		if (key == "_sun_mangle" && parse_vec3(value, angles))
			has_angles = true;
		if (key == "_sunlight" && parse_key_value_number(value, intensity))
			has_intensity = true;
		if (key == "_sunlight_color" && parse_vec3(value, color))
			has_color = true;
	}
	if (!has_intensity)
		return result;
	const float scaled = std::fmax(intensity, 0.0f) * MAP_LIGHTS_SCALE;
	SunLight s;
	s.color[0] = s.color[1] = s.color[2] = scaled;
	if (has_color)
	{
		s.color[0] *= clamp01(color.x / 255.0f);
		s.color[1] *= clamp01(color.y / 255.0f);
		s.color[2] *= clamp01(color.z / 255.0f);
	}
	if (!(s.color[0] > 0.0f || s.color[1] > 0.0f || s.color[2] > 0.0f))
		return result;
	Vec3f dir{0, 0, 1};
	if (has_angles)
	{
		const float deg2rad = 3.14159265358979323846f / 180.0f;
		const float yaw = angles.x * deg2rad, pitch = angles.y * deg2rad, pc = std::cos(pitch);
		dir = -Vec3f{pc * std::cos(yaw), pc * std::sin(yaw), std::sin(pitch)};
	}
	s.dir = dir * bbox_max_dimension;
	result.push_back(s);
	return result;
}

// ---- light_trace.rs -------------------------------------------------------------------------------------------------------
struct Tracer
{
	const BSPMap& map;
	std::vector<float> opacity; // :6-40
	std::vector<char> sky;      // :42-62

	static float signum(float v) { return v != v ? v : (std::signbit(v) ? -1.0f : 1.0f); } // Rust f32::signum

	bool edge_intersects_with_polygon(const Vec3f& v0, const Vec3f& v1, const Polygon& polygon) const // :290-330
	{
		const Plane& plane = polygon.plane;
		const float dist0 = dot(v0, plane.vec) - plane.dist, dist1 = dot(v1, plane.vec) - plane.dist;
		if (signum(dist0) == signum(dist1))
			return false;
		const float dist_sum = dist1 - dist0;
		if (std::fabs(dist_sum) < MIN_POSITIVE_VALUE)
			return false;
		const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
		const Vec3f x = v0 * k1 - v1 * k0;
		for (uint32_t i = 0; i < polygon.num_vertices; ++i)
		{
			const Vec3f v = map.vertices[polygon.first_vertex + i], next_v = map.vertices[polygon.first_vertex + (i + 1) % polygon.num_vertices];
			if (dot(cross(x - v, next_v - v), plane.vec) < 0.0f)
				return false;
		}
		return true;
	}

	float get_shadow_factor_r(const Vec3f& v0, const Vec3f& v1, uint32_t index) const // :83-160
	{
		if (index >= FIRST_LEAF_INDEX)
		{
			const BSPLeaf& leaf = map.leafs[index - FIRST_LEAF_INDEX];
			float acc = 1.0f;
			for (uint32_t i = 0; i < leaf.num_polygons; ++i)
			{
				const Polygon& p = map.polygons[leaf.first_polygon + i];
				const float op = opacity[p.texture];
				if (op >= 1.0f)
					continue;
				if (edge_intersects_with_polygon(v0, v1, p))
				{
					if (op <= 0.0f)
						return 0.0f;
					acc *= op;
				}
			}
			return acc;
		}
		const BSPNode& node = map.nodes[index];
		const float dist0 = dot(v0, node.plane.vec) - node.plane.dist, dist1 = dot(v1, node.plane.vec) - node.plane.dist;
		if (dist0 >= 0.0f && dist1 >= 0.0f)
			return get_shadow_factor_r(v0, v1, node.children[0]);
		if (dist0 <= 0.0f && dist1 <= 0.0f)
			return get_shadow_factor_r(v0, v1, node.children[1]);
		const float dist_sum = dist1 - dist0;
		if (std::fabs(dist_sum) < MIN_POSITIVE_VALUE)
			return 1.0f;
		const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
		const Vec3f x = v0 * k1 - v1 * k0;
		const Vec3f& v_front = dist0 > 0.0f ? v0 : v1;
		const Vec3f& v_back = dist0 > 0.0f ? v1 : v0;
		const float eps = 1.0f / 1024.0f;
		const Vec3f x_front = x * (1.0f - eps) + v_back * eps, x_back = x * (1.0f - eps) + v_front * eps;
		const float f = get_shadow_factor_r(v_front, x_front, node.children[0]);
		if (f <= 0.0f)
			return 0.0f;
		const float b = get_shadow_factor_r(x_back, v_back, node.children[1]);
		if (b <= 0.0f)
			return 0.0f;
		return f * b;
	}
	float get_shadow_factor(const Vec3f& from, const Vec3f& to) const { return get_shadow_factor_r(from, to, map.root_node()); }
	bool can_see(const Vec3f& from, const Vec3f& to) const { return get_shadow_factor(from, to) > 0.0f; }

	bool get_nearest_sky_point_r(const Vec3f& from, const Vec3f& to, uint32_t index, Vec3f& out) const // :178-288
	{
		if (index >= FIRST_LEAF_INDEX)
		{
			const BSPLeaf& leaf = map.leafs[index - FIRST_LEAF_INDEX];
			bool found = false;
			for (uint32_t i = 0; i < leaf.num_polygons; ++i)
			{
				const Polygon& p = map.polygons[leaf.first_polygon + i];
				if (!sky[p.texture])
					continue;
				if (dot(from - to, p.plane.vec) <= 0.0f)
					continue;
				if (edge_intersects_with_polygon(from, to, p))
				{
					const Vec3f pos = get_line_plane_intersection(from, to, p.plane);
					if (!found || magnitude2(from - pos) < magnitude2(from - out))
						out = pos;
					found = true;
				}
			}
			return found;
		}
		const BSPNode& node = map.nodes[index];
		const float dist_from = dot(from, node.plane.vec) - node.plane.dist, dist_to = dot(to, node.plane.vec) - node.plane.dist;
		if (dist_from >= 0.0f && dist_to >= 0.0f)
			return get_nearest_sky_point_r(from, to, node.children[0], out);
		if (dist_from <= 0.0f && dist_to <= 0.0f)
			return get_nearest_sky_point_r(from, to, node.children[1], out);
		const float dist_sum = dist_to - dist_from;
		if (std::fabs(dist_sum) < MIN_POSITIVE_VALUE)
			return false;
		const float k_from = dist_from / dist_sum, k_to = dist_to / dist_sum;
		const Vec3f x = from * k_to - to * k_from;
		const int child_from = dist_from > 0.0f ? 0 : 1, child_to = 1 - child_from;
		const float eps = 1.0f / 1024.0f;
		const Vec3f x_from = x * (1.0f - eps) + from * eps, x_to = x * (1.0f - eps) + to * eps;
		if (get_nearest_sky_point_r(from, x_to, node.children[child_from], out))
			return true;
		return get_nearest_sky_point_r(x_from, to, node.children[child_to], out);
	}

	float get_sun_shadow_factor(const Vec3f& from, const Vec3f& dir) const // :162-176
	{
		Vec3f sky_point;
		if (!get_nearest_sky_point_r(from, from + dir, map.root_node(), sky_point))
			return 0.0f;
		const Vec3f corrected = sky_point - dir / (4.0f * magnitude(dir));
		return get_shadow_factor(from, corrected);
	}
};

// ---- lightmap.rs:28-66 ----------------------------------------------------------------------------------------------------
struct LightmapBasis
{
	Vec3f pos, u_vec, v_vec;
};
bool calculate_lightmap_basis(const Polygon& polygon, LightmapBasis& out)
{
	const float s = 1.0f / (float)LIGHTMAP_SCALE;
	Mat4f m; // columns as in Mat4f::from_cols, then transposed
	m.x = extend(polygon.tex_coord_equation[0].vec, polygon.tex_coord_equation[0].dist) * s;
	m.y = extend(polygon.tex_coord_equation[1].vec, polygon.tex_coord_equation[1].dist) * s;
	m.z = extend(polygon.plane.vec, -polygon.plane.dist);
	m.w = Vec4f{0, 0, 0, 1};
	Mat4f inv;
	if (!m.transpose().invert(inv))
		return false;
	out.u_vec = truncate(inv.x);
	out.v_vec = truncate(inv.y);
	out.pos = truncate(inv.w) + out.u_vec * (float)(polygon.tex_coord_min[0] >> LIGHTMAP_SCALE_LOG2) + out.v_vec * (float)(polygon.tex_coord_min[1] >> LIGHTMAP_SCALE_LOG2);
	return true;
}

Vec3f correct_sample_position(const Tracer& t, const Vec3f& pos, const LightmapBasis& basis, const Vec3f& polygon_center) // lightmaps_builder.rs:2122-2179
{
	if (t.can_see(pos, polygon_center))
		return pos;
	static const float SHIFT_VECS[12][2] = {{0.5f, 0}, {-0.5f, 0}, {0, 0.5f}, {0, -0.5f}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
	for (const float(&shift)[2] : SHIFT_VECS)
	{
		const Vec3f corrected = pos + basis.u_vec * shift[0] + basis.v_vec * shift[1];
		if (t.can_see(corrected, polygon_center))
			return corrected;
	}
	const float max_basis_vec_len = std::fmax(magnitude(basis.u_vec), magnitude(basis.v_vec));
	Vec3f corrected = pos;
	for (int i = 0; i < 16; ++i)
	{
		const Vec3f to_center = polygon_center - corrected;
		const float len = std::fmax(magnitude(to_center), MIN_POSITIVE_VALUE);
		corrected = corrected + (to_center / len) * std::fmin(max_basis_vec_len, len);
		if (t.can_see(corrected, polygon_center))
			return corrected;
	}
	return polygon_center;
}

// ---- light_hemisphere.rs, for lights that each land in ONE pixel (add_point_light) ----------------------------------------------
// The reference keeps a 64 x 64 picture of the hemisphere; with point and sun lights only a handful of pixels are non-zero, so the
// picture is kept as a short list.  extract_ambient_light takes the per-channel MEDIAN over the 3 200-odd pixels of the disc: zero
// unless more than half of them are lit, which such a list never reaches, so ambient = 0 and the pixels stay as they are.
const uint32_t TEXTURE_SIZE = 64;
const float TEXTURE_SIZE_F = 64.0f, HALF_TEXTURE_SIZE_F = 32.0f;
struct LightHemisphere
{
	struct Pixel
	{
		uint32_t x, y;
		float c[3];
	};
	std::vector<Pixel> pixels;

	static uint32_t clamp_to_texture_border(float c) { return (uint32_t)std::fmin(std::fmax(c, 0.0f), TEXTURE_SIZE_F - 1.0f); }
	static Vec3f unproject_texture_coord(uint32_t x, uint32_t y) // :307-325
	{
		const float k = std::sqrt(2.0f) / HALF_TEXTURE_SIZE_F;
		const float cx = ((float)x + 0.5f - HALF_TEXTURE_SIZE_F) * k, cy = ((float)y + 0.5f - HALF_TEXTURE_SIZE_F) * k;
		const float len2 = cx * cx + cy * cy;
		const float xy_scale = std::sqrt(std::fmax(1.0f - len2 * 0.25f, 0.0f));
		return Vec3f{xy_scale * cx, xy_scale * cy, 1.0f - len2 * 0.5f};
	}
	void add_point_light(const Vec3f& direction, const float color[3]) // :24-33, :279-305
	{
		const Vec3f n = direction / magnitude(direction);
		const float k = std::sqrt(2.0f / std::fmax(n.z + 1.0f, 0.0f)) * (HALF_TEXTURE_SIZE_F / std::sqrt(2.0f));
		const uint32_t x = clamp_to_texture_border(n.x * k + HALF_TEXTURE_SIZE_F), y = clamp_to_texture_border(n.y * k + HALF_TEXTURE_SIZE_F);
		for (Pixel& p : pixels)
			if (p.x == x && p.y == y)
			{
				for (int i = 0; i < 3; ++i)
					p.c[i] += color[i];
				return;
			}
		pixels.push_back(Pixel{x, y, {color[0], color[1], color[2]}});
	}
	void calculate_light_direction(sqw_directional_lightmap_element& e) const // :205-262
	{
		float light_sum[3] = {0, 0, 0};
		Vec3f scaled_vecs_sum{0, 0, 0};
		float brightness_sum = 0;
		// the reference scans the picture row by row; same order here so that the float sums agree
		std::vector<Pixel> ordered = pixels;
		std::sort(ordered.begin(), ordered.end(), [](const Pixel& a, const Pixel& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
		for (const Pixel& p : ordered)
		{
			for (int i = 0; i < 3; ++i)
				light_sum[i] += p.c[i];
			const float brightness = p.c[0] * 0.299f + p.c[1] * 0.587f + p.c[2] * 0.114f;
			scaled_vecs_sum = scaled_vecs_sum + unproject_texture_coord(p.x, p.y) * brightness;
			brightness_sum += brightness;
		}
		const float MIN_LEN = 1.0f / (1024.0f * 1024.0f);
		brightness_sum = std::fmax(brightness_sum, MIN_LEN);
		if (scaled_vecs_sum.z < 0.0f)
			scaled_vecs_sum.z = 0.0f;
		float vec_len = magnitude(scaled_vecs_sum);
		if (vec_len <= MIN_LEN)
		{
			scaled_vecs_sum = Vec3f{0, 0, MIN_LEN};
			vec_len = MIN_LEN;
		}
		const float half_deviation_cos = vec_len / brightness_sum;
		e.directional_light_deviation = std::fmin(std::fmax((1.0f - half_deviation_cos) / (1.0f - 0.5f), 0.0f), 1.0f);
		const float lsb = std::fmax(light_sum[0] * 0.299f + light_sum[1] * 0.587f + light_sum[2] * 0.114f, MIN_LEN);
		for (int i = 0; i < 3; ++i)
			e.directional_light_color[i] = light_sum[i] / lsb;
		e.light_direction_vector_scaled[0] = scaled_vecs_sum.x;
		e.light_direction_vector_scaled[1] = scaled_vecs_sum.y;
		e.light_direction_vector_scaled[2] = scaled_vecs_sum.z;
	}
};

// ---- pvs.rs: leaf-to-leaf visibility (LightmappingSettings::use_pvs) ---------------------------------------------------------------
// Per leaf and per portal of it: a breadth-first wave through the portal graph, every portal projected from each vertex of the
// start portal (90 degree views along the portal's normal) and its screen bounds intersected with the bounds under which its leaf
// was reached; a leaf that is ever reached is potentially visible.  The matrix keeps a pair only when each leaf sees the other.
struct Pvs
{
	size_t num_leafs = 0;
	std::vector<std::vector<uint8_t>> rows; // rows[a][b] != 0: b potentially visible from a (symmetrised)
};

size_t clip_polygon_by_z_plane(const Vec3f* polygon, size_t n, float z_dist, Vec3f* out, size_t out_cap) // clipping.rs:4-57
{
	size_t count = 0;
	const Vec3f* prev = &polygon[n - 1];
	for (size_t i = 0; i < n; ++i)
	{
		const Vec3f* v = &polygon[i];
		if (v->z > z_dist)
		{
			if (prev->z < z_dist && count < out_cap)
			{
				const float d0 = prev->z - z_dist, d1 = v->z - z_dist, sum = v->z - prev->z;
				out[count++] = *prev * (d1 / sum) - *v * (d0 / sum);
			}
			if (count < out_cap)
				out[count++] = *v;
		}
		else if (v->z == z_dist)
		{
			if (count < out_cap)
				out[count++] = *v;
		}
		else if (prev->z > z_dist && count < out_cap)
		{
			const float d0 = prev->z - z_dist, d1 = v->z - z_dist, sum = v->z - prev->z;
			out[count++] = *prev * (d1 / sum) - *v * (d0 / sum);
		}
		prev = v;
	}
	return count;
}

bool project_portal(const BSPMap& map, const Portal& portal, const std::vector<Mat4f>& view_matrices, ClippingPolygon& result) // pvs.rs:250-305
{
	const size_t MAX_VERTICES = 24;
	Vec3f transformed[MAX_VERTICES], clipped[MAX_VERTICES + 8];
	bool have = false;
	for (const Mat4f& vm : view_matrices)
	{
		size_t count = std::min<size_t>(portal.num_vertices, MAX_VERTICES);
		for (size_t i = 0; i < count; ++i)
			transformed[i] = view_matrix_transform_vertex(vm, map.vertices[portal.first_vertex + i]);
		count = clip_polygon_by_z_plane(transformed, count, 1.0f / 4096.0f, clipped, MAX_VERTICES + 8);
		if (count < 3)
			continue;
		for (size_t i = 0; i < count; ++i)
		{
			const Vec2f point{clipped[i].x / clipped[i].z, clipped[i].y / clipped[i].z};
			if (have)
				result.extend_with_point(point);
			else
			{
				result = ClippingPolygon::from_point(point);
				have = true;
			}
		}
	}
	return have;
}

void calculate_pvs_for_leaf_portal(const BSPMap& map, uint32_t start_leaf, uint32_t start_portal, std::vector<uint8_t>& visible) // pvs.rs:97-218
{
	const Portal& portal = map.portals[start_portal];
	const uint32_t first = portal.leafs[0] == start_leaf ? portal.leafs[1] : portal.leafs[0];
	struct VisLeaf
	{
		bool has_bounds = false;
		ClippingPolygon bounds;
		size_t last_push_iteration = 0;
	};
	std::vector<VisLeaf> data(map.leafs.size());
	const float inf = 1e36f;
	data[first].has_bounds = true;
	data[first].bounds = ClippingPolygon::from_box(-inf, -inf, inf, inf);
	// calculate_portal_view_matrices :220-248
	Vec3f dir = portal.plane.vec;
	if (portal.leafs[0] == start_leaf)
		dir = -dir;
	const float azimuth = std::atan2(-dir.x, dir.y), elevation = std::atan2(dir.z, std::sqrt(dir.x * dir.x + dir.y * dir.y));
	std::vector<Mat4f> view_matrices;
	for (uint32_t i = 0; i < portal.num_vertices; ++i)
		view_matrices.push_back(build_view_matrix(map.vertices[portal.first_vertex + i], azimuth, elevation, 3.14159265358979323846f * 0.5f, 1.0f, 1.0f).view_matrix);
	std::vector<uint32_t> cur{first}, next;
	size_t iteration = 1;
	while (!cur.empty())
	{
		for (uint32_t leaf_index : cur)
		{
			const ClippingPolygon prev_bounds = data[leaf_index].bounds;
			const BSPLeaf& leaf = map.leafs[leaf_index];
			for (uint32_t k = 0; k < leaf.num_leaf_portals; ++k)
			{
				const Portal& p = map.portals[map.leafs_portals[leaf.first_leaf_portal + k]];
				const uint32_t next_leaf = p.leafs[0] == leaf_index ? p.leafs[1] : p.leafs[0];
				ClippingPolygon bounds;
				if (!project_portal(map, p, view_matrices, bounds))
					continue;
				bounds.intersect(prev_bounds);
				if (!bounds.is_valid_and_non_empty())
					continue;
				VisLeaf& d = data[next_leaf];
				if (d.has_bounds)
				{
					if (d.bounds.contains(bounds))
						continue;
					d.bounds.extend(bounds);
				}
				else
				{
					d.has_bounds = true;
					d.bounds = bounds;
				}
				if (d.last_push_iteration < iteration)
				{
					d.last_push_iteration = iteration;
					next.push_back(next_leaf);
				}
			}
		}
		cur.clear();
		std::swap(cur, next);
		if (++iteration >= 256)
			break;
	}
	for (size_t i = 0; i < data.size(); ++i)
		if (data[i].has_bounds)
			visible[i] = 1;
}

Pvs calculate_visibility_matrix(const BSPMap& map) // pvs.rs:7-67
{
	Pvs pvs;
	pvs.num_leafs = map.leafs.size();
	std::vector<std::vector<uint8_t>> sets(pvs.num_leafs);
	std::atomic<size_t> next{0};
	const unsigned nt = std::min(std::max(std::thread::hardware_concurrency(), 1u), 64u);
	std::vector<std::thread> threads;
	for (unsigned t = 0; t < nt; ++t)
		threads.emplace_back([&] {
			for (;;)
			{
				const size_t leaf_index = next.fetch_add(1);
				if (leaf_index >= pvs.num_leafs)
					break;
				std::vector<uint8_t>& set = sets[leaf_index]; // calculate_pvs_bit_set_for_leaf :71-87
				set.assign(pvs.num_leafs, 0);
				set[leaf_index] = 1;
				const BSPLeaf& leaf = map.leafs[leaf_index];
				for (uint32_t k = 0; k < leaf.num_leaf_portals; ++k)
					calculate_pvs_for_leaf_portal(map, (uint32_t)leaf_index, map.leafs_portals[leaf.first_leaf_portal + k], set);
			}
		});
	for (std::thread& th : threads)
		th.join();
	pvs.rows.assign(pvs.num_leafs, std::vector<uint8_t>(pvs.num_leafs, 0));
	for (size_t x = 0; x < pvs.num_leafs; ++x)
		for (size_t y = 0; y < pvs.num_leafs; ++y)
			pvs.rows[y][x] = sets[x][y] & sets[y][x];
	return pvs;
}

void collect_submodel_leafs_r(const BSPMap& map, const Vec3f (&bbox_vertices)[8], uint32_t node_index, std::vector<uint32_t>& out) // lightmaps_builder.rs:741-775
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		out.push_back(node_index - FIRST_LEAF_INDEX);
		return;
	}
	const BSPNode& node = map.nodes[node_index];
	size_t front = 0;
	for (const Vec3f& v : bbox_vertices)
		if (dot(node.plane.vec, v) > node.plane.dist)
			++front;
	if (front > 0)
		collect_submodel_leafs_r(map, bbox_vertices, node.children[0], out);
	if (front < 8)
		collect_submodel_leafs_r(map, bbox_vertices, node.children[1], out);
}

struct PolygonLighter
{
	const BSPMap& map;
	const Tracer& tracer;
	const std::vector<PointLight>& lights;
	const std::vector<SunLight>& sun_lights;
	const LightmappingSettings& settings;
	uint32_t sample_grid_size;

	// lightmaps_builder.rs:467-603 (simple) and :969-1112 (directional) over the same sample positions
	// `visible_lights`: the lights of the leafs the polygon's leaf (or submodel) can see (build_primary_lightmaps :391-465), or null
	// when the PVS is not used: then every light whose unshadowed contribution can exceed the cut-off
	void light_polygon(const Polygon& polygon, float* simple, sqw_directional_lightmap_element* directional, const std::vector<const PointLight*>* visible_lights = nullptr) const
	{
		const uint32_t size[2] = {get_lightmap_size(polygon.tex_coord_min[0], polygon.tex_coord_max[0]), get_lightmap_size(polygon.tex_coord_min[1], polygon.tex_coord_max[1])};
		const Vec3f normal = polygon.plane.vec / magnitude(polygon.plane.vec);
		Vec3f center{0, 0, 0}; // get_polygon_center :1272-1284
		for (uint32_t i = 0; i < polygon.num_vertices; ++i)
			center = center + map.vertices[polygon.first_vertex + i];
		center = center / (float)polygon.num_vertices + normal * TEXEL_NORMAL_SHIFT;
		LightmapBasis basis;
		if (!calculate_lightmap_basis(polygon, basis))
			return; // (the reference unwraps: a degenerate texture basis does not occur in a compiled map)
		const Vec3f start_pos = basis.pos + normal * TEXEL_NORMAL_SHIFT;
		Vec3f sample_grid[MAX_SAMPLE_GRID_SIZE * MAX_SAMPLE_GRID_SIZE];
		sample_grid[0] = Vec3f{0, 0, 0};
		if (sample_grid_size > 1)
		{
			const float g = (float)sample_grid_size;
			const Vec3f u_step = basis.u_vec / g, v_step = basis.v_vec / g;
			const Vec3f grid_start = (v_step + u_step) * (-0.5f * (g - 1.0f));
			for (uint32_t v = 0; v < sample_grid_size; ++v)
				for (uint32_t u = 0; u < sample_grid_size; ++u)
					sample_grid[u + v * sample_grid_size] = u_step * (float)u + (grid_start + v_step * (float)v);
		}
		const uint32_t num_samples = sample_grid_size * sample_grid_size;
		const float multi_sampling_scale = 1.0f / (float)num_samples;
		const Vec3f u_n = basis.u_vec / std::fmax(magnitude(basis.u_vec), MIN_POSITIVE_VALUE), v_n = basis.v_vec / std::fmax(magnitude(basis.v_vec), MIN_POSITIVE_VALUE);

		// lights that can matter for this polygon (stands in for the reference's per-leaf PVS light list)
		std::vector<const PointLight*> near_lights;
		if (visible_lights)
			near_lights = *visible_lights;
		else
		{
			float radius = 0;
			for (uint32_t i = 0; i < polygon.num_vertices; ++i)
				radius = std::fmax(radius, magnitude(map.vertices[polygon.first_vertex + i] - center));
			radius += 2.0f * (magnitude(basis.u_vec) + magnitude(basis.v_vec));
			for (const PointLight& l : lights)
			{
				if (dot(normal, l.pos - center) <= -radius)
					continue; // behind the polygon's plane for every texel
				const float d = std::fmax(magnitude(l.pos - center) - radius, 1.0f);
				if (std::fmax(std::fmax(l.color[0], l.color[1]), l.color[2]) / (d * d) < settings.light_cutoff)
					continue;
				near_lights.push_back(&l);
			}
		}

		for (uint32_t v = 0; v < size[1]; ++v)
		{
			const Vec3f start_pos_v = start_pos + basis.v_vec * (float)v;
			for (uint32_t u = 0; u < size[0]; ++u)
			{
				float total_light[3] = {0, 0, 0};
				LightHemisphere hemisphere;
				const Vec3f texel_pos = start_pos_v + basis.u_vec * (float)u;
				for (uint32_t s = 0; s < num_samples; ++s)
				{
					const Vec3f pos = correct_sample_position(tracer, texel_pos + sample_grid[s], basis, center);
					for (const PointLight* lp : near_lights)
					{
						const PointLight& light = *lp;
						const Vec3f vec_to_light = light.pos - pos;
						const float len2 = std::fmax(magnitude2(vec_to_light), MIN_POSITIVE_VALUE), len = std::sqrt(len2);
						const float angle_cos = dot(normal, vec_to_light) / len;
						if (angle_cos <= 0.0f)
							continue;
						if (light.has_direction && -dot(vec_to_light, light.direction.dir_normalized) < light.direction.half_angle_cos * len)
							continue;
						const float shadow_factor = tracer.get_shadow_factor(light.pos, pos);
						if (shadow_factor <= 0.0f)
							continue;
						const float light_scale = shadow_factor * angle_cos / len2;
						for (int i = 0; i < 3; ++i)
							total_light[i] += multi_sampling_scale * (light.color[i] * light_scale);
						if (directional)
						{
							const float hs = shadow_factor * multi_sampling_scale / len2;
							const float color_scaled[3] = {light.color[0] * hs, light.color[1] * hs, light.color[2] * hs};
							const Vec3f from_center = light.pos - texel_pos; // one direction for all samples of the texel (sharp specular)
							hemisphere.add_point_light(Vec3f{dot(from_center, u_n), dot(from_center, v_n), dot(from_center, normal)}, color_scaled);
						}
					}
					for (const SunLight& light : sun_lights)
					{
						const float len2 = std::fmax(magnitude2(light.dir), MIN_POSITIVE_VALUE);
						const float angle_cos = dot(normal, light.dir) / std::sqrt(len2);
						if (angle_cos <= 0.0f)
							continue;
						const float shadow_factor = tracer.get_sun_shadow_factor(pos, light.dir);
						if (shadow_factor <= 0.0f)
							continue;
						const float light_scale = shadow_factor * angle_cos;
						for (int i = 0; i < 3; ++i)
							total_light[i] += multi_sampling_scale * (light.color[i] * light_scale);
						if (directional)
						{
							const float hs = shadow_factor * multi_sampling_scale;
							const float color_scaled[3] = {light.color[0] * hs, light.color[1] * hs, light.color[2] * hs};
							hemisphere.add_point_light(Vec3f{dot(light.dir, u_n), dot(light.dir, v_n), dot(light.dir, normal)}, color_scaled);
						}
					}
				}
				const size_t dst = (size_t)polygon.lightmap_data_offset + u + (size_t)v * size[0];
				for (int i = 0; i < 3; ++i)
					simple[dst * 3 + i] = settings.ambient_light + total_light[i]; // build_lightmaps :144-168 with one pass
				if (directional)
				{
					sqw_directional_lightmap_element& e = directional[dst];
					// extract_ambient_light: the median of a nearly empty hemisphere is zero (see LightHemisphere); the tool's
					// constant --ambient-light goes here too so that both lightmap kinds describe the same light
					e.ambient_light[0] = e.ambient_light[1] = e.ambient_light[2] = settings.ambient_light;
					hemisphere.calculate_light_direction(e);
				}
			}
		}
	}
};

// ---- light_cube.rs -------------------------------------------------------------------------------------------------------------
const float MIN_VEC_LEN = 0.00000001f;
struct LightCube
{
	float light_cube[6][3] = {};
	Vec3f sum_scaled_light_dirs[3] = {Vec3f{0, 0, 0}, Vec3f{0, 0, 0}, Vec3f{0, 0, 0}};

	void add_light_sample(const Vec3f& direction, const float light_scaled[3]) // :20-77
	{
		const Vec3f n = direction / std::fmax(magnitude(direction), MIN_VEC_LEN);
		const float comp[3] = {n.x, n.y, n.z};
		for (int a = 0; a < 3; ++a)
		{
			const int side = comp[a] <= 0.0f ? 2 * a : 2 * a + 1;
			for (int i = 0; i < 3; ++i)
				light_cube[side][i] += light_scaled[i] * std::fabs(comp[a]);
		}
		for (int i = 0; i < 3; ++i)
			sum_scaled_light_dirs[i] = sum_scaled_light_dirs[i] + n * light_scaled[i];
	}
	void scale(float k) // :90-103
	{
		for (auto& side : light_cube)
			for (float& c : side)
				c *= k;
		for (Vec3f& v : sum_scaled_light_dirs)
			v = v * k;
	}
	LightGridElement convert_into_light_grid_sample() const // :105-158
	{
		LightGridElement e;
		const Vec3f dir = (sum_scaled_light_dirs[0] + sum_scaled_light_dirs[1] + sum_scaled_light_dirs[2]) * 0.33f;
		const float len2 = std::fmax(magnitude2(dir), MIN_VEC_LEN);
		float color[3];
		for (int i = 0; i < 3; ++i)
			color[i] = std::fmax(dot(sum_scaled_light_dirs[i], dir), 0.0f) / len2;
		const float comp[3] = {dir.x, dir.y, dir.z};
		for (int s = 0; s < 6; ++s)
			for (int i = 0; i < 3; ++i)
				e.light_cube[s][i] = light_cube[s][i];
		for (int i = 0; i < 3; ++i)
			for (int a = 0; a < 3; ++a)
			{
				const int side = comp[a] <= 0.0f ? 2 * a : 2 * a + 1;
				e.light_cube[side][i] = std::fmax(e.light_cube[side][i] - color[i] * std::fabs(comp[a]), 0.0f);
			}
		e.light_direction_vector_scaled = dir;
		for (int i = 0; i < 3; ++i)
			e.directional_light_color[i] = color[i];
		return e;
	}
};

// ---- light grid (lightmaps_builder.rs:1614-2081) -------------------------------------------------------------------------------------
uint32_t leaf_for_point(const BSPMap& map, const Vec3f& p) // bsp_map_compact.rs get_leaf_for_point
{
	uint32_t index = map.root_node();
	while (index < FIRST_LEAF_INDEX)
	{
		const BSPNode& n = map.nodes[index];
		index = dot(n.plane.vec, p) >= n.plane.dist ? n.children[0] : n.children[1];
	}
	return index - FIRST_LEAF_INDEX;
}
bool is_point_inside_leaf_volume(const BSPMap& map, const Vec3f& point) // :2068-2081
{
	const BSPLeaf& leaf = map.leafs[leaf_for_point(map, point)];
	for (uint32_t i = 0; i < leaf.num_polygons; ++i)
	{
		const Polygon& p = map.polygons[leaf.first_polygon + i];
		if (dot(point, p.plane.vec) <= p.plane.dist)
			return false;
	}
	return true;
}
bool correct_light_grid_sample_position(const BSPMap& map, const Vec3f& pos, Vec3f& out) // :2005-2066
{
	if (is_point_inside_leaf_volume(map, pos))
	{
		out = pos;
		return true;
	}
	static const float SHIFTS[26][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1},
		{-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};
	const float* cell = map.light_grid_header.grid_cell_size;
	const int max_step = 3;
	for (int step = 1; step <= max_step; ++step)
	{
		const float f = (float)step / (float)max_step;
		for (const float(&shift)[3] : SHIFTS)
		{
			const Vec3f shifted = pos + Vec3f{shift[0] * (cell[0] * f), shift[1] * (cell[1] * f), shift[2] * (cell[2] * f)};
			if (is_point_inside_leaf_volume(map, shifted))
			{
				out = shifted;
				return true;
			}
		}
	}
	return false;
}
bool light_grid_element_is_zero(const LightGridElement& e) // :1836-1859
{
	const float MIN_LIGHT = 1.0f / 256.0f;
	for (const auto& side : e.light_cube)
		for (float c : side)
			if (c > MIN_LIGHT)
				return false;
	return !(magnitude2(e.light_direction_vector_scaled) > MIN_LIGHT * MIN_LIGHT);
}

uint32_t build_light_grid(const LightmappingSettings& settings, uint32_t sample_grid_size, const Tracer& tracer, const std::vector<PointLight>& lights, const std::vector<SunLight>& sun_lights,
	const Vec3f& bbox_min, const Vec3f& bbox_max, BSPMap& map)
{
	LightGridHeader& h = map.light_grid_header; // prepare_light_grid :1614-1641
	const float width = std::fmin(std::fmax(settings.light_grid_cell_width, 16.0f), 512.0f), height = std::fmin(std::fmax(settings.light_grid_cell_height, 16.0f), 512.0f);
	h.grid_cell_size[0] = h.grid_cell_size[1] = width;
	h.grid_cell_size[2] = height;
	for (int i = 0; i < 3; ++i)
	{
		const float start = std::floor(bbox_min[i] / h.grid_cell_size[i]), end = std::ceil(bbox_max[i] / h.grid_cell_size[i]);
		h.grid_start[i] = start * h.grid_cell_size[i];
		h.grid_size[i] = (uint32_t)(end - start) + 1;
	}
	// calculate_light_grid :1643-1752
	std::vector<Vec3f> shifts(1, Vec3f{0, 0, 0});
	if (sample_grid_size > 1)
	{
		shifts.assign((size_t)sample_grid_size * sample_grid_size * sample_grid_size, Vec3f{0, 0, 0});
		const float g = (float)sample_grid_size;
		const Vec3f step{h.grid_cell_size[0] / g, h.grid_cell_size[1] / g, h.grid_cell_size[2] / g};
		const Vec3f grid_start = step * (-0.5f * (g - 1.0f));
		for (uint32_t x = 0; x < sample_grid_size; ++x)
			for (uint32_t y = 0; y < sample_grid_size; ++y)
				for (uint32_t z = 0; z < sample_grid_size; ++z)
					shifts[x + y * sample_grid_size + z * sample_grid_size * sample_grid_size] = grid_start + Vec3f{(float)x * step.x, (float)y * step.y, (float)z * step.z};
	}
	const float min_light_dist = std::fmin(std::fmin(h.grid_cell_size[0], h.grid_cell_size[1]), h.grid_cell_size[2]) * 0.5f;
	const float min_light_square_dist = min_light_dist * min_light_dist;
	const size_t total = (size_t)h.grid_size[0] * h.grid_size[1] * h.grid_size[2];
	std::vector<LightGridElement> grid(total); // address = z + (x + y * size_x) * size_z (:1754-1758)
	std::atomic<size_t> next{0};
	const unsigned nt = std::min(std::max(std::thread::hardware_concurrency(), 1u), 64u);
	std::vector<std::thread> threads;
	for (unsigned t = 0; t < nt; ++t)
		threads.emplace_back([&] {
			for (;;)
			{
				const size_t first = next.fetch_add(64);
				if (first >= total)
					break;
				for (size_t address = first; address < std::min(first + 64, total); ++address)
				{
					const uint32_t z = (uint32_t)(address % h.grid_size[2]), layer = (uint32_t)(address / h.grid_size[2]);
					const uint32_t x = layer % h.grid_size[0], y = layer / h.grid_size[0];
					const Vec3f pos{h.grid_start[0] + (float)x * h.grid_cell_size[0], h.grid_start[1] + (float)y * h.grid_cell_size[1], h.grid_start[2] + (float)z * h.grid_cell_size[2]};
					LightCube cube;
					uint32_t valid = 0;
					for (const Vec3f& shift : shifts)
					{
						Vec3f p;
						if (!correct_light_grid_sample_position(map, pos + shift, p))
							continue;
						++valid;
						for (const PointLight& light : lights) // calculate_light_for_grid_point :1861-2003 (primary and sun lights)
						{
							const Vec3f vec_to_light = light.pos - p;
							const float len2 = std::fmax(magnitude2(vec_to_light), min_light_square_dist), len = std::sqrt(len2);
							if (std::fmax(std::fmax(light.color[0], light.color[1]), light.color[2]) / len2 < settings.light_cutoff)
								continue; // (stands in for the PVS row of the point's leaf)
							if (light.has_direction && -dot(vec_to_light, light.direction.dir_normalized) < light.direction.half_angle_cos * len)
								continue;
							const float shadow_factor = tracer.get_shadow_factor(light.pos, p);
							if (shadow_factor <= 0.0f)
								continue;
							const float k = shadow_factor / len2;
							const float c[3] = {light.color[0] * k, light.color[1] * k, light.color[2] * k};
							cube.add_light_sample(vec_to_light, c);
						}
						for (const SunLight& sun : sun_lights)
						{
							const float shadow_factor = tracer.get_sun_shadow_factor(p, sun.dir);
							if (shadow_factor <= 0.0f)
								continue;
							const float c[3] = {sun.color[0] * shadow_factor, sun.color[1] * shadow_factor, sun.color[2] * shadow_factor};
							cube.add_light_sample(sun.dir, c);
						}
					}
					if (valid > 0)
					{
						cube.scale(1.0f / (float)valid);
						if (settings.ambient_light > 0.0f) // LightCube::add_constant_light with the tool's ambient term (as in the lightmaps)
							for (auto& side : cube.light_cube)
								for (float& c : side)
									c += settings.ambient_light;
					}
					grid[address] = cube.convert_into_light_grid_sample();
				}
			}
		});
	for (std::thread& th : threads)
		th.join();
	// compress_light_grid :1772-1834
	map.light_grid_columns.assign((size_t)h.grid_size[0] * h.grid_size[1], LightGridColumn{});
	map.light_grid_samples.clear();
	auto sample = [&](uint32_t x, uint32_t y, uint32_t z) -> const LightGridElement& { return grid[z + ((size_t)x + (size_t)y * h.grid_size[0]) * h.grid_size[2]]; };
	for (uint32_t x = 0; x < h.grid_size[0]; ++x)
		for (uint32_t y = 0; y < h.grid_size[1]; ++y)
		{
			uint32_t lo = 0;
			while (lo < h.grid_size[2] && light_grid_element_is_zero(sample(x, y, lo)))
				++lo;
			uint32_t hi = h.grid_size[2] - 1;
			while (hi > lo && light_grid_element_is_zero(sample(x, y, hi)))
				--hi;
			LightGridColumn& col = map.light_grid_columns[x + (size_t)y * h.grid_size[0]];
			col.first_sample = (uint32_t)map.light_grid_samples.size();
			col.num_samples = hi + 1 - lo; // (an all-dark column: lo == size_z, hi == size_z - 1: no samples)
			col.start_z = lo;
			for (uint32_t z = lo; z <= hi && z < h.grid_size[2]; ++z)
				map.light_grid_samples.push_back(sample(x, y, z));
		}
	return (uint32_t)map.light_grid_samples.size();
}
} // namespace

std::vector<uint8_t> calculate_leafs_visibility(const BSPMap& map)
{
	const Pvs pvs = calculate_visibility_matrix(map);
	std::vector<uint8_t> out(pvs.num_leafs * pvs.num_leafs);
	for (size_t y = 0; y < pvs.num_leafs; ++y)
		std::copy(pvs.rows[y].begin(), pvs.rows[y].end(), out.begin() + y * pvs.num_leafs);
	return out;
}

void build_lightmaps(const LightmappingSettings& settings_in, const MaterialTable& materials, BSPMap& map, LightmapStats& stats)
{
	LightmappingSettings settings = settings_in;
	stats = LightmapStats{};
	Vec3f bbox_min{1e24f, 1e24f, 1e24f}, bbox_max{-1e24f, -1e24f, -1e24f}; // bsp_map_compact.rs:219-234
	for (const Vec3f& v : map.vertices)
		for (int a = 0; a < 3; ++a)
		{
			bbox_min.at(a) = std::fmin(bbox_min[a], v[a]);
			bbox_max.at(a) = std::fmax(bbox_max[a], v[a]);
		}
	const uint32_t sample_grid_size = std::max(std::min(settings.sample_grid_size, MAX_SAMPLE_GRID_SIZE), 1u);
	std::vector<PointLight> lights = extract_map_lights(map);
	std::vector<SunLight> sun_lights = extract_sun_lights(map, bbox_min, bbox_max);
	for (PointLight& l : lights)
		for (float& c : l.color)
			c *= settings.light_scale;
	for (SunLight& l : sun_lights)
		for (float& c : l.color)
			c *= settings.light_scale;
	stats.point_lights = (uint32_t)lights.size();
	stats.sun_lights = (uint32_t)sun_lights.size();

	const size_t num_texels = allocate_lightmaps(materials, map);
	stats.lightmap_texels = (uint32_t)num_texels;
	Tracer tracer{map, {}, {}};
	for (const TextureName& t : map.textures) // light_trace.rs:15-40, :52-62
	{
		const MaterialDef* m = material_of_texture(materials, t);
		float op = 0.0f;
		if (m)
		{
			if (m->shadow)
				op = (m->blending_mode == SQW_BLEND_NONE) ? 0.0f : (m->blending_mode == SQW_BLEND_ADDITIVE ? 1.0f : 0.5f);
			else
				op = 1.0f;
		}
		tracer.opacity.push_back(op);
		tracer.sky.push_back(m && m->skybox);
	}

	map.lightmaps_data.assign(num_texels * 3, 0.0f);
	sqw_directional_lightmap_element zero;
	std::memset(&zero, 0, sizeof(zero));
	if (settings.build_directional_lightmap)
		map.directional_lightmaps_data.assign(num_texels, zero);
	else
		map.directional_lightmaps_data.clear();
	for (int i = 0; i < 3; ++i)
		map.lightmaps_data[i] = settings.ambient_light; // element 0 (build_lightmaps fills every element with the ambient term first)

	PolygonLighter lighter{map, tracer, lights, sun_lights, settings, sample_grid_size};
	// With the PVS (the reference's way): a polygon is lit by the lights standing in leafs its own leaf can see
	// (group_lights_by_leafs :229-240, get_visible_leafs :683-700, get_visible_leafs_for_submodel :702-739).
	std::vector<std::vector<const PointLight*>> polygon_lights; // per polygon, only with use_pvs
	if (settings.use_pvs)
	{
		const Pvs pvs = calculate_visibility_matrix(map);
		std::vector<std::vector<const PointLight*>> by_leaf(map.leafs.size());
		for (const PointLight& l : lights)
			by_leaf[leaf_for_point(map, l.pos)].push_back(&l);
		auto lights_of_row = [&](const std::vector<uint8_t>& row) {
			std::vector<const PointLight*> out;
			for (size_t other = 0; other < row.size(); ++other)
				if (row[other])
					out.insert(out.end(), by_leaf[other].begin(), by_leaf[other].end());
			return out;
		};
		polygon_lights.resize(map.polygons.size());
		for (size_t li = 0; li < map.leafs.size(); ++li)
		{
			const std::vector<const PointLight*> ll = lights_of_row(pvs.rows[li]);
			for (uint32_t k = 0; k < map.leafs[li].num_polygons; ++k)
				polygon_lights[map.leafs[li].first_polygon + k] = ll;
		}
		for (const Submodel& sm : map.submodels)
		{
			Vec3f lo{1e24f, 1e24f, 1e24f}, hi{-1e24f, -1e24f, -1e24f}; // get_submodel_bbox (bsp_map_compact.rs:236-258)
			for (uint32_t k = 0; k < sm.num_polygons; ++k)
			{
				const Polygon& p = map.polygons[sm.first_polygon + k];
				for (uint32_t v = 0; v < p.num_vertices; ++v)
					for (int a = 0; a < 3; ++a)
					{
						lo.at(a) = std::fmin(lo[a], map.vertices[p.first_vertex + v][a]);
						hi.at(a) = std::fmax(hi[a], map.vertices[p.first_vertex + v][a]);
					}
			}
			const Vec3f corners[8] = {{lo.x, lo.y, lo.z}, {lo.x, lo.y, hi.z}, {lo.x, hi.y, lo.z}, {lo.x, hi.y, hi.z}, {hi.x, lo.y, lo.z}, {hi.x, lo.y, hi.z}, {hi.x, hi.y, lo.z}, {hi.x, hi.y, hi.z}};
			std::vector<uint32_t> leafs;
			collect_submodel_leafs_r(map, corners, map.root_node(), leafs);
			std::vector<uint8_t> row(map.leafs.size(), 0);
			for (uint32_t l : leafs)
				for (size_t other = 0; other < row.size(); ++other)
					row[other] |= pvs.rows[l][other];
			const std::vector<const PointLight*> ll = lights_of_row(row);
			for (uint32_t k = 0; k < sm.num_polygons; ++k)
				polygon_lights[sm.first_polygon + k] = ll;
		}
	}
	std::atomic<size_t> next{0};
	const unsigned nt = std::min(std::max(std::thread::hardware_concurrency(), 1u), 64u);
	std::vector<std::thread> threads;
	float* simple = map.lightmaps_data.data();
	sqw_directional_lightmap_element* directional = settings.build_directional_lightmap ? map.directional_lightmaps_data.data() : nullptr;
	for (unsigned t = 0; t < nt; ++t)
		threads.emplace_back([&] {
			for (;;)
			{
				const size_t i = next.fetch_add(16);
				if (i >= map.polygons.size())
					break;
				for (size_t k = i; k < std::min(i + 16, map.polygons.size()); ++k)
					if (map.polygons[k].lightmap_data_offset != 0)
						lighter.light_polygon(map.polygons[k], simple, directional, settings.use_pvs ? &polygon_lights[k] : nullptr); // each polygon owns its lightmap range
			}
		});
	for (std::thread& th : threads)
		th.join();

	stats.light_grid_samples = build_light_grid(settings, sample_grid_size, tracer, lights, sun_lights, bbox_min, bbox_max, map);
}

} // namespace mapc
} // namespace sqwh
