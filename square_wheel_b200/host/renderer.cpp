// Host-side mirror of the renderer front end (see renderer.hpp).  Each function cites the reference code it follows.
#include "renderer.hpp"
#include <algorithm>
#include <cstring>

namespace sqwh
{

constexpr size_t MAX_VERTICES = 24; // renderer_utils.rs:4
constexpr float Z_NEAR = 1.0f;      // renderer_utils.rs:7

// ---- common/clipping_polygon.rs ------------------------------------------------------------------------------------
namespace
{
inline bool axis_valid(const ClipAxis& a) { return a.min < a.max; }
inline bool axis_contains(const ClipAxis& a, const ClipAxis& o) { return o.min >= a.min && o.max <= a.max; }
inline void axis_extend(ClipAxis& a, const ClipAxis& o)
{
	if (o.min < a.min)
		a.min = o.min;
	if (o.max > a.max)
		a.max = o.max;
}
inline void axis_extend_point(ClipAxis& a, float p)
{
	if (p < a.min)
		a.min = p;
	if (p > a.max)
		a.max = p;
}
inline void axis_intersect(ClipAxis& a, const ClipAxis& o)
{
	if (o.min > a.min)
		a.min = o.min;
	if (o.max < a.max)
		a.max = o.max;
}
} // namespace

ClippingPolygon ClippingPolygon::from_box(float min_x, float min_y, float max_x, float max_y)
{
	ClippingPolygon c;
	c.x = {min_x, max_x};
	c.y = {min_y, max_y};
	c.x_plus_y = {min_x + min_y, max_x + max_y};
	c.x_minus_y = {min_x - max_y, max_x - min_y};
	return c;
}
ClippingPolygon ClippingPolygon::from_point(Vec2f p)
{
	ClippingPolygon c;
	c.x = {p.x, p.x};
	c.y = {p.y, p.y};
	c.x_plus_y = {p.x + p.y, p.x + p.y};
	c.x_minus_y = {p.x - p.y, p.x - p.y};
	return c;
}
void ClippingPolygon::get_box_clip_planes(Vec3f out[4]) const // clipping_polygon.rs:154-162
{
	out[0] = {-1.0f, 0.0f, -x.max};
	out[1] = {1.0f, 0.0f, x.min};
	out[2] = {0.0f, -1.0f, -y.max};
	out[3] = {0.0f, 1.0f, y.min};
}
bool ClippingPolygon::is_valid_and_non_empty() const { return axis_valid(x) && axis_valid(y) && axis_valid(x_plus_y) && axis_valid(x_minus_y); }
bool ClippingPolygon::contains(const ClippingPolygon& o) const
{
	return axis_contains(x, o.x) && axis_contains(y, o.y) && axis_contains(x_plus_y, o.x_plus_y) && axis_contains(x_minus_y, o.x_minus_y);
}
void ClippingPolygon::extend(const ClippingPolygon& o)
{
	axis_extend(x, o.x);
	axis_extend(y, o.y);
	axis_extend(x_plus_y, o.x_plus_y);
	axis_extend(x_minus_y, o.x_minus_y);
}
void ClippingPolygon::extend_with_point(Vec2f p)
{
	axis_extend_point(x, p.x);
	axis_extend_point(y, p.y);
	axis_extend_point(x_plus_y, p.x + p.y);
	axis_extend_point(x_minus_y, p.x - p.y);
}
void ClippingPolygon::intersect(const ClippingPolygon& o)
{
	axis_intersect(x, o.x);
	axis_intersect(y, o.y);
	axis_intersect(x_plus_y, o.x_plus_y);
	axis_intersect(x_minus_y, o.x_minus_y);
}
void ClippingPolygon::get_clip_planes(Vec3f out[8]) const // clipping_polygon.rs:150-163
{
	out[0] = {-1.0f, 0.0f, -x.max};
	out[1] = {1.0f, 0.0f, x.min};
	out[2] = {0.0f, -1.0f, -y.max};
	out[3] = {0.0f, 1.0f, y.min};
	out[4] = {-1.0f, -1.0f, -x_plus_y.max};
	out[5] = {1.0f, 1.0f, x_plus_y.min};
	out[6] = {-1.0f, 1.0f, -x_minus_y.max};
	out[7] = {1.0f, -1.0f, x_minus_y.min};
}

// ---- common/clipping.rs ------------------------------------------------------------------------------------------
namespace
{

Vec3f get_line_z_intersection(Vec3f v0, Vec3f v1, float z) // clipping.rs:122-130
{
	const float dist0 = v0.z - z;
	const float dist1 = v1.z - z;
	const float dist_sum = v1.z - v0.z;
	const float k0 = dist0 / dist_sum;
	const float k1 = dist1 / dist_sum;
	return v0 * k1 - v1 * k0;
}

size_t clip_3d_polygon_by_z_plane(const Vec3f* polygon, size_t n, float z_dist, Vec3f* out, size_t out_cap) // clipping.rs:4-57
{
	const Vec3f* prev_v = &polygon[n - 1];
	size_t count = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const Vec3f* v = &polygon[i];
		if (v->z > z_dist)
		{
			if (prev_v->z < z_dist)
			{
				out[count++] = get_line_z_intersection(*prev_v, *v, z_dist);
				if (count == out_cap)
					break;
			}
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (v->z == z_dist)
		{
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (prev_v->z > z_dist)
		{
			out[count++] = get_line_z_intersection(*prev_v, *v, z_dist);
			if (count == out_cap)
				break;
		}
		prev_v = v;
	}
	return count;
}

Vec2f get_line_line_intersection(Vec2f v0, Vec2f v1, Vec3f line) // clipping.rs:187-195
{
	const Vec2f l2{line.x, line.y};
	const float dist0 = dot(v0, l2) - line.z;
	const float dist1 = dot(v1, l2) - line.z;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum;
	const float k1 = dist1 / dist_sum;
	return v0 * k1 - v1 * k0;
}

size_t clip_2d_polygon(const Vec2f* polygon, size_t n, Vec3f clip_line_equation, Vec2f* out, size_t out_cap) // clipping.rs:132-185
{
	const Vec2f l2{clip_line_equation.x, clip_line_equation.y};
	const Vec2f* prev_v = &polygon[n - 1];
	float prev_dist = dot(*prev_v, l2) - clip_line_equation.z;
	size_t count = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const Vec2f* v = &polygon[i];
		const float dist = dot(*v, l2) - clip_line_equation.z;
		if (dist > 0.0f)
		{
			if (prev_dist < 0.0f)
			{
				out[count++] = get_line_line_intersection(*prev_v, *v, clip_line_equation);
				if (count == out_cap)
					break;
			}
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (dist == 0.0f)
		{
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (prev_dist > 0.0f)
		{
			out[count++] = get_line_line_intersection(*prev_v, *v, clip_line_equation);
			if (count == out_cap)
				break;
		}
		prev_v = v;
		prev_dist = dist;
	}
	return count;
}

// renderer_utils.rs:10-58
size_t project_and_clip_polygon(const Vec3f clip_planes[8], const Vec3f* vertices_transformed, size_t num_vertices, Vec2f* out_vertices)
{
	size_t vertex_count = std::min(num_vertices, MAX_VERTICES);
	Vec3f z_clipped[MAX_VERTICES];
	vertex_count = clip_3d_polygon_by_z_plane(vertices_transformed, vertex_count, Z_NEAR, z_clipped, MAX_VERTICES);
	if (vertex_count < 3)
		return vertex_count;
	for (size_t i = 0; i < vertex_count; ++i)
		out_vertices[i] = truncate(z_clipped[i]) / z_clipped[i].z;
	Vec2f temp[MAX_VERTICES];
	for (size_t i = 0; i < 4; ++i)
	{
		vertex_count = clip_2d_polygon(out_vertices, vertex_count, clip_planes[i * 2], temp, MAX_VERTICES);
		if (vertex_count < 3)
			return vertex_count;
		vertex_count = clip_2d_polygon(temp, vertex_count, clip_planes[i * 2 + 1], out_vertices, MAX_VERTICES);
		if (vertex_count < 3)
			return vertex_count;
	}
	return vertex_count;
}

// equations.rs
inline float sample_point(const float de[3], Vec2f p) { return de[0] * p.x + de[1] * p.y + de[2]; }

// renderer_utils.rs:201-242
uint32_t calculate_mip(const Vec2f* points, size_t n, const float de[3], const float te[6], float mip_bias)
{
	Vec2f mip_point = points[0];
	float mip_point_inv_z = 0.0f;
	for (size_t i = 0; i < n; ++i)
	{
		const float inv_z = sample_point(de, points[i]);
		if (inv_z > mip_point_inv_z)
		{
			mip_point_inv_z = inv_z;
			mip_point = points[i];
		}
	}
	const float z_2 = 1.0f / (mip_point_inv_z * mip_point_inv_z);
	const float z_4 = z_2 * z_2;
	float d_tc_2[2];
	for (int i = 0; i < 2; ++i)
	{
		// te layout: d_tc_dx[2], d_tc_dy[2], k[2]
		const float d_tc_dx = te[i] * (de[2] + de[1] * mip_point.y) - (te[4 + i] + te[2 + i] * mip_point.y) * de[0];
		const float d_tc_dy = te[2 + i] * (de[2] + de[0] * mip_point.x) - (te[4 + i] + te[i] * mip_point.x) * de[1];
		d_tc_2[i] = (d_tc_dx * d_tc_dx + d_tc_dy * d_tc_dy) * z_4;
	}
	const float max_d_tc_2 = std::fmax(d_tc_2[0], d_tc_2[1]);
	const float mip_f = std::log2(max_d_tc_2) * 0.5f + mip_bias;
	float c = std::ceil(mip_f);
	int32_t ci = (c != c) ? 0 : (c >= 2147483648.0f ? INT32_MAX : (c <= -2147483648.0f ? INT32_MIN : (int32_t)c)); // `as i32`
	return (uint32_t)std::max(0, std::min(ci, (int32_t)MAX_MIP));
}

constexpr float TC_ERROR_THRESHOLD = 0.75f; // renderer_utils.rs:199

// renderer_utils.rs:60-132
bool affine_texture_coordinates_interpolation_may_be_used(const float de[3], const float te[6], Vec2f min_inv_z_point, Vec2f max_inv_z_point)
{
	const Vec2f edge = max_inv_z_point - min_inv_z_point;
	const float edge_square_len = dot(edge, edge);
	if (edge_square_len == 0.0f)
		return true;
	const float edge_len = std::sqrt(edge_square_len);
	const Vec2f edge_vec_normalized = edge / edge_len;
	const float inv_z_clamp = 1.0f / (float)(1 << 20);
	const float min_point_inv_z = std::fmax(sample_point(de, min_inv_z_point), inv_z_clamp);
	const float max_point_inv_z = std::fmax(sample_point(de, max_inv_z_point), inv_z_clamp);
	const float depth_equation_projected_a = dot(Vec2f{de[0], de[1]}, edge_vec_normalized);
	const float depth_equation_projected_b = min_point_inv_z;
	if (std::fabs(depth_equation_projected_a) < 1.0e-10f)
		return true;
	const float depth_b_div_a = depth_equation_projected_b / depth_equation_projected_a;
	const float max_diff_point = std::sqrt((0.0f + depth_b_div_a) * (edge_len + depth_b_div_a)) - depth_b_div_a;
	const float max_diff_point_inv_z = depth_equation_projected_a * max_diff_point + depth_equation_projected_b;
	for (int i = 0; i < 2; ++i)
	{
		const float min_point_tc = te[i] * min_inv_z_point.x + te[2 + i] * min_inv_z_point.y + te[4 + i];
		const float max_point_tc = te[i] * max_inv_z_point.x + te[2 + i] * max_inv_z_point.y + te[4 + i];
		const float tc_projected_a = dot(Vec2f{te[i], te[2 + i]}, edge_vec_normalized);
		const float tc_projected_b = min_point_tc;
		const float min_point_tc_z_mul = min_point_tc / min_point_inv_z;
		const float max_point_tc_z_mul = max_point_tc / max_point_inv_z;
		const float max_diff_point_tc_real = (tc_projected_a * max_diff_point + tc_projected_b) / max_diff_point_inv_z;
		const float max_diff_point_tc_approximate = min_point_tc_z_mul + (max_point_tc_z_mul - min_point_tc_z_mul) * (max_diff_point - 0.0f) / (edge_len - 0.0f);
		const float tc_abs_diff = std::fabs(max_diff_point_tc_real - max_diff_point_tc_approximate);
		if (tc_abs_diff > TC_ERROR_THRESHOLD)
			return false;
	}
	return true;
}

// renderer_utils.rs:134-197
bool line_z_corrected_texture_coordinates_interpolation_may_be_used(const float de[3], const float te[6], Vec2f max_inv_z_point, float min_polygon_x, float max_polygon_x)
{
	if (max_polygon_x - min_polygon_x < 1.0f)
		return true;
	const float a = de[0];
	const float b = de[1] * max_inv_z_point.y + de[2];
	if (std::fabs(a) < 1.0e-10f)
		return true;
	const float depth_b_div_a = b / a;
	const float max_diff_x = std::sqrt((min_polygon_x + depth_b_div_a) * (max_polygon_x + depth_b_div_a)) - depth_b_div_a;
	const float max_diff_point_inv_z = a * max_diff_x + b;
	const float inv_z_at_min_x = a * min_polygon_x + b;
	const float inv_z_at_max_x = a * max_polygon_x + b;
	const float almost_zero = 1e-20f;
	if (inv_z_at_min_x <= almost_zero || inv_z_at_max_x <= almost_zero || max_diff_point_inv_z <= almost_zero)
		return false;
	for (int i = 0; i < 2; ++i)
	{
		const float ta = te[i];
		const float tb = te[2 + i] * max_inv_z_point.y + te[4 + i];
		const float tc_at_min_x = (ta * min_polygon_x + tb) / inv_z_at_min_x;
		const float tc_at_max_x = (ta * max_polygon_x + tb) / inv_z_at_max_x;
		const float max_diff_point_tc_real = (ta * max_diff_x + tb) / max_diff_point_inv_z;
		const float max_diff_point_tc_approximate = tc_at_min_x + (tc_at_max_x - tc_at_min_x) * (max_diff_x - min_polygon_x) / (max_polygon_x - min_polygon_x);
		const float tc_abs_diff = std::fabs(max_diff_point_tc_real - max_diff_point_tc_approximate);
		if (tc_abs_diff > TC_ERROR_THRESHOLD)
			return false;
	}
	return true;
}

inline int32_t f32_to_fixed16_unchecked(float x) // common/fixed_math.rs:36 (cvttss2si semantics)
{
	const float v = x * 65536.0f;
	if (!(v > -2147483904.0f && v < 2147483648.0f))
		return INT32_MIN;
	return (int32_t)v;
}

// surfaces.rs:18-47 PolygonBasisVecs::form_plane_and_tex_coord_equation
void basis_vecs_from_plane_and_tc(const Plane& plane, const Plane tc[2], DrawPolygonData& out)
{
	const Mat4f tex_coord_basis = Mat4f::from_cols(extend(tc[0].vec, tc[0].dist), extend(tc[1].vec, tc[1].dist), extend(plane.vec, -plane.dist), Vec4f{0, 0, 0, 1});
	Mat4f inv;
	if (!tex_coord_basis.transpose().invert(inv))
	{
		out.basis_u = {1, 0, 0};
		out.basis_v = {0, 1, 0};
		out.basis_start = {0, 0, 0};
		out.basis_normal = {0, 0, 1};
		return;
	}
	out.basis_u = truncate(inv.x);
	out.basis_v = truncate(inv.y);
	out.basis_start = truncate(inv.w);
	const float min_positive_value = 1.0f / (float)(1 << 30);
	out.basis_normal = plane.vec / std::fmax(magnitude(plane.vec), min_positive_value);
}

} // namespace

// ---- Renderer ----------------------------------------------------------------------------------------------------------
Renderer::Renderer(const Scene& scene, const RendererConfig& config)
	: scene_(scene), map_(scene.map), config_(config), visibility_(scene.map), shadow_visibility_(scene.map)
{
	// PartialRenderer::new (partial_renderer.rs:56-135)
	polygons_data_.resize(map_.polygons.size());
	for (size_t i = 0; i < map_.polygons.size(); ++i)
		basis_vecs_from_plane_and_tc(map_.polygons[i].plane, map_.polygons[i].tex_coord_equation, polygons_data_[i]);
	for (size_t leaf = 0; leaf < map_.leafs.size(); ++leaf)
		for (uint32_t p = map_.leafs[leaf].first_polygon; p < map_.leafs[leaf].first_polygon + map_.leafs[leaf].num_polygons; ++p)
			polygons_data_[p].parent_leaf = (uint32_t)leaf;
	vertices_transformed_.resize(map_.vertices.size());
	leaf_lights_.resize(map_.leafs.size());
	leaf_models_.resize(map_.leafs.size());
	// get_leaf_clip_planes (renderer_structs.rs:167-226): portal and polygon planes of the leaf, normalised and deduplicated
	leafs_planes_.resize(map_.leafs.size());
	for (size_t leaf_index = 0; leaf_index < map_.leafs.size(); ++leaf_index)
	{
		const BSPLeaf& leaf = map_.leafs[leaf_index];
		std::vector<Plane>& planes = leafs_planes_[leaf_index];
		auto add_and_deduplicate_plane = [&](const Plane& plane) {
			const float normal_length = magnitude(plane.vec);
			if (normal_length < 0.00000000001f)
				return;
			const Plane plane_normalized = {plane.vec / normal_length, plane.dist / normal_length};
			for (Plane& prev_plane : planes)
				if (dot(plane_normalized.vec, prev_plane.vec) >= 1.0f - 1.0f / 256.0f)
				{
					prev_plane.dist = std::fmax(prev_plane.dist, plane_normalized.dist);
					return;
				}
			planes.push_back(plane_normalized);
		};
		for (uint32_t i = leaf.first_leaf_portal; i < leaf.first_leaf_portal + leaf.num_leaf_portals; ++i)
		{
			const Portal& portal = map_.portals[map_.leafs_portals[i]];
			add_and_deduplicate_plane(portal.leafs[0] == leaf_index ? portal.plane : Plane{-portal.plane.vec, -portal.plane.dist});
		}
		for (uint32_t p = leaf.first_polygon; p < leaf.first_polygon + leaf.num_polygons; ++p)
			add_and_deduplicate_plane(map_.polygons[p].plane);
	}
}

VisibilityCalculator::VisibilityCalculator(const BSPMap& map) : map_(map)
{
	leafs_data_.resize(map_.leafs.size());
	portals_data_.resize(map_.portals.size());
}

bool VisibilityCalculator::leaf_bounds(uint32_t leaf, ClippingPolygon& out) const // map_visibility_calculator.rs:116-127
{
	const LeafData& d = leafs_data_[leaf];
	if (d.visible_frame != visibility_frame_)
		return false;
	out = d.current_frame_bounds;
	return true;
}

uint32_t VisibilityCalculator::find_current_leaf() const // :134-156
{
	uint32_t index = map_.root_node();
	const Vec4f w_row = camera_matrices_.planes_matrix.row(3);
	for (;;)
	{
		if (index >= FIRST_LEAF_INDEX)
			return index - FIRST_LEAF_INDEX;
		const BSPNode& node = map_.nodes[index];
		const float plane_transformed_w = dot(w_row, extend(node.plane.vec, -node.plane.dist));
		index = plane_transformed_w >= 0.0f ? node.children[0] : node.children[1];
	}
}

namespace
{
// map_visibility_calculator.rs:282-326 project_portal
bool project_portal(const Portal& portal, const BSPMap& map, const Mat4f& view_matrix, ClippingPolygon& out)
{
	size_t vertex_count = std::min((size_t)portal.num_vertices, MAX_VERTICES);
	Vec3f vt[MAX_VERTICES];
	for (size_t i = 0; i < vertex_count; ++i)
		vt[i] = view_matrix_transform_vertex(view_matrix, map.vertices[portal.first_vertex + i]);
	Vec3f zc[MAX_VERTICES];
	const float PORTAL_Z_NEAR = 1.0f / 4096.0f;
	vertex_count = clip_3d_polygon_by_z_plane(vt, vertex_count, PORTAL_Z_NEAR, zc, MAX_VERTICES);
	if (vertex_count < 3)
		return false;
	out = ClippingPolygon::from_point(truncate(zc[0]) / zc[0].z);
	for (size_t i = 1; i < vertex_count; ++i)
		out.extend_with_point(truncate(zc[i]) / zc[i].z);
	return true;
}
} // namespace

void VisibilityCalculator::mark_reachable_leafs_iterative(const uint32_t* start_leafs, size_t num_start_leafs, const ClippingPolygon& start_bounds) // :158-264
{
	wave_cur_.clear();
	wave_next_.clear();
	for (size_t i = 0; i < num_start_leafs; ++i)
	{
		wave_cur_.push_back(start_leafs[i]);
		leafs_data_[start_leafs[i]].current_frame_bounds = start_bounds;
		leafs_data_[start_leafs[i]].visible_frame = visibility_frame_;
	}
	uint32_t depth = 0;
	while (!wave_cur_.empty())
	{
		for (uint32_t leaf : wave_cur_)
		{
			const ClippingPolygon leaf_bounds_v = leafs_data_[leaf].current_frame_bounds;
			const BSPLeaf& leaf_value = map_.leafs[leaf];
			for (uint32_t k = leaf_value.first_leaf_portal; k < leaf_value.first_leaf_portal + leaf_value.num_leaf_portals; ++k)
			{
				const uint32_t portal = map_.leafs_portals[k];
				const Portal& portal_value = map_.portals[portal];
				const float portal_plane_pos = (camera_matrices_.planes_matrix * extend(portal_value.plane.vec, -portal_value.plane.dist)).w;
				uint32_t next_leaf;
				if (portal_value.leafs[0] == leaf)
				{
					if (portal_plane_pos <= 0.0f)
						continue;
					next_leaf = portal_value.leafs[1];
				}
				else
				{
					if (portal_plane_pos >= 0.0f)
						continue;
					next_leaf = portal_value.leafs[0];
				}
				PortalData& portal_data = portals_data_[portal];
				if (portal_data.visible_frame != visibility_frame_)
				{
					portal_data.visible_frame = visibility_frame_;
					portal_data.has_projection = project_portal(portal_value, map_, camera_matrices_.view_matrix, portal_data.current_frame_projection);
				}
				if (!portal_data.has_projection)
					continue;
				ClippingPolygon bounds_intersection = portal_data.current_frame_projection;
				bounds_intersection.intersect(leaf_bounds_v);
				if (!bounds_intersection.is_valid_and_non_empty())
					continue;
				LeafData& next_leaf_data = leafs_data_[next_leaf];
				if (next_leaf_data.visible_frame != visibility_frame_)
				{
					next_leaf_data.visible_frame = visibility_frame_;
					next_leaf_data.current_frame_bounds = bounds_intersection;
				}
				else
				{
					if (next_leaf_data.current_frame_bounds.contains(bounds_intersection))
						continue;
					next_leaf_data.current_frame_bounds.extend(bounds_intersection);
				}
				wave_next_.push_back(next_leaf);
			}
		}
		wave_cur_.clear();
		std::swap(wave_cur_, wave_next_);
		if (++depth > 1024)
			break;
	}
}

bool VisibilityCalculator::compute_is_inside_leaf_volume(uint32_t leaf_index) const // :266-280
{
	const BSPLeaf& leaf = map_.leafs[leaf_index];
	for (uint32_t p = leaf.first_polygon; p < leaf.first_polygon + leaf.num_polygons; ++p)
	{
		const Polygon& polygon = map_.polygons[p];
		if ((camera_matrices_.planes_matrix * extend(polygon.plane.vec, -polygon.plane.dist)).w < 0.0f)
			return false;
	}
	return true;
}

void VisibilityCalculator::update_visibility(const CameraMatrices& camera_matrices, const ClippingPolygon& frame_bounds) // :52-100
{
	camera_matrices_ = camera_matrices;
	++visibility_frame_;
	const uint32_t current_leaf = find_current_leaf();
	constexpr size_t MAX_START_LEAFS = 32;
	uint32_t start_leafs[MAX_START_LEAFS];
	start_leafs[0] = current_leaf;
	size_t num_start_leafs = 1;
	const BSPLeaf& leaf_value = map_.leafs[current_leaf];
	for (uint32_t k = leaf_value.first_leaf_portal; k < leaf_value.first_leaf_portal + leaf_value.num_leaf_portals; ++k)
	{
		const Portal& portal_value = map_.portals[map_.leafs_portals[k]];
		const float scaled_dist = dot(portal_value.plane.vec, camera_matrices_.position) - portal_value.plane.dist;
		const float eps = Z_NEAR * 2.0f;
		if (std::fabs(scaled_dist) <= eps * magnitude(portal_value.plane.vec))
		{
			const uint32_t next_leaf = portal_value.leafs[0] == current_leaf ? portal_value.leafs[1] : portal_value.leafs[0];
			start_leafs[num_start_leafs++] = next_leaf;
			if (num_start_leafs == MAX_START_LEAFS)
				break;
		}
	}
	mark_reachable_leafs_iterative(start_leafs, num_start_leafs, frame_bounds);
	is_inside_leaf_volume_ = compute_is_inside_leaf_volume(current_leaf);
}

// ---- dynamic lights (dynamic_objects_index.rs, partial_renderer.rs:398-590, depth_renderer.rs) ------------------------------
namespace
{
inline Vec3f quaternion_rotate_vector(Quaternion q, Vec3f v) // cgmath Quaternion::rotate_vector
{
	const Vec3f t = cross(q.v, v) * 2.0f;
	return v + t * q.s + cross(q.v, t);
}
inline Mat4f get_object_matrix(Vec3f position, Quaternion rotation) // common/matrix.rs:112-118
{
	return Mat4f::from_translation(position) * mat4_from_quaternion(rotation);
}
// light.rs:86-130 get_cube_map_side_matrix
Mat4f get_cube_map_side_matrix(uint32_t side)
{
	Mat4f mat = Mat4f::identity();
	const Vec4f ux{1, 0, 0, 0}, uy{0, 1, 0, 0}, uz{0, 0, 1, 0};
	auto neg = [](Vec4f v) { return Vec4f{-v.x, -v.y, -v.z, -v.w}; };
	switch (side)
	{
	case 0: mat.x = uz; mat.y = neg(ux); mat.z = neg(uy); break;      // XPlus
	case 1: mat.x = neg(uz); mat.y = ux; mat.z = neg(uy); break;      // XMinus
	case 2: mat.x = ux; mat.y = uz; mat.z = neg(uy); break;           // YPlus
	case 3: mat.x = neg(ux); mat.y = neg(uz); mat.z = neg(uy); break; // YMinus
	case 4: mat.x = ux; mat.y = uy; mat.z = uz; break;                // ZPlus
	default: mat.x = ux; mat.y = neg(uy); mat.z = neg(uz); break;     // ZMinus
	}
	return mat;
}
// common/matrix.rs:49-78 build_view_matrix_with_full_rotation
CameraMatrices build_view_matrix_with_full_rotation(Vec3f position, Quaternion rotation, float fov, float viewport_width, float viewport_height)
{
	const float inv_half_fov_tan = 1.0f / std::tan(fov * 0.5f);
	const float aspect = viewport_width / viewport_height;
	Mat4f basis_change = Mat4f::identity();
	basis_change.x = {0, 0, 1, 0};
	basis_change.y = {-1, 0, 0, 0};
	basis_change.z = {0, -1, 0, 0};
	return complete_view_matrix(position, basis_change * mat4_from_quaternion(rotation.conjugate()), Vec2f{inv_half_fov_tan / aspect, inv_half_fov_tan}, viewport_width, viewport_height);
}
} // namespace

void Renderer::position_object_sphere_r(uint32_t id, Vec3f center, float radius, uint32_t node_index) // dynamic_objects_index.rs:211-243
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_lights_[leaf_index].push_back(id);
		light_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	const float scaled_radius = radius * magnitude(node.plane.vec);
	const float scaled_dist = dot(center, node.plane.vec);
	if (scaled_dist + scaled_radius >= node.plane.dist)
		position_object_sphere_r(id, center, radius, node.children[0]);
	if (scaled_dist - scaled_radius <= node.plane.dist)
		position_object_sphere_r(id, center, radius, node.children[1]);
}

void Renderer::position_object_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index) // :176-209
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_lights_[leaf_index].push_back(id);
		light_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	size_t vertices_front = 0;
	for (size_t i = 0; i < n; ++i)
		if (dot(node.plane.vec, vertices[i]) > node.plane.dist)
			++vertices_front;
	if (vertices_front > 0)
		position_object_convex_hull_r(id, vertices, n, node.children[0]);
	if (vertices_front < n)
		position_object_convex_hull_r(id, vertices, n, node.children[1]);
}

void Renderer::position_dynamic_lights() // dynamic_objects_index.rs:112-141
{
	for (auto& l : leaf_lights_)
		l.clear();
	light_leafs_.assign(frame_lights_.size(), {});
	for (size_t index = 0; index < frame_lights_.size(); ++index)
	{
		const DynamicLight& light = frame_lights_[index];
		if (light.shadow_type == SQW_SHADOW_PROJECTOR)
		{
			const Mat4f matrix = get_object_matrix(light.position, light.rotation);
			const float half_width = light.radius * std::tan(light.fov * 0.5f);
			const Vec3f local[5] = {{0, 0, 0}, {light.radius, half_width, half_width}, {light.radius, half_width, -half_width}, {light.radius, -half_width, half_width}, {light.radius, -half_width, -half_width}};
			Vec3f vt[5];
			for (int i = 0; i < 5; ++i)
				vt[i] = truncate(matrix * extend(local[i], 1.0f));
			position_object_convex_hull_r((uint32_t)index, vt, 5, map_.root_node());
		}
		else
			position_object_sphere_r((uint32_t)index, light.position, light.radius, map_.root_node());
	}
}

void Renderer::prepare_dynamic_lights() // partial_renderer.rs:398-522
{
	dynamic_lights_info_.assign(frame_lights_.size(), DynamicLightInfo{});
	size_t shadow_map_data_offset = 0;
	for (size_t index = 0; index < frame_lights_.size(); ++index)
	{
		const DynamicLight& light = frame_lights_[index];
		DynamicLightInfo& light_info = dynamic_lights_info_[index];
		for (uint32_t leaf_index : light_leafs_[index])
		{
			ClippingPolygon b;
			if (visibility_.leaf_bounds(leaf_index, b))
			{
				light_info.visible = true;
				break;
			}
		}
		if (!light_info.visible)
			continue;
		size_t shadow_map_data_size = 0;
		if (light.shadow_type == SQW_SHADOW_CUBE)
		{
			const float dist_from_camera = magnitude(camera_matrices_.position - light.position);
			const float dist_to_closest_point = dist_from_camera - light.radius;
			const uint32_t min_shadow_map_size = 64, max_shadow_map_size = 256;
			if (dist_to_closest_point <= 0.0f)
				light_info.shadow_map_size = max_shadow_map_size;
			else
			{
				const float target_size = 128.0f * light.radius / dist_to_closest_point;
				const uint32_t shift = std::min((uint32_t)std::fmax(std::log2(target_size) + config_.shadows_quality, 0.0f), 30u);
				light_info.shadow_map_size = std::max(std::min(1u << shift, max_shadow_map_size), min_shadow_map_size);
			}
			shadow_map_data_size = (size_t)light_info.shadow_map_size * light_info.shadow_map_size * 6;
		}
		else if (light.shadow_type == SQW_SHADOW_PROJECTOR)
		{
			const uint32_t min_shadow_map_size = 32, max_shadow_map_size = 1024;
			const float h = light.radius * std::tan(light.fov * 0.5f);
			const Mat4f light_matrix = get_object_matrix(light.position, light.rotation);
			float closest_square_dist = 1.0e24f;
			const Vec3f corners[5] = {{light.radius, 0, 0}, {light.radius, h, h}, {light.radius, h, -h}, {light.radius, -h, h}, {light.radius, -h, -h}};
			for (const Vec3f& v : corners)
			{
				const Vec3f d = camera_matrices_.position - truncate(light_matrix * extend(v, 1.0f));
				closest_square_dist = std::fmin(closest_square_dist, dot(d, d));
			}
			const float target_size = 1024.0f * h / std::sqrt(closest_square_dist);
			const uint32_t shift = std::min((uint32_t)std::fmax(std::log2(target_size) + config_.shadows_quality, 0.0f), 30u);
			light_info.shadow_map_size = std::max(std::min(1u << shift, max_shadow_map_size), min_shadow_map_size);
			shadow_map_data_size = (size_t)light_info.shadow_map_size * light_info.shadow_map_size;
		}
		light_info.shadow_map_data_offset = shadow_map_data_offset;
		shadow_map_data_offset += shadow_map_data_size;
	}
	shadow_maps_len_ = shadow_map_data_offset;

	// create_dynamic_light_with_shadow (renderer_structs.rs:228-308) flattened into the C-ABI record
	lights_.assign(frame_lights_.size(), sqw_light{});
	for (size_t index = 0; index < frame_lights_.size(); ++index)
	{
		const DynamicLight& light = frame_lights_[index];
		const DynamicLightInfo& light_info = dynamic_lights_info_[index];
		sqw_light& l = lights_[index];
		std::memcpy(l.position, &light.position.x, 12);
		l.radius = light.radius;
		l.inv_square_radius = 1.0f / (light.radius * light.radius);
		std::memcpy(l.color, light.color, 12);
		// An invisible light is in no visible leaf, so no visible polygon can select it (the reference gives it a dummy map).
		l.shadow_type = light_info.visible ? light.shadow_type : (uint32_t)SQW_SHADOW_NONE;
		l.shadow_map_size = light_info.shadow_map_size;
		l.shadow_map_offset = (uint32_t)light_info.shadow_map_data_offset;
		if (light.shadow_type == SQW_SHADOW_PROJECTOR)
		{
			const float inv_half_fov_tan = 1.0f / std::tan(light.fov * 0.5f);
			const Vec3f bx = quaternion_rotate_vector(light.rotation, Vec3f{0, 1, 0}) * inv_half_fov_tan;
			const Vec3f by = quaternion_rotate_vector(light.rotation, Vec3f{0, 0, 1}) * inv_half_fov_tan;
			const Vec3f bz = quaternion_rotate_vector(light.rotation, Vec3f{-1, 0, 0});
			std::memcpy(l.basis_x, &bx.x, 12);
			std::memcpy(l.basis_y, &by.x, 12);
			std::memcpy(l.basis_z, &bz.x, 12);
		}
	}
}

void Renderer::build_shadow_maps() // partial_renderer.rs:524-590: records the fills instead of rasterising on the CPU
{
	depth_maps_.clear();
	depth_cmds_.clear();
	depth_points_.clear();
	for (size_t index = 0; index < frame_lights_.size(); ++index)
	{
		const DynamicLight& light = frame_lights_[index];
		const DynamicLightInfo& light_info = dynamic_lights_info_[index];
		if (!light_info.visible)
			continue;
		const uint32_t size = light_info.shadow_map_size;
		if (light.shadow_type == SQW_SHADOW_CUBE)
		{
			for (uint32_t side = 0; side < 6; ++side)
			{
				const CameraMatrices m = complete_view_matrix(light.position, get_cube_map_side_matrix(side), Vec2f{1.0f, 1.0f}, (float)size, (float)size); // light.rs:72-84
				depth_draw_map((uint32_t)light_info.shadow_map_data_offset + side * size * size, size, SQW_DEPTH_MAP_CLEAR, m);
			}
		}
		else if (light.shadow_type == SQW_SHADOW_PROJECTOR)
		{
			const CameraMatrices m = build_view_matrix_with_full_rotation(light.position, light.rotation, light.fov, (float)size, (float)size); // light.rs:62-70
			depth_draw_map((uint32_t)light_info.shadow_map_data_offset, size, SQW_DEPTH_MAP_CLEAR | SQW_DEPTH_MAP_PROJECTOR_CIRCLE, m);
		}
	}
}

// DepthRenderer::draw_map (depth_renderer.rs:47-117): visibility from the light, BSP walk back to front, one fill per polygon.
// The reference does not clear the map (stale texels stay where no polygon lands); the recorded map is cleared to 0 first,
// which is what a first frame sees (shadow_maps_data is created zeroed, partial_renderer.rs:518-521).
void Renderer::depth_draw_map(uint32_t offset, uint32_t size, uint32_t flags, const CameraMatrices& depth_matrices)
{
	sqw_depth_map map;
	std::memset(&map, 0, sizeof(map));
	map.offset = offset;
	map.width = map.height = size;
	map.first_cmd = (uint32_t)depth_cmds_.size();
	map.flags = flags;
	map.clear_value = 0.0f;

	shadow_visibility_.update_visibility(depth_matrices, ClippingPolygon::from_box(0.0f, 0.0f, (float)size, (float)size));
	const Vec4f planes_matrix_w_row = depth_matrices.planes_matrix.row(3);
	uint32_t nodes_stack[MAX_BSP_TREE_DEPTH];
	nodes_stack[0] = map_.root_node();
	size_t num_nodes_on_stack = 1;
	while (num_nodes_on_stack > 0)
	{
		const uint32_t node_index = nodes_stack[num_nodes_on_stack - 1];
		if (node_index >= FIRST_LEAF_INDEX)
		{
			--num_nodes_on_stack;
			const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
			ClippingPolygon bounds;
			if (shadow_visibility_.leaf_bounds(leaf_index, bounds))
			{
				const BSPLeaf& leaf = map_.leafs[leaf_index];
				Vec3f clip_planes[8];
				bounds.get_clip_planes(clip_planes);
				for (uint32_t p = leaf.first_polygon; p < leaf.first_polygon + leaf.num_polygons; ++p)
					depth_draw_polygon(depth_matrices, clip_planes, p, false); // back to front, no depth test (:133-137)
			}
		}
		else if (num_nodes_on_stack < MAX_BSP_TREE_DEPTH)
		{
			const BSPNode& node = map_.nodes[node_index];
			const float plane_transformed_w = dot(planes_matrix_w_row, extend(node.plane.vec, -node.plane.dist));
			const uint32_t mask = plane_transformed_w >= 0.0f ? 1 : 0;
			nodes_stack[num_nodes_on_stack - 1] = node.children[1 ^ mask];
			nodes_stack[num_nodes_on_stack] = node.children[mask];
			++num_nodes_on_stack;
		}
	}
	map.num_cmds = (uint32_t)depth_cmds_.size() - map.first_cmd;
	depth_maps_.push_back(map);
}

void Renderer::depth_draw_polygon(const CameraMatrices& m, const Vec3f clip_planes[8], uint32_t polygon_index, bool depth_test) // depth_renderer.rs:204-286
{
	const Polygon& polygon = map_.polygons[polygon_index];
	const Material& material = scene_.materials[polygon.texture];
	if (!(material.shadow && material.blending_mode == SQW_BLEND_NONE)) // materials_shadow_table (:21-37)
		return;
	const Vec4f plane_transformed = m.planes_matrix * extend(polygon.plane.vec, -polygon.plane.dist);
	if (plane_transformed.w <= 0.0f)
		return;
	const float plane_transformed_w = -plane_transformed.w;
	const float d_inv_z_dx = plane_transformed.x / plane_transformed_w;
	const float d_inv_z_dy = plane_transformed.y / plane_transformed_w;
	const float DEPTH_BIAS_CONST = -1.0f / (float)(1 << 20);
	const float DEPTH_BIAS_SLOPE = -1.0f;
	const float DEPTH_EQUATION_SCALE = 1.0f - 1.0f / 1024.0f;
	sqw_depth_cmd c;
	std::memset(&c, 0, sizeof(c));
	c.depth_eq[0] = DEPTH_EQUATION_SCALE * d_inv_z_dx;
	c.depth_eq[1] = DEPTH_EQUATION_SCALE * d_inv_z_dy;
	c.depth_eq[2] = DEPTH_EQUATION_SCALE * (plane_transformed.z / plane_transformed_w + DEPTH_BIAS_CONST + DEPTH_BIAS_SLOPE * (std::fabs(d_inv_z_dx) + std::fabs(d_inv_z_dy)));
	c.depth_test = depth_test ? 1 : 0;

	Vec3f vertices_transformed[MAX_VERTICES];
	const size_t vertex_count = std::min((size_t)polygon.num_vertices, MAX_VERTICES);
	for (size_t i = 0; i < vertex_count; ++i)
		vertices_transformed[i] = view_matrix_transform_vertex(m.view_matrix, map_.vertices[polygon.first_vertex + i]);
	if (vertex_count < 3)
		return;
	Vec2f vertices_2d[MAX_VERTICES];
	const size_t n = project_and_clip_polygon(clip_planes, vertices_transformed, vertex_count, vertices_2d);
	if (n < 3)
		return;
	c.num_vertices = (uint32_t)n;
	c.vertex_offset = (uint32_t)depth_points_.size();
	for (size_t i = 0; i < n; ++i)
		depth_points_.push_back(sqw_polygon_point{f32_to_fixed16_unchecked(vertices_2d[i].x), f32_to_fixed16_unchecked(vertices_2d[i].y)});
	depth_cmds_.push_back(c);
}

void Renderer::prepare_frame(uint32_t width, uint32_t height, const CameraPose& camera, float fov)
{
	width_ = width;
	height_ = height;
	camera_matrices_ = build_view_matrix(camera.position, camera.azimuth, camera.elevation, fov, (float)width, (float)height);
	mip_bias_ = config_.textures_mip_bias; // update_mip_bias with dynamic_mip_bias = false (partial_renderer.rs:3786-3804)
	++current_frame_;
	stats_ = RendererDebugStats{};
	visibility_.update_visibility(camera_matrices_, ClippingPolygon::from_box(0.0f, 0.0f, (float)width, (float)height)); // partial_renderer.rs:173-190
	position_dynamic_lights();   // renderer.rs:88 (DynamicObjectsIndex)
	prepare_dynamic_lights();    // partial_renderer.rs:192
	build_shadow_maps();         // :194
	prepare_polygons_surfaces(); // :214
	position_models();              // renderer.rs:84 (dynamic_models_index.position_models)
	prepare_dynamic_models();       // partial_renderer.rs:209
	build_dynamic_models_buffers(); // :227
	build_surface_jobs();        // :221
}

void Renderer::prepare_polygons_surfaces() // partial_renderer.rs:1327-1449 (leaf polygons; submodels are not part of the synthetic maps)
{
	current_frame_visible_polygons_.clear();
	num_visible_surfaces_pixels_ = 0;
	has_current_sky_ = false; // :1336
	for (size_t i = 0; i < map_.leafs.size(); ++i)
	{
		ClippingPolygon bounds;
		if (!visibility_.leaf_bounds((uint32_t)i, bounds))
			continue;
		++stats_.num_visible_leafs;
		const BSPLeaf& leaf = map_.leafs[i];
		Vec3f clip_planes[8];
		bounds.get_clip_planes(clip_planes);
		for (uint32_t p = leaf.first_polygon; p < leaf.first_polygon + leaf.num_polygons; ++p)
			prepare_polygon_surface(clip_planes, p);
	}
	stats_.num_visible_polygons = (uint32_t)current_frame_visible_polygons_.size();
	stats_.num_surfaces_pixels = num_visible_surfaces_pixels_;
}

void Renderer::prepare_polygon_surface(const Vec3f clip_planes[8], size_t polygon_index) // partial_renderer.rs:1464-1638
{
	DrawPolygonData& polygon_data = polygons_data_[polygon_index];
	const Polygon& polygon = map_.polygons[polygon_index];

	const Vec4f plane_transformed = camera_matrices_.planes_matrix * extend(polygon.plane.vec, -polygon.plane.dist);
	if (plane_transformed.w <= 0.0f) // cull back faces
		return;
	const Material& material = scene_.materials[polygon.texture];
	if (!material.draw)
		return;

	const Vec3f* polygon_vertices = &map_.vertices[polygon.first_vertex];
	Vec3f* polygon_vertices_transformed = &vertices_transformed_[polygon.first_vertex];
	for (uint32_t i = 0; i < polygon.num_vertices; ++i)
		polygon_vertices_transformed[i] = view_matrix_transform_vertex(camera_matrices_.view_matrix, polygon_vertices[i]);

	Vec2f vertices_2d[MAX_VERTICES];
	const size_t vertex_count = project_and_clip_polygon(clip_planes, polygon_vertices_transformed, polygon.num_vertices, vertices_2d);
	if (vertex_count < 3)
		return;

	if (material.is_skybox) // :1511-1529: sky polygons are not drawn, they only grow the window the sky box is seen through
	{
		ClippingPolygon polygon_bounds = ClippingPolygon::from_point(vertices_2d[0]);
		for (size_t i = 1; i < vertex_count; ++i)
			polygon_bounds.extend_with_point(vertices_2d[i]);
		if (has_current_sky_)
			current_sky_bounds_.extend(polygon_bounds);
		else
		{
			has_current_sky_ = true;
			current_sky_material_ = polygon.texture;
			current_sky_bounds_ = polygon_bounds;
		}
		return;
	}

	// DepthEquation::from_transformed_plane_equation (equations.rs:22-31)
	const float plane_transformed_w = -plane_transformed.w;
	const float depth_equation[3] = {plane_transformed.x / plane_transformed_w, plane_transformed.y / plane_transformed_w, plane_transformed.z / plane_transformed_w};

	const Vec4f tc_basis_transformed[2] = {
		camera_matrices_.planes_matrix * extend(polygon.tex_coord_equation[0].vec, polygon.tex_coord_equation[0].dist),
		camera_matrices_.planes_matrix * extend(polygon.tex_coord_equation[1].vec, polygon.tex_coord_equation[1].dist)};
	// TexCoordEquation::from_depth_equation_and_transformed_tex_coord_equations (equations.rs:53-75)
	float tc_equation[6];
	for (int i = 0; i < 2; ++i)
	{
		tc_equation[i] = tc_basis_transformed[i].x + tc_basis_transformed[i].w * depth_equation[0];
		tc_equation[2 + i] = tc_basis_transformed[i].y + tc_basis_transformed[i].w * depth_equation[1];
		tc_equation[4 + i] = tc_basis_transformed[i].z + tc_basis_transformed[i].w * depth_equation[2];
	}

	const uint32_t mip = calculate_mip(vertices_2d, vertex_count, depth_equation, tc_equation, mip_bias_);
	const float mip_scale = 1.0f / (float)(1 << mip);
	float tc_equation_scaled[6];
	for (int i = 0; i < 6; ++i)
		tc_equation_scaled[i] = tc_equation[i] * mip_scale;

	const float inf = (float)(1 << 29);
	const float max_z = (float)(1 << 16);
	float tc_min[2] = {inf, inf};
	float tc_max[2] = {-inf, -inf};
	for (size_t k = 0; k < vertex_count; ++k)
	{
		const Vec2f p = vertices_2d[k];
		const float inv_z = std::fmax(sample_point(depth_equation, p), 1.0f / max_z);
		const float z = 1.0f / inv_z;
		for (int i = 0; i < 2; ++i)
		{
			const float tc = z * (tc_equation_scaled[i] * p.x + tc_equation_scaled[2 + i] * p.y + tc_equation_scaled[4 + i]);
			if (tc < tc_min[i])
				tc_min[i] = tc;
			if (tc > tc_max[i])
				tc_max[i] = tc;
		}
	}

	int32_t surface_tc_min[2] = {0, 0};
	int32_t surface_size[2] = {0, 0};
	for (int i = 0; i < 2; ++i)
	{
		const float tc_reduce_eps = 1.0f / 32.0f;
		tc_min[i] += tc_reduce_eps;
		tc_max[i] -= tc_reduce_eps;
		const int32_t round_mask = ~((int32_t)LIGHTMAP_SCALE - 1);
		const int32_t tc_min_round_down = (polygon.tex_coord_min[i] & round_mask) >> mip;
		const int32_t tc_max_round_up = ((polygon.tex_coord_max[i] + (int32_t)LIGHTMAP_SCALE - 1) & round_mask) >> mip;
		int32_t tc_min_int = std::max((int32_t)std::floor(std::fmax(tc_min[i], -inf)), tc_min_round_down);
		int32_t tc_max_int = std::min((int32_t)std::ceil(std::fmin(tc_max[i], inf)), tc_max_round_up);
		if (tc_min_int >= tc_max_int)
		{
			tc_min_int = std::min(tc_min_int, tc_max_round_up - 1);
			tc_max_int = tc_min_int + 1;
		}
		surface_tc_min[i] = tc_min_int;
		surface_size[i] = tc_max_int - tc_min_int;
	}

	const size_t surface_pixels_offset = num_visible_surfaces_pixels_;
	num_visible_surfaces_pixels_ += (size_t)(surface_size[0] * surface_size[1]);

	polygon_data.visible_frame = current_frame_;
	std::memcpy(polygon_data.depth_equation, depth_equation, sizeof(depth_equation));
	std::memcpy(polygon_data.tex_coord_equation, tc_equation_scaled, sizeof(tc_equation_scaled));
	polygon_data.surface_pixels_offset = surface_pixels_offset;
	polygon_data.surface_size[0] = (uint32_t)surface_size[0];
	polygon_data.surface_size[1] = (uint32_t)surface_size[1];
	polygon_data.mip = mip;
	polygon_data.surface_tc_min[0] = surface_tc_min[0];
	polygon_data.surface_tc_min[1] = surface_tc_min[1];
	for (int i = 0; i < 2; ++i) // :1629-1635 shift of the tc equation to the surface rect
	{
		const float tcm = (float)surface_tc_min[i];
		polygon_data.tex_coord_equation[i] -= tcm * depth_equation[0];
		polygon_data.tex_coord_equation[2 + i] -= tcm * depth_equation[1];
		polygon_data.tex_coord_equation[4 + i] -= tcm * depth_equation[2];
	}
	current_frame_visible_polygons_.push_back((uint32_t)polygon_index);
}

namespace
{
// partial_renderer.rs:3877-3985
bool polygon_is_affected_by_light(const Polygon& polygon, const DrawPolygonData& basis_vecs, const sqw_light& light)
{
	const Vec3f light_position{light.position[0], light.position[1], light.position[2]};
	const Vec3f vec_from_light_position_to_tc_start_point = light_position - basis_vecs.basis_start;
	const float signed_distance_to_polygon_plane = dot(vec_from_light_position_to_tc_start_point, basis_vecs.basis_normal);
	if (signed_distance_to_polygon_plane <= 0.0f)
		return false;
	const float square_radius_at_polygon_plane = light.radius * light.radius - signed_distance_to_polygon_plane * signed_distance_to_polygon_plane;
	if (square_radius_at_polygon_plane <= 0.0f)
		return false;
	if (light.shadow_type == SQW_SHADOW_PROJECTOR)
	{
		const Vec3f u0 = basis_vecs.basis_u * (float)polygon.tex_coord_min[0], u1 = basis_vecs.basis_u * (float)polygon.tex_coord_max[0];
		const Vec3f v0 = basis_vecs.basis_v * (float)polygon.tex_coord_min[1], v1 = basis_vecs.basis_v * (float)polygon.tex_coord_max[1];
		const Vec3f vertices[4] = {basis_vecs.basis_start + u0 + v0, basis_vecs.basis_start + u0 + v1, basis_vecs.basis_start + u1 + v0, basis_vecs.basis_start + u1 + v1};
		const Vec3f bx{light.basis_x[0], light.basis_x[1], light.basis_x[2]}, by{light.basis_y[0], light.basis_y[1], light.basis_y[2]}, bz{light.basis_z[0], light.basis_z[1], light.basis_z[2]};
		const Vec3f plane_vecs[5] = {bz, bz - bx, bz + bx, bz - by, bz + by};
		for (const Vec3f& plane_vec : plane_vecs)
		{
			const float dist = dot(plane_vec, light_position);
			int vertices_front = 0;
			for (const Vec3f& vertex : vertices)
				if (dot(plane_vec, vertex) >= dist)
					++vertices_front;
			if (vertices_front >= 4)
				return false;
		}
		return true;
	}
	// texture coordinates of the projection of the light onto the polygon plane: rows of inverse(cols(u, v, n))
	const Vec3f u = basis_vecs.basis_u, v = basis_vecs.basis_v, n = basis_vecs.basis_normal;
	const float det = dot(u, cross(v, n));
	if (det == 0.0f || !std::isfinite(det))
		return false;
	const Vec3f mat_x = cross(v, n) / det, mat_y = cross(n, u) / det; // (M^-1)^T columns x, y
	const float tc_u = dot(vec_from_light_position_to_tc_start_point, mat_x), tc_v = dot(vec_from_light_position_to_tc_start_point, mat_y);
	const float radius_at_polygon_plane = std::sqrt(square_radius_at_polygon_plane);
	const float u_radius = radius_at_polygon_plane * (1.0f / std::sqrt(dot(u, u)));
	const float v_radius = radius_at_polygon_plane * (1.0f / std::sqrt(dot(v, v)));
	return !(tc_u + u_radius < (float)polygon.tex_coord_min[0] || tc_v + v_radius < (float)polygon.tex_coord_min[1] || tc_u - u_radius > (float)polygon.tex_coord_max[0] ||
			 tc_v - v_radius > (float)polygon.tex_coord_max[1]);
}
} // namespace

void Renderer::build_surface_jobs() // the closure of build_polygons_surfaces (partial_renderer.rs:1682-1819) turned into job records
{
	jobs_.clear();
	jobs_.reserve(current_frame_visible_polygons_.size());
	const bool use_directional_lightmap = config_.use_directional_lightmaps && !map_.directional_lightmaps_data.empty();
	for (uint32_t polygon_index : current_frame_visible_polygons_)
	{
		const Polygon& polygon = map_.polygons[polygon_index];
		const DrawPolygonData& polygon_data = polygons_data_[polygon_index];
		const Material& material = scene_.materials[polygon.texture];
		sqw_surface_job j;
		std::memset(&j, 0, sizeof(j));
		const int32_t extra_shift_mip_compensated[2] = {material.texture_shift[0] >> polygon_data.mip, material.texture_shift[1] >> polygon_data.mip};
		j.surface_size[0] = polygon_data.surface_size[0];
		j.surface_size[1] = polygon_data.surface_size[1];
		j.tc_start[0] = polygon_data.surface_tc_min[0] + extra_shift_mip_compensated[0];
		j.tc_start[1] = polygon_data.surface_tc_min[1] + extra_shift_mip_compensated[1];
		j.surface_pixels_offset = (uint32_t)polygon_data.surface_pixels_offset;
		const bool apply_diffuse_layer = material.light && material.has_diffuse;
		if (apply_diffuse_layer)
		{
			j.flags |= SQW_SURF_APPLY_DIFFUSE;
			// get_basis_vecs_for_mip (surfaces.rs:49-63) + shift correction (partial_renderer.rs:1736-1740)
			const float scale = (float)(1 << polygon_data.mip);
			const Vec3f u = polygon_data.mip == 0 ? polygon_data.basis_u : polygon_data.basis_u * scale;
			const Vec3f v = polygon_data.mip == 0 ? polygon_data.basis_v : polygon_data.basis_v * scale;
			const Vec3f start = polygon_data.basis_start - (u * (float)extra_shift_mip_compensated[0] + v * (float)extra_shift_mip_compensated[1]);
			const Vec3f nrm = polygon_data.basis_normal;
			const float* src[4] = {&u.x, &v.x, &start.x, &nrm.x};
			float* dst[4] = {j.basis_u, j.basis_v, j.basis_start, j.basis_normal};
			for (int k = 0; k < 4; ++k)
				std::memcpy(dst[k], src[k], 12);
			j.texture_image = material.texture_images[polygon_data.mip];
			if (material.has_normal_map)
				j.flags |= SQW_SURF_HAS_NORMAL_MAP;
			if (material.has_non_one_roughness)
				j.flags |= SQW_SURF_HAS_NON_ONE_ROUGHNESS;
			if (material.is_metal)
				j.flags |= SQW_SURF_IS_METAL;
			for (int i = 0; i < 2; ++i) // :1744-1752
			{
				const int32_t round_mask = ~((int32_t)LIGHTMAP_SCALE - 1);
				const int32_t shift = polygon_data.surface_tc_min[i] - ((polygon.tex_coord_min[i] & round_mask) >> polygon_data.mip);
				j.lightmap_tc_shift[i] = (uint32_t)shift;
			}
			j.lightmap_size[0] = get_lightmap_size(polygon.tex_coord_min[0], polygon.tex_coord_max[0]);
			j.lightmap_size[1] = get_lightmap_size(polygon.tex_coord_min[1], polygon.tex_coord_max[1]);
			j.lightmap_scale_log2 = LIGHTMAP_SCALE_LOG2 - polygon_data.mip;
			j.lightmap_offset = polygon.lightmap_data_offset;
			if (use_directional_lightmap)
				j.flags |= SQW_SURF_DIRECTIONAL_LIGHTMAP;
			// lights of the polygon's leaf that reach it, at most MAX_POLYGON_LIGHTS (partial_renderer.rs:1705-1734)
			for (uint32_t light_index : leaf_lights_[polygon_data.parent_leaf])
			{
				if (polygon_is_affected_by_light(polygon, polygon_data, lights_[light_index]))
				{
					j.lights[j.num_lights++] = (uint16_t)light_index;
					if (j.num_lights == SQW_MAX_POLYGON_LIGHTS)
						break;
				}
			}
		}
		if (material.has_emissive)
		{
			j.flags |= SQW_SURF_EMISSIVE;
			j.emissive_image = material.emissive_images[polygon_data.mip];
			std::memcpy(j.emissive_light, material.emissive_light, 12);
		}
		jobs_.push_back(j);
	}
}

void Renderer::draw_frame(int color_bits) // PartialRenderer::draw_frame -> perform_rasterization, single-threaded branch (partial_renderer.rs:1207-1233)
{
	cmds_.clear();
	polygon_points_.clear();
	triangle_points_.clear();
	const ClippingPolygon viewport_clipping_polygon = ClippingPolygon::from_box(0.0f, 0.0f, (float)width_, (float)height_);
	draw_skybox(viewport_clipping_polygon, color_bits); // perform_rasterization_for_viewport_part (:1296): sky first
	draw_tree(viewport_clipping_polygon);
	stats_.num_draw_commands = (uint32_t)cmds_.size();
}

void Renderer::draw_skybox(const ClippingPolygon& viewport_clipping_polygon, int color_bits) // partial_renderer.rs:2311-2448 (skybox_rotation = identity)
{
	if (!has_current_sky_)
		return;
	ClippingPolygon bounds = current_sky_bounds_;
	bounds.intersect(viewport_clipping_polygon);
	if (!bounds.is_valid_and_non_empty())
		return;
	const Material& material = scene_.materials[current_sky_material_];
	if (!material.is_skybox)
		return;
	static const float BOX_VERTICES[8][3] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
	struct Side
	{
		int v[4];
		Vec3f normal, u, w;
	};
	static const Side bbox_polygons[6] = {
		{{0, 1, 3, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, -1}},   // -X
		{{4, 6, 7, 5}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}}, // +X
		{{0, 4, 5, 1}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}},  // -Y
		{{2, 3, 7, 6}, {0, -1, 0}, {1, 0, 0}, {0, 0, -1}},  // +Y
		{{0, 2, 6, 4}, {0, 0, 1}, {1, 0, 0}, {0, -1, 0}},   // -Z
		{{1, 5, 7, 3}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}},   // +Z
	};
	const float side_plane_dist = -1.0f, side_tc_shift = 1.0f;
	const Mat4f skybox_matrix = get_object_matrix(camera_matrices_.position, Quaternion{});
	Mat4f skybox_matrix_inverse = Mat4f::identity();
	skybox_matrix.transpose().invert(skybox_matrix_inverse);
	const Mat4f skybox_view_matrix = camera_matrices_.view_matrix * skybox_matrix;
	const Mat4f skybox_planes_matrix = camera_matrices_.planes_matrix * skybox_matrix_inverse;
	const float SKY_SCALE = 16384.0f;
	Vec3f box_vertices_transformed[8];
	for (int i = 0; i < 8; ++i)
		box_vertices_transformed[i] = view_matrix_transform_vertex(skybox_view_matrix, Vec3f{BOX_VERTICES[i][0], BOX_VERTICES[i][1], BOX_VERTICES[i][2]} * SKY_SCALE);
	Vec3f clip_planes[8];
	bounds.get_clip_planes(clip_planes);
	const int w = color_bits == 64 ? 1 : 0;
	for (int side = 0; side < 6; ++side)
	{
		const Side& polygon = bbox_polygons[side];
		const HostImage& mip0 = scene_.images[material.skybox_images[w][side][0]];
		if (mip0.width == 0)
			continue;
		const Vec3f vertices_transformed[4] = {box_vertices_transformed[polygon.v[0]], box_vertices_transformed[polygon.v[1]], box_vertices_transformed[polygon.v[2]], box_vertices_transformed[polygon.v[3]]};
		const Vec4f plane_transformed = skybox_planes_matrix * extend(polygon.normal, -side_plane_dist);
		const float plane_transformed_w = -plane_transformed.w;
		const float depth_equation[3] = {plane_transformed.x / plane_transformed_w, plane_transformed.y / plane_transformed_w, plane_transformed.z / plane_transformed_w};
		const float tc_equation_scale = (float)mip0.width * 0.5f;
		const Vec4f tc_basis_transformed[2] = {skybox_planes_matrix * (extend(polygon.u, side_tc_shift) * tc_equation_scale), skybox_planes_matrix * (extend(polygon.w, side_tc_shift) * tc_equation_scale)};
		float tc_equation[6];
		for (int i = 0; i < 2; ++i)
		{
			tc_equation[i] = tc_basis_transformed[i].x + tc_basis_transformed[i].w * depth_equation[0];
			tc_equation[2 + i] = tc_basis_transformed[i].y + tc_basis_transformed[i].w * depth_equation[1];
			tc_equation[4 + i] = tc_basis_transformed[i].z + tc_basis_transformed[i].w * depth_equation[2];
		}
		Vec2f vertices_2d[MAX_VERTICES];
		const size_t vertex_count = project_and_clip_polygon(clip_planes, vertices_transformed, 4, vertices_2d);
		if (vertex_count < 3)
			continue;
		const uint32_t mip = calculate_mip(vertices_2d, vertex_count, depth_equation, tc_equation, mip_bias_);
		float tc_equation_scaled[6];
		for (int i = 0; i < 6; ++i)
			tc_equation_scaled[i] = tc_equation[i] * (1.0f / (float)(1 << mip));
		const uint32_t image_id = material.skybox_images[w][side][mip];
		const HostImage& side_texture = scene_.images[image_id];
		const uint32_t size[2] = {side_texture.width, side_texture.width};
		draw_polygon(clip_planes, vertices_transformed, 4, depth_equation, tc_equation_scaled, size, SQW_TEX_IMAGE, image_id, SQW_BLEND_NONE);
	}
}

void Renderer::draw_tree(const ClippingPolygon& viewport_clipping_polygon) // partial_renderer.rs:2450-2517
{
	const Vec4f planes_matrix_w_row = camera_matrices_.planes_matrix.row(3);
	uint32_t nodes_stack[MAX_BSP_TREE_DEPTH];
	nodes_stack[0] = map_.root_node();
	size_t num_nodes_on_stack = 1;
	while (num_nodes_on_stack > 0)
	{
		const uint32_t node_index = nodes_stack[num_nodes_on_stack - 1];
		if (node_index >= FIRST_LEAF_INDEX)
		{
			--num_nodes_on_stack;
			const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
			ClippingPolygon bounds;
			if (visibility_.leaf_bounds(leaf_index, bounds))
			{
				bounds.intersect(viewport_clipping_polygon);
				if (bounds.is_valid_and_non_empty())
					draw_leaf(bounds, leaf_index);
			}
		}
		else if (num_nodes_on_stack < MAX_BSP_TREE_DEPTH)
		{
			const BSPNode& node = map_.nodes[node_index];
			const float plane_transformed_w = dot(planes_matrix_w_row, extend(node.plane.vec, -node.plane.dist));
			uint32_t mask = plane_transformed_w >= 0.0f ? 1 : 0;
			if (config_.invert_polygons_order)
				mask ^= 1;
			nodes_stack[num_nodes_on_stack - 1] = node.children[1 ^ mask];
			nodes_stack[num_nodes_on_stack] = node.children[mask];
			++num_nodes_on_stack;
		}
	}
}

void Renderer::draw_leaf(const ClippingPolygon& bounds, uint32_t leaf_index) // partial_renderer.rs:2519-2565 (leaf polygons)
{
	const BSPLeaf& leaf = map_.leafs[leaf_index];
	Vec3f clip_planes[8];
	bounds.get_clip_planes(clip_planes);
	for (uint32_t polygon_index = leaf.first_polygon; polygon_index < leaf.first_polygon + leaf.num_polygons; ++polygon_index)
	{
		const Polygon& polygon = map_.polygons[polygon_index];
		const DrawPolygonData& polygon_data = polygons_data_[polygon_index];
		if (polygon_data.visible_frame != current_frame_)
			continue;
		draw_polygon(
			clip_planes, &vertices_transformed_[polygon.first_vertex], polygon.num_vertices, polygon_data.depth_equation, polygon_data.tex_coord_equation,
			polygon_data.surface_size, SQW_TEX_SURFACE, (uint32_t)polygon_data.surface_pixels_offset, scene_.materials[polygon.texture].blending_mode);
	}
	draw_leaf_objects(bounds, leaf_index);
}

void Renderer::draw_polygon( // partial_renderer.rs:4011-4142, recording instead of rasterising
	const Vec3f clip_planes[8], const Vec3f* vertices_transformed, size_t num_vertices, const float depth_equation[3], const float tex_coord_equation[6],
	const uint32_t texture_size[2], uint8_t tex_kind, uint32_t tex_ref, int blending_mode)
{
	if (num_vertices < 3)
		return;
	Vec2f vertices_2d[MAX_VERTICES];
	const size_t vertex_count = project_and_clip_polygon(clip_planes, vertices_transformed, num_vertices, vertices_2d);
	if (vertex_count < 3)
		return;

	float min_inv_z = 1e24f, max_inv_z = -1e24f, min_x = 1e24f, max_x = -1e24f;
	Vec2f min_inv_z_point = vertices_2d[0], max_inv_z_point = vertices_2d[0];
	for (size_t i = 0; i < vertex_count; ++i)
	{
		const Vec2f point = vertices_2d[i];
		const float inv_z = sample_point(depth_equation, point);
		if (inv_z < min_inv_z)
		{
			min_inv_z = inv_z;
			min_inv_z_point = point;
		}
		if (inv_z > max_inv_z)
		{
			max_inv_z = inv_z;
			max_inv_z_point = point;
		}
		if (point.x < min_x)
			min_x = point.x;
		if (point.x > max_x)
			max_x = point.x;
	}

	sqw_draw_cmd c;
	std::memset(&c, 0, sizeof(c));
	c.kind = SQW_CMD_POLYGON;
	c.blend = (uint8_t)blending_mode;
	c.tex_kind = tex_kind;
	c.tex_ref = tex_ref;
	c.tex_size[0] = (int32_t)texture_size[0];
	c.tex_size[1] = (int32_t)texture_size[1];
	c.num_vertices = (uint32_t)vertex_count;
	c.vertex_offset = (uint32_t)polygon_points_.size();
	for (size_t i = 0; i < vertex_count; ++i)
		polygon_points_.push_back(sqw_polygon_point{f32_to_fixed16_unchecked(vertices_2d[i].x), f32_to_fixed16_unchecked(vertices_2d[i].y)});

	if (affine_texture_coordinates_interpolation_may_be_used(depth_equation, tex_coord_equation, min_inv_z_point, max_inv_z_point))
	{
		c.tc_mode = SQW_TC_AFFINE;
		std::memcpy(c.depth_eq, depth_equation, 12);
		std::memcpy(c.tc_eq, tex_coord_equation, 24);
	}
	else
	{
		// power-of-two renormalisation (:4106-4109)
		const float z_scale = std::exp2(-5.0f - std::ceil(std::log2(std::fmax(max_inv_z, 1.0f / (float)(1 << 20)))));
		for (int i = 0; i < 3; ++i)
			c.depth_eq[i] = depth_equation[i] * z_scale;
		for (int i = 0; i < 6; ++i)
			c.tc_eq[i] = tex_coord_equation[i] * z_scale;
		c.tc_mode = line_z_corrected_texture_coordinates_interpolation_may_be_used(depth_equation, tex_coord_equation, max_inv_z_point, min_x, max_x)
			? SQW_TC_LINE_Z_CORRECTION
			: SQW_TC_FULL_PERSPECTIVE;
	}
	++stats_.mode_counts[c.tc_mode];
	cmds_.push_back(c);
}


// ---- triangle models ---------------------------------------------------------------------------------------------
namespace
{
constexpr size_t MAX_LEAF_CLIP_PLANES = 20; // renderer_utils.rs:5

inline int32_t f32_to_fixed16(float x) // fixed_math.rs:31-34: Rust `as` (saturating, NaN -> 0)
{
	const float v = x * 65536.0f;
	if (!(v == v))
		return 0;
	if (v >= 2147483648.0f)
		return INT32_MAX;
	if (v <= -2147483648.0f)
		return INT32_MIN;
	return (int32_t)v;
}

// triangle_models_rendering.rs:213-273
bool calculate_triangle_model_screen_polygon(const Vec3f v[8], ClippingPolygon& out)
{
	bool have = false;
	int num_front_vertices = 0;
	for (int i = 0; i < 8; ++i)
		if (v[i].z > 0.0f)
		{
			++num_front_vertices;
			const Vec2f point = truncate(v[i]) / v[i].z;
			if (have)
				out.extend_with_point(point);
			else
			{
				out = ClippingPolygon::from_point(point);
				have = true;
			}
		}
	if (num_front_vertices == 8 || num_front_vertices == 0)
		return have;
	const float Z_NEAR_BBOX = 1.0f / 4096.0f;
	for (int i = 0; i < 7; ++i)
		for (int j = i + 1; j < 8; ++j)
		{
			if (v[i].z <= Z_NEAR_BBOX && v[j].z <= Z_NEAR_BBOX)
				continue;
			if (v[i].z >= Z_NEAR_BBOX && v[j].z >= Z_NEAR_BBOX)
				continue;
			const Vec3f p = get_line_z_intersection(v[i], v[j], Z_NEAR_BBOX);
			out.extend_with_point(truncate(p) / p.z);
		}
	return true;
}
} // namespace

void Renderer::position_model_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index) // dynamic_objects_index.rs:172-209
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_models_[leaf_index].push_back(id);
		model_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	size_t vertices_front = 0;
	for (size_t i = 0; i < n; ++i)
		if (dot(node.plane.vec, vertices[i]) > node.plane.dist)
			++vertices_front;
	if (vertices_front > 0)
		position_model_convex_hull_r(id, vertices, n, node.children[0]);
	if (vertices_front < n)
		position_model_convex_hull_r(id, vertices, n, node.children[1]);
}

void Renderer::position_models() // dynamic_objects_index.rs:52-73
{
	for (auto& l : leaf_models_)
		l.clear();
	model_leafs_.assign(frame_models_.size(), {});
	for (size_t index = 0; index < frame_models_.size(); ++index)
	{
		const ModelEntity& model = frame_models_[index];
		if (model.flags & MODEL_FLAG_VIEW_MODEL)
			continue;
		const BBox bbox = get_current_triangle_model_bbox(scene_.models[model.model], model.frames);
		const Mat4f matrix = get_object_matrix(model.position, model.rotation);
		Vec3f corners[8];
		bbox.get_corners_vertices(corners);
		for (Vec3f& c : corners)
			c = truncate(matrix * extend(c, 1.0f));
		position_model_convex_hull_r((uint32_t)index, corners, 8, map_.root_node());
	}
}

void Renderer::prepare_dynamic_models() // partial_renderer.rs:995-1120 (root view: is_third_person_view = false)
{
	visible_dynamic_meshes_list_.clear();
	dynamic_model_to_dynamic_meshes_index_.assign(frame_models_.size(), DynamicModelInfo{});
	size_t vertices_offset = 0, triangles_offset = 0;
	for (size_t entity_index = 0; entity_index < frame_models_.size(); ++entity_index)
	{
		const ModelEntity& model = frame_models_[entity_index];
		const TriangleModel& triangle_model = scene_.models[model.model];
		DynamicModelInfo& info = dynamic_model_to_dynamic_meshes_index_[entity_index];
		const Mat4f model_matrix = get_object_matrix(model.position, model.rotation);
		const Mat4f model_view_matrix = camera_matrices_.view_matrix * model_matrix;
		const BBox bbox = get_current_triangle_model_bbox(triangle_model, model.frames);
		Vec3f bbox_vertices_transformed[8];
		bbox.get_corners_vertices(bbox_vertices_transformed);
		for (Vec3f& c : bbox_vertices_transformed)
			c = view_matrix_transform_vertex(model_view_matrix, c);
		ClippingPolygon clipping_polygon;
		if (!calculate_triangle_model_screen_polygon(bbox_vertices_transformed, clipping_polygon))
			continue; // behind the camera plane
		if (model.flags & MODEL_FLAG_ONLY_THIRD_PERSON_VIEW)
			continue;
		bool visible = (model.flags & MODEL_FLAG_VIEW_MODEL) != 0;
		for (uint32_t leaf_index : model_leafs_[entity_index])
		{
			ClippingPolygon leaf_clipping_polygon;
			if (visibility_.leaf_bounds(leaf_index, leaf_clipping_polygon))
			{
				leaf_clipping_polygon.intersect(clipping_polygon);
				if (leaf_clipping_polygon.is_valid_and_non_empty())
				{
					visible = true;
					break;
				}
			}
		}
		if (!visible)
			continue;
		info.first_visible_mesh = (uint32_t)visible_dynamic_meshes_list_.size();
		CameraMatrices model_camera_matrices;
		model_camera_matrices.view_matrix = model_view_matrix;
		Mat4f inv_t = Mat4f::identity();
		model_matrix.transpose().invert(inv_t);
		model_camera_matrices.planes_matrix = camera_matrices_.planes_matrix * inv_t;
		model_camera_matrices.position = {0, 0, 0};
		for (size_t mesh_index = 0; mesh_index < triangle_model.meshes.size(); ++mesh_index)
		{
			VisibleDynamicMeshInfo m;
			m.entity_index = (uint32_t)entity_index;
			m.mesh_index = (uint32_t)mesh_index;
			m.vertices_offset = vertices_offset;
			m.triangles_offset = triangles_offset;
			std::memcpy(m.bbox_vertices_transformed, bbox_vertices_transformed, sizeof(bbox_vertices_transformed));
			m.clipping_polygon = clipping_polygon;
			m.model_matrix = model_matrix;
			m.camera_matrices = model_camera_matrices;
			visible_dynamic_meshes_list_.push_back(m);
			vertices_offset += triangle_model.meshes[mesh_index].num_vertices();
			triangles_offset += triangle_model.meshes[mesh_index].triangles.size();
		}
		info.num_visible_meshes = (uint32_t)triangle_model.meshes.size();
	}
	if (vertices_offset > dynamic_meshes_vertices_.size())
		dynamic_meshes_vertices_.resize(vertices_offset);
	if (triangles_offset > dynamic_meshes_triangles_.size())
		dynamic_meshes_triangles_.resize(triangles_offset);
	stats_.num_visible_meshes = (uint32_t)visible_dynamic_meshes_list_.size();
}

// get_model_light (triangle_models_rendering.rs:426-572).  Static part: the maps of this mirror carry no light grid, so
// fetch_light_from_grid returns its zero element (:574-586); dynamic part: the eight bbox sample positions against every
// dynamic light WITHOUT a shadow map.  Lights with cube / projector shadow maps need get_light_shadow_factor on shadow maps
// that live on the device here (rendered by depth_kernels.cu) — they do not light models in this mirror.
ModelLightData Renderer::get_model_light(const ModelEntity& model, const Mat4f& model_matrix) const
{
	LightGridElement static_light;
	if (model.lighting == MODEL_LIGHTING_CONSTANT)
	{
		for (auto& side : static_light.light_cube)
			std::memcpy(side, model.light, 12);
	}
	else if (model.lighting == MODEL_LIGHTING_ADVANCED)
	{
		for (auto& side : static_light.light_cube)
			std::memcpy(side, model.light, 12); // + grid_light_scale * (zero grid light)
	}
	LightCube light_cube;
	auto calculate_model_dynamic_light_cube = [&]() { // :494-537
		const BBox bbox = get_current_triangle_model_bbox(scene_.models[model.model], model.frames);
		const Vec3f bbox_center = bbox.get_center();
		Vec3f sample_positions[8];
		bbox.get_corners_vertices(sample_positions);
		for (Vec3f& v : sample_positions)
			v = truncate(model_matrix * extend((v + bbox_center) * 0.5f, 1.0f));
		const Vec3f size = bbox.get_size();
		const float min_square_distance = dot(size, size) * 0.25f;
		for (const DynamicLight& light : frame_lights_)
		{
			if (light.shadow_type != SQW_SHADOW_NONE)
				continue;
			const float inv_square_radius = 1.0f / (light.radius * light.radius);
			for (const Vec3f& position : sample_positions)
			{
				const Vec3f vec_to_light = light.position - position;
				const float square_dist = std::fmax(dot(vec_to_light, vec_to_light), min_square_distance);
				const float inv_square_dist = 1.0f / square_dist;
				if (inv_square_dist < inv_square_radius)
					continue;
				const float scale = 1.0f * (inv_square_dist - inv_square_radius);
				const float light_scaled[3] = {light.color[0] * scale, light.color[1] * scale, light.color[2] * scale};
				light_cube.add_light_sample(vec_to_light, light_scaled);
			}
		}
		light_cube.scale(1.0f / 8.0f);
	};
	switch (model.lighting)
	{
	case MODEL_LIGHTING_DEFAULT: calculate_model_dynamic_light_cube(); break;
	case MODEL_LIGHTING_CONSTANT: light_cube.add_constant_light(model.light); break;
	case MODEL_LIGHTING_ADVANCED:
		calculate_model_dynamic_light_cube();
		light_cube.scale(model.grid_light_scale);
		light_cube.add_constant_light(model.light);
		break;
	}
	return ModelLightData::from_two_light_grid_elements(static_light, light_cube.convert_into_light_grid_sample());
}

void Renderer::build_dynamic_models_buffers() // partial_renderer.rs:1122-1190
{
	for (VisibleDynamicMeshInfo& vm : visible_dynamic_meshes_list_)
	{
		const ModelEntity& model = frame_models_[vm.entity_index];
		const TriangleModel& triangle_model = scene_.models[model.model];
		const HostImage& mip0 = scene_.images[model.texture_images[0]];
		const uint32_t size0[2] = {mip0.width, mip0.height};
		vm.mip = calculate_triangle_model_texture_mip(vm.camera_matrices.view_matrix, get_current_triangle_model_bbox(triangle_model, model.frames), size0, mip_bias_);
		const HostImage& texture = scene_.images[model.texture_images[vm.mip]];
		const TriangleModelMesh& mesh = triangle_model.meshes[vm.mesh_index];
		ModelVertex3d* dst_vertices = dynamic_meshes_vertices_.data() + vm.vertices_offset;
		animate_and_transform_triangle_mesh_vertices(
			triangle_model, mesh, model.frames, model.lerp, get_model_light(model, vm.model_matrix), vm.model_matrix, vm.camera_matrices.view_matrix,
			Vec2f{(float)texture.width, (float)texture.height}, triangle_model.tc_shift, dst_vertices);
		Triangle* dst_triangles = dynamic_meshes_triangles_.data() + vm.triangles_offset;
		vm.num_visible_triangles = reject_triangle_model_back_faces(dst_vertices, mesh.triangles.data(), mesh.triangles.size(), dst_triangles);
		sort_model_triangles(dst_vertices, dst_triangles, vm.num_visible_triangles);
	}
}

// The tail of draw_leaf (partial_renderer.rs:2584-2790) for the object kinds of this mirror (triangle models).
void Renderer::draw_leaf_objects(const ClippingPolygon& bounds, uint32_t leaf_index)
{
	const std::vector<uint32_t>& leaf_dynamic_models = leaf_models_[leaf_index];
	if (leaf_dynamic_models.empty())
		return;
	Plane leaf_clip_planes[MAX_LEAF_CLIP_PLANES];
	const std::vector<Plane>& world = leafs_planes_[leaf_index];
	const size_t num_planes = std::min(MAX_LEAF_CLIP_PLANES, world.size());
	for (size_t i = 0; i < num_planes; ++i)
	{
		const Vec4f t = camera_matrices_.planes_matrix * extend(world[i].vec, -world[i].dist);
		leaf_clip_planes[i] = Plane{truncate(t), -t.w};
	}
	if (leaf_dynamic_models.size() == 1)
	{
		const DynamicModelInfo entry = dynamic_model_to_dynamic_meshes_index_[leaf_dynamic_models[0]];
		for (uint32_t m = entry.first_visible_mesh; m < entry.first_visible_mesh + entry.num_visible_meshes; ++m)
			draw_mesh(bounds, leaf_clip_planes, num_planes, m);
		return;
	}
	objects_sorter_.clear();
	for (uint32_t dynamic_model_index : leaf_dynamic_models)
	{
		const ModelEntity& model = frame_models_[dynamic_model_index];
		const BBox bbox = model.has_ordering_custom_bbox ? model.ordering_custom_bbox : get_current_triangle_model_bbox(scene_.models[model.model], model.frames);
		const DynamicModelInfo entry = dynamic_model_to_dynamic_meshes_index_[dynamic_model_index];
		for (uint32_t m = entry.first_visible_mesh; m < entry.first_visible_mesh + entry.num_visible_meshes; ++m)
			objects_sorter_.add_object(m, project_bbox(bbox, visible_dynamic_meshes_list_[m].camera_matrices));
	}
	objects_sorter_.sort_objects();
	for (size_t i = 0; i < objects_sorter_.num_objects; ++i)
		draw_mesh(bounds, leaf_clip_planes, num_planes, objects_sorter_.objects[i].first);
}

void Renderer::draw_mesh(const ClippingPolygon& clipping_polygon, const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, size_t visible_mesh_index) // partial_renderer.rs:3525-3717
{
	const VisibleDynamicMeshInfo& vm = visible_dynamic_meshes_list_[visible_mesh_index];
	{
		ClippingPolygon mesh_clipping_polygon = vm.clipping_polygon;
		mesh_clipping_polygon.intersect(clipping_polygon);
		if (!mesh_clipping_polygon.is_valid_and_non_empty())
			return;
	}
	const ModelEntity& model = frame_models_[vm.entity_index];
	Plane clip_planes_3d[MAX_LEAF_CLIP_PLANES + 1];
	size_t num_clip_planes_3d = 0;
	const Plane near_z_plane = {{0.0f, 0.0f, 1.0f}, Z_NEAR};
	for (size_t pi = 0; pi <= num_leaf_clip_planes; ++pi)
	{
		const Plane& clip_plane = pi == 0 ? near_z_plane : leaf_clip_planes[pi - 1];
		int vertices_front = 0;
		for (const Vec3f& v : vm.bbox_vertices_transformed)
			if (dot(clip_plane.vec, v) >= clip_plane.dist)
				++vertices_front;
		if (vertices_front == 8)
			continue;
		if (vertices_front == 0)
			return;
		if (num_clip_planes_3d < MAX_LEAF_CLIP_PLANES) // the reference's fixed array holds MAX_LEAF_CLIP_PLANES entries
			clip_planes_3d[num_clip_planes_3d++] = clip_plane;
	}
	Vec3f clip_planes_2d[4];
	size_t num_clip_planes_2d = 0;
	{
		Vec3f mesh_planes[4], cur_planes[4];
		vm.clipping_polygon.get_box_clip_planes(mesh_planes);
		clipping_polygon.get_box_clip_planes(cur_planes);
		for (int i = 0; i < 4; ++i)
			if (cur_planes[i].z > mesh_planes[i].z)
				clip_planes_2d[num_clip_planes_2d++] = cur_planes[i];
	}
	const uint32_t image_id = model.texture_images[vm.mip];
	const HostImage& texture = scene_.images[image_id];
	const ModelVertex3d* vertices_combined = dynamic_meshes_vertices_.data() + vm.vertices_offset;
	const Triangle* triangles = dynamic_meshes_triangles_.data() + vm.triangles_offset;

	auto emit = [&](const sqw_triangle_point pts[3]) {
		sqw_draw_cmd c;
		std::memset(&c, 0, sizeof(c));
		c.kind = SQW_CMD_TRIANGLE;
		c.blend = (uint8_t)model.blending_mode;
		c.tex_kind = SQW_TEX_IMAGE;
		c.tex_ref = image_id;
		c.tex_size[0] = (int32_t)texture.width;
		c.tex_size[1] = (int32_t)texture.height;
		c.num_vertices = 3;
		c.vertex_offset = (uint32_t)triangle_points_.size();
		triangle_points_.insert(triangle_points_.end(), pts, pts + 3);
		cmds_.push_back(c);
		++stats_.num_triangles;
	};
	auto to_point = [](Vec2f pos, Vec2f tc, const float light[3]) {
		sqw_triangle_point p;
		p.x = f32_to_fixed16(pos.x);
		p.y = f32_to_fixed16(pos.y);
		p.tc[0] = f32_to_fixed16(tc.x);
		p.tc[1] = f32_to_fixed16(tc.y);
		for (int i = 0; i < 3; ++i)
			p.light[i] = f32_to_fixed16(light[i]);
		return p;
	};

	if (num_clip_planes_3d == 0 && num_clip_planes_2d == 0)
	{
		for (size_t t = 0; t < vm.num_visible_triangles; ++t)
		{
			sqw_triangle_point pts[3];
			for (int k = 0; k < 3; ++k)
			{
				const ModelVertex3d& v = vertices_combined[triangles[t][k]];
				pts[k] = to_point(truncate(v.pos) / v.pos.z, v.tc, v.light);
			}
			emit(pts);
		}
		return;
	}
	ModelVertex3d vertices_clipped0[MAX_VERTICES], vertices_clipped1[MAX_VERTICES];
	ModelVertex2d vertices_projected0[MAX_VERTICES], vertices_projected1[MAX_VERTICES];
	sqw_triangle_point vertices_fixed[MAX_VERTICES];
	for (size_t t = 0; t < vm.num_visible_triangles; ++t)
	{
		ModelVertex3d *vc_src = vertices_clipped0, *vc_dst = vertices_clipped1;
		for (int k = 0; k < 3; ++k)
			vc_src[k] = vertices_combined[triangles[t][k]];
		size_t num_vertices = 3;
		bool rejected = false;
		for (size_t pi = 0; pi < num_clip_planes_3d; ++pi)
		{
			num_vertices = clip_3d_model_polygon_by_plane(vc_src, num_vertices, clip_planes_3d[pi], vc_dst, MAX_VERTICES);
			if (num_vertices < 3)
			{
				rejected = true;
				break;
			}
			std::swap(vc_src, vc_dst);
		}
		if (rejected)
			continue;
		ModelVertex2d *vp_src = vertices_projected0, *vp_dst = vertices_projected1;
		for (size_t k = 0; k < num_vertices; ++k)
		{
			vp_src[k].pos = truncate(vc_src[k].pos) / vc_src[k].pos.z;
			vp_src[k].tc = vc_src[k].tc;
			std::memcpy(vp_src[k].light, vc_src[k].light, 12);
		}
		for (size_t pi = 0; pi < num_clip_planes_2d; ++pi)
		{
			num_vertices = clip_2d_model_polygon(vp_src, num_vertices, clip_planes_2d[pi], vp_dst, MAX_VERTICES);
			if (num_vertices < 3)
			{
				rejected = true;
				break;
			}
			std::swap(vp_src, vp_dst);
		}
		if (rejected)
			continue;
		for (size_t k = 0; k < num_vertices; ++k)
			vertices_fixed[k] = to_point(vp_src[k].pos, vp_src[k].tc, vp_src[k].light);
		for (size_t k = 0; k + 2 < num_vertices; ++k)
		{
			const sqw_triangle_point pts[3] = {vertices_fixed[0], vertices_fixed[k + 1], vertices_fixed[k + 2]};
			emit(pts);
		}
	}
}

} // namespace sqwh
