// Host-side mirror of the renderer front end (see renderer.hpp).  Each function cites the reference code it follows.
#include "renderer.hpp"
#include <atomic>
#include <thread>
#include <algorithm>
#include <chrono>
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
	leaf_submodels_.resize(map_.leafs.size());
	// InlineModelsIndex::new (inline_models_index.rs:162-186): bounding boxes of the submodels, extended a little; polygons know
	// their submodel (DrawPolygonParent::Submodel, partial_renderer.rs:127-133)
	submodels_bbox_.resize(map_.submodels.size());
	submodel_leafs_.resize(map_.submodels.size());
	submodels_info_.resize(map_.submodels.size());
	submodel_entities_.resize(map_.submodels.size());
	for (size_t i = 0; i < map_.submodels.size(); ++i)
	{
		const Submodel& submodel = map_.submodels[i];
		const float inf = 1e24f;
		BBox bbox = {{inf, inf, inf}, {-inf, -inf, -inf}};
		for (uint32_t p = submodel.first_polygon; p < submodel.first_polygon + submodel.num_polygons; ++p)
		{
			polygons_data_[p].parent_submodel = (uint32_t)i;
			for (uint32_t v = 0; v < map_.polygons[p].num_vertices; ++v)
				bbox.extend_with_point(map_.vertices[map_.polygons[p].first_vertex + v]);
		}
		const Vec3f extend_vec = {0.25f, 0.25f, 0.25f};
		bbox.max = bbox.max + extend_vec;
		bbox.min = bbox.min - extend_vec;
		submodels_bbox_[i] = bbox;
		submodel_entities_[i].position = bbox.get_center();
	}
	leaf_sprites_.resize(map_.leafs.size());
	leaf_portals_.resize(map_.leafs.size());
	leaf_decals_.resize(map_.leafs.size());
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

void VisibilityCalculator::update_visibility_with_start_leafs(const CameraMatrices& camera_matrices, const ClippingPolygon& frame_bounds, const uint32_t* start_leafs, size_t n) // :101-113
{
	camera_matrices_ = camera_matrices;
	++visibility_frame_;
	mark_reachable_leafs_iterative(start_leafs, n, frame_bounds);
	is_inside_leaf_volume_ = true; // "Can't properly determine this."
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
	// submodels atop of the static geometry, with depth test (depth_renderer.rs:64-66, :130-200)
	for (size_t submodel_index = 0; submodel_index < map_.submodels.size(); ++submodel_index)
	{
		Mat4f model_matrix;
		if (!get_submodel_matrix(submodel_index, model_matrix))
			continue;
		bool have_bounds = false;
		ClippingPolygon bounds;
		for (uint32_t leaf_index : submodel_leafs_[submodel_index])
		{
			ClippingPolygon leaf_bounds;
			if (shadow_visibility_.leaf_bounds(leaf_index, leaf_bounds))
			{
				if (have_bounds)
					bounds.extend(leaf_bounds);
				else
				{
					bounds = leaf_bounds;
					have_bounds = true;
				}
			}
		}
		if (!have_bounds)
			continue;
		Vec3f clip_planes[8];
		bounds.get_clip_planes(clip_planes);
		Mat4f model_matrix_inverse = Mat4f::identity();
		model_matrix.transpose().invert(model_matrix_inverse);
		CameraMatrices submodel_camera_matrices;
		submodel_camera_matrices.view_matrix = depth_matrices.view_matrix * model_matrix;
		submodel_camera_matrices.planes_matrix = depth_matrices.planes_matrix * model_matrix_inverse;
		submodel_camera_matrices.position = depth_matrices.position;
		const Submodel& submodel = map_.submodels[submodel_index];
		for (uint32_t p = submodel.first_polygon; p < submodel.first_polygon + submodel.num_polygons; ++p)
			depth_draw_polygon(submodel_camera_matrices, clip_planes, p, true);
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
	prepare_frame_with_matrices(width, height, build_view_matrix(camera.position, camera.azimuth, camera.elevation, fov, (float)width, (float)height), false, nullptr, 0);
}

void Renderer::prepare_frame_with_matrices(uint32_t width, uint32_t height, const CameraMatrices& camera_matrices, bool is_third_person_view, const uint32_t* start_leafs, size_t num_start_leafs)
{
	width_ = width;
	height_ = height;
	camera_matrices_ = camera_matrices;
	is_third_person_view_ = is_third_person_view;
	mip_bias_ = config_.textures_mip_bias; // update_mip_bias with dynamic_mip_bias = false (partial_renderer.rs:3786-3804)
	++current_frame_;
	stats_ = RendererDebugStats{};
	const ClippingPolygon frame_bounds = ClippingPolygon::from_box(0.0f, 0.0f, (float)width, (float)height); // partial_renderer.rs:173-190
	if (start_leafs)
		visibility_.update_visibility_with_start_leafs(camera_matrices_, frame_bounds, start_leafs, num_start_leafs);
	else
		visibility_.update_visibility(camera_matrices_, frame_bounds);
	position_dynamic_lights();   // renderer.rs:88 (DynamicObjectsIndex)
	prepare_dynamic_lights();    // partial_renderer.rs:192
	position_submodels();        // renderer.rs:82 (inline_models_index.position_models)
	build_shadow_maps();         // :194
	prepare_submodels();         // :196
	position_models();              // renderer.rs:84 (dynamic_models_index.position_models)
	prepare_dynamic_models();       // partial_renderer.rs:198-205
	build_dynamic_models_buffers(); // :206
	position_decals();              // renderer.rs:85
	prepare_decals();               // partial_renderer.rs:209
	position_sprites();             // renderer.rs:86
	prepare_sprites();              // partial_renderer.rs:211
	{
		const auto t0 = std::chrono::steady_clock::now();
		prepare_polygons_surfaces(); // :214
		polygons_prepare_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	build_surface_jobs();        // :221
	position_portals();              // renderer.rs:90
	prepare_view_portals_textures(); // partial_renderer.rs:230 (the nested views themselves: prepare_portal_view)
}

void Renderer::prepare_polygons_surfaces() // partial_renderer.rs:1327-1449 (leaf polygons; submodels are not part of the synthetic maps)
{
	current_frame_visible_polygons_.clear();
	polygon_candidates_.clear();
	polygon_clip_planes_.clear();
	polygon_clip_set_leafs_.clear();
	polygon_is_sky_candidate_.clear();
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
		const uint32_t clip_set = (uint32_t)(polygon_clip_planes_.size() / 8);
		polygon_clip_set_leafs_.push_back((uint32_t)i);
		polygon_clip_planes_.insert(polygon_clip_planes_.end(), clip_planes, clip_planes + 8);
		for (uint32_t p = leaf.first_polygon; p < leaf.first_polygon + leaf.num_polygons; ++p)
		{
			polygon_candidates_.push_back(PolygonCandidate{p, clip_set});
			polygon_is_sky_candidate_.push_back(prepare_polygon_surface(camera_matrices_, clip_planes, p) == 2 ? 1 : 0);
		}
	}
	// surfaces of the submodels located in visible leafs (:1362-1447)
	for (size_t index = 0; index < map_.submodels.size(); ++index)
	{
		const SubmodelInfo& info = submodels_info_[index];
		if (!info.has_matrices)
			continue;
		bool have_bounds = false;
		ClippingPolygon bounds;
		for (uint32_t leaf_index : submodel_leafs_[index])
		{
			ClippingPolygon leaf_bounds;
			if (visibility_.leaf_bounds(leaf_index, leaf_bounds))
			{
				if (have_bounds)
					bounds.extend(leaf_bounds);
				else
				{
					bounds = leaf_bounds;
					have_bounds = true;
				}
			}
		}
		if (!have_bounds)
			continue;
		Vec3f clip_planes[8];
		bounds.get_clip_planes(clip_planes);
		const Submodel& submodel = map_.submodels[index];
		for (uint32_t polygon_index = submodel.first_polygon; polygon_index < submodel.first_polygon + submodel.num_polygons; ++polygon_index)
		{
			prepare_polygon_surface(info.camera_matrices, clip_planes, polygon_index);
			// visible: basis vectors in world space under the submodel's transformation, for specular and dynamic lighting
			DrawPolygonData& polygon_data = polygons_data_[polygon_index];
			if (polygon_data.visible_frame == current_frame_)
			{
				const Polygon& polygon = map_.polygons[polygon_index];
				const Vec4f plane_transformed_vec = info.world_planes_matrix * extend(polygon.plane.vec, -polygon.plane.dist);
				const Vec4f t0 = info.world_planes_matrix * extend(polygon.tex_coord_equation[0].vec, polygon.tex_coord_equation[0].dist);
				const Vec4f t1 = info.world_planes_matrix * extend(polygon.tex_coord_equation[1].vec, polygon.tex_coord_equation[1].dist);
				const Plane tc[2] = {Plane{truncate(t0), t0.w}, Plane{truncate(t1), t1.w}};
				basis_vecs_from_plane_and_tc(Plane{truncate(plane_transformed_vec), -plane_transformed_vec.w}, tc, polygon_data);
			}
		}
	}
	stats_.num_visible_polygons = (uint32_t)current_frame_visible_polygons_.size();
	stats_.num_surfaces_pixels = num_visible_surfaces_pixels_;
}

// returns 0: culled or clipped away, 1: visible (surface rectangle recorded), 2: sky polygon (only widened the sky window)
uint32_t Renderer::prepare_polygon_surface(const CameraMatrices& camera_matrices_, const Vec3f clip_planes[8], size_t polygon_index) // partial_renderer.rs:1464-1638
{
	DrawPolygonData& polygon_data = polygons_data_[polygon_index];
	const Polygon& polygon = map_.polygons[polygon_index];

	const Vec4f plane_transformed = camera_matrices_.planes_matrix * extend(polygon.plane.vec, -polygon.plane.dist);
	if (plane_transformed.w <= 0.0f) // cull back faces
		return 0;
	const Material& material = scene_.materials[polygon.texture];
	if (!material.draw)
		return 0;

	const Vec3f* polygon_vertices = &map_.vertices[polygon.first_vertex];
	Vec3f* polygon_vertices_transformed = &vertices_transformed_[polygon.first_vertex];
	for (uint32_t i = 0; i < polygon.num_vertices; ++i)
		polygon_vertices_transformed[i] = view_matrix_transform_vertex(camera_matrices_.view_matrix, polygon_vertices[i]);

	Vec2f vertices_2d[MAX_VERTICES];
	const size_t vertex_count = project_and_clip_polygon(clip_planes, polygon_vertices_transformed, polygon.num_vertices, vertices_2d);
	if (vertex_count < 3)
		return 0;

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
		return 2;
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
	return 1;
}

void Renderer::export_prepared_polygons(PreparedPolygonExport* out, PolygonsFrameExport* frame) const
{
	if (out)
		for (size_t i = 0; i < polygon_candidates_.size(); ++i)
		{
			PreparedPolygonExport& e = out[i];
			std::memset(&e, 0, sizeof(e));
			const DrawPolygonData& d = polygons_data_[polygon_candidates_[i].polygon];
			// (leaf polygons only: a submodel polygon prepared later in the frame is not a candidate)
			if (d.visible_frame == current_frame_ && d.parent_submodel == 0xFFFFFFFFu)
			{
				e.flags = 1;
				e.mip = d.mip;
				std::memcpy(e.depth_equation, d.depth_equation, sizeof(e.depth_equation));
				std::memcpy(e.tex_coord_equation, d.tex_coord_equation, sizeof(e.tex_coord_equation));
				e.surface_tc_min[0] = d.surface_tc_min[0];
				e.surface_tc_min[1] = d.surface_tc_min[1];
				e.surface_size[0] = d.surface_size[0];
				e.surface_size[1] = d.surface_size[1];
				e.surface_pixels_offset = (uint32_t)d.surface_pixels_offset;
			}
			else if (polygon_is_sky_candidate_[i])
				e.flags = 2;
		}
	if (frame)
	{
		std::memset(frame, 0, sizeof(*frame));
		std::memcpy(frame->planes_matrix, &camera_matrices_.planes_matrix, 64);
		std::memcpy(frame->view_matrix, &camera_matrices_.view_matrix, 64);
		frame->mip_bias = mip_bias_;
		frame->has_sky = has_current_sky_ ? 1u : 0u;
		frame->sky_material = current_sky_material_;
		if (has_current_sky_)
			std::memcpy(frame->sky_bounds, &current_sky_bounds_, 32);
		frame->num_visible = 0;
		frame->num_surfaces_pixels = 0;
		for (size_t i = 0; i < polygon_candidates_.size(); ++i)
		{
			const DrawPolygonData& d = polygons_data_[polygon_candidates_[i].polygon];
			if (d.visible_frame == current_frame_ && d.parent_submodel == 0xFFFFFFFFu)
			{
				++frame->num_visible;
				frame->num_surfaces_pixels += (uint64_t)d.surface_size[0] * d.surface_size[1];
			}
		}
	}
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
			const std::vector<uint32_t>& lights_list =
				polygon_data.parent_submodel < submodels_info_.size() ? submodels_info_[polygon_data.parent_submodel].dynamic_lights : leaf_lights_[polygon_data.parent_leaf];
			for (uint32_t light_index : lights_list)
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
	polygon_draw_order_.clear();
	const ClippingPolygon viewport_clipping_polygon = ClippingPolygon::from_box(0.0f, 0.0f, (float)width_, (float)height_);
	draw_skybox(viewport_clipping_polygon, color_bits); // perform_rasterization_for_viewport_part (:1296): sky first
	draw_tree(viewport_clipping_polygon);
	draw_view_models(viewport_clipping_polygon);
	stats_.num_draw_commands = (uint32_t)cmds_.size();
}

// partial_renderer.rs:2284-2309: models flagged VIEW_MODEL (the player's weapon) are drawn after everything else, over the whole
// viewport and with no leaf clip planes — they belong to no leaf (dynamic_objects_index.rs:61 skips them).
void Renderer::draw_view_models(const ClippingPolygon& viewport_clipping_polygon)
{
	for (size_t dynamic_model_index = 0; dynamic_model_index < frame_models_.size(); ++dynamic_model_index)
	{
		if (!(frame_models_[dynamic_model_index].flags & MODEL_FLAG_VIEW_MODEL))
			continue;
		const DynamicModelInfo entry = dynamic_model_to_dynamic_meshes_index_[dynamic_model_index];
		for (uint32_t m = entry.first_visible_mesh; m < entry.first_visible_mesh + entry.num_visible_meshes; ++m)
			draw_mesh(viewport_clipping_polygon, nullptr, 0, m);
	}
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
		polygon_draw_order_.push_back(polygon_index);
		draw_polygon(
			clip_planes, &vertices_transformed_[polygon.first_vertex], polygon.num_vertices, polygon_data.depth_equation, polygon_data.tex_coord_equation,
			polygon_data.surface_size, SQW_TEX_SURFACE, (uint32_t)polygon_data.surface_pixels_offset, scene_.materials[polygon.texture].blending_mode);
	}
	// decals after all leaf polygons (:2567-2582)
	const std::vector<uint32_t>& leaf_decals = leaf_decals_[leaf_index];
	if (!leaf_decals.empty())
		for (uint32_t polygon_index = leaf.first_polygon; polygon_index < leaf.first_polygon + leaf.num_polygons; ++polygon_index)
			draw_polygon_decals(clip_planes, polygon_index, leaf_decals);
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

inline int32_t f32_to_i32_saturating(float v) // Rust `as i32`
{
	if (!(v == v))
		return 0;
	if (v >= 2147483648.0f)
		return INT32_MAX;
	if (v <= -2147483648.0f)
		return INT32_MIN;
	return (int32_t)v;
}

inline void get_light_cube_light_host(const float light_cube[6][3], Vec3f n, float total_light[3]) // light.rs:296-348
{
	total_light[0] = total_light[1] = total_light[2] = 0.0f;
	const float comp[3] = {n.x, n.y, n.z};
	for (int axis = 0; axis < 3; ++axis)
	{
		if (comp[axis] <= 0.0f)
			for (int i = 0; i < 3; ++i)
				total_light[i] += light_cube[axis * 2][i] * (-comp[axis]);
		else
			for (int i = 0; i < 3; ++i)
				total_light[i] += light_cube[axis * 2 + 1][i] * comp[axis];
	}
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
		if ((model.flags & MODEL_FLAG_ONLY_THIRD_PERSON_VIEW) && !is_third_person_view_)
			continue;
		if ((model.flags & MODEL_FLAG_VIEW_MODEL) && is_third_person_view_)
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

// Trilinear fetch from the map's light grid with renormalisation over missing / all-zero samples.
LightGridElement Renderer::fetch_light_from_grid(Vec3f pos) const // triangle_models_rendering.rs:574-699
{
	const LightGridElement zero_light;
	const LightGridHeader& h = map_.light_grid_header;
	if (h.grid_size[0] == 0 || h.grid_size[1] == 0 || h.grid_size[2] == 0 || h.grid_cell_size[0] == 0.0f || h.grid_cell_size[1] == 0.0f || h.grid_cell_size[2] == 0.0f ||
		map_.light_grid_samples.empty() || map_.light_grid_columns.empty())
		return zero_light;
	const float grid_pos[3] = {(pos.x - h.grid_start[0]) / h.grid_cell_size[0], (pos.y - h.grid_start[1]) / h.grid_cell_size[1], (pos.z - h.grid_start[2]) / h.grid_cell_size[2]};
	const int32_t grid_pos_i[3] = {f32_to_i32_saturating(std::floor(grid_pos[0])), f32_to_i32_saturating(std::floor(grid_pos[1])), f32_to_i32_saturating(std::floor(grid_pos[2]))};
	auto is_zero = [](const LightGridElement& e) {
		for (int s = 0; s < 6; ++s)
			for (int c = 0; c < 3; ++c)
				if (e.light_cube[s][c] != 0.0f)
					return false;
		return e.light_direction_vector_scaled.x == 0.0f && e.light_direction_vector_scaled.y == 0.0f && e.light_direction_vector_scaled.z == 0.0f &&
			e.directional_light_color[0] == 0.0f && e.directional_light_color[1] == 0.0f && e.directional_light_color[2] == 0.0f;
	};
	LightGridElement total_light;
	float total_factor = 0.0f;
	for (int dx = 0; dx <= 1; ++dx)
	{
		const int64_t x = (int64_t)grid_pos_i[0] + dx;
		if (x < 0 || x >= (int64_t)h.grid_size[0])
			continue;
		const float factor_x = 1.0f - std::fabs(grid_pos[0] - (float)x);
		for (int dy = 0; dy <= 1; ++dy)
		{
			const int64_t y = (int64_t)grid_pos_i[1] + dy;
			if (y < 0 || y >= (int64_t)h.grid_size[1])
				continue;
			const LightGridColumn& column = map_.light_grid_columns[(size_t)x + (size_t)y * h.grid_size[0]];
			if (column.num_samples == 0)
				continue;
			const float factor_y = 1.0f - std::fabs(grid_pos[1] - (float)y);
			for (int dz = 0; dz <= 1; ++dz)
			{
				const int64_t z = (int64_t)grid_pos_i[2] + dz;
				if (z < 0 || z >= (int64_t)h.grid_size[2])
					continue;
				if ((uint64_t)z < column.start_z || (uint64_t)z >= (uint64_t)column.start_z + column.num_samples)
					continue;
				const size_t address = (size_t)column.first_sample + (size_t)(z - column.start_z);
				if (address >= map_.light_grid_samples.size())
					continue;
				const LightGridElement& sample_value = map_.light_grid_samples[address];
				if (is_zero(sample_value))
					continue;
				const float factor_z = 1.0f - std::fabs(grid_pos[2] - (float)z);
				const float cur_sample_factor = factor_x * factor_y * factor_z;
				for (int i = 0; i < 3; ++i)
				{
					for (int cube_side = 0; cube_side < 6; ++cube_side)
						total_light.light_cube[cube_side][i] += sample_value.light_cube[cube_side][i] * cur_sample_factor;
					total_light.directional_light_color[i] += sample_value.directional_light_color[i] * cur_sample_factor;
				}
				total_light.light_direction_vector_scaled = total_light.light_direction_vector_scaled + sample_value.light_direction_vector_scaled * cur_sample_factor;
				total_factor += cur_sample_factor;
			}
		}
	}
	if (total_factor <= 0.0f)
		return zero_light;
	if (total_factor < 0.995f)
	{
		const float inv_total_factor = 1.0f / total_factor;
		for (int i = 0; i < 3; ++i)
		{
			for (int cube_side = 0; cube_side < 6; ++cube_side)
				total_light.light_cube[cube_side][i] *= inv_total_factor;
			total_light.directional_light_color[i] *= inv_total_factor;
		}
		total_light.light_direction_vector_scaled = total_light.light_direction_vector_scaled * inv_total_factor;
	}
	return total_light;
}

// get_model_light (triangle_models_rendering.rs:426-572).  Static part: the light grid (get_model_static_light, :539-572);
// dynamic part: the eight bbox sample positions against every dynamic light WITHOUT a shadow map.  Lights with cube /
// projector shadow maps need get_light_shadow_factor on shadow maps that live on the device here (rendered by
// depth_kernels.cu) — they do not light models in this mirror.
ModelLightData Renderer::get_model_light(const ModelEntity& model, const Mat4f& model_matrix) const
{
	LightGridElement static_light;
	if (model.lighting == MODEL_LIGHTING_DEFAULT)
		static_light = fetch_light_from_grid(model.position);
	else if (model.lighting == MODEL_LIGHTING_CONSTANT)
	{
		for (auto& side : static_light.light_cube)
			std::memcpy(side, model.light, 12);
	}
	else
	{
		for (auto& side : static_light.light_cube)
			std::memcpy(side, model.light, 12);
		if (model.grid_light_scale > 0.0f)
		{
			const LightGridElement grid_light = fetch_light_from_grid(model.light_position);
			for (int i = 0; i < 6; ++i)
				for (int j = 0; j < 3; ++j)
					static_light.light_cube[i][j] += model.grid_light_scale * grid_light.light_cube[i][j];
			static_light.light_direction_vector_scaled = grid_light.light_direction_vector_scaled * model.grid_light_scale;
			std::memcpy(static_light.directional_light_color, grid_light.directional_light_color, 12);
		}
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

size_t Renderer::export_dynamic_meshes(DynamicMeshExport* out, size_t capacity) const
{
	size_t n = 0;
	for (const VisibleDynamicMeshInfo& vm : visible_dynamic_meshes_list_)
	{
		if (n >= capacity)
			break;
		const ModelEntity& model = frame_models_[vm.entity_index];
		const TriangleModel& triangle_model = scene_.models[model.model];
		const TriangleModelMesh& mesh = triangle_model.meshes[vm.mesh_index];
		const HostImage& texture = scene_.images[model.texture_images[vm.mip]];
		DynamicMeshExport& e = out[n++];
		e.model = model.model;
		e.mesh = vm.mesh_index;
		e.frames[0] = model.frames[0];
		e.frames[1] = model.frames[1];
		e.lerp = model.lerp;
		std::memcpy(e.model_matrix, &vm.model_matrix, 64);
		std::memcpy(e.model_view_matrix, &vm.camera_matrices.view_matrix, 64);
		e.tc_scale[0] = (float)texture.width;
		e.tc_scale[1] = (float)texture.height;
		e.tc_shift[0] = triangle_model.tc_shift.x;
		e.tc_shift[1] = triangle_model.tc_shift.y;
		const ModelLightData light = get_model_light(model, vm.model_matrix);
		std::memcpy(e.light_cube, light.light_cube, sizeof(e.light_cube));
		for (int d = 0; d < 2; ++d)
		{
			e.light_vector_scaled[d][0] = light.vector_scaled[d].x;
			e.light_vector_scaled[d][1] = light.vector_scaled[d].y;
			e.light_vector_scaled[d][2] = light.vector_scaled[d].z;
			std::memcpy(e.light_color[d], light.color[d], 12);
		}
		e.vertices_offset = (uint32_t)vm.vertices_offset;
		e.num_vertices = (uint32_t)mesh.num_vertices();
		e.triangles_offset = (uint32_t)vm.triangles_offset;
		e.num_triangles = (uint32_t)mesh.triangles.size();
		e.num_visible_triangles = (uint32_t)vm.num_visible_triangles;
	}
	return n;
}

void Renderer::build_dynamic_models_buffers() // partial_renderer.rs:1122-1190
{
	// every mesh writes its own region of the vertex / triangle buffers: the reference runs this loop with rayon's par_iter_mut (:1181-1189)
	auto process_mesh = [this](VisibleDynamicMeshInfo& vm) {
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
	};
	const size_t num_meshes = visible_dynamic_meshes_list_.size();
	const size_t num_threads = std::min<size_t>(std::min<size_t>(num_meshes, std::max(1u, std::thread::hardware_concurrency())), 8);
	if (num_threads <= 1)
	{
		for (VisibleDynamicMeshInfo& vm : visible_dynamic_meshes_list_)
			process_mesh(vm);
		return;
	}
	std::atomic<size_t> next{0};
	auto worker = [&]() {
		for (size_t i = next.fetch_add(1); i < num_meshes; i = next.fetch_add(1))
			process_mesh(visible_dynamic_meshes_list_[i]);
	};
	std::vector<std::thread> threads;
	for (size_t t = 1; t < num_threads; ++t)
		threads.emplace_back(worker);
	worker();
	for (std::thread& t : threads)
		t.join();
}

// The tail of draw_leaf (partial_renderer.rs:2584-2790) for the object kinds of this mirror (triangle models and sprites).
void Renderer::draw_leaf_objects(const ClippingPolygon& bounds, uint32_t leaf_index)
{
	const std::vector<uint32_t>& leaf_submodels = leaf_submodels_[leaf_index];
	const std::vector<uint32_t>& leaf_dynamic_models = leaf_models_[leaf_index];
	const std::vector<uint32_t>& leaf_sprites = leaf_sprites_[leaf_index];
	const std::vector<uint32_t>& leaf_portals = leaf_portals_[leaf_index];
	const size_t total_objects = leaf_submodels.size() + leaf_dynamic_models.size() + leaf_sprites.size() + leaf_portals.size();
	if (total_objects == 0)
		return;
	Vec3f clip_planes[8];
	bounds.get_clip_planes(clip_planes);
	Plane leaf_clip_planes[MAX_LEAF_CLIP_PLANES];
	const std::vector<Plane>& world = leafs_planes_[leaf_index];
	const size_t num_planes = std::min(MAX_LEAF_CLIP_PLANES, world.size());
	for (size_t i = 0; i < num_planes; ++i)
	{
		const Vec4f t = camera_matrices_.planes_matrix * extend(world[i].vec, -world[i].dist);
		leaf_clip_planes[i] = Plane{truncate(t), -t.w};
	}
	if (total_objects == 1)
	{
		if (leaf_submodels.size() == 1)
			draw_submodel_in_leaf(clip_planes, leaf_clip_planes, num_planes, leaf_index, leaf_submodels[0]);
		else if (leaf_dynamic_models.size() == 1)
		{
			const DynamicModelInfo entry = dynamic_model_to_dynamic_meshes_index_[leaf_dynamic_models[0]];
			for (uint32_t m = entry.first_visible_mesh; m < entry.first_visible_mesh + entry.num_visible_meshes; ++m)
				draw_mesh(bounds, leaf_clip_planes, num_planes, m);
		}
		else if (leaf_sprites.size() == 1)
			draw_sprite(clip_planes, leaf_clip_planes, num_planes, leaf_sprites[0]);
		else
			draw_view_portal(clip_planes, leaf_clip_planes, num_planes, leaf_portals[0]);
		return;
	}
	const uint32_t SUBMODEL_INDEX_ADD = 65536, DYNAMIC_MESH_INDEX_ADD = SUBMODEL_INDEX_ADD + 65536, SPRITE_INDEX_ADD = DYNAMIC_MESH_INDEX_ADD + 65536,
				   PORTAL_INDEX_ADD = SPRITE_INDEX_ADD + 65536;
	objects_sorter_.clear();
	for (uint32_t model_index : leaf_submodels)
		if (submodels_info_[model_index].has_matrices)
		{
			// get_model_bbox_for_ordering (inline_models_index.rs:68-81): slightly reduced, to order touching models
			BBox bbox = submodels_bbox_[model_index];
			const Vec3f center = bbox.get_center();
			const float eps = 1.0f;
			bbox.max = {std::fmax(bbox.max.x - eps, center.x), std::fmax(bbox.max.y - eps, center.y), std::fmax(bbox.max.z - eps, center.z)};
			bbox.min = {std::fmin(bbox.min.x + eps, center.x), std::fmin(bbox.min.y + eps, center.y), std::fmin(bbox.min.z + eps, center.z)};
			objects_sorter_.add_object(model_index + SUBMODEL_INDEX_ADD, project_bbox(bbox, submodels_info_[model_index].camera_matrices));
		}
	for (uint32_t dynamic_model_index : leaf_dynamic_models)
	{
		const ModelEntity& model = frame_models_[dynamic_model_index];
		const BBox bbox = model.has_ordering_custom_bbox ? model.ordering_custom_bbox : get_current_triangle_model_bbox(scene_.models[model.model], model.frames);
		const DynamicModelInfo entry = dynamic_model_to_dynamic_meshes_index_[dynamic_model_index];
		for (uint32_t m = entry.first_visible_mesh; m < entry.first_visible_mesh + entry.num_visible_meshes; ++m)
			objects_sorter_.add_object(m + DYNAMIC_MESH_INDEX_ADD, project_bbox(bbox, visible_dynamic_meshes_list_[m].camera_matrices));
	}
	for (uint32_t sprite_index : leaf_sprites)
	{
		const Sprite& sprite = frame_sprites_[sprite_index];
		const Vec3f extend_vec = {sprite.radius, sprite.radius, sprite.radius};
		objects_sorter_.add_object(sprite_index + SPRITE_INDEX_ADD, project_bbox(BBox::from_min_max(sprite.position - extend_vec, sprite.position + extend_vec), camera_matrices_));
	}
	for (uint32_t portal_index : leaf_portals)
	{
		const ViewPortal& portal = frame_portals_[portal_index];
		BBox bbox = BBox::from_point(portal.vertices[0]);
		for (size_t v = 1; v < portal.vertices.size(); ++v)
			bbox.extend_with_point(portal.vertices[v]);
		objects_sorter_.add_object(portal_index + PORTAL_INDEX_ADD, project_bbox(bbox, camera_matrices_));
	}
	objects_sorter_.sort_objects();
	for (size_t i = 0; i < objects_sorter_.num_objects; ++i)
	{
		const uint32_t object_index = objects_sorter_.objects[i].first;
		if (object_index >= PORTAL_INDEX_ADD)
			draw_view_portal(clip_planes, leaf_clip_planes, num_planes, object_index - PORTAL_INDEX_ADD);
		else if (object_index >= SPRITE_INDEX_ADD)
			draw_sprite(clip_planes, leaf_clip_planes, num_planes, object_index - SPRITE_INDEX_ADD);
		else if (object_index >= DYNAMIC_MESH_INDEX_ADD)
			draw_mesh(bounds, leaf_clip_planes, num_planes, object_index - DYNAMIC_MESH_INDEX_ADD);
		else
			draw_submodel_in_leaf(clip_planes, leaf_clip_planes, num_planes, leaf_index, object_index - SUBMODEL_INDEX_ADD);
	}
}

// common/clipping.rs: general-plane clipper shared by submodel polygons and decals
namespace
{
Vec3f get_line_plane_intersection(Vec3f v0, Vec3f v1, const Plane& plane) // clipping.rs:187-195
{
	const float dist0 = dot(v0, plane.vec) - plane.dist;
	const float dist1 = dot(v1, plane.vec) - plane.dist;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum;
	const float k1 = dist1 / dist_sum;
	return v0 * k1 - v1 * k0;
}

size_t clip_3d_polygon_by_plane(const Vec3f* polygon, size_t n, const Plane& plane, Vec3f* out, size_t out_cap) // clipping.rs:53-103
{
	const Vec3f* prev_v = &polygon[n - 1];
	float prev_dist = dot(plane.vec, *prev_v);
	size_t count = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const Vec3f* v = &polygon[i];
		const float dist = dot(plane.vec, *v);
		if (dist > plane.dist)
		{
			if (prev_dist < plane.dist)
			{
				out[count++] = get_line_plane_intersection(*prev_v, *v, plane);
				if (count == out_cap)
					break;
			}
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (dist == plane.dist)
		{
			out[count++] = *v;
			if (count == out_cap)
				break;
		}
		else if (prev_dist > plane.dist)
		{
			out[count++] = get_line_plane_intersection(*prev_v, *v, plane);
			if (count == out_cap)
				break;
		}
		prev_v = v;
		prev_dist = dist;
	}
	return count;
}
} // namespace

// ---- submodels (inline models) -----------------------------------------------------------------------------------
void Renderer::set_submodel_entities(const SubmodelEntity* e, size_t n)
{
	for (size_t i = 0; i < submodel_entities_.size() && i < n; ++i)
		submodel_entities_[i] = e[i];
}

bool Renderer::get_submodel_matrix(size_t index, Mat4f& out) const // inline_models_index.rs:188-196
{
	const SubmodelEntity& e = submodel_entities_[index];
	if (!e.present)
		return false;
	const Vec3f center = submodels_bbox_[index].get_center();
	out = Mat4f::from_translation(e.position) * mat4_from_quaternion(e.rotation) * Mat4f::from_translation(-center);
	return true;
}

void Renderer::position_submodel_r(uint32_t model_index, const Vec3f bbox_vertices[8], uint32_t node_index) // inline_models_index.rs:128-158
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_submodels_[leaf_index].push_back(model_index);
		submodel_leafs_[model_index].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	size_t vertices_front = 0;
	for (int i = 0; i < 8; ++i)
		if (dot(node.plane.vec, bbox_vertices[i]) > node.plane.dist)
			++vertices_front;
	if (vertices_front > 0)
		position_submodel_r(model_index, bbox_vertices, node.children[0]);
	if (vertices_front < 8)
		position_submodel_r(model_index, bbox_vertices, node.children[1]);
}

void Renderer::position_submodels() // InlineModelsIndex::position_models (every model repositioned every frame; same result)
{
	for (auto& l : leaf_submodels_)
		l.clear();
	for (size_t index = 0; index < map_.submodels.size(); ++index)
	{
		submodel_leafs_[index].clear();
		Mat4f transform_matrix;
		if (!get_submodel_matrix(index, transform_matrix))
			continue;
		Vec3f bbox_vertices[8];
		submodels_bbox_[index].get_corners_vertices(bbox_vertices);
		for (Vec3f& v : bbox_vertices)
			v = truncate(transform_matrix * extend(v, 1.0f));
		position_submodel_r((uint32_t)index, bbox_vertices, map_.root_node());
	}
}

void Renderer::prepare_submodels() // partial_renderer.rs:593-660
{
	for (size_t index = 0; index < submodels_info_.size(); ++index)
	{
		SubmodelInfo& info = submodels_info_[index];
		Mat4f model_matrix;
		info.has_matrices = get_submodel_matrix(index, model_matrix);
		info.dynamic_lights.clear();
		if (!info.has_matrices)
			continue;
		Mat4f model_matrix_inverse = Mat4f::identity();
		model_matrix.transpose().invert(model_matrix_inverse);
		info.world_planes_matrix = model_matrix_inverse;
		info.camera_matrices.view_matrix = camera_matrices_.view_matrix * model_matrix;
		info.camera_matrices.planes_matrix = camera_matrices_.planes_matrix * model_matrix_inverse;
		info.camera_matrices.position = camera_matrices_.position;
		Vec3f v[8];
		submodels_bbox_[index].get_corners_vertices(v);
		for (Vec3f& c : v)
			c = truncate(model_matrix * extend(c, 1.0f));
		BBox bbox_world_space = BBox::from_point(v[0]);
		for (int i = 1; i < 8; ++i)
			bbox_world_space.extend_with_point(v[i]);
		for (size_t li = 0; li < frame_lights_.size(); ++li)
		{
			const DynamicLight& light = frame_lights_[li];
			if (!dynamic_lights_info_[li].visible)
				continue;
			if (light.position.x - light.radius > bbox_world_space.max.x || light.position.x + light.radius < bbox_world_space.min.x ||
				light.position.y - light.radius > bbox_world_space.max.y || light.position.y + light.radius < bbox_world_space.min.y ||
				light.position.z - light.radius > bbox_world_space.max.z || light.position.z + light.radius < bbox_world_space.min.z)
				continue;
			info.dynamic_lights.push_back((uint32_t)li);
		}
	}
}

void Renderer::draw_submodel_in_leaf(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t submodel_index) // :3098-3125
{
	const SubmodelInfo& info = submodels_info_[submodel_index];
	if (!info.has_matrices)
		return;
	const uint32_t root = map_.submodels[submodel_index].root_node;
	if (root < map_.submodels_bsp_nodes.size())
		draw_submodel_bsp_node_r(info.camera_matrices.planes_matrix.row(3), clip_planes, leaf_clip_planes, num_leaf_clip_planes, leaf_index, root);
}

void Renderer::draw_submodel_bsp_node_r(
	Vec4f submodel_planes_matrix_w, const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t node_index) // :3127-3206
{
	const SubmodelBSPNode node = map_.submodels_bsp_nodes[node_index];
	const float plane_transformed_w = dot(submodel_planes_matrix_w, extend(node.plane.vec, -node.plane.dist));
	uint32_t mask = plane_transformed_w >= 0.0f ? 1 : 0;
	if (config_.invert_polygons_order)
		mask ^= 1;
	const uint32_t c_b = node.children[mask], c_f = node.children[mask ^ 1];
	const uint32_t num_nodes = (uint32_t)map_.submodels_bsp_nodes.size();
	if (c_b < num_nodes)
		draw_submodel_bsp_node_r(submodel_planes_matrix_w, clip_planes, leaf_clip_planes, num_leaf_clip_planes, leaf_index, c_b);
	const std::vector<uint32_t>& leaf_decals = leaf_decals_[leaf_index];
	for (uint32_t polygon_index = node.first_polygon; polygon_index < node.first_polygon + node.num_polygons; ++polygon_index)
	{
		draw_submodel_polygon(clip_planes, leaf_clip_planes, num_leaf_clip_planes, leaf_index, polygon_index);
		if (!leaf_decals.empty())
			draw_polygon_decals(clip_planes, polygon_index, leaf_decals);
	}
	if (c_f < num_nodes)
		draw_submodel_bsp_node_r(submodel_planes_matrix_w, clip_planes, leaf_clip_planes, num_leaf_clip_planes, leaf_index, c_f);
}

void Renderer::draw_submodel_polygon(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t polygon_index) // :3208-3288
{
	const DrawPolygonData& polygon_data = polygons_data_[polygon_index];
	if (polygon_data.visible_frame != current_frame_)
		return;
	const Polygon& polygon = map_.polygons[polygon_index];
	Vec3f vertices_clipped[MAX_VERTICES], vertices_temp[MAX_VERTICES];
	size_t vertex_count = std::min((size_t)polygon.num_vertices, MAX_VERTICES);
	for (size_t i = 0; i < vertex_count; ++i)
		vertices_clipped[i] = vertices_transformed_[polygon.first_vertex + i];
	const std::vector<Plane>& leaf_clip_planes_world_space = leafs_planes_[leaf_index];
	for (size_t pi = 0; pi < num_leaf_clip_planes && pi < leaf_clip_planes_world_space.size(); ++pi)
	{
		// polygons lying on a clip plane: decide in world space instead of clipping (:3234-3270)
		const Plane& clip_plane_world_space = leaf_clip_planes_world_space[pi];
		const float almost_one = 1.0f - 1.0f / 256.0f, dist_eps = 1.0f / 4.0f;
		const float world_space_normals_dot = dot(polygon_data.basis_normal, clip_plane_world_space.vec);
		if (world_space_normals_dot >= almost_one)
		{
			if (std::fabs(dot(clip_plane_world_space.vec, polygon_data.basis_start) - clip_plane_world_space.dist) <= dist_eps)
				continue;
		}
		else if (world_space_normals_dot <= -almost_one)
		{
			if (std::fabs(dot(clip_plane_world_space.vec, polygon_data.basis_start) - clip_plane_world_space.dist) <= dist_eps)
				return;
		}
		vertex_count = clip_3d_polygon_by_plane(vertices_clipped, vertex_count, leaf_clip_planes[pi], vertices_temp, MAX_VERTICES);
		if (vertex_count < 3)
			return;
		for (size_t i = 0; i < vertex_count; ++i)
			vertices_clipped[i] = vertices_temp[i];
	}
	draw_polygon(
		clip_planes, vertices_clipped, vertex_count, polygon_data.depth_equation, polygon_data.tex_coord_equation, polygon_data.surface_size, SQW_TEX_SURFACE,
		(uint32_t)polygon_data.surface_pixels_offset, scene_.materials[polygon.texture].blending_mode);
}

// ---- decals --------------------------------------------------------------------------------------------------

void Renderer::position_decal_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index) // dynamic_objects_index.rs:172-209
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_decals_[leaf_index].push_back(id);
		decal_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	size_t vertices_front = 0;
	for (size_t i = 0; i < n; ++i)
		if (dot(node.plane.vec, vertices[i]) > node.plane.dist)
			++vertices_front;
	if (vertices_front > 0)
		position_decal_convex_hull_r(id, vertices, n, node.children[0]);
	if (vertices_front < n)
		position_decal_convex_hull_r(id, vertices, n, node.children[1]);
}

void Renderer::position_decals() // dynamic_objects_index.rs:76-91
{
	for (auto& l : leaf_decals_)
		l.clear();
	decal_leafs_.assign(frame_decals_.size(), {});
	for (size_t index = 0; index < frame_decals_.size(); ++index)
	{
		const Decal& decal = frame_decals_[index];
		const Mat4f matrix = get_object_matrix(decal.position, decal.rotation) * Mat4f::from_nonuniform_scale(decal.scale.x, decal.scale.y, decal.scale.z);
		Vec3f corners[8];
		BBox::from_min_max({-1, -1, -1}, {1, 1, 1}).get_corners_vertices(corners);
		for (Vec3f& c : corners)
			c = truncate(matrix * extend(c, 1.0f));
		position_decal_convex_hull_r((uint32_t)index, corners, 8, map_.root_node());
	}
}

void Renderer::prepare_decals() // partial_renderer.rs:665-737
{
	decals_info_.assign(frame_decals_.size(), DecalInfo{});
	for (size_t i = 0; i < frame_decals_.size(); ++i)
	{
		const Decal& decal = frame_decals_[i];
		const Mat4f decal_matrix = get_object_matrix(decal.position, decal.rotation) * Mat4f::from_nonuniform_scale(decal.scale.x, decal.scale.y, decal.scale.z);
		Mat4f inv_t = Mat4f::identity();
		decal_matrix.transpose().invert(inv_t);
		decals_info_[i].camera_planes_matrix = camera_matrices_.planes_matrix * inv_t;
		LightCube light_cube;
		if (decal.lightmap_light_scale > 0.0f)
		{
			const float min_square_distance = dot(decal.scale, decal.scale) * 0.25f;
			for (size_t li = 0; li < frame_lights_.size(); ++li)
			{
				const DynamicLight& light = frame_lights_[li];
				// lights with shadow maps: see get_model_light (their maps live on the device)
				if (!dynamic_lights_info_[li].visible || light.shadow_type != SQW_SHADOW_NONE)
					continue;
				const Vec3f vec_to_light = light.position - decal.position;
				const float square_dist = std::fmax(dot(vec_to_light, vec_to_light), min_square_distance);
				const float inv_square_dist = 1.0f / square_dist;
				const float light_inv_square_radius = 1.0f / (light.radius * light.radius);
				if (inv_square_dist < light_inv_square_radius)
					continue;
				const float scale = 1.0f * (inv_square_dist - light_inv_square_radius);
				const float light_scaled[3] = {light.color[0] * scale, light.color[1] * scale, light.color[2] * scale};
				light_cube.add_light_sample(vec_to_light, light_scaled);
			}
		}
		decals_info_[i].dynamic_light = light_cube.convert_into_light_grid_sample();
	}
}

void Renderer::get_polygon_lightmap_light(const Polygon& polygon, Vec2f lightmap_coord, float total_light[3]) const // partial_renderer.rs:3807-3875
{
	total_light[0] = total_light[1] = total_light[2] = 0.0f;
	if (polygon.lightmap_data_offset == 0 || map_.lightmaps_data.empty())
		return;
	const float* polygon_lightmap_data = map_.lightmaps_data.data() + (size_t)polygon.lightmap_data_offset * 3;
	const int32_t lightmap_coord_int[2] = {f32_to_i32_saturating(std::floor(lightmap_coord.x)), f32_to_i32_saturating(std::floor(lightmap_coord.y))};
	const int32_t lightmap_size[2] = {(int32_t)get_lightmap_size(polygon.tex_coord_min[0], polygon.tex_coord_max[0]), (int32_t)get_lightmap_size(polygon.tex_coord_min[1], polygon.tex_coord_max[1])};
	float total_factor = 0.0f;
	for (int dy = 0; dy <= 1; ++dy)
	{
		const int64_t y = (int64_t)lightmap_coord_int[1] + dy;
		if (y < 0 || y >= lightmap_size[1])
			continue;
		const float factor_y = 1.0f - std::fabs(lightmap_coord.y - (float)y);
		for (int dx = 0; dx <= 1; ++dx)
		{
			const int64_t x = (int64_t)lightmap_coord_int[0] + dx;
			if (x < 0 || x >= lightmap_size[0])
				continue;
			const float factor_x = 1.0f - std::fabs(lightmap_coord.x - (float)x);
			const float* lightmap_light = polygon_lightmap_data + (size_t)(x + y * lightmap_size[0]) * 3;
			const float cur_sample_factor = factor_x * factor_y;
			for (int i = 0; i < 3; ++i)
				total_light[i] += cur_sample_factor * lightmap_light[i];
			total_factor += cur_sample_factor;
		}
	}
	if (total_factor < 1.0f)
	{
		const float inv_total_factor = 1.0f / total_factor;
		for (int i = 0; i < 3; ++i)
			total_light[i] *= inv_total_factor;
	}
}

struct Renderer::DecalDrawContext
{
	const Decal* decal;
	const Polygon* polygon;
	const float* depth_equation;
	float tc_equation[6];
	float polygon_lightmap_equation[6];
	float polygon_lightmap_coord_shift[2];
	uint32_t image_id;
	float dynamic_light[3];
};

void Renderer::draw_polygon_decals(const Vec3f clip_planes[8], uint32_t polygon_index, const std::vector<uint32_t>& current_decals) // partial_renderer.rs:2794-2957
{
	const DrawPolygonData& polygon_data = polygons_data_[polygon_index];
	if (polygon_data.visible_frame != current_frame_)
		return;
	const Polygon& polygon = map_.polygons[polygon_index];
	if (!scene_.materials[polygon.texture].decals)
		return;
	static const float CUBE_SIDES[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
	static const float DECAL_TEXTURE_BASIS[2][4] = {{0.0f, -0.5f, 0.0f, 0.5f}, {0.0f, 0.0f, -0.5f, 0.5f}};
	const float MAX_ANGLE_COS = 0.1f;
	Vec3f vertices_clipped0[MAX_VERTICES], vertices_clipped1[MAX_VERTICES];
	Vec2f vertices_projected[MAX_VERTICES];
	for (uint32_t decal_index : current_decals)
	{
		const Decal& decal = frame_decals_[decal_index];
		const DecalInfo& decal_info = decals_info_[decal_index];
		const float decal_angle_cos = dot(polygon_data.basis_normal, quaternion_rotate_vector(decal.rotation, Vec3f{-1, 0, 0}));
		if (decal_angle_cos < MAX_ANGLE_COS)
			continue;
		const Mat4f& decal_planes_matrix = decal_info.camera_planes_matrix;
		const Vec3f* src_vertices = &vertices_transformed_[polygon.first_vertex];
		Vec3f *vc_src = vertices_clipped0, *vc_dst = vertices_clipped1;
		size_t num_vertices = std::min((size_t)polygon.num_vertices, MAX_VERTICES);
		for (size_t i = 0; i < num_vertices; ++i)
			vc_src[i] = src_vertices[i];
		bool clipped_away = false;
		for (const float* n : CUBE_SIDES)
		{
			const Vec4f plane_vec = decal_planes_matrix * Vec4f{n[0], n[1], n[2], 1.0f};
			num_vertices = clip_3d_polygon_by_plane(vc_src, num_vertices, Plane{truncate(plane_vec), -plane_vec.w}, vc_dst, MAX_VERTICES);
			if (num_vertices < 3)
			{
				clipped_away = true;
				break;
			}
			std::swap(vc_src, vc_dst);
		}
		if (clipped_away)
			continue;
		num_vertices = project_and_clip_polygon(clip_planes, vc_src, num_vertices, vertices_projected);
		if (num_vertices < 3)
			continue;
		const HostImage& texture = scene_.images[decal.texture_images[0]];
		const Vec4f b0 = {DECAL_TEXTURE_BASIS[0][0], DECAL_TEXTURE_BASIS[0][1], DECAL_TEXTURE_BASIS[0][2], DECAL_TEXTURE_BASIS[0][3]};
		const Vec4f b1 = {DECAL_TEXTURE_BASIS[1][0], DECAL_TEXTURE_BASIS[1][1], DECAL_TEXTURE_BASIS[1][2], DECAL_TEXTURE_BASIS[1][3]};
		const Vec4f tc_basis_transformed[2] = {decal_planes_matrix * (b0 * (float)texture.width), decal_planes_matrix * (b1 * (float)texture.height)};
		const float* depth_equation = polygon_data.depth_equation;
		float tc_equation[6];
		for (int i = 0; i < 2; ++i)
		{
			tc_equation[i] = tc_basis_transformed[i].x + tc_basis_transformed[i].w * depth_equation[0];
			tc_equation[2 + i] = tc_basis_transformed[i].y + tc_basis_transformed[i].w * depth_equation[1];
			tc_equation[4 + i] = tc_basis_transformed[i].z + tc_basis_transformed[i].w * depth_equation[2];
		}
		const uint32_t mip = calculate_mip(vertices_projected, num_vertices, depth_equation, tc_equation, mip_bias_);
		DecalDrawContext ctx;
		ctx.decal = &decal;
		ctx.polygon = &polygon;
		ctx.depth_equation = depth_equation;
		ctx.image_id = decal.texture_images[mip];
		for (int i = 0; i < 6; ++i)
			ctx.tc_equation[i] = tc_equation[i] * (1.0f / (float)(1 << mip));
		const float polygon_lightmap_coord_scale = (float)(1 << polygon_data.mip) / (float)LIGHTMAP_SCALE;
		for (int i = 0; i < 2; ++i)
			ctx.polygon_lightmap_coord_shift[i] = (float)polygon_data.surface_tc_min[i] * polygon_lightmap_coord_scale - (float)(polygon.tex_coord_min[i] >> LIGHTMAP_SCALE_LOG2);
		for (int i = 0; i < 6; ++i)
			ctx.polygon_lightmap_equation[i] = polygon_data.tex_coord_equation[i] * polygon_lightmap_coord_scale;
		get_light_cube_light_host(decal_info.dynamic_light.light_cube, polygon_data.basis_normal, ctx.dynamic_light);
		const float dynamic_light_dir_normal_dot = std::fmax(dot(polygon_data.basis_normal, decal_info.dynamic_light.light_direction_vector_scaled), 0.0f);
		for (int i = 0; i < 3; ++i)
			ctx.dynamic_light[i] += dynamic_light_dir_normal_dot * decal_info.dynamic_light.directional_light_color[i];
		for (size_t t = 0; t + 2 < num_vertices; ++t)
		{
			const Vec2f points[3] = {vertices_projected[0], vertices_projected[t + 1], vertices_projected[t + 2]};
			subdivide_and_draw_decal_triangle(ctx, points, 0);
		}
	}
}

void Renderer::subdivide_and_draw_decal_triangle(const DecalDrawContext& ctx, const Vec2f points[3], size_t recursion_depth) // partial_renderer.rs:2959-3096
{
	const float* depth_equation = ctx.depth_equation;
	Vec2f min_inv_z_point = points[0], max_inv_z_point = points[0];
	float min_inv_z = sample_point(depth_equation, points[0]), max_inv_z = min_inv_z;
	for (int i = 1; i < 3; ++i)
	{
		const float inv_z = sample_point(depth_equation, points[i]);
		if (inv_z < min_inv_z)
		{
			min_inv_z = inv_z;
			min_inv_z_point = points[i];
		}
		if (inv_z > max_inv_z)
		{
			max_inv_z = inv_z;
			max_inv_z_point = points[i];
		}
	}
	const size_t max_recursion_depth = 2;
	if (recursion_depth >= max_recursion_depth || affine_texture_coordinates_interpolation_may_be_used(depth_equation, ctx.tc_equation, min_inv_z_point, max_inv_z_point))
	{
		const Decal& decal = *ctx.decal;
		const HostImage& texture = scene_.images[ctx.image_id];
		const float tc_eps = 1.0f / 32.0f;
		const float min_tc[2] = {tc_eps, tc_eps};
		const float max_tc[2] = {(float)texture.width - tc_eps, (float)texture.height - tc_eps};
		sqw_triangle_point vertices_fixed[3];
		for (int k = 0; k < 3; ++k)
		{
			const Vec2f src = points[k];
			const float z = 1.0f / sample_point(depth_equation, src);
			float light[3] = {decal.light_add[0], decal.light_add[1], decal.light_add[2]};
			if (decal.lightmap_light_scale > 0.0f)
			{
				const float* le = ctx.polygon_lightmap_equation;
				const Vec2f lightmap_coord = {
					z * (le[0] * src.x + le[2] * src.y + le[4]) + ctx.polygon_lightmap_coord_shift[0], z * (le[1] * src.x + le[3] * src.y + le[5]) + ctx.polygon_lightmap_coord_shift[1]};
				float lightmap_light[3];
				get_polygon_lightmap_light(*ctx.polygon, lightmap_coord, lightmap_light);
				for (int i = 0; i < 3; ++i)
				{
					light[i] += lightmap_light[i] * decal.lightmap_light_scale;
					light[i] += ctx.dynamic_light[i];
				}
			}
			const float* te = ctx.tc_equation;
			vertices_fixed[k].x = f32_to_fixed16(src.x);
			vertices_fixed[k].y = f32_to_fixed16(src.y);
			vertices_fixed[k].tc[0] = f32_to_fixed16(std::fmin(std::fmax(z * (te[0] * src.x + te[2] * src.y + te[4]), min_tc[0]), max_tc[0]));
			vertices_fixed[k].tc[1] = f32_to_fixed16(std::fmin(std::fmax(z * (te[1] * src.x + te[3] * src.y + te[5]), min_tc[1]), max_tc[1]));
			for (int i = 0; i < 3; ++i)
				vertices_fixed[k].light[i] = f32_to_fixed16(light[i]);
		}
		sqw_draw_cmd c;
		std::memset(&c, 0, sizeof(c));
		c.kind = SQW_CMD_TRIANGLE;
		c.blend = (uint8_t)decal.blending_mode;
		c.tex_kind = SQW_TEX_IMAGE;
		c.tex_ref = ctx.image_id;
		c.tex_size[0] = (int32_t)texture.width;
		c.tex_size[1] = (int32_t)texture.height;
		c.num_vertices = 3;
		c.vertex_offset = (uint32_t)triangle_points_.size();
		triangle_points_.insert(triangle_points_.end(), vertices_fixed, vertices_fixed + 3);
		cmds_.push_back(c);
		++stats_.num_triangles;
		++stats_.num_decal_triangles;
		return;
	}
	const Vec2f center0 = (points[0] + points[1]) / 2.0f, center1 = (points[1] + points[2]) / 2.0f, center2 = (points[2] + points[0]) / 2.0f;
	const Vec2f triangles[4][3] = {{points[0], center0, center2}, {points[1], center1, center0}, {points[2], center2, center1}, {center0, center1, center2}};
	for (const auto& triangle : triangles)
		subdivide_and_draw_decal_triangle(ctx, triangle, recursion_depth + 1);
}

// ---- portals and mirrors -------------------------------------------------------------------------------------
void Renderer::position_portal_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index) // dynamic_objects_index.rs:172-209
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_portals_[leaf_index].push_back(id);
		portal_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	size_t vertices_front = 0;
	for (size_t i = 0; i < n; ++i)
		if (dot(node.plane.vec, vertices[i]) > node.plane.dist)
			++vertices_front;
	if (vertices_front > 0)
		position_portal_convex_hull_r(id, vertices, n, node.children[0]);
	if (vertices_front < n)
		position_portal_convex_hull_r(id, vertices, n, node.children[1]);
}

void Renderer::position_portals() // dynamic_objects_index.rs:143-157
{
	for (auto& l : leaf_portals_)
		l.clear();
	portal_leafs_.assign(frame_portals_.size(), {});
	for (size_t index = 0; index < frame_portals_.size(); ++index)
		position_portal_convex_hull_r((uint32_t)index, frame_portals_[index].vertices.data(), frame_portals_[index].vertices.size(), map_.root_node());
}

void Renderer::prepare_view_portals_textures() // partial_renderer.rs:1834-1996
{
	portals_info_.assign(frame_portals_.size(), PortalInfo{});
	for (size_t portal_index = 0; portal_index < frame_portals_.size(); ++portal_index)
	{
		const ViewPortal& portal = frame_portals_[portal_index];
		bool have_bounds = false;
		ClippingPolygon bounds;
		for (uint32_t leaf_index : portal_leafs_[portal_index])
		{
			ClippingPolygon leaf_bounds;
			if (visibility_.leaf_bounds(leaf_index, leaf_bounds))
			{
				if (have_bounds)
					bounds.extend(leaf_bounds);
				else
				{
					bounds = leaf_bounds;
					have_bounds = true;
				}
			}
		}
		if (!have_bounds)
			continue; // in invisible leafs
		Vec3f clip_planes[8];
		bounds.get_clip_planes(clip_planes);
		const Vec4f plane_transformed = camera_matrices_.planes_matrix * extend(portal.plane.vec, -portal.plane.dist);
		if (plane_transformed.w <= 0.0f)
			continue;
		Vec3f vertices_3d[MAX_VERTICES];
		const size_t n = std::min(MAX_VERTICES, portal.vertices.size());
		if (n < 3)
			continue;
		for (size_t i = 0; i < n; ++i)
			vertices_3d[i] = view_matrix_transform_vertex(camera_matrices_.view_matrix, portal.vertices[i]);
		Vec2f vertices_2d[MAX_VERTICES];
		const size_t vertex_count = project_and_clip_polygon(clip_planes, vertices_3d, n, vertices_2d);
		if (vertex_count < 3)
			continue; // fully clipped
		const float inf = (float)(1 << 29);
		float tc_min[2] = {inf, inf}, tc_max[2] = {-inf, -inf};
		for (const Vec3f& vertex : portal.vertices)
			for (int i = 0; i < 2; ++i)
			{
				const float tc = dot(portal.tex_coord_equation[i].vec, vertex) + portal.tex_coord_equation[i].dist;
				tc_min[i] = tc < tc_min[i] ? tc : tc_min[i];
				tc_max[i] = tc > tc_max[i] ? tc : tc_max[i];
			}
		for (int i = 0; i < 2; ++i)
		{
			const float tc_reduce_eps = 1.0f / 32.0f;
			tc_min[i] += tc_reduce_eps;
			tc_max[i] -= tc_reduce_eps;
		}
		PortalInfo info;
		const float plane_transformed_w = -plane_transformed.w;
		info.depth_equation[0] = plane_transformed.x / plane_transformed_w;
		info.depth_equation[1] = plane_transformed.y / plane_transformed_w;
		info.depth_equation[2] = plane_transformed.z / plane_transformed_w;
		const Vec4f tc_basis_transformed[2] = {
			camera_matrices_.planes_matrix * extend(portal.tex_coord_equation[0].vec, portal.tex_coord_equation[0].dist),
			camera_matrices_.planes_matrix * extend(portal.tex_coord_equation[1].vec, portal.tex_coord_equation[1].dist)};
		float tc_equation[6];
		for (int i = 0; i < 2; ++i)
		{
			tc_equation[i] = tc_basis_transformed[i].x + tc_basis_transformed[i].w * info.depth_equation[0];
			tc_equation[2 + i] = tc_basis_transformed[i].y + tc_basis_transformed[i].w * info.depth_equation[1];
			tc_equation[4 + i] = tc_basis_transformed[i].z + tc_basis_transformed[i].w * info.depth_equation[2];
		}
		info.mip = calculate_mip(vertices_2d, vertex_count, info.depth_equation, tc_equation, mip_bias_);
		const float mip_scale = 1.0f / (float)(1 << info.mip);
		for (int i = 0; i < 6; ++i)
			tc_equation[i] *= mip_scale;
		int32_t tex_coord_max[2];
		for (int i = 0; i < 2; ++i)
		{
			info.tc_min[i] = f32_to_i32_saturating(std::floor(tc_min[i]) * mip_scale);
			tex_coord_max[i] = std::max(f32_to_i32_saturating(std::ceil(tc_max[i]) * mip_scale), info.tc_min[i] + 1);
		}
		for (int i = 0; i < 2; ++i) // shift to the surface rect
		{
			const float t = (float)info.tc_min[i];
			tc_equation[i] -= t * info.depth_equation[0];
			tc_equation[2 + i] -= t * info.depth_equation[1];
			tc_equation[4 + i] -= t * info.depth_equation[2];
		}
		std::memcpy(info.tex_coord_equation, tc_equation, sizeof(tc_equation));
		info.resolution[0] = (uint32_t)(tex_coord_max[0] - info.tc_min[0]);
		info.resolution[1] = (uint32_t)(tex_coord_max[1] - info.tc_min[1]);
		if (info.resolution[0] > 16384 || info.resolution[1] > 16384) // the device view limit; nothing sane asks for more
			continue;
		portals_info_[portal_index] = info;
	}
}

Renderer* Renderer::prepare_portal_view(size_t portal_index, int color_bits) // partial_renderer.rs:2017-2225
{
	if (portal_index >= frame_portals_.size())
		return nullptr;
	const ViewPortal& portal = frame_portals_[portal_index];
	const PortalInfo& portal_info = portals_info_[portal_index];
	if (portal_info.resolution[0] * portal_info.resolution[1] == 0)
		return nullptr;
	if (config_.portals_depth == 0)
		return nullptr; // "Portal/mirror depth limit reached": the texture is filled with black
	if (!portal_renderer_)
	{
		RendererConfig child_config = config_;
		child_config.portals_depth = config_.portals_depth - 1;
		portal_renderer_.reset(new Renderer(scene_, child_config));
	}
	Renderer& renderer = *portal_renderer_;
	// the nested renderer sees the same FrameWorldInfo
	renderer.frame_lights_ = frame_lights_;
	renderer.frame_models_ = frame_models_;
	renderer.frame_sprites_ = frame_sprites_;
	renderer.frame_decals_ = frame_decals_;
	renderer.frame_portals_ = frame_portals_;
	renderer.submodel_entities_ = submodel_entities_;

	CameraMatrices portal_camera_matrices;
	const uint32_t* start_leafs = nullptr;
	size_t num_start_leafs = 0;
	const float res_w = (float)portal_info.resolution[0], res_h = (float)portal_info.resolution[1];
	auto tex_coord_space_matrices = [&](Vec3f eye, float dist_normalized, const Mat4f* post) {
		// view matrix that maps the portal plane's texture space onto the texture's pixels, with the plane at Z_NEAR (:2069-2105)
		const float normal_len = magnitude(portal.plane.vec);
		const float mip_scale = 1.0f / (float)(1 << portal_info.mip);
		const Mat4f tex_coord_basis = Mat4f::from_cols(
										 extend(portal.tex_coord_equation[0].vec, portal.tex_coord_equation[0].dist) * mip_scale,
										 extend(portal.tex_coord_equation[1].vec, portal.tex_coord_equation[1].dist) * mip_scale,
										 extend(portal.plane.vec, -portal.plane.dist) * (mip_scale / normal_len), Vec4f{0.0f, 0.0f, 0.0f, 1.0f})
										 .transpose();
		const Vec3f eye_tc = truncate(tex_coord_basis * extend(eye, 1.0f));
		const Mat4f translate = Mat4f::from_translation(-eye_tc);
		Mat4f shift_to_viewport = Mat4f::identity();
		shift_to_viewport.z.x = eye_tc.x - (float)portal_info.tc_min[0];
		shift_to_viewport.z.y = eye_tc.y - (float)portal_info.tc_min[1];
		const Mat4f z_scale = Mat4f::from_nonuniform_scale(Z_NEAR, Z_NEAR, Z_NEAR / (dist_normalized * mip_scale));
		Mat4f m = shift_to_viewport * z_scale * translate * tex_coord_basis;
		if (post)
			m = m * *post;
		CameraMatrices cm;
		cm.view_matrix = m;
		Mat4f inv = Mat4f::identity();
		m.transpose().invert(inv);
		cm.planes_matrix = inv;
		return cm;
	};
	switch (portal.view)
	{
	case PORTAL_VIEW_CAMERA_AT_POSITION:
		portal_camera_matrices = build_view_matrix_with_full_rotation(portal.camera_position, portal.camera_rotation, portal.camera_fov, res_w, res_h);
		break;
	case PORTAL_VIEW_MIRROR:
	{
		const float normal_square_len = dot(portal.plane.vec, portal.plane.vec);
		const float normal_len = std::sqrt(normal_square_len);
		const float dist = dot(portal.plane.vec, camera_matrices_.position) - portal.plane.dist;
		const Vec3f position_reflected = camera_matrices_.position - portal.plane.vec * (2.0f * dist / normal_square_len);
		portal_camera_matrices = tex_coord_space_matrices(position_reflected, dist / normal_len, nullptr);
		portal_camera_matrices.position = position_reflected;
		start_leafs = portal_leafs_[portal_index].data(); // start the visibility search where the mirror is
		num_start_leafs = portal_leafs_[portal_index].size();
		break;
	}
	case PORTAL_VIEW_PARALLAX_PORTAL:
	{
		Vec3f transformed[MAX_VERTICES];
		const size_t n = std::min(portal.vertices.size(), MAX_VERTICES);
		for (size_t i = 0; i < n; ++i)
			transformed[i] = truncate(portal.transform_matrix * extend(portal.vertices[i], 1.0f));
		transformed_portal_leafs_.clear();
		// get_convex_hull_bsp_leafs (bsp_map_compact.rs:294-338)
		std::vector<uint32_t> stack = {map_.root_node()};
		while (!stack.empty())
		{
			const uint32_t node_index = stack.back();
			stack.pop_back();
			if (node_index >= FIRST_LEAF_INDEX)
			{
				transformed_portal_leafs_.push_back(node_index - FIRST_LEAF_INDEX);
				continue;
			}
			const BSPNode& node = map_.nodes[node_index];
			size_t vertices_front = 0;
			for (size_t i = 0; i < n; ++i)
				if (dot(node.plane.vec, transformed[i]) > node.plane.dist)
					++vertices_front;
			if (vertices_front < n)
				stack.push_back(node.children[1]);
			if (vertices_front > 0)
				stack.push_back(node.children[0]); // popped first: the recursion's order
		}
		const float normal_len = magnitude(portal.plane.vec);
		const float dist_normalized = (portal.plane.dist - dot(portal.plane.vec, camera_matrices_.position)) / normal_len;
		Mat4f transform_inverse = Mat4f::identity();
		portal.transform_matrix.invert(transform_inverse);
		portal_camera_matrices = tex_coord_space_matrices(camera_matrices_.position, dist_normalized, &transform_inverse);
		portal_camera_matrices.position = truncate(portal.transform_matrix * extend(camera_matrices_.position, 1.0f));
		start_leafs = transformed_portal_leafs_.data();
		num_start_leafs = transformed_portal_leafs_.size();
		break;
	}
	}
	renderer.prepare_frame_with_matrices(portal_info.resolution[0], portal_info.resolution[1], portal_camera_matrices, true, start_leafs, num_start_leafs);
	renderer.draw_frame(color_bits);
	return &renderer;
}

void Renderer::portal_texture_light(size_t portal_index, float out[3]) const // partial_renderer.rs:2241-2262
{
	const ViewPortal& portal = frame_portals_[portal_index];
	const PortalInfo& portal_info = portals_info_[portal_index];
	std::memcpy(out, portal.texture_light_add, 12);
	if (!(portal.texture_light_scale > 0.0f))
		return;
	DrawPolygonData basis;
	basis_vecs_from_plane_and_tc(portal.plane, portal.tex_coord_equation, basis);
	const float scale = (float)(1 << portal_info.mip);
	const Vec3f u = portal_info.mip == 0 ? basis.basis_u : basis.basis_u * scale, v = portal_info.mip == 0 ? basis.basis_v : basis.basis_v * scale;
	const Vec3f polygon_center = basis.basis_start + u * ((float)portal_info.tc_min[0] + (float)portal_info.resolution[0] * 0.5f) + v * ((float)portal_info.tc_min[1] + (float)portal_info.resolution[1] * 0.5f);
	const LightGridElement grid_light = fetch_light_from_grid(polygon_center);
	float total_light[3];
	get_light_cube_light_host(grid_light.light_cube, basis.basis_normal, total_light);
	const float light_dir_dot = std::fmax(dot(basis.basis_normal, grid_light.light_direction_vector_scaled), 0.0f);
	for (int i = 0; i < 3; ++i)
	{
		total_light[i] += grid_light.directional_light_color[i] * light_dir_dot;
		out[i] = total_light[i] * portal.texture_light_scale + portal.texture_light_add[i];
	}
}

void Renderer::draw_view_portal(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t portal_index) // partial_renderer.rs:3720-3776
{
	const ViewPortal& portal = frame_portals_[portal_index];
	const PortalInfo& portal_info = portals_info_[portal_index];
	if (portal_info.resolution[0] * portal_info.resolution[1] == 0)
		return;
	Vec3f vertices_clipped[MAX_VERTICES], vertices_temp[MAX_VERTICES];
	size_t vertex_count = std::min(portal.vertices.size(), MAX_VERTICES);
	for (size_t i = 0; i < vertex_count; ++i)
		vertices_clipped[i] = view_matrix_transform_vertex(camera_matrices_.view_matrix, portal.vertices[i]);
	for (size_t pi = 0; pi < num_leaf_clip_planes; ++pi)
	{
		vertex_count = clip_3d_polygon_by_plane(vertices_clipped, vertex_count, leaf_clip_planes[pi], vertices_temp, MAX_VERTICES);
		if (vertex_count < 3)
			return;
		for (size_t i = 0; i < vertex_count; ++i)
			vertices_clipped[i] = vertices_temp[i];
	}
	// the texture is another render target: tex_ref = the PORTAL index, replaced by the caller with the device view it rendered
	// the nested frame into (SQW_TEX_VIEW)
	draw_polygon(clip_planes, vertices_clipped, vertex_count, portal_info.depth_equation, portal_info.tex_coord_equation, portal_info.resolution, SQW_TEX_VIEW, portal_index, portal.blending_mode);
}

// ---- sprites -------------------------------------------------------------------------------------------------
void Renderer::position_sprite_sphere_r(uint32_t id, Vec3f center, float radius, uint32_t node_index) // dynamic_objects_index.rs:211-243
{
	if (node_index >= FIRST_LEAF_INDEX)
	{
		const uint32_t leaf_index = node_index - FIRST_LEAF_INDEX;
		leaf_sprites_[leaf_index].push_back(id);
		sprite_leafs_[id].push_back(leaf_index);
		return;
	}
	const BSPNode& node = map_.nodes[node_index];
	const float scaled_radius = radius * magnitude(node.plane.vec);
	const float scaled_dist = dot(center, node.plane.vec);
	if (scaled_dist + scaled_radius >= node.plane.dist)
		position_sprite_sphere_r(id, center, radius, node.children[0]);
	if (scaled_dist - scaled_radius <= node.plane.dist)
		position_sprite_sphere_r(id, center, radius, node.children[1]);
}

void Renderer::position_sprites() // dynamic_objects_index.rs:93-106
{
	for (auto& l : leaf_sprites_)
		l.clear();
	sprite_leafs_.assign(frame_sprites_.size(), {});
	for (size_t index = 0; index < frame_sprites_.size(); ++index)
		position_sprite_sphere_r((uint32_t)index, frame_sprites_[index].position, frame_sprites_[index].radius, map_.root_node());
}

void Renderer::prepare_sprites() // partial_renderer.rs:740-990
{
	sprites_info_.assign(frame_sprites_.size(), SpriteInfo{});
	if (frame_sprites_.empty())
		return;
	Mat4f view_matrix_inverse = Mat4f::identity();
	camera_matrices_.view_matrix.invert(view_matrix_inverse);
	const Vec3f u_vec_base_initial = truncate(view_matrix_inverse * Vec4f{1, 0, 0, 0});
	const Vec3f u_vec_base = u_vec_base_initial / magnitude(u_vec_base_initial);
	const Vec3f v_vec_base_initial = truncate(view_matrix_inverse * Vec4f{0, 1, 0, 0});
	const Vec3f v_vec_base = v_vec_base_initial / magnitude(v_vec_base_initial);
	const Vec3f camera_direction = cross(u_vec_base, v_vec_base);
	const Vec3f unit_x = {1, 0, 0}, minus_unit_z = {0, 0, -1};
	for (size_t sprite_index = 0; sprite_index < frame_sprites_.size(); ++sprite_index)
	{
		if (sprite_leafs_[sprite_index].empty())
			continue;
		const Sprite& sprite = frame_sprites_[sprite_index];
		Vec3f u_vec_normalized = u_vec_base, v_vec_normalized = v_vec_base;
		switch (sprite.orientation)
		{
		case SPRITE_PARALLEL_TO_CAMERA_PLANE: break;
		case SPRITE_FACING_TOWARDS_CAMERA:
		{
			const Vec3f vec_to_camera = camera_matrices_.position - sprite.position;
			const Vec3f plane_normal = vec_to_camera / std::fmax(magnitude(vec_to_camera), 0.001f);
			const Vec3f u_vec_projected_to_plane = u_vec_base - plane_normal * dot(u_vec_base, plane_normal);
			const float u_vec_projected_len = magnitude(u_vec_projected_to_plane);
			u_vec_normalized = u_vec_projected_len < 0.0001f ? unit_x : u_vec_projected_to_plane / u_vec_projected_len;
			v_vec_normalized = cross(u_vec_normalized, plane_normal);
			break;
		}
		case SPRITE_ALIGN_TO_Z_AXIS_PARALLEL_TO_CAMERA_PLANE:
		{
			v_vec_normalized = minus_unit_z;
			const Vec3f u_vec = cross(v_vec_normalized, camera_direction);
			const float u_vec_len = magnitude(u_vec);
			u_vec_normalized = u_vec_len < 0.0001f ? unit_x : u_vec / u_vec_len;
			break;
		}
		case SPRITE_ALIGN_TO_Z_AXIS_FACING_TOWARDS_CAMERA:
		{
			const Vec3f vec_to_camera = camera_matrices_.position - sprite.position;
			v_vec_normalized = minus_unit_z;
			const Vec3f u_vec = cross(vec_to_camera, v_vec_normalized);
			const float u_vec_len = magnitude(u_vec);
			u_vec_normalized = u_vec_len < 0.0001f ? unit_x : u_vec / u_vec_len;
			break;
		}
		}
		if (sprite.angle != 0.0f)
		{
			const float c = std::cos(sprite.angle), sn = std::sin(sprite.angle);
			const Vec3f u2 = u_vec_normalized * c - v_vec_normalized * sn, v2 = u_vec_normalized * sn + v_vec_normalized * c;
			u_vec_normalized = u2;
			v_vec_normalized = v2;
		}
		const HostImage& texture_mip0 = scene_.images[sprite.texture_images[0]];
		const float ratio_h_w = (float)texture_mip0.height / (float)texture_mip0.width;
		const float step_u = sprite.radius * (1.0f / std::sqrt(1.0f + ratio_h_w * ratio_h_w)); // inv_sqrt_fast, portable branch
		const float ratio_w_h = (float)texture_mip0.width / (float)texture_mip0.height;
		const float step_v = sprite.radius * (1.0f / std::sqrt(1.0f + ratio_w_h * ratio_w_h));
		const Vec3f u_vec = u_vec_normalized * step_u, v_vec = v_vec_normalized * step_v;
		const Vec3f vertices[4] = {sprite.position + u_vec + v_vec, sprite.position + u_vec - v_vec, sprite.position - u_vec - v_vec, sprite.position - u_vec + v_vec};
		SpriteInfo& info = sprites_info_[sprite_index];
		for (int i = 0; i < 4; ++i)
			info.vertices_projected[i] = view_matrix_transform_vertex(camera_matrices_.view_matrix, vertices[i]);
		const Vec3f* vp = info.vertices_projected;
		if (vp[0].z > 0.0f && vp[1].z > 0.0f && vp[2].z > 0.0f && vp[3].z > 0.0f)
		{
			Vec2f points[4];
			for (int i = 0; i < 4; ++i)
				points[i] = truncate(vp[i]) / vp[i].z;
			const Vec2f d0 = points[0] - points[2], d1 = points[1] - points[3];
			const float shortest_diagonal_square_len = std::fmin(dot(d0, d0), dot(d1, d1));
			if (shortest_diagonal_square_len <= 0.0f)
				info.mip = MAX_MIP;
			else
			{
				const float tc_diagonal_square_diff = (float)texture_mip0.width * (float)texture_mip0.width + (float)texture_mip0.height * (float)texture_mip0.height;
				const float d_tc_2 = tc_diagonal_square_diff / shortest_diagonal_square_len;
				const float mip_f = std::log2(d_tc_2) * 0.5f + mip_bias_;
				const float m = std::fmax(std::round(mip_f), 0.0f); // f32::round: half away from zero, like std::round
				info.mip = std::min<uint32_t>(m >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)m, MAX_MIP);
			}
		}
		else
			info.mip = 0;
		const float min_z = std::fmin(std::fmin(vp[0].z, vp[1].z), std::fmin(vp[2].z, vp[3].z));
		const float max_z = std::fmax(std::fmax(vp[0].z, vp[1].z), std::fmax(vp[2].z, vp[3].z));
		if (min_z <= 0.0f)
			info.tesselation_level = MAX_SPRITE_TESSELATION_LEVEL;
		else
		{
			const float t = std::fmax((max_z / min_z - 1.0f) * 4.0f, 1.0f);
			info.tesselation_level = std::min<uint32_t>(t >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)t, MAX_SPRITE_TESSELATION_LEVEL);
		}
		std::memcpy(info.light, sprite.light_add, 12);
		if (sprite.light_scale > 0.0f)
		{
			const LightGridElement grid_light = fetch_light_from_grid(sprite.position); // :915-932
			const float cube_side_scale = sprite.light_scale / 6.0f;
			for (const auto& cube_side : grid_light.light_cube)
				for (int i = 0; i < 3; ++i)
					info.light[i] += cube_side[i] * cube_side_scale;
			const float directional_component_scale = sprite.light_scale * 0.5f * magnitude(grid_light.light_direction_vector_scaled);
			for (int i = 0; i < 3; ++i)
				info.light[i] += directional_component_scale * grid_light.directional_light_color[i];
			// dynamic lights without shadow maps (:934-980; see get_model_light for the ones with)
			const float dynamic_lights_scale = sprite.light_scale * 0.5f;
			const float min_square_distance = sprite.radius * sprite.radius;
			for (size_t li = 0; li < frame_lights_.size(); ++li)
			{
				const DynamicLight& light = frame_lights_[li];
				if (!dynamic_lights_info_[li].visible || light.shadow_type != SQW_SHADOW_NONE)
					continue;
				const Vec3f vec_to_light = light.position - sprite.position;
				const float inv_square_dist = 1.0f / std::fmax(dot(vec_to_light, vec_to_light), min_square_distance);
				const float light_inv_square_radius = 1.0f / (light.radius * light.radius);
				if (inv_square_dist < light_inv_square_radius)
					continue;
				const float scale = dynamic_lights_scale * 1.0f * (inv_square_dist - light_inv_square_radius);
				for (int i = 0; i < 3; ++i)
					info.light[i] += scale * light.color[i];
			}
		}
		++stats_.num_visible_sprites;
	}
}

void Renderer::emit_triangle_fan(const ModelVertex2d* vertices, size_t num_vertices, const float* constant_light, uint32_t image_id, int blending_mode)
{
	const HostImage& texture = scene_.images[image_id];
	sqw_triangle_point fixed[MAX_VERTICES];
	for (size_t k = 0; k < num_vertices; ++k)
	{
		const float* light = constant_light ? constant_light : vertices[k].light;
		fixed[k].x = f32_to_fixed16(vertices[k].pos.x);
		fixed[k].y = f32_to_fixed16(vertices[k].pos.y);
		fixed[k].tc[0] = f32_to_fixed16(vertices[k].tc.x);
		fixed[k].tc[1] = f32_to_fixed16(vertices[k].tc.y);
		for (int i = 0; i < 3; ++i)
			fixed[k].light[i] = f32_to_fixed16(light[i]);
	}
	for (size_t t = 0; t + 2 < num_vertices; ++t)
	{
		sqw_draw_cmd c;
		std::memset(&c, 0, sizeof(c));
		c.kind = SQW_CMD_TRIANGLE;
		c.blend = (uint8_t)blending_mode;
		c.tex_kind = SQW_TEX_IMAGE;
		c.tex_ref = image_id;
		c.tex_size[0] = (int32_t)texture.width;
		c.tex_size[1] = (int32_t)texture.height;
		c.num_vertices = 3;
		c.vertex_offset = (uint32_t)triangle_points_.size();
		triangle_points_.push_back(fixed[0]);
		triangle_points_.push_back(fixed[t + 1]);
		triangle_points_.push_back(fixed[t + 2]);
		cmds_.push_back(c);
		++stats_.num_triangles;
	}
}

void Renderer::draw_sprite(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t sprite_index) // partial_renderer.rs:3290-3523
{
	const Sprite& sprite = frame_sprites_[sprite_index];
	const SpriteInfo& sprite_info = sprites_info_[sprite_index];
	const uint32_t image_id = sprite.texture_images[sprite_info.mip];
	const HostImage& texture_mip = scene_.images[image_id];
	const float texture_size[2] = {(float)texture_mip.width, (float)texture_mip.height};
	ModelVertex3d vertices_clipped0[MAX_VERTICES] = {}, vertices_clipped1[MAX_VERTICES] = {};
	ModelVertex3d *vc_src = vertices_clipped0, *vc_dst = vertices_clipped1;
	const float tc_reduce = 1.0f / 16.0f;
	const Vec2f tcs[4] = {{texture_size[0] - tc_reduce, texture_size[1] - tc_reduce}, {texture_size[0] - tc_reduce, tc_reduce}, {tc_reduce, tc_reduce}, {tc_reduce, texture_size[1] - tc_reduce}};
	for (int i = 0; i < 4; ++i)
	{
		vc_src[i].pos = sprite_info.vertices_projected[i];
		vc_src[i].tc = tcs[i];
		std::memcpy(vc_src[i].light, sprite_info.light, 12);
	}
	const Plane z_near_plane = {{0.0f, 0.0f, 1.0f}, Z_NEAR};
	ModelVertex2d vertices_projected0[MAX_VERTICES], vertices_projected1[MAX_VERTICES];
	// clip one convex polygon held in vc_src by the near plane, the leaf planes and the 8 screen-space planes, then emit its fan
	auto clip_and_emit = [&](size_t num_vertices, const float* constant_light) {
		for (size_t pi = 0; pi <= num_leaf_clip_planes; ++pi)
		{
			num_vertices = clip_3d_model_polygon_by_plane(vc_src, num_vertices, pi == 0 ? z_near_plane : leaf_clip_planes[pi - 1], vc_dst, MAX_VERTICES);
			if (num_vertices < 3)
				return;
			std::swap(vc_src, vc_dst);
		}
		ModelVertex2d *vp_src = vertices_projected0, *vp_dst = vertices_projected1;
		for (size_t k = 0; k < num_vertices; ++k)
		{
			vp_src[k].pos = truncate(vc_src[k].pos) / vc_src[k].pos.z;
			vp_src[k].tc = vc_src[k].tc;
			std::memcpy(vp_src[k].light, vc_src[k].light, 12);
		}
		for (int pi = 0; pi < 8; ++pi)
		{
			num_vertices = clip_2d_model_polygon(vp_src, num_vertices, clip_planes[pi], vp_dst, MAX_VERTICES);
			if (num_vertices < 3)
				return;
			std::swap(vp_src, vp_dst);
		}
		emit_triangle_fan(vp_src, num_vertices, constant_light, image_id, sprite.blending_mode);
	};
	if (sprite_info.tesselation_level <= 1)
	{
		clip_and_emit(4, nullptr);
		return;
	}
	// grid tesselation (:3405-3520): bilinear sub-quads, each clipped on its own; the light is the sprite's constant
	ModelVertex3d tesselated_vertices[(MAX_SPRITE_TESSELATION_LEVEL + 1) * (MAX_SPRITE_TESSELATION_LEVEL + 1)] = {};
	const uint32_t level = sprite_info.tesselation_level, num_vertices_in_row = level + 1;
	const float inv_step = 1.0f / (float)level;
	const ModelVertex3d corners[4] = {vc_src[0], vc_src[1], vc_src[2], vc_src[3]};
	for (uint32_t y = 0; y < num_vertices_in_row; ++y)
	{
		const float y_mix = (float)y * inv_step;
		for (uint32_t x = 0; x < num_vertices_in_row; ++x)
		{
			const float x_mix = (float)x * inv_step;
			const float weights[4] = {x_mix * y_mix, (1.0f - x_mix) * y_mix, (1.0f - x_mix) * (1.0f - y_mix), x_mix * (1.0f - y_mix)};
			ModelVertex3d& dst_v = tesselated_vertices[x + y * num_vertices_in_row];
			dst_v.pos = corners[0].pos * weights[0] + corners[1].pos * weights[1] + corners[2].pos * weights[2] + corners[3].pos * weights[3];
			dst_v.tc = corners[0].tc * weights[0] + corners[1].tc * weights[1] + corners[2].tc * weights[2] + corners[3].tc * weights[3];
		}
	}
	for (uint32_t x = 0; x < level; ++x)
		for (uint32_t y = 0; y < level; ++y)
		{
			vc_src[0] = tesselated_vertices[x + y * num_vertices_in_row];
			vc_src[1] = tesselated_vertices[x + (y + 1) * num_vertices_in_row];
			vc_src[2] = tesselated_vertices[1 + x + (y + 1) * num_vertices_in_row];
			vc_src[3] = tesselated_vertices[1 + x + y * num_vertices_in_row];
			clip_and_emit(4, sprite_info.light);
		}
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
