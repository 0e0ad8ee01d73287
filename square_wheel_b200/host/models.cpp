// See models.hpp.  Function-by-function restatement of src/square_wheel/triangle_models_rendering.rs, draw_ordering.rs,
// common/light_cube.rs and the model clippers of common/clipping.rs, plus seeded model fixtures.
#include "models.hpp"
#include "renderer.hpp"
#include <algorithm>
#include <cmath>

namespace sqwh
{

bool Mat3f::invert(Mat3f& out) const // cgmath 0.18 Matrix3::invert
{
	const float det = x.x * (y.y * z.z - z.y * y.z) - y.x * (x.y * z.z - z.y * x.z) + z.x * (x.y * y.z - y.y * x.z);
	if (det == 0.0f)
		return false;
	const Mat3f m = {cross(y, z) / det, cross(z, x) / det, cross(x, y) / det};
	out = m.transpose();
	return true;
}

Mat3f get_normals_matrix(const Mat4f& model_matrix) // triangle_models_rendering.rs:722-731
{
	const Mat3f axis_matrix = {truncate(model_matrix.x), truncate(model_matrix.y), truncate(model_matrix.z)};
	Mat3f inv;
	if (!axis_matrix.transpose().invert(inv))
		return Mat3f::identity();
	return inv;
}

BBox get_current_triangle_model_bbox(const TriangleModel& model, const uint32_t frames[2]) // :204-211
{
	BBox bbox = model.frames_info[frames[0]];
	bbox.extend(model.frames_info[frames[1]]);
	return bbox;
}

ModelLightData ModelLightData::from_two_light_grid_elements(const LightGridElement& a, const LightGridElement& b) // :392-411
{
	ModelLightData r;
	for (int side = 0; side < 6; ++side)
		for (int c = 0; c < 3; ++c)
			r.light_cube[side][c] = a.light_cube[side][c] + b.light_cube[side][c];
	r.vector_scaled[0] = a.light_direction_vector_scaled;
	r.vector_scaled[1] = b.light_direction_vector_scaled;
	for (int c = 0; c < 3; ++c)
	{
		r.color[0][c] = a.directional_light_color[c];
		r.color[1][c] = b.directional_light_color[c];
	}
	return r;
}

namespace
{
constexpr float MIN_VEC_LEN = 0.00000001f; // light_cube.rs:162

// fast_math.rs:376 (portable branch, as everywhere in this repository)
inline float inv_sqrt_fast(float x) { return 1.0f / std::sqrt(x); }

inline void get_light_cube_light(const float light_cube[6][3], Vec3f n, float total_light[3]) // light.rs:296-348
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

inline void get_vertex_light(const ModelLightData& light, Vec3f normal_transformed, float out[3]) // triangle_models_rendering.rs:701-720
{
	const Vec3f normal_normalized = normal_transformed * inv_sqrt_fast(std::fmax(dot(normal_transformed, normal_transformed), 0.00000001f));
	get_light_cube_light(light.light_cube, normal_normalized, out);
	for (int d = 0; d < 2; ++d)
	{
		const float light_dir_dot = std::fmax(dot(normal_normalized, light.vector_scaled[d]), 0.0f);
		for (int i = 0; i < 3; ++i)
			out[i] += light.color[d][i] * light_dir_dot;
	}
}

inline ModelVertex3d make_vertex(const Mat4f& mv, Vec3f position, Vec3f normal_transformed, const float tex_coord[2], Vec2f tc_scale, Vec2f tc_shift, const ModelLightData& light)
{
	ModelVertex3d v;
	v.pos = view_matrix_transform_vertex(mv, position);
	v.tc = Vec2f{tex_coord[0] * tc_scale.x, tex_coord[1] * tc_scale.y} + tc_shift;
	get_vertex_light(light, normal_transformed, v.light);
	return v;
}
} // namespace

void animate_and_transform_triangle_mesh_vertices( // triangle_models_rendering.rs:7-203
	const TriangleModel& model, const TriangleModelMesh& mesh, const uint32_t frames[2], float lerp, const ModelLightData& light, const Mat4f& model_matrix,
	const Mat4f& model_view_matrix, Vec2f tc_scale, Vec2f tc_shift, ModelVertex3d* dst_vertices)
{
	const Mat3f normals_matrix = get_normals_matrix(model_matrix);
	const size_t frame0 = frames[0], frame1 = frames[1];
	const float lerp0 = std::fmin(std::fmax(lerp, 0.0f), 1.0f);
	const float lerp1 = 1.0f - lerp0;
	const bool perform_lerp = lerp0 > 0.01f && lerp0 < 0.99f;

	switch (mesh.kind)
	{
	case VERTEX_DATA_NON_ANIMATED:
		for (size_t i = 0; i < mesh.non_animated.size(); ++i)
		{
			const VertexNonAnimated& v = mesh.non_animated[i];
			dst_vertices[i] = make_vertex(model_view_matrix, v.position, normals_matrix * v.normal, v.tex_coord, tc_scale, tc_shift, light);
		}
		break;
	case VERTEX_DATA_VERTEX_ANIMATED:
	{
		const size_t n = mesh.constant.size();
		const VertexAnimatedVertexVariable* f0 = mesh.variable.data() + frame0 * n;
		const VertexAnimatedVertexVariable* f1 = mesh.variable.data() + frame1 * n;
		if (perform_lerp)
		{
			for (size_t i = 0; i < n; ++i)
			{
				const Vec3f position_lerped = f0[i].position * lerp0 + f1[i].position * lerp1;
				const Vec3f normal_lerped = f0[i].normal * lerp0 + f1[i].normal * lerp1;
				dst_vertices[i] = make_vertex(model_view_matrix, position_lerped, normals_matrix * normal_lerped, mesh.constant[i].tex_coord, tc_scale, tc_shift, light);
			}
		}
		else
		{
			const VertexAnimatedVertexVariable* f = lerp0 > lerp1 ? f0 : f1;
			for (size_t i = 0; i < n; ++i)
				dst_vertices[i] = make_vertex(model_view_matrix, f[i].position, normals_matrix * f[i].normal, mesh.constant[i].tex_coord, tc_scale, tc_shift, light);
		}
		break;
	}
	case VERTEX_DATA_SKELETON_ANIMATED:
	{
		if (model.frame_bones.empty())
		{
			for (size_t i = 0; i < mesh.skeleton.size(); ++i)
			{
				const SkeletonAnimatedVertex& v = mesh.skeleton[i];
				dst_vertices[i] = make_vertex(model_view_matrix, v.position, normals_matrix * v.normal, v.tex_coord, tc_scale, tc_shift, light);
			}
			break;
		}
		const size_t num_bones = std::min(model.bones.size(), MAX_TRIANGLE_MODEL_BONES);
		const Mat4f* frame_bones0 = model.frame_bones.data() + frame0 * model.bones.size();
		const Mat4f* frame_bones1 = model.frame_bones.data() + frame1 * model.bones.size();
		Mat4f matrices[MAX_TRIANGLE_MODEL_BONES];
		Mat3f normals_matrices[MAX_TRIANGLE_MODEL_BONES];
		// relies on the bones being sorted in hierarchy order (:121)
		for (size_t bone_index = 0; bone_index < num_bones; ++bone_index)
		{
			const size_t parent = model.bones[bone_index];
			const Mat4f mat = perform_lerp ? frame_bones0[bone_index] * lerp0 + frame_bones1[bone_index] * lerp1 : (lerp0 > lerp1 ? frame_bones0[bone_index] : frame_bones1[bone_index]);
			matrices[bone_index] = parent < model.bones.size() && parent < bone_index ? matrices[parent] * mat : mat;
		}
		const float weight_scale = 1.0f / 255.0f;
		const Mat3f normals_matrix_weight_scaled = normals_matrix * weight_scale;
		const Mat4f model_view_matrix_weight_scaled = model_view_matrix * weight_scale;
		for (size_t b = 0; b < num_bones; ++b)
		{
			normals_matrices[b] = normals_matrix_weight_scaled * get_normals_matrix(matrices[b]);
			matrices[b] = model_view_matrix_weight_scaled * matrices[b];
		}
		for (size_t i = 0; i < mesh.skeleton.size(); ++i)
		{
			const SkeletonAnimatedVertex& v = mesh.skeleton[i];
			const size_t i0 = std::min<size_t>(v.bone_index[0], num_bones - 1);
			const float w0 = (float)v.weight[0];
			Mat4f mat = matrices[i0] * w0;
			Mat3f normal_mat = normals_matrices[i0] * w0;
			for (int k = 1; k < 4; ++k)
				if (v.weight[k] > 0)
				{
					const size_t ii = std::min<size_t>(v.bone_index[k], num_bones - 1);
					const float wi = (float)v.weight[k];
					mat = mat + matrices[ii] * wi;
					normal_mat = normal_mat + normals_matrices[ii] * wi;
				}
			dst_vertices[i] = make_vertex(mat, v.position, normal_mat * v.normal, v.tex_coord, tc_scale, tc_shift, light);
		}
		break;
	}
	}
}

uint32_t calculate_triangle_model_texture_mip(const Mat4f& model_view_matrix, const BBox& model_bbox, const uint32_t texture_size[2], float mip_bias) // :275-325
{
	const Vec3f bbox_center = model_bbox.get_center();
	const Vec3f bbox_half_size = model_bbox.get_size() * 0.5f;
	const float cube_half_size = (bbox_half_size.x + bbox_half_size.y + bbox_half_size.z) / 3.0f;
	const Vec3f h = {cube_half_size, cube_half_size, cube_half_size};
	const BBox cube_bbox = BBox::from_min_max(bbox_center - h, bbox_center + h);
	Vec3f corners[8];
	cube_bbox.get_corners_vertices(corners);
	Vec2f projected[8];
	for (int i = 0; i < 8; ++i)
	{
		const Vec4f p = model_view_matrix * extend(corners[i], 1.0f);
		if (p.w <= 0.0f)
			return 0;
		projected[i] = Vec2f{p.x, p.y} / p.w;
	}
	float min_x = projected[0].x, max_x = min_x, min_y = projected[0].y, max_y = min_y;
	for (int i = 1; i < 8; ++i)
	{
		min_x = std::fmin(min_x, projected[i].x);
		max_x = std::fmax(max_x, projected[i].x);
		min_y = std::fmin(min_y, projected[i].y);
		max_y = std::fmax(max_y, projected[i].y);
	}
	const float max_dimension = std::fmax(max_x - min_x, max_y - min_y);
	if (max_dimension <= 0.0f)
		return MAX_MIP;
	const float approximate_texels_per_pixel = (float)std::max(texture_size[0], texture_size[1]) / max_dimension;
	const float m = std::fmax(std::floor(std::log2(approximate_texels_per_pixel) + mip_bias), 0.0f);
	return std::min<uint32_t>(m >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)m, MAX_MIP);
}

namespace
{
inline float triangle_plane_dist(const ModelVertex3d* v, const Triangle& t) // get_triangle_plane (:366-374)
{
	const Vec3f v0 = v[t[0]].pos, v1 = v[t[1]].pos, v2 = v[t[2]].pos;
	return dot(cross(v1 - v0, v2 - v1), v0);
}
inline float triangle_max_z(const ModelVertex3d* v, const Triangle& t) { return std::fmax(std::fmax(v[t[0]].pos.z, v[t[1]].pos.z), v[t[2]].pos.z); }
} // namespace

size_t reject_triangle_model_back_faces(const ModelVertex3d* transformed_vertices, const Triangle* triangles, size_t num_triangles, Triangle* out_triangles) // :327-345
{
	size_t n = 0;
	for (size_t i = 0; i < num_triangles; ++i)
		if (triangle_plane_dist(transformed_vertices, triangles[i]) > 0.0f)
			out_triangles[n++] = triangles[i];
	return n;
}

void sort_model_triangles(const ModelVertex3d* transformed_vertices, Triangle* triangles, size_t num_triangles) // :347-364
{
	std::stable_sort(triangles, triangles + num_triangles, [&](const Triangle& a, const Triangle& b) { return triangle_max_z(transformed_vertices, a) > triangle_max_z(transformed_vertices, b); });
}

namespace
{
template <class V, class P> inline V lerp_vertices(const V& v0, const V& v1, float k0, float k1, P pos_lerp)
{
	V r;
	r.pos = pos_lerp(v0.pos, v1.pos);
	r.tc = v0.tc * k1 - v1.tc * k0;
	for (int i = 0; i < 3; ++i)
		r.light[i] = v0.light[i] * k1 - v1.light[i] * k0;
	return r;
}
inline ModelVertex3d get_model_line_plane_intersection(const ModelVertex3d& v0, const ModelVertex3d& v1, const Plane& plane) // clipping.rs:334-351
{
	const float dist0 = dot(v0.pos, plane.vec) - plane.dist, dist1 = dot(v1.pos, plane.vec) - plane.dist;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
	return lerp_vertices(v0, v1, k0, k1, [&](Vec3f a, Vec3f b) { return a * k1 - b * k0; });
}
inline ModelVertex2d get_model_line_line_intersection(const ModelVertex2d& v0, const ModelVertex2d& v1, Vec3f line) // clipping.rs:315-332
{
	const float dist0 = dot(v0.pos, truncate(line)) - line.z, dist1 = dot(v1.pos, truncate(line)) - line.z;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
	return lerp_vertices(v0, v1, k0, k1, [&](Vec2f a, Vec2f b) { return a * k1 - b * k0; });
}
} // namespace

size_t clip_3d_model_polygon_by_plane(const ModelVertex3d* polygon, size_t n, const Plane& plane, ModelVertex3d* out_polygon, size_t out_capacity) // clipping.rs:259-313
{
	const ModelVertex3d* prev_v = &polygon[n - 1];
	float prev_dist = dot(plane.vec, prev_v->pos);
	size_t out_vertex_count = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const ModelVertex3d& v = polygon[i];
		const float dist = dot(plane.vec, v.pos);
		if (dist > plane.dist)
		{
			if (prev_dist < plane.dist)
			{
				out_polygon[out_vertex_count++] = get_model_line_plane_intersection(*prev_v, v, plane);
				if (out_vertex_count == out_capacity)
					break;
			}
			out_polygon[out_vertex_count++] = v;
			if (out_vertex_count == out_capacity)
				break;
		}
		else if (dist == plane.dist)
		{
			out_polygon[out_vertex_count++] = v;
			if (out_vertex_count == out_capacity)
				break;
		}
		else if (prev_dist > plane.dist)
		{
			out_polygon[out_vertex_count++] = get_model_line_plane_intersection(*prev_v, v, plane);
			if (out_vertex_count == out_capacity)
				break;
		}
		prev_v = &v;
		prev_dist = dist;
	}
	return out_vertex_count;
}

size_t clip_2d_model_polygon(const ModelVertex2d* polygon, size_t n, Vec3f line, ModelVertex2d* out_polygon, size_t out_capacity) // clipping.rs:203-257
{
	const ModelVertex2d* prev_v = &polygon[n - 1];
	float prev_dist = dot(prev_v->pos, truncate(line)) - line.z;
	size_t out_vertex_count = 0;
	for (size_t i = 0; i < n; ++i)
	{
		const ModelVertex2d& v = polygon[i];
		const float dist = dot(v.pos, truncate(line)) - line.z;
		if (dist > 0.0f)
		{
			if (prev_dist < 0.0f)
			{
				out_polygon[out_vertex_count++] = get_model_line_line_intersection(*prev_v, v, line);
				if (out_vertex_count == out_capacity)
					break;
			}
			out_polygon[out_vertex_count++] = v;
			if (out_vertex_count == out_capacity)
				break;
		}
		else if (dist == 0.0f)
		{
			out_polygon[out_vertex_count++] = v;
			if (out_vertex_count == out_capacity)
				break;
		}
		else if (prev_dist > 0.0f)
		{
			out_polygon[out_vertex_count++] = get_model_line_line_intersection(*prev_v, v, line);
			if (out_vertex_count == out_capacity)
				break;
		}
		prev_v = &v;
		prev_dist = dist;
	}
	return out_vertex_count;
}

// ---- light_cube.rs -------------------------------------------------------------------------------------------
void LightCube::add_light_sample(Vec3f direction, const float light_scaled[3]) // :22-82
{
	const Vec3f d = direction / std::fmax(magnitude(direction), MIN_VEC_LEN);
	const float comp[3] = {d.x, d.y, d.z};
	for (int axis = 0; axis < 3; ++axis)
	{
		if (comp[axis] <= 0.0f)
			for (int i = 0; i < 3; ++i)
				light_cube[axis * 2][i] += light_scaled[i] * (-comp[axis]);
		else
			for (int i = 0; i < 3; ++i)
				light_cube[axis * 2 + 1][i] += light_scaled[i] * comp[axis];
	}
	for (int i = 0; i < 3; ++i)
		sum_scaled_light_dirs[i] = sum_scaled_light_dirs[i] + d * light_scaled[i];
}

void LightCube::add_constant_light(const float light[3]) // :84-93
{
	for (auto& side : light_cube)
		for (int i = 0; i < 3; ++i)
			side[i] += light[i];
}

void LightCube::scale(float s) // :95-108
{
	for (auto& side : light_cube)
		for (float& c : side)
			c *= s;
	for (auto& v : sum_scaled_light_dirs)
		v = v * s;
}

LightGridElement LightCube::convert_into_light_grid_sample() const // :110-159
{
	LightGridElement r;
	const Vec3f dir = (sum_scaled_light_dirs[0] + sum_scaled_light_dirs[1] + sum_scaled_light_dirs[2]) * 0.33f;
	const float len2 = std::fmax(dot(dir, dir), MIN_VEC_LEN);
	float light_color[3];
	for (int i = 0; i < 3; ++i)
		light_color[i] = std::fmax(dot(sum_scaled_light_dirs[i], dir), 0.0f) / len2;
	const float comp[3] = {dir.x, dir.y, dir.z};
	for (int side = 0; side < 6; ++side)
		for (int i = 0; i < 3; ++i)
			r.light_cube[side][i] = light_cube[side][i];
	for (int i = 0; i < 3; ++i)
		for (int axis = 0; axis < 3; ++axis)
		{
			if (comp[axis] <= 0.0f)
				r.light_cube[axis * 2][i] = std::fmax(r.light_cube[axis * 2][i] - light_color[i] * (-comp[axis]), 0.0f);
			else
				r.light_cube[axis * 2 + 1][i] = std::fmax(r.light_cube[axis * 2 + 1][i] - light_color[i] * comp[axis], 0.0f);
		}
	r.light_direction_vector_scaled = dir;
	for (int i = 0; i < 3; ++i)
		r.directional_light_color[i] = light_color[i];
	return r;
}

// ---- draw_ordering.rs ----------------------------------------------------------------------------------------
namespace
{
inline Plane project_plane(const Plane& plane, const Mat4f& planes_matrix) // :160-167
{
	const Vec4f t = planes_matrix * extend(plane.vec, -plane.dist);
	return Plane{truncate(t), -t.w};
}
inline bool is_bbox_at_front_of_plane(const Plane& plane, const Vec3f vertices[8]) // :198-216
{
	if (plane.dist >= 0.0f)
		return false;
	int vertices_front = 0;
	for (int i = 0; i < 8; ++i)
		if (dot(plane.vec, vertices[i]) >= plane.dist)
			++vertices_front;
	return vertices_front == 8;
}
inline float get_projected_bbox_min_z(const ProjectedBBox& b)
{
	float r = b.vertices[0].z;
	for (const Vec3f& v : b.vertices)
		r = std::fmin(r, v.z);
	return r;
}
inline bool compare_projected_bboxes(const ProjectedBBox& l, const ProjectedBBox& r) // :175-196
{
	for (const Plane& p : l.planes)
		if (is_bbox_at_front_of_plane(p, r.vertices))
			return false;
	for (const Plane& p : r.planes)
		if (is_bbox_at_front_of_plane(p, l.vertices))
			return true;
	return get_projected_bbox_min_z(l) < get_projected_bbox_min_z(r);
}
} // namespace

ProjectedBBox project_bbox(const BBox& bbox, const CameraMatrices& camera_matrices) // :86-158
{
	ProjectedBBox r;
	const Plane pairs[3][2] = {
		{{{1, 0, 0}, bbox.max.x}, {{-1, 0, 0}, -bbox.min.x}},
		{{{0, 1, 0}, bbox.max.y}, {{0, -1, 0}, -bbox.min.y}},
		{{{0, 0, 1}, bbox.max.z}, {{0, 0, -1}, -bbox.min.z}},
	};
	for (int i = 0; i < 3; ++i)
	{
		const Plane p0 = project_plane(pairs[i][0], camera_matrices.planes_matrix), p1 = project_plane(pairs[i][1], camera_matrices.planes_matrix);
		r.planes[i] = p0.dist < 0.0f ? p0 : p1;
	}
	Vec3f corners[8];
	bbox.get_corners_vertices(corners);
	for (int i = 0; i < 8; ++i)
		r.vertices[i] = view_matrix_transform_vertex(camera_matrices.view_matrix, corners[i]);
	return r;
}

void LeafObjectsSorter::sort_objects() // order_bboxes (:93-116): the reference's bubble sort with its non-transitive comparator
{
	if (num_objects <= 1)
		return;
	for (size_t i = 0; i + 1 < num_objects; ++i)
		for (size_t j = i + 1; j < num_objects; ++j)
			if (compare_projected_bboxes(objects[i].second, objects[j].second))
				std::swap(objects[i], objects[j]);
}

// ---- fixtures ------------------------------------------------------------------------------------------------
namespace
{
void finish_bboxes(TriangleModel& m)
{
	const uint32_t num_frames = m.meshes.empty() ? 1 : m.meshes[0].num_frames;
	m.frames_info.clear();
	for (uint32_t f = 0; f < num_frames; ++f)
	{
		bool first = true;
		BBox bbox = BBox::from_point({0, 0, 0});
		auto add = [&](Vec3f p) {
			if (first)
			{
				bbox = BBox::from_point(p);
				first = false;
			}
			else
				bbox.extend_with_point(p);
		};
		for (const TriangleModelMesh& mesh : m.meshes)
		{
			if (mesh.kind == VERTEX_DATA_NON_ANIMATED)
				for (const auto& v : mesh.non_animated)
					add(v.position);
			else if (mesh.kind == VERTEX_DATA_VERTEX_ANIMATED)
				for (size_t i = 0; i < mesh.constant.size(); ++i)
					add(mesh.variable[f * mesh.constant.size() + i].position);
			else
			{
				// animated positions of the frame: the same bone chain as animate_and_transform_triangle_mesh_vertices
				std::vector<Mat4f> matrices(m.bones.size());
				for (size_t b = 0; b < m.bones.size(); ++b)
				{
					const Mat4f mat = m.frame_bones[f * m.bones.size() + b];
					matrices[b] = m.bones[b] < b ? matrices[m.bones[b]] * mat : mat;
				}
				for (const auto& v : mesh.skeleton)
				{
					Vec4f p = {0, 0, 0, 0};
					for (int k = 0; k < 4; ++k)
						if (v.weight[k] > 0)
							p = p + (matrices[v.bone_index[k]] * extend(v.position, 1.0f)) * ((float)v.weight[k] / 255.0f);
					add(truncate(p));
				}
			}
		}
		m.frames_info.push_back(bbox);
	}
}
} // namespace

void generate_triangle_model(uint32_t kind, uint32_t seed, uint32_t detail, TriangleModel& out)
{
	out = TriangleModel{};
	Pcg32 rng(seed * 7919u + kind * 104729u + 17u);
	const uint32_t rings = std::max(detail, 3u), segments = std::max(2 * detail, 6u);
	const float pi = 3.14159265358979f;
	out.tc_shift = {-0.5f, -0.5f}; // MD3 loader convention (triangle_model_md3.rs:107)
	TriangleModelMesh mesh;
	// A lathe grid (rings + 1) x (segments + 1); two triangles per cell, wound so that the outside faces the viewer
	// (reject_triangle_model_back_faces keeps dist > 0).
	auto grid_triangles = [&]() {
		for (uint32_t r = 0; r < rings; ++r)
			for (uint32_t s = 0; s < segments; ++s)
			{
				const uint16_t a = (uint16_t)(r * (segments + 1) + s), b = (uint16_t)(a + 1), c = (uint16_t)(a + segments + 1), d = (uint16_t)(c + 1);
				mesh.triangles.push_back({a, c, b});
				mesh.triangles.push_back({b, c, d});
			}
	};
	const float size = 24.0f + 16.0f * rng.uniform();
	if (kind == 0)
	{
		mesh.kind = VERTEX_DATA_NON_ANIMATED;
		mesh.num_frames = 1;
		const float k1 = 0.35f + 0.3f * rng.uniform();
		for (uint32_t r = 0; r <= rings; ++r)
			for (uint32_t s = 0; s <= segments; ++s)
			{
				const float t = (float)r / (float)rings, a = 2.0f * pi * (float)s / (float)segments;
				const float radius = size * (0.25f + k1 * std::sin(pi * t) + 0.1f * std::sin(3.0f * pi * t));
				const float d_radius = size * (k1 * pi * std::cos(pi * t) + 0.3f * pi * std::cos(3.0f * pi * t));
				VertexNonAnimated v;
				v.position = {radius * std::cos(a), radius * std::sin(a), size * (2.0f * t - 1.0f)};
				v.normal = normalize(Vec3f{2.0f * size * std::cos(a), 2.0f * size * std::sin(a), -d_radius});
				v.tex_coord[0] = (float)s / (float)segments;
				v.tex_coord[1] = t;
				mesh.non_animated.push_back(v);
			}
		grid_triangles();
	}
	else if (kind == 1)
	{
		mesh.kind = VERTEX_DATA_VERTEX_ANIMATED;
		mesh.num_frames = 8;
		const float lobes = (float)(2 + rng.next() % 3);
		for (uint32_t r = 0; r <= rings; ++r)
			for (uint32_t s = 0; s <= segments; ++s)
				mesh.constant.push_back({{(float)s / (float)segments, (float)r / (float)rings}});
		for (uint32_t f = 0; f < mesh.num_frames; ++f)
			for (uint32_t r = 0; r <= rings; ++r)
				for (uint32_t s = 0; s <= segments; ++s)
				{
					const float theta = pi * (float)r / (float)rings, a = 2.0f * pi * (float)s / (float)segments;
					const float phase = 2.0f * pi * (float)f / (float)mesh.num_frames;
					const Vec3f dir = {std::sin(theta) * std::cos(a), std::sin(theta) * std::sin(a), -std::cos(theta)};
					const float radius = size * (1.0f + 0.25f * std::sin(lobes * theta + phase) * std::cos(lobes * a - phase));
					mesh.variable.push_back({dir * radius, dir});
				}
		grid_triangles();
	}
	else
	{
		mesh.kind = VERTEX_DATA_SKELETON_ANIMATED;
		mesh.num_frames = 8;
		const uint32_t num_bones = 6;
		const float length = 4.0f * size, seg = length / (float)num_bones, radius0 = 0.45f * size;
		for (uint32_t b = 0; b < num_bones; ++b)
			out.bones.push_back(b == 0 ? 0xFFFFFFFFu : b - 1);
		for (uint32_t r = 0; r <= rings; ++r)
			for (uint32_t s = 0; s <= segments; ++s)
			{
				const float t = (float)r / (float)rings, a = 2.0f * pi * (float)s / (float)segments;
				const float radius = radius0 * (1.0f - 0.7f * t);
				SkeletonAnimatedVertex v{};
				v.position = {radius * std::cos(a), radius * std::sin(a), length * t};
				v.normal = normalize(Vec3f{std::cos(a), std::sin(a), 0.7f * radius0 / length});
				v.tex_coord[0] = (float)s / (float)segments;
				v.tex_coord[1] = t;
				const float fb = std::fmin(t * (float)num_bones, (float)num_bones - 0.001f);
				const uint32_t b0 = (uint32_t)fb, b1 = std::min(b0 + 1, num_bones - 1);
				const uint32_t w1 = b1 == b0 ? 0u : (uint32_t)(255.0f * (fb - (float)b0)), w0 = 255u - w1;
				v.bone_index[0] = (uint8_t)b0;
				v.weight[0] = (uint8_t)w0;
				v.bone_index[1] = (uint8_t)b1;
				v.weight[1] = (uint8_t)w1;
				mesh.skeleton.push_back(v);
			}
		grid_triangles();
		// Frame matrices relative to the parent's skinning matrix, so that the product along the chain (what the renderer
		// builds, triangle_models_rendering.rs:122-165) is `animated pose x inverse bind pose` of the bone.
		const float bend = 0.25f + 0.2f * rng.uniform();
		for (uint32_t f = 0; f < mesh.num_frames; ++f)
		{
			const float phase = 2.0f * pi * (float)f / (float)mesh.num_frames;
			Mat4f animated = Mat4f::identity(), prev_skin = Mat4f::identity();
			for (uint32_t b = 0; b < num_bones; ++b)
			{
				const float angle = bend * std::sin(phase + 0.9f * (float)b);
				Mat4f rot = Mat4f::identity(); // rotation about x
				rot.y = {0, std::cos(angle), std::sin(angle), 0};
				rot.z = {0, -std::sin(angle), std::cos(angle), 0};
				animated = animated * (b == 0 ? Mat4f::identity() : Mat4f::from_translation({0, 0, seg})) * rot;
				const Mat4f skin = animated * Mat4f::from_translation({0, 0, -seg * (float)b});
				Mat4f prev_inv = Mat4f::identity();
				prev_skin.invert(prev_inv);
				out.frame_bones.push_back(b == 0 ? skin : prev_inv * skin);
				prev_skin = skin;
			}
		}
	}
	out.meshes.push_back(std::move(mesh));
	finish_bboxes(out);
}

} // namespace sqwh
