// Data model of triangle models (src/square_wheel/triangle_model.rs) and the bounding box type (src/common/bbox.rs); kept apart
// from models.hpp so that Scene (map.hpp) can own the model fixtures.
#pragma once
#include "math.hpp"
#include <array>
#include <cmath>
#include <string>
#include <vector>

namespace sqwh
{

struct BBox // common/bbox.rs
{
	Vec3f min, max;
	static BBox from_min_max(Vec3f mn, Vec3f mx) { return {mn, mx}; }
	static BBox from_point(Vec3f p) { return {p, p}; }
	Vec3f get_center() const { return (min + max) * 0.5f; }
	Vec3f get_size() const { return max - min; }
	void get_corners_vertices(Vec3f out[8]) const // bbox.rs:83-95
	{
		out[0] = {min.x, min.y, min.z};
		out[1] = {min.x, min.y, max.z};
		out[2] = {min.x, max.y, min.z};
		out[3] = {min.x, max.y, max.z};
		out[4] = {max.x, min.y, min.z};
		out[5] = {max.x, min.y, max.z};
		out[6] = {max.x, max.y, min.z};
		out[7] = {max.x, max.y, max.z};
	}
	void extend(const BBox& o)
	{
		extend_with_point(o.min);
		extend_with_point(o.max);
	}
	void extend_with_point(Vec3f p)
	{
		min = {std::fmin(min.x, p.x), std::fmin(min.y, p.y), std::fmin(min.z, p.z)};
		max = {std::fmax(max.x, p.x), std::fmax(max.y, p.y), std::fmax(max.z, p.z)};
	}
};

// ---- triangle_model.rs ---------------------------------------------------------------------------------------
struct VertexNonAnimated
{
	Vec3f position, normal;
	float tex_coord[2];
};
struct VertexAnimatedVertexConstant
{
	float tex_coord[2];
};
struct VertexAnimatedVertexVariable
{
	Vec3f position, normal;
};
struct SkeletonAnimatedVertex
{
	float tex_coord[2];
	Vec3f position, normal;
	uint8_t bone_index[4], weight[4]; // bones_description
};
typedef std::array<uint16_t, 3> Triangle; // VertexIndex = u16
constexpr size_t MAX_TRIANGLE_MODEL_BONES = 255;

enum VertexDataKind : uint32_t
{
	VERTEX_DATA_NON_ANIMATED = 0,
	VERTEX_DATA_VERTEX_ANIMATED = 1,
	VERTEX_DATA_SKELETON_ANIMATED = 2,
};
struct TriangleModelMesh
{
	std::string name, material_name; // only file loaders fill these
	std::vector<Triangle> triangles;
	uint32_t num_frames = 1;
	VertexDataKind kind = VERTEX_DATA_NON_ANIMATED;
	std::vector<VertexNonAnimated> non_animated;
	std::vector<VertexAnimatedVertexConstant> constant; // VertexAnimated
	std::vector<VertexAnimatedVertexVariable> variable; // VertexAnimated: num_frames x constant.size()
	std::vector<SkeletonAnimatedVertex> skeleton;
	size_t num_vertices() const { return kind == VERTEX_DATA_NON_ANIMATED ? non_animated.size() : (kind == VERTEX_DATA_VERTEX_ANIMATED ? constant.size() : skeleton.size()); }
};
struct TriangleModelAnimation // triangle_model.rs:18-26
{
	std::string name;
	uint32_t start_frame = 0, num_frames = 0;
	float frames_per_second = 0.0f;
	bool looped = false;
};
struct TriangleModel
{
	std::vector<TriangleModelAnimation> animations;
	std::vector<std::string> bone_names; // TriangleModelBoneInfo::name (file loaders only)
	std::vector<BBox> frames_info; // TriangleModelFrameInfo::bbox
	std::vector<TriangleModelMesh> meshes;
	std::vector<uint32_t> bones;     // TriangleModelBoneInfo::parent (>= bones.size(): none)
	std::vector<Mat4f> frame_bones;  // num_frames x bones.size()
	Vec2f tc_shift = {0.0f, 0.0f};
};

// Seeded in-memory fixtures (no model file ships with the reference except one IQM):
//   kind 0: static lamp-like solid of revolution (NonAnimated), kind 1: wobbling blob (VertexAnimated, MD3-class),
//   kind 2: bending tentacle with a 6-bone chain (SkeletonAnimated, IQM-class).
void generate_triangle_model(uint32_t kind, uint32_t seed, uint32_t detail, TriangleModel& out);

// File loaders (model_file.cpp): triangle_model_md3.rs / triangle_model_iqm.rs on a byte image of the file.  false + message
// where the reference returns Ok(None) or an io::Error.
bool load_model_md3(const uint8_t* data, size_t size, TriangleModel& out, std::string* err);
bool load_model_iqm(const uint8_t* data, size_t size, TriangleModel& out, std::string* err);
bool load_model(const uint8_t* data, size_t size, TriangleModel& out, std::string* err); // by magic
bool load_model_file(const char* path, TriangleModel& out, std::string* err);

// bsp_map_compact.rs:179-189
struct LightGridElement
{
	float light_cube[6][3] = {};
	Vec3f light_direction_vector_scaled = {0, 0, 0};
	float directional_light_color[3] = {0, 0, 0};
};

} // namespace sqwh
