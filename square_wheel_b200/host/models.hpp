// Triangle models of the host-side mirror: data model (src/square_wheel/triangle_model.rs), per-frame preparation
// (triangle_models_rendering.rs: animation, per-vertex light-cube lighting, back-face rejection, depth sort) and the
// ordering of dynamic objects inside a leaf (draw_ordering.rs).  The prepared meshes are turned into SQW_CMD_TRIANGLE
// commands by Renderer::draw_mesh (partial_renderer.rs:3525-3717); the pixels are the GPU rasteriser's job (§8 row a7).
#pragma once
#include "map.hpp"
#include <vector>

namespace sqwh
{

struct Mat3f // cgmath Matrix3 (columns)
{
	Vec3f x, y, z;
	static Mat3f identity() { return {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; }
	Mat3f transpose() const { return {{x.x, y.x, z.x}, {x.y, y.y, z.y}, {x.z, y.z, z.z}}; }
	bool invert(Mat3f& out) const;
};
inline Vec3f operator*(const Mat3f& m, Vec3f v) { return m.x * v.x + m.y * v.y + m.z * v.z; }
inline Mat3f operator*(const Mat3f& m, float s) { return {m.x * s, m.y * s, m.z * s}; }
inline Mat3f operator*(const Mat3f& a, const Mat3f& b) { return {a * b.x, a * b.y, a * b.z}; }
inline Mat3f operator+(const Mat3f& a, const Mat3f& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Mat4f operator*(const Mat4f& m, float s) { return {m.x * s, m.y * s, m.z * s, m.w * s}; }
inline Mat4f operator+(const Mat4f& a, const Mat4f& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

// ---- frame_info.rs:55-110 ------------------------------------------------------------------------------------
enum ModelLightingKind : uint32_t
{
	MODEL_LIGHTING_DEFAULT = 0,
	MODEL_LIGHTING_CONSTANT = 1,
	MODEL_LIGHTING_ADVANCED = 2,
};
constexpr uint32_t MODEL_FLAG_VIEW_MODEL = 1, MODEL_FLAG_ONLY_THIRD_PERSON_VIEW = 2;
struct ModelEntity
{
	Vec3f position;
	Quaternion rotation;
	uint32_t frames[2] = {0, 0}; // AnimationPoint
	float lerp = 0.0f;
	uint32_t model = 0;                  // index into Scene::models
	uint32_t texture_images[NUM_MIPS] = {0, 0, 0, 0, 0}; // TextureLiteWithMips: SQW_IMAGE_COLOR32 image ids
	int blending_mode = SQW_BLEND_NONE;
	ModelLightingKind lighting = MODEL_LIGHTING_DEFAULT;
	float light[3] = {0, 0, 0};          // ConstantLight value / AdvancedLight::light_add
	float grid_light_scale = 0.0f;       // AdvancedLight
	Vec3f light_position;                // AdvancedLight::position
	uint32_t flags = 0;
	bool has_ordering_custom_bbox = false;
	BBox ordering_custom_bbox;
};

// frame_info.rs:125-147
enum SpriteOrientation : uint32_t
{
	SPRITE_PARALLEL_TO_CAMERA_PLANE = 0,
	SPRITE_FACING_TOWARDS_CAMERA = 1,
	SPRITE_ALIGN_TO_Z_AXIS_PARALLEL_TO_CAMERA_PLANE = 2,
	SPRITE_ALIGN_TO_Z_AXIS_FACING_TOWARDS_CAMERA = 3,
};
struct Sprite
{
	Vec3f position;
	float angle = 0.0f;
	float radius = 1.0f;
	uint32_t texture_images[NUM_MIPS] = {0, 0, 0, 0, 0}; // TextureLiteWithMips: SQW_IMAGE_COLOR32 image ids
	int blending_mode = SQW_BLEND_NONE;
	SpriteOrientation orientation = SPRITE_PARALLEL_TO_CAMERA_PLANE;
	float light_scale = 0.0f;
	float light_add[3] = {0, 0, 0};
};
constexpr uint32_t MAX_SPRITE_TESSELATION_LEVEL = 4; // renderer_structs.rs:164

// frame_info.rs:112-123: the decal primitive is a cube with half-size 1, textured along its -X axis
struct Decal
{
	Vec3f position;
	Quaternion rotation;
	Vec3f scale = {1, 1, 1};
	uint32_t texture_images[NUM_MIPS] = {0, 0, 0, 0, 0}; // TextureLiteWithMips: SQW_IMAGE_COLOR32 image ids
	int blending_mode = SQW_BLEND_NONE;
	float lightmap_light_scale = 0.0f;
	float light_add[3] = {0, 0, 0};
};

// frame_info.rs:171-205
enum PortalViewKind : uint32_t
{
	PORTAL_VIEW_CAMERA_AT_POSITION = 0,
	PORTAL_VIEW_MIRROR = 1,
	PORTAL_VIEW_PARALLAX_PORTAL = 2,
};
struct ViewPortal
{
	PortalViewKind view = PORTAL_VIEW_MIRROR;
	Vec3f camera_position;            // CameraAtPosition
	Quaternion camera_rotation;
	float camera_fov = 1.0f;          // radians
	Mat4f transform_matrix = Mat4f::identity(); // ParallaxPortal
	Plane plane;
	Plane tex_coord_equation[2];
	std::vector<Vec3f> vertices;
	int blending_mode = SQW_BLEND_NONE;
	bool has_texture = false;         // ViewPortalTexture
	int texture_blending_mode = SQW_BLEND_AVERAGE;
	uint32_t texture_images[NUM_MIPS] = {0, 0, 0, 0, 0};
	float texture_light_scale = 0.0f;
	float texture_light_add[3] = {1, 1, 1};
};
// renderer_structs.rs:234-243
struct PortalInfo
{
	uint32_t resolution[2] = {0, 0}; // 0 area: not visible this frame
	float depth_equation[3] = {0, 0, 0};
	float tex_coord_equation[6] = {0, 0, 0, 0, 0, 0};
	int32_t tc_min[2] = {0, 0};
	uint32_t mip = 0;
};

// renderer_structs.rs:311-318 / triangle_models_rendering.rs:376-424
struct ModelVertex3d
{
	Vec3f pos;
	Vec2f tc;
	float light[3];
};
struct ModelVertex2d
{
	Vec2f pos;
	Vec2f tc;
	float light[3];
};
struct ModelLightData
{
	float light_cube[6][3];
	Vec3f vector_scaled[2];
	float color[2][3];
	static ModelLightData from_two_light_grid_elements(const LightGridElement& a, const LightGridElement& b);
};

// common/light_cube.rs
struct LightCube
{
	float light_cube[6][3] = {};
	Vec3f sum_scaled_light_dirs[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	void add_light_sample(Vec3f direction, const float light_scaled[3]);
	void add_constant_light(const float light[3]);
	void scale(float s);
	LightGridElement convert_into_light_grid_sample() const;
};

BBox get_current_triangle_model_bbox(const TriangleModel& model, const uint32_t frames[2]);
Mat3f get_normals_matrix(const Mat4f& model_matrix);
void animate_and_transform_triangle_mesh_vertices(
	const TriangleModel& model, const TriangleModelMesh& mesh, const uint32_t frames[2], float lerp, const ModelLightData& light, const Mat4f& model_matrix,
	const Mat4f& model_view_matrix, Vec2f tc_scale, Vec2f tc_shift, ModelVertex3d* dst_vertices);
uint32_t calculate_triangle_model_texture_mip(const Mat4f& model_view_matrix, const BBox& model_bbox, const uint32_t texture_size[2], float mip_bias);
size_t reject_triangle_model_back_faces(const ModelVertex3d* transformed_vertices, const Triangle* triangles, size_t num_triangles, Triangle* out_triangles);
// Back to front by the largest z of the triangle.  The reference uses Rust's unstable sort; the order of triangles with
// EQUAL keys is implementation-defined there — here ties keep the mesh order (stable).
void sort_model_triangles(const ModelVertex3d* transformed_vertices, Triangle* triangles, size_t num_triangles);
size_t clip_3d_model_polygon_by_plane(const ModelVertex3d* polygon, size_t n, const Plane& plane, ModelVertex3d* out_polygon, size_t out_capacity);
size_t clip_2d_model_polygon(const ModelVertex2d* polygon, size_t n, Vec3f clip_line_equation, ModelVertex2d* out_polygon, size_t out_capacity);

// ---- draw_ordering.rs ----------------------------------------------------------------------------------------
struct ProjectedBBox
{
	Plane planes[3];
	Vec3f vertices[8];
};
ProjectedBBox project_bbox(const BBox& bbox, const CameraMatrices& camera_matrices);
constexpr size_t MAX_LEAF_OBJECTS = 48; // draw_ordering.rs:50
struct LeafObjectsSorter
{
	std::pair<uint32_t, ProjectedBBox> objects[MAX_LEAF_OBJECTS];
	size_t num_objects = 0;
	void clear() { num_objects = 0; }
	void add_object(uint32_t id, const ProjectedBBox& b)
	{
		if (num_objects < MAX_LEAF_OBJECTS)
			objects[num_objects++] = {id, b};
	}
	void sort_objects();
};

} // namespace sqwh
