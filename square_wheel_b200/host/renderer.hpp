// Host-side mirror of the reference's renderer front end: everything that "stays host" in SURVEY.md §8 (rows a1, a2, a5):
// portal-flood visibility, polygon surface preparation, job list building, BSP traversal and polygon set-up — emitting
// sqw_surface_job / sqw_draw_cmd lists for the C-ABI instead of calling the CPU kernels.
//
//   Renderer::prepare_frame  <->  Renderer::prepare_frame::<ColorT>  (src/square_wheel/renderer.rs:61)
//                                 -> PartialRenderer::prepare_frame  (partial_renderer.rs:156)
//   Renderer::draw_frame     <->  Renderer::draw_frame::<ColorT>     (renderer.rs:106)
//                                 -> PartialRenderer::draw_frame / perform_rasterization (partial_renderer.rs:236, :1192)
#pragma once
#include "map.hpp"
#include "models.hpp"
#include <memory>
#include <vector>

namespace sqwh
{

// common/clipping_polygon.rs
struct ClipAxis
{
	float min = 0, max = 0;
};
struct ClippingPolygon
{
	ClipAxis x, y, x_plus_y, x_minus_y;
	static ClippingPolygon from_box(float min_x, float min_y, float max_x, float max_y);
	static ClippingPolygon from_point(Vec2f p);
	bool is_valid_and_non_empty() const;
	bool contains(const ClippingPolygon& o) const;
	void get_box_clip_planes(Vec3f out[4]) const; // clipping_polygon.rs:154-162
	void extend(const ClippingPolygon& o);
	void extend_with_point(Vec2f p);
	void intersect(const ClippingPolygon& o);
	void get_clip_planes(Vec3f out[8]) const;
};

// renderer_config.rs:5-31 (fields the path reads)
struct RendererConfig
{
	bool use_directional_lightmaps = true;
	uint32_t portals_depth = 1; // renderer_config.rs `portals_depth`: nested renderers for portals / mirrors (renderer.rs:33 caps it at 8)
	float textures_mip_bias = 0.0f;
	float shadows_quality = 0.0f;
	bool clear_background = true;
	bool invert_polygons_order = false;
};

// renderer_structs.rs:75-91
struct DrawPolygonData
{
	uint32_t parent_leaf = 0;
	uint32_t parent_submodel = 0xFFFFFFFFu; // DrawPolygonParent::Submodel(index) when < number of submodels (renderer_structs.rs:93-98)
	Vec3f basis_u, basis_v, basis_start, basis_normal; // PolygonBasisVecs for mip 0 (surfaces.rs:7-14)
	uint32_t visible_frame = 0;
	float depth_equation[3] = {0, 0, 0};
	float tex_coord_equation[6] = {0, 0, 0, 0, 0, 0};
	size_t surface_pixels_offset = 0;
	uint32_t surface_size[2] = {0, 0};
	uint32_t mip = 0;
	int32_t surface_tc_min[2] = {0, 0};
};

// renderer_structs.rs:36-54 (units of the texels/s metric)
struct RendererDebugStats
{
	uint32_t num_visible_leafs = 0;
	uint32_t num_visible_polygons = 0;
	uint64_t num_surfaces_pixels = 0;
	uint32_t num_draw_commands = 0;
	uint32_t mode_counts[3] = {0, 0, 0};
	uint32_t num_visible_meshes = 0;
	uint32_t num_triangles = 0; // SQW_CMD_TRIANGLE commands recorded
	uint32_t num_visible_sprites = 0;
	uint32_t num_decal_triangles = 0;
};

// frame_info.rs:149-168
struct DynamicLight
{
	Vec3f position;
	float radius = 0;
	float color[3] = {0, 0, 0};
	uint32_t shadow_type = SQW_SHADOW_NONE; // SQW_SHADOW_*
	Quaternion rotation;                    // Projector only
	float fov = 0;                          // Projector only, radians
};

// renderer_structs.rs:207-213
struct DynamicLightInfo
{
	bool visible = false;
	uint32_t shadow_map_size = 0;
	size_t shadow_map_data_offset = 0;
};

// map_visibility_calculator.rs: portal-flood visibility from one view point.  The renderer owns one for the camera and one
// for the shadow-map views (DepthRenderer::visibility_calculator, depth_renderer.rs:11).
class VisibilityCalculator
{
public:
	explicit VisibilityCalculator(const BSPMap& map);
	void update_visibility(const CameraMatrices& camera_matrices, const ClippingPolygon& frame_bounds);
	// for portals / mirrors, whose camera can be far from where the search must start (map_visibility_calculator.rs:101-113)
	void update_visibility_with_start_leafs(const CameraMatrices& camera_matrices, const ClippingPolygon& frame_bounds, const uint32_t* start_leafs, size_t n);
	bool leaf_bounds(uint32_t leaf, ClippingPolygon& out) const; // get_current_frame_leaf_bounds
	bool is_inside_leaf_volume() const { return is_inside_leaf_volume_; }

private:
	uint32_t find_current_leaf() const;
	void mark_reachable_leafs_iterative(const uint32_t* start_leafs, size_t num_start_leafs, const ClippingPolygon& start_bounds);
	bool compute_is_inside_leaf_volume(uint32_t leaf_index) const;
	struct LeafData
	{
		uint32_t visible_frame = 0;
		ClippingPolygon current_frame_bounds;
	};
	struct PortalData
	{
		uint32_t visible_frame = 0;
		bool has_projection = false;
		ClippingPolygon current_frame_projection;
	};
	const BSPMap& map_;
	CameraMatrices camera_matrices_;
	uint32_t visibility_frame_ = 0;
	std::vector<LeafData> leafs_data_;
	std::vector<PortalData> portals_data_;
	std::vector<uint32_t> wave_cur_, wave_next_;
	bool is_inside_leaf_volume_ = true;
};

class Renderer
{
public:
	Renderer(const Scene& scene, const RendererConfig& config);

	// Visibility + surface preparation + job list.  color_bits only selects nothing here (the lists are colour-agnostic) but
	// is kept for symmetry with prepare_frame::<ColorT>.
	void prepare_frame(uint32_t width, uint32_t height, const CameraPose& camera, float fov);
	// PartialRenderer::prepare_frame (partial_renderer.rs:156) with explicit matrices: what nested portal / mirror views use.
	void prepare_frame_with_matrices(uint32_t width, uint32_t height, const CameraMatrices& camera_matrices, bool is_third_person_view, const uint32_t* start_leafs, size_t num_start_leafs);
	// Sky box (if any polygon of a sky material is in view), then BSP traversal -> ordered draw commands for the frame prepared
	// last.  color_bits picks the sky box textures of the target's colour type (SkyboxSideTexture<ColorT>).
	void draw_frame(int color_bits = 64);

	const std::vector<sqw_surface_job>& jobs() const { return jobs_; }
	uint64_t total_surface_pixels() const { return num_visible_surfaces_pixels_; }
	const std::vector<sqw_draw_cmd>& cmds() const { return cmds_; }
	const std::vector<sqw_polygon_point>& polygon_points() const { return polygon_points_; }
	const CameraMatrices& camera_matrices() const { return camera_matrices_; }
	bool needs_background_clear() const { return config_.clear_background && !visibility_.is_inside_leaf_volume(); }
	// FrameInfo::lights for the next prepare_frame (frame_info.rs:24)
	void set_lights(const DynamicLight* lights, size_t n) { frame_lights_.assign(lights, lights + n); }
	// FrameInfo::model_entities for the next prepare_frame (frame_info.rs:14)
	void set_models(const ModelEntity* models, size_t n) { frame_models_.assign(models, models + n); }
	// FrameInfo::submodel_entities for the next prepare_frame (frame_info.rs:13): one optional entity per map submodel.  Before
	// the first call every submodel stands where the map compiler left it (position = bbox centre, no rotation).
	struct SubmodelEntity
	{
		bool present = true;
		Vec3f position;
		Quaternion rotation;
	};
	void set_submodel_entities(const SubmodelEntity* e, size_t n);
	size_t num_submodels() const { return map_.submodels.size(); }
	BBox submodel_bbox(size_t i) const { return submodels_bbox_[i]; }
	// FrameInfo::portals for the next prepare_frame (frame_info.rs:18)
	void set_portals(const ViewPortal* portals, size_t n) { frame_portals_.assign(portals, portals + n); }
	const std::vector<PortalInfo>& portals_info() const { return portals_info_; }
	const std::vector<ViewPortal>& portals() const { return frame_portals_; }
	// The nested view of portal i of the frame prepared last: build_view_portals_textures (partial_renderer.rs:2008-2225).  Prepares
	// and records the child renderer's frame and returns it (valid until the next call), or nullptr when the portal is not visible
	// or the depth limit is reached (the caller then fills the portal texture with black, :2038-2044).
	Renderer* prepare_portal_view(size_t portal_index, int color_bits);
	// light the portal's overlay texture is mixed with (:2227-2262)
	void portal_texture_light(size_t portal_index, float out[3]) const;
	// FrameInfo::decals for the next prepare_frame (frame_info.rs:16)
	void set_decals(const Decal* decals, size_t n) { frame_decals_.assign(decals, decals + n); }
	// FrameInfo::sprites for the next prepare_frame (frame_info.rs:17)
	void set_sprites(const Sprite* sprites, size_t n) { frame_sprites_.assign(sprites, sprites + n); }
	const std::vector<sqw_triangle_point>& triangle_points() const { return triangle_points_; }
	const std::vector<sqw_light>& lights() const { return lights_; }
	const std::vector<sqw_depth_map>& depth_maps() const { return depth_maps_; }
	const std::vector<sqw_depth_cmd>& depth_cmds() const { return depth_cmds_; }
	const std::vector<sqw_polygon_point>& depth_points() const { return depth_points_; }
	uint64_t shadow_maps_len() const { return shadow_maps_len_; }
	const RendererDebugStats& stats() const { return stats_; }
	// Inputs and results of build_dynamic_models_buffers (partial_renderer.rs:1122-1190) for the frame prepared last, one record
	// per visible mesh: what the GPU model-preparation kernels (csrc/model_kernels.cu, sqw_models_prepare) consume, and this
	// mirror's own vertices / sorted triangles for the parity tests.
	struct DynamicMeshExport
	{
		uint32_t model, mesh, frames[2];
		float lerp;
		float model_matrix[16], model_view_matrix[16]; // column-major
		float tc_scale[2], tc_shift[2];
		float light_cube[6][3], light_vector_scaled[2][3], light_color[2][3]; // ModelLightData
		uint32_t vertices_offset, num_vertices, triangles_offset, num_triangles, num_visible_triangles;
	};
	size_t export_dynamic_meshes(DynamicMeshExport* out, size_t capacity) const;
	// Inputs and results of prepare_polygons_surfaces for the leaf polygons of the frame prepared last (what sqw_polygons_prepare
	// computes on the GPU; tests compare the two): the (polygon, clip set) pairs in the order they were prepared, eight planes per set.
	struct PolygonCandidate
	{
		uint32_t polygon, clip_set;
	};
	struct PreparedPolygonExport // layout of sqw_prepared_polygon
	{
		uint32_t flags, mip;
		float depth_equation[3];
		float tex_coord_equation[6];
		int32_t surface_tc_min[2];
		uint32_t surface_size[2];
		uint32_t surface_pixels_offset;
	};
	struct PolygonsFrameExport
	{
		float planes_matrix[16], view_matrix[16];
		float mip_bias;
		uint32_t has_sky, sky_material;
		float sky_bounds[8];
		uint32_t num_visible;
		uint64_t num_surfaces_pixels;
	};
	const std::vector<PolygonCandidate>& polygon_candidates() const { return polygon_candidates_; }
	const std::vector<Vec3f>& polygon_clip_planes() const { return polygon_clip_planes_; }
	const std::vector<uint32_t>& polygon_draw_order() const { return polygon_draw_order_; }
	// lights located in the leaf of every clip set (leaf_lights_), concatenated: offsets [sets + 1], indices into lights()
	void export_clip_set_lights(std::vector<uint32_t>& offsets, std::vector<uint16_t>& list) const
	{
		offsets.assign(1, 0u);
		list.clear();
		for (uint32_t leaf : polygon_clip_set_leafs_)
		{
			for (uint32_t l : leaf_lights_[leaf])
				list.push_back((uint16_t)l);
			offsets.push_back((uint32_t)list.size());
		}
	} // leaf polygons in the order draw_leaf drew them (last draw_frame)
	void export_prepared_polygons(PreparedPolygonExport* out, PolygonsFrameExport* frame) const;
	uint32_t export_polygon_basis(float* out, uint32_t capacity) const
	{
		const uint32_t n = (uint32_t)polygons_data_.size();
		for (uint32_t i = 0; out && i < n && i < capacity; ++i)
		{
			const DrawPolygonData& d = polygons_data_[i];
			const Vec3f v[4] = {d.basis_u, d.basis_v, d.basis_start, d.basis_normal};
			for (int k = 0; k < 4; ++k)
			{
				out[i * 12 + k * 3] = v[k].x;
				out[i * 12 + k * 3 + 1] = v[k].y;
				out[i * 12 + k * 3 + 2] = v[k].z;
			}
		}
		return n;
	}
	double polygons_prepare_seconds() const { return polygons_prepare_seconds_; } // wall time of prepare_polygons_surfaces in the last prepare_frame
	const std::vector<ModelVertex3d>& dynamic_meshes_vertices() const { return dynamic_meshes_vertices_; }
	const std::vector<Triangle>& dynamic_meshes_triangles() const { return dynamic_meshes_triangles_; }

private:
	// dynamic_objects_index.rs / partial_renderer.rs:398-590 / depth_renderer.rs
	void position_dynamic_lights();
	void position_object_sphere_r(uint32_t id, Vec3f center, float radius, uint32_t node_index);
	void position_object_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index);
	void prepare_dynamic_lights();
	void build_shadow_maps();
	void depth_draw_map(uint32_t offset, uint32_t size, uint32_t flags, const CameraMatrices& depth_matrices);
	void depth_draw_polygon(const CameraMatrices& m, const Vec3f clip_planes[8], uint32_t polygon_index, bool depth_test);
	// triangle models: dynamic_objects_index.rs:52-73, partial_renderer.rs:995-1190, :3525-3717, triangle_models_rendering.rs
	void position_models();
	void position_model_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index);
	void prepare_dynamic_models();
	void build_dynamic_models_buffers();
	ModelLightData get_model_light(const ModelEntity& model, const Mat4f& model_matrix) const;
	LightGridElement fetch_light_from_grid(Vec3f pos) const; // triangle_models_rendering.rs:574-699
	void draw_leaf_objects(const ClippingPolygon& bounds, uint32_t leaf_index);
	// portals: dynamic_objects_index.rs:143-157, partial_renderer.rs:1834-1996, :3720-3776
	void position_portals();
	void position_portal_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index);
	void prepare_view_portals_textures();
	void draw_view_portal(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t portal_index);
	// submodels: inline_models_index.rs, partial_renderer.rs:593-660, :1362-1447, :3098-3288, depth_renderer.rs:130-200
	void position_submodels();
	void position_submodel_r(uint32_t model_index, const Vec3f bbox_vertices[8], uint32_t node_index);
	bool get_submodel_matrix(size_t index, Mat4f& out) const;
	void prepare_submodels();
	void draw_submodel_in_leaf(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t submodel_index);
	void draw_submodel_bsp_node_r(
		Vec4f submodel_planes_matrix_w, const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t node_index);
	void draw_submodel_polygon(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t leaf_index, uint32_t polygon_index);
	// decals: dynamic_objects_index.rs:76-91, partial_renderer.rs:665-737, :2794-3096, :3807-3875
	void position_decals();
	void position_decal_convex_hull_r(uint32_t id, const Vec3f* vertices, size_t n, uint32_t node_index);
	void prepare_decals();
	void draw_polygon_decals(const Vec3f clip_planes[8], uint32_t polygon_index, const std::vector<uint32_t>& current_decals);
	struct DecalDrawContext;
	void subdivide_and_draw_decal_triangle(const DecalDrawContext& ctx, const Vec2f points[3], size_t recursion_depth);
	void get_polygon_lightmap_light(const Polygon& polygon, Vec2f lightmap_coord, float out[3]) const;
	// sprites: dynamic_objects_index.rs:93-106, partial_renderer.rs:740-990, :3290-3523
	void position_sprites();
	void position_sprite_sphere_r(uint32_t id, Vec3f center, float radius, uint32_t node_index);
	void prepare_sprites();
	void draw_sprite(const Vec3f clip_planes[8], const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, uint32_t sprite_index);
	void emit_triangle_fan(const ModelVertex2d* vertices, size_t num_vertices, const float* constant_light, uint32_t image_id, int blending_mode);
	void draw_mesh(const ClippingPolygon& clipping_polygon, const Plane* leaf_clip_planes, size_t num_leaf_clip_planes, size_t visible_mesh_index);
	// partial_renderer.rs
	void prepare_polygons_surfaces();
	uint32_t prepare_polygon_surface(const CameraMatrices& camera_matrices, const Vec3f clip_planes[8], size_t polygon_index);
	void build_surface_jobs();
	void draw_skybox(const ClippingPolygon& viewport_clipping_polygon, int color_bits);
	void draw_tree(const ClippingPolygon& viewport_clipping_polygon);
	void draw_view_models(const ClippingPolygon& viewport_clipping_polygon);
	void draw_leaf(const ClippingPolygon& bounds, uint32_t leaf_index);
	void draw_polygon(
		const Vec3f clip_planes[8], const Vec3f* vertices_transformed, size_t num_vertices, const float depth_equation[3], const float tex_coord_equation[6],
		const uint32_t texture_size[2], uint8_t tex_kind, uint32_t tex_ref, int blending_mode);

	const Scene& scene_;
	const BSPMap& map_;
	RendererConfig config_;
	CameraMatrices camera_matrices_;
	uint32_t width_ = 0, height_ = 0;
	uint32_t current_frame_ = 0;
	float mip_bias_ = 0;

	VisibilityCalculator visibility_;
	VisibilityCalculator shadow_visibility_;

	// FrameInfo::lights and what prepare_dynamic_lights / build_shadow_maps derive from them
	std::vector<DynamicLight> frame_lights_;
	std::vector<DynamicLightInfo> dynamic_lights_info_;
	std::vector<std::vector<uint32_t>> leaf_lights_;  // DynamicObjectsIndex::get_leaf_objects
	std::vector<std::vector<uint32_t>> light_leafs_;  // DynamicObjectsIndex::get_object_leafs
	std::vector<sqw_light> lights_;
	std::vector<sqw_depth_map> depth_maps_;
	std::vector<sqw_depth_cmd> depth_cmds_;
	std::vector<sqw_polygon_point> depth_points_;
	uint64_t shadow_maps_len_ = 0;

	// renderer_structs.rs:216-232 (VisibleDynamicMeshInfo) and the model buffers of PartialRenderer
	struct VisibleDynamicMeshInfo
	{
		uint32_t entity_index = 0, mesh_index = 0;
		size_t vertices_offset = 0, triangles_offset = 0, num_visible_triangles = 0;
		Vec3f bbox_vertices_transformed[8];
		ClippingPolygon clipping_polygon;
		Mat4f model_matrix;
		CameraMatrices camera_matrices;
		uint32_t mip = 0;
	};
	struct DynamicModelInfo
	{
		uint32_t first_visible_mesh = 0, num_visible_meshes = 0;
	};
	std::vector<ModelEntity> frame_models_;
	std::vector<std::vector<uint32_t>> leaf_models_, model_leafs_; // DynamicObjectsIndex for models
	std::vector<std::vector<Plane>> leafs_planes_;                 // RenderersCommonData::leafs_planes (renderer.rs:49)
	std::vector<VisibleDynamicMeshInfo> visible_dynamic_meshes_list_;
	std::vector<DynamicModelInfo> dynamic_model_to_dynamic_meshes_index_;
	std::vector<ModelVertex3d> dynamic_meshes_vertices_;
	std::vector<Triangle> dynamic_meshes_triangles_;
	struct SpriteInfo // renderer_structs.rs:199-205
	{
		Vec3f vertices_projected[4];
		float light[3] = {0, 0, 0};
		uint32_t mip = 0, tesselation_level = 0;
	};
	struct SubmodelInfo // renderer_structs.rs:170-189 (VisibleSubmodelInfo / VisibleSubmodelMatrices)
	{
		bool has_matrices = false;
		Mat4f world_planes_matrix;
		CameraMatrices camera_matrices;
		std::vector<uint32_t> dynamic_lights;
	};
	std::vector<BBox> submodels_bbox_;          // InlineModelsIndex::models_info[i].bbox (extended by 1/4)
	std::vector<SubmodelEntity> submodel_entities_;
	std::vector<std::vector<uint32_t>> leaf_submodels_, submodel_leafs_;
	std::vector<SubmodelInfo> submodels_info_;
	std::vector<ViewPortal> frame_portals_;
	std::vector<std::vector<uint32_t>> leaf_portals_, portal_leafs_; // DynamicObjectsIndex for portals
	std::vector<PortalInfo> portals_info_;
	std::unique_ptr<Renderer> portal_renderer_; // PortalsRenderingData::renderer: the nested PartialRenderer, depth - 1
	std::vector<uint32_t> transformed_portal_leafs_;
	bool is_third_person_view_ = false;
	struct DecalInfo // renderer_structs.rs:191-197
	{
		Mat4f camera_planes_matrix;
		LightGridElement dynamic_light;
	};
	std::vector<Decal> frame_decals_;
	std::vector<std::vector<uint32_t>> leaf_decals_, decal_leafs_; // DynamicObjectsIndex for decals
	std::vector<DecalInfo> decals_info_;
	std::vector<Sprite> frame_sprites_;
	std::vector<std::vector<uint32_t>> leaf_sprites_, sprite_leafs_; // DynamicObjectsIndex for sprites
	std::vector<SpriteInfo> sprites_info_;
	std::vector<sqw_triangle_point> triangle_points_;
	LeafObjectsSorter objects_sorter_;

	bool has_current_sky_ = false; // PartialRenderer::current_sky (partial_renderer.rs:35)
	uint32_t current_sky_material_ = 0;
	ClippingPolygon current_sky_bounds_;

	std::vector<DrawPolygonData> polygons_data_;
	double polygons_prepare_seconds_ = 0.0;
	std::vector<PolygonCandidate> polygon_candidates_;
	std::vector<uint32_t> polygon_draw_order_;
	std::vector<uint32_t> polygon_clip_set_leafs_; // leaf of every clip set
	std::vector<Vec3f> polygon_clip_planes_;
	std::vector<uint8_t> polygon_is_sky_candidate_; // per candidate: the polygon reached the sky branch of prepare_polygon_surface
	std::vector<Vec3f> vertices_transformed_;
	std::vector<uint32_t> current_frame_visible_polygons_;
	uint64_t num_visible_surfaces_pixels_ = 0;

	std::vector<sqw_surface_job> jobs_;
	std::vector<sqw_draw_cmd> cmds_;
	std::vector<sqw_polygon_point> polygon_points_;
	RendererDebugStats stats_;
};

} // namespace sqwh
