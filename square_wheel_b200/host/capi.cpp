// C API of the host-side library (libsqw_host.so) for the head-less drivers (tests, bench, Python mirror).
// This library contains NO pixel work: it plays the part of the reference's Rust host around the C-ABI seams
// (scene description, visibility, surface preparation, BSP traversal).
#include "renderer.hpp"
#include "map_compiler.hpp"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

using namespace sqwh;

extern "C"
{

struct sqwh_scene_params // mirrors sqwh::SceneParams (plain C layout for ctypes)
{
	uint32_t seed;
	uint32_t grid[3];
	float cell_size[3];
	uint32_t num_rooms;
	uint32_t num_materials;
	uint32_t texture_size_log2;
	float max_polygon_size;
	uint32_t material_mode;
	float emissive_fraction;
	uint32_t num_models;
	uint32_t model_detail;
	uint32_t skybox;
	uint32_t sky_size_log2;
	uint32_t light_grid;
	uint32_t num_submodels;
};

struct sqwh_scene_info
{
	uint32_t num_nodes, num_leafs, num_polygons, num_portals, num_vertices, num_materials, num_images, num_waypoints;
	uint64_t num_lightmap_elements;
	uint32_t num_models, num_submodels;
};

struct sqwh_renderer_config
{
	uint32_t use_directional_lightmaps;
	float textures_mip_bias;
	uint32_t clear_background;
	uint32_t invert_polygons_order;
	float shadows_quality;
	uint32_t portals_depth;
};

struct sqwh_frame_info
{
	uint32_t num_jobs, num_cmds, num_polygon_points, needs_background_clear;
	uint64_t total_surface_pixels;
	uint32_t num_visible_leafs, num_visible_polygons;
	uint32_t mode_counts[3];
	float cam_pos[3];
	uint32_t num_lights;
	uint32_t num_depth_maps, num_depth_cmds, num_depth_points;
	uint32_t num_triangle_points;
	uint64_t shadow_maps_len;
};

struct sqwh_model_entity // frame_info.rs:55-69 (ModelEntity) in plain C layout
{
	float position[3];
	float rotation[4];    // quaternion s, x, y, z
	uint32_t frames[2];   // AnimationPoint
	float lerp;
	uint32_t model;       // index of the scene's model fixture
	uint32_t skin;        // index of the scene's skin (TextureLiteWithMips)
	uint32_t blending_mode;
	uint32_t lighting;    // 0 Default, 1 ConstantLight, 2 AdvancedLight
	float light[3];       // ConstantLight value / AdvancedLight::light_add
	float grid_light_scale;
	float light_position[3];
	uint32_t flags;       // 1 VIEW_MODEL, 2 ONLY_THIRD_PERSON_VIEW
};

struct sqwh_sprite // frame_info.rs:125-138 (Sprite) in plain C layout
{
	float position[3];
	float angle;
	float radius;
	uint32_t skin;          // index of the scene's skin (TextureLiteWithMips)
	uint32_t blending_mode;
	uint32_t orientation;   // SpriteOrientation 0..3
	float light_scale;
	float light_add[3];
};

struct sqwh_decal // frame_info.rs:112-123 (Decal) in plain C layout
{
	float position[3];
	float rotation[4]; // quaternion s, x, y, z; the decal is projected along its -X axis
	float scale[3];
	uint32_t skin;     // index of the scene's skin (TextureLiteWithMips)
	uint32_t blending_mode;
	float lightmap_light_scale;
	float light_add[3];
};

struct sqwh_model_info
{
	uint32_t num_frames, num_meshes, num_vertices, num_triangles, kind, num_bones;
	float bbox_min[3], bbox_max[3]; // union over all frames
};

struct sqwh_dynamic_light // frame_info.rs:149-168
{
	float position[3];
	float radius;
	float color[3];
	uint32_t shadow_type; // SQW_SHADOW_*
	float rotation[4];    // quaternion s, x, y, z (Projector)
	float fov;            // radians (Projector)
};

void sqwh_default_scene_params(sqwh_scene_params* p)
{
	const SceneParams d;
	p->seed = d.seed;
	for (int i = 0; i < 3; ++i)
	{
		p->grid[i] = d.grid[i];
		p->cell_size[i] = d.cell_size[i];
	}
	p->num_rooms = d.num_rooms;
	p->num_materials = d.num_materials;
	p->texture_size_log2 = d.texture_size_log2;
	p->max_polygon_size = d.max_polygon_size;
	p->material_mode = d.material_mode;
	p->emissive_fraction = d.emissive_fraction;
	p->num_models = d.num_models;
	p->model_detail = d.model_detail;
	p->skybox = d.skybox;
	p->sky_size_log2 = d.sky_size_log2;
	p->light_grid = d.light_grid;
	p->num_submodels = d.num_submodels;
}

Scene* sqwh_scene_create(const sqwh_scene_params* p)
{
	if (!p || p->grid[0] == 0 || p->grid[1] == 0 || p->grid[2] == 0 || p->num_materials == 0 || p->texture_size_log2 < 4 || p->texture_size_log2 > 11)
		return nullptr;
	SceneParams sp;
	sp.seed = p->seed;
	for (int i = 0; i < 3; ++i)
	{
		sp.grid[i] = p->grid[i];
		sp.cell_size[i] = p->cell_size[i];
	}
	sp.num_rooms = p->num_rooms;
	sp.num_materials = p->num_materials;
	sp.texture_size_log2 = p->texture_size_log2;
	sp.max_polygon_size = p->max_polygon_size;
	sp.material_mode = p->material_mode;
	sp.emissive_fraction = p->emissive_fraction;
	sp.num_models = p->num_models;
	sp.model_detail = p->model_detail;
	sp.skybox = p->skybox;
	sp.sky_size_log2 = p->sky_size_log2;
	sp.light_grid = p->light_grid;
	sp.num_submodels = p->num_submodels;
	Scene* s = new (std::nothrow) Scene();
	if (!s)
		return nullptr;
	generate_scene(sp, *s);
	return s;
}
void sqwh_scene_destroy(Scene* s) { delete s; }

void sqwh_scene_get_info(const Scene* s, sqwh_scene_info* out)
{
	out->num_nodes = (uint32_t)s->map.nodes.size();
	out->num_leafs = (uint32_t)s->map.leafs.size();
	out->num_polygons = (uint32_t)s->map.polygons.size();
	out->num_portals = (uint32_t)s->map.portals.size();
	out->num_vertices = (uint32_t)s->map.vertices.size();
	out->num_materials = (uint32_t)s->materials.size();
	out->num_images = (uint32_t)s->images.size();
	out->num_waypoints = (uint32_t)s->camera_waypoints.size();
	out->num_lightmap_elements = s->map.directional_lightmaps_data.size();
	out->num_models = (uint32_t)s->models.size();
	out->num_submodels = (uint32_t)s->map.submodels.size();
}
// .sqwm v11 (bsp_map_save_load.rs): write the scene's map / replace the scene's map by a file's.  Return a MapFileResult
// (0 ok, 1 i/o error, 2 not a map, 3 incompatible version), or 4 when the loaded map fails validate_map against the scene's
// material table (the message is then in sqwh_last_map_error).  No Renderer may exist for the scene while its map is replaced.
static std::string g_last_map_error;
int sqwh_scene_save_map(const Scene* s, const char* path) { return (int)save_map(s->map, path); }
int sqwh_scene_load_map(Scene* s, const char* path)
{
	BSPMap map;
	const MapFileResult r = load_map(path, map);
	if (r != MAP_FILE_OK)
		return (int)r;
	g_last_map_error = validate_map(map, s->materials.size());
	if (!g_last_map_error.empty())
		return 4;
	s->map = std::move(map);
	return 0;
}
const char* sqwh_last_map_error() { return g_last_map_error.c_str(); }
uint32_t sqwh_scene_num_light_grid_samples(const Scene* s) { return (uint32_t)s->map.light_grid_samples.size(); }
const char* sqwh_scene_texture_name(const Scene* s, uint32_t index) { return index < s->map.textures.size() ? s->map.textures[index].name : nullptr; }

const void* sqwh_scene_image(const Scene* s, uint32_t index, int32_t* kind, uint32_t* width, uint32_t* height)
{
	if (index >= s->images.size())
		return nullptr;
	const HostImage& img = s->images[index];
	*kind = img.kind;
	*width = img.width;
	*height = img.height;
	return img.pixels.data();
}
const float* sqwh_scene_lightmaps(const Scene* s) { return s->map.lightmaps_data.data(); }
const sqw_directional_lightmap_element* sqwh_scene_directional_lightmaps(const Scene* s) { return s->map.directional_lightmaps_data.data(); }
const Polygon* sqwh_scene_polygons(const Scene* s) { return s->map.polygons.data(); }
const Vec3f* sqwh_scene_vertices(const Scene* s) { return s->map.vertices.data(); }

// out: num_frames x {position[3], azimuth, elevation}
uint32_t sqwh_scene_camera_path(const Scene* s, uint32_t num_frames, uint32_t jitter_seed, float* out)
{
	std::vector<CameraPose> path;
	make_camera_path(*s, num_frames, jitter_seed, path);
	for (size_t i = 0; i < path.size(); ++i)
	{
		out[i * 5 + 0] = path[i].position.x;
		out[i * 5 + 1] = path[i].position.y;
		out[i * 5 + 2] = path[i].position.z;
		out[i * 5 + 3] = path[i].azimuth;
		out[i * 5 + 4] = path[i].elevation;
	}
	return (uint32_t)path.size();
}

Renderer* sqwh_renderer_create(const Scene* s, const sqwh_renderer_config* cfg)
{
	RendererConfig c;
	if (cfg)
	{
		c.use_directional_lightmaps = cfg->use_directional_lightmaps != 0;
		c.textures_mip_bias = cfg->textures_mip_bias;
		c.clear_background = cfg->clear_background != 0;
		c.invert_polygons_order = cfg->invert_polygons_order != 0;
		c.shadows_quality = cfg->shadows_quality;
		c.portals_depth = std::min(cfg->portals_depth, 8u); // renderer.rs:33
	}
	return new (std::nothrow) Renderer(*s, c);
}
void sqwh_renderer_destroy(Renderer* r) { delete r; }

// pose: {position[3], azimuth, elevation}; fov: vertical field of view in radians (game_map.rs:865 uses 0.375 * pi)
void sqwh_renderer_prepare_frame(Renderer* r, uint32_t width, uint32_t height, const float* pose, float fov)
{
	CameraPose p;
	p.position = {pose[0], pose[1], pose[2]};
	p.azimuth = pose[3];
	p.elevation = pose[4];
	r->prepare_frame(width, height, p, fov);
}
void sqwh_renderer_draw_frame(Renderer* r, int color_bits) { r->draw_frame(color_bits == 32 ? 32 : 64); }

// FrameInfo::lights of the next prepare_frame
void sqwh_renderer_set_lights(Renderer* r, const sqwh_dynamic_light* lights, uint32_t n)
{
	std::vector<DynamicLight> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		v[i].position = {lights[i].position[0], lights[i].position[1], lights[i].position[2]};
		v[i].radius = lights[i].radius;
		std::memcpy(v[i].color, lights[i].color, 12);
		v[i].shadow_type = lights[i].shadow_type;
		v[i].rotation = Quaternion{lights[i].rotation[0], {lights[i].rotation[1], lights[i].rotation[2], lights[i].rotation[3]}};
		v[i].fov = lights[i].fov;
	}
	r->set_lights(v.data(), v.size());
}

// FrameInfo::model_entities of the next prepare_frame.  Returns 0, or -1 when an entity refers to a missing model / skin / frame.
int sqwh_renderer_set_models(Renderer* r, const Scene* s, const sqwh_model_entity* models, uint32_t n)
{
	std::vector<ModelEntity> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		const sqwh_model_entity& m = models[i];
		if (m.model >= s->models.size() || m.skin >= s->model_skins.size() || m.blending_mode > SQW_BLEND_ALPHA_BLEND || m.lighting > MODEL_LIGHTING_ADVANCED)
			return -1;
		const size_t num_frames = s->models[m.model].frames_info.size();
		if (m.frames[0] >= num_frames || m.frames[1] >= num_frames)
			return -1;
		v[i].position = {m.position[0], m.position[1], m.position[2]};
		v[i].rotation = Quaternion{m.rotation[0], {m.rotation[1], m.rotation[2], m.rotation[3]}};
		v[i].frames[0] = m.frames[0];
		v[i].frames[1] = m.frames[1];
		v[i].lerp = m.lerp;
		v[i].model = m.model;
		for (uint32_t k = 0; k < NUM_MIPS; ++k)
			v[i].texture_images[k] = s->model_skins[m.skin][k];
		v[i].blending_mode = (int)m.blending_mode;
		v[i].lighting = (ModelLightingKind)m.lighting;
		std::memcpy(v[i].light, m.light, 12);
		v[i].grid_light_scale = m.grid_light_scale;
		v[i].light_position = {m.light_position[0], m.light_position[1], m.light_position[2]};
		v[i].flags = m.flags;
	}
	r->set_models(v.data(), v.size());
	return 0;
}

// FrameInfo::sprites of the next prepare_frame.  Returns 0, or -1 for a bad skin / mode.
int sqwh_renderer_set_sprites(Renderer* r, const Scene* s, const sqwh_sprite* sprites, uint32_t n)
{
	std::vector<Sprite> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		const sqwh_sprite& sp = sprites[i];
		if (sp.skin >= s->model_skins.size() || sp.blending_mode > SQW_BLEND_ALPHA_BLEND || sp.orientation > SPRITE_ALIGN_TO_Z_AXIS_FACING_TOWARDS_CAMERA || !(sp.radius > 0.0f))
			return -1;
		v[i].position = {sp.position[0], sp.position[1], sp.position[2]};
		v[i].angle = sp.angle;
		v[i].radius = sp.radius;
		for (uint32_t k = 0; k < NUM_MIPS; ++k)
			v[i].texture_images[k] = s->model_skins[sp.skin][k];
		v[i].blending_mode = (int)sp.blending_mode;
		v[i].orientation = (SpriteOrientation)sp.orientation;
		v[i].light_scale = sp.light_scale;
		std::memcpy(v[i].light_add, sp.light_add, 12);
	}
	r->set_sprites(v.data(), v.size());
	return 0;
}

#define SQWH_MAX_PORTAL_VERTICES 8
struct sqwh_portal // frame_info.rs:171-205 (ViewPortal + ViewPortalTexture) in plain C layout
{
	uint32_t view;               // 0 CameraAtPosition, 1 Mirror, 2 ParallaxPortal
	float camera_position[3];    // CameraAtPosition
	float camera_rotation[4];    // quaternion s, x, y, z
	float camera_fov;            // radians
	float transform_matrix[16];  // ParallaxPortal, column-major
	float plane[4];              // vec, dist
	float tex_coord_equation[2][4];
	uint32_t num_vertices;
	float vertices[SQWH_MAX_PORTAL_VERTICES][3];
	uint32_t blending_mode;
	uint32_t has_texture;        // ViewPortalTexture present
	uint32_t texture_skin;       // index of the scene's skin (TextureLiteWithMips)
	uint32_t texture_blending_mode;
	float texture_light_scale;
	float texture_light_add[3];
};
struct sqwh_portal_info // renderer_structs.rs:234-243 + what the caller needs to finish the portal texture
{
	uint32_t resolution[2];      // 0 x 0: not visible this frame
	int32_t tc_min[2];
	uint32_t mip;
	uint32_t has_texture, texture_image, texture_blending_mode; // overlay: SQW_IMAGE_COLOR32 image id of mip `mip`
	float texture_light[3];
};

int sqwh_renderer_set_portals(Renderer* r, const Scene* s, const sqwh_portal* portals, uint32_t n)
{
	std::vector<ViewPortal> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		const sqwh_portal& p = portals[i];
		if (p.view > PORTAL_VIEW_PARALLAX_PORTAL || p.num_vertices < 3 || p.num_vertices > SQWH_MAX_PORTAL_VERTICES || p.blending_mode > SQW_BLEND_ALPHA_BLEND ||
			(p.has_texture && (p.texture_skin >= s->model_skins.size() || p.texture_blending_mode > SQW_BLEND_ALPHA_BLEND)))
			return -1;
		v[i].view = (PortalViewKind)p.view;
		v[i].camera_position = {p.camera_position[0], p.camera_position[1], p.camera_position[2]};
		v[i].camera_rotation = Quaternion{p.camera_rotation[0], {p.camera_rotation[1], p.camera_rotation[2], p.camera_rotation[3]}};
		v[i].camera_fov = p.camera_fov;
		const float* m = p.transform_matrix;
		v[i].transform_matrix = Mat4f::from_cols({m[0], m[1], m[2], m[3]}, {m[4], m[5], m[6], m[7]}, {m[8], m[9], m[10], m[11]}, {m[12], m[13], m[14], m[15]});
		v[i].plane = Plane{{p.plane[0], p.plane[1], p.plane[2]}, p.plane[3]};
		for (int k = 0; k < 2; ++k)
			v[i].tex_coord_equation[k] = Plane{{p.tex_coord_equation[k][0], p.tex_coord_equation[k][1], p.tex_coord_equation[k][2]}, p.tex_coord_equation[k][3]};
		for (uint32_t k = 0; k < p.num_vertices; ++k)
			v[i].vertices.push_back({p.vertices[k][0], p.vertices[k][1], p.vertices[k][2]});
		v[i].blending_mode = (int)p.blending_mode;
		v[i].has_texture = p.has_texture != 0;
		if (v[i].has_texture)
		{
			for (uint32_t k = 0; k < NUM_MIPS; ++k)
				v[i].texture_images[k] = s->model_skins[p.texture_skin][k];
			v[i].texture_blending_mode = (int)p.texture_blending_mode;
			v[i].texture_light_scale = p.texture_light_scale;
			std::memcpy(v[i].texture_light_add, p.texture_light_add, 12);
		}
	}
	r->set_portals(v.data(), v.size());
	return 0;
}
uint32_t sqwh_renderer_num_portals(const Renderer* r) { return (uint32_t)r->portals_info().size(); }
int sqwh_renderer_portal_info(const Renderer* r, uint32_t index, sqwh_portal_info* out)
{
	if (index >= r->portals_info().size())
		return -1;
	const PortalInfo& pi = r->portals_info()[index];
	const ViewPortal& p = r->portals()[index];
	std::memset(out, 0, sizeof(*out));
	out->resolution[0] = pi.resolution[0];
	out->resolution[1] = pi.resolution[1];
	out->tc_min[0] = pi.tc_min[0];
	out->tc_min[1] = pi.tc_min[1];
	out->mip = pi.mip;
	out->has_texture = p.has_texture ? 1 : 0;
	if (p.has_texture)
	{
		out->texture_image = p.texture_images[pi.mip];
		out->texture_blending_mode = (uint32_t)p.texture_blending_mode;
		r->portal_texture_light(index, out->texture_light);
	}
	return 0;
}
// The nested view of portal `index` (build_view_portals_textures): prepares and records the child renderer's frame and returns the
// child (read its lists with the usual getters before the next call), or NULL: portal invisible / depth limit (black texture).
Renderer* sqwh_renderer_prepare_portal_view(Renderer* r, uint32_t index, int color_bits) { return r->prepare_portal_view(index, color_bits == 32 ? 32 : 64); }

struct sqwh_submodel_entity // frame_info.rs:13 (SubmodelEntityOpt): present = 0 hides the submodel
{
	uint32_t present;
	float position[3];
	float rotation[4]; // quaternion s, x, y, z
};
// FrameInfo::submodel_entities of the next prepare_frame: one record per map submodel (n must equal the number of submodels).
int sqwh_renderer_set_submodel_entities(Renderer* r, const sqwh_submodel_entity* e, uint32_t n)
{
	if (n != r->num_submodels())
		return -1;
	std::vector<Renderer::SubmodelEntity> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		v[i].present = e[i].present != 0;
		v[i].position = {e[i].position[0], e[i].position[1], e[i].position[2]};
		v[i].rotation = Quaternion{e[i].rotation[0], {e[i].rotation[1], e[i].rotation[2], e[i].rotation[3]}};
	}
	r->set_submodel_entities(v.data(), n);
	return 0;
}
// bbox centre of a submodel = the position that leaves it where the map has it
void sqwh_renderer_submodel_center(const Renderer* r, uint32_t index, float out[3])
{
	const Vec3f c = r->submodel_bbox(index).get_center();
	out[0] = c.x;
	out[1] = c.y;
	out[2] = c.z;
}

// FrameInfo::decals of the next prepare_frame.  Returns 0, or -1 for a bad skin / mode.
int sqwh_renderer_set_decals(Renderer* r, const Scene* s, const sqwh_decal* decals, uint32_t n)
{
	std::vector<Decal> v(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		const sqwh_decal& d = decals[i];
		if (d.skin >= s->model_skins.size() || d.blending_mode > SQW_BLEND_ALPHA_BLEND)
			return -1;
		v[i].position = {d.position[0], d.position[1], d.position[2]};
		v[i].rotation = Quaternion{d.rotation[0], {d.rotation[1], d.rotation[2], d.rotation[3]}};
		v[i].scale = {d.scale[0], d.scale[1], d.scale[2]};
		for (uint32_t k = 0; k < NUM_MIPS; ++k)
			v[i].texture_images[k] = s->model_skins[d.skin][k];
		v[i].blending_mode = (int)d.blending_mode;
		v[i].lightmap_light_scale = d.lightmap_light_scale;
		std::memcpy(v[i].light_add, d.light_add, 12);
	}
	r->set_decals(v.data(), v.size());
	return 0;
}

// build_dynamic_models_buffers of the frame prepared last: records (Renderer::DynamicMeshExport, 304 bytes), vertices (8 f32 each)
// and sorted triangles (3 u16 each) of this mirror.  NULL buffers: only the counts are returned.
int sqwh_renderer_dynamic_meshes(const Renderer* r, void* records, uint32_t capacity, float* vertices, uint64_t vertices_capacity, uint16_t* triangles, uint64_t triangles_capacity, uint64_t counts_out[3])
{
	if (!r)
		return -1;
	static_assert(sizeof(Renderer::DynamicMeshExport) == 304, "DynamicMeshExport layout");
	const size_t n = r->export_dynamic_meshes(nullptr, 0);
	(void)n;
	std::vector<Renderer::DynamicMeshExport> all(4096);
	const size_t count = r->export_dynamic_meshes(all.data(), all.size());
	const auto& v = r->dynamic_meshes_vertices();
	const auto& t = r->dynamic_meshes_triangles();
	if (counts_out)
	{
		counts_out[0] = count;
		counts_out[1] = v.size();
		counts_out[2] = t.size();
	}
	if (records)
		std::memcpy(records, all.data(), std::min<size_t>(count, capacity) * sizeof(Renderer::DynamicMeshExport));
	if (vertices)
		std::memcpy(vertices, v.data(), std::min<uint64_t>(v.size(), vertices_capacity) * sizeof(ModelVertex3d));
	if (triangles)
		std::memcpy(triangles, t.data(), std::min<uint64_t>(t.size(), triangles_capacity) * sizeof(Triangle));
	return (int)count;
}

// prepare_polygons_surfaces of the frame prepared last, leaf polygons: the (polygon, clip set) candidates in preparation order, the
// clip planes (8 x 3 floats per set), this mirror's results in the layout of sqw_prepared_polygon (64 bytes) and the frame record
// (Renderer::PolygonsFrameExport).  NULL buffers are skipped; counts_out = {candidates, clip sets}.
int sqwh_renderer_prepared_polygons(const Renderer* r, void* candidates, float* clip_planes, void* prepared, void* frame, uint64_t counts_out[2])
{
	if (!r)
		return -1;
	static_assert(sizeof(Renderer::PreparedPolygonExport) == 64 && sizeof(Renderer::PolygonCandidate) == 8, "export layouts");
	const auto& c = r->polygon_candidates();
	const auto& p = r->polygon_clip_planes();
	if (counts_out)
	{
		counts_out[0] = c.size();
		counts_out[1] = p.size() / 8;
	}
	if (candidates && !c.empty())
		std::memcpy(candidates, c.data(), c.size() * sizeof(Renderer::PolygonCandidate));
	if (clip_planes && !p.empty())
		std::memcpy(clip_planes, p.data(), p.size() * sizeof(Vec3f));
	r->export_prepared_polygons(static_cast<Renderer::PreparedPolygonExport*>(prepared), static_cast<Renderer::PolygonsFrameExport*>(frame));
	return (int)c.size();
}
double sqwh_renderer_polygons_prepare_seconds(const Renderer* r) { return r ? r->polygons_prepare_seconds() : 0.0; }
// map polygon indices of the leaf polygons drawn by the last draw_frame, in draw order (the BSP walk); returns the count
uint32_t sqwh_renderer_polygon_draw_order(const Renderer* r, uint32_t* out, uint32_t capacity)
{
	const auto& o = r->polygon_draw_order();
	if (out)
		std::memcpy(out, o.data(), std::min<size_t>(o.size(), capacity) * 4);
	return (uint32_t)o.size();
}
// sqw_map_material records (64 bytes each) for sqw_upload_map_materials
uint32_t sqwh_scene_map_materials(const Scene* s, void* out, uint32_t capacity)
{
	struct Rec
	{
		uint32_t flags, texture_images[5], emissive_images[5];
		float emissive_light[3];
		int32_t texture_shift[2];
	};
	static_assert(sizeof(Rec) == 64, "sqw_map_material layout");
	const uint32_t n = (uint32_t)s->materials.size();
	Rec* o = static_cast<Rec*>(out);
	for (uint32_t i = 0; i < n && i < capacity; ++i)
	{
		const Material& m = s->materials[i];
		Rec r{};
		r.flags = (m.light ? 1u : 0u) | (m.has_diffuse ? 2u : 0u) | (m.has_normal_map ? 4u : 0u) | (m.has_non_one_roughness ? 8u : 0u) | (m.is_metal ? 16u : 0u) | (m.has_emissive ? 32u : 0u);
		for (int k = 0; k < 5; ++k)
		{
			r.texture_images[k] = m.texture_images[k];
			r.emissive_images[k] = m.emissive_images[k];
		}
		std::memcpy(r.emissive_light, m.emissive_light, 12);
		r.texture_shift[0] = m.texture_shift[0];
		r.texture_shift[1] = m.texture_shift[1];
		o[i] = r;
	}
	return n;
}
// PolygonBasisVecs of every map polygon for mip 0: [num_polygons][12] floats {u, v, start, normal}
uint32_t sqwh_renderer_polygon_basis(const Renderer* r, float* out, uint32_t capacity) { return r->export_polygon_basis(out, capacity); }
// the lights located in the leaf of every clip set of the last prepare_frame: offsets [sets + 1] (u32) and light indices (u16); NULL buffers: counts only
uint32_t sqwh_renderer_clip_set_lights(const Renderer* r, uint32_t* offsets, uint16_t* list, uint32_t counts_out[2])
{
	std::vector<uint32_t> off;
	std::vector<uint16_t> li;
	r->export_clip_set_lights(off, li);
	if (counts_out)
	{
		counts_out[0] = (uint32_t)off.size();
		counts_out[1] = (uint32_t)li.size();
	}
	if (offsets)
		std::memcpy(offsets, off.data(), off.size() * 4);
	if (list && !li.empty())
		std::memcpy(list, li.data(), li.size() * 2);
	return (uint32_t)off.size();
}
// per-material SQW_MATERIAL_* flags for sqw_upload_map_geometry
uint32_t sqwh_scene_material_flags(const Scene* s, uint8_t* out, uint32_t capacity)
{
	const uint32_t n = (uint32_t)s->materials.size();
	for (uint32_t i = 0; i < n && i < capacity; ++i)
		out[i] = (uint8_t)((s->materials[i].draw ? 1 : 0) | (s->materials[i].is_skybox ? 2 : 0) | ((s->materials[i].blending_mode & 7) << 4));
	return n;
}

int sqwh_scene_model_info(const Scene* s, uint32_t index, sqwh_model_info* out)
{
	if (index >= s->models.size())
		return -1;
	const TriangleModel& m = s->models[index];
	std::memset(out, 0, sizeof(*out));
	out->num_frames = (uint32_t)m.frames_info.size();
	out->num_meshes = (uint32_t)m.meshes.size();
	out->num_bones = (uint32_t)m.bones.size();
	for (const TriangleModelMesh& mesh : m.meshes)
	{
		out->num_vertices += (uint32_t)mesh.num_vertices();
		out->num_triangles += (uint32_t)mesh.triangles.size();
		out->kind = mesh.kind;
	}
	BBox b = m.frames_info[0];
	for (const BBox& f : m.frames_info)
		b.extend(f);
	const float mn[3] = {b.min.x, b.min.y, b.min.z}, mx[3] = {b.max.x, b.max.y, b.max.z};
	std::memcpy(out->bbox_min, mn, 12);
	std::memcpy(out->bbox_max, mx, 12);
	return 0;
}

void sqwh_renderer_frame_info(const Renderer* r, sqwh_frame_info* out)
{
	std::memset(out, 0, sizeof(*out));
	out->num_jobs = (uint32_t)r->jobs().size();
	out->num_cmds = (uint32_t)r->cmds().size();
	out->num_polygon_points = (uint32_t)r->polygon_points().size();
	out->needs_background_clear = r->needs_background_clear() ? 1 : 0;
	out->total_surface_pixels = r->total_surface_pixels();
	out->num_visible_leafs = r->stats().num_visible_leafs;
	out->num_visible_polygons = r->stats().num_visible_polygons;
	for (int i = 0; i < 3; ++i)
		out->mode_counts[i] = r->stats().mode_counts[i];
	out->cam_pos[0] = r->camera_matrices().position.x;
	out->cam_pos[1] = r->camera_matrices().position.y;
	out->cam_pos[2] = r->camera_matrices().position.z;
	out->num_lights = (uint32_t)r->lights().size();
	out->num_depth_maps = (uint32_t)r->depth_maps().size();
	out->num_depth_cmds = (uint32_t)r->depth_cmds().size();
	out->num_depth_points = (uint32_t)r->depth_points().size();
	out->num_triangle_points = (uint32_t)r->triangle_points().size();
	out->shadow_maps_len = r->shadow_maps_len();
}
const sqw_light* sqwh_renderer_lights(const Renderer* r) { return r->lights().data(); }
const sqw_depth_map* sqwh_renderer_depth_maps(const Renderer* r) { return r->depth_maps().data(); }
const sqw_depth_cmd* sqwh_renderer_depth_cmds(const Renderer* r) { return r->depth_cmds().data(); }
const sqw_polygon_point* sqwh_renderer_depth_points(const Renderer* r) { return r->depth_points().data(); }
const sqw_surface_job* sqwh_renderer_jobs(const Renderer* r) { return r->jobs().data(); }
const sqw_draw_cmd* sqwh_renderer_cmds(const Renderer* r) { return r->cmds().data(); }
const sqw_polygon_point* sqwh_renderer_polygon_points(const Renderer* r) { return r->polygon_points().data(); }
const sqw_triangle_point* sqwh_renderer_triangle_points(const Renderer* r) { return r->triangle_points().data(); }

// ---- triangle-model files (model_file.cpp: triangle_model_md3.rs / triangle_model_iqm.rs) ----------------------------------
// Loads an .md3 / .iqm file (recognised by its magic) and appends it to the scene's models; returns the model index the
// entities of sqwh_renderer_set_models refer to, or -1 with a message in err (when err_capacity > 0).
static int append_loaded_model(Scene* s, bool ok, TriangleModel& m, const std::string& message, char* err, size_t err_capacity)
{
	if (!ok || m.meshes.empty() || m.frames_info.empty())
	{
		if (err && err_capacity)
			std::snprintf(err, err_capacity, "%s", ok ? "model without meshes or frames" : message.c_str());
		return -1;
	}
	s->models.push_back(std::move(m));
	return (int)s->models.size() - 1;
}
int sqwh_scene_load_model_file(Scene* s, const char* path, char* err, size_t err_capacity)
{
	TriangleModel m;
	std::string message;
	const bool ok = load_model_file(path, m, &message);
	return append_loaded_model(s, ok, m, message, err, err_capacity);
}
int sqwh_scene_load_model_memory(Scene* s, const void* data, size_t size, char* err, size_t err_capacity)
{
	TriangleModel m;
	std::string message;
	const bool ok = load_model(static_cast<const uint8_t*>(data), size, m, &message);
	return append_loaded_model(s, ok, m, message, err, err_capacity);
}

// Raw arrays of a model for inspection (tests): returns the byte size of the array, copies it when it fits into `capacity`.
//   what: 0 triangles u16[n][3] | 1 non-animated vertices f32[n][8] (position, normal, tc) | 2 vertex-animated constant f32[n][2]
//         3 vertex-animated variable f32[frames*n][6] | 4 skeleton vertices {f32 tc[2], position[3], normal[3], u8 bone[4], weight[4]} (40 B)
//         5 frame bones f32[frames*bones][16] (column-major) | 6 frame bounding boxes f32[frames][6] | 7 bone parents u32[bones]
//         8 tc_shift f32[2] | 9 names: mesh name, material name, then bone names and animation names, each NUL-terminated
//         10 animations {u32 start_frame, num_frames, f32 fps, u32 looped}[n]
int64_t sqwh_scene_model_copy(const Scene* s, uint32_t index, uint32_t mesh_index, uint32_t what, void* out, uint64_t capacity)
{
	if (index >= s->models.size())
		return -1;
	const TriangleModel& m = s->models[index];
	if (mesh_index >= m.meshes.size())
		return -1;
	const TriangleModelMesh& mesh = m.meshes[mesh_index];
	std::vector<uint8_t> tmp;
	const void* src = nullptr;
	size_t bytes = 0;
	auto pack_floats = [&](const std::vector<float>& f) {
		tmp.resize(f.size() * 4);
		if (!f.empty())
			std::memcpy(tmp.data(), f.data(), tmp.size());
		src = tmp.data();
		bytes = tmp.size();
	};
	static_assert(sizeof(Triangle) == 6 && sizeof(VertexNonAnimated) == 32 && sizeof(VertexAnimatedVertexConstant) == 8 && sizeof(VertexAnimatedVertexVariable) == 24 &&
					  sizeof(SkeletonAnimatedVertex) == 40 && sizeof(Mat4f) == 64 && sizeof(BBox) == 24,
				  "model arrays are copied as they lie");
	switch (what)
	{
	case 0: src = mesh.triangles.data(), bytes = mesh.triangles.size() * sizeof(Triangle); break;
	case 1: src = mesh.non_animated.data(), bytes = mesh.non_animated.size() * sizeof(VertexNonAnimated); break;
	case 2: src = mesh.constant.data(), bytes = mesh.constant.size() * sizeof(VertexAnimatedVertexConstant); break;
	case 3: src = mesh.variable.data(), bytes = mesh.variable.size() * sizeof(VertexAnimatedVertexVariable); break;
	case 4: src = mesh.skeleton.data(), bytes = mesh.skeleton.size() * sizeof(SkeletonAnimatedVertex); break;
	case 5: src = m.frame_bones.data(), bytes = m.frame_bones.size() * sizeof(Mat4f); break;
	case 6: src = m.frames_info.data(), bytes = m.frames_info.size() * sizeof(BBox); break;
	case 7: src = m.bones.data(), bytes = m.bones.size() * 4; break;
	case 8: pack_floats({m.tc_shift.x, m.tc_shift.y}); break;
	case 9:
	{
		std::string all = mesh.name + '\0' + mesh.material_name + '\0';
		for (const std::string& n : m.bone_names)
			all += n + '\0';
		for (const TriangleModelAnimation& a : m.animations)
			all += a.name + '\0';
		tmp.assign(all.begin(), all.end());
		src = tmp.data();
		bytes = tmp.size();
		break;
	}
	case 10:
	{
		tmp.resize(m.animations.size() * 16);
		for (size_t i = 0; i < m.animations.size(); ++i)
		{
			const TriangleModelAnimation& a = m.animations[i];
			const uint32_t looped = a.looped ? 1u : 0u;
			std::memcpy(&tmp[i * 16], &a.start_frame, 4);
			std::memcpy(&tmp[i * 16 + 4], &a.num_frames, 4);
			std::memcpy(&tmp[i * 16 + 8], &a.frames_per_second, 4);
			std::memcpy(&tmp[i * 16 + 12], &looped, 4);
		}
		src = tmp.data();
		bytes = tmp.size();
		break;
	}
	default: return -1;
	}
	if (out && bytes <= capacity && bytes)
		std::memcpy(out, src, bytes);
	return (int64_t)bytes;
}

// ---- offline map tools (map_compiler.cpp, lightmapper.cpp; the reference's map_compiler / lightmapper binaries) -----------------
// materials_table: the text form of the reference's materials/*.json that mapc::parse_material_table reads.
// flags: bit 0 --perform-csg, bit 1 --perform-advanced-splitter-plane-selection.  stats_out: 11 x u32 (mapc::CompileStats).
// Returns 0, 1 (i/o error writing the map) or 5 (the map text does not parse / holds no entity); the tool's messages go to log.
int sqwh_map_compile(const char* map_text, const char* materials_table, uint32_t flags, const char* out_path, uint32_t* stats_out, char* log, uint32_t log_capacity)
{
	if (!out_path)
		return (int)MAP_FILE_IO_ERROR;
	mapc::CompileOptions opt;
	opt.perform_csg = (flags & 1u) != 0;
	opt.advanced_splitter_plane_selection = (flags & 2u) != 0;
	const mapc::MaterialTable table = mapc::parse_material_table(materials_table ? materials_table : "");
	BSPMap map;
	mapc::CompileStats st;
	std::string messages;
	const bool ok = mapc::compile_map(map_text ? map_text : "", table, opt, map, st, messages);
	if (log && log_capacity)
	{
		const size_t n = std::min<size_t>(messages.size(), log_capacity - 1);
		std::memcpy(log, messages.data(), n);
		log[n] = 0;
	}
	if (stats_out)
	{
		const uint32_t v[11] = {st.initial_polygons, st.polygons_after_csg, st.leafs_initial, st.leafs_removed, st.tree_depth, st.nodes, st.leafs, st.polygons, st.portals, st.submodels, st.lightmap_texels};
		std::memcpy(stats_out, v, sizeof(v));
	}
	if (!ok)
		return 5;
	return (int)save_map(map, out_path);
}

// lightmapper -i in_path -o out_path --sample-grid-size .. --light-scale .. --ambient-light .. [--no-directional-lightmap]
// (direct light only: lightmapper.cpp).  stats_out: 4 x u32 (mapc::LightmapStats).  Returns a MapFileResult.
int sqwh_map_light(const char* in_path, const char* materials_table, uint32_t sample_grid_size, float light_scale, float ambient_light, int32_t directional, float light_cutoff, const char* out_path, uint32_t* stats_out)
{
	if (!in_path || !out_path)
		return (int)MAP_FILE_IO_ERROR;
	BSPMap map;
	const MapFileResult r = load_map(in_path, map);
	if (r != MAP_FILE_OK)
		return (int)r;
	const mapc::MaterialTable table = mapc::parse_material_table(materials_table ? materials_table : "");
	mapc::LightmappingSettings settings;
	settings.sample_grid_size = sample_grid_size;
	settings.light_scale = light_scale;
	settings.ambient_light = ambient_light;
	settings.build_directional_lightmap = (directional & 1) != 0;
	settings.use_pvs = (directional & 2) != 0; // bit 1 of `directional`: --use-pvs
	if (light_cutoff > 0.0f)
		settings.light_cutoff = light_cutoff;
	mapc::LightmapStats st;
	mapc::build_lightmaps(settings, table, map, st);
	if (stats_out)
	{
		const uint32_t v[4] = {st.point_lights, st.sun_lights, st.lightmap_texels, st.light_grid_samples};
		std::memcpy(stats_out, v, sizeof(v));
	}
	return (int)save_map(map, out_path);
}

// Leaf-to-leaf visibility matrix of a map file (pvs.rs calculate_visibility_matrix): writes num_leafs^2 bytes when `capacity` allows
// and returns num_leafs, or a negative MapFileResult.
int64_t sqwh_map_pvs(const char* in_path, uint8_t* out, uint64_t capacity)
{
	if (!in_path)
		return -(int64_t)MAP_FILE_IO_ERROR;
	BSPMap map;
	const MapFileResult r = load_map(in_path, map);
	if (r != MAP_FILE_OK)
		return -(int64_t)r;
	const std::vector<uint8_t> m = mapc::calculate_leafs_visibility(map);
	if (out && m.size() <= capacity)
		std::memcpy(out, m.data(), m.size());
	return (int64_t)map.leafs.size();
}

// classname of the map entity that owns a submodel ("" when none): the game decides by it whether the inline model is drawn
// (test_game/world_spawn.rs: triggers are invisible, func_mirror / func_camera_portal / func_parallax_portal become view portals)
const char* sqwh_scene_submodel_classname(const Scene* s, uint32_t submodel_index)
{
	static thread_local std::string name;
	name.clear();
	const BSPMap& m = s->map;
	for (const Entity& e : m.entities)
		if (e.submodel_index == submodel_index)
			for (uint32_t i = 0; i < e.num_key_value_pairs; ++i)
			{
				const KeyValuePair& kv = m.key_value_pairs[e.first_key_value_pair + i];
				if (kv.key.size == 9 && std::memcmp(m.strings_data.data() + kv.key.offset, "classname", 9) == 0)
					name.assign((const char*)m.strings_data.data() + kv.value.offset, kv.value.size);
			}
	return name.c_str();
}

// Replace the scene by a compiled map with materials bound by texture name (bind_compiled_map).  Returns a MapFileResult, or 4
// when the map fails validate_map (message in sqwh_last_map_error).  No Renderer may exist for the scene meanwhile.
int sqwh_scene_load_compiled_map(Scene* s, const char* path, const char* materials_table, uint32_t texture_size_log2, uint32_t seed)
{
	if (!s || !path)
		return (int)MAP_FILE_IO_ERROR;
	BSPMap map;
	const MapFileResult r = load_map(path, map);
	if (r != MAP_FILE_OK)
		return (int)r;
	g_last_map_error = validate_map(map, map.textures.size());
	if (!g_last_map_error.empty())
		return 4;
	bind_compiled_map(*s, std::move(map), mapc::parse_material_table(materials_table ? materials_table : ""), std::min(std::max(texture_size_log2, 4u), 10u), seed);
	return 0;
}

} // extern "C"
