// In-memory map (BSPMap, src/common/bsp_map_compact.rs:7-35), materials and textures of the host-side mirror, plus the
// seeded synthetic map generator that stood in for map_compiler + lightmapper output before those tools were restated
// (map_compiler.hpp; SURVEY.md §0: the compiled demo map and its textures are not shipped); it remains the fixture of most tests
// and the fallback workload of bench.py.
#pragma once
#include "../../include/sqw_cuda.h"
#include "math.hpp"
#include "model_data.hpp"
#include <string>
#include <vector>

namespace sqwh
{

constexpr uint32_t FIRST_LEAF_INDEX = 1u << 31; // bsp_map_compact.rs:46
constexpr uint32_t MAX_BSP_TREE_DEPTH = 128;
constexpr uint32_t LIGHTMAP_SCALE_LOG2 = 4; // common/lightmap.rs:4
constexpr uint32_t LIGHTMAP_SCALE = 1u << LIGHTMAP_SCALE_LOG2;
constexpr uint32_t MAX_MIP = 4; // textures.rs:6
constexpr uint32_t NUM_MIPS = MAX_MIP + 1;

struct BSPNode // bsp_map_compact.rs:37-44 (24 bytes)
{
	uint32_t children[2];
	Plane plane;
};
struct BSPLeaf // :48-56
{
	uint32_t first_polygon, num_polygons, first_leaf_portal, num_leaf_portals;
};
struct Polygon // :58-75 (80 bytes)
{
	uint32_t first_vertex, num_vertices;
	Plane plane;
	Plane tex_coord_equation[2];
	int32_t tex_coord_min[2];
	int32_t tex_coord_max[2];
	uint32_t lightmap_data_offset;
	uint32_t texture;
};
static_assert(sizeof(Polygon) == 80, "Polygon layout (.sqwm lump 2)");
struct Portal // :77-85
{
	uint32_t leafs[2];
	Plane plane;
	uint32_t first_vertex, num_vertices;
};

struct Submodel // bsp_map_compact.rs:87-96
{
	uint32_t root_node, first_polygon, num_polygons;
};
struct SubmodelBSPNode // :98-107
{
	Plane plane;
	uint32_t first_polygon, num_polygons;
	uint32_t children[2];
};
struct Entity // :109-117
{
	uint32_t first_key_value_pair, num_key_value_pairs, submodel_index;
};
struct StringRef // :127-135
{
	uint16_t offset, size;
};
struct KeyValuePair // :119-125
{
	StringRef key, value;
};
constexpr size_t MAX_TEXTURE_NAME_LEN = 64; // :137
struct TextureName
{
	char name[MAX_TEXTURE_NAME_LEN]; // NUL-padded material name
};
struct LightGridHeader // :158-167
{
	float grid_cell_size[3] = {0, 0, 0};
	float grid_start[3] = {0, 0, 0};
	uint32_t grid_size[3] = {0, 0, 0};
};
struct LightGridColumn // :169-177
{
	uint32_t start_z = 0, first_sample = 0, num_samples = 0;
};
static_assert(sizeof(BSPNode) == 24 && sizeof(BSPLeaf) == 16 && sizeof(Portal) == 32 && sizeof(Submodel) == 12 && sizeof(SubmodelBSPNode) == 32, ".sqwm lump element sizes");
static_assert(sizeof(Entity) == 12 && sizeof(KeyValuePair) == 8 && sizeof(TextureName) == 64 && sizeof(LightGridHeader) == 36 && sizeof(LightGridColumn) == 12, ".sqwm lump element sizes");

struct BSPMap
{
	std::vector<BSPNode> nodes; // last node is the root
	std::vector<BSPLeaf> leafs;
	std::vector<Polygon> polygons;
	std::vector<Portal> portals;
	std::vector<uint32_t> leafs_portals;
	std::vector<Vec3f> vertices;
	std::vector<TextureName> textures;               // material names, indexed by Polygon::texture
	std::vector<Submodel> submodels;                 // inline models (kept for the file format; the mirror does not draw them)
	std::vector<SubmodelBSPNode> submodels_bsp_nodes;
	std::vector<Entity> entities;
	std::vector<KeyValuePair> key_value_pairs;
	std::vector<uint8_t> strings_data;
	LightGridHeader light_grid_header;
	std::vector<LightGridColumn> light_grid_columns; // grid_size[0] x grid_size[1]
	std::vector<LightGridElement> light_grid_samples;
	std::vector<float> lightmaps_data; // [n][3]
	std::vector<sqw_directional_lightmap_element> directional_lightmaps_data;
	uint32_t root_node() const { return nodes.empty() ? FIRST_LEAF_INDEX : (uint32_t)nodes.size() - 1; } // :289-292
};

// .sqwm v11 (src/common/bsp_map_save_load.rs): "SqwM", version, 32 lumps {offset, element_size,
// element_count}, raw little-endian #[repr(C)] elements; lightmaps, directional lightmaps and light grid samples are stored
// quantised (lightmap_compression.rs) and are de-quantised f32 in memory.
enum MapFileResult
{
	MAP_FILE_OK = 0,
	MAP_FILE_IO_ERROR = 1,     // cannot open / short read or write
	MAP_FILE_NOT_A_MAP = 2,    // bad id ("File is not a valid BSP map")
	MAP_FILE_BAD_VERSION = 3,  // "Can't load incompatible map version"
};
MapFileResult save_map(const BSPMap& map, const char* file_path);
MapFileResult load_map(const char* file_path, BSPMap& out);
// The structural checks the renderer relies on (indices inside their arrays, lightmap ranges, root node): "" or what is wrong.
std::string validate_map(const BSPMap& map, size_t num_materials);

inline uint32_t get_lightmap_size(int32_t tc_min, int32_t tc_max) // common/lightmap.rs:19-27
{
	return (uint32_t)(((tc_max + (int32_t)(LIGHTMAP_SCALE - 1)) >> LIGHTMAP_SCALE_LOG2) - (tc_min >> LIGHTMAP_SCALE_LOG2) + 1);
}

// Host image (one mip).  kind = SQW_IMAGE_*; pixels are TextureElement (8 B) or Color32 (4 B) or Color64 (8 B).
struct HostImage
{
	int kind = 0;
	uint32_t width = 0, height = 0;
	std::vector<uint8_t> pixels;
};

// The subset of common/material.rs:7-98 the world-drawing path reads.
struct Material
{
	bool draw = true;
	bool light = true;
	bool decals = true; // receives decals (material.rs `decals`, partial_renderer.rs:2814)
	bool shadow = true; // casts shadows in shadow maps (material.rs `shadow`, depth_renderer.rs:31)
	bool has_diffuse = true;
	int blending_mode = SQW_BLEND_NONE;
	bool has_normal_map = false, has_non_one_roughness = false, is_metal = false; // Texture flags (textures.rs:14-16)
	uint32_t texture_images[NUM_MIPS] = {0, 0, 0, 0, 0};                          // image ids of the TextureElement mips
	bool has_emissive = false;
	uint32_t emissive_images[NUM_MIPS] = {0, 0, 0, 0, 0};
	float emissive_light[3] = {0, 0, 0};
	int32_t texture_shift[2] = {0, 0}; // MapMaterialsProcessor::get_texture_shift (scrolling), 0 when animations are off
	bool is_skybox = false; // SpecialMaterialEffect::Skybox (material.rs:118-132): polygons are not drawn, they open a window onto the sky box
	uint32_t skybox_images[2][6][NUM_MIPS] = {}; // SkyboxSideTexture<Color32> / <Color64> mips of the six sides (-X +X -Y +Y -Z +Z)
};

struct SceneParams
{
	uint32_t seed = 1;
	uint32_t grid[3] = {12, 12, 3};
	float cell_size[3] = {256.0f, 256.0f, 192.0f};
	uint32_t num_rooms = 7;
	uint32_t num_materials = 16;
	uint32_t texture_size_log2 = 8; // 256 x 256 world textures (demo/export_textures.py:27)
	float max_polygon_size = 128.0f; // split so that a lightmap stays <= 17 texels per side (bsp_builder.rs:1282)
	uint32_t material_mode = 0;      // 0: mix per demo_materials.json (normal map + dielectric/metal specular), 1: plain (no normal map, roughness 1), 2: as 0 plus translucent / alpha-tested / average / additive materials
	float emissive_fraction = 0.15f;
	uint32_t num_submodels = 0;    // inline models (crates beside the camera path), each a one-node submodel BSP of 6+ lit polygons
	uint32_t light_grid = 0;       // 1: generate a light grid (64-unit cells over the map) for model / sprite lighting
	uint32_t skybox = 0;           // 1: the first ceiling material becomes a sky box material (6 sides, 32- and 64-bit, 5 mips)
	uint32_t sky_size_log2 = 7;    // side of a sky box texture
	uint32_t num_models = 0;       // triangle-model fixtures to generate (models.hpp)
	uint32_t model_detail = 10;    // rings of the lathe grids: a model has about 4 * detail^2 triangles
};

struct Scene
{
	SceneParams params;
	BSPMap map;
	std::vector<Material> materials;
	std::vector<HostImage> images; // image id = index; the caller uploads them in this order
	// Triangle-model fixtures (SceneParams::num_models of them, kinds cycling NonAnimated / VertexAnimated / SkeletonAnimated)
	// and their skins: TextureLiteWithMips as NUM_MIPS consecutive SQW_IMAGE_COLOR32 image ids per skin.
	std::vector<TriangleModel> models;
	std::vector<std::array<uint32_t, NUM_MIPS>> model_skins;
	std::vector<Vec3f> camera_waypoints; // centres of the cells the fly-through visits
	// Compiled maps (bind_compiled_map): the waypoints are a poly-line through leaf portals that make_camera_path follows
	// linearly; waypoint_breaks[i] != 0 = the path jumps to waypoint i (a teleport: no interpolation across it)
	bool polyline_path = false;
	std::vector<uint8_t> waypoint_breaks;
	std::vector<uint8_t> solid;          // grid occupancy
	bool is_solid(int x, int y, int z) const;
	Vec3f cell_center(int x, int y, int z) const;
};

void generate_scene(const SceneParams& params, Scene& out);

// 64-pose (or n-pose) Catmull-Rom fly-through of SURVEY.md §8(d): position + azimuth + elevation per frame.
struct CameraPose
{
	Vec3f position;
	float azimuth, elevation;
};
void make_camera_path(const Scene& scene, uint32_t num_frames, uint32_t jitter_seed, std::vector<CameraPose>& out);

// textures.rs:309-365 / :388-421
void build_texture_mip(HostImage& mip, const HostImage& prev);
void build_texture_lite_mip(HostImage& mip, const HostImage& prev);

} // namespace sqwh
