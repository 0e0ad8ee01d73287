// Offline map tools of the host mirror: map compiler (map_compiler.cpp) and lightmapper (lightmapper.cpp),
// C++ restatements of /root/reference/src/map_compiler + src/lightmapper and the common/ modules they drive.
#pragma once
#include "map.hpp"
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace sqwh
{
namespace mapc
{

typedef std::vector<std::pair<std::string, std::string>> KeyList; // entity keys in file order (the reference: HashMap)

struct TextureInfo // map_polygonizer.rs:3-8
{
	Plane tex_coord_equation[2];
	std::string texture;
};
struct MapPolygon // map_polygonizer.rs:10-16
{
	Plane plane;
	TextureInfo texture_info;
	std::vector<Vec3f> vertices;
};
typedef std::vector<MapPolygon> Brush;
struct PolygonizedEntity
{
	std::vector<Brush> brushes;
	KeyList keys;
};
struct CsgEntity // map_csg.rs:9-14
{
	std::vector<MapPolygon> polygons;
	KeyList keys;
};

// The fields of common/material.rs:7-98 the tools and the renderer's material table read
struct MaterialDef
{
	std::string name;
	bool bsp = true, twosided = false, blocks_view = true, light = true, shadow = true, draw = true, decals = true, skybox = false;
	int blending_mode = 0; // SQW_BLEND_*
	bool normal_map = false, roughness_map = false, is_metal = false;
	float roughness = 1.0f;
	float emissive_light[3] = {0, 0, 0};
	bool emissive_layer = false; // material.rs EmissiveLayer {image, light}
	float emissive_layer_light[3] = {0, 0, 0};
};
typedef std::map<std::string, MaterialDef> MaterialTable;
MaterialTable parse_material_table(const std::string& text);
const MaterialDef* material_of_texture(const MaterialTable& materials, const TextureName& t);

enum PointPosition
{
	POINT_FRONT,
	POINT_BACK,
	POINT_ON_PLANE
};
enum PolygonPosition
{
	POLY_FRONT,
	POLY_BACK,
	POLY_COPLANAR_FRONT,
	POLY_COPLANAR_BACK,
	POLY_SPLITTED
};
PointPosition get_point_position_relative_plane(const Vec3f& point, const Plane& plane);
PolygonPosition get_polygon_position_relative_plane(const MapPolygon& polygon, const Plane& plane);
Vec3f get_line_plane_intersection(const Vec3f& v0, const Vec3f& v1, const Plane& plane);
void split_polygon(const MapPolygon& in, const Plane& plane, MapPolygon& front, MapPolygon& back);
std::vector<Vec3f> remove_duplicate_vertices(const std::vector<Vec3f>& in);
std::vector<Vec3f> sort_convex_polygon_vertices(std::vector<Vec3f> in, const Plane& plane);
bool parse_vec3(const std::string& s, Vec3f& out);
bool parse_key_value_number(const std::string& s, float& out);

struct CompileOptions // map_compiler/main.rs:9-36
{
	bool perform_csg = false;
	bool advanced_splitter_plane_selection = false;
};
struct CompileStats
{
	uint32_t initial_polygons = 0, polygons_after_csg = 0, leafs_initial = 0, leafs_removed = 0, tree_depth = 0;
	uint32_t nodes = 0, leafs = 0, polygons = 0, portals = 0, submodels = 0, lightmap_texels = 0;
};
bool compile_map(const std::string& map_text, const MaterialTable& materials, const CompileOptions& options, BSPMap& out_map, CompileStats& stats, std::string& log);

size_t allocate_lightmaps(const MaterialTable& materials, BSPMap& map);
void build_dummy_lightmaps(const MaterialTable& materials, BSPMap& map);

struct LightmappingSettings // lightmaps_builder.rs:8-20 (the subset of a direct-light run)
{
	uint32_t sample_grid_size = 1;
	float light_scale = 1.0f;
	float ambient_light = 0.0f;
	bool build_directional_lightmap = true;
	float light_grid_cell_width = 64.0f, light_grid_cell_height = 64.0f;
	float light_cutoff = 1.0f / 512.0f; // (not in the reference) a point light is skipped where its unshadowed contribution stays below this
	bool use_pvs = false;               // the reference's way instead: lights of the leafs the polygon's leaf can see (pvs.rs)
};
struct LightmapStats
{
	uint32_t point_lights = 0, sun_lights = 0, lightmap_texels = 0, light_grid_samples = 0;
};
void build_lightmaps(const LightmappingSettings& settings, const MaterialTable& materials, BSPMap& map, LightmapStats& stats);
// pvs.rs calculate_visibility_matrix: [leaf a][leaf b] != 0 when each of the two leafs can potentially see the other
std::vector<uint8_t> calculate_leafs_visibility(const BSPMap& map);

} // namespace mapc

// A compiled (and lit) map as the Scene the renderer mirror draws: materials of the map's texture names from the table (seeded
// procedural textures of the kind each definition asks for), sky box sides for sky materials, camera path through the map's
// player start and teleport destinations (map.cpp).
void bind_compiled_map(Scene& s, BSPMap&& map, const mapc::MaterialTable& table, uint32_t texture_size_log2, uint32_t seed);
} // namespace sqwh
