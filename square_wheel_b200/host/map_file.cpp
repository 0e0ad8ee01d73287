// .sqwm v11 reader / writer (src/common/bsp_map_save_load.rs) and the lossy light compression it applies to lightmaps and
// light grid samples (src/common/lightmap_compression.rs).  A map compiled and lit by the reference's own tools loads into
// the same in-memory BSPMap the synthetic generator produces.
#include "map.hpp"
#include <cstdio>
#include <cstring>

namespace sqwh
{

namespace
{
constexpr char BSP_MAP_ID[4] = {'S', 'q', 'w', 'M'};
constexpr uint32_t BSP_MAP_VERSION = 11; // bsp_map_save_load.rs:244
constexpr size_t MAX_LUMPS = 32;
enum
{
	LUMP_NODES = 0,
	LUMP_LEAFS,
	LUMP_POLYGONS,
	LUMP_PORTALS,
	LUMP_LEAFS_PORTALS,
	LUMP_VERTICES,
	LUMP_TEXTURES,
	LUMP_SUBMODELS,
	LUMP_SUBMODELS_BSP_NODES,
	LUMP_ENTITIES,
	LUMP_KEY_VALUE_PAIRS,
	LUMP_STRINGS_DATA,
	LUMP_LIGHTMAPS_DATA,
	LUMP_DIRECTIONAL_LIGHTMAPS_DATA,
	LUMP_LIGHT_GRID_HEADER,
	LUMP_LIGHT_GRID_COLUMNS,
	LUMP_LIGHT_GRID_SAMPLES,
};
struct Lump
{
	uint32_t offset, element_size, element_count;
};
struct BspMapHeader
{
	char id[4];
	uint32_t version;
	Lump lumps[MAX_LUMPS];
};
static_assert(sizeof(BspMapHeader) == 392, ".sqwm header");

// ---- lightmap_compression.rs -------------------------------------------------------------------------------------
constexpr float COLOR_SCALE = 255.0f, VECTOR_SCALE = 127.0f, LIGHT_DEVIATION_SCALE = 255.0f;
struct CompressedColor
{
	uint8_t components[3], scale;
};
struct CompressedVector
{
	int8_t components[3];
	uint8_t scale;
};
struct DirectionalLightmapElementCompressed
{
	CompressedColor ambient_light;
	CompressedVector light_direction_vector_scaled;
	uint8_t directional_light_deviation;
	CompressedColor directional_light_color;
};
struct LightGridElementCompressed
{
	CompressedColor light_cube[6];
	CompressedVector light_direction_vector_scaled;
	CompressedColor directional_light_color;
};
static_assert(sizeof(CompressedColor) == 4 && sizeof(DirectionalLightmapElementCompressed) == 13 && sizeof(LightGridElementCompressed) == 32, "compressed light element sizes");

inline uint8_t f32_as_u8(float v) // Rust `as u8`: saturating, NaN -> 0
{
	if (!(v == v) || v <= 0.0f)
		return 0;
	return v >= 255.0f ? 255 : (uint8_t)v;
}
inline int8_t f32_as_i8(float v)
{
	if (!(v == v))
		return 0;
	return v >= 127.0f ? 127 : (v <= -128.0f ? -128 : (int8_t)v);
}
CompressedColor compress_color(const float color[3]) // :48-60
{
	const float max_component_value = std::fmax(std::fmax(color[0], color[1]), color[2]);
	const float scale = std::fmin(std::fmax(std::ceil(max_component_value), 1.0f), 255.0f);
	const float multiplier = COLOR_SCALE / scale;
	CompressedColor c;
	for (int i = 0; i < 3; ++i)
		c.components[i] = f32_as_u8(std::fmin(std::fmax(color[i] * multiplier, 0.0f), 255.0f));
	c.scale = f32_as_u8(scale);
	return c;
}
void decompress_color(const CompressedColor& c, float out[3]) // :62-66
{
	const float multiplier = (float)c.scale / COLOR_SCALE;
	for (int i = 0; i < 3; ++i)
		out[i] = (float)c.components[i] * multiplier;
}
CompressedVector compress_vector(Vec3f v) // :71-83
{
	const float max_component_value = std::fmax(std::fmax(std::fabs(v.x), std::fabs(v.y)), std::fabs(v.z));
	const float scale = std::fmin(std::fmax(std::ceil(max_component_value), 1.0f), 255.0f);
	const float multiplier = VECTOR_SCALE / scale;
	const float comp[3] = {v.x, v.y, v.z};
	CompressedVector c;
	for (int i = 0; i < 3; ++i)
		c.components[i] = f32_as_i8(std::fmin(std::fmax(comp[i] * multiplier, -127.0f), 127.0f));
	c.scale = f32_as_u8(scale);
	return c;
}
Vec3f decompress_vector(const CompressedVector& c) // :85-94
{
	const float multiplier = (float)c.scale / VECTOR_SCALE;
	return Vec3f{(float)c.components[0], (float)c.components[1], (float)c.components[2]} * multiplier;
}

template <class T> bool write_lump(FILE* f, const T* data, size_t count, Lump& lump, size_t& offset) // :262-285
{
	lump.offset = (uint32_t)offset;
	lump.element_size = (uint32_t)sizeof(T);
	lump.element_count = (uint32_t)count;
	if (count == 0)
		return true;
	if (std::fwrite(data, sizeof(T), count, f) != count)
		return false;
	offset += sizeof(T) * count;
	return true;
}
template <class T> bool read_lump(FILE* f, const Lump& lump, std::vector<T>& out) // :287-313: a lump of another element size reads as EMPTY
{
	out.clear();
	if (lump.element_size != sizeof(T))
		return true;
	out.resize(lump.element_count);
	if (out.empty())
		return true;
	if (std::fseek(f, (long)lump.offset, SEEK_SET) != 0)
		return false;
	return std::fread(out.data(), sizeof(T), out.size(), f) == out.size();
}
} // namespace

MapFileResult save_map(const BSPMap& map, const char* file_path)
{
	FILE* f = std::fopen(file_path, "wb");
	if (!f)
		return MAP_FILE_IO_ERROR;
	BspMapHeader header;
	std::memset(&header, 0, sizeof(header));
	std::memcpy(header.id, BSP_MAP_ID, 4);
	header.version = BSP_MAP_VERSION;
	bool ok = std::fwrite(&header, sizeof(header), 1, f) == 1; // written again at the end with the lump table filled in
	size_t offset = sizeof(header);
	ok = ok && write_lump(f, map.nodes.data(), map.nodes.size(), header.lumps[LUMP_NODES], offset);
	ok = ok && write_lump(f, map.leafs.data(), map.leafs.size(), header.lumps[LUMP_LEAFS], offset);
	ok = ok && write_lump(f, map.polygons.data(), map.polygons.size(), header.lumps[LUMP_POLYGONS], offset);
	ok = ok && write_lump(f, map.portals.data(), map.portals.size(), header.lumps[LUMP_PORTALS], offset);
	ok = ok && write_lump(f, map.leafs_portals.data(), map.leafs_portals.size(), header.lumps[LUMP_LEAFS_PORTALS], offset);
	ok = ok && write_lump(f, map.vertices.data(), map.vertices.size(), header.lumps[LUMP_VERTICES], offset);
	ok = ok && write_lump(f, map.textures.data(), map.textures.size(), header.lumps[LUMP_TEXTURES], offset);
	ok = ok && write_lump(f, map.submodels.data(), map.submodels.size(), header.lumps[LUMP_SUBMODELS], offset);
	ok = ok && write_lump(f, map.submodels_bsp_nodes.data(), map.submodels_bsp_nodes.size(), header.lumps[LUMP_SUBMODELS_BSP_NODES], offset);
	ok = ok && write_lump(f, map.entities.data(), map.entities.size(), header.lumps[LUMP_ENTITIES], offset);
	ok = ok && write_lump(f, map.key_value_pairs.data(), map.key_value_pairs.size(), header.lumps[LUMP_KEY_VALUE_PAIRS], offset);
	ok = ok && write_lump(f, map.strings_data.data(), map.strings_data.size(), header.lumps[LUMP_STRINGS_DATA], offset);
	{
		std::vector<CompressedColor> c(map.lightmaps_data.size() / 3);
		for (size_t i = 0; i < c.size(); ++i)
			c[i] = compress_color(&map.lightmaps_data[i * 3]);
		ok = ok && write_lump(f, c.data(), c.size(), header.lumps[LUMP_LIGHTMAPS_DATA], offset);
	}
	{
		std::vector<DirectionalLightmapElementCompressed> c(map.directional_lightmaps_data.size());
		for (size_t i = 0; i < c.size(); ++i) // lightmap_compression.rs:99-110
		{
			const sqw_directional_lightmap_element& e = map.directional_lightmaps_data[i];
			c[i].ambient_light = compress_color(e.ambient_light);
			c[i].light_direction_vector_scaled = compress_vector(Vec3f{e.light_direction_vector_scaled[0], e.light_direction_vector_scaled[1], e.light_direction_vector_scaled[2]});
			c[i].directional_light_deviation = f32_as_u8(std::fmin(std::fmax(e.directional_light_deviation * LIGHT_DEVIATION_SCALE, 0.0f), 255.0f));
			c[i].directional_light_color = compress_color(e.directional_light_color);
		}
		ok = ok && write_lump(f, c.data(), c.size(), header.lumps[LUMP_DIRECTIONAL_LIGHTMAPS_DATA], offset);
	}
	ok = ok && write_lump(f, &map.light_grid_header, 1, header.lumps[LUMP_LIGHT_GRID_HEADER], offset); // write_single_element_lump
	ok = ok && write_lump(f, map.light_grid_columns.data(), map.light_grid_columns.size(), header.lumps[LUMP_LIGHT_GRID_COLUMNS], offset);
	{
		std::vector<LightGridElementCompressed> c(map.light_grid_samples.size());
		for (size_t i = 0; i < c.size(); ++i) // :128-137
		{
			const LightGridElement& e = map.light_grid_samples[i];
			for (int s = 0; s < 6; ++s)
				c[i].light_cube[s] = compress_color(e.light_cube[s]);
			c[i].light_direction_vector_scaled = compress_vector(e.light_direction_vector_scaled);
			c[i].directional_light_color = compress_color(e.directional_light_color);
		}
		ok = ok && write_lump(f, c.data(), c.size(), header.lumps[LUMP_LIGHT_GRID_SAMPLES], offset);
	}
	ok = ok && std::fseek(f, 0, SEEK_SET) == 0 && std::fwrite(&header, sizeof(header), 1, f) == 1;
	ok = (std::fclose(f) == 0) && ok;
	return ok ? MAP_FILE_OK : MAP_FILE_IO_ERROR;
}

MapFileResult load_map(const char* file_path, BSPMap& out)
{
	FILE* f = std::fopen(file_path, "rb");
	if (!f)
		return MAP_FILE_IO_ERROR;
	BspMapHeader header;
	if (std::fread(&header, sizeof(header), 1, f) != 1)
	{
		std::fclose(f);
		return MAP_FILE_IO_ERROR; // "Can't read BSP map header"
	}
	if (std::memcmp(header.id, BSP_MAP_ID, 4) != 0)
	{
		std::fclose(f);
		return MAP_FILE_NOT_A_MAP;
	}
	if (header.version != BSP_MAP_VERSION)
	{
		std::fclose(f);
		return MAP_FILE_BAD_VERSION;
	}
	BSPMap map;
	bool ok = read_lump(f, header.lumps[LUMP_NODES], map.nodes);
	ok = ok && read_lump(f, header.lumps[LUMP_LEAFS], map.leafs);
	ok = ok && read_lump(f, header.lumps[LUMP_POLYGONS], map.polygons);
	ok = ok && read_lump(f, header.lumps[LUMP_PORTALS], map.portals);
	ok = ok && read_lump(f, header.lumps[LUMP_LEAFS_PORTALS], map.leafs_portals);
	ok = ok && read_lump(f, header.lumps[LUMP_VERTICES], map.vertices);
	ok = ok && read_lump(f, header.lumps[LUMP_TEXTURES], map.textures);
	ok = ok && read_lump(f, header.lumps[LUMP_SUBMODELS], map.submodels);
	ok = ok && read_lump(f, header.lumps[LUMP_SUBMODELS_BSP_NODES], map.submodels_bsp_nodes);
	ok = ok && read_lump(f, header.lumps[LUMP_ENTITIES], map.entities);
	ok = ok && read_lump(f, header.lumps[LUMP_KEY_VALUE_PAIRS], map.key_value_pairs);
	ok = ok && read_lump(f, header.lumps[LUMP_STRINGS_DATA], map.strings_data);
	{
		std::vector<CompressedColor> c;
		ok = ok && read_lump(f, header.lumps[LUMP_LIGHTMAPS_DATA], c);
		map.lightmaps_data.resize(c.size() * 3);
		for (size_t i = 0; i < c.size(); ++i)
			decompress_color(c[i], &map.lightmaps_data[i * 3]);
	}
	{
		std::vector<DirectionalLightmapElementCompressed> c;
		ok = ok && read_lump(f, header.lumps[LUMP_DIRECTIONAL_LIGHTMAPS_DATA], c);
		map.directional_lightmaps_data.resize(c.size());
		for (size_t i = 0; i < c.size(); ++i) // lightmap_compression.rs:112-128
		{
			sqw_directional_lightmap_element& e = map.directional_lightmaps_data[i];
			decompress_color(c[i].ambient_light, e.ambient_light);
			Vec3f v = decompress_vector(c[i].light_direction_vector_scaled);
			if (dot(v, v) <= 0.0f)
				v = Vec3f{0.0f, 0.0f, 1.0f / 1024.0f}; // must be non-zero; quantisation can zero it
			e.light_direction_vector_scaled[0] = v.x;
			e.light_direction_vector_scaled[1] = v.y;
			e.light_direction_vector_scaled[2] = v.z;
			e.directional_light_deviation = (float)c[i].directional_light_deviation / LIGHT_DEVIATION_SCALE;
			decompress_color(c[i].directional_light_color, e.directional_light_color);
		}
	}
	{
		std::vector<LightGridHeader> h;
		ok = ok && read_lump(f, header.lumps[LUMP_LIGHT_GRID_HEADER], h); // read_single_element_lump: zeroed when absent / wrong size
		map.light_grid_header = h.empty() ? LightGridHeader{} : h[0];
	}
	ok = ok && read_lump(f, header.lumps[LUMP_LIGHT_GRID_COLUMNS], map.light_grid_columns);
	{
		std::vector<LightGridElementCompressed> c;
		ok = ok && read_lump(f, header.lumps[LUMP_LIGHT_GRID_SAMPLES], c);
		map.light_grid_samples.resize(c.size());
		for (size_t i = 0; i < c.size(); ++i) // :139-146
		{
			LightGridElement& e = map.light_grid_samples[i];
			for (int s = 0; s < 6; ++s)
				decompress_color(c[i].light_cube[s], e.light_cube[s]);
			e.light_direction_vector_scaled = decompress_vector(c[i].light_direction_vector_scaled);
			decompress_color(c[i].directional_light_color, e.directional_light_color);
		}
	}
	std::fclose(f);
	if (!ok)
		return MAP_FILE_IO_ERROR;
	out = std::move(map);
	return MAP_FILE_OK;
}

std::string validate_map(const BSPMap& map, size_t num_materials)
{
	if (map.leafs.empty())
		return "no leafs";
	const size_t num_lm = map.directional_lightmaps_data.size(), num_simple = map.lightmaps_data.size() / 3;
	for (const BSPNode& n : map.nodes)
		for (uint32_t c : n.children)
			if (c >= FIRST_LEAF_INDEX ? (c - FIRST_LEAF_INDEX >= map.leafs.size()) : (c >= map.nodes.size()))
				return "BSP node child out of range";
	for (const BSPLeaf& l : map.leafs)
		if ((uint64_t)l.first_polygon + l.num_polygons > map.polygons.size() || (uint64_t)l.first_leaf_portal + l.num_leaf_portals > map.leafs_portals.size())
			return "leaf polygon / portal range out of range";
	for (uint32_t p : map.leafs_portals)
		if (p >= map.portals.size())
			return "leaf portal index out of range";
	for (const Portal& p : map.portals)
		if (p.leafs[0] >= map.leafs.size() || p.leafs[1] >= map.leafs.size() || (uint64_t)p.first_vertex + p.num_vertices > map.vertices.size() || p.num_vertices < 3)
			return "portal out of range";
	for (const Polygon& p : map.polygons)
	{
		if ((uint64_t)p.first_vertex + p.num_vertices > map.vertices.size() || p.num_vertices < 3)
			return "polygon vertices out of range";
		if (p.texture >= num_materials)
			return "polygon texture index has no material";
		if (p.tex_coord_max[0] < p.tex_coord_min[0] || p.tex_coord_max[1] < p.tex_coord_min[1])
			return "polygon texture coordinate bounds inverted";
		if (p.lightmap_data_offset != 0)
		{
			const uint64_t n = (uint64_t)get_lightmap_size(p.tex_coord_min[0], p.tex_coord_max[0]) * get_lightmap_size(p.tex_coord_min[1], p.tex_coord_max[1]);
			if ((num_simple && p.lightmap_data_offset + n > num_simple) || (num_lm && p.lightmap_data_offset + n > num_lm))
				return "polygon lightmap outside the lightmap data";
		}
	}
	const LightGridHeader& h = map.light_grid_header;
	if (!map.light_grid_columns.empty() && !map.light_grid_samples.empty()) // without samples the grid is simply unused (fetch_light_from_grid :586)
	{
		if ((uint64_t)h.grid_size[0] * h.grid_size[1] != map.light_grid_columns.size())
			return "light grid columns do not match the grid size";
		for (const LightGridColumn& c : map.light_grid_columns)
			if ((uint64_t)c.first_sample + c.num_samples > map.light_grid_samples.size())
				return "light grid column outside the samples";
	}
	return "";
}

} // namespace sqwh
