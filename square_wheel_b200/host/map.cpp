// Synthetic map generator + texture / mip construction (see map.hpp).
//
// The generated map has the structure the reference's renderer expects from map_compiler + lightmapper
// (docs in SURVEY.md appendix A): a leaf BSP whose leaves are convex (axis-aligned boxes of a carved voxel grid),
// portals between neighbouring empty leaves, wall polygons split so that their lightmaps stay <= 17 texels per side,
// 2 texels per map unit, one lightmap texel per 16 surface texels, simple + directional lightmaps sharing offsets
// (offset 0 = no lightmap).  Lighting is a direct-light-only stand-in for the lightmapper with voxel shadow rays.
#include "map.hpp"
#include "map_compiler.hpp"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>

namespace sqwh
{

bool Scene::is_solid(int x, int y, int z) const
{
	const int nx = (int)params.grid[0], ny = (int)params.grid[1], nz = (int)params.grid[2];
	if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz)
		return true;
	return solid[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z)] != 0;
}
Vec3f Scene::cell_center(int x, int y, int z) const
{
	return {((float)x + 0.5f) * params.cell_size[0], ((float)y + 0.5f) * params.cell_size[1], ((float)z + 0.5f) * params.cell_size[2]};
}

// ---- textures ---------------------------------------------------------------------------------------------------
namespace
{

struct TextureElement
{
	uint32_t diffuse;
	int32_t packed;
};

int32_t pack_normal_roughness(Vec3f n, float roughness) // textures.rs:34-40, fast_math.rs:418-427
{
	const float v[4] = {n.x, n.y, n.z, roughness};
	int32_t res = 0;
	for (int i = 0; i < 4; ++i)
		res |= (((int32_t)std::fmin(std::fmax(v[i] * 127.0f, -127.0f), 127.0f)) & 0xFF) << (i * 8);
	return res;
}
void unpack_normal_roughness(int32_t b, Vec3f& n, float& roughness) // textures.rs:42-58
{
	const float s = 1.0f / 127.0f;
	n.x = (float)((int32_t)((uint32_t)b << 24) >> 24) * s;
	n.y = (float)((int32_t)((uint32_t)b << 16) >> 24) * s;
	n.z = (float)((int32_t)((uint32_t)b << 8) >> 24) * s;
	roughness = (float)(b >> 24) * s;
}
uint32_t color32_average_4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) // common/color.rs:51-74
{
	uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
	for (uint32_t col : {a, b, c, d})
	{
		s0 += col & 0x000000FFu;
		s1 += col & 0x0000FF00u;
		s2 += col & 0x00FF0000u;
		s3 += (col & 0xFF000000u) >> 2;
	}
	return (s0 >> 2) | ((s1 >> 2) & 0x0000FF00u) | ((s2 >> 2) & 0x00FF0000u) | (s3 & 0xFF000000u);
}

// Tileable value noise in [0, 1].
float value_noise(Pcg32&, const std::vector<float>& lattice, uint32_t cells, float x, float y)
{
	const float fx = x * (float)cells, fy = y * (float)cells;
	const uint32_t x0 = (uint32_t)fx % cells, y0 = (uint32_t)fy % cells;
	const uint32_t x1 = (x0 + 1) % cells, y1 = (y0 + 1) % cells;
	float tx = fx - std::floor(fx), ty = fy - std::floor(fy);
	tx = tx * tx * (3 - 2 * tx);
	ty = ty * ty * (3 - 2 * ty);
	const float a = lattice[x0 + y0 * cells], b = lattice[x1 + y0 * cells], c = lattice[x0 + y1 * cells], d = lattice[x1 + y1 * cells];
	return (a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty;
}

// alpha_mode: 0 opaque; 1 smooth translucency (noise, 48..208: alpha_blend / stained_glass of demo_materials.json); 2 a grid of
// holes (alpha 0 inside the tiles, 255 on the bars: alpha_grid)
void make_world_texture(Pcg32& rng, uint32_t size, bool normal_map, bool roughness_map, float base_roughness, HostImage& out, int alpha_mode = 0)
{
	out.kind = SQW_IMAGE_TEXTURE_ELEMENT;
	out.width = out.height = size;
	out.pixels.resize((size_t)size * size * 8);
	TextureElement* px = reinterpret_cast<TextureElement*>(out.pixels.data());
	const uint32_t c1 = 4, c2 = 16, c3 = 64;
	std::vector<float> l1(c1 * c1), l2(c2 * c2), l3(c3 * c3);
	for (auto* l : {&l1, &l2, &l3})
		for (float& v : *l)
			v = rng.uniform();
	const float base[3] = {rng.uniform(0.25f, 0.9f), rng.uniform(0.25f, 0.9f), rng.uniform(0.25f, 0.9f)};
	const float tint[3] = {rng.uniform(-0.3f, 0.3f), rng.uniform(-0.3f, 0.3f), rng.uniform(-0.3f, 0.3f)};
	const uint32_t tiles = 1u << rng.below(4); // brick / tile pattern
	const float groove = rng.uniform(0.02f, 0.06f);
	std::vector<float> height((size_t)size * size);
	for (uint32_t y = 0; y < size; ++y)
		for (uint32_t x = 0; x < size; ++x)
		{
			const float u = ((float)x + 0.5f) / (float)size, v = ((float)y + 0.5f) / (float)size;
			const float n = 0.5f * value_noise(rng, l1, c1, u, v) + 0.3f * value_noise(rng, l2, c2, u, v) + 0.2f * value_noise(rng, l3, c3, u, v);
			float fu = u * (float)tiles, fv = v * (float)tiles;
			fu -= std::floor(fu);
			fv -= std::floor(fv);
			const float edge = std::fmin(std::fmin(fu, 1 - fu), std::fmin(fv, 1 - fv));
			const float g = edge < groove ? edge / groove : 1.0f;
			height[x + (size_t)y * size] = 0.6f * g + 0.4f * n;
			uint32_t rgb[3];
			for (int c = 0; c < 3; ++c)
			{
				float val = (base[c] + tint[c] * (n - 0.5f) * 2.0f) * (0.55f + 0.45f * g) * (0.75f + 0.5f * n);
				rgb[c] = (uint32_t)std::fmin(std::fmax(val * 255.0f, 0.0f), 255.0f);
			}
			uint32_t alpha = 255;
			if (alpha_mode == 1)
				alpha = (uint32_t)std::fmin(std::fmax(48.0f + 160.0f * n, 0.0f), 255.0f);
			else if (alpha_mode == 2)
				alpha = edge < 4.0f * groove ? 255u : 0u;
			px[x + (size_t)y * size].diffuse = (alpha << 24) | (rgb[0] << 16) | (rgb[1] << 8) | rgb[2];
		}
	const float bump = rng.uniform(1.5f, 5.0f);
	for (uint32_t y = 0; y < size; ++y)
		for (uint32_t x = 0; x < size; ++x)
		{
			Vec3f n{0, 0, 1};
			if (normal_map)
			{
				const float h = height[x + (size_t)y * size];
				const float hx = height[(x + 1) % size + (size_t)y * size], hy = height[x + (size_t)((y + 1) % size) * size];
				n = normalize(Vec3f{-(hx - h) * bump, -(hy - h) * bump, 1.0f});
			}
			float roughness = base_roughness;
			if (roughness_map)
				roughness = std::fmin(std::fmax(base_roughness * (0.5f + height[x + (size_t)y * size]), 1.0f / 126.0f), 1.0f); // textures.rs:418
			px[x + (size_t)y * size].packed = pack_normal_roughness(n, roughness);
		}
}

void make_emissive_texture(Pcg32& rng, uint32_t size, HostImage& out)
{
	out.kind = SQW_IMAGE_COLOR32;
	out.width = out.height = size;
	out.pixels.resize((size_t)size * size * 4);
	uint32_t* px = reinterpret_cast<uint32_t*>(out.pixels.data());
	const uint32_t cells = 8;
	std::vector<float> lat(cells * cells);
	for (float& v : lat)
		v = rng.uniform();
	const float col[3] = {rng.uniform(0.3f, 1.0f), rng.uniform(0.3f, 1.0f), rng.uniform(0.3f, 1.0f)};
	for (uint32_t y = 0; y < size; ++y)
		for (uint32_t x = 0; x < size; ++x)
		{
			const float n = value_noise(rng, lat, cells, ((float)x + 0.5f) / (float)size, ((float)y + 0.5f) / (float)size);
			const float glow = n > 0.62f ? (n - 0.62f) / 0.38f : 0.0f; // glowing veins / buttons
			uint32_t rgb[3];
			for (int c = 0; c < 3; ++c)
				rgb[c] = (uint32_t)(col[c] * glow * 255.0f);
			px[x + (size_t)y * size] = 0xFF000000u | (rgb[0] << 16) | (rgb[1] << 8) | rgb[2];
		}
}

} // namespace

void build_texture_mip(HostImage& mip, const HostImage& prev) // textures.rs:309-365
{
	mip.kind = SQW_IMAGE_TEXTURE_ELEMENT;
	mip.width = prev.width >> 1;
	mip.height = prev.height >> 1;
	mip.pixels.resize((size_t)mip.width * mip.height * 8);
	const TextureElement* src = reinterpret_cast<const TextureElement*>(prev.pixels.data());
	TextureElement* dst = reinterpret_cast<TextureElement*>(mip.pixels.data());
	const float MIN_VALID_ROUGHNESS = 1.0f / (127.0f - 1.0f);
	for (uint32_t y = 0; y < mip.height; ++y)
		for (uint32_t x = 0; x < mip.width; ++x)
		{
			const TextureElement p00 = src[x * 2 + (size_t)(y * 2) * prev.width], p01 = src[x * 2 + (size_t)(y * 2 + 1) * prev.width];
			const TextureElement p10 = src[x * 2 + 1 + (size_t)(y * 2) * prev.width], p11 = src[x * 2 + 1 + (size_t)(y * 2 + 1) * prev.width];
			TextureElement& d = dst[x + (size_t)y * mip.width];
			d.diffuse = color32_average_4(p00.diffuse, p01.diffuse, p10.diffuse, p11.diffuse);
			Vec3f n00, n01, n10, n11;
			float r00, r01, r10, r11;
			unpack_normal_roughness(p00.packed, n00, r00);
			unpack_normal_roughness(p01.packed, n01, r01);
			unpack_normal_roughness(p10.packed, n10, r10);
			unpack_normal_roughness(p11.packed, n11, r11);
			const Vec3f normals_sum = n00 + n01 + n10 + n11;
			const float normals_sum_len = std::fmax(magnitude(normals_sum), 0.000001f);
			const Vec3f dst_normal = normals_sum / normals_sum_len;
			const float half_normal_deviation_cos = normals_sum_len * 0.25f;
			const float normal_deviation = (1.0f - half_normal_deviation_cos) / (1.0f - 0.5f);
			const float roughness_sum = r00 + r01 + r10 + r11;
			const float dst_roughness = std::fmin(std::fmax(roughness_sum * 0.25f + normal_deviation, MIN_VALID_ROUGHNESS), 1.0f);
			d.packed = pack_normal_roughness(dst_normal, dst_roughness);
		}
}

void build_texture_lite_mip(HostImage& mip, const HostImage& prev) // textures.rs:388-421
{
	mip.kind = SQW_IMAGE_COLOR32;
	mip.width = prev.width >> 1;
	mip.height = prev.height >> 1;
	mip.pixels.resize((size_t)mip.width * mip.height * 4);
	const uint32_t* src = reinterpret_cast<const uint32_t*>(prev.pixels.data());
	uint32_t* dst = reinterpret_cast<uint32_t*>(mip.pixels.data());
	for (uint32_t y = 0; y < mip.height; ++y)
		for (uint32_t x = 0; x < mip.width; ++x)
			dst[x + (size_t)y * mip.width] = color32_average_4(
				src[x * 2 + (size_t)(y * 2) * prev.width], src[x * 2 + (size_t)(y * 2 + 1) * prev.width], src[x * 2 + 1 + (size_t)(y * 2) * prev.width],
				src[x * 2 + 1 + (size_t)(y * 2 + 1) * prev.width]);
}

// ---- map generation ------------------------------------------------------------------------------------------------
namespace
{

struct Box
{
	int lo[3], hi[3]; // cell index range [lo, hi)
};

struct MapLight
{
	Vec3f pos;
	float color[3];
};

struct Builder
{
	Scene& s;
	Pcg32 rng;
	int n[3];
	std::vector<uint32_t> cell_leaf;
	std::vector<Box> leaf_boxes;
	std::vector<uint8_t> leaf_solid;
	std::vector<MapLight> lights;

	Builder(Scene& scene) : s(scene), rng(scene.params.seed)
	{
		for (int i = 0; i < 3; ++i)
			n[i] = (int)s.params.grid[i];
	}
	size_t cell_index(int x, int y, int z) const { return (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z); }
	void carve(const Box& b)
	{
		for (int z = b.lo[2]; z < b.hi[2]; ++z)
			for (int y = b.lo[1]; y < b.hi[1]; ++y)
				for (int x = b.lo[0]; x < b.hi[0]; ++x)
					s.solid[cell_index(x, y, z)] = 0;
	}

	void carve_layout()
	{
		s.solid.assign((size_t)n[0] * n[1] * n[2], 1);
		std::vector<Box> rooms;
		const uint32_t num_rooms = std::max(2u, s.params.num_rooms);
		for (uint32_t attempt = 0; rooms.size() < num_rooms && attempt < num_rooms * 40; ++attempt)
		{
			Box b;
			int size[3] = {2 + (int)rng.below(3), 2 + (int)rng.below(3), 1 + (int)rng.below(2)};
			bool ok = true;
			for (int a = 0; a < 3; ++a)
			{
				size[a] = std::min(size[a], n[a]);
				b.lo[a] = (int)rng.below((uint32_t)(n[a] - size[a] + 1));
				b.hi[a] = b.lo[a] + size[a];
			}
			for (const Box& o : rooms) // keep rooms apart by one cell in x/y
				if (b.lo[0] < o.hi[0] + 1 && o.lo[0] < b.hi[0] + 1 && b.lo[1] < o.hi[1] + 1 && o.lo[1] < b.hi[1] + 1 && b.lo[2] < o.hi[2] && o.lo[2] < b.hi[2])
					ok = false;
			if (ok)
				rooms.push_back(b);
		}
		for (const Box& b : rooms)
			carve(b);
		// pillars inside large rooms
		for (const Box& b : rooms)
			if (b.hi[0] - b.lo[0] >= 3 && b.hi[1] - b.lo[1] >= 3 && rng.uniform() < 0.7f)
			{
				const int px = b.lo[0] + 1 + (int)rng.below((uint32_t)(b.hi[0] - b.lo[0] - 2));
				const int py = b.lo[1] + 1 + (int)rng.below((uint32_t)(b.hi[1] - b.lo[1] - 2));
				for (int z = b.lo[2]; z < b.hi[2]; ++z)
					s.solid[cell_index(px, py, z)] = 1;
			}
		// L-shaped corridors between consecutive rooms (and back to the first one: the fly-through is a loop)
		auto room_anchor = [&](const Box& b, int out[3]) {
			out[0] = b.lo[0];
			out[1] = b.lo[1];
			out[2] = b.lo[2];
		};
		std::vector<std::array<int, 3>> anchors;
		for (size_t i = 0; i < rooms.size(); ++i)
		{
			int a[3], b[3];
			room_anchor(rooms[i], a);
			room_anchor(rooms[(i + 1) % rooms.size()], b);
			anchors.push_back({a[0], a[1], a[2]});
			int c[3] = {a[0], a[1], a[2]};
			auto step_to = [&](int axis, int target) {
				while (c[axis] != target)
				{
					c[axis] += target > c[axis] ? 1 : -1;
					s.solid[cell_index(c[0], c[1], c[2])] = 0;
				}
			};
			s.solid[cell_index(c[0], c[1], c[2])] = 0;
			if (rng.uniform() < 0.5f)
			{
				step_to(0, b[0]);
				step_to(1, b[1]);
			}
			else
			{
				step_to(1, b[1]);
				step_to(0, b[0]);
			}
			step_to(2, b[2]);
		}
		// camera waypoints: BFS path through empty cells from anchor to anchor
		auto bfs = [&](std::array<int, 3> from, std::array<int, 3> to, std::vector<std::array<int, 3>>& path) {
			std::vector<int32_t> prev((size_t)n[0] * n[1] * n[2], -2);
			std::vector<std::array<int, 3>> queue{from};
			prev[cell_index(from[0], from[1], from[2])] = -1;
			for (size_t qi = 0; qi < queue.size(); ++qi)
			{
				auto c = queue[qi];
				if (c == to)
					break;
				const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
				for (auto& dd : d)
				{
					std::array<int, 3> nb{c[0] + dd[0], c[1] + dd[1], c[2] + dd[2]};
					if (s.is_solid(nb[0], nb[1], nb[2]))
						continue;
					const size_t idx = cell_index(nb[0], nb[1], nb[2]);
					if (prev[idx] != -2)
						continue;
					prev[idx] = (int32_t)cell_index(c[0], c[1], c[2]);
					queue.push_back(nb);
				}
			}
			std::vector<std::array<int, 3>> rev;
			int32_t cur = (int32_t)cell_index(to[0], to[1], to[2]);
			if (prev[(size_t)cur] == -2)
				return;
			while (cur != -1)
			{
				rev.push_back({cur % n[0], (cur / n[0]) % n[1], cur / (n[0] * n[1])});
				cur = prev[(size_t)cur];
			}
			for (size_t i = rev.size(); i-- > 1;) // exclude the target itself (it starts the next leg)
				path.push_back(rev[i]);
		};
		std::vector<std::array<int, 3>> cells;
		for (size_t i = 0; i < anchors.size(); ++i)
			bfs(anchors[i], anchors[(i + 1) % anchors.size()], cells);
		if (cells.empty())
			cells.push_back(anchors[0]);
		s.camera_waypoints.clear();
		for (auto& c : cells)
			s.camera_waypoints.push_back(s.cell_center(c[0], c[1], c[2]));
		// lights: one per room under the ceiling + one every few corridor cells
		for (const Box& b : rooms)
		{
			MapLight l;
			l.pos = {(float)(b.lo[0] + b.hi[0]) * 0.5f * s.params.cell_size[0], (float)(b.lo[1] + b.hi[1]) * 0.5f * s.params.cell_size[1],
					 ((float)b.hi[2] - 0.25f) * s.params.cell_size[2]};
			// keep the light out of a pillar
			const int cx = (int)(l.pos.x / s.params.cell_size[0]), cy = (int)(l.pos.y / s.params.cell_size[1]);
			if (s.is_solid(cx, cy, b.hi[2] - 1))
				l.pos.x += s.params.cell_size[0];
			const float power = rng.uniform(60000.0f, 160000.0f);
			const float warm = rng.uniform();
			l.color[0] = power * (0.7f + 0.3f * warm);
			l.color[1] = power * 0.8f;
			l.color[2] = power * (1.0f - 0.4f * warm);
			lights.push_back(l);
		}
		for (size_t i = 0; i < cells.size(); i += 3)
		{
			MapLight l;
			l.pos = s.cell_center(cells[i][0], cells[i][1], cells[i][2]);
			l.pos.z += s.params.cell_size[2] * 0.3f;
			const float power = rng.uniform(15000.0f, 40000.0f);
			l.color[0] = power * rng.uniform(0.5f, 1.0f);
			l.color[1] = power * rng.uniform(0.5f, 1.0f);
			l.color[2] = power * rng.uniform(0.5f, 1.0f);
			lights.push_back(l);
		}
	}

	// kd-style BSP over the grid; a box whose cells are all solid or all empty becomes one (convex) leaf.
	bool uniform_box(const Box& b, bool& solid_out) const
	{
		const bool first = s.solid[cell_index(b.lo[0], b.lo[1], b.lo[2])] != 0;
		for (int z = b.lo[2]; z < b.hi[2]; ++z)
			for (int y = b.lo[1]; y < b.hi[1]; ++y)
				for (int x = b.lo[0]; x < b.hi[0]; ++x)
					if ((s.solid[cell_index(x, y, z)] != 0) != first)
						return false;
		solid_out = first;
		return true;
	}
	uint32_t build_bsp(const Box& b)
	{
		bool solid;
		if (uniform_box(b, solid))
		{
			const uint32_t leaf = (uint32_t)leaf_boxes.size();
			leaf_boxes.push_back(b);
			leaf_solid.push_back(solid ? 1 : 0);
			for (int z = b.lo[2]; z < b.hi[2]; ++z)
				for (int y = b.lo[1]; y < b.hi[1]; ++y)
					for (int x = b.lo[0]; x < b.hi[0]; ++x)
						cell_leaf[cell_index(x, y, z)] = leaf;
			return leaf + FIRST_LEAF_INDEX;
		}
		// split along the axis with the largest world extent
		int axis = 0;
		float best = -1;
		for (int a = 0; a < 3; ++a)
		{
			const float ext = (float)(b.hi[a] - b.lo[a]) * s.params.cell_size[a];
			if (b.hi[a] - b.lo[a] > 1 && ext > best)
			{
				best = ext;
				axis = a;
			}
		}
		const int mid = (b.lo[axis] + b.hi[axis]) / 2;
		Box lo_box = b, hi_box = b;
		lo_box.hi[axis] = mid;
		hi_box.lo[axis] = mid;
		const uint32_t front = build_bsp(hi_box); // dot(normal, p) - dist >= 0
		const uint32_t back = build_bsp(lo_box);
		BSPNode node;
		node.children[0] = front;
		node.children[1] = back;
		node.plane.vec = {axis == 0 ? 1.0f : 0.0f, axis == 1 ? 1.0f : 0.0f, axis == 2 ? 1.0f : 0.0f};
		node.plane.dist = (float)mid * s.params.cell_size[axis];
		s.map.nodes.push_back(node);
		return (uint32_t)s.map.nodes.size() - 1;
	}

	// Rectangle on the plane {pos[axis] = coord}, spanning [b0,b1] x [c0,c1] on the two other axes, facing `sign` along the axis.
	// Vertices are clockwise when seen from the front (the reference's polygon winding, rasterizer.rs:34).
	void rect_vertices(int axis, float coord, int sign, float b0, float b1, float c0, float c1, Vec3f out[4]) const
	{
		const int ba = (axis + 1) % 3, ca = (axis + 2) % 3;
		const float bc[4][2] = {{b0, c0}, {b1, c0}, {b1, c1}, {b0, c1}};
		for (int i = 0; i < 4; ++i)
		{
			const int k = sign < 0 ? i : 3 - i;
			Vec3f v;
			v.at(axis) = coord;
			v.at(ba) = bc[k][0];
			v.at(ca) = bc[k][1];
			out[i] = v;
		}
	}

	uint32_t pick_material(int axis, int sign, uint32_t leaf)
	{
		// floors, ceilings and walls draw from different parts of the material table; one material per leaf face
		const uint32_t nm = (uint32_t)s.materials.size();
		uint32_t h = leaf * 2654435761u ^ (uint32_t)(axis * 97 + sign * 31 + 1000);
		h ^= h >> 15;
		h *= 2246822519u;
		h ^= h >> 13;
		if (axis == 2)
			return sign > 0 ? (h % std::max(1u, nm / 4)) : (nm / 4 + h % std::max(1u, nm / 4)) % nm;
		return (nm / 2 + h % std::max(1u, nm - nm / 2)) % nm;
	}

	void add_wall(uint32_t leaf, int axis, float coord, int sign, float b0, float b1, float c0, float c1)
	{
		const int ba = (axis + 1) % 3, ca = (axis + 2) % 3;
		const float max_size = s.params.max_polygon_size;
		const int nb = std::max(1, (int)std::ceil((b1 - b0) / max_size - 1e-4f)), nc = std::max(1, (int)std::ceil((c1 - c0) / max_size - 1e-4f));
		const uint32_t material = pick_material(axis, sign, leaf);
		for (int ic = 0; ic < nc; ++ic)
			for (int ib = 0; ib < nb; ++ib)
			{
				const float pb0 = b0 + (b1 - b0) * (float)ib / (float)nb, pb1 = b0 + (b1 - b0) * (float)(ib + 1) / (float)nb;
				const float pc0 = c0 + (c1 - c0) * (float)ic / (float)nc, pc1 = c0 + (c1 - c0) * (float)(ic + 1) / (float)nc;
				Vec3f v[4];
				rect_vertices(axis, coord, sign, pb0, pb1, pc0, pc1, v);
				Polygon p{};
				p.first_vertex = (uint32_t)s.map.vertices.size();
				p.num_vertices = 4;
				for (auto& vv : v)
					s.map.vertices.push_back(vv);
				p.plane.vec = Vec3f{};
				p.plane.vec.at(axis) = (float)sign;
				p.plane.dist = (float)sign * coord;
				// 2 texels per unit; walls: u horizontal, v down; floors/ceilings: u = x, v = y
				Vec3f tu{}, tv{};
				if (axis == 2)
				{
					tu.x = 2.0f;
					tv.y = 2.0f;
				}
				else
				{
					tu.at(axis == 0 ? 1 : 0) = 2.0f;
					tv.z = -2.0f;
				}
				p.tex_coord_equation[0] = Plane{tu, 0.0f};
				p.tex_coord_equation[1] = Plane{tv, 0.0f};
				for (int i = 0; i < 2; ++i)
				{
					float mn = 1e30f, mx = -1e30f;
					for (auto& vv : v)
					{
						const float tc = dot(p.tex_coord_equation[i].vec, vv) + p.tex_coord_equation[i].dist;
						mn = std::fmin(mn, tc);
						mx = std::fmax(mx, tc);
					}
					p.tex_coord_min[i] = (int32_t)std::floor(mn);
					p.tex_coord_max[i] = (int32_t)std::ceil(mx);
				}
				p.texture = material;
				p.lightmap_data_offset = 0;
				s.map.polygons.push_back(p);
			}
		(void)ba;
		(void)ca;
	}

	void build_leaf_geometry()
	{
		BSPMap& m = s.map;
		m.leafs.resize(leaf_boxes.size());
		std::vector<std::vector<uint32_t>> leaf_portal_lists(leaf_boxes.size());
		std::map<std::pair<uint32_t, uint32_t>, bool> portal_done;
		for (uint32_t leaf = 0; leaf < leaf_boxes.size(); ++leaf)
		{
			BSPLeaf& L = m.leafs[leaf];
			L.first_polygon = (uint32_t)m.polygons.size();
			L.num_polygons = 0;
			L.first_leaf_portal = L.num_leaf_portals = 0;
			if (leaf_solid[leaf])
				continue;
			const Box& b = leaf_boxes[leaf];
			for (int axis = 0; axis < 3; ++axis)
				for (int side = 0; side < 2; ++side)
				{
					const int ba = (axis + 1) % 3, ca = (axis + 2) % 3;
					const int outside = side == 0 ? b.lo[axis] - 1 : b.hi[axis];
					const float coord = (float)(side == 0 ? b.lo[axis] : b.hi[axis]) * s.params.cell_size[axis];
					const int sign = side == 0 ? 1 : -1; // wall normal points into the leaf
					// group the face cells by the leaf on the other side
					std::map<uint32_t, Box> groups; // neighbour leaf (or ~0 for outside the grid) -> cell rectangle
					for (int c = b.lo[ca]; c < b.hi[ca]; ++c)
						for (int bb = b.lo[ba]; bb < b.hi[ba]; ++bb)
						{
							int cell[3];
							cell[axis] = outside;
							cell[ba] = bb;
							cell[ca] = c;
							uint32_t nb = ~0u;
							if (outside >= 0 && outside < n[axis])
								nb = cell_leaf[cell_index(cell[0], cell[1], cell[2])];
							auto it = groups.find(nb);
							if (it == groups.end())
							{
								Box g;
								g.lo[0] = bb;
								g.hi[0] = bb + 1;
								g.lo[1] = c;
								g.hi[1] = c + 1;
								g.lo[2] = g.hi[2] = 0;
								groups[nb] = g;
							}
							else
							{
								it->second.lo[0] = std::min(it->second.lo[0], bb);
								it->second.hi[0] = std::max(it->second.hi[0], bb + 1);
								it->second.lo[1] = std::min(it->second.lo[1], c);
								it->second.hi[1] = std::max(it->second.hi[1], c + 1);
							}
						}
					for (auto& [nb, g] : groups)
					{
						const float b0 = (float)g.lo[0] * s.params.cell_size[ba], b1 = (float)g.hi[0] * s.params.cell_size[ba];
						const float c0 = (float)g.lo[1] * s.params.cell_size[ca], c1 = (float)g.hi[1] * s.params.cell_size[ca];
						if (nb == ~0u || leaf_solid[nb])
							add_wall(leaf, axis, coord, sign, b0, b1, c0, c1);
						else
						{
							const auto key = std::make_pair(std::min(leaf, nb), std::max(leaf, nb));
							// two boxes share at most one face rectangle
							if (portal_done.count(key))
								continue;
							portal_done[key] = true;
							Portal p{};
							p.leafs[0] = leaf;
							p.leafs[1] = nb;
							// looking from leafs[0] through the portal requires the camera on the positive side
							// (map_visibility_calculator.rs:193-211): the normal points back into leafs[0]
							p.plane.vec = Vec3f{};
							p.plane.vec.at(axis) = (float)sign;
							p.plane.dist = (float)sign * coord;
							Vec3f v[4];
							rect_vertices(axis, coord, sign, b0, b1, c0, c1, v);
							p.first_vertex = (uint32_t)m.vertices.size();
							p.num_vertices = 4;
							for (auto& vv : v)
								m.vertices.push_back(vv);
							const uint32_t pi = (uint32_t)m.portals.size();
							m.portals.push_back(p);
							leaf_portal_lists[leaf].push_back(pi);
							leaf_portal_lists[nb].push_back(pi);
						}
					}
				}
			L.num_polygons = (uint32_t)m.polygons.size() - L.first_polygon;
		}
		for (uint32_t leaf = 0; leaf < leaf_boxes.size(); ++leaf)
		{
			m.leafs[leaf].first_leaf_portal = (uint32_t)m.leafs_portals.size();
			m.leafs[leaf].num_leaf_portals = (uint32_t)leaf_portal_lists[leaf].size();
			for (uint32_t p : leaf_portal_lists[leaf])
				m.leafs_portals.push_back(p);
		}
	}

	bool occluded(Vec3f from, Vec3f to) const
	{
		const Vec3f d = to - from;
		const float len = magnitude(d);
		const float step = 0.45f * std::fmin(s.params.cell_size[0], std::fmin(s.params.cell_size[1], s.params.cell_size[2]));
		const int steps = (int)(len / step);
		for (int i = 1; i < steps; ++i)
		{
			const Vec3f p = from + d * ((float)i / (float)steps);
			if (s.is_solid((int)std::floor(p.x / s.params.cell_size[0]), (int)std::floor(p.y / s.params.cell_size[1]), (int)std::floor(p.z / s.params.cell_size[2])))
				return true;
		}
		return false;
	}

	// Direct-light stand-in for lightmaps_builder.rs: fills both lightmap arrays (same offsets, allocate_lightmaps :299-330).
	void build_lightmaps()
	{
		BSPMap& m = s.map;
		m.lightmaps_data.assign(3, 0.0f); // element 0 is reserved: offset 0 means "no lightmap"
		m.directional_lightmaps_data.assign(1, sqw_directional_lightmap_element{});
		for (Polygon& p : m.polygons)
		{
			const uint32_t lw = get_lightmap_size(p.tex_coord_min[0], p.tex_coord_max[0]);
			const uint32_t lh = get_lightmap_size(p.tex_coord_min[1], p.tex_coord_max[1]);
			p.lightmap_data_offset = (uint32_t)m.directional_lightmaps_data.size();
			const Vec3f tu = p.tex_coord_equation[0].vec, tv = p.tex_coord_equation[1].vec;
			const Vec3f u_dir = tu / dot(tu, tu), v_dir = tv / dot(tv, tv); // world units per texel along u / v
			const Vec3f nrm = p.plane.vec;
			const Vec3f un = normalize(tu), vn = normalize(tv);
			const Vec3f origin = nrm * p.plane.dist; // tc (0,0) is at the plane point closest to the world origin for these equations
			for (uint32_t j = 0; j < lh; ++j)
				for (uint32_t i = 0; i < lw; ++i)
				{
					const float tcu = (float)(((p.tex_coord_min[0] >> LIGHTMAP_SCALE_LOG2) + (int32_t)i) * (int32_t)LIGHTMAP_SCALE);
					const float tcv = (float)(((p.tex_coord_min[1] >> LIGHTMAP_SCALE_LOG2) + (int32_t)j) * (int32_t)LIGHTMAP_SCALE);
					const Vec3f pos = origin + u_dir * tcu + v_dir * tcv + nrm * 2.0f;
					float total[3] = {0, 0, 0};
					Vec3f dir_sum{};
					float weight_sum = 0;
					for (const MapLight& l : lights)
					{
						const Vec3f to_light = l.pos - pos;
						const float d2 = std::fmax(dot(to_light, to_light), 64.0f);
						if (d2 > 2200.0f * 2200.0f)
							continue;
						const float d = std::sqrt(d2);
						const float cosine = dot(to_light, nrm) / d;
						if (cosine <= 0.0f)
							continue;
						if (occluded(pos, l.pos))
							continue;
						const float k = cosine / d2;
						const float c[3] = {l.color[0] * k, l.color[1] * k, l.color[2] * k};
						const float w = c[0] + c[1] + c[2];
						for (int ch = 0; ch < 3; ++ch)
							total[ch] += c[ch];
						const Vec3f dir_ts{dot(to_light, un) / d, dot(to_light, vn) / d, cosine};
						dir_sum = dir_sum + dir_ts * w;
						weight_sum += w;
					}
					const float ambient = 0.06f;
					sqw_directional_lightmap_element e{};
					Vec3f mean{0, 0, 1.0f / 1024.0f}; // zero direction is patched like lightmap_compression.rs:108-124
					if (weight_sum > 0)
						mean = dir_sum / weight_sum;
					const float mz = std::fmax(mean.z, 0.05f);
					for (int ch = 0; ch < 3; ++ch)
					{
						e.ambient_light[ch] = ambient + 0.12f * total[ch];
						e.directional_light_color[ch] = 0.88f * total[ch] / mz;
					}
					e.light_direction_vector_scaled[0] = mean.x;
					e.light_direction_vector_scaled[1] = mean.y;
					e.light_direction_vector_scaled[2] = mean.z;
					e.directional_light_deviation = std::fmin(std::fmax(1.0f - magnitude(mean), 0.0f), 1.0f);
					m.directional_lightmaps_data.push_back(e);
					for (int ch = 0; ch < 3; ++ch)
						m.lightmaps_data.push_back(ambient + total[ch]);
				}
		}
	}

	void build_materials()
	{
		const uint32_t size = 1u << s.params.texture_size_log2;
		s.materials.resize(s.params.num_materials);
		for (uint32_t mi = 0; mi < s.params.num_materials; ++mi)
		{
			Material& mat = s.materials[mi];
			if (s.params.material_mode == 0 || s.params.material_mode == 2)
			{
				// demo_materials.json: everything has a normal map; roughness < 1 for most; a quarter are metals
				mat.has_normal_map = true;
				mat.has_non_one_roughness = (mi % 4) != 3;
				mat.is_metal = (mi % 4) == 1;
			}
			int alpha_mode = 0;
			if (s.params.material_mode == 2)
			{
				// BASELINE configs[3]: a third of the world materials are translucent or alpha-tested like alpha_blend / alpha_grid /
				// stained_glass of demo_materials.json (material.rs:101-116); they neither cast shadows nor hide what lies behind
				switch (mi % 12)
				{
				case 2: mat.blending_mode = SQW_BLEND_ALPHA_BLEND; alpha_mode = 1; break;
				case 5: mat.blending_mode = SQW_BLEND_ALPHA_TEST; alpha_mode = 2; break;
				case 8: mat.blending_mode = SQW_BLEND_AVERAGE; break;
				case 11: mat.blending_mode = SQW_BLEND_ADDITIVE; break;
				default: break;
				}
			}
			const float base_roughness = mat.has_non_one_roughness ? rng.uniform(0.15f, 0.7f) : 1.0f;
			HostImage mips[NUM_MIPS];
			make_world_texture(rng, size, mat.has_normal_map, mat.has_non_one_roughness, base_roughness, mips[0], alpha_mode);
			for (uint32_t k = 1; k < NUM_MIPS; ++k)
				build_texture_mip(mips[k], mips[k - 1]);
			for (uint32_t k = 0; k < NUM_MIPS; ++k)
			{
				mat.texture_images[k] = (uint32_t)s.images.size();
				s.images.push_back(std::move(mips[k]));
			}
			if (rng.uniform() < s.params.emissive_fraction)
			{
				mat.has_emissive = true;
				const float power = rng.uniform(2.0f, 40.0f); // demo_materials.json glowing_button: up to 80
				for (int c = 0; c < 3; ++c)
					mat.emissive_light[c] = power * rng.uniform(0.4f, 1.0f);
				HostImage em[NUM_MIPS];
				make_emissive_texture(rng, size, em[0]);
				for (uint32_t k = 1; k < NUM_MIPS; ++k)
					build_texture_lite_mip(em[k], em[k - 1]);
				for (uint32_t k = 0; k < NUM_MIPS; ++k)
				{
					mat.emissive_images[k] = (uint32_t)s.images.size();
					s.images.push_back(std::move(em[k]));
				}
			}
		}
	}

	void run()
	{
		build_materials();
		carve_layout();
		cell_leaf.assign((size_t)n[0] * n[1] * n[2], 0);
		Box all;
		for (int a = 0; a < 3; ++a)
		{
			all.lo[a] = 0;
			all.hi[a] = n[a];
		}
		build_bsp(all);
		build_leaf_geometry();
		build_submodels();
		build_lightmaps();
	}

	// Inline models (bsp_map_compact.rs:87-107): axis-aligned crates standing beside the fly-through path, their polygons
	// appended after all leaf polygons, one SubmodelBSPNode each (a convex solid needs no ordering between its faces).
	void build_submodels()
	{
		BSPMap& m = s.map;
		const size_t nw = s.camera_waypoints.size();
		for (uint32_t k = 0; k < s.params.num_submodels && nw > 0; ++k)
		{
			const Vec3f c = s.camera_waypoints[(size_t)((uint64_t)k * nw / s.params.num_submodels) % nw];
			const float half[3] = {rng.uniform(14.0f, 26.0f), rng.uniform(14.0f, 26.0f), rng.uniform(18.0f, 34.0f)};
			const float side = (k & 1) ? 1.0f : -1.0f;
			const Vec3f center = {c.x + side * 0.28f * s.params.cell_size[0], c.y - side * 0.22f * s.params.cell_size[1], c.z - 0.5f * s.params.cell_size[2] + half[2] + 6.0f};
			const float lo[3] = {center.x - half[0], center.y - half[1], center.z - half[2]}, hi[3] = {center.x + half[0], center.y + half[1], center.z + half[2]};
			Submodel sm;
			sm.first_polygon = (uint32_t)m.polygons.size();
			for (int axis = 0; axis < 3; ++axis)
			{
				const int ba = (axis + 1) % 3, ca = (axis + 2) % 3;
				add_wall(1000003u + k, axis, hi[axis], +1, lo[ba], hi[ba], lo[ca], hi[ca]);
				add_wall(1000003u + k, axis, lo[axis], -1, lo[ba], hi[ba], lo[ca], hi[ca]);
			}
			sm.num_polygons = (uint32_t)m.polygons.size() - sm.first_polygon;
			sm.root_node = (uint32_t)m.submodels_bsp_nodes.size();
			SubmodelBSPNode node;
			node.plane = m.polygons[sm.first_polygon].plane;
			node.first_polygon = sm.first_polygon;
			node.num_polygons = sm.num_polygons;
			node.children[0] = node.children[1] = 0xFFFFFFFFu; // >= number of nodes: no child
			m.submodels_bsp_nodes.push_back(node);
			m.submodels.push_back(sm);
		}
	}
};

} // namespace

namespace
{
// Model skin (TextureLite): coloured panels with seams, every fourth skin with transparent holes (alpha 0) for the alpha-test
// and alpha-blend modes.  64 x 64 with 4 more mips (TextureLiteWithMips, textures.rs:367-386).
void make_model_skin(Pcg32& rng, uint32_t size, bool with_holes, HostImage& out)
{
	out.kind = SQW_IMAGE_COLOR32;
	out.width = out.height = size;
	out.pixels.resize((size_t)size * size * 4);
	uint32_t* px = reinterpret_cast<uint32_t*>(out.pixels.data());
	const uint32_t cells = 8;
	std::vector<float> lat(cells * cells);
	for (float& v : lat)
		v = rng.uniform();
	const float base[3] = {rng.uniform(0.35f, 1.0f), rng.uniform(0.35f, 1.0f), rng.uniform(0.35f, 1.0f)};
	for (uint32_t y = 0; y < size; ++y)
		for (uint32_t x = 0; x < size; ++x)
		{
			const float n = value_noise(rng, lat, cells, ((float)x + 0.5f) / (float)size, ((float)y + 0.5f) / (float)size);
			const bool seam = (x % (size / 4)) == 0 || (y % (size / 4)) == 0;
			const float shade = seam ? 0.35f : 0.6f + 0.4f * n;
			uint32_t rgb[3];
			for (int c = 0; c < 3; ++c)
				rgb[c] = (uint32_t)(base[c] * shade * 255.0f);
			uint32_t alpha = 255;
			if (with_holes)
				alpha = n > 0.6f ? 0u : (n > 0.5f ? (uint32_t)(255.0f * (0.6f - n) * 10.0f) : 255u);
			px[x + (size_t)y * size] = (alpha << 24) | (rgb[0] << 16) | (rgb[1] << 8) | rgb[2];
		}
}
} // namespace

namespace
{
// One side of the sky box in both colour widths (SkyboxSideTexture<ColorT>, textures.rs:82-87): a vertical gradient with
// value-noise clouds and, on side +X, a sun disc far above LDR white in the 64-bit version (HDR unit: 255 = white).
void make_sky_side(Pcg32& rng, uint32_t size, uint32_t side, HostImage& out32, HostImage& out64)
{
	out32.kind = SQW_IMAGE_COLOR32;
	out64.kind = SQW_IMAGE_COLOR64;
	out32.width = out32.height = out64.width = out64.height = size;
	out32.pixels.resize((size_t)size * size * 4);
	out64.pixels.resize((size_t)size * size * 8);
	uint32_t* p32 = reinterpret_cast<uint32_t*>(out32.pixels.data());
	uint64_t* p64 = reinterpret_cast<uint64_t*>(out64.pixels.data());
	const uint32_t cells = 8;
	std::vector<float> lat(cells * cells);
	for (float& v : lat)
		v = rng.uniform();
	for (uint32_t y = 0; y < size; ++y)
		for (uint32_t x = 0; x < size; ++x)
		{
			const float u = ((float)x + 0.5f) / (float)size, v = ((float)y + 0.5f) / (float)size;
			const float cloud = value_noise(rng, lat, cells, u, v);
			const float up = side == 5 ? 1.0f : (side == 4 ? 0.0f : 1.0f - v); // sides: v runs downwards (-Z basis of draw_skybox)
			float rgb[3] = {0.25f + 0.35f * (1.0f - up), 0.45f + 0.3f * (1.0f - up), 0.95f};
			const float c = cloud > 0.55f ? (cloud - 0.55f) / 0.45f : 0.0f;
			for (float& ch : rgb)
				ch = ch * (1.0f - c) + c;
			float hdr_scale = 1.6f;
			if (side == 1)
			{
				const float dx = u - 0.6f, dy = v - 0.3f;
				if (dx * dx + dy * dy < 0.004f)
				{
					rgb[0] = rgb[1] = 1.0f;
					rgb[2] = 0.85f;
					hdr_scale = 24.0f;
				}
			}
			uint32_t c8[3];
			uint64_t c16[3];
			for (int i = 0; i < 3; ++i)
			{
				c8[i] = (uint32_t)std::fmin(rgb[i] * 255.0f, 255.0f);
				c16[i] = (uint64_t)std::fmin(rgb[i] * 255.0f * hdr_scale, 65535.0f);
			}
			p32[x + (size_t)y * size] = 0xFF000000u | (c8[0] << 16) | (c8[1] << 8) | c8[2];
			p64[x + (size_t)y * size] = (c16[0] << 32) | (c16[1] << 16) | c16[2];
		}
}

// 2x2 box filter of a Color64 image (the 64-bit twin of build_texture_lite_mip, textures.rs:388-421: channel average)
void build_color64_mip(HostImage& mip, const HostImage& prev)
{
	mip.kind = SQW_IMAGE_COLOR64;
	mip.width = prev.width >> 1;
	mip.height = prev.height >> 1;
	mip.pixels.resize((size_t)mip.width * mip.height * 8);
	const uint64_t* src = reinterpret_cast<const uint64_t*>(prev.pixels.data());
	uint64_t* dst = reinterpret_cast<uint64_t*>(mip.pixels.data());
	for (uint32_t y = 0; y < mip.height; ++y)
		for (uint32_t x = 0; x < mip.width; ++x)
		{
			uint64_t r = 0;
			for (int ch = 0; ch < 4; ++ch)
			{
				uint64_t sum = 0;
				for (int dy = 0; dy < 2; ++dy)
					for (int dx = 0; dx < 2; ++dx)
						sum += (src[x * 2 + dx + (size_t)(y * 2 + dy) * prev.width] >> (16 * ch)) & 0xFFFF;
				r |= ((sum >> 2) & 0xFFFF) << (16 * ch);
			}
			dst[x + (size_t)y * mip.width] = r;
		}
}
} // namespace

void generate_scene(const SceneParams& params, Scene& out)
{
	out = Scene{};
	out.params = params;
	Builder b(out);
	b.run();
	// material names of the file format's texture table (bsp_map_compact.rs:137-139)
	out.map.textures.resize(out.materials.size());
	for (size_t i = 0; i < out.materials.size(); ++i)
	{
		std::memset(out.map.textures[i].name, 0, MAX_TEXTURE_NAME_LEN);
		std::snprintf(out.map.textures[i].name, MAX_TEXTURE_NAME_LEN, "synthetic/material_%02u", (unsigned)i);
	}
	if (params.light_grid)
	{
		// Light grid fixture (lightmaps_builder.rs:299-330 layout): 64-unit cells over the whole grid volume, one column per
		// (x, y) holding the samples of the non-solid z range; sample = a smooth seeded light cube + one directional component.
		// Samples inside solid cells are the all-zero element, which fetch_light_from_grid skips (:644-647).
		BSPMap& m = out.map;
		const float cell = 64.0f;
		const float extent[3] = {params.grid[0] * params.cell_size[0], params.grid[1] * params.cell_size[1], params.grid[2] * params.cell_size[2]};
		for (int a = 0; a < 3; ++a)
		{
			m.light_grid_header.grid_cell_size[a] = cell;
			m.light_grid_header.grid_start[a] = 0.5f * cell;
			m.light_grid_header.grid_size[a] = std::max(1u, (uint32_t)(extent[a] / cell));
		}
		const uint32_t* gs = m.light_grid_header.grid_size;
		m.light_grid_columns.resize((size_t)gs[0] * gs[1]);
		for (uint32_t y = 0; y < gs[1]; ++y)
			for (uint32_t x = 0; x < gs[0]; ++x)
			{
				LightGridColumn& col = m.light_grid_columns[x + (size_t)y * gs[0]];
				col.start_z = 0;
				col.first_sample = (uint32_t)m.light_grid_samples.size();
				col.num_samples = gs[2];
				for (uint32_t z = 0; z < gs[2]; ++z)
				{
					const Vec3f pos = {m.light_grid_header.grid_start[0] + x * cell, m.light_grid_header.grid_start[1] + y * cell, m.light_grid_header.grid_start[2] + z * cell};
					LightGridElement e;
					const int cx = (int)(pos.x / params.cell_size[0]), cy = (int)(pos.y / params.cell_size[1]), cz = (int)(pos.z / params.cell_size[2]);
					if (!out.is_solid(cx, cy, cz))
					{
						const float w = 0.5f + 0.5f * std::sin(pos.x * 0.011f) * std::cos(pos.y * 0.013f);
						for (int side = 0; side < 6; ++side)
						{
							const float up = side == 5 ? 1.6f : (side == 4 ? 0.5f : 1.0f); // brighter from above
							e.light_cube[side][0] = 0.35f * up * (0.6f + 0.4f * w);
							e.light_cube[side][1] = 0.32f * up;
							e.light_cube[side][2] = 0.30f * up * (1.0f - 0.3f * w);
						}
						e.light_direction_vector_scaled = Vec3f{std::cos(pos.x * 0.02f), std::sin(pos.y * 0.02f), 0.7f} * 0.8f;
						e.directional_light_color[0] = 0.9f;
						e.directional_light_color[1] = 0.8f * (0.5f + 0.5f * w);
						e.directional_light_color[2] = 0.6f;
					}
					m.light_grid_samples.push_back(e);
				}
			}
	}
	if (params.skybox && out.materials.size() >= 4)
	{
		// appended after the world images, so scenes without a sky keep their image ids
		Material& sky = out.materials[out.materials.size() / 4]; // the first ceiling material (Builder::pick_material)
		sky.is_skybox = true;
		Pcg32 sky_rng(params.seed * 40503u + 7u);
		const uint32_t size = 1u << std::min(std::max(params.sky_size_log2, 5u), 10u);
		for (uint32_t side = 0; side < 6; ++side)
		{
			HostImage m32[NUM_MIPS], m64[NUM_MIPS];
			make_sky_side(sky_rng, size, side, m32[0], m64[0]);
			for (uint32_t k = 1; k < NUM_MIPS; ++k)
			{
				build_texture_lite_mip(m32[k], m32[k - 1]);
				build_color64_mip(m64[k], m64[k - 1]);
			}
			for (uint32_t k = 0; k < NUM_MIPS; ++k)
			{
				sky.skybox_images[0][side][k] = (uint32_t)out.images.size();
				out.images.push_back(std::move(m32[k]));
				sky.skybox_images[1][side][k] = (uint32_t)out.images.size();
				out.images.push_back(std::move(m64[k]));
			}
		}
	}
	// triangle-model fixtures and their skins (appended after the world images, so scenes without models keep their image ids)
	Pcg32 rng(params.seed * 2654435761u + 99u);
	for (uint32_t i = 0; i < params.num_models; ++i)
	{
		TriangleModel model;
		generate_triangle_model(i % 3, params.seed + i, params.model_detail, model);
		out.models.push_back(std::move(model));
		HostImage mips[NUM_MIPS];
		make_model_skin(rng, 64, (i % 4) == 3, mips[0]);
		for (uint32_t k = 1; k < NUM_MIPS; ++k)
			build_texture_lite_mip(mips[k], mips[k - 1]);
		std::array<uint32_t, NUM_MIPS> ids;
		for (uint32_t k = 0; k < NUM_MIPS; ++k)
		{
			ids[k] = (uint32_t)out.images.size();
			out.images.push_back(std::move(mips[k]));
		}
		out.model_skins.push_back(ids);
	}
}

// ---- a compiled map (map_compiler.cpp + lightmapper.cpp, or the reference's own tools) as a Scene ------------------------------
namespace
{
uint32_t leaf_of_point(const BSPMap& map, Vec3f p) // bsp_map_compact.rs get_leaf_for_point
{
	uint32_t index = map.root_node();
	while (index < FIRST_LEAF_INDEX)
	{
		const BSPNode& n = map.nodes[index];
		index = dot(n.plane.vec, p) >= n.plane.dist ? n.children[0] : n.children[1];
	}
	return index - FIRST_LEAF_INDEX;
}
Vec3f portal_center(const BSPMap& map, const Portal& p)
{
	Vec3f c{0, 0, 0};
	for (uint32_t i = 0; i < p.num_vertices; ++i)
		c = c + map.vertices[p.first_vertex + i];
	return c / (float)p.num_vertices;
}
} // namespace

void bind_compiled_map(Scene& s, BSPMap&& map, const mapc::MaterialTable& table, uint32_t texture_size_log2, uint32_t seed)
{
	s.map = std::move(map);
	s.materials.clear();
	s.images.clear();
	s.models.clear();
	s.model_skins.clear();
	s.params.seed = seed;
	s.params.texture_size_log2 = texture_size_log2;
	s.params.num_materials = (uint32_t)s.map.textures.size();
	const uint32_t size = 1u << texture_size_log2;
	// Materials by texture name (map_materials_processor.rs:29-120 reads the same table).  The demo's texture IMAGES are not
	// shipped (procedural .ptex sources), so every material gets a seeded texture of the kind its definition asks for.
	for (size_t ti = 0; ti < s.map.textures.size(); ++ti)
	{
		const mapc::MaterialDef* def = mapc::material_of_texture(table, s.map.textures[ti]);
		Material mat;
		Pcg32 rng(seed * 7919u + (uint32_t)ti * 104729u + 13u);
		float base_roughness = 1.0f;
		int alpha_mode = 0;
		if (def)
		{
			mat.draw = def->draw;
			mat.light = def->light;
			mat.decals = def->decals;
			mat.shadow = def->shadow;
			mat.blending_mode = def->blending_mode;
			mat.has_normal_map = def->normal_map;
			mat.has_non_one_roughness = def->roughness_map || def->roughness < 1.0f;
			mat.is_metal = def->is_metal;
			base_roughness = def->roughness_map ? 0.5f : std::fmin(std::fmax(def->roughness, 1.0f / 126.0f), 1.0f);
			alpha_mode = def->blending_mode == SQW_BLEND_ALPHA_BLEND ? 1 : (def->blending_mode == SQW_BLEND_ALPHA_TEST ? 2 : 0);
			mat.is_skybox = def->skybox;
		}
		HostImage mips[NUM_MIPS];
		make_world_texture(rng, size, mat.has_normal_map, mat.has_non_one_roughness, base_roughness, mips[0], alpha_mode);
		for (uint32_t k = 1; k < NUM_MIPS; ++k)
			build_texture_mip(mips[k], mips[k - 1]);
		for (uint32_t k = 0; k < NUM_MIPS; ++k)
		{
			mat.texture_images[k] = (uint32_t)s.images.size();
			s.images.push_back(std::move(mips[k]));
		}
		if (def && def->emissive_layer) // material.rs EmissiveLayer: a TextureLite mixed in with the layer's light (surfaces.rs:854)
		{
			mat.has_emissive = true;
			for (int c = 0; c < 3; ++c)
				mat.emissive_light[c] = def->emissive_layer_light[c];
			HostImage em[NUM_MIPS];
			make_emissive_texture(rng, size, em[0]);
			for (uint32_t k = 1; k < NUM_MIPS; ++k)
				build_texture_lite_mip(em[k], em[k - 1]);
			for (uint32_t k = 0; k < NUM_MIPS; ++k)
			{
				mat.emissive_images[k] = (uint32_t)s.images.size();
				s.images.push_back(std::move(em[k]));
			}
		}
		s.materials.push_back(mat);
	}
	for (Material& sky : s.materials)
	{
		if (!sky.is_skybox)
			continue;
		Pcg32 sky_rng(seed * 40503u + 7u);
		const uint32_t sky_size = 1u << std::min(std::max(s.params.sky_size_log2, 5u), 10u);
		for (uint32_t side = 0; side < 6; ++side)
		{
			HostImage m32[NUM_MIPS], m64[NUM_MIPS];
			make_sky_side(sky_rng, sky_size, side, m32[0], m64[0]);
			for (uint32_t k = 1; k < NUM_MIPS; ++k)
			{
				build_texture_lite_mip(m32[k], m32[k - 1]);
				build_color64_mip(m64[k], m64[k - 1]);
			}
			for (uint32_t k = 0; k < NUM_MIPS; ++k)
			{
				sky.skybox_images[0][side][k] = (uint32_t)s.images.size();
				s.images.push_back(std::move(m32[k]));
				sky.skybox_images[1][side][k] = (uint32_t)s.images.size();
				s.images.push_back(std::move(m64[k]));
			}
		}
	}

	// Camera path: from info_player_start through the info_teleport_destination origins, in entity order.
	// Between two of them the path follows the leaf graph (breadth-first over portals) through the portal centres — inside a
	// convex leaf the straight piece between two points of its boundary stays in empty space; where the graph does not connect
	// them (areas joined only by a teleporter) the path jumps.
	s.camera_waypoints.clear();
	s.waypoint_breaks.clear();
	s.polyline_path = true;
	std::vector<Vec3f> stops;
	{
		const BSPMap& m = s.map;
		std::vector<Vec3f> destinations, room_points;
		for (const Entity& e : m.entities)
		{
			std::string classname;
			Vec3f origin{0, 0, 0};
			bool has_origin = false;
			for (uint32_t i = 0; i < e.num_key_value_pairs; ++i)
			{
				const KeyValuePair& kv = m.key_value_pairs[e.first_key_value_pair + i];
				const std::string key((const char*)m.strings_data.data() + kv.key.offset, kv.key.size), value((const char*)m.strings_data.data() + kv.value.offset, kv.value.size);
				if (key == "classname")
					classname = value;
				if (key == "origin")
					has_origin = mapc::parse_vec3(value, origin);
			}
			if (!has_origin)
				continue;
			if (classname == "light") // free-standing point lights hang in the open space of the rooms
				room_points.push_back(origin);
			origin.z += 40.0f; // eye above the entity's origin
			if (classname == "info_player_start")
				stops.insert(stops.begin(), origin);
			else if (classname == "info_teleport_destination")
				destinations.push_back(origin);
		}
		// A tour of the part of the map the player start lies in (the demo's teleport destinations are rooms of their own that
		// the leaf graph does not connect to it): after the start, ten times the room — a free-standing `light` entity stands for
		// one — that is farthest, in portal hops, from every stop chosen so far.  Deterministic: ties go to the earlier entity.
		if (!stops.empty() && !m.leafs.empty())
		{
			const size_t nl = m.leafs.size();
			std::vector<uint32_t> dist(nl, ~0u);
			auto relax_from = [&](uint32_t source) {
				std::vector<uint32_t> queue{source};
				dist[source] = 0;
				for (size_t qi = 0; qi < queue.size(); ++qi)
				{
					const BSPLeaf& leaf = m.leafs[queue[qi]];
					for (uint32_t k = 0; k < leaf.num_leaf_portals; ++k)
					{
						const Portal& portal = m.portals[m.leafs_portals[leaf.first_leaf_portal + k]];
						const uint32_t other = portal.leafs[0] == queue[qi] ? portal.leafs[1] : portal.leafs[0];
						if (dist[other] > dist[queue[qi]] + 1)
						{
							dist[other] = dist[queue[qi]] + 1;
							queue.push_back(other);
						}
					}
				}
			};
			relax_from(leaf_of_point(m, stops[0]));
			std::vector<uint32_t> room_leafs;
			for (const Vec3f& p : room_points)
				room_leafs.push_back(leaf_of_point(m, p));
			std::vector<Vec3f> tour;
			for (int n = 0; n < 10; ++n)
			{
				size_t best = (size_t)-1;
				for (size_t r = 0; r < room_points.size(); ++r)
					if (dist[room_leafs[r]] != ~0u && dist[room_leafs[r]] > 0 && (best == (size_t)-1 || dist[room_leafs[r]] > dist[room_leafs[best]]))
						best = r;
				if (best == (size_t)-1)
					break;
				tour.push_back(room_points[best]);
				relax_from(room_leafs[best]);
			}
			stops.insert(stops.end(), tour.begin(), tour.end());
		}
		stops.insert(stops.end(), destinations.begin(), destinations.end());
	}
	const BSPMap& m = s.map;
	for (size_t si = 0; si < stops.size(); ++si)
	{
		if (si == 0)
		{
			s.camera_waypoints.push_back(stops[0]);
			s.waypoint_breaks.push_back(1);
			continue;
		}
		const uint32_t from = leaf_of_point(m, stops[si - 1]), to = leaf_of_point(m, stops[si]);
		std::vector<int32_t> via(m.leafs.size(), -1); // portal through which the leaf was reached
		std::vector<uint32_t> prev(m.leafs.size(), ~0u), queue{from};
		std::vector<uint8_t> seen(m.leafs.size(), 0);
		seen[from] = 1;
		for (size_t qi = 0; qi < queue.size() && !seen[to]; ++qi)
		{
			const BSPLeaf& leaf = m.leafs[queue[qi]];
			for (uint32_t k = 0; k < leaf.num_leaf_portals; ++k)
			{
				const uint32_t pi = m.leafs_portals[leaf.first_leaf_portal + k];
				const Portal& portal = m.portals[pi];
				const uint32_t other = portal.leafs[0] == queue[qi] ? portal.leafs[1] : portal.leafs[0];
				if (seen[other])
					continue;
				seen[other] = 1;
				via[other] = (int32_t)pi;
				prev[other] = queue[qi];
				queue.push_back(other);
			}
		}
		if (!seen[to])
		{
			s.camera_waypoints.push_back(stops[si]);
			s.waypoint_breaks.push_back(1);
			continue;
		}
		std::vector<Vec3f> centres;
		for (uint32_t l = to; l != from; l = prev[l])
			centres.push_back(portal_center(m, m.portals[(size_t)via[l]]));
		for (size_t k = centres.size(); k-- > 0;)
		{
			s.camera_waypoints.push_back(centres[k]);
			s.waypoint_breaks.push_back(0);
		}
		s.camera_waypoints.push_back(stops[si]);
		s.waypoint_breaks.push_back(0);
	}
}

namespace
{
// make_camera_path for Scene::polyline_path: poses at equal steps of arc length along the poly-line, looking at the point 160
// units further along it (never across a break)
void make_polyline_camera_path(const Scene& scene, uint32_t num_frames, uint32_t jitter_seed, std::vector<CameraPose>& out)
{
	const std::vector<Vec3f>& w = scene.camera_waypoints;
	const size_t n = w.size();
	std::vector<float> start(n, 0.0f); // arc length at waypoint i (breaks add no length)
	for (size_t i = 1; i < n; ++i)
		start[i] = start[i - 1] + (scene.waypoint_breaks[i] ? 0.0f : magnitude(w[i] - w[i - 1]));
	const float total = start[n - 1];
	Pcg32 rng(jitter_seed ? jitter_seed : 1, 0x5157);
	auto at = [&](float t, size_t& seg) {
		// last segment whose start is <= t
		size_t i = (size_t)(std::upper_bound(start.begin(), start.end(), t) - start.begin());
		i = i == 0 ? 0 : i - 1;
		while (i + 1 < n && scene.waypoint_breaks[i + 1] && start[i + 1] <= t)
			++i;
		seg = i;
		if (i + 1 >= n || scene.waypoint_breaks[i + 1])
			return w[i];
		const float len = start[i + 1] - start[i];
		const float f = len > 0.0f ? std::fmin(std::fmax((t - start[i]) / len, 0.0f), 1.0f) : 0.0f;
		return w[i] + (w[i + 1] - w[i]) * f;
	};
	for (uint32_t k = 0; k < num_frames; ++k)
	{
		const float t = total * ((float)k + 0.5f) / (float)num_frames;
		size_t seg, seg_ahead;
		CameraPose pose;
		pose.position = at(t, seg);
		Vec3f d{0, 1, 0};
		for (float ahead : {160.0f, 64.0f, 16.0f, -64.0f})
		{
			const Vec3f target = at(std::fmin(std::fmax(t + ahead, 0.0f), total), seg_ahead);
			bool crosses_break = false;
			for (size_t i = std::min(seg, seg_ahead) + 1; i <= std::max(seg, seg_ahead); ++i)
				crosses_break = crosses_break || scene.waypoint_breaks[i];
			const Vec3f dd = ahead > 0.0f ? target - pose.position : pose.position - target;
			if (!crosses_break && magnitude(dd) > 1.0f)
			{
				d = dd;
				break;
			}
		}
		pose.azimuth = std::atan2(d.y, d.x) - 3.14159265358979323846f * 0.5f; // azimuth 0 looks along +Y (matrix.rs:22)
		pose.elevation = std::atan2(d.z, std::fmax(std::sqrt(d.x * d.x + d.y * d.y), 1.0f)) * 0.5f;
		if (jitter_seed)
		{
			pose.position.x += rng.uniform(-6.0f, 6.0f);
			pose.position.y += rng.uniform(-6.0f, 6.0f);
			pose.position.z += rng.uniform(-4.0f, 4.0f);
			pose.azimuth += rng.uniform(-0.25f, 0.25f);
			pose.elevation += rng.uniform(-0.15f, 0.15f);
		}
		else
		{
			pose.azimuth += 0.5f * std::sin((float)k * 0.37f);
			pose.elevation += 0.25f * std::sin((float)k * 0.23f + 1.0f);
		}
		out.push_back(pose);
	}
}
} // namespace

void make_camera_path(const Scene& scene, uint32_t num_frames, uint32_t jitter_seed, std::vector<CameraPose>& out)
{
	out.clear();
	const std::vector<Vec3f>& w = scene.camera_waypoints;
	const size_t n = w.size();
	if (n == 0 || num_frames == 0)
		return;
	if (scene.polyline_path)
	{
		make_polyline_camera_path(scene, num_frames, jitter_seed, out);
		return;
	}
	Pcg32 rng(jitter_seed ? jitter_seed : 1, 0x5157);
	auto wp = [&](long i) { return w[(size_t)(((i % (long)n) + (long)n) % (long)n)]; };
	auto position_at = [&](float t) {
		// Catmull-Rom through the (closed) waypoint loop; eye a little below the cell centre
		const long i = (long)std::floor(t);
		const float f = t - (float)i;
		const Vec3f p0 = wp(i - 1), p1 = wp(i), p2 = wp(i + 1), p3 = wp(i + 2);
		const float f2 = f * f, f3 = f2 * f;
		Vec3f r = (p1 * 2.0f + (p2 - p0) * f + (p0 * 2.0f - p1 * 5.0f + p2 * 4.0f - p3) * f2 + (p1 * 3.0f - p0 - p2 * 3.0f + p3) * f3) * 0.5f;
		// Catmull-Rom overshoots at corridor corners: pull the eye back into the segment's bounding box so it stays in empty space
		for (int a = 0; a < 3; ++a)
		{
			const float lo = std::fmin(p1[a], p2[a]) - 40.0f, hi = std::fmax(p1[a], p2[a]) + 40.0f;
			r.at(a) = std::fmin(std::fmax(r[a], lo), hi);
		}
		return r;
	};
	for (uint32_t k = 0; k < num_frames; ++k)
	{
		const float t = (float)k * (float)n / (float)num_frames;
		CameraPose pose;
		pose.position = position_at(t);
		const Vec3f ahead = position_at(t + 0.35f);
		const Vec3f d = ahead - pose.position;
		pose.azimuth = std::atan2(d.y, d.x) - 3.14159265358979323846f * 0.5f; // azimuth 0 looks along +Y (matrix.rs:22)
		const float dxy = std::sqrt(d.x * d.x + d.y * d.y);
		pose.elevation = std::atan2(d.z, std::fmax(dxy, 1.0f)) * 0.5f;
		pose.position.z -= scene.params.cell_size[2] * 0.12f;
		if (jitter_seed)
		{
			pose.position.x += rng.uniform(-12.0f, 12.0f);
			pose.position.y += rng.uniform(-12.0f, 12.0f);
			pose.position.z += rng.uniform(-8.0f, 8.0f);
			pose.azimuth += rng.uniform(-0.25f, 0.25f);
			pose.elevation += rng.uniform(-0.15f, 0.15f);
		}
		else
		{
			// deterministic look-around so that floors, ceilings and walls all take part
			pose.azimuth += 0.5f * std::sin((float)k * 0.37f);
			pose.elevation += 0.25f * std::sin((float)k * 0.23f + 1.0f);
		}
		out.push_back(pose);
	}
}

} // namespace sqwh
