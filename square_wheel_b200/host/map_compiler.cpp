// Map compiler: Quake-style .map text -> BSPMap.  A C++ restatement of the reference's offline tool
// chain, function by function (all paths relative to /root/reference/src/):
//   map_file_common.rs / map_file_q1.rs   parse_map_file_content       text -> entities with brushes of three-point planes
//   map_polygonizer.rs                    polygonize_map               brush planes -> convex polygons + texture equations
//   map_csg.rs                            perform_csg_for_map_brushes  parts of polygons inside other solid brushes removed
//   bsp_builder.rs                        build_leaf_bsp_tree          leaf BSP tree, portals between leafs, unreachable leafs dropped
//                                         build_submodel_bsp_tree      one node BSP per brush entity
//   bsp_map_compact_conversion.rs         convert_bsp_map_to_compact_format (+ split_long_polygons, bsp_builder.rs:1282)
//   lightmaps_builder.rs:216              build_dummy_lightmaps        (the lightmapper, lightmapper.cpp, replaces them)
// Host-only, offline; nothing here runs per frame.  Where the reference iterates a HashMap (entity keys, portals grouped by node,
// texture indices) the order is unspecified there; here it is the insertion / tree order, so the output is deterministic.
// Parity with the Rust tool's output is UNPINNED (no Rust toolchain in this image: nothing to compare a compiled map with); what
// is checked is structural (tests/test_map_compiler_cpu.py: closed convex leafs, portal symmetry, every vertex on its plane,
// lightmap sizes within the splitter's limit, rendering through the oracle).
#include "map.hpp"
#include "map_compiler.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <unordered_map>

namespace sqwh
{
namespace mapc
{

// ---- map_file_common.rs ---------------------------------------------------------------------------------------------
namespace
{
struct ParseError
{
	std::string text_left;
};
struct Cursor
{
	const char* p;
	const char* end;
	bool empty() const { return p >= end; }
	bool starts_with(char c) const { return p < end && *p == c; }
	bool starts_with2(char a, char b) const { return p + 1 < end && p[0] == a && p[1] == b; }
};
[[noreturn]] void fail(const Cursor& it) { throw ParseError{std::string(it.p, std::min<size_t>((size_t)(it.end - it.p), 128))}; }
bool is_ascii_whitespace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\x0C'; }

void skip_whitespaces(Cursor& it) // :87-110
{
	while (!it.empty())
	{
		if (is_ascii_whitespace(*it.p))
			++it.p;
		else if (it.starts_with2('/', '/'))
		{
			it.p += 2;
			while (!it.empty() && *it.p != '\n')
				++it.p;
		}
		else
			break;
	}
}
std::string parse_quoted_string(Cursor& it) // :57-77
{
	++it.p; // skip "
	std::string result;
	while (!it.empty() && *it.p != '"')
		result.push_back(*it.p++);
	if (!it.starts_with('"'))
		fail(it);
	++it.p;
	return result;
}
std::string parse_whitespace_separated_string(Cursor& it) // :34-55
{
	std::string result;
	skip_whitespaces(it);
	while (!it.empty())
	{
		const char c = *it.p++;
		if (is_ascii_whitespace(c))
			break;
		result.push_back(c);
	}
	return result;
}
float parse_number(Cursor& it) // :127-152: the characters '.', '0'..'9', '-' and Rust's f32 parse of them
{
	std::string s;
	skip_whitespaces(it);
	while (!it.empty() && (*it.p == '.' || (*it.p >= '0' && *it.p <= '9') || *it.p == '-'))
		s.push_back(*it.p++);
	if (s.empty())
		fail(it);
	char* e = nullptr;
	const float v = std::strtof(s.c_str(), &e);
	if (e == s.c_str() || *e != 0)
		fail(it);
	return v;
}
Vec3f parse_brush_plane_vertex(Cursor& it) // :154-176
{
	skip_whitespaces(it);
	if (!it.starts_with('('))
		fail(it);
	++it.p;
	Vec3f r;
	r.x = parse_number(it);
	r.y = parse_number(it);
	r.z = parse_number(it);
	skip_whitespaces(it);
	if (!it.starts_with(')'))
		fail(it);
	++it.p;
	return r;
}
} // namespace

bool parse_vec3(const std::string& s, Vec3f& out) // :112-120
{
	Cursor it{s.data(), s.data() + s.size()};
	try
	{
		out.x = parse_number(it);
		out.y = parse_number(it);
		out.z = parse_number(it);
		return true;
	}
	catch (const ParseError&)
	{
		return false;
	}
}
bool parse_key_value_number(const std::string& s, float& out) // :122-125
{
	Cursor it{s.data(), s.data() + s.size()};
	try
	{
		out = parse_number(it);
		return true;
	}
	catch (const ParseError&)
	{
		return false;
	}
}

// ---- map_file_q1.rs ---------------------------------------------------------------------------------------------------
namespace
{
struct BrushPlane // :4-11
{
	Vec3f vertices[3];
	std::string texture;
	Vec2f tc_offset, tc_scale;
	float tc_angle = 0;
};
typedef std::vector<BrushPlane> RawBrush;
struct RawEntity
{
	std::vector<RawBrush> brushes;
	KeyList keys;
};

void set_key(KeyList& keys, std::string k, std::string v) // HashMap::insert: a repeated key keeps its slot, takes the new value
{
	for (auto& kv : keys)
		if (kv.first == k)
		{
			kv.second = std::move(v);
			return;
		}
	keys.emplace_back(std::move(k), std::move(v));
}

BrushPlane parse_brush_plane(Cursor& it) // :103-115
{
	BrushPlane bp;
	for (int i = 0; i < 3; ++i)
		bp.vertices[i] = parse_brush_plane_vertex(it);
	bp.texture = parse_whitespace_separated_string(it);
	bp.tc_offset.x = parse_number(it);
	bp.tc_offset.y = parse_number(it);
	bp.tc_angle = parse_number(it);
	bp.tc_scale.x = parse_number(it);
	bp.tc_scale.y = parse_number(it);
	return bp;
}
RawBrush parse_brush(Cursor& it) // :82-101
{
	++it.p; // '{'
	RawBrush result;
	skip_whitespaces(it); // (the reference enters parse_brush_plane, whose first step skips them)
	while (!it.empty() && !it.starts_with('}'))
	{
		result.push_back(parse_brush_plane(it));
		skip_whitespaces(it);
	}
	if (!it.starts_with('}'))
		fail(it);
	++it.p;
	return result;
}
RawEntity parse_entity(Cursor& it) // :46-80
{
	++it.p; // '{'
	RawEntity result;
	while (!it.empty() && !it.starts_with('}'))
	{
		skip_whitespaces(it);
		if (it.starts_with('{'))
			result.brushes.push_back(parse_brush(it));
		else if (it.starts_with('"'))
		{
			std::string k = parse_quoted_string(it);
			skip_whitespaces(it);
			std::string v = parse_quoted_string(it);
			skip_whitespaces(it);
			set_key(result.keys, std::move(k), std::move(v));
		}
		else if (it.starts_with('}'))
			break;
		else
			fail(it);
		skip_whitespaces(it);
	}
	if (!it.starts_with('}'))
		fail(it);
	++it.p;
	return result;
}
std::vector<RawEntity> parse_map_file_content(const std::string& content) // :24-44
{
	std::vector<RawEntity> result;
	Cursor it{content.data(), content.data() + content.size()};
	while (!it.empty())
	{
		skip_whitespaces(it);
		if (it.empty())
			break;
		if (it.starts_with('{'))
			result.push_back(parse_entity(it));
		else
			fail(it);
		skip_whitespaces(it);
	}
	return result;
}

// ---- small linear algebra ----------------------------------------------------------------------------------------------
inline float magnitude2(Vec3f v) { return dot(v, v); }
inline bool plane_eq(const Plane& a, const Plane& b) { return a.vec.x == b.vec.x && a.vec.y == b.vec.y && a.vec.z == b.vec.z && a.dist == b.dist; }
inline Plane inverted(const Plane& p) { return Plane{-p.vec, -p.dist}; }

// Point where three planes meet: solves rows (a.vec, b.vec, c.vec) * p = (a.dist, b.dist, c.dist) through the inverse matrix like
// the reference's `Mat3f::from_cols(..).transpose().invert()` (cgmath: None exactly when the determinant is 0).
bool three_planes_point(const Plane& a, const Plane& b, const Plane& c, Vec3f& out)
{
	const Vec3f r0 = a.vec, r1 = b.vec, r2 = c.vec;
	const Vec3f c12 = cross(r1, r2), c20 = cross(r2, r0), c01 = cross(r0, r1);
	const float det = dot(r0, c12);
	if (det == 0.0f || !(det == det))
		return false;
	const float inv = 1.0f / det;
	out = (c12 * a.dist + c20 * b.dist + c01 * c.dist) * inv;
	return true;
}

// ---- map_polygonizer.rs -------------------------------------------------------------------------------------------------
bool get_brush_side_plane(const BrushPlane& s, Plane& out) // :294-306
{
	const Vec3f vec = cross(s.vertices[0] - s.vertices[1], s.vertices[2] - s.vertices[1]);
	if (vec.x == 0.0f && vec.y == 0.0f && vec.z == 0.0f)
		return false;
	out.vec = vec;
	out.dist = dot(vec, s.vertices[0]);
	return true;
}
} // namespace

std::vector<Vec3f> remove_duplicate_vertices(const std::vector<Vec3f>& in) // :308-336
{
	const float DIST_EPS = 1.0f / 16.0f;
	std::vector<Vec3f> result;
	for (const Vec3f& v : in)
	{
		bool duplicate = false;
		for (const Vec3f& e : result)
			if (magnitude2(e - v) <= DIST_EPS * DIST_EPS)
			{
				duplicate = true;
				break;
			}
		if (!duplicate)
			result.push_back(v);
	}
	return result;
}

std::vector<Vec3f> sort_convex_polygon_vertices(std::vector<Vec3f> in, const Plane& plane) // :338-400
{
	Vec3f sum{0, 0, 0};
	for (const Vec3f& v : in)
		sum = sum + v;
	const Vec3f middle = sum / (float)in.size();
	std::vector<Vec3f> result;
	result.push_back(in.back());
	in.pop_back();
	while (!in.empty())
	{
		const Vec3f v0 = result.back() - middle;
		int best = -1;
		float best_cotan = 0;
		for (size_t i = 0; i < in.size(); ++i)
		{
			const Vec3f v1 = in[i] - middle;
			const float d = dot(v0, v1);
			const float cpd = dot(cross(v0, v1), plane.vec);
			if (cpd >= 0.0f)
				continue; // wrong direction
			const float scaled_angle_cotan = d / cpd;
			if (best < 0 || scaled_angle_cotan < best_cotan)
			{
				best = (int)i;
				best_cotan = scaled_angle_cotan;
			}
		}
		if (best < 0)
			break;
		result.push_back(in[(size_t)best]);
		in.erase(in.begin() + best);
	}
	return result;
}

namespace
{
void get_texture_basis(const Vec3f& n, Vec3f out[2]) // :447-497 (QBSP TextureAxisFromPlane)
{
	static const float B[6][3][3] = {
		{{0, 0, 1}, {1, 0, 0}, {0, -1, 0}}, {{0, 0, -1}, {1, 0, 0}, {0, -1, 0}}, {{1, 0, 0}, {0, 1, 0}, {0, 0, -1}},
		{{-1, 0, 0}, {0, 1, 0}, {0, 0, -1}}, {{0, 1, 0}, {1, 0, 0}, {0, 0, -1}}, {{0, -1, 0}, {1, 0, 0}, {0, 0, -1}}};
	float best_dot = 0.0f;
	int best = 0;
	for (int i = 0; i < 6; ++i)
	{
		const float d = n.x * B[i][0][0] + n.y * B[i][0][1] + n.z * B[i][0][2];
		if (d > best_dot)
		{
			best = i;
			best_dot = d;
		}
	}
	out[0] = Vec3f{B[best][1][0], B[best][1][1], B[best][1][2]};
	out[1] = Vec3f{B[best][2][0], B[best][2][1], B[best][2][2]};
}

TextureInfo get_polygon_texture_info(const BrushPlane& bp, const Vec3f& normal) // :402-428
{
	Vec3f basis[2];
	get_texture_basis(normal, basis);
	const float angle_rad = bp.tc_angle * (3.14159265358979323846f / 180.0f);
	const float c = std::cos(angle_rad), s = std::sin(angle_rad);
	const Vec3f r0 = basis[0] * c - basis[1] * s, r1 = basis[0] * s + basis[1] * c;
	TextureInfo t;
	t.tex_coord_equation[0] = Plane{r0 / bp.tc_scale.x, bp.tc_offset.x};
	t.tex_coord_equation[1] = Plane{r1 / bp.tc_scale.y, bp.tc_offset.y};
	t.texture = bp.texture;
	return t;
}

Brush polygonize_brush(const RawBrush& brush, std::string& log) // :112-199
{
	Brush result;
	const size_t n = brush.size();
	std::vector<Plane> planes(n);
	std::vector<char> ok(n);
	for (size_t i = 0; i < n; ++i)
		ok[i] = get_brush_side_plane(brush[i], planes[i]);
	for (size_t i = 0; i < n; ++i)
	{
		if (!ok[i])
			continue;
		std::vector<Vec3f> vertices;
		for (size_t j = 0; j < n; ++j)
		{
			if (j == i || !ok[j])
				continue;
			for (size_t k = j + 1; k < n; ++k)
			{
				if (k == i || !ok[k])
					continue;
				Vec3f p;
				if (!three_planes_point(planes[i], planes[j], planes[k], p))
					continue;
				bool behind = false;
				for (size_t l = 0; l < n; ++l)
				{
					if (l == i || l == j || l == k || !ok[l])
						continue;
					if (dot(p, planes[l].vec) > planes[l].dist)
					{
						behind = true;
						break;
					}
				}
				if (!behind)
					vertices.push_back(p);
			}
		}
		vertices = remove_duplicate_vertices(vertices);
		if (vertices.size() < 3)
		{
			log += "Wrong polygon with only " + std::to_string(vertices.size()) + " vertices\n";
			continue;
		}
		std::vector<Vec3f> sorted = sort_convex_polygon_vertices(std::move(vertices), planes[i]);
		if (sorted.size() < 3)
		{
			log += "Wrong polygon with only " + std::to_string(sorted.size()) + " vertices_sorted\n";
			continue;
		}
		MapPolygon poly;
		poly.plane = planes[i];
		poly.texture_info = get_polygon_texture_info(brush[i], planes[i].vec);
		poly.vertices = std::move(sorted);
		result.push_back(std::move(poly));
	}
	return result;
}
} // namespace

// ---- clipping_bsp.rs ----------------------------------------------------------------------------------------------------
const float POINT_POSITION_EPS = 1.0f / 16.0f;

PointPosition get_point_position_relative_plane(const Vec3f& point, const Plane& plane) // :67-85
{
	const float dist_scaled = dot(point, plane.vec) - plane.dist;
	const float eps_scaled = POINT_POSITION_EPS * magnitude(plane.vec);
	if (dist_scaled > eps_scaled)
		return POINT_FRONT;
	if (dist_scaled < -eps_scaled)
		return POINT_BACK;
	return POINT_ON_PLANE;
}

PolygonPosition get_polygon_position_relative_plane(const MapPolygon& polygon, const Plane& plane) // :15-55
{
	int front = 0, back = 0;
	const float eps_scaled = POINT_POSITION_EPS * magnitude(plane.vec);
	for (const Vec3f& v : polygon.vertices)
	{
		const float dist_scaled = dot(v, plane.vec) - plane.dist;
		if (dist_scaled > eps_scaled)
			++front;
		else if (dist_scaled < -eps_scaled)
			++back;
	}
	if (front != 0 && back != 0)
		return POLY_SPLITTED;
	if (front != 0)
		return POLY_FRONT;
	if (back != 0)
		return POLY_BACK;
	return dot(polygon.plane.vec, plane.vec) >= 0.0f ? POLY_COPLANAR_FRONT : POLY_COPLANAR_BACK;
}

Vec3f get_line_plane_intersection(const Vec3f& v0, const Vec3f& v1, const Plane& plane) // clipping.rs:177-185
{
	const float dist0 = dot(v0, plane.vec) - plane.dist, dist1 = dot(v1, plane.vec) - plane.dist;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
	return v0 * k1 - v1 * k0;
}

void split_polygon(const MapPolygon& in, const Plane& plane, MapPolygon& front, MapPolygon& back) // :88-147
{
	front.plane = back.plane = in.plane;
	front.texture_info = back.texture_info = in.texture_info;
	front.vertices.clear();
	back.vertices.clear();
	const Vec3f* prev = &in.vertices.back();
	PointPosition prev_pos = get_point_position_relative_plane(*prev, plane);
	for (const Vec3f& vert : in.vertices)
	{
		const PointPosition pos = get_point_position_relative_plane(vert, plane);
		if (pos == POINT_FRONT)
		{
			if (prev_pos == POINT_BACK)
			{
				const Vec3f x = get_line_plane_intersection(*prev, vert, plane);
				back.vertices.push_back(x);
				front.vertices.push_back(x);
			}
			front.vertices.push_back(vert);
		}
		else if (pos == POINT_BACK)
		{
			if (prev_pos == POINT_FRONT)
			{
				const Vec3f x = get_line_plane_intersection(*prev, vert, plane);
				front.vertices.push_back(x);
				back.vertices.push_back(x);
			}
			back.vertices.push_back(vert);
		}
		else
		{
			front.vertices.push_back(vert);
			back.vertices.push_back(vert);
		}
		prev = &vert;
		prev_pos = pos;
	}
}

// ---- map_csg.rs -----------------------------------------------------------------------------------------------------------
namespace
{
const MaterialDef* find_material(const MaterialTable& materials, const std::string& name)
{
	auto it = materials.find(name);
	return it == materials.end() ? nullptr : &it->second;
}

MapPolygon cut_polygon_by_planes(const MapPolygon& polygon, const Plane* planes, size_t num_planes) // :330-360
{
	MapPolygon result = polygon;
	for (size_t i = 0; i < num_planes; ++i)
	{
		switch (get_polygon_position_relative_plane(result, planes[i]))
		{
		case POLY_FRONT:
		case POLY_COPLANAR_FRONT: break;
		case POLY_BACK:
		case POLY_COPLANAR_BACK: result.vertices.clear(); return result;
		case POLY_SPLITTED:
		{
			MapPolygon f, b;
			split_polygon(result, planes[i], f, b);
			if (f.vertices.size() < 3)
				return f;
			result = std::move(f);
		}
		}
	}
	return result;
}

// :162-328: the polygon minus a convex hole inside it, as pieces cut along the polygon's texture axes
std::vector<MapPolygon> make_hole_in_polygon(const MapPolygon& polygon, const MapPolygon& hole)
{
	const Vec3f normals[4] = {polygon.texture_info.tex_coord_equation[0].vec, -polygon.texture_info.tex_coord_equation[0].vec, polygon.texture_info.tex_coord_equation[1].vec,
		-polygon.texture_info.tex_coord_equation[1].vec};
	const size_t nh = hole.vertices.size();
	struct VertexCut
	{
		bool is_double;
		Vec3f a, b; // Single(a) | Double(a = vec_prev, b = vec)
	};
	std::vector<VertexCut> cuts;
	cuts.reserve(nh);
	for (size_t i = 0; i < nh; ++i)
	{
		const Vec3f v_prev = hole.vertices[(i + nh - 1) % nh], v = hole.vertices[i], v_next = hole.vertices[(i + 1) % nh];
		int selected = -1;
		for (int c = 0; c < 4; ++c)
		{
			const float dist = dot(normals[c], v);
			if (dot(normals[c], v_prev) <= dist && dot(normals[c], v_next) >= dist)
			{
				selected = c;
				break;
			}
		}
		if (selected >= 0)
			cuts.push_back(VertexCut{false, normals[selected], normals[selected]});
		else
		{
			int sel_prev = -1, sel = -1;
			for (int c = 0; c < 4; ++c)
			{
				const float dist = dot(normals[c], v);
				if (dot(normals[c], v_prev) <= dist && sel_prev < 0)
					sel_prev = c;
				if (dot(normals[c], v_next) >= dist && sel < 0)
					sel = c;
			}
			if (sel_prev >= 0 && sel >= 0)
				cuts.push_back(VertexCut{true, normals[sel_prev], normals[sel]});
			else
				return {polygon};
		}
	}
	std::vector<MapPolygon> result;
	for (size_t i = 0; i < nh; ++i)
	{
		const Vec3f v = hole.vertices[i], v_next = hole.vertices[(i + 1) % nh];
		const Vec3f edge_cut_plane_normal = cross(hole.plane.vec, v_next - v);
		if (cuts[i].is_double)
		{
			const Plane planes[2] = {Plane{cuts[i].a, dot(cuts[i].a, v)}, Plane{-cuts[i].b, -dot(cuts[i].b, v)}};
			MapPolygon corner = cut_polygon_by_planes(polygon, planes, 2);
			if (corner.vertices.size() >= 3)
				result.push_back(std::move(corner));
		}
		const Vec3f cut_vec = cuts[i].is_double ? cuts[i].b : cuts[i].a;
		const VertexCut& nc = cuts[(i + 1) % nh];
		const Vec3f cut_vec_next = -nc.a; // Single(vec) -> -vec, Double(vec_prev, _) -> -vec_prev
		const Plane planes[3] = {Plane{cut_vec, dot(cut_vec, v)}, Plane{cut_vec_next, dot(cut_vec_next, v_next)}, Plane{edge_cut_plane_normal, dot(edge_cut_plane_normal, v)}};
		MapPolygon edge_polygon = cut_polygon_by_planes(polygon, planes, 3);
		if (edge_polygon.vertices.size() >= 3)
			result.push_back(std::move(edge_polygon));
	}
	if (result.empty())
		return {polygon};
	return result;
}

std::vector<MapPolygon> cut_polygon_by_brush_planes(const MapPolygon& polygon, const Brush& brush, bool preserve_coplanar) // :103-160
{
	for (const MapPolygon& bp : brush)
		if (get_polygon_position_relative_plane(polygon, bp.plane) == POLY_FRONT)
			return {polygon};
	std::vector<MapPolygon> result;
	MapPolygon leftover = polygon;
	for (const MapPolygon& bp : brush)
	{
		switch (get_polygon_position_relative_plane(leftover, bp.plane))
		{
		case POLY_FRONT: return {polygon};
		case POLY_BACK: break;
		case POLY_COPLANAR_FRONT:
			if (preserve_coplanar)
				return {polygon};
			break;
		case POLY_COPLANAR_BACK: break;
		case POLY_SPLITTED:
		{
			MapPolygon f, b;
			split_polygon(leftover, bp.plane, f, b);
			if (f.vertices.size() >= 3)
				result.push_back(std::move(f)); // outside the brush: kept
			if (b.vertices.size() >= 3)
				leftover = std::move(b);
			else
				return {polygon};
		}
		}
	}
	if (result.size() <= 1)
		return result;
	return make_hole_in_polygon(polygon, leftover);
}

std::vector<MapPolygon> perform_csg_for_entity_brushes(const std::vector<Brush>& brushes, const MaterialTable& materials) // :48-101
{
	std::vector<MapPolygon> result;
	std::vector<char> solid(brushes.size(), 1);
	for (size_t b = 0; b < brushes.size(); ++b)
		for (const MapPolygon& p : brushes[b])
		{
			const MaterialDef* m = find_material(materials, p.texture_info.texture);
			if (m && !m->blocks_view)
			{
				solid[b] = 0;
				break;
			}
		}
	// bounding boxes: a brush that cannot touch the polygon's brush leaves every polygon as it is (the first loop of
	// cut_polygon_by_brush_planes finds a separating face plane for disjoint convex solids; this only skips that work)
	struct Box
	{
		Vec3f lo, hi;
	};
	std::vector<Box> boxes(brushes.size());
	for (size_t b = 0; b < brushes.size(); ++b)
	{
		Box bx{Vec3f{1e30f, 1e30f, 1e30f}, Vec3f{-1e30f, -1e30f, -1e30f}};
		for (const MapPolygon& p : brushes[b])
			for (const Vec3f& v : p.vertices)
				for (int a = 0; a < 3; ++a)
				{
					bx.lo.at(a) = std::fmin(bx.lo[a], v[a]);
					bx.hi.at(a) = std::fmax(bx.hi[a], v[a]);
				}
		boxes[b] = bx;
	}
	for (size_t b = 0; b < brushes.size(); ++b)
	{
		if (brushes[b].empty())
			continue;
		std::vector<MapPolygon> brush_polygons = brushes[b];
		bool preserve_coplanar = false;
		for (size_t o = 0; o < brushes.size(); ++o)
		{
			if (o == b)
			{
				preserve_coplanar = true;
				continue;
			}
			if (!solid[o])
				continue;
			bool apart = false;
			for (int a = 0; a < 3; ++a)
				if (boxes[o].lo[a] > boxes[b].hi[a] + 1.0f || boxes[o].hi[a] < boxes[b].lo[a] - 1.0f)
					apart = true;
			if (apart)
				continue;
			std::vector<MapPolygon> clipped;
			for (MapPolygon& p : brush_polygons)
			{
				std::vector<MapPolygon> pieces = cut_polygon_by_brush_planes(p, brushes[o], preserve_coplanar);
				for (MapPolygon& piece : pieces)
					clipped.push_back(std::move(piece));
			}
			brush_polygons = std::move(clipped);
		}
		for (MapPolygon& p : brush_polygons)
			result.push_back(std::move(p));
	}
	return result;
}

// ---- bsp_builder.rs ---------------------------------------------------------------------------------------------------------
unsigned worker_count()
{
	const unsigned n = std::thread::hardware_concurrency();
	return std::min(std::max(n, 1u), 32u);
}
template <class F> void parallel_for(size_t count, F&& body, bool serial = false) // static interleaved split; results are written per index
{
	const unsigned nt = serial ? 1u : (unsigned)std::min<size_t>(worker_count(), count);
	if (nt <= 1)
	{
		for (size_t i = 0; i < count; ++i)
			body(i);
		return;
	}
	std::vector<std::thread> threads;
	for (unsigned t = 0; t < nt; ++t)
		threads.emplace_back([&, t] {
			for (size_t i = t; i < count; i += nt)
				body(i);
		});
	for (std::thread& th : threads)
		th.join();
}

int num_zero_normal_components(const Plane& plane)
{
	return (plane.vec.x == 0.0f) + (plane.vec.y == 0.0f) + (plane.vec.z == 0.0f);
}

// :282-360: smaller is better; false = not a splitter
bool get_splitter_plane_score(const std::vector<MapPolygon>& polygons, const Plane& plane, float& score)
{
	int front = 0, back = 0, cfront = 0, cback = 0, splitted = 0;
	for (const MapPolygon& p : polygons)
		switch (get_polygon_position_relative_plane(p, plane))
		{
		case POLY_FRONT: ++front; break;
		case POLY_BACK: ++back; break;
		case POLY_COPLANAR_FRONT: ++cfront; break;
		case POLY_COPLANAR_BACK: ++cback; break;
		case POLY_SPLITTED: ++splitted; break;
		}
	const int front_total = front + cfront, back_total = back + cback;
	if (splitted == 0 && (front_total == 0 || back_total == 0))
		return false;
	const int SPLITTED_POLYGON_SCALE = 5;
	const int base_score = std::abs(front_total - back_total) + SPLITTED_POLYGON_SCALE * splitted;
	const int nz = num_zero_normal_components(plane);
	score = (float)base_score;
	if (nz == 0)
		score *= 2.0f;
	if (nz == 1)
		score *= 1.5f;
	return true;
}

// :197-278.  The candidates are scored in parallel; the winner is the FIRST candidate (in the reference's order) with the lowest
// score, exactly what the sequential `if score < prev_score` scan keeps.
bool choose_best_splitter_plane(const std::vector<MapPolygon>& polygons, bool check_edges, Plane& out)
{
	std::vector<Plane> candidates;
	candidates.reserve(polygons.size());
	for (const MapPolygon& p : polygons)
		candidates.push_back(p.plane);
	const size_t num_own = candidates.size();
	// Equal planes (coplanar polygons, axis-aligned edges) score equally and only the first of them can win, so each distinct
	// plane is scored once.
	struct PlaneKey
	{
		uint32_t w[4];
		bool operator==(const PlaneKey& o) const { return std::memcmp(w, o.w, sizeof(w)) == 0; }
	};
	struct PlaneKeyHash
	{
		size_t operator()(const PlaneKey& k) const { return (size_t)(((uint64_t)k.w[0] * 0x9E3779B1u) ^ ((uint64_t)k.w[1] * 0x85EBCA77u) ^ ((uint64_t)k.w[2] * 0xC2B2AE3Du) ^ ((uint64_t)k.w[3] * 0x27D4EB2Fu)); }
	};
	std::unordered_map<PlaneKey, char, PlaneKeyHash> seen;
	auto best_of = [&](size_t first, size_t last, float& best_score, size_t& best_index) {
		std::vector<size_t> unique;
		for (size_t i = first; i < last; ++i)
		{
			PlaneKey k;
			std::memcpy(k.w, &candidates[i], sizeof(k.w));
			if (seen.emplace(k, 1).second)
				unique.push_back(i);
		}
		const size_t count = unique.size();
		std::vector<float> scores(count);
		std::vector<char> valid(count);
		parallel_for(count, [&](size_t i) { valid[i] = get_splitter_plane_score(polygons, candidates[unique[i]], scores[i]); }, count * polygons.size() < 400000);
		bool found = best_index != (size_t)-1;
		for (size_t i = 0; i < count; ++i)
			if (valid[i] && (!found || scores[i] < best_score))
			{
				found = true;
				best_score = scores[i];
				best_index = unique[i];
			}
		return found;
	};
	float best_score = 0;
	size_t best_index = (size_t)-1;
	if (!best_of(0, num_own, best_score, best_index))
		return false; // the polygons form a convex hull: a leaf
	if (check_edges)
	{
		const Vec3f basis_vecs[3] = {Vec3f{1, 0, 0}, Vec3f{0, 1, 0}, Vec3f{0, 0, 1}};
		for (const MapPolygon& p : polygons)
		{
			Vec3f prev_v = p.vertices.back();
			for (const Vec3f& v : p.vertices)
			{
				const Vec3f edge = v - prev_v;
				prev_v = v;
				for (const Vec3f& bv : basis_vecs)
				{
					const Vec3f normal = cross(edge, bv);
					const float sin2 = magnitude2(normal) / magnitude2(edge);
					if (!(sin2 > 0.001f)) // parallel to this basis vector (a zero-length edge gives NaN: `<=` is false for it in the reference)
					{
						if (sin2 == sin2)
							continue;
					}
					candidates.push_back(Plane{normal, dot(normal, v)});
				}
			}
		}
		best_of(num_own, candidates.size(), best_score, best_index);
	}
	out = candidates[best_index];
	return true;
}

struct TreeNode;
struct TreeLeaf
{
	std::vector<MapPolygon> polygons;
	std::vector<size_t> portals; // indices into the portal list
	bool reachable = false;
	uint32_t index = 0;
};
struct TreeChild
{
	std::unique_ptr<TreeNode> node;
	std::unique_ptr<TreeLeaf> leaf;
};
struct TreeNode
{
	Plane plane;
	TreeChild children[2];
};
struct LeafsPortal // :42-50
{
	TreeLeaf* leaf_front;
	TreeLeaf* leaf_back;
	Plane plane;
	std::vector<Vec3f> vertices;
};

TreeChild make_leaf(std::vector<MapPolygon>&& polygons)
{
	TreeChild c;
	c.leaf.reset(new TreeLeaf);
	c.leaf->polygons = std::move(polygons);
	return c;
}

TreeChild build_leaf_bsp_tree_r(std::vector<MapPolygon>&& in_polygons, bool advanced) // :131-194
{
	Plane splitter;
	if (!choose_best_splitter_plane(in_polygons, advanced, splitter))
		return make_leaf(std::move(in_polygons));
	std::vector<MapPolygon> front, back;
	for (MapPolygon& p : in_polygons)
	{
		switch (get_polygon_position_relative_plane(p, splitter))
		{
		case POLY_FRONT:
		case POLY_COPLANAR_FRONT: front.push_back(std::move(p)); break;
		case POLY_BACK:
		case POLY_COPLANAR_BACK: back.push_back(std::move(p)); break;
		case POLY_SPLITTED:
		{
			MapPolygon f, b;
			split_polygon(p, splitter, f, b);
			if (f.vertices.size() >= 3)
				front.push_back(std::move(f));
			if (b.vertices.size() >= 3)
				back.push_back(std::move(b));
		}
		}
	}
	in_polygons.clear();
	if (front.empty())
		return make_leaf(std::move(back));
	if (back.empty())
		return make_leaf(std::move(front));
	TreeChild c;
	c.node.reset(new TreeNode);
	c.node->plane = splitter;
	c.node->children[0] = build_leaf_bsp_tree_r(std::move(front), advanced);
	c.node->children[1] = build_leaf_bsp_tree_r(std::move(back), advanced);
	return c;
}

std::vector<MapPolygon> preprocess_input_polygons(const std::vector<MapPolygon>& polygons, const MaterialTable& materials) // :555-582
{
	std::vector<MapPolygon> result;
	for (const MapPolygon& p : polygons)
	{
		if (const MaterialDef* m = find_material(materials, p.texture_info.texture))
		{
			if (!m->bsp)
				continue;
			if (m->twosided)
			{
				MapPolygon copy = p;
				copy.plane = inverted(copy.plane);
				std::reverse(copy.vertices.begin(), copy.vertices.end());
				result.push_back(std::move(copy));
			}
		}
		result.push_back(p);
	}
	return result;
}

bool are_planes_almost_parallel(const Plane& a, const Plane& b) // :1013-1016
{
	return std::fabs(magnitude(cross(a.vec, b.vec)) / dot(a.vec, b.vec)) < 0.0001f;
}

struct LeafPortalInitial // :629-635
{
	std::vector<Vec3f> vertices;
	Plane plane;
	TreeLeaf* leaf;
	bool is_front;
	const TreeNode* node;
};
struct SplitterRef
{
	const TreeNode* node;
	bool is_front;
};

bool blocks_view(const MaterialTable& materials, const MapPolygon& p)
{
	const MaterialDef* m = find_material(materials, p.texture_info.texture);
	return !m || m->blocks_view;
}

// :700-881: for every splitter plane above the leaf, the convex polygon on it bounded by all other cut planes of the leaf
void build_leaf_portals(TreeLeaf* leaf, const std::vector<SplitterRef>& splitters, const Vec3f& bbox_min, const Vec3f& bbox_max, const MaterialTable& materials, std::vector<LeafPortalInitial>& out)
{
	std::vector<Plane> cut_planes;
	for (const SplitterRef& s : splitters)
		cut_planes.push_back(s.is_front ? inverted(s.node->plane) : s.node->plane);
	for (const MapPolygon& p : leaf->polygons)
		if (blocks_view(materials, p))
			cut_planes.push_back(Plane{-p.plane.vec, -p.plane.dist});
	cut_planes.push_back(Plane{Vec3f{1, 0, 0}, bbox_max.x});
	cut_planes.push_back(Plane{Vec3f{-1, 0, 0}, -bbox_min.x});
	cut_planes.push_back(Plane{Vec3f{0, 1, 0}, bbox_max.y});
	cut_planes.push_back(Plane{Vec3f{0, -1, 0}, -bbox_min.y});
	cut_planes.push_back(Plane{Vec3f{0, 0, 1}, bbox_max.z});
	cut_planes.push_back(Plane{Vec3f{0, 0, -1}, -bbox_min.z});
	const size_t n = cut_planes.size();
	std::vector<float> lens(n);
	for (size_t i = 0; i < n; ++i)
		lens[i] = magnitude(cut_planes[i].vec);

	for (const SplitterRef& s : splitters)
	{
		const Plane portal_plane = s.is_front ? inverted(s.node->plane) : s.node->plane;
		std::vector<Vec3f> vertices;
		for (size_t i = 0; i < n; ++i)
		{
			const Plane& ci = cut_planes[i];
			if (plane_eq(ci, portal_plane) || are_planes_almost_parallel(portal_plane, ci))
				continue;
			for (size_t j = i + 1; j < n; ++j)
			{
				const Plane& cj = cut_planes[j];
				if (plane_eq(cj, portal_plane) || plane_eq(cj, ci) || are_planes_almost_parallel(portal_plane, cj) || are_planes_almost_parallel(ci, cj))
					continue;
				Vec3f p;
				if (!three_planes_point(portal_plane, ci, cj, p))
					continue;
				bool behind = false;
				for (size_t k = 0; k < n; ++k)
				{
					if (k == i || k == j)
						continue;
					const Plane& ck = cut_planes[k];
					if (dot(p, ck.vec) > ck.dist + lens[k] * POINT_POSITION_EPS)
					{
						if (plane_eq(ck, portal_plane))
							continue;
						behind = true;
						break;
					}
				}
				if (!behind)
					vertices.push_back(p);
			}
		}
		if (vertices.size() < 3)
			continue;
		std::vector<Vec3f> dedup = remove_duplicate_vertices(vertices);
		if (dedup.size() < 3)
			continue;
		std::vector<Vec3f> sorted = sort_convex_polygon_vertices(std::move(dedup), s.node->plane);
		if (sorted.size() < 3)
			continue;
		out.push_back(LeafPortalInitial{std::move(sorted), s.node->plane, leaf, s.is_front, s.node});
	}
}

std::vector<Vec3f> build_portals_intersection(const Plane& plane, const std::vector<Vec3f>& v0, const std::vector<Vec3f>& v1) // :938-1011
{
	std::vector<Plane> clip_planes;
	for (const std::vector<Vec3f>* vs : {&v0, &v1})
	{
		const Vec3f* prev = &vs->back();
		for (const Vec3f& v : *vs)
		{
			const Vec3f vec = cross(*prev - v, plane.vec);
			clip_planes.push_back(Plane{vec, dot(vec, v)});
			prev = &v;
		}
	}
	const size_t n = clip_planes.size();
	std::vector<float> lens(n);
	for (size_t i = 0; i < n; ++i)
		lens[i] = magnitude(clip_planes[i].vec);
	std::vector<Vec3f> vertices;
	for (size_t i = 0; i < n; ++i)
		for (size_t j = i + 1; j < n; ++j)
		{
			const Plane &pi = clip_planes[i], &pj = clip_planes[j];
			if (plane_eq(pj, pi) || are_planes_almost_parallel(pi, pj))
				continue;
			Vec3f p;
			if (!three_planes_point(plane, pi, pj, p))
				continue;
			bool behind = false;
			for (size_t k = 0; k < n; ++k)
			{
				if (k == i || k == j)
					continue;
				const Plane& pk = clip_planes[k];
				if (plane_eq(pk, pi) || plane_eq(pk, pj))
					continue;
				if (dot(p, pk.vec) > pk.dist + lens[k] * POINT_POSITION_EPS)
				{
					behind = true;
					break;
				}
			}
			if (!behind)
				vertices.push_back(p);
		}
	if (vertices.size() < 3)
		return vertices;
	std::vector<Vec3f> dedup = remove_duplicate_vertices(vertices);
	if (dedup.size() < 3)
		return dedup;
	return sort_convex_polygon_vertices(std::move(dedup), plane);
}

bool portal_is_fully_covered_by_leaf_polygons(const Plane& portal_plane, const std::vector<Vec3f>& portal_vertices, const TreeLeaf& leaf, const MaterialTable& materials) // :1018-1080
{
	const float PORTAL_POLYGON_COVERAGE_EPS = 0.25f;
	for (const MapPolygon& polygon : leaf.polygons)
	{
		if (!blocks_view(materials, polygon))
			continue;
		const PolygonPosition pos = get_polygon_position_relative_plane(polygon, portal_plane);
		if (!(pos == POLY_COPLANAR_FRONT || pos == POLY_COPLANAR_BACK))
			continue;
		const Vec3f* prev = &polygon.vertices.back();
		bool inside = true;
		for (const Vec3f& pv : polygon.vertices)
		{
			const Vec3f vec = cross(*prev - pv, polygon.plane.vec);
			const float eps_scaled = PORTAL_POLYGON_COVERAGE_EPS * magnitude(vec);
			const float dist = dot(vec, pv);
			bool all_inside = true;
			for (const Vec3f& v : portal_vertices)
				if (dot(v, vec) > dist + eps_scaled)
				{
					all_inside = false;
					break;
				}
			prev = &pv;
			if (!all_inside)
			{
				inside = false;
				break;
			}
		}
		if (inside)
			return true;
	}
	return false;
}

struct Box3
{
	Vec3f lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
	void add(const Vec3f& v)
	{
		for (int a = 0; a < 3; ++a)
		{
			lo.at(a) = std::fmin(lo[a], v[a]);
			hi.at(a) = std::fmax(hi[a], v[a]);
		}
	}
	bool apart(const Box3& o, float margin) const
	{
		for (int a = 0; a < 3; ++a)
			if (o.lo[a] > hi[a] + margin || o.hi[a] < lo[a] - margin)
				return true;
		return false;
	}
};

// :883-936: portals of one node, front side x back side
void build_leafs_portals(const std::vector<const LeafPortalInitial*>& in_portals, const MaterialTable& materials, std::vector<LeafsPortal>& result)
{
	std::vector<Box3> boxes(in_portals.size());
	for (size_t i = 0; i < in_portals.size(); ++i)
		for (const Vec3f& v : in_portals[i]->vertices)
			boxes[i].add(v);
	for (size_t fi = 0; fi < in_portals.size(); ++fi)
	{
		const LeafPortalInitial& pf = *in_portals[fi];
		if (!pf.is_front)
			continue;
		const Plane plane = pf.plane;
		for (size_t bi = 0; bi < in_portals.size(); ++bi)
		{
			const LeafPortalInitial& pb = *in_portals[bi];
			if (pb.is_front)
				continue;
			if (boxes[fi].apart(boxes[bi], 1.0f)) // disjoint convex polygons have no common area (only skips the work below)
				continue;
			std::vector<Vec3f> x = build_portals_intersection(plane, pb.vertices, pf.vertices);
			if (x.size() < 3)
				continue;
			if (portal_is_fully_covered_by_leaf_polygons(plane, x, *pf.leaf, materials) || portal_is_fully_covered_by_leaf_polygons(plane, x, *pb.leaf, materials))
				continue;
			result.push_back(LeafsPortal{pf.leaf, pb.leaf, plane, std::move(x)});
		}
	}
}

void collect_leafs_r(TreeChild& c, std::vector<SplitterRef>& stack, std::vector<std::pair<TreeLeaf*, std::vector<SplitterRef>>>& out, std::vector<const TreeNode*>& node_order)
{
	if (c.node)
	{
		node_order.push_back(c.node.get());
		stack.push_back(SplitterRef{c.node.get(), true});
		collect_leafs_r(c.node->children[0], stack, out, node_order);
		stack.back().is_front = false;
		collect_leafs_r(c.node->children[1], stack, out, node_order);
		stack.pop_back();
	}
	else
		out.emplace_back(c.leaf.get(), stack);
}

std::vector<LeafsPortal> build_portals(TreeChild& root, const Vec3f& bbox_min, const Vec3f& bbox_max, const MaterialTable& materials) // :604-627
{
	std::vector<std::pair<TreeLeaf*, std::vector<SplitterRef>>> leafs;
	std::vector<const TreeNode*> node_order;
	std::vector<SplitterRef> stack;
	collect_leafs_r(root, stack, leafs, node_order);
	std::vector<std::vector<LeafPortalInitial>> per_leaf(leafs.size());
	parallel_for(leafs.size(), [&](size_t i) { build_leaf_portals(leafs[i].first, leafs[i].second, bbox_min, bbox_max, materials, per_leaf[i]); });
	// grouped by node: the reference keys a HashMap by node address (iteration order unspecified); here nodes go in pre-order and
	// a node's portals in leaf order
	std::unordered_map<const TreeNode*, std::vector<const LeafPortalInitial*>> by_node;
	for (const std::vector<LeafPortalInitial>& v : per_leaf)
		for (const LeafPortalInitial& p : v)
			by_node[p.node].push_back(&p);
	std::vector<std::vector<LeafsPortal>> per_node(node_order.size());
	parallel_for(node_order.size(), [&](size_t i) {
		auto it = by_node.find(node_order[i]);
		if (it != by_node.end())
			build_leafs_portals(it->second, materials, per_node[i]);
	});
	std::vector<LeafsPortal> result;
	for (std::vector<LeafsPortal>& v : per_node)
		for (LeafsPortal& p : v)
			result.push_back(std::move(p));
	return result;
}

TreeLeaf* get_leaf_for_point(TreeChild& c, const Vec3f& point) // :1063-1083
{
	TreeChild* cur = &c;
	while (cur->node)
		cur = dot(cur->node->plane.vec, point) >= cur->node->plane.dist ? &cur->node->children[0] : &cur->node->children[1];
	return cur->leaf.get();
}

// :1128-1180: true = keep this child
bool remove_unreachable_leafs_r(TreeChild& c, uint32_t& total, uint32_t& removed)
{
	if (c.node)
	{
		const bool keep_front = remove_unreachable_leafs_r(c.node->children[0], total, removed);
		const bool keep_back = remove_unreachable_leafs_r(c.node->children[1], total, removed);
		if (!keep_front && !keep_back)
			return false;
		if (!keep_back)
		{
			TreeChild child = std::move(c.node->children[0]);
			c = std::move(child);
		}
		else if (!keep_front)
		{
			TreeChild child = std::move(c.node->children[1]);
			c = std::move(child);
		}
		return true;
	}
	++total;
	if (!c.leaf->reachable)
		++removed;
	return c.leaf->reachable;
}

size_t get_tree_max_depth_r(const TreeChild& c)
{
	if (c.node)
		return std::max(get_tree_max_depth_r(c.node->children[0]), get_tree_max_depth_r(c.node->children[1])) + 1;
	return 1;
}

// ---- submodel BSP trees (:362-553) -----------------------------------------------------------------------------------------
struct SubmodelNode
{
	Plane plane;
	std::vector<MapPolygon> polygons;
	std::unique_ptr<SubmodelNode> children[2];
};

bool get_submodel_splitter_plane_score(const std::vector<MapPolygon>& polygons, const Plane& plane, float& score) // :482-553
{
	int front = 0, back = 0, splitted = 0, cfront = 0, cback = 0;
	for (const MapPolygon& p : polygons)
		switch (get_polygon_position_relative_plane(p, plane))
		{
		case POLY_FRONT: ++front; break;
		case POLY_BACK: ++back; break;
		case POLY_COPLANAR_FRONT: ++cfront; break;
		case POLY_COPLANAR_BACK: ++cback; break;
		case POLY_SPLITTED: ++splitted; break;
		}
	if (cfront + cback == (int)polygons.size())
		return false;
	if (splitted == 0 && (cfront == 0 || cback == 0) && (front == 0 || back == 0))
		return false;
	const int base_score = std::abs(front - back) + 5 * splitted;
	const int nz = num_zero_normal_components(plane);
	score = (float)base_score;
	if (nz == 0)
		score *= 2.0f;
	if (nz == 1)
		score *= 1.5f;
	return true;
}

std::unique_ptr<SubmodelNode> build_submodel_bsp_tree_r(std::vector<MapPolygon>&& in) // :380-455
{
	std::unique_ptr<SubmodelNode> node(new SubmodelNode);
	bool found = false;
	float best_score = 0;
	Plane splitter;
	for (const MapPolygon& p : in)
	{
		float score;
		if (get_submodel_splitter_plane_score(in, p.plane, score) && (!found || score < best_score))
		{
			found = true;
			best_score = score;
			splitter = p.plane;
		}
	}
	if (!found)
	{
		node->plane = in[0].plane; // dummy plane of a leaf node
		node->polygons = std::move(in);
		return node;
	}
	std::vector<MapPolygon> front, back, coplanar;
	for (MapPolygon& p : in)
		switch (get_polygon_position_relative_plane(p, splitter))
		{
		case POLY_FRONT: front.push_back(std::move(p)); break;
		case POLY_BACK: back.push_back(std::move(p)); break;
		case POLY_COPLANAR_FRONT:
		case POLY_COPLANAR_BACK: coplanar.push_back(std::move(p)); break;
		case POLY_SPLITTED:
		{
			MapPolygon f, b;
			split_polygon(p, splitter, f, b);
			if (f.vertices.size() >= 3)
				front.push_back(std::move(f));
			if (b.vertices.size() >= 3)
				back.push_back(std::move(b));
		}
		}
	node->plane = splitter;
	node->polygons = std::move(coplanar);
	if (!front.empty())
		node->children[0] = build_submodel_bsp_tree_r(std::move(front));
	if (!back.empty())
		node->children[1] = build_submodel_bsp_tree_r(std::move(back));
	return node;
}

// ---- split_long_polygons (bsp_builder.rs:1282-1362) --------------------------------------------------------------------------
const uint32_t MAX_LIGHTMAP_SIZE = 17; // common/lightmap.rs:9

void split_long_polygon_r(const MapPolygon& polygon, std::vector<MapPolygon>& out, size_t depth)
{
	if (polygon.vertices.size() < 3 || depth > 10)
		return;
	const float inf = (float)(1 << 29);
	for (int i = 0; i < 2; ++i)
	{
		float tc_min = inf, tc_max = -inf;
		const Plane& eq = polygon.texture_info.tex_coord_equation[i];
		for (const Vec3f& v : polygon.vertices)
		{
			const float tc = dot(eq.vec, v) + eq.dist;
			if (tc < tc_min)
				tc_min = tc;
			if (tc > tc_max)
				tc_max = tc;
		}
		const int32_t tc_min_int = (int32_t)std::floor(tc_min), tc_max_int = (int32_t)std::ceil(tc_max);
		if (get_lightmap_size(tc_min_int, tc_max_int) <= MAX_LIGHTMAP_SIZE)
			continue;
		const int32_t middle_tc = (tc_max_int + tc_min_int) >> 1;
		const float shift = (float)(middle_tc & ~(int32_t)(LIGHTMAP_SCALE - 1));
		const Plane split_plane{eq.vec, -eq.dist + shift};
		MapPolygon p0, p1;
		split_polygon(polygon, split_plane, p0, p1);
		split_long_polygon_r(p0, out, depth + 1);
		split_long_polygon_r(p1, out, depth + 1);
		return;
	}
	out.push_back(polygon);
}

std::vector<MapPolygon> split_long_polygons(const std::vector<MapPolygon>& polygons, const MaterialTable& materials)
{
	std::vector<MapPolygon> result;
	for (const MapPolygon& p : polygons)
	{
		bool need_split = true;
		if (const MaterialDef* m = find_material(materials, p.texture_info.texture))
			if (!m->light || m->skybox)
				need_split = false;
		if (need_split)
			split_long_polygon_r(p, result, 0);
		else
			result.push_back(p);
	}
	return result;
}

// ---- bsp_map_compact_conversion.rs ------------------------------------------------------------------------------------------
struct Converter
{
	BSPMap& out;
	const MaterialTable& materials;
	std::vector<std::string> texture_names; // index = texture index (first use)
	std::unordered_map<std::string, uint32_t> texture_index;
	std::unordered_map<std::string, StringRef> strings_cache;

	uint32_t get_texture_index(const std::string& name) // :261-270
	{
		auto it = texture_index.find(name);
		if (it != texture_index.end())
			return it->second;
		const uint32_t index = (uint32_t)texture_names.size();
		texture_index.emplace(name, index);
		texture_names.push_back(name);
		return index;
	}

	Polygon convert_polygon(const MapPolygon& polygon) // :205-259
	{
		Polygon p;
		p.first_vertex = (uint32_t)out.vertices.size();
		p.num_vertices = (uint32_t)polygon.vertices.size();
		out.vertices.insert(out.vertices.end(), polygon.vertices.begin(), polygon.vertices.end());
		const float inf = (float)(1 << 29);
		float tc_min[2] = {inf, inf}, tc_max[2] = {-inf, -inf};
		for (const Vec3f& v : polygon.vertices)
			for (int i = 0; i < 2; ++i)
			{
				const float tc = dot(polygon.texture_info.tex_coord_equation[i].vec, v) + polygon.texture_info.tex_coord_equation[i].dist;
				if (tc < tc_min[i])
					tc_min[i] = tc;
				if (tc > tc_max[i])
					tc_max[i] = tc;
			}
		const float tc_reduce_eps = 1.0f / 32.0f;
		for (int i = 0; i < 2; ++i)
		{
			tc_min[i] += tc_reduce_eps;
			tc_max[i] -= tc_reduce_eps;
			p.tex_coord_min[i] = (int32_t)std::floor(tc_min[i]);
			p.tex_coord_max[i] = std::max((int32_t)std::ceil(tc_max[i]), p.tex_coord_min[i] + 1);
			p.tex_coord_equation[i] = polygon.texture_info.tex_coord_equation[i];
		}
		p.plane = polygon.plane;
		p.lightmap_data_offset = 0;
		p.texture = get_texture_index(polygon.texture_info.texture);
		return p;
	}

	uint32_t convert_child(const TreeChild& c, const std::vector<LeafsPortal>& portals) // :82-205
	{
		if (c.node)
		{
			BSPNode n;
			n.children[0] = convert_child(c.node->children[0], portals);
			n.children[1] = convert_child(c.node->children[1], portals);
			n.plane = c.node->plane;
			out.nodes.push_back(n);
			return (uint32_t)out.nodes.size() - 1;
		}
		TreeLeaf& leaf = *c.leaf;
		const uint32_t leaf_index = (uint32_t)out.leafs.size();
		leaf.index = leaf_index;
		const std::vector<MapPolygon> splitted = split_long_polygons(leaf.polygons, materials);
		BSPLeaf l;
		l.first_polygon = (uint32_t)out.polygons.size();
		for (const MapPolygon& p : splitted)
			out.polygons.push_back(convert_polygon(p));
		l.num_polygons = (uint32_t)splitted.size();
		l.first_leaf_portal = (uint32_t)out.leafs_portals.size();
		for (size_t pi : leaf.portals)
			out.leafs_portals.push_back((uint32_t)pi);
		l.num_leaf_portals = (uint32_t)leaf.portals.size();
		out.leafs.push_back(l);
		return leaf_index + FIRST_LEAF_INDEX;
	}

	uint32_t convert_submodel_node_r(const SubmodelNode& node) // :371-410
	{
		uint32_t children[2] = {~0u, ~0u};
		for (int i = 0; i < 2; ++i)
			if (node.children[i])
				children[i] = convert_submodel_node_r(*node.children[i]);
		SubmodelBSPNode n;
		n.plane = node.plane;
		n.first_polygon = (uint32_t)out.polygons.size();
		const std::vector<MapPolygon> splitted = split_long_polygons(node.polygons, materials);
		for (const MapPolygon& p : splitted)
			out.polygons.push_back(convert_polygon(p));
		n.num_polygons = (uint32_t)splitted.size();
		n.children[0] = children[0];
		n.children[1] = children[1];
		out.submodels_bsp_nodes.push_back(n);
		return (uint32_t)out.submodels_bsp_nodes.size() - 1;
	}

	StringRef convert_string(const std::string& s) // :436-459
	{
		auto it = strings_cache.find(s);
		if (it != strings_cache.end())
			return it->second;
		if (out.strings_data.size() > 65535)
			return StringRef{0, 0};
		const size_t offset = out.strings_data.size();
		out.strings_data.insert(out.strings_data.end(), s.begin(), s.end());
		const StringRef r{(uint16_t)offset, (uint16_t)s.size()};
		strings_cache.emplace(s, r);
		return r;
	}
};

} // namespace

// ---- the tool (map_compiler/main.rs:38-125) ---------------------------------------------------------------------------------------
bool compile_map(const std::string& map_text, const MaterialTable& materials, const CompileOptions& options, BSPMap& out_map, CompileStats& stats, std::string& log)
{
	out_map = BSPMap{};
	stats = CompileStats{};
	std::vector<RawEntity> parsed;
	try
	{
		parsed = parse_map_file_content(map_text);
	}
	catch (const ParseError& e)
	{
		log += "parse error near: " + e.text_left + "\n";
		return false;
	}
	if (parsed.empty())
	{
		log += "no entities\n";
		return false;
	}
	// polygonize_map (map_polygonizer.rs:28-66): brushes that produce no polygon are dropped
	std::vector<PolygonizedEntity> polygonized(parsed.size());
	for (size_t e = 0; e < parsed.size(); ++e)
	{
		for (const RawBrush& rb : parsed[e].brushes)
		{
			Brush b = polygonize_brush(rb, log);
			if (!b.empty())
				polygonized[e].brushes.push_back(std::move(b));
		}
		polygonized[e].keys = parsed[e].keys;
	}
	for (const Brush& b : polygonized[0].brushes)
		stats.initial_polygons += (uint32_t)b.size();

	// map_csg.rs:19-46
	std::vector<CsgEntity> csg(polygonized.size());
	parallel_for(polygonized.size(), [&](size_t e) {
		if (options.perform_csg)
			csg[e].polygons = perform_csg_for_entity_brushes(polygonized[e].brushes, materials);
		else
			for (const Brush& b : polygonized[e].brushes)
				csg[e].polygons.insert(csg[e].polygons.end(), b.begin(), b.end());
		csg[e].keys = polygonized[e].keys;
	});
	stats.polygons_after_csg = (uint32_t)csg[0].polygons.size();

	// build_leaf_bsp_tree (bsp_builder.rs:66-129)
	Vec3f bbox_min{1.0e8f, 1.0e8f, 1.0e8f}, bbox_max{-1.0e8f, -1.0e8f, -1.0e8f}; // build_bounding_box :584-602
	for (const MapPolygon& p : csg[0].polygons)
		for (const Vec3f& v : p.vertices)
			for (int a = 0; a < 3; ++a)
			{
				bbox_min.at(a) = std::fmin(bbox_min[a], v[a]);
				bbox_max.at(a) = std::fmax(bbox_max[a], v[a]);
			}
	for (int a = 0; a < 3; ++a)
	{
		bbox_min.at(a) -= 128.0f;
		bbox_max.at(a) += 128.0f;
	}
	TreeChild root = build_leaf_bsp_tree_r(preprocess_input_polygons(csg[0].polygons, materials), options.advanced_splitter_plane_selection);
	std::vector<LeafsPortal> portals = build_portals(root, bbox_min, bbox_max, materials);
	for (size_t i = 0; i < portals.size(); ++i) // set_leafs_portals :1018-1027
	{
		portals[i].leaf_front->portals.push_back(i);
		portals[i].leaf_back->portals.push_back(i);
	}
	// collect_reachable_leafs (:1029-1126): flood from the leafs of the entities' origins through the portals
	{
		std::vector<TreeLeaf*> todo;
		for (size_t e = 1; e < csg.size(); ++e)
			for (const auto& kv : csg[e].keys)
				if (kv.first == "origin")
				{
					Vec3f origin;
					if (parse_vec3(kv.second, origin))
						todo.push_back(get_leaf_for_point(root, origin));
				}
		while (!todo.empty())
		{
			TreeLeaf* leaf = todo.back();
			todo.pop_back();
			if (leaf->reachable)
				continue;
			leaf->reachable = true;
			for (size_t pi : leaf->portals)
			{
				todo.push_back(portals[pi].leaf_front);
				todo.push_back(portals[pi].leaf_back);
			}
		}
	}
	// remove_unreachable_portals (:1182-1198) first (it needs the leafs alive), then the leafs (:1128-1180)
	{
		std::vector<size_t> remap(portals.size(), (size_t)-1);
		std::vector<LeafsPortal> kept;
		for (size_t i = 0; i < portals.size(); ++i)
			if (portals[i].leaf_front->reachable && portals[i].leaf_back->reachable)
			{
				remap[i] = kept.size();
				kept.push_back(std::move(portals[i]));
			}
		// remove_expired_portals_from_leafs_r (:1200-1222)
		std::vector<SplitterRef> stack;
		std::vector<std::pair<TreeLeaf*, std::vector<SplitterRef>>> leafs;
		std::vector<const TreeNode*> node_order;
		collect_leafs_r(root, stack, leafs, node_order);
		for (auto& lf : leafs)
		{
			std::vector<size_t> np;
			for (size_t pi : lf.first->portals)
				if (remap[pi] != (size_t)-1)
					np.push_back(remap[pi]);
			lf.first->portals = std::move(np);
		}
		portals = std::move(kept);
	}
	const bool root_reachable = remove_unreachable_leafs_r(root, stats.leafs_initial, stats.leafs_removed);
	if (!root_reachable)
		log += "Warning, root node is unreachable!\n";
	stats.tree_depth = (uint32_t)get_tree_max_depth_r(root);
	if (stats.tree_depth >= 72) // MAX_BSP_TREE_DEPTH :54
		log += "Warning, BSP tree max depth " + std::to_string(stats.tree_depth) + " is greater than limit 72\n";

	// submodels (main.rs:104-108)
	std::vector<std::unique_ptr<SubmodelNode>> submodel_trees;
	for (size_t e = 1; e < csg.size(); ++e)
	{
		std::vector<MapPolygon> pre = preprocess_input_polygons(csg[e].polygons, materials);
		if (pre.empty())
			submodel_trees.emplace_back(new SubmodelNode{Plane{Vec3f{0, 0, 0}, 0.0f}, {}, {}});
		else
			submodel_trees.push_back(build_submodel_bsp_tree_r(std::move(pre)));
	}

	// convert_bsp_map_to_compact_format (bsp_map_compact_conversion.rs:4-40)
	Converter cv{out_map, materials, {}, {}, {}};
	for (const LeafsPortal& p : portals) // :44-75
	{
		Portal cp;
		cp.first_vertex = (uint32_t)out_map.vertices.size();
		cp.num_vertices = (uint32_t)p.vertices.size();
		cp.leafs[0] = cp.leafs[1] = 0;
		cp.plane = p.plane;
		out_map.vertices.insert(out_map.vertices.end(), p.vertices.begin(), p.vertices.end());
		out_map.portals.push_back(cp);
	}
	if (root_reachable)
		cv.convert_child(root, portals);
	for (size_t i = 0; i < portals.size(); ++i) // fill_portals_leafs :272-293
	{
		out_map.portals[i].leafs[0] = portals[i].leaf_front->index;
		out_map.portals[i].leafs[1] = portals[i].leaf_back->index;
	}
	for (size_t e = 0; e < csg.size(); ++e) // convert_entities_to_compact_format :316-349
	{
		uint32_t submodel_index = ~0u;
		if (e != 0)
		{
			const SubmodelNode& tree = *submodel_trees[e - 1];
			if (!tree.polygons.empty() || tree.children[0] || tree.children[1])
			{
				submodel_index = (uint32_t)out_map.submodels.size();
				Submodel sm;
				sm.first_polygon = (uint32_t)out_map.polygons.size();
				sm.root_node = cv.convert_submodel_node_r(tree);
				sm.num_polygons = (uint32_t)out_map.polygons.size() - sm.first_polygon;
				out_map.submodels.push_back(sm);
			}
		}
		Entity ent;
		ent.first_key_value_pair = (uint32_t)out_map.key_value_pairs.size();
		for (const auto& kv : csg[e].keys)
		{
			KeyValuePair pair;
			pair.key = cv.convert_string(kv.first);
			pair.value = cv.convert_string(kv.second);
			out_map.key_value_pairs.push_back(pair);
		}
		ent.num_key_value_pairs = (uint32_t)csg[e].keys.size();
		ent.submodel_index = submodel_index;
		out_map.entities.push_back(ent);
	}
	out_map.textures.resize(cv.texture_names.size()); // fill_textures :295-306
	for (size_t i = 0; i < cv.texture_names.size(); ++i)
	{
		std::memset(out_map.textures[i].name, 0, MAX_TEXTURE_NAME_LEN);
		std::memcpy(out_map.textures[i].name, cv.texture_names[i].data(), std::min(cv.texture_names[i].size(), MAX_TEXTURE_NAME_LEN - 1));
	}
	build_dummy_lightmaps(materials, out_map);
	stats.portals = (uint32_t)out_map.portals.size();
	stats.leafs = (uint32_t)out_map.leafs.size();
	stats.nodes = (uint32_t)out_map.nodes.size();
	stats.polygons = (uint32_t)out_map.polygons.size();
	stats.submodels = (uint32_t)out_map.submodels.size();
	stats.lightmap_texels = (uint32_t)(out_map.lightmaps_data.size() / 3);
	return true;
}

const MaterialDef* material_of_texture(const MaterialTable& materials, const TextureName& t)
{
	return find_material(materials, std::string(t.name, strnlen(t.name, MAX_TEXTURE_NAME_LEN)));
}

// lightmaps_builder.rs:299-330: offsets of the per-polygon lightmaps; element 0 is the "no lightmap" marker
size_t allocate_lightmaps(const MaterialTable& materials, BSPMap& map)
{
	size_t offset = 1;
	for (Polygon& p : map.polygons)
	{
		const MaterialDef* m = material_of_texture(materials, map.textures[p.texture]);
		if (!m || m->light)
		{
			p.lightmap_data_offset = (uint32_t)offset;
			offset += (size_t)get_lightmap_size(p.tex_coord_min[0], p.tex_coord_max[0]) * get_lightmap_size(p.tex_coord_min[1], p.tex_coord_max[1]);
		}
		else
			p.lightmap_data_offset = 0;
	}
	return offset;
}

void build_dummy_lightmaps(const MaterialTable& materials, BSPMap& map) // lightmaps_builder.rs:216-225
{
	const size_t n = allocate_lightmaps(materials, map);
	map.lightmaps_data.assign(n * 3, 1.0f);
	map.directional_lightmaps_data.clear();
}

// One line per material: name bsp twosided blocks_view light shadow draw decals skybox blending normal_map roughness roughness_map
// is_metal emissive_r emissive_g emissive_b [emissive_layer layer_r layer_g layer_b] (written by tools/map_compiler.py from the reference's materials/*.json, material.rs:7-98)
MaterialTable parse_material_table(const std::string& text)
{
	// (no iostreams here: the library may be linked with a static libstdc++ next to a process-wide shared one)
	MaterialTable table;
	size_t pos = 0;
	while (pos < text.size())
	{
		size_t eol = text.find('\n', pos);
		if (eol == std::string::npos)
			eol = text.size();
		std::vector<std::string> tok;
		size_t i = pos;
		while (i < eol)
		{
			while (i < eol && (text[i] == ' ' || text[i] == '\t' || text[i] == '\r'))
				++i;
			size_t j = i;
			while (j < eol && !(text[j] == ' ' || text[j] == '\t' || text[j] == '\r'))
				++j;
			if (j > i)
				tok.push_back(text.substr(i, j - i));
			i = j;
		}
		pos = eol + 1;
		if (tok.size() < 17)
			continue;
		MaterialDef m;
		m.name = tok[0];
		auto flag = [&](size_t k) { return std::strtol(tok[k].c_str(), nullptr, 10) != 0; };
		m.bsp = flag(1);
		m.twosided = flag(2);
		m.blocks_view = flag(3);
		m.light = flag(4);
		m.shadow = flag(5);
		m.draw = flag(6);
		m.decals = flag(7);
		m.skybox = flag(8);
		m.blending_mode = (int)std::strtol(tok[9].c_str(), nullptr, 10);
		m.normal_map = flag(10);
		m.roughness = std::strtof(tok[11].c_str(), nullptr);
		m.roughness_map = flag(12);
		m.is_metal = flag(13);
		for (int c = 0; c < 3; ++c)
			m.emissive_light[c] = std::strtof(tok[14 + (size_t)c].c_str(), nullptr);
		if (tok.size() >= 21)
		{
			m.emissive_layer = flag(17);
			for (int c = 0; c < 3; ++c)
				m.emissive_layer_light[c] = std::strtof(tok[18 + (size_t)c].c_str(), nullptr);
		}
		table[m.name] = m;
	}
	return table;
}

} // namespace mapc
} // namespace sqwh
