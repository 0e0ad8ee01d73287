// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).
//
// CPU restatement of the shadow-map depth rasteriser: DepthRasterizer::fill_polygon / fill_polygon_part
// (src/square_wheel/rasterizer.rs:1295-1474) and make_shadow_map_circle (src/square_wheel/light.rs:133-157), driven the way
// DepthRenderer::draw_map (src/square_wheel/depth_renderer.rs:47-67) drives them: world polygons back to front without depth
// test, submodel polygons afterwards with DEPTH_TEST = true.
#include "oracle_common.h"

namespace oracle
{

struct DepthSide
{
	Fixed16 x_start, dx_dy;
};

// rasterizer.rs:1428-1473
template <bool DEPTH_TEST>
static void depth_fill_polygon_part(float* depth_buffer, int32_t width, int32_t height, Fixed16 y_start, Fixed16 y_end, DepthSide left_side, DepthSide right_side, const float de[3])
{
	const int32_t y_start_int = std::max(fixed16_round_to_int(y_start), 0);
	const int32_t y_end_int = std::min(fixed16_round_to_int(y_end), height);
	const Fixed16 y_start_delta = wsub(wadd(int_to_fixed16(y_start_int), FIXED16_HALF), y_start);
	Fixed16 x_left = wadd(wadd(left_side.x_start, fixed16_mul(y_start_delta, left_side.dx_dy)), FIXED16_HALF);
	Fixed16 x_right = wadd(wadd(right_side.x_start, fixed16_mul(y_start_delta, right_side.dx_dy)), FIXED16_HALF);
	const float d_inv_z_dx = de[0], d_inv_z_dy = de[1], k = de[2];
	float line_inv_z = (float)y_start_int * d_inv_z_dy + 0.5f * (d_inv_z_dx + d_inv_z_dy) + k;

	for (int32_t y_int = y_start_int; y_int < y_end_int; ++y_int)
	{
		const int32_t x_start_int = std::max(fixed16_floor_to_int(x_left), 0);
		const int32_t x_end_int = std::min(fixed16_floor_to_int(x_right), width);
		if (x_start_int < x_end_int)
		{
			float* line_dst = depth_buffer + (size_t)y_int * width;
			float inv_z = line_inv_z + (float)x_start_int * d_inv_z_dx;
			for (int32_t x = x_start_int; x < x_end_int; ++x)
			{
				if (DEPTH_TEST)
					line_dst[x] = std::fmax(inv_z, line_dst[x]); // f32::max
				else
					line_dst[x] = inv_z;
				inv_z += d_inv_z_dx;
			}
		}
		x_left = wadd(x_left, left_side.dx_dy);
		x_right = wadd(x_right, right_side.dx_dy);
		line_inv_z += d_inv_z_dy;
	}
}

// rasterizer.rs:1314-1426
template <bool DEPTH_TEST>
static void depth_fill_polygon(float* depth_buffer, int32_t width, int32_t height, const sqw_polygon_point* vertices, uint32_t num_vertices, const float de[3])
{
	uint32_t lower_vertex_index = 0;
	Fixed16 min_y = vertices[0].y;
	for (uint32_t index = 0; index < num_vertices; ++index)
		if (vertices[index].y < min_y)
		{
			min_y = vertices[index].y;
			lower_vertex_index = index;
		}

	uint32_t left_index = lower_vertex_index, right_index = lower_vertex_index;
	Fixed16 cur_y = min_y;
	for (;;)
	{
		uint32_t next_left_index = left_index + num_vertices - 1;
		if (next_left_index >= num_vertices)
			next_left_index -= num_vertices;
		uint32_t next_right_index = right_index + 1;
		if (next_right_index >= num_vertices)
			next_right_index -= num_vertices;

		const sqw_polygon_point& left_vertex = vertices[left_index];
		const sqw_polygon_point& next_left_vertex = vertices[next_left_index];
		const sqw_polygon_point& right_vertex = vertices[right_index];
		const sqw_polygon_point& next_right_vertex = vertices[next_right_index];

		const Fixed16 dy_left = wsub(next_left_vertex.y, left_vertex.y);
		const Fixed16 dy_right = wsub(next_right_vertex.y, right_vertex.y);
		const Fixed16 next_y = std::min(next_left_vertex.y, next_right_vertex.y);
		if (dy_left > FIXED16_HALF && dy_right > FIXED16_HALF)
		{
			const Fixed16 dx_dy_left = fixed16_div(wsub(next_left_vertex.x, left_vertex.x), dy_left);
			const Fixed16 dx_dy_right = fixed16_div(wsub(next_right_vertex.x, right_vertex.x), dy_right);
			depth_fill_polygon_part<DEPTH_TEST>(
				depth_buffer, width, height, cur_y, next_y, DepthSide{wadd(left_vertex.x, fixed16_mul(dx_dy_left, wsub(cur_y, left_vertex.y))), dx_dy_left},
				DepthSide{wadd(right_vertex.x, fixed16_mul(dx_dy_right, wsub(cur_y, right_vertex.y))), dx_dy_right}, de);
		}
		else if (dy_left > 0 && dy_right > 0)
		{
			const int32_t cur_y_int = fixed16_round_to_int(cur_y);
			const int32_t next_y_int = fixed16_round_to_int(next_y);
			if (cur_y_int < next_y_int)
			{
				const Fixed16 thin_line_y = wadd(int_to_fixed16(cur_y_int), FIXED16_HALF);
				const Fixed16 x_start_left = wadd(left_vertex.x, fixed16_mul_div(wsub(thin_line_y, left_vertex.y), wsub(next_left_vertex.x, left_vertex.x), dy_left));
				const Fixed16 x_start_right = wadd(right_vertex.x, fixed16_mul_div(wsub(thin_line_y, right_vertex.y), wsub(next_right_vertex.x, right_vertex.x), dy_right));
				depth_fill_polygon_part<DEPTH_TEST>(depth_buffer, width, height, cur_y, next_y, DepthSide{x_start_left, 0}, DepthSide{x_start_right, 0}, de);
			}
		}

		if (next_left_index == next_right_index)
			break;
		if (vertices[next_right_index].y < vertices[next_left_index].y)
			right_index = next_right_index;
		else
			left_index = next_left_index;
		cur_y = next_y;
	}
}

// light.rs:133-157
static void make_shadow_map_circle(float* data, uint32_t size)
{
	const float inf = 1.0e24f;
	const int32_t size_signed = (int32_t)size;
	const int32_t radius = size_signed / 2;
	for (int32_t y = 0; y < size_signed; ++y)
	{
		const int32_t dy = y - radius;
		const int32_t dx = radius - std::min((int32_t)std::ceil(std::sqrt((float)std::max(radius * radius - dy * dy, 0))), radius);
		float* line = data + (size_t)y * size_signed;
		for (int32_t x = 0; x < dx; ++x)
			line[x] = inf;
		for (int32_t x = size_signed - dx; x < size_signed; ++x)
			line[x] = inf;
	}
}

} // namespace oracle

// Render every map of the call into `shadow_maps` (floats).  Same inputs as sqw_frame_set_shadow_map_draw.
extern "C" int oracle_draw_shadow_maps(
	float* shadow_maps, const sqw_depth_map* maps, uint32_t num_maps, const sqw_depth_cmd* cmds, uint32_t num_cmds, const sqw_polygon_point* points)
{
	for (uint32_t m = 0; m < num_maps; ++m)
	{
		const sqw_depth_map& map = maps[m];
		float* buf = shadow_maps + map.offset;
		if (map.flags & SQW_DEPTH_MAP_CLEAR)
			for (size_t i = 0; i < (size_t)map.width * map.height; ++i)
				buf[i] = map.clear_value;
		for (uint32_t ci = map.first_cmd; ci < map.first_cmd + map.num_cmds && ci < num_cmds; ++ci)
		{
			const sqw_depth_cmd& c = cmds[ci];
			if (c.depth_test)
				oracle::depth_fill_polygon<true>(buf, (int32_t)map.width, (int32_t)map.height, points + c.vertex_offset, c.num_vertices, c.depth_eq);
			else
				oracle::depth_fill_polygon<false>(buf, (int32_t)map.width, (int32_t)map.height, points + c.vertex_offset, c.num_vertices, c.depth_eq);
		}
		if ((map.flags & SQW_DEPTH_MAP_PROJECTOR_CIRCLE) && map.width == map.height)
			oracle::make_shadow_map_circle(buf, map.width);
	}
	return 0;
}
