// Edge walk shared by the colour rasteriser (Rasterizer::fill_polygon, src/square_wheel/rasterizer.rs:160-270) and the depth
// rasteriser (DepthRasterizer::fill_polygon, :1314-1426): both split a convex clockwise polygon into trapezoids between
// successive vertex y's with identical code, so one closed-form search serves both.
#pragma once
#include "sqw_device.cuh"
#include "../../include/sqw_cuda.h"

namespace sqw
{

struct PolygonSide
{
	Fixed16 x_start, dx_dy;
};

// The trapezoid [y_start, y_end) handed to fill_polygon_part* that owns a scan-line.
struct PolygonPart
{
	Fixed16 y_start, y_end;
	int32_t y_start_int; // max(round(y_start), 0)
	PolygonSide left, right;
};

// Re-walks the left / right vertex chains from the min-y vertex exactly like the reference's loop and stops at the
// trapezoid whose clipped row range [max(round(y_start), 0), min(round(y_end), clip_h)) contains `row`.
// Returns false when no trapezoid owns the row (thin parts the reference skips, rows outside the polygon).
__device__ inline bool find_polygon_part(const sqw_polygon_point* __restrict__ vertices, uint32_t num_vertices, int32_t row, int32_t clip_h, PolygonPart& out)
{
	uint32_t lower_vertex_index = 0;
	Fixed16 min_y = vertices[0].y;
	for (uint32_t index = 1; index < num_vertices; ++index)
	{
		const Fixed16 y = vertices[index].y;
		if (y < min_y)
		{
			min_y = y;
			lower_vertex_index = index;
		}
	}

	uint32_t left_index = lower_vertex_index;
	uint32_t right_index = lower_vertex_index;
	Fixed16 cur_y = min_y;
	for (;;)
	{
		uint32_t next_left_index = left_index + num_vertices - 1;
		if (next_left_index >= num_vertices)
			next_left_index -= num_vertices;
		uint32_t next_right_index = right_index + 1;
		if (next_right_index >= num_vertices)
			next_right_index -= num_vertices;

		const sqw_polygon_point left_vertex = vertices[left_index];
		const sqw_polygon_point next_left_vertex = vertices[next_left_index];
		const sqw_polygon_point right_vertex = vertices[right_index];
		const sqw_polygon_point next_right_vertex = vertices[next_right_index];

		const Fixed16 dy_left = wsub(next_left_vertex.y, left_vertex.y);
		const Fixed16 dy_right = wsub(next_right_vertex.y, right_vertex.y);
		const Fixed16 next_y = min(next_left_vertex.y, next_right_vertex.y);

		const int32_t y_start_int = max(fixed16_round_to_int(cur_y), 0);
		const int32_t y_end_int = min(fixed16_round_to_int(next_y), clip_h);
		const bool owns_row = row >= y_start_int && row < y_end_int;

		if (dy_left > SQW_FIXED16_HALF && dy_right > SQW_FIXED16_HALF)
		{
			if (owns_row)
			{
				const Fixed16 dx_dy_left = fixed16_div(wsub(next_left_vertex.x, left_vertex.x), dy_left);
				const Fixed16 dx_dy_right = fixed16_div(wsub(next_right_vertex.x, right_vertex.x), dy_right);
				out.y_start = cur_y;
				out.y_end = next_y;
				out.y_start_int = y_start_int;
				out.left.x_start = wadd(left_vertex.x, fixed16_mul(dx_dy_left, wsub(cur_y, left_vertex.y)));
				out.left.dx_dy = dx_dy_left;
				out.right.x_start = wadd(right_vertex.x, fixed16_mul(dx_dy_right, wsub(cur_y, right_vertex.y)));
				out.right.dx_dy = dx_dy_right;
				return true;
			}
		}
		else if (dy_left > 0 && dy_right > 0)
		{
			const int32_t cur_y_int = fixed16_round_to_int(cur_y);
			const int32_t next_y_int = fixed16_round_to_int(next_y);
			if (cur_y_int < next_y_int && owns_row)
			{
				// thin trapezoid: a single line evaluated at its centre (rasterizer.rs:226-252, :1380-1410)
				const Fixed16 thin_line_y = wadd(int_to_fixed16(cur_y_int), SQW_FIXED16_HALF);
				out.y_start = cur_y;
				out.y_end = next_y;
				out.y_start_int = y_start_int;
				out.left.x_start = wadd(left_vertex.x, fixed16_mul_div(wsub(thin_line_y, left_vertex.y), wsub(next_left_vertex.x, left_vertex.x), dy_left));
				out.left.dx_dy = 0;
				out.right.x_start = wadd(right_vertex.x, fixed16_mul_div(wsub(thin_line_y, right_vertex.y), wsub(next_right_vertex.x, right_vertex.x), dy_right));
				out.right.dx_dy = 0;
				return true;
			}
		}

		if (next_left_index == next_right_index)
			break;
		if (next_right_vertex.y < next_left_vertex.y)
			right_index = next_right_index;
		else
			left_index = next_left_index;
		cur_y = next_y;
	}
	return false;
}

} // namespace sqw
