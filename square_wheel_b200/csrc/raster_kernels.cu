// Stage 2 — span rasteriser on sm_100a.
//
// Replaces Rasterizer::fill_polygon / fill_triangle (src/square_wheel/rasterizer.rs:35-1292) as driven by
// PartialRenderer::perform_rasterization (partial_renderer.rs:1192-1283) with the single-threaded reference's
// clip rect (the whole surface).
//
// The reference is an incremental integer DDA painted back to front without a depth buffer.  To reproduce its
// pixels bit for bit on a GPU the incremental state is turned into closed forms with the same wrap-around
// arithmetic, and the painter's order is kept per pixel:
//
//   raster_setup_kernel   one THREAD per (command, scan-line), one warp per 32 scan-lines of one command: re-walks the polygon edges (or the triangle
//                         halves) to the trapezoid that owns the line, evaluates x_left/x_right for that line as
//                         x0 + n*dx_dy (mod 2^32, exactly the value the reference reaches after n additions) and
//                         does the per-line set-up of the reference's inner loops (i64 equations, span end-point
//                         clamps, the i32 divisions) -> one 48-byte SpanRec.
//   (binning)             tail CTAs of the same grid, one per 8-row band: order-preserving compaction (ballot + popc
//                         prefix) of the commands whose row range touches the band.
//   raster_tile_kernel    one CTA per 128x8 tile, one warp per row, 4 consecutive pixels per lane kept in
//                         registers; the warp walks the band's command list in submission order, filters it
//                         32 commands at a time against the tile, loads the SpanRec of its row (warp-uniform,
//                         one L1 transaction) and applies texel fetch + blend to the covered pixels.  The tile
//                         is written once with 128-bit stores — every framebuffer byte is written exactly once
//                         per frame no matter how much overdraw the painter's order has.
//
// Bound: HBM/L2 — per covered pixel one texel gather (C bytes, L2-resident surfaces) and per tile one C-byte
// store; the full-perspective mode adds one u32 divide per pixel (rasterizer.rs:430).
#include "sqw_device.cuh"
#include "sqw_internal.h"
#include "sqw_raster_walk.cuh"
#include <algorithm>
#include <cstdlib>

namespace sqw
{
namespace
{

// rasterizer.rs:1573-1580
constexpr int TC_SHIFT = 24;
constexpr int INV_Z_SHIFT = 29;
constexpr int INV_Z_PRE_SHIFT = 8;
constexpr int Z_CALC_SHIFT = 31;
constexpr int TC_FINAL_SHIFT = TC_SHIFT + Z_CALC_SHIFT - INV_Z_SHIFT + INV_Z_PRE_SHIFT; // 34
constexpr float TC_SCALE = 16777216.0f;
constexpr float INV_Z_SCALE = 536870912.0f;

__device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
__device__ __forceinline__ int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

struct Eq64
{
	int64_t d_inv_z_dx, d_inv_z_dy, inv_z_k;
	int64_t d_tc_dx[2], d_tc_dy[2], tc_k[2];
};
__device__ __forceinline__ Eq64 make_eq(const sqw_draw_cmd& c) // rasterizer.rs:287-310 / :472-495
{
	Eq64 e;
	const float dzx = c.depth_eq[0], dzy = c.depth_eq[1], dzk = c.depth_eq[2];
	e.d_inv_z_dx = f32_to_i64_trunc(INV_Z_SCALE * dzx);
	e.d_inv_z_dy = f32_to_i64_trunc(INV_Z_SCALE * dzy);
	e.inv_z_k = f32_to_i64_trunc(INV_Z_SCALE * (dzk + (dzx + dzy) * 0.5f));
#pragma unroll
	for (int i = 0; i < 2; ++i)
	{
		const float tx = c.tc_eq[i], ty = c.tc_eq[2 + i], tk = c.tc_eq[4 + i];
		e.d_tc_dx[i] = f32_to_i64_trunc(TC_SCALE * tx);
		e.d_tc_dy[i] = f32_to_i64_trunc(TC_SCALE * ty);
		e.tc_k[i] = f32_to_i64_trunc(TC_SCALE * (tk + (tx + ty) * 0.5f));
	}
	return e;
}

__device__ __forceinline__ void write_empty(SpanRec* out)
{
	uint4* o = reinterpret_cast<uint4*>(out);
	o[0] = make_uint4(0, 0, 0, 0);
	o[1] = make_uint4(0, 0, 0, 0);
	o[2] = make_uint4(0, 0, 0, 0);
}

// One scan-line `row` of one trapezoid [y_start, y_end) of a polygon: the body of the `for y_int` loop of
// fill_polygon_part / _line_z_corrected / _affine up to (not including) the pixel loop.
__device__ void polygon_part_line(
	const sqw_draw_cmd& c, const CmdDev& cd, int32_t row, int32_t y_start_int, Fixed16 y_start, Fixed16 y_end, PolygonSide left_side, PolygonSide right_side,
	int32_t clip_w, SpanRec* out)
{
	const int32_t n = row - y_start_int;
	const Fixed16 y_start_delta = wsub(wadd(int_to_fixed16(y_start_int), SQW_FIXED16_HALF), y_start);
	// x_left after n additions of dx_dy (rasterizer.rs:319-320, :448-449)
	const Fixed16 x_left = wadd(wadd(wadd(left_side.x_start, fixed16_mul(y_start_delta, left_side.dx_dy)), SQW_FIXED16_HALF), wmul(n, left_side.dx_dy));
	const Fixed16 x_right = wadd(wadd(wadd(right_side.x_start, fixed16_mul(y_start_delta, right_side.dx_dy)), SQW_FIXED16_HALF), wmul(n, right_side.dx_dy));
	const int32_t x_start_int = max(fixed16_floor_to_int(x_left), 0);
	const int32_t x_end_int = min(fixed16_floor_to_int(x_right), clip_w);
	if (x_start_int >= x_end_int)
	{
		write_empty(out);
		return;
	}

	SpanRec r;
	r.x0 = (int16_t)x_start_int;
	r.x1 = (int16_t)x_end_int;
#pragma unroll
	for (int i = 0; i < 11; ++i)
		r.v[i] = 0;

	const int64_t span_start_x = x_start_int;
	const int64_t span_end_x = x_end_int - 1;
	const int32_t span_length_minus_one = (x_end_int - 1) - x_start_int;

	if (cd.tc_mode == SQW_TC_AFFINE)
	{
		// rasterizer.rs:597-703
		const float dzx = c.depth_eq[0], dzy = c.depth_eq[1], dzk = c.depth_eq[2];
		const float y_start_f32 = fixed16_to_f32(y_start);
		const float y_end_f32 = fixed16_to_f32(y_end);
		const Fixed16 y_delta = wsub(y_end, y_start);
		const float x_start_left = fixed16_to_f32(left_side.x_start);
		const float x_end_left = fixed16_to_f32(wadd(left_side.x_start, fixed16_mul(y_delta, left_side.dx_dy)));
		const float x_start_right = fixed16_to_f32(right_side.x_start);
		const float x_end_right = fixed16_to_f32(wadd(right_side.x_start, fixed16_mul(y_delta, right_side.dx_dy)));
		const float z_start_left = 1.0f / (dzx * x_start_left + dzy * y_start_f32 + dzk);
		const float z_end_left = 1.0f / (dzx * x_end_left + dzy * y_end_f32 + dzk);
		const float z_start_right = 1.0f / (dzx * x_start_right + dzy * y_start_f32 + dzk);
		const float z_end_right = 1.0f / (dzx * x_end_right + dzy * y_end_f32 + dzk);
		const Fixed16 y_delta_for_tc = max(y_delta, SQW_FIXED16_ONE);
		const Fixed16 x_delta_for_tc = max(wsub(x_right, x_left), SQW_FIXED16_ONE);
#pragma unroll
		for (int i = 0; i < 2; ++i)
		{
			const Fixed16 max_tc = wsub(int_to_fixed16(c.tex_size[i]), 1);
			const float tx = c.tc_eq[i], ty = c.tc_eq[2 + i], tk = c.tc_eq[4 + i];
			const Fixed16 tc_start_left = min(max(f32_to_fixed16_unchecked(z_start_left * (x_start_left * tx + y_start_f32 * ty + tk)), 0), max_tc);
			const Fixed16 tc_end_left = min(max(f32_to_fixed16_unchecked(z_end_left * (x_end_left * tx + y_end_f32 * ty + tk)), 0), max_tc);
			const Fixed16 tc_start_right = min(max(f32_to_fixed16_unchecked(z_start_right * (x_start_right * tx + y_start_f32 * ty + tk)), 0), max_tc);
			const Fixed16 tc_end_right = min(max(f32_to_fixed16_unchecked(z_end_right * (x_end_right * tx + y_end_f32 * ty + tk)), 0), max_tc);
			const Fixed16 d_tc_left = fixed16_div(wsub(tc_end_left, tc_start_left), y_delta_for_tc);
			const Fixed16 d_tc_right = fixed16_div(wsub(tc_end_right, tc_start_right), y_delta_for_tc);
			// :680-681 + n row steps (:724-727)
			const Fixed16 tc_left = wadd(wadd(tc_start_left, fixed16_mul(y_start_delta, d_tc_left)), wmul(n, d_tc_left));
			const Fixed16 tc_right = wadd(wadd(tc_start_right, fixed16_mul(y_start_delta, d_tc_right)), wmul(n, d_tc_right));
			r.v[2 + i] = tc_left;
			r.v[4 + i] = fixed16_div(wsub(tc_right, tc_left), x_delta_for_tc);
		}
	}
	else
	{
		const Eq64 e = make_eq(c);
		const int64_t line_inv_z = (int64_t)row * e.d_inv_z_dy + e.inv_z_k; // :321 + n * d_inv_z_dy (:450)
		const int64_t line_tc[2] = {(int64_t)row * e.d_tc_dy[0] + e.tc_k[0], (int64_t)row * e.d_tc_dy[1] + e.tc_k[1]};
		const int64_t inv_z_min = 1 << INV_Z_PRE_SHIFT;
		const int64_t inv_z_max = 1 << 29;
		if (cd.tc_mode == SQW_TC_FULL_PERSPECTIVE)
		{
			// :345-422
			const int32_t span_inv_z_start = (int32_t)max64(inv_z_min, min64(span_start_x * e.d_inv_z_dx + line_inv_z, inv_z_max));
			const int32_t span_inv_z_end = (int32_t)max64(inv_z_min, min64(span_end_x * e.d_inv_z_dx + line_inv_z, inv_z_max));
			const int32_t span_d_inv_z = span_length_minus_one > 0 ? (span_inv_z_end - span_inv_z_start) / span_length_minus_one : 0;
			r.v[0] = span_inv_z_start;
			r.v[1] = span_d_inv_z;
#pragma unroll
			for (int i = 0; i < 2; ++i)
			{
				const int tc_max_shift = INV_Z_SHIFT - TC_SHIFT;
				const int64_t tc_max_max = 1 << 29;
				const int64_t span_inv_z_corrected = (int64_t)(span_inv_z_start - (1 << INV_Z_PRE_SHIFT));
				const int64_t tc_max_start = max64(0, min64(((span_inv_z_corrected * (int64_t)c.tex_size[i]) >> tc_max_shift) - 1, tc_max_max));
				const int32_t span_tc_start = (int32_t)max64(0, min64(span_start_x * e.d_tc_dx[i] + line_tc[i], tc_max_start));
				if (span_length_minus_one > 0)
				{
					const int64_t tc_max_end = max64(
						0, min64((((span_inv_z_corrected + (int64_t)wmul(span_d_inv_z, span_length_minus_one)) * (int64_t)c.tex_size[i]) >> tc_max_shift) - 1, tc_max_max));
					const int32_t span_tc_end = (int32_t)max64(0, min64(span_end_x * e.d_tc_dx[i] + line_tc[i], tc_max_end));
					const int32_t d = (span_tc_end - span_tc_start) / span_length_minus_one;
					r.v[4 + i] = d;
					r.v[2 + i] = span_tc_end > span_tc_start ? span_tc_start : wsub(span_tc_end, wmul(d, span_length_minus_one));
				}
				else
					r.v[2 + i] = span_tc_start;
			}
		}
		else
		{
			// line-z corrected, :532-556
			const uint32_t span_inv_z_start = (uint32_t)min64(max64(span_start_x * e.d_inv_z_dx + line_inv_z, inv_z_min), inv_z_max);
			const uint32_t span_inv_z_end = (uint32_t)min64(max64(span_end_x * e.d_inv_z_dx + line_inv_z, inv_z_min), inv_z_max);
			const uint64_t span_z_start = (1u << Z_CALC_SHIFT) / (span_inv_z_start >> INV_Z_PRE_SHIFT);
			const uint64_t span_z_end = (1u << Z_CALC_SHIFT) / (span_inv_z_end >> INV_Z_PRE_SHIFT);
#pragma unroll
			for (int i = 0; i < 2; ++i)
			{
				const uint64_t tc_start = (uint64_t)max64(line_tc[i] + span_start_x * e.d_tc_dx[i], 0);
				const uint64_t tc_end = (uint64_t)max64(line_tc[i] + span_end_x * e.d_tc_dx[i], 0);
				const uint64_t max_tc = ((uint64_t)(int64_t)c.tex_size[i] << 16) - 1;
				uint64_t a = (span_z_start * tc_start) >> (TC_FINAL_SHIFT - 16);
				uint64_t b = (span_z_end * tc_end) >> (TC_FINAL_SHIFT - 16);
				const int32_t tc_mul_z_start = (int32_t)(a < max_tc ? a : max_tc);
				const int32_t tc_mul_z_end = (int32_t)(b < max_tc ? b : max_tc);
				r.v[2 + i] = tc_mul_z_start;
				if (span_length_minus_one > 0)
					r.v[4 + i] = (tc_mul_z_end - tc_mul_z_start) / span_length_minus_one;
			}
		}
	}
	uint4* o = reinterpret_cast<uint4*>(out);
	const uint4* s = reinterpret_cast<const uint4*>(&r);
	o[0] = s[0];
	o[1] = s[1];
	o[2] = s[2];
}

// fill_polygon_impl_2_static_params (rasterizer.rs:160-270) for the line `row`: the edge walk lives in sqw_raster_walk.cuh.
__device__ void setup_polygon_line(const sqw_draw_cmd& c, const CmdDev& cd, const sqw_polygon_point* __restrict__ vertices, int32_t row, int32_t clip_w, int32_t clip_h, SpanRec* out)
{
	PolygonPart part;
	if (find_polygon_part(vertices, c.num_vertices, row, clip_h, part))
		polygon_part_line(c, cd, row, part.y_start_int, part.y_start, part.y_end, part.left, part.right, clip_w, out);
	else
		write_empty(out);
}

// ---- triangles -----------------------------------------------------------------------------------------------
// fill_triangle (rasterizer.rs:731-982) does almost all of its work once per triangle HALF, before the row loop: edge
// slopes, texture-coordinate and light gradients (two dozen 16.16 divisions) and the range-correction block
// (:1018-1192).  Only `x_left/x_right += dx_dy`, `tc_left += d_tc_dy_left`, `light_left += d_light_left` and the first
// pixel offset depend on the row.  So the set-up is split the same way: raster_triangle_setup_kernel runs FIVE LANES
// PER TRIANGLE (one per interpolant) and stores a TriRec (both halves, corrections applied); the per-row set-up in raster_setup_kernel reads
// the half that owns its row (a warp-uniform 96-byte broadcast) and evaluates the closed forms of the row loop.
struct TriPart // 96 bytes
{
	int32_t row0, row1;  // rows [row0, row1) this half owns (clipped)
	int32_t y_start_int; // first row of the half's row loop (:1000)
	Fixed16 x_left0, dx_dy_left, x_right0, dx_dy_right;
	Fixed16 tc_left[2], d_tc_dy_left[2], d_tc_dx[2];
	Fixed16 light_left[3], d_light_left[3], d_light_dx[3];
	int32_t pad[2];
};
struct TriRec
{
	TriPart part[2]; // lower, upper
};
static_assert(sizeof(TriPart) == 96 && sizeof(TriRec) == SQW_TRI_REC_BYTES, "TriRec layout");

// The row-independent part of fill_triangle_part (rasterizer.rs:984-1192), split by what it depends on so that the five
// interpolants of a triangle (two texture coordinates, three light components) can be set up by five lanes:
//   TriPartGeometry      rows and edges of the half, shared by every interpolant (:984-1016);
//   tc_channel_part      range correction of ONE texture coordinate (:1018-1138);
//   light_channel_part   underflow correction of ONE light component (:1140-1192).
struct TriPartGeometry
{
	int32_t y_start_int, row1;
	Fixed16 x_left0, x_right0, dx_dy_left, dx_dy_right;
	Fixed16 y_start_delta, y_end_delta, d_x_lower, d_x_upper;
};
__device__ __forceinline__ TriPartGeometry triangle_part_geometry(const Fixed16 y_start, const Fixed16 y_end, const PolygonSide left_side, const PolygonSide right_side, const int32_t clip_h)
{
	TriPartGeometry g;
	g.y_start_int = max(fixed16_round_to_int(y_start), 0);
	g.row1 = min(fixed16_round_to_int(y_end), clip_h);
	g.y_start_delta = wsub(wadd(int_to_fixed16(g.y_start_int), SQW_FIXED16_HALF), y_start);
	g.x_left0 = wadd(wadd(left_side.x_start, fixed16_mul(g.y_start_delta, left_side.dx_dy)), SQW_FIXED16_HALF);
	g.x_right0 = wadd(wadd(right_side.x_start, fixed16_mul(g.y_start_delta, right_side.dx_dy)), SQW_FIXED16_HALF);
	g.dx_dy_left = left_side.dx_dy;
	g.dx_dy_right = right_side.dx_dy;
	g.y_end_delta = wsub(y_end, wadd(y_start, g.y_start_delta));
	g.d_x_lower = max(wsub(g.x_right0, g.x_left0), 0);
	const Fixed16 x_upper_left = wadd(g.x_left0, fixed16_mul(g.y_end_delta, left_side.dx_dy));
	const Fixed16 x_upper_right = wadd(g.x_right0, fixed16_mul(g.y_end_delta, right_side.dx_dy));
	g.d_x_upper = max(wsub(x_upper_right, x_upper_left), 0);
	return g;
}

// :1018-1138 texture coordinate range correction of one coordinate -> {tc_left, d_tc_dy_left, d_tc_dx}
__device__ __forceinline__ void tc_channel_part(const TriPartGeometry& g, const Fixed16 max_tc, const Fixed16 tc_start_left, Fixed16 d_tc_dy_left, Fixed16 d_tc_dx, Fixed16 (&out)[3])
{
	const Fixed16 y_start_delta = g.y_start_delta, y_end_delta = g.y_end_delta, d_x_lower = g.d_x_lower, d_x_upper = g.d_x_upper;
	Fixed16 tc_left = min(max(wadd(tc_start_left, fixed16_mul(y_start_delta, d_tc_dy_left)), 0), max_tc);
	{
		const Fixed16 tc_lower_delta = fixed16_mul(d_tc_dx, d_x_lower);
		if (tc_lower_delta > max_tc)
		{
			tc_left = 0;
			d_tc_dx = fixed16_div(max_tc, d_x_lower);
		}
		else if (tc_lower_delta < -max_tc)
		{
			tc_left = max_tc;
			d_tc_dx = fixed16_div(-max_tc, d_x_lower);
		}
		else
		{
			const Fixed16 tc_lower_right = wadd(tc_left, tc_lower_delta);
			if (tc_lower_right > max_tc)
				tc_left = wsub(tc_left, wsub(tc_lower_right, max_tc));
			else if (tc_lower_right < 0)
				tc_left = wsub(tc_left, tc_lower_right);
		}
	}
	{
		Fixed16 tc_upper_left = min(max(wadd(tc_left, fixed16_mul(y_end_delta, d_tc_dy_left)), 0), max_tc);
		const Fixed16 tc_upper_delta = fixed16_mul(d_tc_dx, d_x_upper);
		if (tc_upper_delta > max_tc)
		{
			tc_upper_left = 0;
			d_tc_dx = fixed16_div(max_tc, d_x_upper);
		}
		else if (tc_upper_delta < -max_tc)
		{
			tc_upper_left = max_tc;
			d_tc_dx = fixed16_div(-max_tc, d_x_upper);
		}
		else
		{
			const Fixed16 tc_upper_right = wadd(tc_upper_left, tc_upper_delta);
			if (tc_upper_right > max_tc)
				tc_upper_left = wsub(tc_upper_left, wsub(tc_upper_right, max_tc));
			else if (tc_upper_right < 0)
				tc_upper_left = wsub(tc_upper_left, tc_upper_right);
		}
		d_tc_dy_left = y_end_delta > 0 ? fixed16_div(wsub(tc_upper_left, tc_left), y_end_delta) : 0;

		const Fixed16 tc_upper_left2 = wadd(tc_left, fixed16_mul(y_end_delta, d_tc_dy_left));
		const Fixed16 tc_upper_delta2 = fixed16_mul(d_tc_dx, d_x_upper);
		const Fixed16 tc_upper_right2 = wadd(tc_upper_left2, tc_upper_delta2);
		if (tc_upper_right2 < 0)
			d_tc_dx = d_x_upper > 0 ? fixed16_div(-tc_upper_left2, d_x_upper) : 0;
		else if (tc_upper_right2 > max_tc)
			d_tc_dx = d_x_upper > 0 ? fixed16_div(wsub(max_tc, tc_upper_left2), d_x_upper) : 0;
	}
	out[0] = tc_left;
	out[1] = d_tc_dy_left;
	out[2] = d_tc_dx;
}

// :1140-1192 light underflow correction of one component -> {light_left, d_light_left, d_light_dx}
__device__ __forceinline__ void light_channel_part(const TriPartGeometry& g, const Fixed16 light_start_left, Fixed16 d_light_left, Fixed16 d_light_dx, Fixed16 (&out)[3])
{
	const Fixed16 y_start_delta = g.y_start_delta, y_end_delta = g.y_end_delta, d_x_lower = g.d_x_lower, d_x_upper = g.d_x_upper;
	Fixed16 light_left = max(wadd(light_start_left, fixed16_mul(y_start_delta, d_light_left)), 0);
	const Fixed16 light_lower_right = wadd(light_left, fixed16_mul(d_x_lower, d_light_dx));
	if (light_lower_right < 0)
		light_left = wsub(light_left, light_lower_right);
	{
		Fixed16 light_upper_left = max(wadd(light_left, fixed16_mul(y_end_delta, d_light_left)), 0);
		const Fixed16 light_upper_right = wadd(light_upper_left, fixed16_mul(d_x_upper, d_light_dx));
		if (light_upper_right < 0)
			light_upper_left = wsub(light_upper_left, light_upper_right);
		d_light_left = y_end_delta > 0 ? fixed16_div(wsub(light_upper_left, light_left), y_end_delta) : 0;

		const Fixed16 light_upper_left2 = wadd(light_left, fixed16_mul(y_end_delta, d_light_left));
		const Fixed16 light_upper_right2 = wadd(light_upper_left2, fixed16_mul(d_x_upper, d_light_dx));
		if (light_upper_right2 < 0)
			d_light_dx = d_x_upper > 0 ? fixed16_div(-light_upper_left2, d_x_upper) : 0;
	}
	out[0] = light_left;
	out[1] = d_light_left;
	out[2] = d_light_dx;
}

// The row loop of fill_triangle_part (rasterizer.rs:1194-1291) in closed form for the line `row` of the half `p`.
__device__ __forceinline__ void triangle_row(const TriPart& p, int32_t row, int32_t clip_w, SpanRec* out)
{
	const int32_t n = row - p.y_start_int;
	const Fixed16 x_left = wadd(p.x_left0, wmul(n, p.dx_dy_left));
	const Fixed16 x_right = wadd(p.x_right0, wmul(n, p.dx_dy_right));
	const int32_t x_start_int = max(fixed16_floor_to_int(x_left), 0);
	const int32_t x_end_int = min(fixed16_floor_to_int(x_right), clip_w);
	if (x_start_int >= x_end_int)
	{
		write_empty(out);
		return;
	}
	const Fixed16 x_start_delta = wsub(wadd(int_to_fixed16(x_start_int), SQW_FIXED16_ONE), x_left); // :1202
	SpanRec r;
	r.x0 = (int16_t)x_start_int;
	r.x1 = (int16_t)x_end_int;
	r.v[10] = 0;
#pragma unroll
	for (int i = 0; i < 2; ++i)
	{
		const Fixed16 tc_left_row = wadd(p.tc_left[i], wmul(n, p.d_tc_dy_left[i]));  // :1285 applied n times
		r.v[0 + i] = wadd(tc_left_row, fixed16_mul(x_start_delta, p.d_tc_dx[i]));      // :1205-1208
		r.v[2 + i] = p.d_tc_dx[i];
	}
#pragma unroll
	for (int i = 0; i < 3; ++i)
	{
		const Fixed16 light_left_row = wadd(p.light_left[i], wmul(n, p.d_light_left[i])); // :1289 applied n times
		r.v[4 + i] = wadd(light_left_row, fixed16_mul(x_start_delta, p.d_light_dx[i]));     // :1210-1214
		r.v[7 + i] = p.d_light_dx[i];
	}
	uint4* o = reinterpret_cast<uint4*>(out);
	const uint4* s = reinterpret_cast<const uint4*>(&r);
	o[0] = s[0];
	o[1] = s[1];
	o[2] = s[2];
}

// fill_triangle_impl (rasterizer.rs:765-966) for ONE interpolant of one triangle: lane role `ch` 0-1 = texture coordinate, 2-4 = light
// component.  The geometry (vertex order, edge slopes, the two halves' rows) is recomputed by every lane of the triangle — it is a
// fraction of the work — so the lanes never communicate; lane 0 also stores the fields the interpolants share.  The channel only
// selects DATA (which vertex attribute, which output slot), never code, except for the two correction routines.
__device__ void triangle_channel_setup(const sqw_draw_cmd& c, const sqw_triangle_point* __restrict__ vertices, const int32_t clip_h, const int ch, TriRec* __restrict__ rec)
{
	const sqw_triangle_point v0 = vertices[0], v1 = vertices[1], v2 = vertices[2];
	int upper_index, lower_index;
	if (v0.y >= v1.y && v0.y >= v2.y)
	{
		upper_index = 0;
		lower_index = v1.y < v2.y ? 1 : 2;
	}
	else if (v1.y >= v0.y && v1.y >= v2.y)
	{
		upper_index = 1;
		lower_index = v0.y < v2.y ? 0 : 2;
	}
	else
	{
		upper_index = 2;
		lower_index = v0.y < v1.y ? 0 : 1;
	}
	const int middle_index = 3 - upper_index - lower_index;
	const sqw_triangle_point upper_vertex = upper_index == 0 ? v0 : (upper_index == 1 ? v1 : v2);
	const sqw_triangle_point lower_vertex = lower_index == 0 ? v0 : (lower_index == 1 ? v1 : v2);
	const sqw_triangle_point middle_vertex = middle_index == 0 ? v0 : (middle_index == 1 ? v1 : v2);
	// this lane's attribute at the three vertices
	auto attribute = [ch](const sqw_triangle_point& v) -> Fixed16 { return ch == 0 ? v.tc[0] : (ch == 1 ? v.tc[1] : (ch == 2 ? v.light[0] : (ch == 3 ? v.light[1] : v.light[2]))); };
	const Fixed16 a_upper = attribute(upper_vertex), a_lower = attribute(lower_vertex), a_middle = attribute(middle_vertex);

	const Fixed16 long_edge_dy = max(wsub(upper_vertex.y, lower_vertex.y), SQW_FIXED16_HALF);
	const Fixed16 lower_part_dy = max(wsub(middle_vertex.y, lower_vertex.y), SQW_FIXED16_HALF);
	const Fixed16 upper_part_dy = max(wsub(upper_vertex.y, middle_vertex.y), SQW_FIXED16_HALF);

	const Fixed16 long_edge_dx_dy = fixed16_div(wsub(upper_vertex.x, lower_vertex.x), long_edge_dy);
	const Fixed16 long_edge_x_in_middle = wadd(lower_vertex.x, fixed16_mul(long_edge_dx_dy, wsub(middle_vertex.y, lower_vertex.y)));
	const Fixed16 lower_part_dx_dy = fixed16_div(wsub(middle_vertex.x, lower_vertex.x), lower_part_dy);
	const Fixed16 upper_part_dx_dy = fixed16_div(wsub(upper_vertex.x, middle_vertex.x), upper_part_dy);

	const Fixed16 long_edge_d_dy = fixed16_div(wsub(a_upper, a_lower), long_edge_dy);
	const Fixed16 long_edge_in_middle = wadd(a_lower, fixed16_mul(long_edge_d_dy, lower_part_dy));
	const Fixed16 d_dy_lower = fixed16_div(wsub(a_middle, a_lower), lower_part_dy);
	const Fixed16 d_dy_upper = fixed16_div(wsub(a_upper, a_middle), upper_part_dy);

	// the two halves: {y range, left / right edge, start value and slope along the LEFT edge}
	const bool long_edge_right = long_edge_x_in_middle >= middle_vertex.x;
	Fixed16 d_dx;
	PolygonSide l[2], r[2];
	Fixed16 start[2], d_dy[2];
	if (long_edge_right)
	{
		const Fixed16 middle_dx = max(wsub(long_edge_x_in_middle, middle_vertex.x), SQW_FIXED16_HALF);
		d_dx = fixed16_div(wsub(long_edge_in_middle, a_middle), middle_dx);
		l[0].x_start = lower_vertex.x;
		l[0].dx_dy = lower_part_dx_dy;
		r[0].x_start = lower_vertex.x;
		r[0].dx_dy = long_edge_dx_dy;
		start[0] = a_lower;
		d_dy[0] = d_dy_lower;
		l[1].x_start = middle_vertex.x;
		l[1].dx_dy = upper_part_dx_dy;
		r[1].x_start = long_edge_x_in_middle;
		r[1].dx_dy = long_edge_dx_dy;
		start[1] = a_middle;
		d_dy[1] = d_dy_upper;
	}
	else
	{
		const Fixed16 middle_dx = max(wsub(middle_vertex.x, long_edge_x_in_middle), SQW_FIXED16_HALF);
		d_dx = fixed16_div(wsub(a_middle, long_edge_in_middle), middle_dx);
		l[0].x_start = lower_vertex.x;
		l[0].dx_dy = long_edge_dx_dy;
		r[0].x_start = lower_vertex.x;
		r[0].dx_dy = lower_part_dx_dy;
		start[0] = a_lower;
		d_dy[0] = long_edge_d_dy;
		l[1].x_start = long_edge_x_in_middle;
		l[1].dx_dy = long_edge_dx_dy;
		r[1].x_start = middle_vertex.x;
		r[1].dx_dy = upper_part_dx_dy;
		start[1] = long_edge_in_middle;
		d_dy[1] = long_edge_d_dy;
	}
	const Fixed16 y_range[3] = {lower_vertex.y, middle_vertex.y, upper_vertex.y};
#pragma unroll
	for (int half = 0; half < 2; ++half)
	{
		const TriPartGeometry g = triangle_part_geometry(y_range[half], y_range[half + 1], l[half], r[half], clip_h);
		TriPart* const p = &rec->part[half];
		if (ch == 0)
		{
			p->row0 = g.y_start_int;
			p->row1 = g.row1;
			p->y_start_int = g.y_start_int;
			p->x_left0 = g.x_left0;
			p->dx_dy_left = g.dx_dy_left;
			p->x_right0 = g.x_right0;
			p->dx_dy_right = g.dx_dy_right;
			p->pad[0] = p->pad[1] = 0;
		}
		Fixed16 out[3];
		if (ch < 2)
		{
			tc_channel_part(g, wsub(int_to_fixed16(ch == 0 ? c.tex_size[0] : c.tex_size[1]), 1), start[half], d_dy[half], d_dx, out);
			p->tc_left[ch] = out[0];
			p->d_tc_dy_left[ch] = out[1];
			p->d_tc_dx[ch] = out[2];
		}
		else
		{
			light_channel_part(g, start[half], d_dy[half], d_dx, out);
			p->light_left[ch - 2] = out[0];
			p->d_light_left[ch - 2] = out[1];
			p->d_light_dx[ch - 2] = out[2];
		}
	}
}

// Eight lanes per command, five of them at work: the TriRec of every triangle command, one interpolant per lane (polygon commands
// leave their slot untouched).  A triangle's set-up is a chain of two dozen 16.16 divisions; five lanes cut it to a third and
// four triangles per warp put four times as many warps on the machine as one thread per triangle did.
__global__ void __launch_bounds__(128) raster_triangle_setup_kernel(
	const sqw_draw_cmd* __restrict__ cmds, uint32_t num_cmds, const sqw_triangle_point* __restrict__ triangle_points, int32_t clip_h, TriRec* __restrict__ tri_recs)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i = t >> 3;
	const int ch = (int)(t & 7);
	if (i >= num_cmds || ch >= 5)
		return;
	const sqw_draw_cmd c = cmds[i];
	if (c.kind != SQW_CMD_TRIANGLE)
		return;
	triangle_channel_setup(c, triangle_points + c.vertex_offset, clip_h, ch, tri_recs + i);
}

// The line `row` of a triangle command: the half that owns the row (fill_triangle_part returns early when its clipped
// row range is empty, :1000-1005).
__device__ __forceinline__ void setup_triangle_line(const TriRec* __restrict__ rec, int32_t row, int32_t clip_w, SpanRec* out)
{
	const uint4* src = reinterpret_cast<const uint4*>(rec);
	const uint4 h0 = __ldg(src), h1 = __ldg(src + 6); // {row0, row1, ...} of the two halves
	int part;
	if (row >= (int32_t)h0.x && row < (int32_t)h0.y)
		part = 0;
	else if (row >= (int32_t)h1.x && row < (int32_t)h1.y)
		part = 1;
	else
	{
		write_empty(out);
		return;
	}
	TriPart p;
	uint4* d = reinterpret_cast<uint4*>(&p);
#pragma unroll
	for (int k = 0; k < 6; ++k)
		d[k] = __ldg(src + part * 6 + k);
	triangle_row(p, row, clip_w, out);
}

// ---- binning -------------------------------------------------------------------------------------------------
// One CTA per 8-row band, two steps:
//   1. the ordered list of the commands whose row range touches the band is built by the HOST (sqw_frame_set_draw walks those
//      row ranges anyway); the CTA stages it in shared memory (read from band_lists beyond SQW_BIN_CACHE entries);
//   2. every warp takes some of the band's 128-pixel columns and compacts, again in order, the band list entries whose
//      conservative column range touches the column into that tile's list.  Tile (band, col) owns the N slots starting at
//      band_offsets[band] * tiles_x + col * N (N = band list length), so no global prefix sum is needed; tile_desc gets
//      {first slot, count, tile origin}.  The tile kernel then reads ready-made ordered lists and never synchronises its CTA.
// Runs as the tail CTAs of the set-up grid (the two jobs are independent; one launch instead of two).
#define SQW_BIN_CACHE 1024
__device__ void bin_band(
	const CmdDev* __restrict__ cmd_dev, uint32_t num_cmds, const uint32_t* __restrict__ band_offsets, const uint32_t* __restrict__ band_lists, uint32_t* __restrict__ tile_lists,
	uint4* __restrict__ tile_desc, uint32_t* __restrict__ general_tiles, uint32_t tiles_x, int band)
{
	constexpr int WARPS = SQW_SETUP_THREADS / 32;
	__shared__ uint32_t s_idx[SQW_BIN_CACHE];
	__shared__ uint32_t s_bx[SQW_BIN_CACHE];
	__shared__ uint8_t s_cx[SQW_BIN_CACHE]; // 1: the command is not an opaque polygon (triangle, or a blend that reads the destination / the texel's alpha)
	const int32_t band_y0 = band * SQW_TILE_H;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t band_begin = __ldg(band_offsets + band);
	const uint32_t n = __ldg(band_offsets + band + 1) - band_begin;
	// 1. the band's ordered command list comes from the host (it walks every command's row range anyway to size the lists): stage
	//    the indices and the conservative column ranges of up to SQW_BIN_CACHE entries in shared memory — two dependent loads
	//    per entry, all entries independent of each other.
	for (uint32_t e = threadIdx.x; e < min(n, (uint32_t)SQW_BIN_CACHE); e += SQW_SETUP_THREADS)
	{
		const uint32_t idx = __ldg(band_lists + band_begin + e);
		s_idx[e] = idx;
		s_bx[e] = __ldg(reinterpret_cast<const uint32_t*>(cmd_dev + idx) + 1);
		s_cx[e] = (__ldg(reinterpret_cast<const uint32_t*>(cmd_dev + idx) + 3) & 0xFFFFu) != (uint32_t)SQW_CMD_POLYGON; // kind | blend << 8
	}
	__syncthreads();
	for (uint32_t col = (uint32_t)warp; col < tiles_x; col += WARPS)
	{
		const int32_t col_x0 = (int32_t)col * SQW_TILE_W, col_x1 = col_x0 + SQW_TILE_W;
		const uint32_t list_start = band_begin * tiles_x + col * n;
		uint32_t count = 0;
		bool complex = false;
		for (uint32_t base = 0; base < n; base += 32)
		{
			const uint32_t e = base + (uint32_t)lane;
			bool hit = false, cx = false;
			uint32_t idx = 0;
			if (e < n)
			{
				uint32_t bx;
				if (e < SQW_BIN_CACHE)
				{
					idx = s_idx[e];
					bx = s_bx[e];
					cx = s_cx[e] != 0;
				}
				else
				{
					idx = __ldg(band_lists + band_begin + e);
					bx = __ldg(reinterpret_cast<const uint32_t*>(cmd_dev + idx) + 1);
					cx = (__ldg(reinterpret_cast<const uint32_t*>(cmd_dev + idx) + 3) & 0xFFFFu) != (uint32_t)SQW_CMD_POLYGON;
				}
				const int32_t bx0 = (int16_t)(bx & 0xFFFF), bx1 = (int16_t)(bx >> 16);
				hit = bx0 < col_x1 && bx1 > col_x0;
			}
			const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
			complex |= __any_sync(0xFFFFFFFFu, hit && cx);
			if (hit)
				tile_lists[list_start + count + __popc(ballot & ((1u << lane) - 1))] = idx;
			count += __popc(ballot);
		}
		// .w: tile class.  SQW_TILE_CLASS_OPAQUE = every command of the tile is a polygon drawn with BlendingMode::None: only the LAST
		// command covering a pixel shows, which raster_tile_opaque_kernel exploits; everything else goes to raster_tile_kernel.
		if (lane == 0)
		{
			const uint32_t tile = (uint32_t)band * tiles_x + col;
			tile_desc[tile] = make_uint4(list_start, count, (uint32_t)col_x0 | ((uint32_t)band_y0 << 16), complex ? SQW_TILE_CLASS_GENERAL : SQW_TILE_CLASS_OPAQUE);
			// The general tiles of the frame, as a compact list for raster_tile_kernel (in a frame of world geometry they are a
			// few per cent of the tiles: that kernel no longer scans every descriptor to find them).  general_tiles[0] = count
			// (zeroed by launch_raster_setup), entries from [SQW_GENERAL_LIST_HEADER]; any order will do.
			if (complex)
				general_tiles[SQW_GENERAL_LIST_HEADER + atomicAdd(general_tiles, 1u)] = tile;
		}
	}
}

__global__ void __launch_bounds__(SQW_SETUP_THREADS) raster_setup_kernel(
	const sqw_draw_cmd* __restrict__ cmds, const CmdDev* __restrict__ cmd_dev, uint32_t num_cmds, const sqw_polygon_point* __restrict__ polygon_points,
	const TriRec* __restrict__ tri_recs, const uint2* __restrict__ chunks, uint32_t num_chunks, SpanRec* __restrict__ spans,
	int32_t clip_w, int32_t clip_h, uint32_t setup_blocks, const uint32_t* __restrict__ band_offsets, const uint32_t* __restrict__ band_lists, uint32_t* __restrict__ tile_lists,
	uint4* __restrict__ tile_desc, uint32_t* __restrict__ general_tiles, uint32_t tiles_x, uint32_t num_bands,
	unsigned long long* __restrict__ sums_to_zero, uint8_t* __restrict__ tonemap_lut, float isb, float isg, float isr)
{
	// Grid order: binning CTAs FIRST (each is one long dependent chain over the whole command list — started last they were the
	// tail of the kernel), then the tonemapping-table CTAs, then the span set-up CTAs.
	if (blockIdx.x < num_bands)
	{
		bin_band(cmd_dev, num_cmds, band_offsets, band_lists, tile_lists, tile_desc, general_tiles, tiles_x, (int)blockIdx.x);
		return;
	}
	const uint32_t lut_blocks = tonemap_lut ? (3 * SQW_TONEMAP_LUT_SIZE + SQW_SETUP_THREADS - 1) / SQW_SETUP_THREADS : 0;
	if (blockIdx.x < num_bands + lut_blocks)
	{
		// Tonemapping table of the fused tile epilogue: TonemappingFunction::do_it (postprocessor.rs:1019-1023) of every
		// channel value below SQW_TONEMAP_LUT_SIZE, computed with the same IEEE operations as tonemap_pixel.
		const uint32_t idx = (blockIdx.x - num_bands) * SQW_SETUP_THREADS + threadIdx.x;
		if (idx < 3 * SQW_TONEMAP_LUT_SIZE)
		{
			const uint32_t ch = idx / SQW_TONEMAP_LUT_SIZE;
			const float c = (float)(idx % SQW_TONEMAP_LUT_SIZE);
			const float inv_scale = ch == 0 ? isb : (ch == 1 ? isg : isr);
			tonemap_lut[idx] = (uint8_t)f32_to_u_clamped(__fdiv_rn(c, f32_mul_add(c, 1.0f / 255.0f, inv_scale)), 255.0f);
		}
		return;
	}
	const uint32_t setup_block = blockIdx.x - num_bands - lut_blocks;
	if (sums_to_zero && setup_block == 0) // colour sums of the fused tonemapping epilogue of the tile kernel that follows
		for (uint32_t i = threadIdx.x; i < SQW_SUM_WORDS; i += SQW_SETUP_THREADS)
			sums_to_zero[i] = 0;
	// One warp per chunk = up to 32 consecutive scan-lines of one command (chunk list built by the host while it computes
	// the row ranges): the command and its vertices are warp-uniform, so their loads are broadcasts served by one L1 line.
	const uint32_t chunk = setup_block * (SQW_SETUP_THREADS / 32) + (threadIdx.x >> 5);
	if (chunk >= num_chunks)
		return;
	const uint2 ch = __ldg(chunks + chunk); // {command index, first row of the chunk relative to the command's y0}
	const CmdDev cd = cmd_dev[ch.x];
	const uint32_t r = ch.y + (threadIdx.x & 31);
	if (r >= (uint32_t)((int32_t)cd.y1 - (int32_t)cd.y0))
		return;
	const uint32_t s = cd.span_base + r;
	const int32_t row = (int32_t)cd.y0 + (int32_t)(s - cd.span_base);
	if (cd.kind == SQW_CMD_POLYGON)
	{
		const sqw_draw_cmd c = cmds[ch.x];
		setup_polygon_line(c, cd, polygon_points + c.vertex_offset, row, clip_w, clip_h, spans + s);
	}
	else
		setup_triangle_line(tri_recs + ch.x, row, clip_w, spans + s);
}

// ---- tile kernel ----------------------------------------------------------------------------------------------
// z = 2^31 / d exactly as the u32 division of rasterizer.rs:430, for d in [1, 2^24) (inv_z >> 8 of a 32-bit value), in 13
// instructions instead of the ~20 of the generic u32 division: two correction steps from the hardware reciprocal.
//   q0 = trunc(2^31 * rcp(d) * (1 - 2^-21)) never exceeds the quotient and misses it by less than 2^11 + 1, so the
//   remainder 2^31 - q0*d is below 2^11*d + 2d: its quotient by d (same reciprocal, same downward bias) is at most one short.
// sqw_self_test compares it with the plain division for every d.
__device__ __forceinline__ uint32_t z_from_inv_z(uint32_t d)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_rz(d))); // d < 2^24: the conversion is exact
	r = __fmul_rz(r, 0x1.fffffp-1f);                                        // (1 - 2^-21): biased low, never above 1/d
	const uint32_t q0 = __float2uint_rz(__fmul_rz(r, 2147483648.0f));
	const uint32_t rem0 = (1u << Z_CALC_SHIFT) - q0 * d;
	const uint32_t k = __float2uint_rz(__fmul_rz(__uint2float_rz(rem0), r));
	const uint32_t rem1 = rem0 - k * d;
	return q0 + k + (rem1 >= d ? 1u : 0u);
}

// The same for d >= 2^11, where the first quotient misses by at most one: q - q0 < 2^31 / d * (2^-21 + 2^-22) + 1 < 2, so one
// correction step is enough (7 instructions fewer per pixel).  A polygon's largest inv_z is normalised into [2^23, 2^24)
// (draw_polygon, partial_renderer.rs:4106-4109), so d >= 2^11 whenever the depth range inside it is below 16 : 1; the opaque
// tile kernel checks the bound per warp and falls back to z_from_inv_z.  sqw_self_test compares it for every d >= 2^11.
#define SQW_Z_ONE_STEP_MIN_D 2048u
__device__ __forceinline__ uint32_t z_from_inv_z_one_step(uint32_t d)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_rz(d)));
	r = __fmul_rz(r, 0x1.fffffp-1f);
	const uint32_t q0 = __float2uint_rz(__fmul_rz(r, 2147483648.0f));
	const uint32_t rem0 = (1u << Z_CALC_SHIFT) - q0 * d;
	return q0 + (rem0 >= d ? 1u : 0u);
}

// 2^26 operand pairs for fixed16_div against the plain 64-bit division: a hash of the thread index picks both operands over the
// whole i32 range, with every eighth pair forced onto the edges (0, +-1, +-2^15, +-2^16, INT_MIN, INT_MAX, equal magnitudes).
__global__ void fixed16_div_test_kernel(unsigned long long* mismatches)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t h = i * 2654435761u + 0x9E3779B9u;
	h ^= h >> 16;
	h *= 0x85EBCA6Bu;
	h ^= h >> 13;
	uint32_t g = (i ^ 0xA5A5A5A5u) * 0xC2B2AE35u;
	g ^= g >> 15;
	g *= 0x27D4EB2Fu;
	g ^= g >> 16;
	int32_t x = (int32_t)h, y = (int32_t)g;
	const int32_t edges[12] = {0, 1, -1, 32768, -32768, 65536, -65536, INT32_MIN, INT32_MAX, 3, -3, 65535};
	if ((i & 7) == 0)
		x = edges[(i >> 3) % 12];
	if ((i & 7) == 1)
		y = edges[(i >> 3) % 12];
	if ((i & 7) == 2)
		y = (i & 8) ? x : -x;
	if ((i & 15) == 3)
		y >>= (h & 31); // small divisors: quotients far beyond 32 bits (the `as i32` wrap)
	if (y == 0)
		return;
	if (fixed16_div(x, y) != fixed16_div_reference(x, y))
		atomicAdd(mismatches, 1ull);
}

__global__ void z_from_inv_z_test_kernel(unsigned long long* mismatches)
{
	const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
	if (d >= 1 && d < (1u << 24) && z_from_inv_z(d) != (1u << Z_CALC_SHIFT) / d)
		atomicAdd(mismatches, 1ull);
	if (d >= SQW_Z_ONE_STEP_MIN_D && d < (1u << 24) && z_from_inv_z_one_step(d) != (1u << Z_CALC_SHIFT) / d)
		atomicAdd(mismatches, 1ull);
}

template <class C>
__device__ __forceinline__ typename C::Raw fetch_texel(const void* tex, uint32_t address)
{
	return __ldg(static_cast<const typename C::Raw*>(tex) + address);
}


#define SQW_PX_OFS(i) (((i) & 1) + 16 * ((i) >> 1))
#define SQW_NEED_DST(i)                       \
	if (!(have & (1u << (i))))                \
	{                                         \
		px[i] = dst[SQW_PX_OFS(i)];           \
		have |= 1u << (i);                    \
	}
// fill_triangle_part pixel loop (rasterizer.rs:1223-1279) of one command on the lane's pixels [I0, I1) of one row (0-1 = the pair in
// the left 16 pixels of the block, 2-3 = the pair in the right 16): small triangles usually touch one side only, and the
// caller skips the other (warp-uniform), which halves the shading of a model's triangles.
template <class C, int I0, int I1>
__device__ __forceinline__ void shade_triangle(
	const uint4 w0, const void* const tex, const uint32_t tex_width, const uint4 r0, const uint4 r1, const uint4 r2, const int32_t px0,
	const typename C::Raw* __restrict__ const dst, typename C::Raw (&px)[4], uint32_t& have)
{
	const int32_t x0 = (int16_t)(r0.x & 0xFFFF), x1 = (int16_t)(r0.x >> 16);
	const uint32_t blend = (w0.w >> 8) & 0xFF, tex64 = (w0.w >> 24) & 1;
	// fill_triangle_part pixel loop (rasterizer.rs:1223-1279), in three branch-free steps like the polygon path: the four
	// texel addresses, the four gathers in flight together, then light and blend with the (warp-uniform) blend mode
	// decided once outside the pixel loop.
	const int32_t tc0 = (int32_t)r0.y, tc1 = (int32_t)r0.z, dtc0 = (int32_t)r0.w, dtc1 = (int32_t)r1.x;
	const int32_t l0 = (int32_t)r1.y, l1 = (int32_t)r1.z, l2 = (int32_t)r1.w;
	const int32_t dl0 = (int32_t)r2.x, dl1 = (int32_t)r2.y, dl2 = (int32_t)r2.z;
	bool in[4];
	uint32_t address[4];
#pragma unroll
	for (int i = I0; i < I1; ++i)
	{
		const int32_t x = px0 + SQW_PX_OFS(i);
		in[i] = x >= x0 && x < x1;
		const int32_t k = x - x0;
		const uint32_t u = (uint32_t)(wadd(tc0, wmul(k, dtc0)) >> 16);
		const uint32_t v = (uint32_t)(wadd(tc1, wmul(k, dtc1)) >> 16);
		address[i] = u + v * tex_width; // garbage for uncovered pixels, never dereferenced
	}
	ColorVecI texel_vec[4];
	int32_t alpha[4];
	bool alpha_pass[4];
	if (tex64)
	{
		uint64_t t[4];
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			t[i] = 0;
			if (in[i])
				t[i] = __ldg(static_cast<const uint64_t*>(tex) + address[i]);
		}
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			texel_vec[i] = cvi_from_color64(t[i]);
			alpha[i] = Color64::get_alpha(t[i]);
			alpha_pass[i] = Color64::test_alpha(t[i]);
		}
	}
	else
	{
		uint32_t t[4];
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			t[i] = 0;
			if (in[i])
				t[i] = __ldg(static_cast<const uint32_t*>(tex) + address[i]);
		}
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			texel_vec[i] = cvi_from_color32(t[i]);
			alpha[i] = Color32::get_alpha(t[i]);
			alpha_pass[i] = Color32::test_alpha(t[i]);
		}
	}
	ColorVecI lighted[4];
#pragma unroll
	for (int i = I0; i < I1; ++i)
	{
		const int32_t k = px0 + SQW_PX_OFS(i) - x0;
		// line_light = from_color_i32x3([l0,l1,l2]) -> lanes [B,G,R,A] = [l2,l1,l0,0]
		const ColorVecI light = cvi(wadd(l2, wmul(k, dl2)), wadd(l1, wmul(k, dl1)), wadd(l0, wmul(k, dl0)), 0);
		lighted[i] = cvi_shr<16>(cvi_mul(texel_vec[i], light));
	}
	if (blend == SQW_BLEND_NONE)
	{
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			px[i] = in[i] ? C::from_veci(lighted[i]) : px[i];
			have |= in[i] ? (1u << i) : 0u;
		}
	}
	else if (blend == SQW_BLEND_ALPHA_TEST)
	{
#pragma unroll
		for (int i = I0; i < I1; ++i)
		{
			const bool pass = in[i] && alpha_pass[i];
			px[i] = pass ? C::from_veci(lighted[i]) : px[i];
			have |= pass ? (1u << i) : 0u;
		}
	}
	else
	{
#pragma unroll
		for (int i = I0; i < I1; ++i)
			if (in[i])
			{
				SQW_NEED_DST(i)
				if (blend == SQW_BLEND_AVERAGE)
					px[i] = C::average(px[i], C::from_veci(lighted[i]));
				else if (blend == SQW_BLEND_ADDITIVE)
					px[i] = C::saturated_sum(px[i], C::from_veci(lighted[i]));
				else
					px[i] = C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(C::to_veci(px[i]), 255 - alpha[i]), cvi_mul_scalar(lighted[i], alpha[i]))));
			}
	}
}

// One command on the four pixels x = px0 + {0, 1, 16, 17} of one row: the pixel loops of fill_polygon_part*
// (rasterizer.rs:424-445, :558-572, :705-719) or fill_triangle_part (:1223-1279) and write_into_framebuffer (:1511-1536).
//   w0 = CmdDev words {y0,y1 | bx0,bx1 | span_base | kind, blend, tc_mode, tex_flags}; r0..r2 = the SpanRec of the row.
// `have` bit i: px[i] holds the current value of the pixel; the framebuffer (dst) is only read for pixels whose first
// covering command blends with the destination.
template <class C>
__device__ __forceinline__ void shade_quad(
	const uint4 w0, const void* const tex, const uint32_t tex_width, const uint4 r0, const uint4 r1, const uint4 r2, const int32_t px0,
	const typename C::Raw* __restrict__ const dst, typename C::Raw (&px)[4], uint32_t& have)
{
	typedef typename C::Raw Raw;
	const int32_t x0 = (int16_t)(r0.x & 0xFFFF), x1 = (int16_t)(r0.x >> 16);
	const uint32_t kind = w0.w & 0xFF, blend = (w0.w >> 8) & 0xFF, tc_mode = (w0.w >> 16) & 0xFF;

	if (kind == SQW_CMD_POLYGON)
	{
		const int32_t tc0 = (int32_t)r0.w, tc1 = (int32_t)r1.x, dtc0 = (int32_t)r1.y, dtc1 = (int32_t)r1.z;
		const int32_t inv_z0 = (int32_t)r0.y, d_inv_z = (int32_t)r0.z;
		// The four pixels of the lane are evaluated WITHOUT branches (addresses of uncovered pixels are computed from
		// wrapped garbage and never dereferenced), so the four dependent chains — one u32 divide each in the full
		// perspective mode — interleave and the four texel gathers are in flight together.
		bool in[4];
		uint32_t address[4];
		if (tc_mode == SQW_TC_FULL_PERSPECTIVE)
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const int32_t x = px0 + SQW_PX_OFS(i);
				in[i] = x >= x0 && x < x1;
				const int32_t k = x - x0;
				const uint32_t span_inv_z = (uint32_t)wadd(inv_z0, wmul(k, d_inv_z));
				const uint32_t z = z_from_inv_z(max(span_inv_z >> INV_Z_PRE_SHIFT, 1u)); // rasterizer.rs:430 (covered pixels: >= 1 by the span clamp)
				const uint64_t t0 = (uint64_t)(int64_t)wadd(tc0, wmul(k, dtc0));
				const uint64_t t1 = (uint64_t)(int64_t)wadd(tc1, wmul(k, dtc1));
				const uint32_t u = (uint32_t)(((uint64_t)z * t0) >> TC_FINAL_SHIFT);
				const uint32_t v = (uint32_t)(((uint64_t)z * t1) >> TC_FINAL_SHIFT);
				address[i] = u + v * tex_width;
			}
		}
		else
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const int32_t x = px0 + SQW_PX_OFS(i);
				in[i] = x >= x0 && x < x1;
				const int32_t k = x - x0;
				const uint32_t u = (uint32_t)(wadd(tc0, wmul(k, dtc0)) >> 16);
				const uint32_t v = (uint32_t)(wadd(tc1, wmul(k, dtc1)) >> 16);
				address[i] = u + v * tex_width;
			}
		}
		Raw texel[4];
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			texel[i] = 0;
			if (in[i])
				texel[i] = fetch_texel<C>(tex, address[i]);
		}
		if (blend == SQW_BLEND_NONE)
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				px[i] = in[i] ? texel[i] : px[i];
				have |= in[i] ? (1u << i) : 0u;
			}
		}
		else if (blend == SQW_BLEND_ALPHA_TEST)
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const bool pass = in[i] && C::test_alpha(texel[i]);
				px[i] = pass ? texel[i] : px[i];
				have |= pass ? (1u << i) : 0u;
			}
		}
		else
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
				if (in[i])
				{
					SQW_NEED_DST(i)
					px[i] = blend_pixel<C>(blend, px[i], texel[i]);
				}
		}
	}
	// (triangle commands are shaded pair by pair: shade_triangle)
}

// The finished quad of one row: one 128-bit store per pixel pair, and with TONEMAP the 32-bit picture and colour sums.
template <class C, int TONEMAP>
__device__ __forceinline__ void finish_quad(
	typename C::Raw (&px)[4], uint32_t have, const int32_t px0, const int32_t width, const bool vec_ok, typename C::Raw* __restrict__ const dst,
	uint32_t* __restrict__ const o, const uint8_t* const s_lut, const float isb, const float isg, const float isr, uint32_t& sb, uint32_t& sg, uint32_t& sr)
{
	typedef typename C::Raw Raw;
	if (vec_ok && have == 0xF)
	{
		// the usual case: all four pixels are final and inside the picture
		if (sizeof(Raw) == 8)
		{
			*reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)px[0], (uint32_t)((uint64_t)px[0] >> 32), (uint32_t)px[1], (uint32_t)((uint64_t)px[1] >> 32));
			*reinterpret_cast<uint4*>(dst + 16) = make_uint4((uint32_t)px[2], (uint32_t)((uint64_t)px[2] >> 32), (uint32_t)px[3], (uint32_t)((uint64_t)px[3] >> 32));
		}
		else
		{
			*reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)px[0], (uint32_t)px[1]);
			*reinterpret_cast<uint2*>(dst + 16) = make_uint2((uint32_t)px[2], (uint32_t)px[3]);
		}
		if (TONEMAP)
		{
			// sizeof(Raw) == 8 here (the host only fuses 64-bit views)
			uint32_t out[4];
			const uint64_t any = (uint64_t)px[0] | (uint64_t)px[1] | (uint64_t)px[2] | (uint64_t)px[3];
			uint32_t cb[4], cg[4], cr[4];
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const uint64_t p = (uint64_t)px[i];
				cb[i] = (uint32_t)(p & 0xFFFF);
				cg[i] = (uint32_t)((p >> 16) & 0xFFFF);
				cr[i] = (uint32_t)((p >> 32) & 0xFFFF);
				sb += cb[i];
				sg += cg[i];
				sr += cr[i];
			}
			// one test for the twelve channels: no B, G or R bit at or above SQW_TONEMAP_LUT_SIZE
			if ((any & (0x0000FFFFFFFFFFFFull & ~(0x0001000100010001ull * (SQW_TONEMAP_LUT_SIZE - 1)))) == 0)
			{
#pragma unroll
				for (int i = 0; i < 4; ++i)
					out[i] = (uint32_t)s_lut[cb[i]] | ((uint32_t)s_lut[SQW_TONEMAP_LUT_SIZE + cg[i]] << 8) | ((uint32_t)s_lut[2 * SQW_TONEMAP_LUT_SIZE + cr[i]] << 16) | 0xFF000000u;
			}
			else
			{
#pragma unroll
				for (int i = 0; i < 4; ++i)
					out[i] = tonemap_pixel(cv_from_color64((uint64_t)px[i]), isb, isg, isr);
			}
			__stcs(reinterpret_cast<uint2*>(o), make_uint2(out[0], out[1]));
			__stcs(reinterpret_cast<uint2*>(o + 16), make_uint2(out[2], out[3]));
		}
	}
	else
	{
#pragma unroll
		for (int i = 0; i < 4; ++i)
			if (px0 + SQW_PX_OFS(i) < width && (have & (1u << i)))
				dst[SQW_PX_OFS(i)] = px[i];
		if (TONEMAP)
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
				if (px0 + SQW_PX_OFS(i) < width)
				{
					SQW_NEED_DST(i) // untouched pixel: tonemap what the buffer holds
					const uint64_t p = (uint64_t)px[i];
					sb += (uint32_t)(p & 0xFFFF);
					sg += (uint32_t)((p >> 16) & 0xFFFF);
					sr += (uint32_t)((p >> 32) & 0xFFFF);
					o[SQW_PX_OFS(i)] = tonemap_pixel(cv_from_color64(p), isb, isg, isr);
				}
		}
	}
}

// TONEMAP != 0 fuses Postprocessor::perform_tonemapping (postprocessor.rs:226-256) into the tile epilogue: the finished
// HDR pixels are still in registers, so the 32-bit picture and the colour sums are produced without reading the 64-bit
// buffer back (saves 8 of the 12 bytes per pixel of a separate tonemapping pass).
//
// Warp-autonomous tiles.  A CTA of 4 warps walks 128x8 tiles (persistent grid); warp w owns the 32x8 pixel block w of the
// tile and lane l eight pixels: four of block row l >> 3 and four of row 4 + (l >> 3), all kept in registers until the
// tile is finished, so each framebuffer byte is written once per frame however much overdraw the painter's order has.
// A compact footprint meets far fewer polygon edges than a 128x1 strip, and every global store is still a run of whole
// 32-byte sectors.  Nothing in the tile loop synchronises the CTA:
//   list     the ordered command list of the tile comes ready-made from the binning CTAs of the set-up grid; a lane takes
//            one entry, rejects it against the warp's block and the survivors are compacted IN LIST ORDER (ballot + popc)
//            into the warp's slice of shared memory, texture pointer resolved.  The descriptor and the first 32 entries
//            of the NEXT tile are fetched while the current one is evaluated.
//   spans    the span records of 4 listed commands x 8 block rows are fetched by the 32 lanes at once into the warp's
//            slice of shared memory (one memory latency per 4 commands), then the warp walks them in order: span records
//            of the lane's two rows from shared memory, pixels.  Only the texel gathers go to global memory.
template <class C, int TONEMAP>
__global__ void __launch_bounds__(SQW_RASTER_THREADS, SQW_TILE_CTAS_PER_SM) raster_tile_kernel(
	const CmdDev* __restrict__ cmd_dev, const SpanRec* __restrict__ spans, const uint32_t* __restrict__ tile_lists, const uint4* __restrict__ tile_desc,
	const char* __restrict__ surfaces_base, typename C::Raw* __restrict__ color_buffer, int32_t width, int32_t height, int32_t pitch, int clear,
	typename C::Raw clear_color, uint32_t* __restrict__ ldr, float isb, float isg, float isr, unsigned long long* __restrict__ sums,
	const uint8_t* __restrict__ tonemap_lut, uint32_t num_tiles, const uint32_t* __restrict__ general_tiles)
{
	typedef typename C::Raw Raw;
	constexpr int WARPS = SQW_RASTER_THREADS / 32;
	static_assert(WARPS * 32 == SQW_TILE_W && SQW_TILE_H == 8 && SQW_TILE_STAGE_HITS * SQW_TILE_H == 32, "tile geometry");
	// The frame's general tiles as a list written by the binning CTAs ([0] count, entries from [SQW_GENERAL_LIST_HEADER]; the count
	// is zeroed by launch_raster_setup before every binning).  A CTA with nothing to do leaves before it pays for anything (the
	// tonemapping table copy is 12 KB per CTA).
	const uint32_t num_general = __ldg(general_tiles);
	if (blockIdx.x >= num_general)
		return;
	__shared__ uint4 s_cmd_all[WARPS][2][32];                                // per warp: [0] CmdDev words 0-3, [1] {texture pointer, tex_width, tex_height}
	__shared__ uint4 s_span_all[WARPS][3][SQW_TILE_STAGE_HITS][SQW_TILE_H]; // per warp: SpanRec words of (staged command, block row)
	// Per-channel tonemapping table (exact results for channel values < SQW_TONEMAP_LUT_SIZE, built by the set-up grid of
	// this frame): three byte look-ups replace three IEEE divides for all but very bright pixels.
	__shared__ __align__(16) uint8_t s_lut[TONEMAP ? 3 * SQW_TONEMAP_LUT_SIZE : 16];
	// The table is copied with asynchronous copies (LDGSTS): the first tile's list, span records and texels are fetched and
	// evaluated while it arrives; the CTA waits for it once, before its first tonemapping epilogue.
	if (TONEMAP)
	{
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_lut);
		for (uint32_t i = threadIdx.x; i < 3 * SQW_TONEMAP_LUT_SIZE / 16; i += SQW_RASTER_THREADS)
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(tonemap_lut + (size_t)i * 16) : "memory");
		asm volatile("cp.async.commit_group;" ::: "memory");
	}
	bool lut_pending = TONEMAP != 0;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint4(*const s_cmd)[32] = s_cmd_all[warp];
	uint4(*const s_span)[SQW_TILE_STAGE_HITS][SQW_TILE_H] = s_span_all[warp];
	uint32_t sb = 0, sg = 0, sr = 0; // colour sums of the fused tonemapping (all tiles of this warp)

	// Persistent CTAs: the grid is sized to what the GPU holds at once and every CTA strides over the tiles, so the
	// per-CTA costs (tonemapping table, sum reduction, launch) are paid a few hundred times per frame, not once per tile.
	uint4 desc_next = __ldg(tile_desc + __ldg(general_tiles + SQW_GENERAL_LIST_HEADER + blockIdx.x));
	uint32_t idx_next = 0;
	if ((uint32_t)lane < desc_next.y)
		idx_next = __ldg(tile_lists + desc_next.x + lane);
	const int ra = lane >> 3, rb = 4 + (lane >> 3); // this lane's two rows inside the block
	for (uint32_t tile = blockIdx.x; tile < num_general; tile += gridDim.x) // (position in the list of general tiles)
	{
	const uint4 desc = desc_next; // {first entry in tile_lists, number of entries, tile_x0 | band_y0 << 16, -}
	uint32_t idx = idx_next;
	if (tile + gridDim.x < num_general)
		desc_next = __ldg(tile_desc + __ldg(general_tiles + SQW_GENERAL_LIST_HEADER + tile + gridDim.x));
	// this warp's block and this lane's pixels
	const int32_t blk_x0 = (int32_t)(desc.z & 0xFFFF) + warp * 32, blk_x1 = blk_x0 + 32;
	const int32_t blk_y0 = (int32_t)(desc.z >> 16), blk_y1 = min(blk_y0 + SQW_TILE_H, height);
	const int32_t row_a = blk_y0 + ra, row_b = blk_y0 + rb;
	// The lane's four pixels of a row are two PAIRS, 16 pixels apart: x = px0 + {0, 1, 16, 17}.  That way the 8 lanes of a
	// block row write 8 consecutive 16-byte chunks per store instruction (whole 32-byte sectors, 128 contiguous bytes per
	// row) instead of half-sectors with a 32-byte stride.
	const int32_t px0 = blk_x0 + (lane & 7) * 2;
	const uint32_t pix_a = (uint32_t)row_a * (uint32_t)pitch + (uint32_t)px0; // < 2^32: the view size is limited to 16384^2
	const uint32_t pix_b = pix_a + 4u * (uint32_t)pitch;
	Raw px_a[4], px_b[4];
	// `have` bit i: px[i] holds the current value of the pixel.  The framebuffer is only read for pixels whose first
	// covering command blends with the destination, and pixels no command touches are not written back: a frame whose
	// polygons cover the screen (the usual case inside a map, partial_renderer.rs:249-256) never reads the buffer.
	uint32_t have_a = clear ? 0xFu : 0u, have_b = have_a;
#pragma unroll
	for (int i = 0; i < 4; ++i)
		px_a[i] = px_b[i] = clear_color;
	const bool vec_ok = (blk_x1 <= width) && ((pitch & 1) == 0);

	if (blk_x0 < width) // (warp-uniform) blocks outside the picture have nothing to do
	for (uint32_t base = 0; base < desc.y; base += 32)
	{
		// ---- the warp's list: entries of the tile list that can touch this 32x8 block, in order ----
		if (base)
			idx = (base + lane < desc.y) ? __ldg(tile_lists + desc.x + base + lane) : 0u;
		bool hit = false;
		uint4 c0 = make_uint4(0, 0, 0, 0);
		const uint4* cdp = nullptr;
		if (base + lane < desc.y)
		{
			cdp = reinterpret_cast<const uint4*>(cmd_dev + idx);
			c0 = __ldg(cdp); // y0,y1 | bx0,bx1 | span_base | kind, blend, tc_mode, tex_flags
			const int32_t bx0 = (int16_t)(c0.y & 0xFFFF), bx1 = (int16_t)(c0.y >> 16);
			hit = bx0 < blk_x1 && bx1 > blk_x0; // the row range was tested by the binning CTA (block rows = band rows)
		}
		const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
		if (ballot == 0)
			continue;
		const uint32_t num_hits = __popc(ballot);
		if (hit)
		{
			uint4 c1 = __ldg(cdp + 1); // texture ptr | tex_width | tex_height
			// bit 1 of the last flag byte: `texture` is a byte offset into the view's surface pool (the pool may be
			// re-allocated between recording and submission, so surface references stay relative).
			if ((c0.w >> 25) & 1)
			{
				const uint64_t t = (uint64_t)reinterpret_cast<uintptr_t>(surfaces_base) + ((uint64_t)c1.x | ((uint64_t)c1.y << 32));
				c1.x = (uint32_t)t;
				c1.y = (uint32_t)(t >> 32);
			}
			const uint32_t slot = __popc(ballot & ((1u << lane) - 1));
			s_cmd[0][slot] = c0;
			s_cmd[1][slot] = c1;
		}
		__syncwarp();

		// ---- walk the warp's list in submission order, SQW_TILE_STAGE_HITS commands at a time ----
		for (uint32_t h0 = 0; h0 < num_hits; h0 += SQW_TILE_STAGE_HITS)
		{
			// Stage the span records of these commands for the 8 rows of the block: one (command, row) pair per lane, all
			// loads independent, so the warp pays ONE memory latency for the span records of 4 commands.
			{
				const uint32_t hs = (uint32_t)lane >> 3, rs = (uint32_t)lane & 7;
				if (h0 + hs < num_hits)
				{
					const uint4 w0 = s_cmd[0][h0 + hs];
					const int32_t y0 = (int16_t)(w0.x & 0xFFFF), y1 = (int16_t)(w0.x >> 16);
					const int32_t srow = blk_y0 + (int32_t)rs;
					uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0;
					if (srow >= y0 && srow < y1)
					{
						const uint4* sp = reinterpret_cast<const uint4*>(spans + w0.z + (uint32_t)(srow - y0));
						r0 = __ldg(sp);
						r1 = __ldg(sp + 1);
						if ((w0.w & 0xFF) == SQW_CMD_TRIANGLE)
							r2 = __ldg(sp + 2);
					}
					s_span[0][hs][rs] = r0;
					s_span[1][hs][rs] = r1;
					s_span[2][hs][rs] = r2;
				}
			}
			__syncwarp();
			const uint32_t h1 = min(num_hits - h0, (uint32_t)SQW_TILE_STAGE_HITS);
			for (uint32_t hh = 0; hh < h1; ++hh)
			{
				const uint4 ra0 = s_span[0][hh][ra], rb0 = s_span[0][hh][rb];
				const int32_t ax0 = (int16_t)(ra0.x & 0xFFFF), ax1 = (int16_t)(ra0.x >> 16);
				const int32_t bx0 = (int16_t)(rb0.x & 0xFFFF), bx1 = (int16_t)(rb0.x >> 16);
				// coverage of the lane's left pair (x = px0, px0 + 1) and right pair (x = px0 + 16, px0 + 17) of both rows
				const bool cov_al = ax0 < px0 + 2 && ax1 > px0, cov_ar = ax0 < px0 + 18 && ax1 > px0 + 16;
				const bool cov_bl = bx0 < px0 + 2 && bx1 > px0, cov_br = bx0 < px0 + 18 && bx1 > px0 + 16;
				// one ballot for the four: bit groups of 8 lanes would need four ballots; a packed vote does it in one reduction
				const uint32_t mine = (cov_al ? 1u : 0u) | (cov_ar ? 2u : 0u) | (cov_bl ? 4u : 0u) | (cov_br ? 8u : 0u);
				const uint32_t any = __reduce_or_sync(0xFFFFFFFFu, mine);
				if (any == 0)
					continue;
				const uint4 w0 = s_cmd[0][h0 + hh];
				const uint4 w1 = s_cmd[1][h0 + hh];
				const void* const tex = reinterpret_cast<const void*>((uint64_t)w1.x | ((uint64_t)w1.y << 32));
				if ((w0.w & 0xFF) == SQW_CMD_TRIANGLE) // (warp-uniform) small triangles mostly touch one 16-pixel side of the block: shade only that pair
				{
					const uint4 ra1 = s_span[1][hh][ra], ra2 = s_span[2][hh][ra], rb1 = s_span[1][hh][rb], rb2 = s_span[2][hh][rb];
					if (any & 1u)
						shade_triangle<C, 0, 2>(w0, tex, w1.z, ra0, ra1, ra2, px0, color_buffer + pix_a, px_a, have_a);
					if (any & 2u)
						shade_triangle<C, 2, 4>(w0, tex, w1.z, ra0, ra1, ra2, px0, color_buffer + pix_a, px_a, have_a);
					if (any & 4u)
						shade_triangle<C, 0, 2>(w0, tex, w1.z, rb0, rb1, rb2, px0, color_buffer + pix_b, px_b, have_b);
					if (any & 8u)
						shade_triangle<C, 2, 4>(w0, tex, w1.z, rb0, rb1, rb2, px0, color_buffer + pix_b, px_b, have_b);
					continue;
				}
				if (any & 3u) // (warp-uniform)
					shade_quad<C>(w0, tex, w1.z, ra0, s_span[1][hh][ra], s_span[2][hh][ra], px0, color_buffer + pix_a, px_a, have_a);
				if (any & 12u)
					shade_quad<C>(w0, tex, w1.z, rb0, s_span[1][hh][rb], s_span[2][hh][rb], px0, color_buffer + pix_b, px_b, have_b);
			}
			__syncwarp(); // the span stage (and, after the last group, the warp's list) is rewritten next
		}
	}

	// first entries of the next tile's list: the latency overlaps the epilogue below
	idx_next = ((uint32_t)lane < desc_next.y && tile + gridDim.x < num_general) ? __ldg(tile_lists + desc_next.x + lane) : 0u;

	if (lut_pending) // (CTA-uniform: first tile of the CTA) the only CTA-wide barrier of the kernel
	{
		asm volatile("cp.async.wait_all;" ::: "memory");
		__syncthreads();
		lut_pending = false;
	}
	if (row_a < height)
		finish_quad<C, TONEMAP>(px_a, have_a, px0, width, vec_ok, color_buffer + pix_a, ldr + pix_a, s_lut, isb, isg, isr, sb, sg, sr);
	if (row_b < height)
		finish_quad<C, TONEMAP>(px_b, have_b, px0, width, vec_ok, color_buffer + pix_b, ldr + pix_b, s_lut, isb, isg, isr, sb, sg, sr);
	(void)blk_y1;
	} // tile loop

	if (TONEMAP)
	{
		// colour sums for update_exposure (postprocessor.rs:176-179): exact integers, 3 atomics per warp (no CTA barrier: a
		// warp that is done leaves).  A lane adds at most 8 x 65535 per tile and a warp handles a few dozen tiles, so 32-bit
		// partial sums cannot overflow before the warp reduction.
		unsigned long long wb = sb, wg = sg, wr = sr;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			wb += __shfl_xor_sync(0xFFFFFFFFu, wb, o);
			wg += __shfl_xor_sync(0xFFFFFFFFu, wg, o);
			wr += __shfl_xor_sync(0xFFFFFFFFu, wr, o);
		}
		if (lane < 3)
			atomicAdd(sums + ((blockIdx.x * WARPS + warp) % SQW_SUM_SLOTS) * 4 + lane, lane == 0 ? wb : (lane == 1 ? wg : wr));
	}
}

// ---- opaque-polygon tile kernel ---------------------------------------------------------------------------------
// Tiles whose every command is a polygon drawn with BlendingMode::None (class SQW_TILE_CLASS_OPAQUE; all world geometry of a
// frame without translucent surfaces) need no painter's loop: write_into_framebuffer overwrites (rasterizer.rs:1516), so a pixel
// shows the LAST command of the list that covers it, whatever lay below.  The kernel therefore works in two passes per 32x8 block:
//   winners   walks the block's commands in order and only compares the lane's eight pixel positions with the span [x0, x1) of
//             their row: three integer instructions per (pixel, command);
//   resolve   evaluates the reference's pixel loop (rasterizer.rs:424-445 / :558-572 / :705-719) ONCE per pixel, for the winning
//             command, and issues the eight texel gathers of the lane together.
// Shading cost no longer depends on how many commands touch a block (1.7 on average at 4K in the stand-in map: most of them
// only clip a corner of it) nor on overdraw.
// Memory latency is taken off the critical path by a software pipeline over the tiles of a warp: while tile i is evaluated the
// descriptor of tile i + 4, the list of i + 3 and the command records of i + 2 are in flight in registers, and the span records
// of tile i + 1 travel to shared memory with asynchronous copies (LDGSTS, zero-filled for rows a command does not own) into
// the other half of the warp's double buffer.  Warps are autonomous (two 128x8 tiles per CTA of eight warps, no barrier in
// the tile loop); only the tonemapping table is shared by the CTA.
struct OpaqueStage // one warp, one pipeline stage: 2240 bytes
{
	uint4 span[SQW_OPAQUE_STAGE_HITS][SQW_TILE_H][2]; // SpanRec words 0-7 of (hit, block row)
	uint4 cmd[SQW_OPAQUE_STAGE_HITS];                   // texture pointer (2 words) | tex_width | tc_mode
	uint2 rows[SQW_OPAQUE_STAGE_HITS];                  // y0 | y1 << 16, span_base
};
#define SQW_NO_WINNER 0xFFu

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gmem_src, bool valid)
{
	const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
	const int src_size = valid ? 16 : 0; // 0: nothing is read, the 16 bytes are filled with zeros
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_size) : "memory");
}

// Compaction of one 32-entry piece of a tile list for one warp: entries whose column range meets the warp's block, in list
// order, slots [first_slot, first_slot + SQW_OPAQUE_STAGE_HITS) go to the stage.  Returns the number of hits (bits 0-29) and two flags.
__device__ __forceinline__ uint32_t opaque_stage_commands(
	OpaqueStage& st, const uint4 c0, const uint4 c1_in, const bool entry_valid, const int32_t blk_x0, const int32_t blk_w, const char* __restrict__ const surfaces_base,
	const uint32_t first_slot, const int lane)
{
	const int32_t bx0 = (int16_t)(c0.y & 0xFFFF), bx1 = (int16_t)(c0.y >> 16);
	const bool hit = entry_valid && bx0 < blk_x0 + blk_w && bx1 > blk_x0; // the row range was tested by the binning CTA (block rows = band rows)
	const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
	const uint32_t slot = __popc(ballot & ((1u << lane) - 1)) - first_slot; // (wraps for slots before first_slot)
	if (hit && slot < SQW_OPAQUE_STAGE_HITS)
	{
		uint4 c1 = c1_in;
		if ((c0.w >> 25) & 1) // `texture` is a byte offset into the view's surface pool
		{
			const uint64_t t = (uint64_t)reinterpret_cast<uintptr_t>(surfaces_base) + ((uint64_t)c1.x | ((uint64_t)c1.y << 32));
			c1.x = (uint32_t)t;
			c1.y = (uint32_t)(t >> 32);
		}
		st.cmd[slot] = make_uint4(c1.x, c1.y, c1.z, (c0.w >> 16) & 0xFFu);
		st.rows[slot] = make_uint2(c0.x, c0.z);
	}
	// bit 31: some staged command uses the full-perspective pixel loop (the resolve pass then evaluates the division);
	// bit 30: one of those may meet divisors below SQW_Z_ONE_STEP_MIN_D (no SQW_CMDF_Z_ONE_STEP flag): general division
	const bool full = hit && slot < SQW_OPAQUE_STAGE_HITS && ((c0.w >> 16) & 0xFFu) == SQW_TC_FULL_PERSPECTIVE;
	const bool slow = full && !((c0.w >> 24) & SQW_CMDF_Z_ONE_STEP);
	return __popc(ballot) | (__any_sync(0xFFFFFFFFu, full) ? 0x80000000u : 0u) | (__any_sync(0xFFFFFFFFu, slow) ? 0x40000000u : 0u);
}

// Asynchronous copies of the span records of the staged commands (after a __syncwarp that follows opaque_stage_commands).
__device__ __forceinline__ void opaque_stage_spans(OpaqueStage& st, const uint32_t num_staged, const int32_t blk_y0, const SpanRec* __restrict__ const spans, const int lane)
{
#pragma unroll
	for (int j = 0; j < SQW_OPAQUE_STAGE_HITS * SQW_TILE_H / 32; ++j)
	{
		const uint32_t p = (uint32_t)lane + 32u * j, h = p >> 3, r = p & 7;
		if (h < num_staged)
		{
			const uint2 rows = st.rows[h];
			const int32_t y0 = (int16_t)(rows.x & 0xFFFF), y1 = (int16_t)(rows.x >> 16);
			const int32_t srow = blk_y0 + (int32_t)r;
			const bool ok = srow >= y0 && srow < y1;
			const uint4* src = reinterpret_cast<const uint4*>(spans + rows.y + (ok ? (uint32_t)(srow - y0) : 0u));
			cp_async_16_zfill(&st.span[h][r][0], src, ok);
			cp_async_16_zfill(&st.span[h][r][1], src + 1, ok);
		}
	}
}

// Address of the texel the reference's pixel loop fetches for pixel x of a span (polygon commands), as straight-line code:
// the eight pixels of a lane are independent dependency chains that the scheduler interleaves, so nothing here may branch.
// FULL = some lane of the warp has a full-perspective winner among these pixels: the division is then evaluated for every pixel
// and the result selected by the winner's mode (line-z corrected and affine spans step in 16.16, rasterizer.rs:558-572, :705-719).
template <class C, int FULL>
__device__ __forceinline__ const typename C::Raw* opaque_texel_address(const uint4 r0, const uint4 r1, const uint4 cm, const int32_t x)
{
	const int32_t k = x - (int32_t)(int16_t)(r0.x & 0xFFFF);
	const int32_t t0 = wadd((int32_t)r0.w, wmul(k, (int32_t)r1.y)), t1 = wadd((int32_t)r1.x, wmul(k, (int32_t)r1.z));
	uint32_t u = (uint32_t)(t0 >> 16), v = (uint32_t)(t1 >> 16);
	if (FULL)
	{
		const uint32_t span_inv_z = (uint32_t)wadd((int32_t)r0.y, wmul(k, (int32_t)r0.z));
		// rasterizer.rs:430; the divisor is >= 1 inside a full-perspective span by the clamp of the set-up
		const uint32_t z = FULL == 1 ? z_from_inv_z_one_step(span_inv_z >> INV_Z_PRE_SHIFT) : z_from_inv_z(span_inv_z >> INV_Z_PRE_SHIFT);
#ifdef SQW_DEBUG_Z_ONE_STEP
		// validation build: a flagged command met a divisor the one-step form does not cover (the host's bound would be wrong)
		if (FULL == 1 && cm.w == SQW_TC_FULL_PERSPECTIVE && k >= 0 && x < (int32_t)(int16_t)(r0.x >> 16) && z != z_from_inv_z(span_inv_z >> INV_Z_PRE_SHIFT))
			__trap(); // (pixels outside the span have no winner here: their address is never used)
#endif
		// tc = (z as u64 * span_tc as u64) >> 34 (:431-434).  span_tc never leaves [0, tc_max] inside a span (both span ends are
		// clamped to it and the steps are truncated towards zero), so the sign extension of the reference's cast is a no-op
		// and the product is the unsigned 32 x 32 -> 64 one: its high word shifted by two.
		static_assert(TC_FINAL_SHIFT == 34, "tc shift");
		const bool full = cm.w == SQW_TC_FULL_PERSPECTIVE;
		u = full ? __umulhi(z, (uint32_t)t0) >> 2 : u;
		v = full ? __umulhi(z, (uint32_t)t1) >> 2 : v;
	}
	const typename C::Raw* const tex = reinterpret_cast<const typename C::Raw*>((uint64_t)cm.x | ((uint64_t)cm.y << 32));
	return tex + (u + v * cm.z);
}

// 128-bit shared-memory load under a predicate, never a branch: a reload that no lane needs costs an issue slot but no
// shared-memory wavefront (a 128-bit load costs four, broadcast or not: the load/store pipe is what bounds this kernel).
__device__ __forceinline__ void lds128_if(const bool pred, const void* smem_src, uint4& v)
{
	const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem_src);
	asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"((uint32_t)pred), "r"(src));
}

// Texel addresses of the lane's four consecutive pixels of one row.  Neighbouring pixels nearly always share their winner, so
// the winner's records (span record of the row, command) are loaded once and only reloaded — under a predicate — where the
// winner changes inside the quad (lanes on a polygon edge).  Pixels without a winner keep the previous record; their gather is
// predicated off by the caller.  Straight-line code: the pixels are independent chains for the scheduler to interleave.
template <class C, int FULL>
__device__ __forceinline__ void opaque_resolve_quad(const OpaqueStage& st, const uint32_t (&w)[4], const int row, const int32_t px0, const typename C::Raw* (&p)[4])
{
	uint32_t sprev = w[0] == SQW_NO_WINNER ? 0u : w[0];
	uint4 r0 = st.span[sprev][row][0], r1 = st.span[sprev][row][1], cm = st.cmd[sprev];
	p[0] = opaque_texel_address<C, FULL>(r0, r1, cm, px0);
#pragma unroll
	for (int i = 1; i < 4; ++i)
	{
		const uint32_t si = w[i] == SQW_NO_WINNER ? sprev : w[i];
		const bool reload = si != sprev;
		lds128_if(reload, &st.span[si][row][0], r0);
		lds128_if(reload, &st.span[si][row][1], r1);
		lds128_if(reload, &st.cmd[si], cm);
		sprev = si;
		p[i] = opaque_texel_address<C, FULL>(r0, r1, cm, px0 + i);
	}
}

// Lane layout of the opaque kernel: a lane owns FOUR CONSECUTIVE pixels of block row lane >> 3 and of row 4 + (lane >> 3):
// x = px0 + {0, 1, 2, 3}, px0 = block x + 4 * (lane & 7).  A quad is one 256-bit store of 64-bit pixels (STG.256, new with
// sm_100) and one 128-bit store of the 32-bit picture; 8 lanes write 256 contiguous bytes of a row.
//
// Winners of the lane's eight pixels among the staged commands [0, nh), then one texel gather per pixel that has one.
template <class C>
__device__ __forceinline__ void opaque_chunk(
	const OpaqueStage& st, const uint32_t nh, const bool any_full, const bool one_step, const int ra, const int rb, const int32_t px0, typename C::Raw (&px_a)[4], typename C::Raw (&px_b)[4],
	uint32_t& have_a, uint32_t& have_b)
{
	typedef typename C::Raw Raw;
	uint32_t wa[4], wb[4];
#pragma unroll
	for (int i = 0; i < 4; ++i)
		wa[i] = wb[i] = SQW_NO_WINNER;
	for (uint32_t h = 0; h < nh; ++h)
	{
		const uint32_t xa = st.span[h][ra][0].x, xb = st.span[h][rb][0].x;
		const int32_t a0 = (int16_t)(xa & 0xFFFF), b0 = (int16_t)(xb & 0xFFFF);
		const uint32_t alen = (uint32_t)max((int32_t)(int16_t)(xa >> 16) - a0, 0), blen = (uint32_t)max((int32_t)(int16_t)(xb >> 16) - b0, 0);
		const int32_t da = px0 - a0, db = px0 - b0;
#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			wa[i] = (uint32_t)(da + i) < alen ? h : wa[i];
			wb[i] = (uint32_t)(db + i) < blen ? h : wb[i];
		}
	}
	uint32_t valid = 0;
	const Raw* pa[4];
	const Raw* pb[4];
#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		valid |= (wa[i] != SQW_NO_WINNER ? 1u : 0u) << i;
		valid |= (wb[i] != SQW_NO_WINNER ? 16u : 0u) << i;
	}
	if (any_full) // (warp-uniform) some staged command is drawn with full perspective correction
	{
		if (one_step) // (warp-uniform) every staged full-perspective command keeps its divisors >= SQW_Z_ONE_STEP_MIN_D (CmdDev flag, set by the host)
		{
			opaque_resolve_quad<C, 1>(st, wa, ra, px0, pa);
			opaque_resolve_quad<C, 1>(st, wb, rb, px0, pb);
		}
		else
		{
			opaque_resolve_quad<C, 2>(st, wa, ra, px0, pa);
			opaque_resolve_quad<C, 2>(st, wb, rb, px0, pb);
		}
	}
	else
	{
		opaque_resolve_quad<C, 0>(st, wa, ra, px0, pa);
		opaque_resolve_quad<C, 0>(st, wb, rb, px0, pb);
	}
	// the eight gathers of the lane are in flight together
	Raw ta[4], tb[4];
#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		ta[i] = tb[i] = 0;
		if (valid & (1u << i))
			ta[i] = __ldg(pa[i]);
		if (valid & (16u << i))
			tb[i] = __ldg(pb[i]);
	}
#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		px_a[i] = (valid & (1u << i)) ? ta[i] : px_a[i];
		px_b[i] = (valid & (16u << i)) ? tb[i] : px_b[i];
	}
	have_a |= valid & 0xFu;
	have_b |= valid >> 4;
}

// The finished quad of one row (four consecutive pixels): one 256-bit store of the 64-bit pixels (128-bit for 32-bit ones), and
// with TONEMAP one 128-bit store of the 32-bit picture plus the colour sums.  Same arithmetic as finish_quad.
template <class C, int TONEMAP>
__device__ __forceinline__ void opaque_finish_quad(
	typename C::Raw (&px)[4], uint32_t have, const int32_t px0, const int32_t width, const bool vec_ok, typename C::Raw* __restrict__ const dst,
	uint32_t* __restrict__ const o, const uint8_t* const s_lut, const float isb, const float isg, const float isr, uint32_t& sb, uint32_t& sg, uint32_t& sr)
{
	typedef typename C::Raw Raw;
	if (vec_ok && have == 0xF)
	{
		if (sizeof(Raw) == 8)
			asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"((uint64_t)px[0]), "l"((uint64_t)px[1]), "l"((uint64_t)px[2]), "l"((uint64_t)px[3]) : "memory");
		else
			*reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)px[0], (uint32_t)px[1], (uint32_t)px[2], (uint32_t)px[3]);
		if (TONEMAP)
		{
			uint32_t out[4];
			const uint64_t any = (uint64_t)px[0] | (uint64_t)px[1] | (uint64_t)px[2] | (uint64_t)px[3];
			uint32_t cb[4], cg[4], cr[4];
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const uint64_t p = (uint64_t)px[i];
				cb[i] = (uint32_t)(p & 0xFFFF);
				cg[i] = (uint32_t)((p >> 16) & 0xFFFF);
				cr[i] = (uint32_t)((p >> 32) & 0xFFFF);
				sb += cb[i];
				sg += cg[i];
				sr += cr[i];
			}
			if ((any & (0x0000FFFFFFFFFFFFull & ~(0x0001000100010001ull * (SQW_TONEMAP_LUT_SIZE - 1)))) == 0)
			{
#pragma unroll
				for (int i = 0; i < 4; ++i)
					out[i] = (uint32_t)s_lut[cb[i]] | ((uint32_t)s_lut[SQW_TONEMAP_LUT_SIZE + cg[i]] << 8) | ((uint32_t)s_lut[2 * SQW_TONEMAP_LUT_SIZE + cr[i]] << 16) | 0xFF000000u;
			}
			else
			{
#pragma unroll
				for (int i = 0; i < 4; ++i)
					out[i] = tonemap_pixel(cv_from_color64((uint64_t)px[i]), isb, isg, isr);
			}
			__stcs(reinterpret_cast<uint4*>(o), make_uint4(out[0], out[1], out[2], out[3]));
		}
	}
	else
	{
#pragma unroll
		for (int i = 0; i < 4; ++i)
			if (px0 + i < width && (have & (1u << i)))
				dst[i] = px[i];
		if (TONEMAP)
		{
#pragma unroll
			for (int i = 0; i < 4; ++i)
				if (px0 + i < width)
				{
					if (!(have & (1u << i)))
						px[i] = dst[i]; // untouched pixel: tonemap what the buffer holds
					const uint64_t p = (uint64_t)px[i];
					sb += (uint32_t)(p & 0xFFFF);
					sg += (uint32_t)((p >> 16) & 0xFFFF);
					sr += (uint32_t)((p >> 32) & 0xFFFF);
					o[i] = tonemap_pixel(cv_from_color64(p), isb, isg, isr);
				}
		}
	}
}

template <class C, int TONEMAP>
__global__ void __launch_bounds__(SQW_OPAQUE_WARPS * 32, SQW_OPAQUE_CTAS_PER_SM) raster_tile_opaque_kernel(
	const CmdDev* __restrict__ cmd_dev, const SpanRec* __restrict__ spans, const uint32_t* __restrict__ tile_lists, const uint4* __restrict__ tile_desc,
	const char* __restrict__ surfaces_base, typename C::Raw* __restrict__ color_buffer, int32_t width, int32_t height, int32_t pitch, int clear,
	typename C::Raw clear_color, uint32_t* __restrict__ ldr, float isb, float isg, float isr, unsigned long long* __restrict__ sums,
	const uint8_t* __restrict__ tonemap_lut, uint32_t num_tiles)
{
	typedef typename C::Raw Raw;
	constexpr int THREADS = SQW_OPAQUE_WARPS * 32;
	constexpr uint32_t K = SQW_OPAQUE_STAGE_HITS;
	// A warp renders BLOCKS adjacent 32x8 blocks of its tile one after the other from ONE staged command list (hits against the
	// 32 * BLOCKS pixel range): the per-tile work — list compaction, span staging, the prefetch steps — is paid once per BLOCKS
	// blocks, while the registers hold one block at a time.
	constexpr int BLOCKS = SQW_OPAQUE_BLOCKS, WPT = 4 / BLOCKS; // warps per tile
	static_assert((BLOCKS == 1 || BLOCKS == 2 || BLOCKS == 4) && SQW_OPAQUE_WARPS % WPT == 0 && SQW_TILE_W == 128 && SQW_TILE_H == 8 && (K * SQW_TILE_H) % 32 == 0, "tile geometry");
	__shared__ OpaqueStage s_stage_all[SQW_OPAQUE_WARPS][2];
	__shared__ __align__(16) uint8_t s_lut[TONEMAP ? 3 * SQW_TONEMAP_LUT_SIZE : 16];
	if (TONEMAP)
	{
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_lut);
		for (uint32_t i = threadIdx.x; i < 3 * SQW_TONEMAP_LUT_SIZE / 16; i += THREADS)
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(tonemap_lut + (size_t)i * 16) : "memory");
		asm volatile("cp.async.commit_group;" ::: "memory");
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	OpaqueStage(&stage)[2] = s_stage_all[warp];
	uint32_t sb = 0, sg = 0, sr = 0;
	const int ra = lane >> 3, rb = 4 + (lane >> 3);
	const int32_t warp_x = (warp % WPT) * 32 * BLOCKS;  // first pixel of the warp's range relative to the tile
	const int32_t lane_x = warp_x + (lane & 7) * 4;     // first pixel of the lane in block 0
	const bool pitch_vec = (pitch & 3) == 0; // rows start on 32-byte boundaries: the quad stores are aligned

	// Tiles of this warp: the tile group (WPT warps) g of CTA b takes tiles (b * G + g) + n * (gridDim.x * G).
	constexpr uint32_t G = SQW_OPAQUE_WARPS / WPT;
	const uint32_t tile_first = blockIdx.x * G + (uint32_t)(warp / WPT), tile_stride = gridDim.x * G;

	// ---- pipeline registers ----
	// A tile travels through four loads, one per iteration of the loop below: descriptor -> first 32 list entries -> their
	// command records -> span records (shared memory).  `inf*` = the descriptor fields the later steps need:
	// origin (x0 | y0 << 16) in .x; list length, or ~0 for a tile that is absent / of the other class in .y.
	auto load_desc = [&](uint32_t tile) { return tile < num_tiles ? __ldg(tile_desc + tile) : make_uint4(0, 0, 0, SQW_TILE_CLASS_GENERAL); };
	auto desc_info = [&](const uint4 d, uint32_t tile) { return make_uint2(d.z, (tile < num_tiles && d.w == SQW_TILE_CLASS_OPAQUE) ? d.y : 0xFFFFFFFFu); };
	auto load_idx = [&](const uint4 d, const uint2 inf) { return (inf.y != 0xFFFFFFFFu && (uint32_t)lane < inf.y) ? __ldg(tile_lists + d.x + lane) : 0xFFFFFFFFu; };
	uint4 c0n = make_uint4(0, 0, 0, 0), c1n = c0n; // command records of the entries of tile i + 2 (one entry per lane)
	bool c_valid = false;
	auto load_cmds = [&](uint32_t idx) {
		c_valid = idx != 0xFFFFFFFFu;
		if (c_valid)
		{
			const uint4* cdp = reinterpret_cast<const uint4*>(cmd_dev + idx);
			c0n = __ldg(cdp);
			c1n = __ldg(cdp + 1);
		}
	};
	// hits of a tile's first list piece -> stage buffer + asynchronous span copies; returns the hit count
	auto stage_tile = [&](OpaqueStage& st, const uint2 inf) -> uint32_t {
		uint32_t nh = 0;
		if (inf.y != 0xFFFFFFFFu) // (warp-uniform)
		{
			nh = opaque_stage_commands(st, c0n, c1n, c_valid, (int32_t)(inf.x & 0xFFFF) + warp_x, 32 * BLOCKS, surfaces_base, 0, lane);
			__syncwarp();
			opaque_stage_spans(st, min(nh & 0x3FFFFFFFu, K), (int32_t)(inf.x >> 16), spans, lane);
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		return nh;
	};

	// ---- prologue: fill the pipeline for tiles 0 .. 3 of this warp ----
	uint32_t tile = tile_first;
	uint4 d0 = load_desc(tile), d1 = load_desc(tile + tile_stride), d2 = load_desc(tile + 2 * tile_stride), d3 = load_desc(tile + 3 * tile_stride);
	uint2 inf0 = desc_info(d0, tile), inf1 = desc_info(d1, tile + tile_stride), inf2 = desc_info(d2, tile + 2 * tile_stride), inf3 = desc_info(d3, tile + 3 * tile_stride);
	uint32_t idx0 = load_idx(d0, inf0), idx1 = load_idx(d1, inf1), idxn = load_idx(d2, inf2);
	load_cmds(idx0);
	uint32_t nh0 = stage_tile(stage[0], inf0);
	load_cmds(idx1);
	// from here on: c0n/c1n belong to tile + 1, idxn to tile + 2, dn to tile + 3
	uint4 dn = d3;
	uint2 infn = inf3;
	if (TONEMAP)
	{
		// the table (and the first tile's span records) have arrived; the only CTA-wide barrier of the kernel
		asm volatile("cp.async.wait_all;" ::: "memory");
		__syncthreads();
	}

	for (uint32_t it = 0; tile < num_tiles; tile += tile_stride, ++it)
	{
		OpaqueStage& st = stage[it & 1];
		// (A) tile + 1: hits -> other stage buffer, span records on their way
		const uint32_t nh1 = stage_tile(stage[(it & 1) ^ 1], inf1);
		// (B) tile + 2: command records of its list entries
		load_cmds(idxn);
		// (C) tile + 3: first 32 list entries
		idxn = load_idx(dn, infn);
		const uint2 inf3_new = infn;
		// (D) tile + 4: descriptor
		dn = load_desc(tile + 4 * tile_stride);
		infn = desc_info(dn, tile + 4 * tile_stride);
		// (E) the span records of this tile have landed (everything but the group committed in (A))
		asm volatile("cp.async.wait_group 1;" ::: "memory");
		__syncwarp();

		if (inf0.y != 0xFFFFFFFFu) // (warp-uniform) a tile of this kernel's class
		for (int blk = 0; blk < BLOCKS; ++blk)
		{
			const int32_t blk_x0 = (int32_t)(inf0.x & 0xFFFF) + warp_x + 32 * blk;
			const int32_t blk_y0 = (int32_t)(inf0.x >> 16);
			const int32_t row_a = blk_y0 + ra, row_b = blk_y0 + rb;
			const int32_t px0 = (int32_t)(inf0.x & 0xFFFF) + lane_x + 32 * blk;
			const uint32_t pix_a = (uint32_t)row_a * (uint32_t)pitch + (uint32_t)px0;
			const uint32_t pix_b = pix_a + 4u * (uint32_t)pitch;
			Raw px_a[4], px_b[4];
			uint32_t have_a = clear ? 0xFu : 0u, have_b = have_a;
#pragma unroll
			for (int i = 0; i < 4; ++i)
				px_a[i] = px_b[i] = clear_color;
			const bool vec_ok = (blk_x0 + 32 <= width) && pitch_vec;
			if (blk_x0 < width) // (warp-uniform)
			{
				if (inf0.y <= 32 && (nh0 & 0x3FFFFFFFu) <= K)
				{
					if (nh0)
						opaque_chunk<C>(st, nh0 & 0x3FFFFFFFu, (nh0 >> 31) != 0, !((nh0 >> 30) & 1), ra, rb, px0, px_a, px_b, have_a, have_b);
				}
				else
				{
					// Rare: a list longer than one piece, or more hits than a stage holds (far, finely tessellated geometry).  The
					// same two passes, chunk by chunk in list order (a later chunk's winners replace earlier ones), staged
					// synchronously for THIS block into the tile's stage buffer.
					const uint32_t list_start = __ldg(&tile_desc[tile].x);
					for (uint32_t base = 0; base < inf0.y; base += 32)
					{
						const bool valid = base + lane < inf0.y;
						uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
						if (valid)
						{
							const uint4* cdp = reinterpret_cast<const uint4*>(cmd_dev + __ldg(tile_lists + list_start + base + lane));
							c0 = __ldg(cdp);
							c1 = __ldg(cdp + 1);
						}
						uint32_t nh = 1;
						for (uint32_t first = 0; first < nh; first += K)
						{
							__syncwarp();
							const uint32_t nhf = opaque_stage_commands(st, c0, c1, valid, blk_x0, 32, surfaces_base, first, lane);
							nh = nhf & 0x3FFFFFFFu;
							if (first >= nh)
								break;
							__syncwarp();
							const uint32_t n = min(nh - first, K);
							opaque_stage_spans(st, n, blk_y0, spans, lane);
							asm volatile("cp.async.commit_group;" ::: "memory");
							asm volatile("cp.async.wait_all;" ::: "memory");
							__syncwarp();
							opaque_chunk<C>(st, n, (nhf >> 31) != 0, !((nhf >> 30) & 1), ra, rb, px0, px_a, px_b, have_a, have_b);
						}
					}
					__syncwarp();
				}
			}
			if (row_a < height)
				opaque_finish_quad<C, TONEMAP>(px_a, have_a, px0, width, vec_ok, color_buffer + pix_a, ldr + pix_a, s_lut, isb, isg, isr, sb, sg, sr);
			if (row_b < height)
				opaque_finish_quad<C, TONEMAP>(px_b, have_b, px0, width, vec_ok, color_buffer + pix_b, ldr + pix_b, s_lut, isb, isg, isr, sb, sg, sr);
		}
		__syncwarp(); // this stage buffer is rewritten by step (A) of the next iteration but one; the other one by the next
		inf0 = inf1;
		inf1 = inf2;
		inf2 = inf3_new;
		nh0 = nh1;
	}
	asm volatile("cp.async.wait_all;" ::: "memory"); // nothing may still be writing this CTA's shared memory when it retires

	if (TONEMAP)
	{
		unsigned long long wb = sb, wg = sg, wr = sr;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			wb += __shfl_xor_sync(0xFFFFFFFFu, wb, o);
			wg += __shfl_xor_sync(0xFFFFFFFFu, wg, o);
			wr += __shfl_xor_sync(0xFFFFFFFFu, wr, o);
		}
		if (lane < 3 && (wb | wg | wr))
			atomicAdd(sums + ((blockIdx.x * SQW_OPAQUE_WARPS + warp) % SQW_SUM_SLOTS) * 4 + lane, lane == 0 ? wb : (lane == 1 ? wg : wr));
	}
}
#undef SQW_NO_WINNER
#undef SQW_NEED_DST
#undef SQW_PX_OFS


} // namespace

cudaError_t launch_raster_setup(const RasterLaunch& l, cudaStream_t stream)
{
	if (l.num_bands == 0)
		return cudaSuccess;
	const uint32_t chunks_per_block = SQW_SETUP_THREADS / 32;
	const uint32_t setup_blocks = std::max<uint32_t>((l.num_chunks + chunks_per_block - 1) / chunks_per_block, 1); // block 0 also zeroes the sums
	const uint32_t bin_blocks = l.num_bands; // also with an empty command list: the tile kernel reads every tile descriptor
	const uint32_t lut_blocks = l.tonemap_dst ? (3 * SQW_TONEMAP_LUT_SIZE + SQW_SETUP_THREADS - 1) / SQW_SETUP_THREADS : 0;
	// the binning CTAs append the frame's general tiles behind this counter (no general command -> no general tile, and
	// launch_raster_tiles does not start the kernel that reads the list: frames of plain world geometry skip the reset)
	if (l.num_general_cmds)
	{
		const cudaError_t e = cudaMemsetAsync(l.general_tiles, 0, sizeof(uint32_t), stream);
		if (e != cudaSuccess)
			return e;
	}
	if (l.num_triangle_cmds)
		raster_triangle_setup_kernel<<<(l.num_cmds * 8 + 127) / 128, 128, 0, stream>>> // eight lanes per command, one interpolant per lane
		(l.cmds, l.num_cmds, l.triangle_points, l.height, static_cast<TriRec*>(l.tri_recs));
	raster_setup_kernel<<<setup_blocks + bin_blocks + lut_blocks, SQW_SETUP_THREADS, 0, stream>>>(
		l.cmds, l.cmd_dev, l.num_cmds, l.polygon_points, static_cast<const TriRec*>(l.tri_recs), l.chunks, l.num_chunks, l.spans, l.width, l.height, setup_blocks, l.band_offsets,
		l.band_lists, l.tile_lists, l.tile_desc, l.general_tiles, (uint32_t)((l.width + SQW_TILE_W - 1) / SQW_TILE_W), bin_blocks, l.tonemap_dst ? l.color_sums : nullptr, l.tonemap_dst ? l.tonemap_lut : nullptr, l.inv_scale_bgr[0], l.inv_scale_bgr[1], l.inv_scale_bgr[2]);
	return cudaGetLastError();
}

cudaError_t launch_z_from_inv_z_test(unsigned long long* mismatches, cudaStream_t stream)
{
	z_from_inv_z_test_kernel<<<(1u << 24) / 256, 256, 0, stream>>>(mismatches);
	fixed16_div_test_kernel<<<(1u << 26) / 256, 256, 0, stream>>>(mismatches);
	return cudaGetLastError();
}

cudaError_t launch_raster_tiles(const RasterLaunch& l, cudaStream_t stream)
{
	if (l.num_bands == 0)
		return cudaSuccess;
	const uint32_t tiles_x = (uint32_t)((l.width + SQW_TILE_W - 1) / SQW_TILE_W), num_tiles = tiles_x * l.num_bands;
	// persistent grid: resident CTAs per SM (registers and shared memory allow 4) x SM count
	static int sm_count = 0;
	if (sm_count == 0)
	{
		int dev = 0;
		cudaGetDevice(&dev);
		if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
			sm_count = 148;
	}
#ifdef SQW_TILE_NONPERSISTENT
	const uint32_t grid = num_tiles;
#else
	const uint32_t grid = std::min<uint32_t>(num_tiles, (uint32_t)sm_count * SQW_TILE_CTAS_PER_SM);
#endif
#define SQW_TILE_ARGS(RAW) \
	l.cmd_dev, l.spans, l.tile_lists, l.tile_desc, static_cast<const char*>(l.surfaces_base), static_cast<RAW*>(l.color_buffer), l.width, l.height, l.pitch, l.clear, \
		(RAW)l.clear_color, l.tonemap_dst, l.inv_scale_bgr[0], l.inv_scale_bgr[1], l.inv_scale_bgr[2], l.color_sums, l.tonemap_lut, num_tiles
	// Tiles of class SQW_TILE_CLASS_OPAQUE (only polygons with BlendingMode::None: all of them in a frame of plain world geometry).
	{
		constexpr uint32_t G = SQW_OPAQUE_WARPS / (4 / SQW_OPAQUE_BLOCKS); // tiles a CTA renders at a time
		const uint32_t ogrid = std::min<uint32_t>((num_tiles + G - 1) / G, (uint32_t)sm_count * SQW_OPAQUE_CTAS_PER_SM);
		if (l.color64 && l.tonemap_dst)
			raster_tile_opaque_kernel<Color64, 1><<<ogrid, SQW_OPAQUE_WARPS * 32, 0, stream>>>(SQW_TILE_ARGS(uint64_t));
		else if (l.color64)
			raster_tile_opaque_kernel<Color64, 0><<<ogrid, SQW_OPAQUE_WARPS * 32, 0, stream>>>(SQW_TILE_ARGS(uint64_t));
		else
			raster_tile_opaque_kernel<Color32, 0><<<ogrid, SQW_OPAQUE_WARPS * 32, 0, stream>>>(SQW_TILE_ARGS(uint32_t));
		if (l.num_general_cmds == 0)
			return cudaGetLastError();
	}
	// The other tiles: triangles, blended and alpha-tested polygons (the kernel skips the opaque class).
	if (l.color64 && l.tonemap_dst)
		raster_tile_kernel<Color64, 1><<<grid, SQW_RASTER_THREADS, 0, stream>>>(SQW_TILE_ARGS(uint64_t), l.general_tiles);
	else if (l.color64)
		raster_tile_kernel<Color64, 0><<<grid, SQW_RASTER_THREADS, 0, stream>>>(SQW_TILE_ARGS(uint64_t), l.general_tiles);
	else
		raster_tile_kernel<Color32, 0><<<grid, SQW_RASTER_THREADS, 0, stream>>>(SQW_TILE_ARGS(uint32_t), l.general_tiles);
#undef SQW_TILE_ARGS
	return cudaGetLastError();
}

} // namespace sqw
