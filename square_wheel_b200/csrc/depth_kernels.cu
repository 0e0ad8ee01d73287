// Shadow-map depth rasteriser on sm_100a.
//
// Replaces DepthRasterizer::fill_polygon / fill_polygon_part (src/square_wheel/rasterizer.rs:1295-1474) and
// make_shadow_map_circle (src/square_wheel/light.rs:133-157) as driven by DepthRenderer::draw_map
// (src/square_wheel/depth_renderer.rs:47-67) from PartialRenderer::build_shadow_maps (partial_renderer.rs:524-590), which the
// reference runs single-threaded on the CPU (:527 "TODO - use multithreading").
//
// The reference accumulates inverse depth with float additions along y (line_inv_z += d_inv_z_dy, rasterizer.rs:1470) and
// along x (inv_z += d_inv_z_dx, :1463); float addition does not associate, so the only way to get its bits is to replay the
// same additions in the same order.  The replay is cheap (shadow maps are at most 1024 texels wide) and it parallelises over
// what IS independent:
//   depth_setup_kernel    one thread per (fill call, scan-line): edge walk to the owning trapezoid (shared with the colour
//                         rasteriser, sqw_raster_walk.cuh), x range, and the row's starting inverse depth replayed with the
//                         reference's n additions -> one 16-byte DepthSpan;
//   depth_resolve_kernel  one thread per (map, row): optional clear, then the fill calls of the map IN ORDER — a row is owned
//                         by one thread, so painter's order (world polygons back to front, depth-tested submodels after)
//                         needs no atomics — each span written left to right with the running `inv_z += d_inv_z_dx`,
//                         then the projector circle mask.
// A frame with 16 lights has ~100 maps x 64..256 rows = tens of thousands of independent rows.
#include "sqw_device.cuh"
#include "sqw_internal.h"
#include "sqw_raster_walk.cuh"

namespace sqw
{
namespace
{

__global__ void __launch_bounds__(128) depth_setup_kernel(
	const sqw_depth_cmd* __restrict__ cmds, const DepthCmdDev* __restrict__ cmd_dev, const sqw_polygon_point* __restrict__ points, const sqw_depth_map* __restrict__ maps,
	const uint2* __restrict__ chunks, uint32_t num_chunks, DepthSpan* __restrict__ spans)
{
	const uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (chunk >= num_chunks)
		return;
	const uint2 ch = __ldg(chunks + chunk); // {command, first row relative to the command's y0}
	const DepthCmdDev cd = cmd_dev[ch.x];
	const uint32_t r = ch.y + (threadIdx.x & 31);
	if (r >= (uint32_t)((int32_t)cd.y1 - (int32_t)cd.y0))
		return;
	const sqw_depth_cmd c = cmds[ch.x];
	const int32_t width = (int32_t)maps[cd.map].width, height = (int32_t)maps[cd.map].height;
	const int32_t row = (int32_t)cd.y0 + (int32_t)r;

	DepthSpan out;
	out.x0 = out.x1 = 0;
	out.inv_z = 0.0f;
	out.d_inv_z_dx = c.depth_eq[0];
	out.pad = 0;
	PolygonPart part;
	if (find_polygon_part(points + c.vertex_offset, c.num_vertices, row, height, part))
	{
		// fill_polygon_part (rasterizer.rs:1437-1471) up to the pixel loop of line `row`
		const int32_t n = row - part.y_start_int;
		const Fixed16 y_start_delta = wsub(wadd(int_to_fixed16(part.y_start_int), SQW_FIXED16_HALF), part.y_start);
		const Fixed16 x_left = wadd(wadd(wadd(part.left.x_start, fixed16_mul(y_start_delta, part.left.dx_dy)), SQW_FIXED16_HALF), wmul(n, part.left.dx_dy));
		const Fixed16 x_right = wadd(wadd(wadd(part.right.x_start, fixed16_mul(y_start_delta, part.right.dx_dy)), SQW_FIXED16_HALF), wmul(n, part.right.dx_dy));
		const int32_t x_start_int = max(fixed16_floor_to_int(x_left), 0);
		const int32_t x_end_int = min(fixed16_floor_to_int(x_right), width);
		if (x_start_int < x_end_int)
		{
			const float d_inv_z_dx = c.depth_eq[0], d_inv_z_dy = c.depth_eq[1], k = c.depth_eq[2];
			// :1443-1445, then one `line_inv_z += d_inv_z_dy` per line already passed (:1470) — replayed, not closed-formed
			float line_inv_z = __fadd_rn(__fadd_rn(__fmul_rn((float)part.y_start_int, d_inv_z_dy), __fmul_rn(0.5f, __fadd_rn(d_inv_z_dx, d_inv_z_dy))), k);
			for (int32_t i = 0; i < n; ++i)
				line_inv_z = __fadd_rn(line_inv_z, d_inv_z_dy);
			out.x0 = (int16_t)x_start_int;
			out.x1 = (int16_t)x_end_int;
			out.inv_z = __fadd_rn(line_inv_z, __fmul_rn((float)x_start_int, d_inv_z_dx)); // :1456
		}
	}
	spans[cd.span_base + r] = out;
}

// One WARP per shadow-map row (lane = pixel x mod 32, so a pixel always belongs to the same lane and the fills of a row
// stay in submission order without any synchronisation).  The fills that own a span on the row are found 32 at a time
// (one candidate per lane, ballot); each span is then replayed exactly like DepthRasterizer::fill_polygon's pixel loop
// (rasterizer.rs:1455-1466): the running sum `inv_z += d_inv_z_dx` is carried ALONG the span by warp-uniform additions —
// one per pixel, each lane keeping the value of its own pixel — so the float rounding sequence is the reference's, and
// every store / max is a coalesced 128-byte row segment.
__global__ void __launch_bounds__(128) depth_resolve_kernel(
	const sqw_depth_cmd* __restrict__ cmds, const DepthCmdDev* __restrict__ cmd_dev, const sqw_depth_map* __restrict__ maps, const uint2* __restrict__ row_chunks,
	uint32_t num_row_chunks, const DepthSpan* __restrict__ spans, float* __restrict__ shadow_maps)
{
	const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const uint32_t chunk = warp_global >> 5;
	if (chunk >= num_row_chunks)
		return;
	const int lane = threadIdx.x & 31;
	const uint2 ch = __ldg(row_chunks + chunk); // {map, first row of the 32-row chunk}
	const sqw_depth_map map = maps[ch.x];
	const int32_t row = (int32_t)(ch.y + (warp_global & 31));
	if (row >= (int32_t)map.height)
		return;
	float* const line = shadow_maps + (size_t)map.offset + (size_t)row * map.width;
	const int32_t width = (int32_t)map.width;
	if (map.flags & SQW_DEPTH_MAP_CLEAR)
		for (int32_t x = lane; x < width; x += 32)
			line[x] = map.clear_value;

	const uint32_t cmd_end = map.first_cmd + map.num_cmds;
	for (uint32_t base = map.first_cmd; base < cmd_end; base += 32)
	{
		const uint32_t ci = base + (uint32_t)lane;
		DepthSpan s;
		s.x0 = s.x1 = 0;
		s.inv_z = s.d_inv_z_dx = 0.0f;
		uint32_t depth_test = 0;
		if (ci < cmd_end)
		{
			const DepthCmdDev cd = cmd_dev[ci];
			if (row >= (int32_t)cd.y0 && row < (int32_t)cd.y1)
			{
				s = spans[cd.span_base + (uint32_t)(row - (int32_t)cd.y0)];
				depth_test = cd.depth_test;
			}
		}
		uint32_t pending = __ballot_sync(0xFFFFFFFFu, s.x0 < s.x1);
		while (pending)
		{
			const int src = __ffs(pending) - 1;
			pending &= pending - 1;
			const int32_t x0 = __shfl_sync(0xFFFFFFFFu, (int32_t)s.x0, src), x1 = __shfl_sync(0xFFFFFFFFu, (int32_t)s.x1, src);
			const float d = __shfl_sync(0xFFFFFFFFu, s.d_inv_z_dx, src);
			const uint32_t test = __shfl_sync(0xFFFFFFFFu, depth_test, src);
			float v = __shfl_sync(0xFFFFFFFFu, s.inv_z, src); // value of pixel x0
			for (int32_t cx = x0 & ~31; cx < x1; cx += 32)
			{
				const int32_t start = max(x0, cx), end = min(cx + 32, x1); // v = value of pixel `start`
				const int32_t x = cx + lane;
				float mine = 0.0f;
				for (int32_t px = start; px < end; ++px) // warp-uniform trip count
				{
					mine = px == x ? v : mine;
					v = __fadd_rn(v, d);
				}
				if (x >= start && x < end)
					line[x] = test ? fmaxf(mine, line[x]) : mine; // f32::max when depth-tested (rasterizer.rs:1459)
			}
		}
	}

	if ((map.flags & SQW_DEPTH_MAP_PROJECTOR_CIRCLE) && map.width == map.height)
	{
		// make_shadow_map_circle (light.rs:133-157)
		const int32_t size = width, radius = size / 2;
		const int32_t dy = row - radius;
		const int32_t dx = radius - min((int32_t)ceilf(__fsqrt_rn((float)max(radius * radius - dy * dy, 0))), radius);
		for (int32_t x = lane; x < dx; x += 32)
			line[x] = 1.0e24f;
		for (int32_t x = size - dx + lane; x < size; x += 32)
			line[x] = 1.0e24f;
	}
}

} // namespace

cudaError_t launch_depth_setup(const DepthLaunch& l, cudaStream_t stream)
{
	if (l.num_chunks == 0)
		return cudaSuccess;
	depth_setup_kernel<<<(l.num_chunks + 3) / 4, 128, 0, stream>>>(l.cmds, l.cmd_dev, l.points, l.maps, l.chunks, l.num_chunks, l.spans);
	return cudaGetLastError();
}

cudaError_t launch_depth_resolve(const DepthLaunch& l, cudaStream_t stream)
{
	if (l.num_row_chunks == 0)
		return cudaSuccess;
	depth_resolve_kernel<<<l.num_row_chunks * 8, 128, 0, stream>>>(l.cmds, l.cmd_dev, l.maps, l.row_chunks, l.num_row_chunks, l.spans, l.shadow_maps);
	return cudaGetLastError();
}

} // namespace sqw
