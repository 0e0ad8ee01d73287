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

// ---- k rounded float additions in O(1) ------------------------------------------------------------------------------
// The reference accumulates `v += d` once per pixel / per line, and float addition does not associate.  But inside one binade
// (same sign and exponent, normal numbers) the grid of representable values is uniform, so fl(v + d) - v is the same whole
// number of ulps at every step — with one exception, exact ties, where round-to-nearest-even alternates by the parity of v;
// from a value that itself came out of such a rounding the step is constant again.  So: three real additions (the second and
// third results give the steady step q, in ulps of the magnitude), then n steps at once as an integer addition on the bits,
// n chosen so that the magnitude stays strictly inside the binade (mantissa field in [1, 2^23 - 1]: landing on a power of two
// from above could round on the finer grid below it); a step that leaves the binade is again a real addition.  Zero,
// denormals, infinities and NaNs never jump.  The host-side exhaustive comparison against the plain loop (3.2e8 random, tie,
// binade-edge, zero-crossing and denormal cases, tools/f32_add_repeated_check.c) found no difference.
__device__ __forceinline__ float f32_add_repeated(float v, const float d, int32_t k)
{
	while (k > 0)
	{
		const float v1 = __fadd_rn(v, d);
		if (--k == 0)
			return v1;
		const float v2 = __fadd_rn(v1, d);
		if (--k == 0)
			return v2;
		const float v3 = __fadd_rn(v2, d);
		if (--k == 0)
			return v3;
		v = v3;
		const uint32_t b1 = __float_as_uint(v1), b2 = __float_as_uint(v2), b3 = __float_as_uint(v3);
		const uint32_t e = (b3 >> 23) & 0xFFu;
		if ((b1 >> 23) == (b2 >> 23) && (b2 >> 23) == (b3 >> 23) && e != 0 && e != 0xFFu)
		{
			const int32_t m3 = (int32_t)(b3 & 0x7FFFFFu);
			const int32_t q = m3 - (int32_t)(b2 & 0x7FFFFFu);
			int32_t n = k;
			if (q > 0)
				n = min((0x7FFFFF - m3) / q, k);
			else if (q < 0)
				n = min((m3 - 1) / (-q), k);
			if (n > 0)
			{
				v = __uint_as_float(b3 + (uint32_t)(n * q));
				k -= n;
			}
		}
	}
	return v;
}

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
			// :1443-1445, then one `line_inv_z += d_inv_z_dy` per line already passed (:1470): the same bits in O(1) (f32_add_repeated)
			float line_inv_z = __fadd_rn(__fadd_rn(__fmul_rn((float)part.y_start_int, d_inv_z_dy), __fmul_rn(0.5f, __fadd_rn(d_inv_z_dx, d_inv_z_dy))), k);
			line_inv_z = f32_add_repeated(line_inv_z, d_inv_z_dy, n);
			out.x0 = (int16_t)x_start_int;
			out.x1 = (int16_t)x_end_int;
			out.inv_z = __fadd_rn(line_inv_z, __fmul_rn((float)x_start_int, d_inv_z_dx)); // :1456
		}
	}
	spans[cd.span_base + r] = out;
}

// Value DepthRasterizer::fill_polygon's pixel loop (rasterizer.rs:1455-1466) writes to pixel x of a span.
__device__ __forceinline__ float depth_span_value(const DepthSpan* __restrict__ spans, const uint32_t ref, const int32_t x)
{
	const uint4 w = __ldg(reinterpret_cast<const uint4*>(spans + ref));
	return f32_add_repeated(__uint_as_float(w.y), __uint_as_float(w.z), x - (int32_t)(int16_t)(w.x & 0xFFFF));
}

// One CTA per 8 consecutive rows of a shadow map, one WARP per row (lane = pixel x mod 32: a pixel always belongs to the same
// lane, so the fills of a row stay in submission order without synchronisation).
//   list      the CTA compacts, in order, the fill calls of the map that own a row of the group (the row warps used to scan
//             the whole command list of the map each);
//   winners   world polygons overwrite (depth_test = 0, painter's order): a row warp only RECORDS the last span covering a
//             pixel (shared memory), and the running sum `inv_z += d_inv_z_dx` of the reference is evaluated once per pixel,
//             in closed form (f32_add_repeated), for that span: the cost no longer depends on the overdraw of the shadow map
//             nor on span lengths;
//   max       depth-tested fills (submodels, f32::max with the stored value, :1459) resolve the pixel's pending winner first.
// Rows are processed in segments of DEPTH_SEG pixels (shared-memory state: 8 bytes per pixel).
#define DEPTH_GROUP_ROWS 8
#define DEPTH_SEG 512
#define DEPTH_LIST_CAP 1024
#define DEPTH_NO_SPAN 0xFFFFFFFFu
__global__ void __launch_bounds__(DEPTH_GROUP_ROWS * 32) depth_resolve_kernel(
	const sqw_depth_cmd* __restrict__ cmds, const DepthCmdDev* __restrict__ cmd_dev, const sqw_depth_map* __restrict__ maps, const uint2* __restrict__ row_chunks,
	uint32_t num_row_chunks, const DepthSpan* __restrict__ spans, float* __restrict__ shadow_maps)
{
	__shared__ uint32_t s_list[DEPTH_LIST_CAP];
	__shared__ uint32_t s_warp_count[DEPTH_GROUP_ROWS + 1];
	__shared__ float s_val[DEPTH_GROUP_ROWS][DEPTH_SEG];
	__shared__ uint32_t s_win[DEPTH_GROUP_ROWS][DEPTH_SEG];
	(void)cmds;
	const uint32_t chunk = blockIdx.x / (32 / DEPTH_GROUP_ROWS);
	if (chunk >= num_row_chunks) // (CTA-uniform)
		return;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint2 ch = __ldg(row_chunks + chunk); // {map, first row of the 32-row chunk}
	const sqw_depth_map map = maps[ch.x];
	const int32_t group_row0 = (int32_t)ch.y + (int32_t)(blockIdx.x % (32 / DEPTH_GROUP_ROWS)) * DEPTH_GROUP_ROWS;
	if (group_row0 >= (int32_t)map.height) // (CTA-uniform)
		return;
	const int32_t group_row1 = min(group_row0 + DEPTH_GROUP_ROWS, (int32_t)map.height);
	const int32_t row = group_row0 + warp;
	const bool row_ok = row < (int32_t)map.height; // idle warps still take part in the CTA's barriers
	const int32_t width = (int32_t)map.width;
	float* const line = shadow_maps + (size_t)map.offset + (size_t)(row_ok ? row : 0) * map.width;
	float* const val = s_val[warp];
	uint32_t* const win = s_win[warp];
	// make_shadow_map_circle (light.rs:133-157): columns [0, circle_dx) and [size - circle_dx, size) of the row
	int32_t circle_dx = 0;
	if ((map.flags & SQW_DEPTH_MAP_PROJECTOR_CIRCLE) && map.width == map.height)
	{
		const int32_t radius = width / 2, dy = row - radius;
		circle_dx = radius - min((int32_t)ceilf(__fsqrt_rn((float)max(radius * radius - dy * dy, 0))), radius);
	}
	const uint32_t cmd_end = map.first_cmd + map.num_cmds;

	for (int32_t seg0 = 0; seg0 < width; seg0 += DEPTH_SEG)
	{
		const int32_t seg1 = min(seg0 + DEPTH_SEG, width);
		if (row_ok)
			for (int32_t x = seg0 + lane; x < seg1; x += 32)
			{
				val[x - seg0] = (map.flags & SQW_DEPTH_MAP_CLEAR) ? map.clear_value : line[x];
				win[x - seg0] = DEPTH_NO_SPAN;
			}
		// the map's fill calls, DEPTH_LIST_CAP list entries at a time
		uint32_t next_cmd = map.first_cmd;
		while (next_cmd < cmd_end) // (CTA-uniform)
		{
			// ---- ordered compaction of the calls that own a row of this group ----
			uint32_t count = 0;
			__syncthreads(); // the previous piece of the list has been consumed
			while (next_cmd < cmd_end && count + DEPTH_GROUP_ROWS * 32 <= DEPTH_LIST_CAP)
			{
				const uint32_t ci = next_cmd + threadIdx.x;
				bool hit = false;
				if (ci < cmd_end)
				{
					const uint32_t rows = __ldg(reinterpret_cast<const uint32_t*>(cmd_dev + ci)); // y0 | y1 << 16
					hit = (int32_t)(int16_t)(rows & 0xFFFF) < group_row1 && (int32_t)(int16_t)(rows >> 16) > group_row0;
				}
				const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
				if (lane == 0)
					s_warp_count[warp] = __popc(ballot);
				__syncthreads();
				uint32_t before = count, total = count;
#pragma unroll
				for (int w = 0; w < DEPTH_GROUP_ROWS; ++w)
				{
					const uint32_t c = s_warp_count[w];
					before += w < warp ? c : 0u;
					total += c;
				}
				if (hit)
					s_list[before + __popc(ballot & ((1u << lane) - 1))] = ci;
				count = total;
				next_cmd += DEPTH_GROUP_ROWS * 32;
				__syncthreads(); // s_warp_count is rewritten by the next round
			}
			next_cmd = min(next_cmd, cmd_end);
			// ---- every row warp walks the list in order ----
			if (row_ok)
				for (uint32_t base = 0; base < count; base += 32)
				{
					uint4 sp = make_uint4(0, 0, 0, 0);
					uint32_t ref = 0, depth_test = 0;
					if (base + lane < count)
					{
						const DepthCmdDev cd = cmd_dev[s_list[base + lane]];
						if (row >= (int32_t)cd.y0 && row < (int32_t)cd.y1)
						{
							ref = cd.span_base + (uint32_t)(row - (int32_t)cd.y0);
							sp = __ldg(reinterpret_cast<const uint4*>(spans + ref));
							depth_test = cd.depth_test;
						}
					}
					const int32_t my_x0 = (int16_t)(sp.x & 0xFFFF), my_x1 = (int16_t)(sp.x >> 16);
					uint32_t pending = __ballot_sync(0xFFFFFFFFu, max(my_x0, seg0) < min(my_x1, seg1));
					while (pending)
					{
						const int src = __ffs(pending) - 1;
						pending &= pending - 1;
						const int32_t x0 = max(__shfl_sync(0xFFFFFFFFu, my_x0, src), seg0), x1 = min(__shfl_sync(0xFFFFFFFFu, my_x1, src), seg1);
						const uint32_t r = __shfl_sync(0xFFFFFFFFu, ref, src);
						const uint32_t test = __shfl_sync(0xFFFFFFFFu, depth_test, src);
						if (!test)
						{
							for (int32_t x = (x0 & ~31) + lane; x < x1; x += 32)
								if (x >= x0)
									win[x - seg0] = r;
						}
						else
						{
							const float v0 = __uint_as_float(__shfl_sync(0xFFFFFFFFu, sp.y, src)), d = __uint_as_float(__shfl_sync(0xFFFFFFFFu, sp.z, src));
							const int32_t sx0 = __shfl_sync(0xFFFFFFFFu, my_x0, src);
							for (int32_t x = (x0 & ~31) + lane; x < x1; x += 32)
								if (x >= x0)
								{
									const uint32_t w = win[x - seg0];
									const float stored = w != DEPTH_NO_SPAN ? depth_span_value(spans, w, x) : val[x - seg0];
									val[x - seg0] = fmaxf(f32_add_repeated(v0, d, x - sx0), stored); // f32::max (rasterizer.rs:1459)
									win[x - seg0] = DEPTH_NO_SPAN;
								}
						}
					}
				}
		}
		// ---- one evaluation and one coalesced store per pixel ----
		if (row_ok)
			for (int32_t x = seg0 + lane; x < seg1; x += 32)
			{
				const uint32_t w = win[x - seg0];
				float out = w != DEPTH_NO_SPAN ? depth_span_value(spans, w, x) : val[x - seg0];
				if (x < circle_dx || x >= width - circle_dx)
					out = 1.0e24f;
				line[x] = out;
			}
		__syncwarp();
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
	depth_resolve_kernel<<<l.num_row_chunks * (32 / DEPTH_GROUP_ROWS), DEPTH_GROUP_ROWS * 32, 0, stream>>>(l.cmds, l.cmd_dev, l.maps, l.row_chunks, l.num_row_chunks, l.spans, l.shadow_maps);
	return cudaGetLastError();
}

} // namespace sqw
