// Device-side frame set-up of the rasteriser (the commands never visit the host).
//
// sqw_frame_set_draw builds, on the host, what the raster kernels need besides the command list itself: per command the row /
// column ranges, the span-table offset and the resolved texture reference (CmdDev), the list of 32-scan-line chunks the span
// set-up warps take, and per 8-row band the ordered list of the commands touching it.  When the command list was produced on the
// device (sqw_polygons_draw), sqw_frame_set_draw_from_polygons runs the same construction here:
//
//   draw_cmd_dev_kernel      one thread per command: vertex bounding box -> rows [y0, y1) and conservative columns, the one-step
//                            division flag (the same double-precision bound as the host), rows and chunk count of the command;
//   draw_scan_kernel         one CTA: exclusive prefix sums -> span_base of every command, first chunk of every command, totals;
//   draw_band_count / _fill  one warp per band: the commands whose rows touch the band, IN COMMAND ORDER (ballot / popc
//                            compaction over the command list, 32 at a time): painter's order per band, as the host builds it;
//   draw_band_scan_kernel    one warp: band list offsets;
//   draw_chunks_kernel       one thread per command: its chunk records.
//
// The host only reads four totals back (to size the span table and the tile lists of the view) and launches.
#include "sqw_device.cuh"
#include "sqw_internal.h"

namespace sqw
{
namespace
{

__global__ void __launch_bounds__(128) draw_cmd_dev_kernel(const DrawSetupLaunch l)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= l.num_cmds)
		return;
	const sqw_draw_cmd c = l.cmds[i];
	CmdDev d;
	d.kind = c.kind;
	d.blend = c.blend;
	d.tc_mode = c.tc_mode;
	d.tex_width = c.tex_size[0];
	d.tex_height = c.tex_size[1];
	int32_t min_x = INT32_MAX, max_x = INT32_MIN, min_y = INT32_MAX, max_y = INT32_MIN;
	const sqw_polygon_point* const pts = l.points + c.vertex_offset;
	for (uint32_t k = 0; k < c.num_vertices; ++k)
	{
		const sqw_polygon_point p = pts[k];
		min_x = min(min_x, p.x);
		max_x = max(max_x, p.x);
		min_y = min(min_y, p.y);
		max_y = max(max_y, p.y);
	}
	// the same bound as sqw_frame_set_draw: smallest inverse depth of the polygon lowered by two pixels of gradient, in the
	// rasteriser's 2^29 scale; when half of it stays above 2^11 << 8 the opaque tile kernel may use its one-step division
	uint8_t z_flag = 0;
	if (c.tc_mode == SQW_TC_FULL_PERSPECTIVE)
	{
		double min_iz = 1e300;
		for (uint32_t k = 0; k < c.num_vertices; ++k)
		{
			const sqw_polygon_point p = pts[k];
			min_iz = fmin(min_iz, (double)c.depth_eq[0] * (p.x / 65536.0) + (double)c.depth_eq[1] * (p.y / 65536.0) + (double)c.depth_eq[2]);
		}
		min_iz -= 2.0 * (fabs((double)c.depth_eq[0]) + fabs((double)c.depth_eq[1]));
		if (min_iz * 536870912.0 * 0.5 >= 2048.0 * 256.0)
			z_flag = SQW_CMDF_Z_ONE_STEP;
	}
	// polygon textures have the framebuffer's colour type; the list holds surface references only (sqw_polygons_draw)
	d.texture = reinterpret_cast<const void*>((uintptr_t)c.tex_ref * l.bytes_per_pixel);
	d.tex_flags = (uint8_t)((l.bytes_per_pixel == 8 ? 1 : 0) | 2 | z_flag);
	const uint64_t tex_elems = (uint64_t)c.tex_size[0] * (uint64_t)c.tex_size[1];
	if (c.kind != SQW_CMD_POLYGON || c.tex_kind != SQW_TEX_SURFACE || c.num_vertices < 3 || c.tex_size[0] <= 0 || c.tex_size[1] <= 0 ||
		(uint64_t)c.tex_ref + tex_elems > l.surface_pool_elements)
		atomicOr(l.totals + 4, 1u); // a command this path cannot take: the host reports the error

	// rows that can own a span: [round(min_y), round(max_y)) clipped (rasterizer.rs:316-317)
	int32_t y0 = max(fixed16_round_to_int(min_y), 0);
	int32_t y1 = min(fixed16_round_to_int(max_y), l.height);
	int32_t bx0 = max((min_x >> 16) - 1, 0);
	int32_t bx1 = min((max_x >> 16) + 2, l.width);
	if (y0 >= y1 || bx0 >= bx1)
		y0 = y1 = bx0 = bx1 = 0;
	d.y0 = (int16_t)y0;
	d.y1 = (int16_t)y1;
	d.bx0 = (int16_t)bx0;
	d.bx1 = (int16_t)bx1;
	d.span_base = 0;
	l.cmd_dev[i] = d;
	l.rows[i] = (uint32_t)(y1 - y0);
	if (c.blend != SQW_BLEND_NONE)
		atomicAdd(l.totals + 3, 1u);
}

// span_base[i] = rows of the commands before i; chunk_base[i] = chunks (32 scan-lines) of the commands before i
__global__ void __launch_bounds__(1024) draw_scan_kernel(CmdDev* __restrict__ cmd_dev, const uint32_t* __restrict__ rows, uint32_t* __restrict__ chunk_base, const uint32_t n, uint32_t* __restrict__ totals)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t r = i < n ? rows[i] : 0u;
		// rows in the low word, chunks in the high word: one scan for both
		const unsigned long long own = (unsigned long long)r | ((unsigned long long)((r + 31u) / 32u) << 32);
		unsigned long long incl = own;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= o)
				incl += t;
		}
		if (lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		unsigned long long before = s_carry, all = 0;
		for (int w = 0; w < 32; ++w)
		{
			before += w < warp ? s_warp[w] : 0ull;
			all += s_warp[w];
		}
		if (i < n)
		{
			const unsigned long long excl = before + incl - own;
			cmd_dev[i].span_base = (uint32_t)excl;
			chunk_base[i] = (uint32_t)(excl >> 32);
		}
		__syncthreads();
		if (threadIdx.x == 0)
			s_carry += all;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals[0] = (uint32_t)s_carry;         // scan-lines (the host rejects frames past 2^32 before they get here: rows <= height per command)
		totals[1] = (uint32_t)(s_carry >> 32); // chunks
	}
}

__global__ void __launch_bounds__(128) draw_chunks_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ chunk_base, const uint32_t n, uint2* __restrict__ chunks)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint32_t r = rows[i];
	uint2* o = chunks + chunk_base[i];
	for (uint32_t r0 = 0; r0 < r; r0 += 32)
		*o++ = make_uint2(i, r0);
}

// One warp per 8-row band: FILL = 0 counts the commands touching the band, FILL = 1 writes them in command order.
template <int FILL>
__global__ void __launch_bounds__(128) draw_band_lists_kernel(const CmdDev* __restrict__ cmd_dev, const uint32_t n, const uint32_t num_bands, uint32_t* __restrict__ band_counts, const uint32_t* __restrict__ band_offsets, uint32_t* __restrict__ band_list)
{
	const uint32_t band = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (band >= num_bands)
		return;
	const int lane = threadIdx.x & 31;
	const int32_t band_y0 = (int32_t)band * SQW_TILE_H, band_y1 = band_y0 + SQW_TILE_H;
	uint32_t count = 0;
	const uint32_t out0 = FILL ? band_offsets[band] : 0u;
	for (uint32_t base = 0; base < n; base += 32)
	{
		const uint32_t i = base + (uint32_t)lane;
		bool hit = false;
		if (i < n)
		{
			const uint32_t rows = __ldg(reinterpret_cast<const uint32_t*>(cmd_dev + i)); // y0 | y1 << 16
			const int32_t y0 = (int16_t)(rows & 0xFFFF), y1 = (int16_t)(rows >> 16);
			hit = y0 < y1 && y0 < band_y1 && y1 > band_y0;
		}
		const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
		if (FILL && hit)
			band_list[out0 + count + __popc(ballot & ((1u << lane) - 1))] = i;
		count += __popc(ballot);
	}
	if (!FILL && lane == 0)
		band_counts[band] = count;
}

__global__ void __launch_bounds__(32) draw_band_scan_kernel(const uint32_t* __restrict__ band_counts, const uint32_t num_bands, uint32_t* __restrict__ band_offsets, uint32_t* __restrict__ totals)
{
	const int lane = threadIdx.x;
	uint32_t carry = 0;
	for (uint32_t base = 0; base < num_bands; base += 32)
	{
		const uint32_t b = base + (uint32_t)lane;
		const uint32_t own = b < num_bands ? band_counts[b] : 0u;
		uint32_t incl = own;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= o)
				incl += t;
		}
		if (b < num_bands)
			band_offsets[b] = carry + incl - own;
		carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
	}
	if (lane == 0)
	{
		band_offsets[num_bands] = carry;
		totals[2] = carry;
	}
}


// ---- surface jobs and bands from the prepared polygons ----------------------------------------------------------------------------
// visible_rank[i] = visible polygons before candidate i (the job index: current_frame_visible_polygons is in candidate order)
__global__ void __launch_bounds__(1024) surface_rank_kernel(const sqw_prepared_polygon* __restrict__ prepared, const uint32_t n, uint32_t* __restrict__ visible_rank, unsigned long long* __restrict__ totals)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t own = (i < n && (prepared[i].flags & SQW_PREPARED_VISIBLE)) ? 1u : 0u;
		uint32_t incl = own;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if (lane >= o)
				incl += t;
		}
		if (lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		uint32_t before = s_carry, all = 0;
		for (int w = 0; w < 32; ++w)
		{
			before += w < warp ? s_warp[w] : 0u;
			all += s_warp[w];
		}
		if (i < n)
			visible_rank[i] = before + incl - own;
		__syncthreads();
		if (threadIdx.x == 0)
			s_carry += all;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		totals[0] = s_carry;
}

// get_lightmap_size (common/lightmap.rs:19-27)
__device__ __forceinline__ uint32_t lightmap_size_of(const int32_t tc_min, const int32_t tc_max) { return (uint32_t)(((tc_max + 15) >> 4) - (tc_min >> 4) + 1); }

// polygon_is_affected_by_light (partial_renderer.rs:3877-3985), operand order of the host mirror: the light's sphere against the polygon's
// plane and texture-coordinate rectangle, or, for a projector, the polygon's corners against the five planes of its pyramid.
__device__ bool polygon_is_affected_by_light(const sqw_map_polygon& polygon, const float* __restrict__ basis, const sqw_light& light)
{
	const float bu[3] = {basis[0], basis[1], basis[2]}, bv[3] = {basis[3], basis[4], basis[5]}, bs[3] = {basis[6], basis[7], basis[8]}, bn[3] = {basis[9], basis[10], basis[11]};
	const float d[3] = {light.position[0] - bs[0], light.position[1] - bs[1], light.position[2] - bs[2]};
	auto dot3 = [](const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
	const float signed_distance = dot3(d, bn);
	if (signed_distance <= 0.0f)
		return false;
	const float square_radius_at_plane = light.radius * light.radius - signed_distance * signed_distance;
	if (square_radius_at_plane <= 0.0f)
		return false;
	if (light.shadow_type == SQW_SHADOW_PROJECTOR)
	{
		const float tmin0 = (float)polygon.tex_coord_min[0], tmax0 = (float)polygon.tex_coord_max[0], tmin1 = (float)polygon.tex_coord_min[1], tmax1 = (float)polygon.tex_coord_max[1];
		float vertices[4][3];
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			const float u0 = bu[k] * tmin0, u1 = bu[k] * tmax0, v0 = bv[k] * tmin1, v1 = bv[k] * tmax1;
			vertices[0][k] = (bs[k] + u0) + v0;
			vertices[1][k] = (bs[k] + u0) + v1;
			vertices[2][k] = (bs[k] + u1) + v0;
			vertices[3][k] = (bs[k] + u1) + v1;
		}
		for (int p = 0; p < 5; ++p)
		{
			float plane_vec[3];
#pragma unroll
			for (int k = 0; k < 3; ++k)
				plane_vec[k] = p == 0 ? light.basis_z[k] : (p == 1 ? light.basis_z[k] - light.basis_x[k] : (p == 2 ? light.basis_z[k] + light.basis_x[k] : (p == 3 ? light.basis_z[k] - light.basis_y[k] : light.basis_z[k] + light.basis_y[k])));
			const float dist = dot3(plane_vec, light.position);
			int vertices_front = 0;
			for (int v = 0; v < 4; ++v)
				if (dot3(plane_vec, vertices[v]) >= dist)
					++vertices_front;
			if (vertices_front >= 4)
				return false;
		}
		return true;
	}
	// texture coordinates of the projection of the light onto the polygon plane: rows of inverse(cols(u, v, n))
	const float vxn[3] = {bv[1] * bn[2] - bv[2] * bn[1], bv[2] * bn[0] - bv[0] * bn[2], bv[0] * bn[1] - bv[1] * bn[0]};
	const float nxu[3] = {bn[1] * bu[2] - bn[2] * bu[1], bn[2] * bu[0] - bn[0] * bu[2], bn[0] * bu[1] - bn[1] * bu[0]};
	const float det = dot3(bu, vxn);
	if (det == 0.0f || !isfinite(det))
		return false;
	const float mat_x[3] = {vxn[0] / det, vxn[1] / det, vxn[2] / det}, mat_y[3] = {nxu[0] / det, nxu[1] / det, nxu[2] / det};
	const float tc_u = dot3(d, mat_x), tc_v = dot3(d, mat_y);
	const float radius_at_plane = sqrtf(square_radius_at_plane);
	const float u_radius = radius_at_plane * (1.0f / sqrtf(dot3(bu, bu)));
	const float v_radius = radius_at_plane * (1.0f / sqrtf(dot3(bv, bv)));
	return !(tc_u + u_radius < (float)polygon.tex_coord_min[0] || tc_v + v_radius < (float)polygon.tex_coord_min[1] || tc_u - u_radius > (float)polygon.tex_coord_max[0] ||
			 tc_v - v_radius > (float)polygon.tex_coord_max[1]);
}

// One thread per candidate: the job of a visible polygon (the closure of build_polygons_surfaces; dynamic lights when the frame has them).
__global__ void __launch_bounds__(128) surface_jobs_kernel(const SurfaceSetupLaunch l)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= l.num_candidates)
		return;
	const sqw_prepared_polygon rec = l.prepared[i];
	if (!(rec.flags & SQW_PREPARED_VISIBLE))
		return;
	const uint32_t polygon_index = l.candidates[i].polygon;
	const sqw_map_polygon& polygon = l.polygons[polygon_index];
	const sqw_map_material& material = l.materials[polygon.texture];
	const float* const basis = l.polygon_basis + (size_t)polygon_index * 12;
	sqw_surface_job j;
	memset(&j, 0, sizeof(j));
	const int32_t shift0 = material.texture_shift[0] >> rec.mip, shift1 = material.texture_shift[1] >> rec.mip;
	j.surface_size[0] = rec.surface_size[0];
	j.surface_size[1] = rec.surface_size[1];
	j.tc_start[0] = rec.surface_tc_min[0] + shift0;
	j.tc_start[1] = rec.surface_tc_min[1] + shift1;
	j.surface_pixels_offset = rec.surface_pixels_offset;
	const unsigned long long texels = (unsigned long long)rec.surface_size[0] * rec.surface_size[1];
	const bool diffuse = (material.flags & SQW_MAT_LIGHT) && (material.flags & SQW_MAT_HAS_DIFFUSE);
	if (diffuse)
	{
		j.flags |= SQW_SURF_APPLY_DIFFUSE;
		// get_basis_vecs_for_mip (surfaces.rs:49-63) + shift correction (partial_renderer.rs:1736-1740)
		const float scale = (float)(1 << rec.mip);
		float u[3], v[3];
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			u[k] = rec.mip == 0 ? basis[k] : basis[k] * scale;
			v[k] = rec.mip == 0 ? basis[3 + k] : basis[3 + k] * scale;
		}
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			j.basis_u[k] = u[k];
			j.basis_v[k] = v[k];
			j.basis_start[k] = basis[6 + k] - (u[k] * (float)shift0 + v[k] * (float)shift1);
			j.basis_normal[k] = basis[9 + k];
		}
		j.texture_image = material.texture_images[rec.mip];
		if (material.flags & SQW_MAT_HAS_NORMAL_MAP)
			j.flags |= SQW_SURF_HAS_NORMAL_MAP;
		if (material.flags & SQW_MAT_HAS_NON_ONE_ROUGHNESS)
			j.flags |= SQW_SURF_HAS_NON_ONE_ROUGHNESS;
		if (material.flags & SQW_MAT_IS_METAL)
			j.flags |= SQW_SURF_IS_METAL;
#pragma unroll
		for (int k = 0; k < 2; ++k) // :1744-1752
			j.lightmap_tc_shift[k] = (uint32_t)(rec.surface_tc_min[k] - ((polygon.tex_coord_min[k] & ~15) >> rec.mip));
		j.lightmap_size[0] = lightmap_size_of(polygon.tex_coord_min[0], polygon.tex_coord_max[0]);
		j.lightmap_size[1] = lightmap_size_of(polygon.tex_coord_min[1], polygon.tex_coord_max[1]);
		j.lightmap_scale_log2 = 4u - rec.mip;
		j.lightmap_offset = polygon.lightmap_data_offset;
		if (l.use_directional)
			j.flags |= SQW_SURF_DIRECTIONAL_LIGHTMAP;
		// lights of the polygon's leaf that reach it, at most SQW_MAX_POLYGON_LIGHTS, in list order (partial_renderer.rs:1705-1734)
		if (l.lights)
		{
			const uint32_t set = l.candidates[i].clip_set;
			for (uint32_t k = l.set_light_offsets[set]; k < l.set_light_offsets[set + 1]; ++k)
			{
				const uint32_t light_index = l.set_lights[k];
				if (polygon_is_affected_by_light(polygon, basis, l.lights[light_index]))
				{
					j.lights[j.num_lights++] = (uint16_t)light_index;
					if (j.num_lights == SQW_MAX_POLYGON_LIGHTS)
						break;
				}
			}
		}
		if (j.lightmap_offset != 0)
			atomicAdd(l.totals + 6, (unsigned long long)j.lightmap_size[0] * j.lightmap_size[1] * (l.use_directional ? 40ull : 12ull));
	}
	if (material.flags & SQW_MAT_HAS_EMISSIVE)
	{
		j.flags |= SQW_SURF_EMISSIVE;
		j.emissive_image = material.emissive_images[rec.mip];
		j.emissive_light[0] = material.emissive_light[0];
		j.emissive_light[1] = material.emissive_light[1];
		j.emissive_light[2] = material.emissive_light[2];
		if (texels)
			atomicAdd(l.totals + 5, texels);
	}
	l.jobs[l.visible_rank[i]] = j;
	if ((j.flags & (SQW_SURF_APPLY_DIFFUSE | SQW_SURF_EMISSIVE)) && texels)
	{
		atomicAdd(l.totals + 4, texels);
		if (texels <= SQW_SURF_WARP_JOB_TEXELS)
			atomicAdd(l.totals + 1, texels);
	}
}

// Band geometry of one job, as sqw_frame_set_surfaces cuts it: rows per band, and whether the job goes to warp bands.
__device__ __forceinline__ uint32_t surface_band_rows(const sqw_surface_job& j, const uint32_t warp_band_texels, bool& warp_bands)
{
	const unsigned long long texels = (unsigned long long)j.surface_size[0] * j.surface_size[1];
	warp_bands = texels <= SQW_SURF_WARP_JOB_TEXELS;
	uint32_t band_rows;
	if (warp_bands)
	{
		band_rows = max(warp_band_texels / j.surface_size[0], 1u);
		if ((j.flags & SQW_SURF_APPLY_DIFFUSE) && j.lightmap_offset != 0)
			band_rows = min(band_rows, max((uint32_t)SQW_SURF_WARP_LM_ELEMENTS / min(j.lightmap_size[0], 19u), 1u));
		band_rows = min(band_rows, (uint32_t)SQW_SURF_MAX_BAND_ROWS);
	}
	else
		band_rows = min(max((uint32_t)SQW_SURF_BAND_TEXELS / j.surface_size[0], 1u), (uint32_t)SQW_SURF_MAX_BAND_ROWS);
	return band_rows;
}
__device__ __forceinline__ uint32_t surface_warp_band_texels(const unsigned long long small_texels)
{
	uint32_t t = SQW_SURF_WARP_BAND_TEXELS;
	const unsigned long long resident_warps = 148ull * 4 * (SQW_SURF_THREADS / 32);
	while (t > SQW_SURF_WARP_BAND_TEXELS_MIN && small_texels < resident_warps * t)
		t /= 2;
	return t;
}
__device__ __forceinline__ uint32_t surface_variant_bits(const sqw_surface_job& j)
{
	if (!(j.flags & SQW_SURF_APPLY_DIFFUSE))
		return 16u; // emissive-only
	const uint32_t spec = (j.flags & SQW_SURF_HAS_NON_ONE_ROUGHNESS) ? ((j.flags & SQW_SURF_IS_METAL) ? 2u : 1u) : 0u;
	return ((j.flags & SQW_SURF_DIRECTIONAL_LIGHTMAP) ? 1u : 0u) | ((j.flags & SQW_SURF_HAS_NORMAL_MAP) ? 2u : 0u) | (spec << 2);
}

// FILL = 0: bands of every job (CTA kind, warp kind); FILL = 1: the band records at the scanned offsets.
template <int FILL>
__global__ void __launch_bounds__(128) surface_bands_kernel(const sqw_surface_job* __restrict__ jobs, const uint32_t num_jobs, const unsigned long long* __restrict__ totals, uint32_t* __restrict__ band_counts, const uint32_t* __restrict__ band_base, SurfaceBand* __restrict__ bands)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= num_jobs)
		return;
	const sqw_surface_job j = jobs[i];
	const unsigned long long texels = (unsigned long long)j.surface_size[0] * j.surface_size[1];
	uint32_t n = 0;
	bool warp_bands = false;
	uint32_t band_rows = 1;
	if ((j.flags & (SQW_SURF_APPLY_DIFFUSE | SQW_SURF_EMISSIVE)) && texels)
	{
		band_rows = surface_band_rows(j, surface_warp_band_texels(totals[1]), warp_bands);
		n = (j.surface_size[1] + band_rows - 1) / band_rows;
	}
	const uint32_t kind = (j.num_lights ? 2u : 0u) + (warp_bands ? 1u : 0u);
	if (!FILL)
	{
#pragma unroll
		for (uint32_t k = 0; k < 4; ++k)
			band_counts[4 * i + k] = k == kind ? n : 0u;
		return;
	}
	// the four kinds one after the other: CTA, warp (no dynamic lights), CTA, warp (dynamic lights)
	const uint32_t kind_base[4] = {0u, (uint32_t)totals[2], (uint32_t)(totals[2] + totals[3]), (uint32_t)(totals[2] + totals[3] + (totals[7] & 0xFFFFFFFFull))};
	SurfaceBand* o = bands + kind_base[kind] + band_base[4 * i + kind];
	const uint32_t variant = surface_variant_bits(j);
	for (uint32_t row0 = 0; row0 < j.surface_size[1]; row0 += band_rows)
		*o++ = SurfaceBand{i, row0, min(band_rows, j.surface_size[1] - row0), variant};
}

// exclusive scans of the four band counts of every job (one CTA): totals[2], [3] = CTA / warp bands without dynamic lights,
// totals[7] = CTA | warp << 32 with
__global__ void __launch_bounds__(1024) surface_band_scan_kernel(const uint32_t* __restrict__ band_counts, const uint32_t num_jobs, uint32_t* __restrict__ band_base, unsigned long long* __restrict__ totals)
{
	__shared__ unsigned long long s_warp[2][32];
	__shared__ unsigned long long s_carry[2];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 2)
		s_carry[threadIdx.x] = 0;
	__syncthreads();
	for (uint32_t base = 0; base < num_jobs; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		unsigned long long own[2] = {0ull, 0ull}, incl[2];
		if (i < num_jobs)
		{
			own[0] = (unsigned long long)band_counts[4 * i] | ((unsigned long long)band_counts[4 * i + 1] << 32);
			own[1] = (unsigned long long)band_counts[4 * i + 2] | ((unsigned long long)band_counts[4 * i + 3] << 32);
		}
		incl[0] = own[0];
		incl[1] = own[1];
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long t0 = __shfl_up_sync(0xFFFFFFFFu, incl[0], o), t1 = __shfl_up_sync(0xFFFFFFFFu, incl[1], o);
			if (lane >= o)
			{
				incl[0] += t0;
				incl[1] += t1;
			}
		}
		if (lane == 31)
		{
			s_warp[0][warp] = incl[0];
			s_warp[1][warp] = incl[1];
		}
		__syncthreads();
		unsigned long long before[2] = {s_carry[0], s_carry[1]}, all[2] = {0ull, 0ull};
		for (int w = 0; w < 32; ++w)
		{
			before[0] += w < warp ? s_warp[0][w] : 0ull;
			before[1] += w < warp ? s_warp[1][w] : 0ull;
			all[0] += s_warp[0][w];
			all[1] += s_warp[1][w];
		}
		if (i < num_jobs)
		{
			const unsigned long long e0 = before[0] + incl[0] - own[0], e1 = before[1] + incl[1] - own[1];
			band_base[4 * i] = (uint32_t)e0;
			band_base[4 * i + 1] = (uint32_t)(e0 >> 32);
			band_base[4 * i + 2] = (uint32_t)e1;
			band_base[4 * i + 3] = (uint32_t)(e1 >> 32);
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			s_carry[0] += all[0];
			s_carry[1] += all[1];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals[2] = s_carry[0] & 0xFFFFFFFFull;
		totals[3] = s_carry[0] >> 32;
		totals[7] = s_carry[1];
	}
}

} // namespace

// Phase 1: everything whose size is known (CmdDev, rows, prefix sums, band counts and offsets); totals = {scan-lines, chunks, band
// list entries, commands that are not opaque polygons, error flag}.
cudaError_t launch_draw_setup_phase1(const DrawSetupLaunch& l, cudaStream_t stream)
{
	cudaMemsetAsync(l.totals, 0, 8 * sizeof(uint32_t), stream);
	if (l.num_cmds)
	{
		draw_cmd_dev_kernel<<<(l.num_cmds + 127) / 128, 128, 0, stream>>>(l);
		draw_scan_kernel<<<1, 1024, 0, stream>>>(l.cmd_dev, l.rows, l.chunk_base, l.num_cmds, l.totals);
	}
	draw_band_lists_kernel<0><<<(l.num_bands + 3) / 4, 128, 0, stream>>>(l.cmd_dev, l.num_cmds, l.num_bands, l.band_counts, nullptr, nullptr);
	draw_band_scan_kernel<<<1, 32, 0, stream>>>(l.band_counts, l.num_bands, l.band_offsets, l.totals);
	return cudaGetLastError();
}

// Phase 2, once the host has sized the lists: chunk records and the ordered band lists.
cudaError_t launch_draw_setup_phase2(const DrawSetupLaunch& l, cudaStream_t stream)
{
	if (l.num_cmds)
	{
		draw_chunks_kernel<<<(l.num_cmds + 127) / 128, 128, 0, stream>>>(l.rows, l.chunk_base, l.num_cmds, l.chunks);
		draw_band_lists_kernel<1><<<(l.num_bands + 3) / 4, 128, 0, stream>>>(l.cmd_dev, l.num_cmds, l.num_bands, nullptr, l.band_offsets, l.band_list);
	}
	return cudaGetLastError();
}

// Surfaces, phase 1: jobs of the visible polygons and their band counts; totals = {visible polygons, small texels, CTA bands, warp bands,
// surface texels, emissive texels, lightmap bytes}.
cudaError_t launch_surface_setup_phase1(const SurfaceSetupLaunch& l, cudaStream_t stream)
{
	cudaMemsetAsync(l.totals, 0, 8 * sizeof(unsigned long long), stream);
	if (l.num_candidates == 0)
		return cudaGetLastError();
	surface_rank_kernel<<<1, 1024, 0, stream>>>(l.prepared, l.num_candidates, l.visible_rank, l.totals);
	surface_jobs_kernel<<<(l.num_candidates + 127) / 128, 128, 0, stream>>>(l);
	// (jobs are at most num_candidates: the kernels below take the visible count from the device)
	return cudaGetLastError();
}
// Surfaces, phase 1b (the host knows the number of jobs now) and phase 2 (it has sized the band list).
cudaError_t launch_surface_bands_count(const SurfaceSetupLaunch& l, uint32_t num_jobs, cudaStream_t stream)
{
	if (num_jobs == 0)
		return cudaGetLastError();
	surface_bands_kernel<0><<<(num_jobs + 127) / 128, 128, 0, stream>>>(l.jobs, num_jobs, l.totals, l.band_counts, nullptr, nullptr);
	surface_band_scan_kernel<<<1, 1024, 0, stream>>>(l.band_counts, num_jobs, l.band_base, l.totals);
	return cudaGetLastError();
}
cudaError_t launch_surface_setup_phase2(const SurfaceSetupLaunch& l, uint32_t num_jobs, cudaStream_t stream)
{
	if (num_jobs)
		surface_bands_kernel<1><<<(num_jobs + 127) / 128, 128, 0, stream>>>(l.jobs, num_jobs, l.totals, nullptr, l.band_base, l.bands);
	return cudaGetLastError();
}

} // namespace sqw
