// World-polygon preparation on sm_100a (the geometry half of the host front end).
//
// Replaces the loop body of PartialRenderer::prepare_polygons_surfaces (src/square_wheel/partial_renderer.rs:1327-1360), i.e.
// prepare_polygon_surface (:1464-1638), for the polygons of the leafs the host's visibility flood marked visible:
//
//   polygons_prepare_kernel   one thread per candidate (polygon, leaf clip planes): back-face cull, material flags, vertex
//                             transform, z-near clip (clipping.rs:4-57), projection, the eight half-plane clips of the leaf's
//                             octagon (clipping.rs:132-185), sky bounds, depth / texture-coordinate equations
//                             (equations.rs:22-75), mip (renderer_utils.rs:201-242), surface rectangle, equation shift;
//   polygons_offsets_kernel   one CTA: the reference's running `surface_pixels_offset` (:1617-1618) is a prefix sum over the
//                             visible polygons in candidate order; also the frame totals.
// and draw_polygon (:4011-4142) for the same polygons in the order of the host's BSP walk (draw_tree / draw_leaf, :2450-2565):
//   polygons_draw_kernel      one thread per polygon: clipping again (the reference clips twice too), inverse-depth extremes,
//                             16.16 vertices, the interpolation-mode predicates (renderer_utils.rs:60-197) and the power-of-two
//                             renormalisation of the equations -> one sqw_draw_cmd + its vertices;
//   polygons_draw_offsets / _compact   command index and vertex offset by a prefix sum, packed lists in draw order — the
//                             arrays sqw_draw / sqw_frame_set_draw take.
//
// The sky window (:1511-1529) is the min / max of four linear forms over every clipped sky vertex: min and max are exact and
// associative, so device atomics give the bits of the host's sequential ClippingPolygon::extend.
//
// Every float expression keeps the operand order of the host mirror (square_wheel_b200/host/renderer.cpp, math.hpp) and the file
// is compiled with -fmad=false -prec-div=true, so the records are bit-identical to it (tests/test_polygons_gpu.py).  The one
// library call, log2 in calculate_mip, is evaluated in double and rounded to float: the correctly rounded result, which is what
// glibc's log2f returns for all but a vanishing set of arguments (its error bound is 2^-26 ulp of the float before rounding).
#include "sqw_device.cuh"
#include "sqw_internal.h"

namespace sqw
{
namespace
{

constexpr int MAX_POLY_VERTICES = 24; // renderer_utils.rs:4
constexpr float POLY_Z_NEAR = 1.0f;   // renderer_utils.rs:7
constexpr int POLY_MAX_MIP = 4;       // textures.rs:6
constexpr int POLY_LIGHTMAP_SCALE = 16; // common/lightmap.rs:4

struct P2
{
	float x, y;
};
struct P3
{
	float x, y, z;
};
struct P4
{
	float x, y, z, w;
};

// cgmath: m * v = sum of the columns scaled by the components, added left to right
__device__ __forceinline__ P4 mat_mul(const float* __restrict__ m, const float vx, const float vy, const float vz, const float vw)
{
	P4 r;
	r.x = ((m[0] * vx + m[4] * vy) + m[8] * vz) + m[12] * vw;
	r.y = ((m[1] * vx + m[5] * vy) + m[9] * vz) + m[13] * vw;
	r.z = ((m[2] * vx + m[6] * vy) + m[10] * vz) + m[14] * vw;
	r.w = ((m[3] * vx + m[7] * vy) + m[11] * vz) + m[15] * vw;
	return r;
}

// the `as i32` of the host mirror's x86 build for floats that may be out of range (cvttss2si: the indefinite value)
__device__ __forceinline__ int32_t f32_to_i32_x86(const float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)f : INT32_MIN; }

// v0 * k1 - v1 * k0 with k = dist / (dist1 - dist0)  (clipping.rs:122-130, :187-195)
__device__ __forceinline__ P3 z_intersection(const P3 v0, const P3 v1, const float z)
{
	const float dist0 = v0.z - z, dist1 = v1.z - z;
	const float dist_sum = v1.z - v0.z;
	const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
	return P3{v0.x * k1 - v1.x * k0, v0.y * k1 - v1.y * k0, v0.z * k1 - v1.z * k0};
}
__device__ __forceinline__ P2 line_intersection(const P2 v0, const P2 v1, const float lx, const float ly, const float lz)
{
	const float dist0 = (v0.x * lx + v0.y * ly) - lz, dist1 = (v1.x * lx + v1.y * ly) - lz;
	const float dist_sum = dist1 - dist0;
	const float k0 = dist0 / dist_sum, k1 = dist1 / dist_sum;
	return P2{v0.x * k1 - v1.x * k0, v0.y * k1 - v1.y * k0};
}

// Sutherland-Hodgman step against one half-plane dot(v, l.xy) - l.z >= 0, output capped like the reference's fixed arrays.
__device__ __forceinline__ int clip_by_line(const P2* __restrict__ in, const int n, const float lx, const float ly, const float lz, P2* __restrict__ out)
{
	P2 prev = in[n - 1];
	float prev_dist = (prev.x * lx + prev.y * ly) - lz;
	int count = 0;
	for (int i = 0; i < n; ++i)
	{
		const P2 v = in[i];
		const float dist = (v.x * lx + v.y * ly) - lz;
		if (dist > 0.0f)
		{
			if (prev_dist < 0.0f)
			{
				out[count++] = line_intersection(prev, v, lx, ly, lz);
				if (count == MAX_POLY_VERTICES)
					break;
			}
			out[count++] = v;
			if (count == MAX_POLY_VERTICES)
				break;
		}
		else if (dist == 0.0f)
		{
			out[count++] = v;
			if (count == MAX_POLY_VERTICES)
				break;
		}
		else if (prev_dist > 0.0f)
		{
			out[count++] = line_intersection(prev, v, lx, ly, lz);
			if (count == MAX_POLY_VERTICES)
				break;
		}
		prev = v;
		prev_dist = dist;
	}
	return count;
}

// view_matrix transform of the polygon's vertices (matrix.rs:125), clip_3d_polygon_by_z_plane (clipping.rs:4-57), projection and the
// eight half-plane clips of project_and_clip_polygon (renderer_utils.rs:10-58).  Returns the vertex count (< 3: nothing left); the
// clipped polygon is in `va`.
__device__ __forceinline__ int transform_clip_project(
	const sqw_map_polygon& polygon, const float* __restrict__ vertices, const float* __restrict__ vm, const float* __restrict__ planes, P2* __restrict__ va, P2* __restrict__ vb)
{
	P3 zc[MAX_POLY_VERTICES];
	int count = 0;
	{
		const int n = (int)min(polygon.num_vertices, (uint32_t)MAX_POLY_VERTICES);
		const float* const vsrc = vertices + (size_t)polygon.first_vertex * 3;
		auto transformed = [&](int i) {
			const P4 t = mat_mul(vm, vsrc[i * 3], vsrc[i * 3 + 1], vsrc[i * 3 + 2], 1.0f);
			return P3{t.x, t.y, t.z};
		};
		P3 prev = transformed(n - 1);
		for (int i = 0; i < n; ++i)
		{
			const P3 v = transformed(i);
			if (v.z > POLY_Z_NEAR)
			{
				if (prev.z < POLY_Z_NEAR)
				{
					zc[count++] = z_intersection(prev, v, POLY_Z_NEAR);
					if (count == MAX_POLY_VERTICES)
						break;
				}
				zc[count++] = v;
				if (count == MAX_POLY_VERTICES)
					break;
			}
			else if (v.z == POLY_Z_NEAR)
			{
				zc[count++] = v;
				if (count == MAX_POLY_VERTICES)
					break;
			}
			else if (prev.z > POLY_Z_NEAR)
			{
				zc[count++] = z_intersection(prev, v, POLY_Z_NEAR);
				if (count == MAX_POLY_VERTICES)
					break;
			}
			prev = v;
		}
	}
	if (count < 3)
		return count;
	for (int i = 0; i < count; ++i)
		va[i] = P2{zc[i].x / zc[i].z, zc[i].y / zc[i].z};
	for (int i = 0; i < 4 && count >= 3; ++i)
	{
		count = clip_by_line(va, count, planes[i * 6], planes[i * 6 + 1], planes[i * 6 + 2], vb);
		if (count < 3)
			break;
		count = clip_by_line(vb, count, planes[i * 6 + 3], planes[i * 6 + 4], planes[i * 6 + 5], va);
	}
	return count;
}

// float min / max as integer atomics on the bit pattern (non-negative floats order like ints, negative ones in reverse)
__device__ __forceinline__ void atomic_min_f32(float* addr, float v)
{
	v = __fadd_rn(v, 0.0f); // -0.0 -> +0.0
	if (v >= 0.0f)
		atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
	else
		atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* addr, float v)
{
	v = __fadd_rn(v, 0.0f);
	if (v >= 0.0f)
		atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
	else
		atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(128) polygons_prepare_kernel(const PolygonsLaunch l)
{
	const uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
	if (ci >= l.num_candidates)
		return;
	sqw_prepared_polygon out;
	out.flags = 0;
	out.mip = 0;
	out.depth_equation[0] = out.depth_equation[1] = out.depth_equation[2] = 0.0f;
#pragma unroll
	for (int i = 0; i < 6; ++i)
		out.tex_coord_equation[i] = 0.0f;
	out.surface_tc_min[0] = out.surface_tc_min[1] = 0;
	out.surface_size[0] = out.surface_size[1] = 0;
	out.surface_pixels_offset = 0;
	sqw_prepared_polygon* const dst = l.out + ci;

	const sqw_polygon_candidate cand = l.candidates[ci];
	const sqw_map_polygon& polygon = l.polygons[cand.polygon];
	const float* const pm = l.frame.planes_matrix;
	const float* const vm = l.frame.view_matrix;

	// :1477-1482 back faces
	const P4 plane_transformed = mat_mul(pm, polygon.plane[0], polygon.plane[1], polygon.plane[2], -polygon.plane[3]);
	const uint32_t material = polygon.texture;
	const uint32_t mflags = material < l.num_materials ? l.material_flags[material] : 0u;
	if (plane_transformed.w <= 0.0f || !(mflags & SQW_MATERIAL_DRAW))
	{
		*dst = out;
		return;
	}

	// :1494-1505 transform, clip by the z-near plane, project, clip by the leaf's octagon
	P2 va[MAX_POLY_VERTICES], vb[MAX_POLY_VERTICES];
	const int count = transform_clip_project(polygon, l.vertices, vm, l.clip_planes + (size_t)cand.clip_set * 24, va, vb);
	if (count < 3)
	{
		*dst = out;
		return;
	}
	const P2* const vertices_2d = va;

	if (mflags & SQW_MATERIAL_SKYBOX) // :1511-1529
	{
		PolygonsTotalsDev* const t = l.totals;
		for (int i = 0; i < count; ++i)
		{
			const P2 p = vertices_2d[i];
			const float s = p.x + p.y, d = p.x - p.y;
			// (a -0.0 would order below every negative value as an integer: the helpers canonicalise it)
			atomic_min_f32(&t->sky_bounds[0], p.x);
			atomic_max_f32(&t->sky_bounds[1], p.x);
			atomic_min_f32(&t->sky_bounds[2], p.y);
			atomic_max_f32(&t->sky_bounds[3], p.y);
			atomic_min_f32(&t->sky_bounds[4], s);
			atomic_max_f32(&t->sky_bounds[5], s);
			atomic_min_f32(&t->sky_bounds[6], d);
			atomic_max_f32(&t->sky_bounds[7], d);
		}
		atomicMin(&t->first_sky_candidate, ci);
		out.flags = SQW_PREPARED_SKY;
		*dst = out;
		return;
	}

	// equations.rs:22-31, :53-75
	const float plane_transformed_w = -plane_transformed.w;
	const float de0 = plane_transformed.x / plane_transformed_w, de1 = plane_transformed.y / plane_transformed_w, de2 = plane_transformed.z / plane_transformed_w;
	float te[6];
#pragma unroll
	for (int i = 0; i < 2; ++i)
	{
		const float* const e = polygon.tex_coord_equation[i];
		const P4 b = mat_mul(pm, e[0], e[1], e[2], e[3]);
		te[i] = b.x + b.w * de0;
		te[2 + i] = b.y + b.w * de1;
		te[4 + i] = b.z + b.w * de2;
	}

	// calculate_mip (renderer_utils.rs:201-242): at the point of the largest inverse depth
	uint32_t mip;
	{
		P2 mip_point = vertices_2d[0];
		float mip_point_inv_z = 0.0f;
		for (int i = 0; i < count; ++i)
		{
			const float inv_z = (de0 * vertices_2d[i].x + de1 * vertices_2d[i].y) + de2;
			if (inv_z > mip_point_inv_z)
			{
				mip_point_inv_z = inv_z;
				mip_point = vertices_2d[i];
			}
		}
		const float z_2 = 1.0f / (mip_point_inv_z * mip_point_inv_z);
		const float z_4 = z_2 * z_2;
		float d_tc_2[2];
#pragma unroll
		for (int i = 0; i < 2; ++i)
		{
			const float d_tc_dx = te[i] * (de2 + de1 * mip_point.y) - (te[4 + i] + te[2 + i] * mip_point.y) * de0;
			const float d_tc_dy = te[2 + i] * (de2 + de0 * mip_point.x) - (te[4 + i] + te[i] * mip_point.x) * de1;
			d_tc_2[i] = (d_tc_dx * d_tc_dx + d_tc_dy * d_tc_dy) * z_4;
		}
		const float max_d_tc_2 = fmaxf(d_tc_2[0], d_tc_2[1]);
		const float mip_f = (float)log2((double)max_d_tc_2) * 0.5f + l.frame.mip_bias;
		const float c = ceilf(mip_f);
		const int32_t ci32 = (c != c) ? 0 : (c >= 2147483648.0f ? INT32_MAX : (c <= -2147483648.0f ? INT32_MIN : (int32_t)c)); // Rust `as i32`
		mip = (uint32_t)max(0, min(ci32, POLY_MAX_MIP));
	}
	const float mip_scale = 1.0f / (float)(1 << mip);
	float tes[6];
#pragma unroll
	for (int i = 0; i < 6; ++i)
		tes[i] = te[i] * mip_scale;

	// :1558-1615 texture-coordinate range of the clipped polygon -> surface rectangle
	const float inf = (float)(1 << 29);
	const float max_z = (float)(1 << 16);
	float tc_min[2] = {inf, inf}, tc_max[2] = {-inf, -inf};
	for (int k = 0; k < count; ++k)
	{
		const P2 p = vertices_2d[k];
		const float inv_z = fmaxf((de0 * p.x + de1 * p.y) + de2, 1.0f / max_z);
		const float z = 1.0f / inv_z;
#pragma unroll
		for (int i = 0; i < 2; ++i)
		{
			const float tc = z * ((tes[i] * p.x + tes[2 + i] * p.y) + tes[4 + i]);
			if (tc < tc_min[i])
				tc_min[i] = tc;
			if (tc > tc_max[i])
				tc_max[i] = tc;
		}
	}
	int32_t surface_tc_min[2], surface_size[2];
#pragma unroll
	for (int i = 0; i < 2; ++i)
	{
		const float tc_reduce_eps = 1.0f / 32.0f;
		tc_min[i] += tc_reduce_eps;
		tc_max[i] -= tc_reduce_eps;
		const int32_t round_mask = ~(POLY_LIGHTMAP_SCALE - 1);
		const int32_t tc_min_round_down = (polygon.tex_coord_min[i] & round_mask) >> mip;
		const int32_t tc_max_round_up = ((polygon.tex_coord_max[i] + POLY_LIGHTMAP_SCALE - 1) & round_mask) >> mip;
		int32_t tc_min_int = max(f32_to_i32_x86(floorf(fmaxf(tc_min[i], -inf))), tc_min_round_down);
		int32_t tc_max_int = min(f32_to_i32_x86(ceilf(fminf(tc_max[i], inf))), tc_max_round_up);
		if (tc_min_int >= tc_max_int)
		{
			tc_min_int = min(tc_min_int, tc_max_round_up - 1);
			tc_max_int = tc_min_int + 1;
		}
		surface_tc_min[i] = tc_min_int;
		surface_size[i] = tc_max_int - tc_min_int;
	}

	out.flags = SQW_PREPARED_VISIBLE;
	out.mip = mip;
	out.depth_equation[0] = de0;
	out.depth_equation[1] = de1;
	out.depth_equation[2] = de2;
#pragma unroll
	for (int i = 0; i < 2; ++i) // :1629-1635 shift of the equation to the surface rectangle
	{
		const float tcm = (float)surface_tc_min[i];
		out.tex_coord_equation[i] = tes[i] - tcm * de0;
		out.tex_coord_equation[2 + i] = tes[2 + i] - tcm * de1;
		out.tex_coord_equation[4 + i] = tes[4 + i] - tcm * de2;
	}
	out.surface_tc_min[0] = surface_tc_min[0];
	out.surface_tc_min[1] = surface_tc_min[1];
	out.surface_size[0] = (uint32_t)surface_size[0];
	out.surface_size[1] = (uint32_t)surface_size[1];
	*dst = out;
}

// surface_pixels_offset of every visible polygon = texels of the visible polygons before it (candidate order), frame totals.
__global__ void __launch_bounds__(1024) polygons_offsets_kernel(sqw_prepared_polygon* __restrict__ out, const uint32_t n, PolygonsTotalsDev* __restrict__ totals)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ uint32_t s_warp_vis[32];
	__shared__ unsigned long long s_carry;
	__shared__ uint32_t s_carry_vis;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
	{
		s_carry = 0;
		s_carry_vis = 0;
	}
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		unsigned long long texels = 0;
		uint32_t vis = 0;
		if (i < n && (out[i].flags & SQW_PREPARED_VISIBLE))
		{
			texels = (unsigned long long)out[i].surface_size[0] * out[i].surface_size[1];
			vis = 1;
		}
		unsigned long long incl = texels;
		uint32_t incl_vis = vis;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			const uint32_t tv = __shfl_up_sync(0xFFFFFFFFu, incl_vis, o);
			if (lane >= o)
			{
				incl += t;
				incl_vis += tv;
			}
		}
		if (lane == 31)
		{
			s_warp[warp] = incl;
			s_warp_vis[warp] = incl_vis;
		}
		__syncthreads();
		unsigned long long before = s_carry, all = 0;
		uint32_t all_vis = 0;
		for (int w = 0; w < 32; ++w)
		{
			before += w < warp ? s_warp[w] : 0ull;
			all += s_warp[w];
			all_vis += s_warp_vis[w];
		}
		if (vis)
			out[i].surface_pixels_offset = (uint32_t)(before + incl - texels);
		__syncthreads();
		if (threadIdx.x == 0)
		{
			s_carry += all;
			s_carry_vis += all_vis;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals->num_surfaces_pixels = s_carry;
		totals->num_visible = s_carry_vis;
	}
}

__global__ void polygons_totals_init_kernel(PolygonsTotalsDev* totals)
{
	// ClippingPolygon::from_point of the first sky vertex followed by extend: the identity of min / max
	for (int i = 0; i < 4; ++i)
	{
		totals->sky_bounds[2 * i] = __int_as_float(0x7F800000);     // +inf
		totals->sky_bounds[2 * i + 1] = __int_as_float(0xFF800000); // -inf
	}
	totals->first_sky_candidate = 0xFFFFFFFFu;
	totals->num_visible = 0;
	totals->num_surfaces_pixels = 0;
}


// ---- draw_polygon (partial_renderer.rs:4011-4142) for the world polygons, in the order of the host's BSP walk ---------------------
// renderer_utils.rs:60-132
__device__ __forceinline__ bool affine_interpolation_may_be_used(const float* de, const float* te, const P2 min_inv_z_point, const P2 max_inv_z_point)
{
	const float TC_ERROR_THRESHOLD = 0.75f; // renderer_utils.rs:199
	const P2 edge{max_inv_z_point.x - min_inv_z_point.x, max_inv_z_point.y - min_inv_z_point.y};
	const float edge_square_len = edge.x * edge.x + edge.y * edge.y;
	if (edge_square_len == 0.0f)
		return true;
	const float edge_len = sqrtf(edge_square_len);
	const P2 en{edge.x / edge_len, edge.y / edge_len};
	const float inv_z_clamp = 1.0f / (float)(1 << 20);
	const float min_point_inv_z = fmaxf((de[0] * min_inv_z_point.x + de[1] * min_inv_z_point.y) + de[2], inv_z_clamp);
	const float max_point_inv_z = fmaxf((de[0] * max_inv_z_point.x + de[1] * max_inv_z_point.y) + de[2], inv_z_clamp);
	const float depth_a = de[0] * en.x + de[1] * en.y;
	const float depth_b = min_point_inv_z;
	if (fabsf(depth_a) < 1.0e-10f)
		return true;
	const float depth_b_div_a = depth_b / depth_a;
	const float max_diff_point = sqrtf((0.0f + depth_b_div_a) * (edge_len + depth_b_div_a)) - depth_b_div_a;
	const float max_diff_point_inv_z = depth_a * max_diff_point + depth_b;
	for (int i = 0; i < 2; ++i)
	{
		const float min_point_tc = (te[i] * min_inv_z_point.x + te[2 + i] * min_inv_z_point.y) + te[4 + i];
		const float max_point_tc = (te[i] * max_inv_z_point.x + te[2 + i] * max_inv_z_point.y) + te[4 + i];
		const float tc_a = te[i] * en.x + te[2 + i] * en.y;
		const float tc_b = min_point_tc;
		const float min_point_tc_z_mul = min_point_tc / min_point_inv_z;
		const float max_point_tc_z_mul = max_point_tc / max_point_inv_z;
		const float real = (tc_a * max_diff_point + tc_b) / max_diff_point_inv_z;
		const float approximate = min_point_tc_z_mul + (max_point_tc_z_mul - min_point_tc_z_mul) * (max_diff_point - 0.0f) / (edge_len - 0.0f);
		if (fabsf(real - approximate) > TC_ERROR_THRESHOLD)
			return false;
	}
	return true;
}
// renderer_utils.rs:134-197
__device__ __forceinline__ bool line_z_corrected_interpolation_may_be_used(const float* de, const float* te, const P2 max_inv_z_point, const float min_x, const float max_x)
{
	const float TC_ERROR_THRESHOLD = 0.75f;
	if (max_x - min_x < 1.0f)
		return true;
	const float a = de[0];
	const float b = de[1] * max_inv_z_point.y + de[2];
	if (fabsf(a) < 1.0e-10f)
		return true;
	const float depth_b_div_a = b / a;
	const float max_diff_x = sqrtf((min_x + depth_b_div_a) * (max_x + depth_b_div_a)) - depth_b_div_a;
	const float max_diff_point_inv_z = a * max_diff_x + b;
	const float inv_z_at_min_x = a * min_x + b;
	const float inv_z_at_max_x = a * max_x + b;
	const float almost_zero = 1e-20f;
	if (inv_z_at_min_x <= almost_zero || inv_z_at_max_x <= almost_zero || max_diff_point_inv_z <= almost_zero)
		return false;
	for (int i = 0; i < 2; ++i)
	{
		const float ta = te[i];
		const float tb = te[2 + i] * max_inv_z_point.y + te[4 + i];
		const float tc_at_min_x = (ta * min_x + tb) / inv_z_at_min_x;
		const float tc_at_max_x = (ta * max_x + tb) / inv_z_at_max_x;
		const float real = (ta * max_diff_x + tb) / max_diff_point_inv_z;
		const float approximate = tc_at_min_x + (tc_at_max_x - tc_at_min_x) * (max_diff_x - min_x) / (max_x - min_x);
		if (fabsf(real - approximate) > TC_ERROR_THRESHOLD)
			return false;
	}
	return true;
}

// One thread per polygon of the draw order: the command (mode choice, power-of-two renormalisation of the equations, :4106-4109)
// and its 16.16 vertices into per-item scratch; polygons_draw_compact_kernel packs them in order.
__global__ void __launch_bounds__(128) polygons_draw_kernel(const PolygonsDrawLaunch l)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= l.num_order)
		return;
	sqw_draw_cmd c;
	c.kind = SQW_CMD_POLYGON;
	c.blend = 0;
	c.tc_mode = 0;
	c.tex_kind = SQW_TEX_SURFACE;
	c.num_vertices = 0;
	c.vertex_offset = 0;
	c.tex_ref = 0;
	c.tex_size[0] = c.tex_size[1] = 0;
	c.depth_eq[0] = c.depth_eq[1] = c.depth_eq[2] = 0.0f;
#pragma unroll
	for (int k = 0; k < 6; ++k)
		c.tc_eq[k] = 0.0f;
	c.reserved = 0;
	const uint32_t ci = l.order[i];
	const sqw_prepared_polygon rec = l.prepared[ci];
	if (!(rec.flags & SQW_PREPARED_VISIBLE)) // DrawPolygonData::visible_frame != current frame (:2530)
	{
		l.scratch_cmds[i] = c;
		return;
	}
	const sqw_polygon_candidate cand = l.candidates[ci];
	const sqw_map_polygon& polygon = l.polygons[cand.polygon];
	P2 va[MAX_POLY_VERTICES], vb[MAX_POLY_VERTICES];
	const int count = transform_clip_project(polygon, l.vertices, l.frame.view_matrix, l.clip_planes + (size_t)cand.clip_set * 24, va, vb);
	if (count < 3)
	{
		l.scratch_cmds[i] = c;
		return;
	}
	const float* const de = rec.depth_equation;
	const float* const te = rec.tex_coord_equation;
	float min_inv_z = 1e24f, max_inv_z = -1e24f, min_x = 1e24f, max_x = -1e24f;
	P2 min_inv_z_point = va[0], max_inv_z_point = va[0];
	sqw_polygon_point* const pts = l.scratch_points + (size_t)i * MAX_POLY_VERTICES;
	for (int k = 0; k < count; ++k)
	{
		const P2 point = va[k];
		const float inv_z = (de[0] * point.x + de[1] * point.y) + de[2];
		if (inv_z < min_inv_z)
		{
			min_inv_z = inv_z;
			min_inv_z_point = point;
		}
		if (inv_z > max_inv_z)
		{
			max_inv_z = inv_z;
			max_inv_z_point = point;
		}
		if (point.x < min_x)
			min_x = point.x;
		if (point.x > max_x)
			max_x = point.x;
		pts[k] = sqw_polygon_point{f32_to_i32_x86(point.x * 65536.0f), f32_to_i32_x86(point.y * 65536.0f)}; // f32_to_fixed16_unchecked (fixed_math.rs:36)
	}
	const uint32_t material = polygon.texture;
	c.blend = (uint8_t)((material < l.num_materials ? l.material_flags[material] : 0u) >> 4);
	c.tex_ref = rec.surface_pixels_offset;
	c.tex_size[0] = (int32_t)rec.surface_size[0];
	c.tex_size[1] = (int32_t)rec.surface_size[1];
	c.num_vertices = (uint32_t)count;
	if (affine_interpolation_may_be_used(de, te, min_inv_z_point, max_inv_z_point))
	{
		c.tc_mode = SQW_TC_AFFINE;
#pragma unroll
		for (int k = 0; k < 3; ++k)
			c.depth_eq[k] = de[k];
#pragma unroll
		for (int k = 0; k < 6; ++k)
			c.tc_eq[k] = te[k];
	}
	else
	{
		// z_scale = 2^(-5 - ceil(log2(max inv_z))): an exact power of two
		const float e = ceilf((float)log2((double)fmaxf(max_inv_z, 1.0f / (float)(1 << 20))));
		const float z_scale = ldexpf(1.0f, (int)(-5.0f - e));
#pragma unroll
		for (int k = 0; k < 3; ++k)
			c.depth_eq[k] = de[k] * z_scale;
#pragma unroll
		for (int k = 0; k < 6; ++k)
			c.tc_eq[k] = te[k] * z_scale;
		c.tc_mode = line_z_corrected_interpolation_may_be_used(de, te, max_inv_z_point, min_x, max_x) ? SQW_TC_LINE_Z_CORRECTION : SQW_TC_FULL_PERSPECTIVE;
	}
	l.scratch_cmds[i] = c;
}

// Command index and first vertex of every item = number of commands / vertices before it in draw order (one CTA), then the totals.
__global__ void __launch_bounds__(1024) polygons_draw_offsets_kernel(const sqw_draw_cmd* __restrict__ scratch_cmds, const uint32_t n, uint2* __restrict__ offsets, uint32_t* __restrict__ totals)
{
	__shared__ uint32_t s_warp[2][32];
	__shared__ uint32_t s_carry[2];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 2)
		s_carry[threadIdx.x] = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t nv = i < n ? scratch_cmds[i].num_vertices : 0u;
		uint32_t incl_c = nv ? 1u : 0u, incl_v = nv;
		const uint32_t own_c = incl_c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t tc = __shfl_up_sync(0xFFFFFFFFu, incl_c, o), tv = __shfl_up_sync(0xFFFFFFFFu, incl_v, o);
			if (lane >= o)
			{
				incl_c += tc;
				incl_v += tv;
			}
		}
		if (lane == 31)
		{
			s_warp[0][warp] = incl_c;
			s_warp[1][warp] = incl_v;
		}
		__syncthreads();
		uint32_t before_c = s_carry[0], before_v = s_carry[1], all_c = 0, all_v = 0;
		for (int w = 0; w < 32; ++w)
		{
			before_c += w < warp ? s_warp[0][w] : 0u;
			before_v += w < warp ? s_warp[1][w] : 0u;
			all_c += s_warp[0][w];
			all_v += s_warp[1][w];
		}
		if (i < n)
			offsets[i] = make_uint2(before_c + incl_c - own_c, before_v + incl_v - nv);
		__syncthreads();
		if (threadIdx.x == 0)
		{
			s_carry[0] += all_c;
			s_carry[1] += all_v;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals[0] = s_carry[0];
		totals[1] = s_carry[1];
	}
}

__global__ void __launch_bounds__(128) polygons_draw_compact_kernel(
	const sqw_draw_cmd* __restrict__ scratch_cmds, const sqw_polygon_point* __restrict__ scratch_points, const uint2* __restrict__ offsets, const uint32_t n,
	sqw_draw_cmd* __restrict__ cmds, sqw_polygon_point* __restrict__ points)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	sqw_draw_cmd c = scratch_cmds[i];
	if (c.num_vertices == 0)
		return;
	const uint2 o = offsets[i];
	c.vertex_offset = o.y;
	cmds[o.x] = c;
	for (uint32_t k = 0; k < c.num_vertices; ++k)
		points[o.y + k] = scratch_points[(size_t)i * MAX_POLY_VERTICES + k];
}

} // namespace

cudaError_t launch_polygons_prepare(const PolygonsLaunch& l, cudaStream_t stream)
{
	polygons_totals_init_kernel<<<1, 1, 0, stream>>>(l.totals);
	if (l.num_candidates)
	{
		polygons_prepare_kernel<<<(l.num_candidates + 127) / 128, 128, 0, stream>>>(l);
		polygons_offsets_kernel<<<1, 1024, 0, stream>>>(l.out, l.num_candidates, l.totals);
	}
	return cudaGetLastError();
}

cudaError_t launch_polygons_draw(const PolygonsDrawLaunch& l, cudaStream_t stream)
{
	if (l.num_order == 0)
	{
		cudaMemsetAsync(l.totals, 0, 8, stream);
		return cudaGetLastError();
	}
	polygons_draw_kernel<<<(l.num_order + 127) / 128, 128, 0, stream>>>(l);
	polygons_draw_offsets_kernel<<<1, 1024, 0, stream>>>(l.scratch_cmds, l.num_order, l.offsets, l.totals);
	polygons_draw_compact_kernel<<<(l.num_order + 127) / 128, 128, 0, stream>>>(l.scratch_cmds, l.scratch_points, l.offsets, l.num_order, l.cmds, l.points);
	return cudaGetLastError();
}

} // namespace sqw
