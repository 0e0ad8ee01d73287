// Stage 3 — tonemapping, bloom and average colour on sm_100a.
//
// Replaces Postprocessor::perform_postprocessing (src/square_wheel/postprocessor.rs:68-204):
//   tonemap_kernel          perform_tonemapping (:226-256): Reinhard c / (c/255 + 1/(exposure*modulate)) per channel,
//                           fused with the colour sum that feeds update_exposure (:176-179).
//   bloom_downscale_kernel  downscale_hdr_buffer (:792-844) + colour sum of the downscaled buffer.
//   bloom_blur_{h,v}_kernel perform_blur_impl (:846-956), integer weights x 2^14, clamp-to-edge.
//   tonemap_bloom_kernel    perform_tonemapping_with_bloom_{nearest,linear} (:346-726) as a per-pixel formulation of
//                           the reference's region loops (main area, four borders, four corners).
//
// All kernels are pure streaming: 8 B in + 4 B out per pixel, 128-bit loads/stores, grid sized in multiples of the
// SM count with a grid-stride loop.  The colour sum is accumulated in exact u64 integers (the reference sums f32
// per row and then across rows; the integer sum is the value that float sum approximates).
#include "sqw_device.cuh"
#include "sqw_internal.h"

namespace sqw
{
namespace
{

__device__ __forceinline__ void block_sum3_atomic(unsigned long long b, unsigned long long g, unsigned long long r, unsigned long long* sums)
{
	__shared__ unsigned long long s_part[3][8];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
		g += __shfl_xor_sync(0xFFFFFFFFu, g, o);
		r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0)
	{
		s_part[0][warp] = b;
		s_part[1][warp] = g;
		s_part[2][warp] = r;
	}
	__syncthreads();
	if (threadIdx.x < 3)
	{
		unsigned long long t = 0;
		for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
			t += s_part[threadIdx.x][w];
		atomicAdd(sums + (blockIdx.x % SQW_SUM_SLOTS) * 4 + threadIdx.x, t);
	}
}

// Two pixels per thread per iteration: one 128-bit load, one 64-bit store.
__global__ void __launch_bounds__(256) tonemap_kernel(
	const uint64_t* __restrict__ hdr, uint32_t* __restrict__ dst, uint64_t num_pixels, float isb, float isg, float isr, unsigned long long* __restrict__ sums)
{
	unsigned long long sb = 0, sg = 0, sr = 0;
	const uint64_t num_pairs = num_pixels >> 1;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_pairs; i += stride)
	{
		const uint4 v = __ldcs(reinterpret_cast<const uint4*>(hdr) + i);
		const uint64_t p0 = (uint64_t)v.x | ((uint64_t)v.y << 32);
		const uint64_t p1 = (uint64_t)v.z | ((uint64_t)v.w << 32);
		const ColorVec c0 = cv_from_color64(p0), c1 = cv_from_color64(p1);
		sb += (p0 & 0xFFFF) + (p1 & 0xFFFF);
		sg += ((p0 >> 16) & 0xFFFF) + ((p1 >> 16) & 0xFFFF);
		sr += ((p0 >> 32) & 0xFFFF) + ((p1 >> 32) & 0xFFFF);
		__stcs(reinterpret_cast<uint2*>(dst) + i, make_uint2(tonemap_pixel(c0, isb, isg, isr), tonemap_pixel(c1, isb, isg, isr)));
	}
	if ((num_pixels & 1) && blockIdx.x == 0 && threadIdx.x == 0)
	{
		const uint64_t p0 = hdr[num_pixels - 1];
		sb += p0 & 0xFFFF;
		sg += (p0 >> 16) & 0xFFFF;
		sr += (p0 >> 32) & 0xFFFF;
		dst[num_pixels - 1] = tonemap_pixel(cv_from_color64(p0), isb, isg, isr);
	}
	block_sum3_atomic(sb, sg, sr, sums);
}

template <int S_LOG2>
__global__ void __launch_bounds__(256) bloom_downscale_kernel(
	const uint64_t* __restrict__ hdr, uint32_t hdr_w, uint64_t* __restrict__ bloom0, uint32_t bw, uint32_t bh, unsigned long long* __restrict__ sums)
{
	constexpr int S = 1 << S_LOG2;
	constexpr int32_t average_scaler = (1 << 8) / (S * S); // :799
	unsigned long long sb = 0, sg = 0, sr = 0;
	const uint32_t chunks_per_row = (bw + 255u) >> 8;
	const uint32_t num_chunks = chunks_per_row * bh;
	for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x)
	{
		const uint32_t dst_y = chunk / chunks_per_row;
		const uint32_t dst_x = ((chunk - dst_y * chunks_per_row) << 8) + threadIdx.x;
		if (dst_x >= bw)
			continue;
		const size_t i = (size_t)dst_y * bw + dst_x;
		ColorVecI sum = cvi(0, 0, 0, 0);
#pragma unroll
		for (int dy = 0; dy < S; ++dy)
		{
			const uint64_t* line = hdr + (size_t)(dst_y * S + dy) * hdr_w + (size_t)dst_x * S;
#pragma unroll
			for (int dx = 0; dx < S; ++dx) // 64-bit loads: rows of an odd-width buffer are only 8-byte aligned
				sum = cvi_add(sum, cvi_from_color64(__ldg(line + dx)));
		}
		const ColorVecI average = cvi_shr<8>(cvi_mul_scalar(sum, average_scaler));
		bloom0[i] = Color64::from_veci(average);
		sb += (uint32_t)average.b;
		sg += (uint32_t)average.g;
		sr += (uint32_t)average.r;
	}
	block_sum3_atomic(sb, sg, sr, sums);
}

struct BlurKernel
{
	int32_t w[33];
};

// HORIZONTAL = true: bloom_buffers[0] -> [1] (:863-900); false: [1] -> [0] (:910-955).
template <bool HORIZONTAL>
__global__ void __launch_bounds__(256) bloom_blur_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int32_t bw, int32_t bh, int32_t radius, BlurKernel kernel)
{
	// work item = 256 consecutive pixels of one row (32-bit index arithmetic, one division per item instead of one 64-bit
	// division per pixel)
	const uint32_t chunks_per_row = ((uint32_t)bw + 255u) >> 8;
	const uint32_t num_chunks = chunks_per_row * (uint32_t)bh;
	for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x)
	{
		const int32_t y = (int32_t)(chunk / chunks_per_row);
		const int32_t x = (int32_t)((chunk - (uint32_t)y * chunks_per_row) << 8) + (int32_t)threadIdx.x;
		if (x >= bw)
			continue;
		const uint64_t* const row = src + (size_t)y * bw;
		ColorVecI sum = cvi_mul_scalar(cvi_from_color64(__ldg(row + x)), kernel.w[radius]);
		for (int32_t d = 1; d <= radius; ++d)
		{
			uint64_t a, b;
			if (HORIZONTAL)
			{
				a = __ldg(row + max(x - d, 0));
				b = __ldg(row + min(x + d, bw - 1));
			}
			else
			{
				a = __ldg(src + (size_t)max(y - d, 0) * bw + x);
				b = __ldg(src + (size_t)min(y + d, bh - 1) * bw + x);
			}
			const ColorVecI two_sum = cvi_add(cvi_from_color64(a), cvi_from_color64(b));
			sum = cvi_add(sum, cvi_mul_scalar(two_sum, kernel.w[d + radius]));
		}
		dst[(size_t)y * bw + x] = Color64::from_veci(cvi_shr<14>(sum));
	}
}

__device__ __forceinline__ ColorVec bloom_at(const uint64_t* __restrict__ bloom0, uint32_t bw, uint32_t x, uint32_t y) { return cv_from_color64(__ldg(bloom0 + (size_t)y * bw + x)); }

// Region classification reproduces the loop nests of the reference exactly:
//   nearest (:346-462): bloom texel = (min(x / S, bw - 1), min(y / S, bh - 1)), sum = bloom + hdr.
//   linear  (:464-726): x < S/2 -> left border, x >= (bw-1)*S + S/2 -> right border, else between texels
//                       sx = (x - S/2) / S and sx + 1 with mix_factors[(x - S/2) % S]; same for y; the order of
//                       the mul_scalar_add chain is the reference's.
template <int S_LOG2, bool LINEAR>
__device__ __forceinline__ ColorVec bloom_sum(const ColorVec c, const uint32_t x, const uint32_t y, const uint64_t* __restrict__ bloom0, const uint32_t bw, const uint32_t bh)
{
	constexpr uint32_t S = 1u << S_LOG2;
	ColorVec sum;
	if (!LINEAR)
	{
		const uint32_t sx = min(x >> S_LOG2, bw - 1), sy = min(y >> S_LOG2, bh - 1);
		sum = cv_add(bloom_at(bloom0, bw, sx, sy), c);
	}
	else
	{
		const uint32_t right_border_begin_x = (bw - 1) * S + S / 2;
		const uint32_t upper_border_begin_y = (bh - 1) * S + S / 2;
		// x classification
		int xr; // 0 left border, 1 main, 2 right border
		uint32_t sx = 0;
		float xm = 0.0f;
		if (x < S / 2)
			xr = 0;
		else if (x >= right_border_begin_x)
		{
			xr = 2;
			sx = bw - 1;
		}
		else
		{
			xr = 1;
			sx = (x - S / 2) >> S_LOG2;
			const uint32_t dx = (x - S / 2) & (S - 1);
			xm = (float)((S + 1) & 1) / (2.0f * (float)S) + (float)dx / (float)S; // :477
		}
		int yr;
		uint32_t sy = 0;
		float ym = 0.0f;
		if (y < S / 2)
			yr = 0;
		else if (y >= upper_border_begin_y)
		{
			yr = 2;
			sy = bh - 1;
		}
		else
		{
			yr = 1;
			sy = (y - S / 2) >> S_LOG2;
			const uint32_t dy = (y - S / 2) & (S - 1);
			ym = (float)((S + 1) & 1) / (2.0f * (float)S) + (float)dy / (float)S;
		}

		if (yr != 1)
		{
			if (xr != 1)
				sum = cv_add(bloom_at(bloom0, bw, sx, sy), c); // corners (:672-725)
			else
			{
				// lower / upper border (:488-522, :632-669)
				const ColorVec b00 = bloom_at(bloom0, bw, sx, sy), b10 = bloom_at(bloom0, bw, sx + 1, sy);
				sum = cv_mul_scalar_add(b00, 1.0f - xm, cv_mul_scalar_add(b10, xm, c));
			}
		}
		else
		{
			const float oym = 1.0f - ym;
			if (xr != 1)
			{
				// left / right border (:536-556, :609-629)
				const ColorVec b00 = bloom_at(bloom0, bw, sx, sy), b01 = bloom_at(bloom0, bw, sx, sy + 1);
				const ColorVec bs = cv_mul_scalar_add(b00, oym, cv_scalar_mul(b01, ym));
				sum = cv_add(bs, c);
			}
			else
			{
				// main image (:558-607)
				const ColorVec b00 = bloom_at(bloom0, bw, sx, sy), b01 = bloom_at(bloom0, bw, sx, sy + 1);
				const ColorVec b10 = bloom_at(bloom0, bw, sx + 1, sy), b11 = bloom_at(bloom0, bw, sx + 1, sy + 1);
				const ColorVec left = cv_mul_scalar_add(b00, oym, cv_scalar_mul(b01, ym));
				const ColorVec right = cv_mul_scalar_add(b10, oym, cv_scalar_mul(b11, ym));
				sum = cv_mul_scalar_add(left, 1.0f - xm, cv_mul_scalar_add(right, xm, c));
			}
		}
	}
	return sum;
}

// Work item = one run of consecutive pixels of one row (32-bit index arithmetic, the row classification uniform per item).
// Rows whose width is a multiple of four (every usual resolution) take FOUR pixels per thread: two 128-bit loads of the HDR
// buffer in flight per thread and one 128-bit store of the picture; neighbouring pixels share their bloom texels through L1.
template <int S_LOG2, bool LINEAR>
__global__ void __launch_bounds__(256) tonemap_bloom_kernel(
	const uint64_t* __restrict__ hdr, uint32_t* __restrict__ dst, uint32_t w, uint32_t h, const uint64_t* __restrict__ bloom0, uint32_t bw, uint32_t bh, float isb,
	float isg, float isr)
{
	if ((w & 3u) == 0)
	{
		const uint32_t quads_per_row = w >> 2;
		const uint32_t chunks_per_row = (quads_per_row + 255u) >> 8;
		const uint32_t num_chunks = chunks_per_row * h;
		for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x)
		{
			const uint32_t y = chunk / chunks_per_row;
			const uint32_t q = ((chunk - y * chunks_per_row) << 8) + threadIdx.x;
			if (q >= quads_per_row)
				continue;
			const uint32_t x = q << 2;
			const size_t i = (size_t)y * w + x; // a multiple of four: 32-byte aligned in hdr, 16-byte aligned in dst
			const uint4 v0 = __ldcs(reinterpret_cast<const uint4*>(hdr + i)), v1 = __ldcs(reinterpret_cast<const uint4*>(hdr + i) + 1);
			const uint64_t p[4] = {(uint64_t)v0.x | ((uint64_t)v0.y << 32), (uint64_t)v0.z | ((uint64_t)v0.w << 32), (uint64_t)v1.x | ((uint64_t)v1.y << 32), (uint64_t)v1.z | ((uint64_t)v1.w << 32)};
			uint32_t out[4];
#pragma unroll
			for (int k = 0; k < 4; ++k)
				out[k] = tonemap_pixel(bloom_sum<S_LOG2, LINEAR>(cv_from_color64(p[k]), x + k, y, bloom0, bw, bh), isb, isg, isr);
			__stcs(reinterpret_cast<uint4*>(dst + i), make_uint4(out[0], out[1], out[2], out[3]));
		}
		return;
	}
	const uint32_t chunks_per_row = (w + 255u) >> 8;
	const uint32_t num_chunks = chunks_per_row * h;
	for (uint32_t chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x)
	{
		const uint32_t y = chunk / chunks_per_row;
		const uint32_t x = ((chunk - y * chunks_per_row) << 8) + threadIdx.x;
		if (x >= w)
			continue;
		const size_t i = (size_t)y * w + x;
		__stcs(dst + i, tonemap_pixel(bloom_sum<S_LOG2, LINEAR>(cv_from_color64(__ldcs(hdr + i)), x, y, bloom0, bw, bh), isb, isg, isr));
	}
}

inline uint32_t grid_for(uint64_t work_items, int threads, int max_ctas_per_sm = 8)
{
	// multiples of the SM count (148 on B200)
	const uint64_t needed = (work_items + threads - 1) / threads;
	const uint64_t cap = 148ull * max_ctas_per_sm;
	return (uint32_t)(needed < cap ? (needed ? needed : 1) : cap);
}


// ---- compressed read-back of the 32-bit picture (host half: readback_decode.cu) --------------------------------------------------
// One CTA per picture row.  A run of equal pixels becomes one word B | G << 8 | R << 16 | (length - 1) << 24; runs never cross a
// 256-pixel boundary, so the length always fits the byte the constant alpha (255) leaves free.  The row's runs go to
// `rows_out + row * row_capacity` (page-locked host memory mapped into the device: only what is written crosses PCIe) and their
// number to row_counts[row].  A row with any alpha != 255 is copied raw (count = 0x80000000 | width): lossless in every case.
#define SQW_RLE_THREADS 128
#define SQW_RLE_SEGMENT (SQW_RLE_THREADS * 8)
__global__ void __launch_bounds__(SQW_RLE_THREADS) picture_rle_kernel(
	const uint32_t* __restrict__ ldr, uint32_t width, uint32_t pitch, uint32_t* __restrict__ rows_out, uint32_t row_capacity, uint32_t* __restrict__ row_counts)
{
	__shared__ uint16_t s_start[SQW_RLE_SEGMENT + 1];
	__shared__ uint32_t s_val[SQW_RLE_SEGMENT];
	__shared__ uint32_t s_warp[SQW_RLE_THREADS / 32];
	const uint32_t row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t* const src = ldr + (size_t)row * pitch;
	uint32_t* const dst = rows_out + (size_t)row * row_capacity;
	// alpha check of the whole row first (rare path: raw copy)
	{
		uint32_t bad = 0;
		for (uint32_t x = tid; x < width; x += SQW_RLE_THREADS)
			bad |= (~__ldg(src + x)) & 0xFF000000u;
		if (__syncthreads_or(bad != 0))
		{
			for (uint32_t x = tid; x < width; x += SQW_RLE_THREADS)
				dst[x] = __ldg(src + x);
			if (tid == 0)
				row_counts[row] = 0x80000000u | width;
			return;
		}
	}
	uint32_t total = 0;
	for (uint32_t seg0 = 0; seg0 < width; seg0 += SQW_RLE_SEGMENT)
	{
		const uint32_t x0 = seg0 + tid * 8;
		uint32_t p[8];
		uint32_t prev = 0;
		if (x0 < width && x0 > 0)
			prev = __ldg(src + x0 - 1);
		uint32_t flags = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			const uint32_t x = x0 + i;
			p[i] = x < width ? __ldg(src + x) : 0u;
			const bool start = x < width && (((x & 255u) == 0) || p[i] != (i ? p[i - 1] : prev));
			flags |= start ? (1u << i) : 0u;
		}
		const uint32_t cnt = __popc(flags);
		uint32_t incl = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
			if ((int)lane >= o)
				incl += t;
		}
		if (lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		uint32_t base = incl - cnt, seg_total = 0;
#pragma unroll
		for (int w = 0; w < SQW_RLE_THREADS / 32; ++w)
		{
			base += w < (int)warp ? s_warp[w] : 0u;
			seg_total += s_warp[w];
		}
		__syncthreads(); // every warp has read the warp totals before the run records are written (and before the next segment rewrites the totals)
#pragma unroll
		for (int i = 0; i < 8; ++i)
			if (flags & (1u << i))
			{
#ifdef SQW_DEBUG_RLE_BOUNDS
				if (base >= SQW_RLE_SEGMENT)
					__trap();
#endif
				s_start[base] = (uint16_t)(tid * 8 + i);
				s_val[base] = p[i];
				++base;
			}
		if (tid == 0)
			s_start[seg_total] = (uint16_t)min((uint32_t)SQW_RLE_SEGMENT, width - seg0);
		__syncthreads();
		for (uint32_t r = tid; r < seg_total; r += SQW_RLE_THREADS)
			dst[total + r] = (s_val[r] & 0x00FFFFFFu) | (((uint32_t)s_start[r + 1] - (uint32_t)s_start[r] - 1u) << 24);
		total += seg_total;
		__syncthreads(); // s_start / s_val / s_warp are rewritten by the next segment
	}
	if (tid == 0)
		row_counts[row] = total;
}

} // namespace

cudaError_t launch_tonemap(const PostLaunch& l, cudaStream_t stream)
{
	const uint64_t n = (uint64_t)l.width * l.height;
	tonemap_kernel<<<grid_for(n / 2 + 1, 256), 256, 0, stream>>>(l.hdr, l.dst, n, l.inv_scale_bgr[0], l.inv_scale_bgr[1], l.inv_scale_bgr[2], l.color_sums);
	return cudaGetLastError();
}

cudaError_t launch_bloom_downscale(const PostLaunch& l, cudaStream_t stream)
{
	const uint64_t n = (uint64_t)l.bloom_w * l.bloom_h;
	if (n == 0)
		return cudaSuccess;
	const uint32_t g = grid_for(n, 256);
	switch (l.bloom_scale_log2)
	{
	case 1: bloom_downscale_kernel<1><<<g, 256, 0, stream>>>(l.hdr, l.width, l.bloom0, l.bloom_w, l.bloom_h, l.color_sums); break;
	case 2: bloom_downscale_kernel<2><<<g, 256, 0, stream>>>(l.hdr, l.width, l.bloom0, l.bloom_w, l.bloom_h, l.color_sums); break;
	default: bloom_downscale_kernel<3><<<g, 256, 0, stream>>>(l.hdr, l.width, l.bloom0, l.bloom_w, l.bloom_h, l.color_sums); break;
	}
	return cudaGetLastError();
}

cudaError_t launch_bloom_blur(const PostLaunch& l, cudaStream_t stream)
{
	const uint64_t n = (uint64_t)l.bloom_w * l.bloom_h;
	if (n == 0 || l.blur_radius <= 0)
		return cudaSuccess;
	BlurKernel kh, kv;
	for (int i = 0; i < 33; ++i)
	{
		kh.w[i] = l.kernel_h[i];
		kv.w[i] = l.kernel_v[i];
	}
	const uint32_t g = grid_for(n, 256);
	bloom_blur_kernel<true><<<g, 256, 0, stream>>>(l.bloom0, l.bloom1, (int32_t)l.bloom_w, (int32_t)l.bloom_h, l.blur_radius, kh);
	bloom_blur_kernel<false><<<g, 256, 0, stream>>>(l.bloom1, l.bloom0, (int32_t)l.bloom_w, (int32_t)l.bloom_h, l.blur_radius, kv);
	return cudaGetLastError();
}

cudaError_t launch_tonemap_bloom(const PostLaunch& l, cudaStream_t stream)
{
	const uint64_t n = (uint64_t)l.width * l.height;
	const uint32_t g = grid_for(n, 256);
	const float b = l.inv_scale_bgr[0], gg = l.inv_scale_bgr[1], r = l.inv_scale_bgr[2];
#define SQW_TB(SL, LIN) tonemap_bloom_kernel<SL, LIN><<<g, 256, 0, stream>>>(l.hdr, l.dst, l.width, l.height, l.bloom0, l.bloom_w, l.bloom_h, b, gg, r)
	if (l.linear)
	{
		switch (l.bloom_scale_log2)
		{
		case 1: SQW_TB(1, true); break;
		case 2: SQW_TB(2, true); break;
		default: SQW_TB(3, true); break;
		}
	}
	else
	{
		switch (l.bloom_scale_log2)
		{
		case 1: SQW_TB(1, false); break;
		case 2: SQW_TB(2, false); break;
		default: SQW_TB(3, false); break;
		}
	}
#undef SQW_TB
	return cudaGetLastError();
}

cudaError_t launch_picture_rle(const uint32_t* ldr, uint32_t width, uint32_t height, uint32_t pitch, uint32_t* rows_out, uint32_t row_capacity, uint32_t* row_counts, cudaStream_t stream)
{
	picture_rle_kernel<<<height, SQW_RLE_THREADS, 0, stream>>>(ldr, width, pitch, rows_out, row_capacity, row_counts);
	return cudaGetLastError();
}

} // namespace sqw
