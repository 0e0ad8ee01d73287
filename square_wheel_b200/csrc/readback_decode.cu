// Host half of the compressed read-back of the 32-bit picture (sqw_read_pixels32_async with SQW_CTX_COMPRESSED_READBACK).
//
// The reference hands its picture to the window as `pixels: &mut [Color32]` in host memory (host.rs:476-492); through the C-ABI
// that is a device -> host copy of W x H x 4 bytes, and at 4K it is THE end-to-end limit: 33 MB per frame over PCIe.  The alpha
// byte of the tonemapped picture is 255 by construction (postprocessor.rs:997-1024; every producer of the picture here writes
// it), and at 4K a surface texel covers several pixels, so a row is mostly short runs of one colour.  picture_rle_kernel
// (post_kernels.cu) therefore writes one 32-bit word per run — B | G << 8 | R << 16 | (length - 1) << 24, runs cut at 256-pixel
// boundaries — straight into page-locked, device-mapped host memory (only the bytes written cross the bus, no size has to be
// known in advance), and this file expands the rows into the caller's buffer on a small pool of host threads.  Lossless: a row
// whose alpha is not 255 everywhere travels raw (count word with bit 31 set).
//
// Compiled by nvcc as host-only code with -mavx2 (the GPU boxes are x86-64-v3; the scalar path below is the same algorithm).
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace sqw
{

// One row of runs -> `width` pixels at `out` (which has room for 16 more pixels when `slack` is set: a scratch row).
static inline void decode_row(const uint32_t* src, uint32_t c, uint32_t width, uint32_t* out, bool slack)
{
	uint32_t* const end = out + width;
	uint32_t i = 0;
#if defined(__AVX2__)
	// Fast part: every run is two unconditional 8-pixel stores (the runs that follow overwrite the excess) and only runs longer
	// than 16 pixels loop: run lengths are data, so a branch per run would be mispredicted all the time.  The position of the
	// next run depends on the previous one only through an add.  Without slack behind the row the last runs take the exact loop.
	uint32_t* const fast_end = slack ? end : end - 16;
	for (; i < c; ++i)
	{
		const uint32_t w = src[i];
		const uint32_t len = (w >> 24) + 1;
		if (out + len > fast_end)
			break;
		const __m256i v = _mm256_set1_epi32((int)(w | 0xFF000000u));
		_mm256_storeu_si256(reinterpret_cast<__m256i*>(out), v);
		_mm256_storeu_si256(reinterpret_cast<__m256i*>(out + 8), v);
		if (__builtin_expect(len > 16, 0))
			for (uint32_t k = 16; k < len; k += 8)
				_mm256_storeu_si256(reinterpret_cast<__m256i*>(out + k), v);
		out += len;
	}
#else
	(void)slack;
#endif
	for (; i < c; ++i) // exact stores
	{
		const uint32_t w = src[i];
		const uint32_t len = (w >> 24) + 1;
		const uint32_t px = w | 0xFF000000u;
		for (uint32_t k = 0; k < len && out + k < end; ++k)
			out[k] = px;
		out += len;
	}
}

// rows [row0, row1) of the picture: `rle` holds row r at r * row_capacity words, `counts[r]` = number of run words (bit 31: raw row).
// Rows are expanded in a cache-resident scratch row and leave for the caller's buffer with non-temporal stores: the picture is
// written once and not read back by this thread, so the lines need not be fetched first (half the memory traffic of plain stores).
void decode_picture_rows(const uint32_t* rle, const uint32_t* counts, uint32_t row_capacity, uint32_t width, uint32_t row0, uint32_t row1, uint32_t* dst, size_t dst_pitch)
{
	constexpr uint32_t SCRATCH_PIXELS = 8192;
	alignas(64) uint32_t scratch[SCRATCH_PIXELS + 32];
	for (uint32_t r = row0; r < row1; ++r)
	{
		const uint32_t* src = rle + (size_t)r * row_capacity;
		uint32_t* out = dst + (size_t)r * dst_pitch;
		const uint32_t c = counts[r];
		if (c & 0x80000000u)
		{
			std::memcpy(out, src, (size_t)width * 4);
			continue;
		}
#if defined(__AVX2__)
		if (width <= SCRATCH_PIXELS && width >= 64)
		{
			decode_row(src, c, width, scratch, true);
			// head up to the first 32-byte boundary of the destination, aligned body with streaming stores, tail
			uint32_t x = 0;
			const uint32_t head = (uint32_t)(((32u - ((uintptr_t)out & 31u)) & 31u) / 4u);
			for (; x < head; ++x)
				out[x] = scratch[x];
			for (; x + 8 <= width; x += 8)
				_mm256_stream_si256(reinterpret_cast<__m256i*>(out + x), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(scratch + x)));
			for (; x < width; ++x)
				out[x] = scratch[x];
			continue;
		}
#endif
		decode_row(src, c, width, out, false);
	}
#if defined(__AVX2__)
	_mm_sfence(); // the streaming stores are visible before the job is reported done
#endif
}

} // namespace sqw
