// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).  PARITY UNPINNED.
//
// Restatement of src/square_wheel/postprocessor.rs (stage 3: tonemapping + bloom + average colour).
#include "oracle_common.h"
#include "oracle_pool.h"
#include <vector>

namespace oracle
{
namespace
{

constexpr size_t MAX_GAUSSIAN_KERNEL_RADIUS = 16;                          // postprocessor.rs:991
constexpr size_t MAX_GAUSSIAN_KERNEL_SIZE = 1 + 2 * MAX_GAUSSIAN_KERNEL_RADIUS; // :992

struct TonemappingFunction // :997-1024
{
	ColorVec inv_scale_vec, inv_255_vec;
	TonemappingFunction(float exposure, const float color_modulate[3])
	{
		const float inv_255 = 1.0f / 255.0f;
		const float s[3] = {1.0f / (exposure * color_modulate[0]), 1.0f / (exposure * color_modulate[1]), 1.0f / (exposure * color_modulate[2])};
		inv_scale_vec = cv_from_f32x3(s);
		const float i[3] = {inv_255, inv_255, inv_255};
		inv_255_vec = cv_from_f32x3(i);
	}
	ColorVec do_it(ColorVec c) const { return cv_div(c, cv_mul_add(c, inv_255_vec, inv_scale_vec)); } // :1019
};

void compute_gaussian_kernel(float sigma, size_t radius, float result[MAX_GAUSSIAN_KERNEL_SIZE]) // :1026
{
	for (size_t i = 0; i < MAX_GAUSSIAN_KERNEL_SIZE; ++i)
		result[i] = 0.0f;
	for (int32_t x = -(int32_t)radius; x <= (int32_t)radius; ++x)
	{
		const float SAMPLES[4] = {-0.375f, -0.125f, 0.125f, 0.375f};
		float val = 0.0f;
		for (float sample : SAMPLES)
		{
			const float coord = (float)x + sample;
			val += std::exp(-0.5f * (coord / sigma) * (coord / sigma)) / (sigma * std::sqrt(6.28318530717958647692f));
		}
		result[(size_t)(x + (int32_t)radius)] = val / 4.0f;
	}
}

struct Post
{
	const uint64_t* hdr_buffer;
	size_t hdr_size[2];
	uint32_t* pixels;
	size_t pitch;
	std::vector<uint64_t> bloom_buffers[2];
	size_t bloom_size[2];

	// :226-256
	ColorVec perform_tonemapping(size_t y_start, size_t y_end, const TonemappingFunction& f) const
	{
#ifdef ORACLE_NUMERICS_SSE
		// The release build's ColorVec IS an __m128 (fast_math.rs:95-98): the loop below is postprocessor.rs:226-256 with the same
		// intrinsics per pixel — from_color64 (:123-132), add, mul_add = vfmadd (:208), div (:213), into_color32 = min + cvtps2dq +
		// pshufb (:134-146).  Lane for lane the same IEEE operations as the scalar statement of this mode further down.
		{
			const __m128 inv_scale = _mm_loadu_ps(f.inv_scale_vec.v), inv_255 = _mm_loadu_ps(f.inv_255_vec.v);
			const __m128 max_val = _mm_set1_ps(255.0f);
			const __m128i shuffle_mask = _mm_set_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12, 8, 4, 0);
			__m128 colors_sum = _mm_setzero_ps();
			for (size_t y = y_start; y < y_end; ++y)
			{
				__m128 line_colors_sum = _mm_setzero_ps();
				const uint64_t* src_line = hdr_buffer + y * hdr_size[0];
				uint32_t* dst_line = pixels + y * pitch;
				for (size_t x = 0; x < hdr_size[0]; ++x)
				{
					const __m128 c = _mm_cvtepi32_ps(_mm_cvtepu16_epi32(_mm_cvtsi64_si128((long long)src_line[x])));
					line_colors_sum = _mm_add_ps(line_colors_sum, c);
					const __m128 mapped = _mm_div_ps(c, _mm_fmadd_ps(c, inv_255, inv_scale));
					const __m128i values_32bit = _mm_cvtps_epi32(_mm_min_ps(mapped, max_val));
					dst_line[x] = (uint32_t)_mm_cvtsi128_si32(_mm_shuffle_epi8(values_32bit, shuffle_mask));
				}
				colors_sum = _mm_add_ps(colors_sum, line_colors_sum);
			}
			ColorVec r;
			_mm_storeu_ps(r.v, colors_sum);
			return r;
		}
#endif
		ColorVec colors_sum = cv_zero();
		for (size_t y = y_start; y < y_end; ++y)
		{
			ColorVec line_colors_sum = cv_zero();
			const uint64_t* src_line = hdr_buffer + y * hdr_size[0];
			uint32_t* dst_line = pixels + y * pitch;
			for (size_t x = 0; x < hdr_size[0]; ++x)
			{
				const ColorVec c = cv_from_color64(src_line[x]);
				line_colors_sum = cv_add(line_colors_sum, c);
				dst_line[x] = cv_into_color32(f.do_it(c));
			}
			colors_sum = cv_add(colors_sum, line_colors_sum);
		}
		return colors_sum;
	}

	// :792-844
	ColorVec downscale_hdr_buffer(size_t S)
	{
		const int32_t average_scaler = (1 << 8) / (int32_t)(S * S);
		ColorVec colors_sum = cv_zero();
		for (size_t dst_y = 0; dst_y < bloom_size[1]; ++dst_y)
		{
			const size_t src_y_base = dst_y * S;
			ColorVecI line_colors_sum = {{0, 0, 0, 0}};
			for (size_t dst_x = 0; dst_x < bloom_size[0]; ++dst_x)
			{
				const size_t src_x_base = dst_x * S;
				ColorVecI sum = {{0, 0, 0, 0}};
				for (size_t dy = 0; dy < S; ++dy)
					for (size_t dx = 0; dx < S; ++dx)
						sum = cvi_add(sum, cvi_from_color64(hdr_buffer[src_x_base + dx + (src_y_base + dy) * hdr_size[0]]));
				const ColorVecI average = cvi_shr<8>(cvi_mul_scalar(sum, average_scaler));
				bloom_buffers[0][dst_x + dst_y * bloom_size[0]] = cvi_into_color64(average);
				line_colors_sum = cvi_add(line_colors_sum, average);
			}
			const ColorVec lf = {{(float)line_colors_sum.v[0], (float)line_colors_sum.v[1], (float)line_colors_sum.v[2], (float)line_colors_sum.v[3]}};
			colors_sum = cv_add(colors_sum, lf);
		}
		return cv_scalar_mul(colors_sum, 1.0f / (float)(bloom_size[0] * bloom_size[1]));
	}

	// :846-956
	void perform_blur(size_t RADIUS, const float blur_kernel[MAX_GAUSSIAN_KERNEL_SIZE], float bloom_scale)
	{
		const int COLOR_SHIFT = 14;
		int32_t blur_kernel_i[MAX_GAUSSIAN_KERNEL_SIZE];
		for (size_t i = 0; i < MAX_GAUSSIAN_KERNEL_SIZE; ++i)
			blur_kernel_i[i] = (int32_t)(blur_kernel[i] * (float)(1 << COLOR_SHIFT));
		const int32_t radius_i = (int32_t)RADIUS;
		const int32_t size_minus_one[2] = {(int32_t)bloom_size[0] - 1, (int32_t)bloom_size[1] - 1};

		for (size_t dst_y = 0; dst_y < bloom_size[1]; ++dst_y)
		{
			const size_t line_offset = dst_y * bloom_size[0];
			for (size_t dst_x = 0; dst_x < bloom_size[0]; ++dst_x)
			{
				ColorVecI sum = cvi_mul_scalar(cvi_from_color64(bloom_buffers[0][dst_x + line_offset]), blur_kernel_i[radius_i]);
				for (int32_t dx = 1; dx <= radius_i; ++dx)
				{
					const int32_t src_x_0 = std::max((int32_t)dst_x - dx, 0);
					const int32_t src_x_1 = std::min((int32_t)dst_x + dx, size_minus_one[0]);
					const ColorVecI two_sum =
						cvi_add(cvi_from_color64(bloom_buffers[0][(size_t)src_x_0 + line_offset]), cvi_from_color64(bloom_buffers[0][(size_t)src_x_1 + line_offset]));
					sum = cvi_add(sum, cvi_mul_scalar(two_sum, blur_kernel_i[dx + radius_i]));
				}
				bloom_buffers[1][dst_x + line_offset] = cvi_into_color64(cvi_shr<COLOR_SHIFT>(sum));
			}
		}

		for (size_t i = 0; i < MAX_GAUSSIAN_KERNEL_SIZE; ++i)
			blur_kernel_i[i] = (int32_t)(blur_kernel[i] * bloom_scale * (float)(1 << COLOR_SHIFT)); // :906

		for (size_t dst_y = 0; dst_y < bloom_size[1]; ++dst_y)
		{
			for (size_t dst_x = 0; dst_x < bloom_size[0]; ++dst_x)
			{
				ColorVecI sum = cvi_mul_scalar(cvi_from_color64(bloom_buffers[1][dst_x + dst_y * bloom_size[0]]), blur_kernel_i[radius_i]);
				for (int32_t dy = 1; dy <= radius_i; ++dy)
				{
					const int32_t src_y_0 = std::max((int32_t)dst_y - dy, 0);
					const int32_t src_y_1 = std::min((int32_t)dst_y + dy, size_minus_one[1]);
					const ColorVecI two_sum = cvi_add(
						cvi_from_color64(bloom_buffers[1][dst_x + (size_t)src_y_0 * bloom_size[0]]),
						cvi_from_color64(bloom_buffers[1][dst_x + (size_t)src_y_1 * bloom_size[0]]));
					sum = cvi_add(sum, cvi_mul_scalar(two_sum, blur_kernel_i[dy + radius_i]));
				}
				bloom_buffers[0][dst_x + dst_y * bloom_size[0]] = cvi_into_color64(cvi_shr<COLOR_SHIFT>(sum));
			}
		}
	}

	// :729-789.  Returns bloom buffer scale; average colour through out parameter.
	size_t perform_bloom(const sqw_post_params& p, ColorVec& average_color)
	{
		const size_t scale_log2 = std::min<size_t>(std::max<size_t>(p.bloom_buffer_scale_log2, 1), 3);
		const size_t bloom_buffer_scale = (size_t)1 << scale_log2;
		bloom_size[0] = hdr_size[0] / bloom_buffer_scale;
		bloom_size[1] = hdr_size[1] / bloom_buffer_scale;
		for (auto& b : bloom_buffers)
			b.assign(bloom_size[0] * bloom_size[1], 0);

		average_color = downscale_hdr_buffer(bloom_buffer_scale);

		const float bloom_sigma_corrected = p.bloom_sigma / (float)bloom_buffer_scale;
		const size_t blur_radius = std::min<size_t>((size_t)fmax_(std::ceil(3.0f * bloom_sigma_corrected - 0.5f), 0.0f), MAX_GAUSSIAN_KERNEL_RADIUS);
		float blur_kernel[MAX_GAUSSIAN_KERNEL_SIZE];
		compute_gaussian_kernel(bloom_sigma_corrected, blur_radius, blur_kernel);
		if (blur_radius > 0)
			perform_blur(blur_radius, blur_kernel, p.bloom_scale);
		return bloom_buffer_scale;
	}

	inline void put(size_t x, size_t y, ColorVec sum, const TonemappingFunction& f) const { pixels[x + y * pitch] = cv_into_color32(f.do_it(sum)); }
	inline ColorVec hdr(size_t x, size_t y) const { return cv_from_color64(hdr_buffer[x + y * hdr_size[0]]); }
	inline ColorVec bloom(size_t x, size_t y) const { return cv_from_color64(bloom_buffers[0][x + y * bloom_size[0]]); }

	// :346-462
	void tonemap_bloom_nearest(size_t S, size_t src_y_start, size_t src_y_end, const TonemappingFunction& f) const
	{
		const size_t columns_left = hdr_size[0] - bloom_size[0] * S;
		const size_t lines_left = hdr_size[1] - bloom_size[1] * S;
		for (size_t src_y = src_y_start; src_y < src_y_end; ++src_y)
		{
			const size_t dst_y_base = src_y * S;
			for (size_t src_x = 0; src_x < bloom_size[0]; ++src_x)
			{
				const ColorVec bloom_c = bloom(src_x, src_y);
				for (size_t dy = 0; dy < S; ++dy)
					for (size_t dx = 0; dx < S; ++dx)
						put(src_x * S + dx, dst_y_base + dy, cv_add(bloom_c, hdr(src_x * S + dx, dst_y_base + dy)), f);
			}
			if (columns_left > 0)
			{
				const size_t dst_x_base = bloom_size[0] * S;
				const ColorVec bloom_c = bloom(bloom_size[0] - 1, src_y);
				for (size_t dy = 0; dy < S; ++dy)
					for (size_t dx = 0; dx < columns_left; ++dx)
						put(dst_x_base + dx, dst_y_base + dy, cv_add(bloom_c, hdr(dst_x_base + dx, dst_y_base + dy)), f);
			}
		}
		if (lines_left > 0 && src_y_end >= bloom_size[1])
		{
			const size_t src_y = bloom_size[1] - 1;
			const size_t dst_y_base = bloom_size[1] * S;
			for (size_t src_x = 0; src_x < bloom_size[0]; ++src_x)
			{
				const ColorVec bloom_c = bloom(src_x, src_y);
				for (size_t dy = 0; dy < lines_left; ++dy)
					for (size_t dx = 0; dx < S; ++dx)
						put(src_x * S + dx, dst_y_base + dy, cv_add(bloom_c, hdr(src_x * S + dx, dst_y_base + dy)), f);
			}
			if (columns_left > 0)
			{
				const size_t dst_x_base = bloom_size[0] * S;
				const ColorVec bloom_c = bloom(bloom_size[0] - 1, src_y);
				for (size_t dy = 0; dy < lines_left; ++dy)
					for (size_t dx = 0; dx < columns_left; ++dx)
						put(dst_x_base + dx, dst_y_base + dy, cv_add(bloom_c, hdr(dst_x_base + dx, dst_y_base + dy)), f);
			}
		}
	}

	// :464-726
	void tonemap_bloom_linear(size_t S, size_t src_y_start, size_t src_y_end, const TonemappingFunction& f) const
	{
		float mix_factors[8];
		const float bloom_buffer_scale_f = (float)S;
		for (size_t i = 0; i < S; ++i)
			mix_factors[i] = (float)((S + 1) & 1) / (2.0f * bloom_buffer_scale_f) + (float)i / bloom_buffer_scale_f;

		const size_t right_border_begin_x = (bloom_size[0] - 1) * S + S / 2;
		const size_t right_border_size = hdr_size[0] - right_border_begin_x;
		const size_t upper_border_begin_y = (bloom_size[1] - 1) * S + S / 2;
		const size_t upper_border_size = hdr_size[1] - upper_border_begin_y;

		auto horizontal_band = [&](size_t src_row, size_t dst_y_base, size_t num_rows) { // lower (:488-522) and upper (:632-669) borders
			ColorVec b00 = bloom(0, src_row);
			for (size_t src_x = 0; src_x + 1 < bloom_size[0]; ++src_x)
			{
				const size_t dst_x_base = src_x * S + S / 2;
				const ColorVec b10 = bloom(src_x + 1, src_row);
				for (size_t dy = 0; dy < num_rows; ++dy)
					for (size_t dx = 0; dx < S; ++dx)
					{
						const float xm = mix_factors[dx];
						const float oxm = 1.0f - xm;
						const ColorVec sum = cv_mul_scalar_add(b00, oxm, cv_mul_scalar_add(b10, xm, hdr(dst_x_base + dx, dst_y_base + dy)));
						put(dst_x_base + dx, dst_y_base + dy, sum, f);
					}
				b00 = b10;
			}
		};

		if (src_y_start == 0)
			horizontal_band(0, 0, S / 2);

		for (size_t src_y = src_y_start; src_y < std::min(src_y_end, bloom_size[1] - 1); ++src_y)
		{
			const size_t dst_y_base = src_y * S + S / 2;
			ColorVec b00 = bloom(0, src_y);
			ColorVec b01 = bloom(0, src_y + 1);
			for (size_t dy = 0; dy < S; ++dy) // left border :536-556
			{
				const float ym = mix_factors[dy];
				const float oym = 1.0f - ym;
				const ColorVec bs = cv_mul_scalar_add(b00, oym, cv_scalar_mul(b01, ym));
				for (size_t dx = 0; dx < S / 2; ++dx)
					put(dx, dst_y_base + dy, cv_add(bs, hdr(dx, dst_y_base + dy)), f);
			}
			for (size_t src_x = 0; src_x + 1 < bloom_size[0]; ++src_x) // main :558-607
			{
				const size_t dst_x_base = src_x * S + S / 2;
				const ColorVec b10 = bloom(src_x + 1, src_y);
				const ColorVec b11 = bloom(src_x + 1, src_y + 1);
				for (size_t dy = 0; dy < S; ++dy)
				{
					const float ym = mix_factors[dy];
					const float oym = 1.0f - ym;
					const ColorVec left = cv_mul_scalar_add(b00, oym, cv_scalar_mul(b01, ym));
					const ColorVec right = cv_mul_scalar_add(b10, oym, cv_scalar_mul(b11, ym));
					for (size_t dx = 0; dx < S; ++dx)
					{
						const float xm = mix_factors[dx];
						const float oxm = 1.0f - xm;
						const ColorVec sum = cv_mul_scalar_add(left, oxm, cv_mul_scalar_add(right, xm, hdr(dst_x_base + dx, dst_y_base + dy)));
						put(dst_x_base + dx, dst_y_base + dy, sum, f);
					}
				}
				b00 = b10;
				b01 = b11;
			}
			for (size_t dy = 0; dy < S; ++dy) // right border :609-629
			{
				const float ym = mix_factors[dy];
				const float oym = 1.0f - ym;
				const ColorVec bs = cv_mul_scalar_add(b00, oym, cv_scalar_mul(b01, ym));
				for (size_t dx = 0; dx < right_border_size; ++dx)
					put(right_border_begin_x + dx, dst_y_base + dy, cv_add(bs, hdr(right_border_begin_x + dx, dst_y_base + dy)), f);
			}
		}

		if (src_y_end >= bloom_size[1])
			horizontal_band(bloom_size[1] - 1, upper_border_begin_y, upper_border_size);

		auto process_corner = [&](size_t x0, size_t x1, size_t y0, size_t y1, ColorVec bloom_value) { // :672-685
			for (size_t y = y0; y < y1; ++y)
				for (size_t x = x0; x < x1; ++x)
					put(x, y, cv_add(bloom_value, hdr(x, y)), f);
		};
		if (src_y_start == 0)
		{
			process_corner(0, S / 2, 0, S / 2, bloom(0, 0));
			process_corner(right_border_begin_x, hdr_size[0], 0, S / 2, bloom(bloom_size[0] - 1, 0));
		}
		if (src_y_end >= bloom_size[1])
		{
			process_corner(0, S / 2, upper_border_begin_y, hdr_size[1], bloom(0, bloom_size[1] - 1));
			process_corner(right_border_begin_x, hdr_size[0], upper_border_begin_y, hdr_size[1], bloom(bloom_size[0] - 1, bloom_size[1] - 1));
		}
	}
};

} // namespace
} // namespace oracle

// Postprocessor::perform_postprocessing (postprocessor.rs:68-204) without the exposure update.
// avg_color_out = ColorVec lanes [B, G, R].  bloom_out (optional) receives bloom_buffers[0] after the blur.
extern "C" int oracle_postprocess(
	const uint64_t* hdr_buffer, uint32_t width, uint32_t height, const sqw_post_params* params, uint32_t* dst, size_t dst_pitch, float avg_color_out[3],
	uint64_t* bloom_out, int num_threads)
{
	using namespace oracle;
	Post p;
	p.hdr_buffer = hdr_buffer;
	p.hdr_size[0] = width;
	p.hdr_size[1] = height;
	p.pixels = dst;
	p.pitch = dst_pitch;
	p.bloom_size[0] = p.bloom_size[1] = 0;

	const bool use_bloom = params->bloom_sigma > 0.0f;
	size_t bloom_buffer_scale = 1;
	ColorVec average_color = cv_zero();
	if (use_bloom)
		bloom_buffer_scale = p.perform_bloom(*params, average_color);

	const TonemappingFunction f(params->exposure, params->color_modulate);
	const int nt = std::max(1, std::min(num_threads, 64));

	if (use_bloom)
	{
		auto run = [&](size_t y0, size_t y1) {
			if (!params->linear_bloom_filter)
				p.tonemap_bloom_nearest(bloom_buffer_scale, y0, y1, f);
			else
				p.tonemap_bloom_linear(bloom_buffer_scale, y0, y1, f);
		};
		if (nt == 1)
			run(0, p.bloom_size[1]);
		else
		{
			// :127-145
			ThreadPool::instance().parallel_for((uint32_t)nt, nt, [&](uint32_t i) { run(p.bloom_size[1] * (size_t)i / (size_t)nt, p.bloom_size[1] * (size_t)(i + 1) / (size_t)nt); });
		}
		if (bloom_out)
			std::memcpy(bloom_out, p.bloom_buffers[0].data(), p.bloom_buffers[0].size() * sizeof(uint64_t));
	}
	else
	{
		ColorVec colors_sum = cv_zero();
		if (nt == 1)
			colors_sum = p.perform_tonemapping(0, height, f);
		else
		{
			std::vector<ColorVec> sums((size_t)nt, cv_zero()); // :156-174
			ThreadPool::instance().parallel_for((uint32_t)nt, nt, [&](uint32_t i) { sums[(size_t)i] = p.perform_tonemapping((size_t)height * (size_t)i / (size_t)nt, (size_t)height * (size_t)(i + 1) / (size_t)nt, f); });
			for (auto& s : sums)
				colors_sum = cv_add(colors_sum, s);
		}
		average_color = cv_scalar_mul(colors_sum, 1.0f / (float)((size_t)width * (size_t)height));
	}
	avg_color_out[0] = average_color.v[0];
	avg_color_out[1] = average_color.v[1];
	avg_color_out[2] = average_color.v[2];
	return 0;
}

// Postprocessor::update_exposure (postprocessor.rs:206-223); config values as in postprocessor_config.rs.
extern "C" float oracle_update_exposure(
	float current_exposure, const float average_color_bgr[3], float frame_duration_s, float zero_level_brightness, float brightness_scale_power,
	float base_brightness, float min_exposure, float max_exposure, float exposure_update_speed)
{
	using namespace oracle;
	const float brightness = fmin_(fmax_((average_color_bgr[0] + average_color_bgr[1] + average_color_bgr[2]) * 0.33f, 1.0f / 1024.0f), 65536.0f) / 255.0f;
	const float target_exposure = std::pow(brightness / zero_level_brightness, brightness_scale_power) * 0.5f * base_brightness / brightness;
	const float clamped = fmin_(fmax_(target_exposure, min_exposure), max_exposure);
	const float mix_factor = std::exp(-exposure_update_speed * frame_duration_s);
	return std::exp(std::log(current_exposure) * mix_factor + std::log(clamped) * (1.0f - mix_factor));
}
