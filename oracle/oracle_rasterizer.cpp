// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_common.h).  PARITY UNPINNED.
//
// Restatement of src/square_wheel/rasterizer.rs (stage 2: span rasteriser) — Rasterizer::fill_polygon with its
// three texture-coordinate interpolation modes and five blending modes, Rasterizer::fill_triangle, and the
// thread/rect split of PartialRenderer::perform_rasterization (partial_renderer.rs:1192-1283, rect_splitting.rs).
// Compiled with -fwrapv: Rust release builds wrap on signed overflow and so does this file.
#include "oracle_common.h"
#include "oracle_pool.h"
#include <vector>

namespace oracle
{
namespace
{

struct ClipRect // rasterizer.rs:1476
{
	int32_t min_x, min_y, max_x, max_y;
};
struct PolygonSide // :1505
{
	Fixed16 x_start, dx_dy;
};
struct DepthEquation // equations.rs:5
{
	float d_inv_z_dx, d_inv_z_dy, k;
};
struct TexCoordEquation // equations.rs:14
{
	float d_tc_dx[2], d_tc_dy[2], k[2];
};
struct TextureInfo // rasterizer.rs:1500
{
	int32_t size[2];
};

constexpr int32_t TC_SHIFT = 24;        // :1573
constexpr int32_t INV_Z_SHIFT = 29;     // :1574
constexpr int32_t INV_Z_PRE_SHIFT = 8;  // :1575
constexpr int32_t Z_CALC_SHIFT = 31;    // :1576
constexpr int64_t TC_FINAL_SHIFT = TC_SHIFT + Z_CALC_SHIFT - INV_Z_SHIFT + INV_Z_PRE_SHIFT; // :1577 = 34
constexpr float TC_SCALE = (float)(1 << TC_SHIFT);       // :1579
constexpr float INV_Z_SCALE = (float)(1 << INV_Z_SHIFT); // :1580

inline uint32_t unchecked_div(uint32_t x, uint32_t y) { return x / y; } // :1547

template <class C, int BLENDING_MODE>
inline void write_into_framebuffer(typename C::Raw& dst_pixel, typename C::Raw texel) // :1511
{
	if (BLENDING_MODE == BLEND_NONE)
		dst_pixel = texel;
	else if (BLENDING_MODE == BLEND_AVERAGE)
		dst_pixel = C::average(dst_pixel, texel);
	else if (BLENDING_MODE == BLEND_ADDITIVE)
		dst_pixel = C::saturated_sum(dst_pixel, texel);
	else if (BLENDING_MODE == BLEND_ALPHA_TEST)
	{
		if (C::test_alpha(texel))
			dst_pixel = texel;
	}
	else if (BLENDING_MODE == BLEND_ALPHA_BLEND)
		dst_pixel = alpha_blend<C>(dst_pixel, texel);
}

template <class C>
struct Rasterizer
{
	typedef typename C::Raw Raw;
	Raw* color_buffer;
	int32_t row_size;
	ClipRect clip_rect;

	struct Eq64
	{
		int64_t d_inv_z_dx, d_inv_z_dy, inv_z_k;
		int64_t d_tc_dx[2], d_tc_dy[2], tc_k[2];
	};
	static Eq64 make_eq(const DepthEquation& de, const TexCoordEquation& te) // :287-310 / :472-495
	{
		Eq64 e;
		e.d_inv_z_dx = f32_to_i64_trunc(INV_Z_SCALE * de.d_inv_z_dx);
		e.d_inv_z_dy = f32_to_i64_trunc(INV_Z_SCALE * de.d_inv_z_dy);
		e.inv_z_k = f32_to_i64_trunc(INV_Z_SCALE * (de.k + (de.d_inv_z_dx + de.d_inv_z_dy) * 0.5f));
		for (int i = 0; i < 2; ++i)
		{
			e.d_tc_dx[i] = f32_to_i64_trunc(TC_SCALE * te.d_tc_dx[i]);
			e.d_tc_dy[i] = f32_to_i64_trunc(TC_SCALE * te.d_tc_dy[i]);
			e.tc_k[i] = f32_to_i64_trunc(TC_SCALE * (te.k[i] + (te.d_tc_dx[i] + te.d_tc_dy[i]) * 0.5f));
		}
		return e;
	}

	// :273-454
	template <int BLENDING_MODE>
	void fill_polygon_part(
		Fixed16 y_start, Fixed16 y_end, PolygonSide left_side, PolygonSide right_side, const DepthEquation& depth_equation,
		const TexCoordEquation& tex_coord_equation, const TextureInfo& texture_info, const Raw* texture_data)
	{
		const Eq64 e = make_eq(depth_equation, tex_coord_equation);
		const int32_t texture_size_minus_one[2] = {texture_info.size[0] - 1, texture_info.size[1] - 1};
		const uint32_t texture_width = (uint32_t)texture_info.size[0];

		const int32_t y_start_int = std::max(fixed16_round_to_int(y_start), clip_rect.min_y);
		const int32_t y_end_int = std::min(fixed16_round_to_int(y_end), clip_rect.max_y);
		const Fixed16 y_start_delta = int_to_fixed16(y_start_int) + FIXED16_HALF - y_start;
		Fixed16 x_left = left_side.x_start + fixed16_mul(y_start_delta, left_side.dx_dy) + FIXED16_HALF;
		Fixed16 x_right = right_side.x_start + fixed16_mul(y_start_delta, right_side.dx_dy) + FIXED16_HALF;
		int64_t line_inv_z = (int64_t)y_start_int * e.d_inv_z_dy + e.inv_z_k;
		int64_t line_tc[2] = {(int64_t)y_start_int * e.d_tc_dy[0] + e.tc_k[0], (int64_t)y_start_int * e.d_tc_dy[1] + e.tc_k[1]};

		for (int32_t y_int = y_start_int; y_int < y_end_int; ++y_int)
		{
			const int32_t x_start_int = std::max(fixed16_floor_to_int(x_left), clip_rect.min_x);
			const int32_t x_end_int = std::min(fixed16_floor_to_int(x_right), clip_rect.max_x);
			if (x_start_int < x_end_int)
			{
				Raw* line_dst = color_buffer + (size_t)(x_start_int + y_int * row_size);
				const int64_t span_start_x = x_start_int;
				const int64_t span_end_x = x_end_int - 1;
				const int32_t span_length_minus_one = (x_end_int - 1) - x_start_int;

				int32_t span_inv_z, span_d_inv_z;
				{
					const int64_t inv_z_min = 1 << INV_Z_PRE_SHIFT;
					const int64_t inv_z_max = 1 << 29;
					const int32_t span_inv_z_start = (int32_t)std::max(inv_z_min, std::min(span_start_x * e.d_inv_z_dx + line_inv_z, inv_z_max));
					const int32_t span_inv_z_end = (int32_t)std::max(inv_z_min, std::min(span_end_x * e.d_inv_z_dx + line_inv_z, inv_z_max));
					span_d_inv_z = span_length_minus_one > 0 ? (span_inv_z_end - span_inv_z_start) / span_length_minus_one : 0;
					span_inv_z = span_inv_z_start;
				}

				int32_t span_tc[2] = {0, 0};
				int32_t span_d_tc[2] = {0, 0};
				for (int i = 0; i < 2; ++i)
				{
					const int32_t tc_max_shift = INV_Z_SHIFT - TC_SHIFT;
					const int64_t tc_max_max = 1 << 29;
					const int64_t span_inv_z_corrected = (int64_t)(span_inv_z - (1 << INV_Z_PRE_SHIFT));
					const int64_t tc_max_start =
						std::max<int64_t>(0, std::min<int64_t>(((span_inv_z_corrected * (int64_t)texture_info.size[i]) >> tc_max_shift) - 1, tc_max_max));
					const int32_t span_tc_start = (int32_t)std::max<int64_t>(0, std::min<int64_t>(span_start_x * e.d_tc_dx[i] + line_tc[i], tc_max_start));

					if (span_length_minus_one > 0)
					{
						const int64_t tc_max_end = std::max<int64_t>(
							0, std::min<int64_t>(
								   (((span_inv_z_corrected + (int64_t)(int32_t)(span_d_inv_z * span_length_minus_one)) * (int64_t)texture_info.size[i]) >> tc_max_shift) - 1,
								   tc_max_max));
						const int32_t span_tc_end = (int32_t)std::max<int64_t>(0, std::min<int64_t>(span_end_x * e.d_tc_dx[i] + line_tc[i], tc_max_end));

						span_d_tc[i] = (span_tc_end - span_tc_start) / span_length_minus_one;
						if (span_tc_end > span_tc_start)
							span_tc[i] = span_tc_start;
						else
							span_tc[i] = span_tc_end - span_d_tc[i] * span_length_minus_one;
						ORACLE_ASSERT(span_tc[i] >= 0);                                                  // :413
						ORACLE_ASSERT(span_tc[i] <= span_tc_start);                                     // :414
						ORACLE_ASSERT(span_tc[i] + span_d_tc[i] * span_length_minus_one >= 0);          // :415
						ORACLE_ASSERT(span_tc[i] + span_d_tc[i] * span_length_minus_one <= span_tc_end); // :416
					}
					else
						span_tc[i] = span_tc_start;
				}

				const int32_t n = x_end_int - x_start_int;
				for (int32_t p = 0; p < n; ++p)
				{
					const uint32_t z = unchecked_div(1u << Z_CALC_SHIFT, (uint32_t)span_inv_z >> INV_Z_PRE_SHIFT);
					const uint32_t pix_tc[2] = {
						(uint32_t)(((uint64_t)z * (uint64_t)(int64_t)span_tc[0]) >> TC_FINAL_SHIFT),
						(uint32_t)(((uint64_t)z * (uint64_t)(int64_t)span_tc[1]) >> TC_FINAL_SHIFT)};
					ORACLE_ASSERT(pix_tc[0] <= (uint32_t)texture_size_minus_one[0]); // :436
					ORACLE_ASSERT(pix_tc[1] <= (uint32_t)texture_size_minus_one[1]); // :437
					const size_t texel_address = (size_t)(pix_tc[0] + pix_tc[1] * texture_width);
					write_into_framebuffer<C, BLENDING_MODE>(line_dst[p], texture_data[texel_address]);
					span_inv_z += span_d_inv_z;
					span_tc[0] += span_d_tc[0];
					span_tc[1] += span_d_tc[1];
				}
			}
			x_left += left_side.dx_dy;
			x_right += right_side.dx_dy;
			line_inv_z += e.d_inv_z_dy;
			line_tc[0] += e.d_tc_dy[0];
			line_tc[1] += e.d_tc_dy[1];
		}
	}

	// :456-581
	template <int BLENDING_MODE>
	void fill_polygon_part_line_z_corrected(
		Fixed16 y_start, Fixed16 y_end, PolygonSide left_side, PolygonSide right_side, const DepthEquation& depth_equation,
		const TexCoordEquation& tex_coord_equation, const TextureInfo& texture_info, const Raw* texture_data)
	{
		const int64_t LINE_TC_SHIFT = 16;
		const Eq64 e = make_eq(depth_equation, tex_coord_equation);
		const uint32_t texture_width = (uint32_t)texture_info.size[0];

		const int32_t y_start_int = std::max(fixed16_round_to_int(y_start), clip_rect.min_y);
		const int32_t y_end_int = std::min(fixed16_round_to_int(y_end), clip_rect.max_y);
		const Fixed16 y_start_delta = int_to_fixed16(y_start_int) + FIXED16_HALF - y_start;
		Fixed16 x_left = left_side.x_start + fixed16_mul(y_start_delta, left_side.dx_dy) + FIXED16_HALF;
		Fixed16 x_right = right_side.x_start + fixed16_mul(y_start_delta, right_side.dx_dy) + FIXED16_HALF;
		int64_t line_inv_z = (int64_t)y_start_int * e.d_inv_z_dy + e.inv_z_k;
		int64_t line_tc[2] = {(int64_t)y_start_int * e.d_tc_dy[0] + e.tc_k[0], (int64_t)y_start_int * e.d_tc_dy[1] + e.tc_k[1]};

		for (int32_t y_int = y_start_int; y_int < y_end_int; ++y_int)
		{
			const int32_t x_start_int = std::max(fixed16_floor_to_int(x_left), clip_rect.min_x);
			const int32_t x_end_int = std::min(fixed16_floor_to_int(x_right), clip_rect.max_x);
			if (x_start_int < x_end_int)
			{
				Raw* line_dst = color_buffer + (size_t)(x_start_int + y_int * row_size);
				const int64_t span_start_x = x_start_int;
				const int64_t span_end_x = x_end_int - 1;
				const int32_t span_length_minus_one = (x_end_int - 1) - x_start_int;

				const int64_t inv_z_min = 1 << INV_Z_PRE_SHIFT;
				const int64_t inv_z_max = 1 << 29;
				const uint32_t span_inv_z_start = (uint32_t)std::min(std::max(span_start_x * e.d_inv_z_dx + line_inv_z, inv_z_min), inv_z_max);
				const uint32_t span_inv_z_end = (uint32_t)std::min(std::max(span_end_x * e.d_inv_z_dx + line_inv_z, inv_z_min), inv_z_max);
				const uint64_t span_z_start = unchecked_div(1u << Z_CALC_SHIFT, span_inv_z_start >> INV_Z_PRE_SHIFT);
				const uint64_t span_z_end = unchecked_div(1u << Z_CALC_SHIFT, span_inv_z_end >> INV_Z_PRE_SHIFT);

				int32_t span_tc[2] = {0, 0};
				int32_t span_d_tc[2] = {0, 0};
				for (int i = 0; i < 2; ++i)
				{
					const uint64_t tc_start = (uint64_t)std::max<int64_t>(line_tc[i] + span_start_x * e.d_tc_dx[i], 0);
					const uint64_t tc_end = (uint64_t)std::max<int64_t>(line_tc[i] + span_end_x * e.d_tc_dx[i], 0);
					const uint64_t max_tc = ((uint64_t)(int64_t)texture_info.size[i] << LINE_TC_SHIFT) - 1;
					const int32_t tc_mul_z_start = (int32_t)std::min((span_z_start * tc_start) >> (TC_FINAL_SHIFT - LINE_TC_SHIFT), max_tc);
					const int32_t tc_mul_z_end = (int32_t)std::min((span_z_end * tc_end) >> (TC_FINAL_SHIFT - LINE_TC_SHIFT), max_tc);
					span_tc[i] = tc_mul_z_start;
					if (span_length_minus_one > 0)
						span_d_tc[i] = (tc_mul_z_end - tc_mul_z_start) / span_length_minus_one;
				}

				const int32_t n = x_end_int - x_start_int;
				for (int32_t p = 0; p < n; ++p)
				{
					const int32_t tc_int[2] = {span_tc[0] >> LINE_TC_SHIFT, span_tc[1] >> LINE_TC_SHIFT};
					ORACLE_ASSERT(tc_int[0] >= 0);
					ORACLE_ASSERT(tc_int[1] >= 0);
					ORACLE_ASSERT(tc_int[0] < texture_info.size[0]);
					ORACLE_ASSERT(tc_int[1] < texture_info.size[1]);
					const size_t texel_address = (size_t)((uint32_t)tc_int[0] + (uint32_t)tc_int[1] * texture_width);
					write_into_framebuffer<C, BLENDING_MODE>(line_dst[p], texture_data[texel_address]);
					span_tc[0] += span_d_tc[0];
					span_tc[1] += span_d_tc[1];
				}
			}
			x_left += left_side.dx_dy;
			x_right += right_side.dx_dy;
			line_inv_z += e.d_inv_z_dy;
			line_tc[0] += e.d_tc_dy[0];
			line_tc[1] += e.d_tc_dy[1];
		}
	}

	// :583-729
	template <int BLENDING_MODE>
	void fill_polygon_part_affine(
		Fixed16 y_start, Fixed16 y_end, PolygonSide left_side, PolygonSide right_side, const DepthEquation& de, const TexCoordEquation& te,
		const TextureInfo& texture_info, const Raw* texture_data)
	{
		const float y_start_f32 = fixed16_to_f32(y_start);
		const float y_end_f32 = fixed16_to_f32(y_end);
		const Fixed16 y_delta = y_end - y_start;
		const float x_start_left = fixed16_to_f32(left_side.x_start);
		const float x_end_left = fixed16_to_f32(left_side.x_start + fixed16_mul(y_delta, left_side.dx_dy));
		const float x_start_right = fixed16_to_f32(right_side.x_start);
		const float x_end_right = fixed16_to_f32(right_side.x_start + fixed16_mul(y_delta, right_side.dx_dy));
		const float z_start_left = 1.0f / (de.d_inv_z_dx * x_start_left + de.d_inv_z_dy * y_start_f32 + de.k);
		const float z_end_left = 1.0f / (de.d_inv_z_dx * x_end_left + de.d_inv_z_dy * y_end_f32 + de.k);
		const float z_start_right = 1.0f / (de.d_inv_z_dx * x_start_right + de.d_inv_z_dy * y_start_f32 + de.k);
		const float z_end_right = 1.0f / (de.d_inv_z_dx * x_end_right + de.d_inv_z_dy * y_end_f32 + de.k);

		const Fixed16 y_delta_for_tc_interpoltion = std::max(y_delta, FIXED16_ONE);

		Fixed16 tc_left[2], tc_right[2], d_tc_left[2], d_tc_right[2];
		for (int i = 0; i < 2; ++i)
		{
			const Fixed16 max_tc = int_to_fixed16(texture_info.size[i]) - 1;
			auto corner = [&](float z, float x, float y) {
				return std::min(std::max(f32_to_fixed16_unchecked(z * (x * te.d_tc_dx[i] + y * te.d_tc_dy[i] + te.k[i])), 0), max_tc);
			};
			const Fixed16 tc_start_left = corner(z_start_left, x_start_left, y_start_f32);
			const Fixed16 tc_end_left = corner(z_end_left, x_end_left, y_end_f32);
			const Fixed16 tc_start_right = corner(z_start_right, x_start_right, y_start_f32);
			const Fixed16 tc_end_right = corner(z_end_right, x_end_right, y_end_f32);
			d_tc_left[i] = fixed16_div(tc_end_left - tc_start_left, y_delta_for_tc_interpoltion);
			d_tc_right[i] = fixed16_div(tc_end_right - tc_start_right, y_delta_for_tc_interpoltion);
			tc_left[i] = tc_start_left;
			tc_right[i] = tc_start_right;
		}

		const int32_t y_start_int = std::max(fixed16_round_to_int(y_start), clip_rect.min_y);
		const int32_t y_end_int = std::min(fixed16_round_to_int(y_end), clip_rect.max_y);
		const Fixed16 y_start_delta = int_to_fixed16(y_start_int) + FIXED16_HALF - y_start;
		Fixed16 x_left = left_side.x_start + fixed16_mul(y_start_delta, left_side.dx_dy) + FIXED16_HALF;
		Fixed16 x_right = right_side.x_start + fixed16_mul(y_start_delta, right_side.dx_dy) + FIXED16_HALF;
		for (int i = 0; i < 2; ++i)
		{
			tc_left[i] += fixed16_mul(y_start_delta, d_tc_left[i]);
			tc_right[i] += fixed16_mul(y_start_delta, d_tc_right[i]);
		}

		for (int32_t y_int = y_start_int; y_int < y_end_int; ++y_int)
		{
			const int32_t x_start_int = std::max(fixed16_floor_to_int(x_left), clip_rect.min_x);
			const int32_t x_end_int = std::min(fixed16_floor_to_int(x_right), clip_rect.max_x);
			if (x_start_int < x_end_int)
			{
				Raw* line_dst = color_buffer + (size_t)(x_start_int + y_int * row_size);
				Fixed16 tc[2] = {tc_left[0], tc_left[1]};
				const Fixed16 x_delta_for_tc_interpolation = std::max(x_right - x_left, FIXED16_ONE);
				const Fixed16 d_tc[2] = {
					fixed16_div(tc_right[0] - tc_left[0], x_delta_for_tc_interpolation), fixed16_div(tc_right[1] - tc_left[1], x_delta_for_tc_interpolation)};
				const int32_t n = x_end_int - x_start_int;
				for (int32_t p = 0; p < n; ++p)
				{
					const int32_t tc_int[2] = {fixed16_floor_to_int(tc[0]), fixed16_floor_to_int(tc[1])};
					ORACLE_ASSERT(tc_int[0] >= 0);
					ORACLE_ASSERT(tc_int[1] >= 0);
					ORACLE_ASSERT(tc_int[0] < texture_info.size[0]);
					ORACLE_ASSERT(tc_int[1] < texture_info.size[1]);
					const size_t texel_address = (size_t)(tc_int[0] + tc_int[1] * texture_info.size[0]);
					write_into_framebuffer<C, BLENDING_MODE>(line_dst[p], texture_data[texel_address]);
					tc[0] += d_tc[0];
					tc[1] += d_tc[1];
				}
			}
			x_left += left_side.dx_dy;
			x_right += right_side.dx_dy;
			tc_left[0] += d_tc_left[0];
			tc_left[1] += d_tc_left[1];
			tc_right[0] += d_tc_right[0];
			tc_right[1] += d_tc_right[1];
		}
	}

	// :137-271 fill_polygon_impl_2_static_params
	template <int MODE, int BLENDING_MODE>
	void fill_polygon_impl(
		const sqw_polygon_point* vertices, size_t num_vertices, const DepthEquation& de, const TexCoordEquation& te, const TextureInfo& ti,
		const Raw* texture_data)
	{
		auto draw_func = [&](Fixed16 y0, Fixed16 y1, PolygonSide l, PolygonSide r) {
			if (MODE == SQW_TC_FULL_PERSPECTIVE)
				fill_polygon_part<BLENDING_MODE>(y0, y1, l, r, de, te, ti, texture_data);
			else if (MODE == SQW_TC_LINE_Z_CORRECTION)
				fill_polygon_part_line_z_corrected<BLENDING_MODE>(y0, y1, l, r, de, te, ti, texture_data);
			else
				fill_polygon_part_affine<BLENDING_MODE>(y0, y1, l, r, de, te, ti, texture_data);
		};

		size_t lower_vertex_index = 0;
		Fixed16 min_y = vertices[0].y;
		for (size_t index = 0; index < num_vertices; ++index)
			if (vertices[index].y < min_y)
			{
				min_y = vertices[index].y;
				lower_vertex_index = index;
			}

		size_t left_index = lower_vertex_index;
		size_t right_index = lower_vertex_index;
		Fixed16 cur_y = min_y;
		for (;;)
		{
			size_t next_left_index = left_index + num_vertices - 1;
			if (next_left_index >= num_vertices)
				next_left_index -= num_vertices;
			size_t next_right_index = right_index + 1;
			if (next_right_index >= num_vertices)
				next_right_index -= num_vertices;

			const sqw_polygon_point& left_vertex = vertices[left_index];
			const sqw_polygon_point& next_left_vertex = vertices[next_left_index];
			const sqw_polygon_point& right_vertex = vertices[right_index];
			const sqw_polygon_point& next_right_vertex = vertices[next_right_index];

			const Fixed16 dy_left = next_left_vertex.y - left_vertex.y;
			const Fixed16 dy_right = next_right_vertex.y - right_vertex.y;
			const Fixed16 next_y = std::min(next_left_vertex.y, next_right_vertex.y);
			if (dy_left > FIXED16_HALF && dy_right > FIXED16_HALF)
			{
				const Fixed16 dx_dy_left = fixed16_div(next_left_vertex.x - left_vertex.x, dy_left);
				const Fixed16 dx_dy_right = fixed16_div(next_right_vertex.x - right_vertex.x, dy_right);
				draw_func(
					cur_y, next_y, PolygonSide{left_vertex.x + fixed16_mul(dx_dy_left, cur_y - left_vertex.y), dx_dy_left},
					PolygonSide{right_vertex.x + fixed16_mul(dx_dy_right, cur_y - right_vertex.y), dx_dy_right});
			}
			else if (dy_left > 0 && dy_right > 0)
			{
				const int32_t cur_y_int = fixed16_round_to_int(cur_y);
				const int32_t next_y_int = fixed16_round_to_int(next_y);
				if (cur_y_int < next_y_int)
				{
					const Fixed16 thin_line_y = int_to_fixed16(cur_y_int) + FIXED16_HALF;
					const Fixed16 x_start_left = left_vertex.x + fixed16_mul_div(thin_line_y - left_vertex.y, next_left_vertex.x - left_vertex.x, dy_left);
					const Fixed16 x_start_right =
						right_vertex.x + fixed16_mul_div(thin_line_y - right_vertex.y, next_right_vertex.x - right_vertex.x, dy_right);
					draw_func(cur_y, next_y, PolygonSide{x_start_left, 0}, PolygonSide{x_start_right, 0});
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
	}

	template <int MODE>
	void fill_polygon_blend(
		int blend, const sqw_polygon_point* v, size_t n, const DepthEquation& de, const TexCoordEquation& te, const TextureInfo& ti, const Raw* tex)
	{
		switch (blend) // :92-134
		{
		case BLEND_NONE: fill_polygon_impl<MODE, BLEND_NONE>(v, n, de, te, ti, tex); break;
		case BLEND_AVERAGE: fill_polygon_impl<MODE, BLEND_AVERAGE>(v, n, de, te, ti, tex); break;
		case BLEND_ADDITIVE: fill_polygon_impl<MODE, BLEND_ADDITIVE>(v, n, de, te, ti, tex); break;
		case BLEND_ALPHA_TEST: fill_polygon_impl<MODE, BLEND_ALPHA_TEST>(v, n, de, te, ti, tex); break;
		case BLEND_ALPHA_BLEND: fill_polygon_impl<MODE, BLEND_ALPHA_BLEND>(v, n, de, te, ti, tex); break;
		}
	}

	// :35-80
	void fill_polygon(
		const sqw_polygon_point* v, size_t n, const DepthEquation& de, const TexCoordEquation& te, const TextureInfo& ti, const Raw* tex, int mode,
		int blend)
	{
		switch (mode)
		{
		case SQW_TC_FULL_PERSPECTIVE: fill_polygon_blend<SQW_TC_FULL_PERSPECTIVE>(blend, v, n, de, te, ti, tex); break;
		case SQW_TC_LINE_Z_CORRECTION: fill_polygon_blend<SQW_TC_LINE_Z_CORRECTION>(blend, v, n, de, te, ti, tex); break;
		default: fill_polygon_blend<SQW_TC_AFFINE>(blend, v, n, de, te, ti, tex); break;
		}
	}

	// :984-1292
	template <class TC, int BLENDING_MODE>
	void fill_triangle_part(
		Fixed16 y_start, Fixed16 y_end, PolygonSide left_side, PolygonSide right_side, const Fixed16 tc_start_left[2], const Fixed16 d_tc_dy_left_in[2],
		const Fixed16 d_tc_dx_in[2], const Fixed16 light_start_left[3], const Fixed16 d_light_left_in[3], const Fixed16 d_light_dx_in[3],
		const TextureInfo& texture_info, const typename TC::Raw* texture_data)
	{
		Fixed16 d_tc_dy_left[2] = {d_tc_dy_left_in[0], d_tc_dy_left_in[1]};
		Fixed16 d_tc_dx[2] = {d_tc_dx_in[0], d_tc_dx_in[1]};
		Fixed16 d_light_left[3] = {d_light_left_in[0], d_light_left_in[1], d_light_left_in[2]};
		Fixed16 d_light_dx[3] = {d_light_dx_in[0], d_light_dx_in[1], d_light_dx_in[2]};

		const int32_t y_start_int = std::max(fixed16_round_to_int(y_start), clip_rect.min_y);
		const int32_t y_end_int = std::min(fixed16_round_to_int(y_end), clip_rect.max_y);
		if (y_start_int >= y_end_int)
			return;

		const Fixed16 y_start_delta = int_to_fixed16(y_start_int) + FIXED16_HALF - y_start;
		Fixed16 x_left = left_side.x_start + fixed16_mul(y_start_delta, left_side.dx_dy) + FIXED16_HALF;
		Fixed16 x_right = right_side.x_start + fixed16_mul(y_start_delta, right_side.dx_dy) + FIXED16_HALF;

		const Fixed16 y_end_delta = y_end - (y_start + y_start_delta);
		const Fixed16 d_x_lower = std::max(x_right - x_left, 0);
		const Fixed16 x_upper_left = x_left + fixed16_mul(y_end_delta, left_side.dx_dy);
		const Fixed16 x_upper_right = x_right + fixed16_mul(y_end_delta, right_side.dx_dy);
		const Fixed16 d_x_upper = std::max(x_upper_right - x_upper_left, 0);

		Fixed16 tc_left[2] = {0, 0};
		for (int i = 0; i < 2; ++i)
		{
			const Fixed16 max_tc = int_to_fixed16(texture_info.size[i]) - 1;
			{
				tc_left[i] = std::min(std::max(tc_start_left[i] + fixed16_mul(y_start_delta, d_tc_dy_left[i]), 0), max_tc);
				const Fixed16 tc_lower_delta = fixed16_mul(d_tc_dx[i], d_x_lower);
				if (tc_lower_delta > max_tc)
				{
					tc_left[i] = 0;
					d_tc_dx[i] = fixed16_div(max_tc, d_x_lower);
				}
				else if (tc_lower_delta < -max_tc)
				{
					tc_left[i] = max_tc;
					d_tc_dx[i] = fixed16_div(-max_tc, d_x_lower);
				}
				else
				{
					const Fixed16 tc_lower_right = tc_left[i] + tc_lower_delta;
					if (tc_lower_right > max_tc)
						tc_left[i] -= tc_lower_right - max_tc;
					else if (tc_lower_right < 0)
						tc_left[i] -= tc_lower_right;
				}
			}
			ORACLE_ASSERT(tc_left[i] >= 0);
			ORACLE_ASSERT(tc_left[i] <= max_tc);
			{
				Fixed16 tc_upper_left = std::min(std::max(tc_left[i] + fixed16_mul(y_end_delta, d_tc_dy_left[i]), 0), max_tc);
				const Fixed16 tc_upper_delta = fixed16_mul(d_tc_dx[i], d_x_upper);
				if (tc_upper_delta > max_tc)
				{
					tc_upper_left = 0;
					d_tc_dx[i] = fixed16_div(max_tc, d_x_upper);
				}
				else if (tc_upper_delta < -max_tc)
				{
					tc_upper_left = max_tc;
					d_tc_dx[i] = fixed16_div(-max_tc, d_x_upper);
				}
				else
				{
					const Fixed16 tc_upper_right = tc_upper_left + tc_upper_delta;
					if (tc_upper_right > max_tc)
						tc_upper_left -= tc_upper_right - max_tc;
					else if (tc_upper_right < 0)
						tc_upper_left -= tc_upper_right;
				}
				ORACLE_ASSERT(tc_upper_left >= 0);
				ORACLE_ASSERT(tc_upper_left <= max_tc);
				if (y_end_delta > 0)
					d_tc_dy_left[i] = fixed16_div(tc_upper_left - tc_left[i], y_end_delta);
				else
					d_tc_dy_left[i] = 0;

				const Fixed16 tc_upper_left2 = tc_left[i] + fixed16_mul(y_end_delta, d_tc_dy_left[i]);
				const Fixed16 tc_upper_delta2 = fixed16_mul(d_tc_dx[i], d_x_upper);
				const Fixed16 tc_upper_right2 = tc_upper_left2 + tc_upper_delta2;
				if (tc_upper_right2 < 0)
					d_tc_dx[i] = d_x_upper > 0 ? fixed16_div(-tc_upper_left2, d_x_upper) : 0;
				else if (tc_upper_right2 > max_tc)
					d_tc_dx[i] = d_x_upper > 0 ? fixed16_div(max_tc - tc_upper_left2, d_x_upper) : 0;
			}
		}

		Fixed16 light_left[3] = {0, 0, 0};
		for (int i = 0; i < 3; ++i)
		{
			light_left[i] = std::max(light_start_left[i] + fixed16_mul(y_start_delta, d_light_left[i]), 0);
			const Fixed16 light_lower_right = light_left[i] + fixed16_mul(d_x_lower, d_light_dx[i]);
			if (light_lower_right < 0)
				light_left[i] -= light_lower_right;
			ORACLE_ASSERT(light_left[i] >= 0);
			{
				Fixed16 light_upper_left = std::max(light_left[i] + fixed16_mul(y_end_delta, d_light_left[i]), 0);
				const Fixed16 light_upper_right = light_upper_left + fixed16_mul(d_x_upper, d_light_dx[i]);
				if (light_upper_right < 0)
					light_upper_left -= light_upper_right;
				ORACLE_ASSERT(light_upper_left >= 0);
				if (y_end_delta > 0)
					d_light_left[i] = fixed16_div(light_upper_left - light_left[i], y_end_delta);
				else
					d_light_left[i] = 0;

				const Fixed16 light_upper_left2 = light_left[i] + fixed16_mul(y_end_delta, d_light_left[i]);
				const Fixed16 light_upper_right2 = light_upper_left2 + fixed16_mul(d_x_upper, d_light_dx[i]);
				if (light_upper_right2 < 0)
					d_light_dx[i] = d_x_upper > 0 ? fixed16_div(-light_upper_left2, d_x_upper) : 0;
			}
		}

		const ColorVecI d_light_dx_vec = cvi_from_i32x3(d_light_dx);

		for (int32_t y_int = y_start_int; y_int < y_end_int; ++y_int)
		{
			const int32_t x_start_int = std::max(fixed16_floor_to_int(x_left), clip_rect.min_x);
			const int32_t x_end_int = std::min(fixed16_floor_to_int(x_right), clip_rect.max_x);
			if (x_start_int < x_end_int)
			{
				const Fixed16 x_start_delta = int_to_fixed16(x_start_int) + FIXED16_ONE - x_left;
				ORACLE_ASSERT(x_start_delta >= 0);
				Fixed16 line_tc[2] = {tc_left[0] + fixed16_mul(x_start_delta, d_tc_dx[0]), tc_left[1] + fixed16_mul(x_start_delta, d_tc_dx[1])};
				const Fixed16 ll[3] = {
					light_left[0] + fixed16_mul(x_start_delta, d_light_dx[0]), light_left[1] + fixed16_mul(x_start_delta, d_light_dx[1]),
					light_left[2] + fixed16_mul(x_start_delta, d_light_dx[2])};
				ColorVecI line_light = cvi_from_i32x3(ll);

				Raw* line_dst = color_buffer + (size_t)(x_start_int + y_int * row_size);
				const int32_t n = x_end_int - x_start_int;
				for (int32_t p = 0; p < n; ++p)
				{
					Raw& dst_pixel = line_dst[p];
					const int32_t u = fixed16_floor_to_int(line_tc[0]);
					const int32_t v = fixed16_floor_to_int(line_tc[1]);
					ORACLE_ASSERT(u >= 0);
					ORACLE_ASSERT(u < texture_info.size[0]);
					ORACLE_ASSERT(v >= 0);
					ORACLE_ASSERT(v < texture_info.size[1]);
					const size_t texel_address = (size_t)(u + v * texture_info.size[0]);
					const typename TC::Raw texel = texture_data[texel_address];
					const ColorVecI texel_vec = TC::to_veci(texel);
					const ColorVecI texel_vec_lighted = cvi_shr<16>(cvi_mul(texel_vec, line_light));

					if (BLENDING_MODE == BLEND_NONE)
						dst_pixel = C::from_veci(texel_vec_lighted);
					else if (BLENDING_MODE == BLEND_AVERAGE)
						dst_pixel = C::average(dst_pixel, C::from_veci(texel_vec_lighted));
					else if (BLENDING_MODE == BLEND_ADDITIVE)
						dst_pixel = C::saturated_sum(dst_pixel, C::from_veci(texel_vec_lighted));
					else if (BLENDING_MODE == BLEND_ALPHA_TEST)
					{
						const Raw texel_converted = C::from_veci(texel_vec_lighted);
						if (TC::test_alpha(texel))
							dst_pixel = texel_converted;
					}
					else if (BLENDING_MODE == BLEND_ALPHA_BLEND)
					{
						const int32_t alpha = TC::get_alpha(texel);
						const ColorVecI dst_vec = C::to_veci(dst_pixel);
						dst_pixel = C::from_veci(cvi_shr<8>(cvi_add(cvi_mul_scalar(dst_vec, 255 - alpha), cvi_mul_scalar(texel_vec_lighted, alpha))));
					}

					line_tc[0] += d_tc_dx[0];
					line_tc[1] += d_tc_dx[1];
					line_light = cvi_add(line_light, d_light_dx_vec);
				}
			}
			x_left += left_side.dx_dy;
			x_right += right_side.dx_dy;
			for (int i = 0; i < 2; ++i)
				tc_left[i] += d_tc_dy_left[i];
			for (int i = 0; i < 3; ++i)
				light_left[i] += d_light_left[i];
		}
	}

	// :765-982
	template <class TC, int BLENDING_MODE>
	void fill_triangle_impl(const sqw_triangle_point vertices[3], const TextureInfo& texture_info, const typename TC::Raw* texture_data)
	{
		int upper_index, lower_index;
		if (vertices[0].y >= vertices[1].y && vertices[0].y >= vertices[2].y)
		{
			upper_index = 0;
			lower_index = vertices[1].y < vertices[2].y ? 1 : 2;
		}
		else if (vertices[1].y >= vertices[0].y && vertices[1].y >= vertices[2].y)
		{
			upper_index = 1;
			lower_index = vertices[0].y < vertices[2].y ? 0 : 2;
		}
		else
		{
			upper_index = 2;
			lower_index = vertices[0].y < vertices[1].y ? 0 : 1;
		}
		const int middle_index = 3 - upper_index - lower_index;

		const sqw_triangle_point& upper_vertex = vertices[upper_index];
		const sqw_triangle_point& lower_vertex = vertices[lower_index];
		const sqw_triangle_point& middle_vertex = vertices[middle_index];

		const Fixed16 long_edge_dy = std::max(upper_vertex.y - lower_vertex.y, FIXED16_HALF);
		const Fixed16 lower_part_dy = std::max(middle_vertex.y - lower_vertex.y, FIXED16_HALF);
		const Fixed16 upper_part_dy = std::max(upper_vertex.y - middle_vertex.y, FIXED16_HALF);

		const Fixed16 long_edge_dx_dy = fixed16_div(upper_vertex.x - lower_vertex.x, long_edge_dy);
		const Fixed16 long_edge_x_in_middle = lower_vertex.x + fixed16_mul(long_edge_dx_dy, middle_vertex.y - lower_vertex.y);

		const Fixed16 lower_part_dx_dy = fixed16_div(middle_vertex.x - lower_vertex.x, lower_part_dy);
		const Fixed16 upper_part_dx_dy = fixed16_div(upper_vertex.x - middle_vertex.x, upper_part_dy);

		Fixed16 long_edge_d_tc_dy[2], long_edge_tc_in_middle[2], d_tc_dy_lower[2], d_tc_dy_upper[2];
		for (int i = 0; i < 2; ++i)
		{
			long_edge_d_tc_dy[i] = fixed16_div(upper_vertex.tc[i] - lower_vertex.tc[i], long_edge_dy);
			long_edge_tc_in_middle[i] = lower_vertex.tc[i] + fixed16_mul(long_edge_d_tc_dy[i], lower_part_dy);
			d_tc_dy_lower[i] = fixed16_div(middle_vertex.tc[i] - lower_vertex.tc[i], lower_part_dy);
			d_tc_dy_upper[i] = fixed16_div(upper_vertex.tc[i] - middle_vertex.tc[i], upper_part_dy);
		}
		Fixed16 long_edge_d_light_dy[3], long_edge_light_in_middle[3], d_light_dy_lower[3], d_light_dy_upper[3];
		for (int i = 0; i < 3; ++i)
		{
			long_edge_d_light_dy[i] = fixed16_div(upper_vertex.light[i] - lower_vertex.light[i], long_edge_dy);
			long_edge_light_in_middle[i] = lower_vertex.light[i] + fixed16_mul(long_edge_d_light_dy[i], lower_part_dy);
			d_light_dy_lower[i] = fixed16_div(middle_vertex.light[i] - lower_vertex.light[i], lower_part_dy);
			d_light_dy_upper[i] = fixed16_div(upper_vertex.light[i] - middle_vertex.light[i], upper_part_dy);
		}

		if (long_edge_x_in_middle >= middle_vertex.x)
		{
			const Fixed16 middle_dx = std::max(long_edge_x_in_middle - middle_vertex.x, FIXED16_HALF);
			Fixed16 d_tc_dx[2], d_light_dx[3];
			for (int i = 0; i < 2; ++i)
				d_tc_dx[i] = fixed16_div(long_edge_tc_in_middle[i] - middle_vertex.tc[i], middle_dx);
			for (int i = 0; i < 3; ++i)
				d_light_dx[i] = fixed16_div(long_edge_light_in_middle[i] - middle_vertex.light[i], middle_dx);

			fill_triangle_part<TC, BLENDING_MODE>(
				lower_vertex.y, middle_vertex.y, PolygonSide{lower_vertex.x, lower_part_dx_dy}, PolygonSide{lower_vertex.x, long_edge_dx_dy}, lower_vertex.tc,
				d_tc_dy_lower, d_tc_dx, lower_vertex.light, d_light_dy_lower, d_light_dx, texture_info, texture_data);
			fill_triangle_part<TC, BLENDING_MODE>(
				middle_vertex.y, upper_vertex.y, PolygonSide{middle_vertex.x, upper_part_dx_dy}, PolygonSide{long_edge_x_in_middle, long_edge_dx_dy},
				middle_vertex.tc, d_tc_dy_upper, d_tc_dx, middle_vertex.light, d_light_dy_upper, d_light_dx, texture_info, texture_data);
		}
		else
		{
			const Fixed16 middle_dx = std::max(middle_vertex.x - long_edge_x_in_middle, FIXED16_HALF);
			Fixed16 d_tc_dx[2], d_light_dx[3];
			for (int i = 0; i < 2; ++i)
				d_tc_dx[i] = fixed16_div(middle_vertex.tc[i] - long_edge_tc_in_middle[i], middle_dx);
			for (int i = 0; i < 3; ++i)
				d_light_dx[i] = fixed16_div(middle_vertex.light[i] - long_edge_light_in_middle[i], middle_dx);

			fill_triangle_part<TC, BLENDING_MODE>(
				lower_vertex.y, middle_vertex.y, PolygonSide{lower_vertex.x, long_edge_dx_dy}, PolygonSide{lower_vertex.x, lower_part_dx_dy}, lower_vertex.tc,
				long_edge_d_tc_dy, d_tc_dx, lower_vertex.light, long_edge_d_light_dy, d_light_dx, texture_info, texture_data);
			fill_triangle_part<TC, BLENDING_MODE>(
				middle_vertex.y, upper_vertex.y, PolygonSide{long_edge_x_in_middle, long_edge_dx_dy}, PolygonSide{middle_vertex.x, upper_part_dx_dy},
				long_edge_tc_in_middle, long_edge_d_tc_dy, d_tc_dx, long_edge_light_in_middle, long_edge_d_light_dy, d_light_dx, texture_info, texture_data);
		}
	}

	template <class TC>
	void fill_triangle(const sqw_triangle_point v[3], const TextureInfo& ti, const typename TC::Raw* tex, int blend) // :731-763
	{
		switch (blend)
		{
		case BLEND_NONE: fill_triangle_impl<TC, BLEND_NONE>(v, ti, tex); break;
		case BLEND_AVERAGE: fill_triangle_impl<TC, BLEND_AVERAGE>(v, ti, tex); break;
		case BLEND_ADDITIVE: fill_triangle_impl<TC, BLEND_ADDITIVE>(v, ti, tex); break;
		case BLEND_ALPHA_TEST: fill_triangle_impl<TC, BLEND_ALPHA_TEST>(v, ti, tex); break;
		case BLEND_ALPHA_BLEND: fill_triangle_impl<TC, BLEND_ALPHA_BLEND>(v, ti, tex); break;
		}
	}
};

struct DrawCall
{
	void* color_buffer;
	int32_t width, height, pitch;
	const sqw_draw_cmd* cmds;
	uint32_t num_cmds;
	const sqw_polygon_point* polygon_points;
	const sqw_triangle_point* triangle_points;
	const void* surfaces;
	const OracleImage* images;
	uint32_t num_images;
	const void* const* view_textures; // colour buffers of other views (portal textures), indexed by tex_ref
};

template <class C>
void run_cmds(const DrawCall& d, ClipRect clip)
{
	typedef typename C::Raw Raw;
	Rasterizer<C> r{(Raw*)d.color_buffer, d.pitch, clip};
	for (uint32_t i = 0; i < d.num_cmds; ++i)
	{
		const sqw_draw_cmd& c = d.cmds[i];
		const TextureInfo ti{{c.tex_size[0], c.tex_size[1]}};
		if (c.kind == SQW_CMD_POLYGON)
		{
			const Raw* tex = nullptr;
			if (c.tex_kind == SQW_TEX_SURFACE)
				tex = (const Raw*)d.surfaces + c.tex_ref;
			else if (c.tex_kind == SQW_TEX_IMAGE)
				tex = (const Raw*)d.images[c.tex_ref].pixels;
			else
				tex = (const Raw*)d.view_textures[c.tex_ref];
			const DepthEquation de{c.depth_eq[0], c.depth_eq[1], c.depth_eq[2]};
			const TexCoordEquation te{{c.tc_eq[0], c.tc_eq[1]}, {c.tc_eq[2], c.tc_eq[3]}, {c.tc_eq[4], c.tc_eq[5]}};
			r.fill_polygon(d.polygon_points + c.vertex_offset, c.num_vertices, de, te, ti, tex, c.tc_mode, c.blend);
		}
		else
		{
			const OracleImage& img = d.images[c.tex_ref];
			if (img.kind == SQW_IMAGE_COLOR64)
				r.template fill_triangle<Color64>(d.triangle_points + c.vertex_offset, ti, (const uint64_t*)img.pixels, c.blend);
			else
				r.template fill_triangle<Color32>(d.triangle_points + c.vertex_offset, ti, (const uint32_t*)img.pixels, c.blend);
		}
	}
}

// rect_splitting.rs:23-78
struct RectF
{
	float min_x, min_y, max_x, max_y;
};
void split_rect_r(const RectF& r, uint32_t num_parts, std::vector<RectF>& out)
{
	if (num_parts <= 1)
	{
		out.push_back(r);
		return;
	}
	const uint32_t num_parts0 = num_parts / 2;
	const uint32_t num_parts1 = num_parts - num_parts0;
	const float ratio0 = (float)num_parts0 / (float)num_parts;
	const float ratio1 = (float)num_parts1 / (float)num_parts;
	RectF r0, r1;
	const float width = r.max_x - r.min_x;
	const float height = r.max_y - r.min_y;
	if (width >= height)
	{
		const float middle = r.max_x * ratio0 + r.min_x * ratio1;
		r0 = {r.min_x, r.min_y, middle, r.max_y};
		r1 = {middle, r.min_y, r.max_x, r.max_y};
	}
	else
	{
		const float middle = r.max_y * ratio0 + r.min_y * ratio1;
		r0 = {r.min_x, r.min_y, r.max_x, middle};
		r1 = {r.min_x, middle, r.max_x, r.max_y};
	}
	split_rect_r(r0, num_parts0, out);
	split_rect_r(r1, num_parts1, out);
}

template <class C>
void run_draw(const DrawCall& d, int num_threads)
{
	if (num_threads <= 1)
	{
		run_cmds<C>(d, ClipRect{0, 0, d.width, d.height}); // partial_renderer.rs:1207-1233
		return;
	}
	// partial_renderer.rs:1237-1281: one Rasterizer per thread with its own clip rect; every thread walks the
	// whole command list (the reference walks the whole BSP per thread).
	std::vector<RectF> rects;
	split_rect_r(RectF{0.0f, 0.0f, (float)d.width, (float)d.height}, (uint32_t)std::min(num_threads, 64), rects);
	ThreadPool::instance().parallel_for((uint32_t)rects.size(), num_threads, [&](uint32_t i) {
		const RectF& rc = rects[i];
		run_cmds<C>(d, ClipRect{(int32_t)rc.min_x, (int32_t)rc.min_y, (int32_t)rc.max_x, (int32_t)rc.max_y});
	});
}

} // namespace
} // namespace oracle

extern "C" int oracle_draw(
	int color_bits, void* color_buffer, int32_t width, int32_t height, int32_t pitch, const sqw_draw_cmd* cmds, uint32_t num_cmds,
	const sqw_polygon_point* polygon_points, const sqw_triangle_point* triangle_points, const void* surfaces, const oracle::OracleImage* images,
	uint32_t num_images, const void* const* view_textures, const int32_t* clip_rect_or_null, int num_threads)
{
	oracle::DrawCall d{color_buffer, width, height, pitch, cmds, num_cmds, polygon_points, triangle_points, surfaces, images, num_images, view_textures};
	if (clip_rect_or_null)
	{
		oracle::ClipRect cr{clip_rect_or_null[0], clip_rect_or_null[1], clip_rect_or_null[2], clip_rect_or_null[3]};
		if (color_bits == 32)
			oracle::run_cmds<oracle::Color32>(d, cr);
		else
			oracle::run_cmds<oracle::Color64>(d, cr);
		return 0;
	}
	if (color_bits == 32)
		oracle::run_draw<oracle::Color32>(d, num_threads);
	else if (color_bits == 64)
		oracle::run_draw<oracle::Color64>(d, num_threads);
	else
		return -1;
	return 0;
}

// draw_background (partial_renderer.rs:3976-4009)
extern "C" void oracle_fill(int color_bits, void* color_buffer, uint64_t num_pixels, uint64_t raw_color)
{
	if (color_bits == 32)
		std::fill_n((uint32_t*)color_buffer, num_pixels, (uint32_t)raw_color);
	else
		std::fill_n((uint64_t*)color_buffer, num_pixels, raw_color);
}
