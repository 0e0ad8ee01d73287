// C-ABI implementation (include/sqw_cuda.h): contexts, device memory, host-side job/command preparation and the
// kernel launch sequence of one frame.  No CPU fallback: every entry fails with SQW_ERR_NO_DEVICE / SQW_ERR_CUDA
// when the CUDA device is unavailable.
#include "sqw_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace sqw;

static_assert(sizeof(sqw_surface_job) == 136, "sqw_surface_job layout");
static_assert(sizeof(sqw_light) == 80, "sqw_light layout");
static_assert(sizeof(sqw_draw_cmd) == 64, "sqw_draw_cmd layout");
static_assert(sizeof(sqw_polygon_point) == 8, "sqw_polygon_point layout");
static_assert(sizeof(sqw_triangle_point) == 28, "sqw_triangle_point layout");
static_assert(sizeof(sqw_directional_lightmap_element) == 40, "sqw_directional_lightmap_element layout");
static_assert(sizeof(sqw_post_params) == 32, "sqw_post_params layout");

namespace
{

thread_local std::string g_create_error;

// Grow-only device buffer.
struct DevBuf
{
	void* ptr = nullptr;
	size_t capacity = 0;
	cudaError_t reserve(size_t bytes)
	{
		if (bytes <= capacity)
			return cudaSuccess;
		if (ptr)
		{
			cudaError_t e = cudaFree(ptr);
			ptr = nullptr;
			capacity = 0;
			if (e != cudaSuccess)
				return e;
		}
		const size_t rounded = (bytes + (bytes >> 2) + 255) & ~size_t(255);
		cudaError_t e = cudaMalloc(&ptr, rounded);
		if (e != cudaSuccess)
		{
			ptr = nullptr;
			return e;
		}
		capacity = rounded;
		return cudaSuccess;
	}
	void release()
	{
		if (ptr)
			cudaFree(ptr);
		ptr = nullptr;
		capacity = 0;
	}
};

struct ImageHost
{
	void* pixels;
	uint32_t width, height;
	int kind;
};

} // namespace

struct sqw_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	float* d_lightmaps = nullptr;
	sqw_directional_lightmap_element* d_directional_lightmaps = nullptr;
	uint64_t num_lightmap_elements = 0;
	std::vector<ImageHost> images;
	ImageDev* d_images = nullptr;
	uint32_t max_images = 0;
	std::vector<sqw_view*> views;
	std::string last_error;
	// per-kernel timing (sqw_profile_begin / _end)
	bool profiling = false;
	struct TimedLaunch
	{
		cudaEvent_t start, stop;
		int kind;
	};
	std::vector<TimedLaunch> timed;
	std::vector<cudaEvent_t> event_pool;
	cudaEvent_t get_event()
	{
		if (!event_pool.empty())
		{
			cudaEvent_t e = event_pool.back();
			event_pool.pop_back();
			return e;
		}
		cudaEvent_t e = nullptr;
		cudaEventCreate(&e);
		return e;
	}
};

namespace
{
// RAII bracket around one kernel launch when profiling is on.
struct LaunchTimer
{
	sqw_ctx* ctx;
	sqw_ctx::TimedLaunch t{};
	LaunchTimer(sqw_ctx* c, int kind) : ctx(c)
	{
		if (!ctx->profiling)
			return;
		t.kind = kind;
		t.start = ctx->get_event();
		t.stop = ctx->get_event();
		cudaEventRecord(t.start, ctx->stream);
	}
	~LaunchTimer()
	{
		if (!ctx->profiling)
			return;
		cudaEventRecord(t.stop, ctx->stream);
		ctx->timed.push_back(t);
	}
};
} // namespace

struct sqw_view
{
	sqw_ctx* ctx = nullptr;
	uint32_t index = 0;
	uint32_t width = 0, height = 0;
	int color_bits = 0;
	void* d_color = nullptr;
	uint32_t* d_ldr = nullptr; // post-processed picture (64-bit views)
	DevBuf surfaces;
	DevBuf spans;
	DevBuf band_lists;
	uint64_t* d_bloom[2] = {nullptr, nullptr};
	unsigned long long* d_sums = nullptr; // 3 x u64
	uint64_t sums_divisor = 1;
	sqw_frame* immediate = nullptr;
};

struct SurfaceGroup
{
	SurfaceVariant variant;
	uint32_t band_offset, num_bands;
};

struct sqw_frame
{
	sqw_view* view = nullptr;
	// surfaces
	DevBuf d_jobs, d_bands, d_lights, d_shadow_maps;
	std::vector<SurfaceBand> h_bands;
	std::vector<SurfaceGroup> groups;
	float cam_pos[3] = {0, 0, 0};
	bool has_surfaces = false;
	// draw
	DevBuf d_cmds, d_cmd_dev, d_polygon_points, d_triangle_points, d_cmd_span_base, d_band_offsets;
	std::vector<CmdDev> h_cmd_dev;
	std::vector<uint32_t> h_cmd_span_base, h_band_offsets;
	uint32_t num_cmds = 0, total_spans = 0, num_bands = 0, total_band_entries = 0;
	bool clear = false;
	uint64_t clear_color = 0;
	bool has_draw = false;
	// post
	sqw_post_params post{};
	bool has_post = false;
	sqw_frame_stats stats{};
};

namespace
{

int fail(sqw_ctx* ctx, int code, const std::string& msg)
{
	if (ctx)
		ctx->last_error = msg;
	else
		g_create_error = msg;
	return code;
}
int fail_cuda(sqw_ctx* ctx, cudaError_t e, const char* what)
{
	return fail(ctx, e == cudaErrorMemoryAllocation ? SQW_ERR_OUT_OF_MEMORY : SQW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define SQW_CUDA(ctx, call)                      \
	do                                           \
	{                                            \
		cudaError_t e__ = (call);                \
		if (e__ != cudaSuccess)                  \
			return fail_cuda((ctx), e__, #call); \
	} while (0)

size_t image_element_size(int kind)
{
	switch (kind)
	{
	case SQW_IMAGE_TEXTURE_ELEMENT: return 8;
	case SQW_IMAGE_COLOR32: return 4;
	case SQW_IMAGE_COLOR64: return 8;
	default: return 0;
	}
}

inline int32_t fixed16_round_to_int_h(int32_t x) { return (int32_t)((uint32_t)x + 0x8000u) >> 16; }

uint32_t variant_key(const SurfaceVariant& v) // bucket index: 24 lit variants + emissive-only
{
	if (v.emissive_only)
		return 24;
	return (uint32_t)v.directional * 12 + (uint32_t)v.dynamic_lights * 6 + (uint32_t)v.normal_map * 3 + (uint32_t)v.specular;
}

int upload(sqw_ctx* ctx, DevBuf& buf, const void* src, size_t bytes)
{
	if (bytes == 0)
		return SQW_OK;
	SQW_CUDA(ctx, buf.reserve(bytes));
	SQW_CUDA(ctx, cudaMemcpyAsync(buf.ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return SQW_OK;
}

} // namespace

extern "C"
{

int sqw_abi_version(void) { return SQW_ABI_VERSION; }

const char* sqw_last_error(const sqw_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int sqw_create(sqw_ctx** out_ctx, int device, const sqw_map_desc* map, void* external_stream)
{
	if (!out_ctx || !map)
		return fail(nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_create: null argument");
	*out_ctx = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0)
		return fail(nullptr, SQW_ERR_NO_DEVICE, std::string("sqw_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
	if (device < 0 || device >= count)
		return fail(nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_create: bad device index");
	SQW_CUDA(nullptr, cudaSetDevice(device));

	sqw_ctx* ctx = new (std::nothrow) sqw_ctx();
	if (!ctx)
		return fail(nullptr, SQW_ERR_OUT_OF_MEMORY, "sqw_create: host allocation failed");
	ctx->device = device;
	if (external_stream)
		ctx->stream = static_cast<cudaStream_t>(external_stream);
	else
	{
		e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
		if (e != cudaSuccess)
		{
			delete ctx;
			return fail_cuda(nullptr, e, "cudaStreamCreate");
		}
		ctx->own_stream = true;
	}
	ctx->max_images = map->max_images ? map->max_images : 4096;
	ctx->num_lightmap_elements = map->num_lightmap_elements;
	auto bail = [&](cudaError_t err, const char* what) {
		int rc = fail_cuda(nullptr, err, what);
		sqw_destroy(ctx);
		return rc;
	};
	e = cudaMalloc(&ctx->d_images, sizeof(ImageDev) * ctx->max_images);
	if (e != cudaSuccess)
		return bail(e, "cudaMalloc(images)");
	const size_t n = (size_t)map->num_lightmap_elements;
	if (n && map->lightmaps_data)
	{
		e = cudaMalloc(&ctx->d_lightmaps, n * 12);
		if (e != cudaSuccess)
			return bail(e, "cudaMalloc(lightmaps)");
		e = cudaMemcpy(ctx->d_lightmaps, map->lightmaps_data, n * 12, cudaMemcpyHostToDevice);
		if (e != cudaSuccess)
			return bail(e, "cudaMemcpy(lightmaps)");
	}
	if (n && map->directional_lightmaps_data)
	{
		e = cudaMalloc(&ctx->d_directional_lightmaps, n * sizeof(sqw_directional_lightmap_element));
		if (e != cudaSuccess)
			return bail(e, "cudaMalloc(directional lightmaps)");
		e = cudaMemcpy(ctx->d_directional_lightmaps, map->directional_lightmaps_data, n * sizeof(sqw_directional_lightmap_element), cudaMemcpyHostToDevice);
		if (e != cudaSuccess)
			return bail(e, "cudaMemcpy(directional lightmaps)");
	}
	*out_ctx = ctx;
	return SQW_OK;
}

void sqw_destroy(sqw_ctx* ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	if (ctx->stream)
		cudaStreamSynchronize(ctx->stream);
	while (!ctx->views.empty())
		sqw_view_destroy(ctx->views.back());
	for (auto& img : ctx->images)
		if (img.pixels)
			cudaFree(img.pixels);
	if (ctx->d_images)
		cudaFree(ctx->d_images);
	if (ctx->d_lightmaps)
		cudaFree(ctx->d_lightmaps);
	if (ctx->d_directional_lightmaps)
		cudaFree(ctx->d_directional_lightmaps);
	for (auto& t : ctx->timed)
	{
		cudaEventDestroy(t.start);
		cudaEventDestroy(t.stop);
	}
	for (cudaEvent_t e : ctx->event_pool)
		cudaEventDestroy(e);
	if (ctx->own_stream && ctx->stream)
		cudaStreamDestroy(ctx->stream);
	delete ctx;
}

int sqw_upload_image(sqw_ctx* ctx, int kind, uint32_t width, uint32_t height, const void* pixels)
{
	if (!ctx || !pixels || width == 0 || height == 0 || image_element_size(kind) == 0)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_image: bad argument");
	if (ctx->images.size() >= ctx->max_images)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_image: image table full");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t bytes = (size_t)width * height * image_element_size(kind);
	void* d = nullptr;
	SQW_CUDA(ctx, cudaMalloc(&d, bytes));
	cudaError_t e = cudaMemcpy(d, pixels, bytes, cudaMemcpyHostToDevice);
	if (e != cudaSuccess)
	{
		cudaFree(d);
		return fail_cuda(ctx, e, "cudaMemcpy(image)");
	}
	const int id = (int)ctx->images.size();
	ctx->images.push_back(ImageHost{d, width, height, kind});
	const ImageDev dev{d, width, height};
	e = cudaMemcpy(ctx->d_images + id, &dev, sizeof(dev), cudaMemcpyHostToDevice);
	if (e != cudaSuccess)
		return fail_cuda(ctx, e, "cudaMemcpy(image table)");
	return id;
}

int sqw_update_image(sqw_ctx* ctx, int image_id, const void* pixels)
{
	if (!ctx || !pixels || image_id < 0 || (size_t)image_id >= ctx->images.size())
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_update_image: bad argument");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const ImageHost& img = ctx->images[(size_t)image_id];
	SQW_CUDA(ctx, cudaMemcpyAsync(img.pixels, pixels, (size_t)img.width * img.height * image_element_size(img.kind), cudaMemcpyHostToDevice, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return SQW_OK;
}

int sqw_view_create(sqw_ctx* ctx, sqw_view** out_view, uint32_t width, uint32_t height, int color_bits)
{
	if (!ctx || !out_view || width == 0 || height == 0 || width > 16384 || height > 16384 || (color_bits != 32 && color_bits != 64))
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_view_create: bad argument");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	sqw_view* v = new (std::nothrow) sqw_view();
	if (!v)
		return fail(ctx, SQW_ERR_OUT_OF_MEMORY, "sqw_view_create: host allocation failed");
	v->ctx = ctx;
	v->width = width;
	v->height = height;
	v->color_bits = color_bits;
	const size_t px = (size_t)width * height;
	cudaError_t e = cudaMalloc(&v->d_color, px * (color_bits / 8));
	if (e == cudaSuccess)
		e = cudaMemset(v->d_color, 0, px * (color_bits / 8));
	if (e == cudaSuccess && color_bits == 64)
		e = cudaMalloc(&v->d_ldr, px * 4);
	if (e == cudaSuccess && color_bits == 64)
	{
		// bloom buffers at the largest size (scale 2)
		const size_t bpx = (size_t)(width / 2) * (height / 2) + 1;
		e = cudaMalloc(&v->d_bloom[0], bpx * 8);
		if (e == cudaSuccess)
			e = cudaMalloc(&v->d_bloom[1], bpx * 8);
	}
	if (e == cudaSuccess)
		e = cudaMalloc(&v->d_sums, 3 * sizeof(unsigned long long));
	if (e == cudaSuccess)
		e = cudaMemset(v->d_sums, 0, 3 * sizeof(unsigned long long));
	if (e != cudaSuccess)
	{
		int rc = fail_cuda(ctx, e, "sqw_view_create: device allocation");
		v->index = (uint32_t)ctx->views.size();
		ctx->views.push_back(v);
		sqw_view_destroy(v);
		return rc;
	}
	v->index = (uint32_t)ctx->views.size();
	ctx->views.push_back(v);
	*out_view = v;
	return SQW_OK;
}

void sqw_view_destroy(sqw_view* view)
{
	if (!view)
		return;
	sqw_ctx* ctx = view->ctx;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (view->immediate)
		sqw_frame_destroy(view->immediate);
	if (view->d_color)
		cudaFree(view->d_color);
	if (view->d_ldr)
		cudaFree(view->d_ldr);
	view->surfaces.release();
	view->spans.release();
	view->band_lists.release();
	for (auto& b : view->d_bloom)
		if (b)
			cudaFree(b);
	if (view->d_sums)
		cudaFree(view->d_sums);
	// keep indices of the other views stable: leave a hole
	for (auto& p : ctx->views)
		if (p == view)
			p = nullptr;
	while (!ctx->views.empty() && ctx->views.back() == nullptr)
		ctx->views.pop_back();
	delete view;
}

int sqw_frame_create(sqw_view* view, sqw_frame** out_frame)
{
	if (!view || !out_frame)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_frame* f = new (std::nothrow) sqw_frame();
	if (!f)
		return fail(view->ctx, SQW_ERR_OUT_OF_MEMORY, "sqw_frame_create: host allocation failed");
	f->view = view;
	*out_frame = f;
	return SQW_OK;
}

void sqw_frame_destroy(sqw_frame* f)
{
	if (!f)
		return;
	sqw_ctx* ctx = f->view->ctx;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	DevBuf* bufs[] = {&f->d_jobs, &f->d_bands, &f->d_lights, &f->d_shadow_maps, &f->d_cmds, &f->d_cmd_dev, &f->d_polygon_points, &f->d_triangle_points,
					  &f->d_cmd_span_base, &f->d_band_offsets};
	for (DevBuf* b : bufs)
		b->release();
	if (f->view->immediate == f)
		f->view->immediate = nullptr;
	delete f;
}

int sqw_frame_set_surfaces(
	sqw_frame* f, const sqw_surface_job* jobs, uint32_t num_jobs, const sqw_light* lights, uint32_t num_lights, const float* shadow_maps,
	uint64_t shadow_maps_len, const float cam_pos[3], uint64_t total_surface_pixels)
{
	if (!f || (num_jobs && !jobs) || (num_lights && !lights) || !cam_pos)
		return fail(f ? f->view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: null argument");
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t C = (size_t)view->color_bits / 8;

	// Validate and group jobs by kernel variant (the reference's const-generic dispatch, surfaces.rs:124-423).
	std::vector<SurfaceBand> buckets[25];
	sqw_frame_stats& st = f->stats;
	st.surface_texels = st.emissive_texels = st.lightmap_bytes = 0;
	for (uint32_t i = 0; i < num_jobs; ++i)
	{
		const sqw_surface_job& j = jobs[i];
		const uint64_t texels = (uint64_t)j.surface_size[0] * j.surface_size[1];
		if ((uint64_t)j.surface_pixels_offset + texels > total_surface_pixels)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: job writes outside the surface pool");
		const bool diffuse = (j.flags & SQW_SURF_APPLY_DIFFUSE) != 0;
		const bool emissive = (j.flags & SQW_SURF_EMISSIVE) != 0;
		if (!diffuse && !emissive)
			continue; // nothing is written (partial_renderer.rs:1700-1819)
		if (diffuse && (j.texture_image >= ctx->images.size() || ctx->images[j.texture_image].kind != SQW_IMAGE_TEXTURE_ELEMENT))
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: bad texture image");
		if (emissive && (j.emissive_image >= ctx->images.size() || ctx->images[j.emissive_image].kind != SQW_IMAGE_COLOR32))
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: bad emissive image");
		if (j.num_lights > SQW_MAX_POLYGON_LIGHTS)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: too many lights for a polygon");
		for (uint32_t l = 0; l < j.num_lights; ++l)
			if (j.lights[l] >= num_lights)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: light index out of range");
		SurfaceVariant v{};
		v.color64 = view->color_bits == 64;
		if (!diffuse)
			v.emissive_only = 1;
		else
		{
			v.directional = (j.flags & SQW_SURF_DIRECTIONAL_LIGHTMAP) != 0;
			v.dynamic_lights = j.num_lights != 0;
			v.normal_map = (j.flags & SQW_SURF_HAS_NORMAL_MAP) != 0;
			v.specular = (j.flags & SQW_SURF_HAS_NON_ONE_ROUGHNESS) ? ((j.flags & SQW_SURF_IS_METAL) ? 2 : 1) : 0;
			if (j.lightmap_offset != 0)
			{
				const uint64_t lm = (uint64_t)j.lightmap_size[0] * j.lightmap_size[1];
				if (j.lightmap_offset + lm > ctx->num_lightmap_elements || j.lightmap_scale_log2 > 4 || j.lightmap_size[0] < 2 || j.lightmap_size[1] < 2)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: bad lightmap reference");
				if (v.directional ? !ctx->d_directional_lightmaps : !ctx->d_lightmaps)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: lightmap kind not uploaded");
				// last sampled lightmap texel must exist (debug_assert surfaces.rs:482,541)
				const uint32_t max_u = ((j.surface_size[0] - 1 + j.lightmap_tc_shift[0]) >> j.lightmap_scale_log2) + 1;
				const uint32_t max_v = ((j.surface_size[1] - 1 + j.lightmap_tc_shift[1]) >> j.lightmap_scale_log2) + 1;
				if (texels && (max_u >= j.lightmap_size[0] || max_v >= j.lightmap_size[1] || max_u > 18))
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: surface reads outside its lightmap");
				st.lightmap_bytes += lm * (v.directional ? 40 : 12);
			}
		}
		const uint32_t key = variant_key(v);
		for (uint32_t row0 = 0; row0 < j.surface_size[1]; row0 += SQW_SURF_BAND_ROWS)
			buckets[key].push_back(SurfaceBand{i, row0});
		st.surface_texels += texels;
		if (emissive)
			st.emissive_texels += texels;
	}
	f->h_bands.clear();
	f->groups.clear();
	for (uint32_t key = 0; key < 25; ++key)
	{
		if (buckets[key].empty())
			continue;
		SurfaceGroup g{};
		g.variant.color64 = view->color_bits == 64;
		if (key == 24)
			g.variant.emissive_only = 1;
		else
		{
			g.variant.directional = (key / 12) & 1;
			g.variant.dynamic_lights = (key / 6) & 1;
			g.variant.normal_map = (key / 3) & 1;
			g.variant.specular = key % 3;
		}
		g.band_offset = (uint32_t)f->h_bands.size();
		g.num_bands = (uint32_t)buckets[key].size();
		f->h_bands.insert(f->h_bands.end(), buckets[key].begin(), buckets[key].end());
		f->groups.push_back(g);
	}
	st.num_surface_variants = (uint32_t)f->groups.size();

	SQW_CUDA(ctx, view->surfaces.reserve((size_t)total_surface_pixels * C + 16));
	int rc;
	if ((rc = upload(ctx, f->d_jobs, jobs, (size_t)num_jobs * sizeof(sqw_surface_job))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_bands, f->h_bands.data(), f->h_bands.size() * sizeof(SurfaceBand))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_lights, lights, (size_t)num_lights * sizeof(sqw_light))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_shadow_maps, shadow_maps, (size_t)shadow_maps_len * sizeof(float))) != SQW_OK)
		return rc;
	f->cam_pos[0] = cam_pos[0];
	f->cam_pos[1] = cam_pos[1];
	f->cam_pos[2] = cam_pos[2];
	f->has_surfaces = true;
	return SQW_OK;
}

int sqw_frame_set_draw(
	sqw_frame* f, const sqw_draw_cmd* cmds, uint32_t num_cmds, const sqw_polygon_point* polygon_points, uint32_t num_polygon_points,
	const sqw_triangle_point* triangle_points, uint32_t num_triangle_points, const void* clear_color)
{
	if (!f || (num_cmds && !cmds))
		return fail(f ? f->view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: null argument");
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const int32_t W = (int32_t)view->width, H = (int32_t)view->height;
	const size_t C = (size_t)view->color_bits / 8;
	const uint32_t num_bands = (uint32_t)((H + SQW_TILE_H - 1) / SQW_TILE_H);

	f->h_cmd_dev.resize(num_cmds);
	f->h_cmd_span_base.assign((size_t)num_cmds + 1, 0);
	f->h_band_offsets.assign((size_t)num_bands + 1, 0);
	uint64_t span_total = 0;
	for (uint32_t i = 0; i < num_cmds; ++i)
	{
		const sqw_draw_cmd& c = cmds[i];
		CmdDev& d = f->h_cmd_dev[i];
		std::memset(&d, 0, sizeof(d));
		d.kind = c.kind;
		d.blend = c.blend;
		d.tc_mode = c.tc_mode;
		if (c.blend > SQW_BLEND_ALPHA_BLEND || c.tc_mode > SQW_TC_AFFINE || c.tex_size[0] <= 0 || c.tex_size[1] <= 0)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad command");
		d.tex_width = c.tex_size[0];
		d.tex_height = c.tex_size[1];
		const uint64_t tex_elems = (uint64_t)c.tex_size[0] * (uint64_t)c.tex_size[1];
		int32_t min_x = INT32_MAX, max_x = INT32_MIN, min_y = INT32_MAX, max_y = INT32_MIN;
		if (c.kind == SQW_CMD_POLYGON)
		{
			if (c.num_vertices < 3 || c.num_vertices > 64 || (uint64_t)c.vertex_offset + c.num_vertices > num_polygon_points || !polygon_points)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad polygon vertices");
			for (uint32_t k = 0; k < c.num_vertices; ++k)
			{
				const sqw_polygon_point& p = polygon_points[c.vertex_offset + k];
				min_x = std::min(min_x, p.x);
				max_x = std::max(max_x, p.x);
				min_y = std::min(min_y, p.y);
				max_y = std::max(max_y, p.y);
			}
			// polygon textures have the framebuffer's colour type (rasterizer.rs:41)
			if (c.tex_kind == SQW_TEX_SURFACE)
			{
				if ((uint64_t)c.tex_ref + tex_elems > view->surfaces.capacity / C)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: surface reference outside the pool (build surfaces first)");
				d.texture = reinterpret_cast<const void*>((uintptr_t)c.tex_ref * C);
				d.tex_flags = (view->color_bits == 64 ? 1 : 0) | 2;
			}
			else if (c.tex_kind == SQW_TEX_IMAGE)
			{
				if (c.tex_ref >= ctx->images.size())
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad image reference");
				const ImageHost& img = ctx->images[c.tex_ref];
				if (img.kind != (view->color_bits == 64 ? SQW_IMAGE_COLOR64 : SQW_IMAGE_COLOR32) || tex_elems > (uint64_t)img.width * img.height)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: polygon image must match the view colour type and size");
				d.texture = img.pixels;
				d.tex_flags = view->color_bits == 64 ? 1 : 0;
			}
			else if (c.tex_kind == SQW_TEX_VIEW)
			{
				if (c.tex_ref >= ctx->views.size() || !ctx->views[c.tex_ref] || ctx->views[c.tex_ref] == view)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad view reference");
				const sqw_view* tv = ctx->views[c.tex_ref];
				if (tv->color_bits != view->color_bits || tex_elems > (uint64_t)tv->width * tv->height)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: view texture must match colour type and size");
				d.texture = tv->d_color;
				d.tex_flags = view->color_bits == 64 ? 1 : 0;
			}
			else
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad tex_kind");
		}
		else if (c.kind == SQW_CMD_TRIANGLE)
		{
			if ((uint64_t)c.vertex_offset + 3 > num_triangle_points || !triangle_points)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad triangle vertices");
			for (uint32_t k = 0; k < 3; ++k)
			{
				const sqw_triangle_point& p = triangle_points[c.vertex_offset + k];
				min_x = std::min(min_x, p.x);
				max_x = std::max(max_x, p.x);
				min_y = std::min(min_y, p.y);
				max_y = std::max(max_y, p.y);
			}
			if (c.tex_kind != SQW_TEX_IMAGE || c.tex_ref >= ctx->images.size())
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: triangles need an image texture");
			const ImageHost& img = ctx->images[c.tex_ref];
			if (img.kind == SQW_IMAGE_TEXTURE_ELEMENT || tex_elems > (uint64_t)img.width * img.height)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad triangle texture");
			d.texture = img.pixels;
			d.tex_flags = img.kind == SQW_IMAGE_COLOR64 ? 1 : 0;
		}
		else
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad command kind");

		// Rows that can own a span: [round(min_y), round(max_y)) clipped (rasterizer.rs:316-317, :1000-1001).
		int32_t y0 = std::max(fixed16_round_to_int_h(min_y), 0);
		int32_t y1 = std::min(fixed16_round_to_int_h(max_y), H);
		// Conservative column range: edges stay inside the vertex hull up to fixed-point slope rounding (< 1 px).
		int32_t bx0 = std::max((min_x >> 16) - 1, 0);
		int32_t bx1 = std::min((max_x >> 16) + 2, W);
		if (y0 >= y1 || bx0 >= bx1)
			y0 = y1 = bx0 = bx1 = 0;
		d.y0 = (int16_t)y0;
		d.y1 = (int16_t)y1;
		d.bx0 = (int16_t)bx0;
		d.bx1 = (int16_t)bx1;
		d.span_base = (uint32_t)span_total;
		f->h_cmd_span_base[i] = (uint32_t)span_total;
		span_total += (uint64_t)(y1 - y0);
		if (span_total > 0xFFFFFFFFull)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: too many scan-lines");
		if (y0 < y1)
			for (int32_t b = y0 / SQW_TILE_H; b <= (y1 - 1) / SQW_TILE_H; ++b)
				f->h_band_offsets[(size_t)b + 1]++;
	}
	f->h_cmd_span_base[num_cmds] = (uint32_t)span_total;
	for (uint32_t b = 0; b < num_bands; ++b)
		f->h_band_offsets[(size_t)b + 1] += f->h_band_offsets[b];
	f->num_cmds = num_cmds;
	f->total_spans = (uint32_t)span_total;
	f->num_bands = num_bands;
	f->total_band_entries = f->h_band_offsets[num_bands];
	f->stats.num_spans = span_total;
	f->stats.num_tiles = num_bands * (uint32_t)((W + SQW_TILE_W - 1) / SQW_TILE_W);

	SQW_CUDA(ctx, view->spans.reserve((size_t)span_total * sizeof(SpanRec) + 64));
	SQW_CUDA(ctx, view->band_lists.reserve((size_t)f->total_band_entries * 4 + 64));
	int rc;
	if ((rc = upload(ctx, f->d_cmds, cmds, (size_t)num_cmds * sizeof(sqw_draw_cmd))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_cmd_dev, f->h_cmd_dev.data(), (size_t)num_cmds * sizeof(CmdDev))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_polygon_points, polygon_points, (size_t)num_polygon_points * sizeof(sqw_polygon_point))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_triangle_points, triangle_points, (size_t)num_triangle_points * sizeof(sqw_triangle_point))) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_cmd_span_base, f->h_cmd_span_base.data(), f->h_cmd_span_base.size() * 4)) != SQW_OK)
		return rc;
	if ((rc = upload(ctx, f->d_band_offsets, f->h_band_offsets.data(), f->h_band_offsets.size() * 4)) != SQW_OK)
		return rc;
	f->clear = clear_color != nullptr;
	f->clear_color = 0;
	if (clear_color)
		std::memcpy(&f->clear_color, clear_color, C);
	f->has_draw = true;
	return SQW_OK;
}

int sqw_frame_set_post(sqw_frame* f, const sqw_post_params* params)
{
	if (!f || !params)
		return fail(f ? f->view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_post: null argument");
	if (f->view->color_bits != 64)
		return fail(f->view->ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_post: post-processing needs a 64-bit view");
	f->post = *params;
	f->has_post = true;
	return SQW_OK;
}

// compute_gaussian_kernel (postprocessor.rs:1026-1045) — host-side set-up like in the reference.
static void compute_gaussian_kernel(float sigma, int radius, float result[33])
{
	for (int i = 0; i < 33; ++i)
		result[i] = 0.0f;
	for (int x = -radius; x <= radius; ++x)
	{
		const float samples[4] = {-0.375f, -0.125f, 0.125f, 0.375f};
		float val = 0.0f;
		for (float sample : samples)
		{
			const float coord = (float)x + sample;
			val += std::exp(-0.5f * (coord / sigma) * (coord / sigma)) / (sigma * std::sqrt(6.28318530717958647692f));
		}
		result[x + radius] = val / 4.0f;
	}
}

int sqw_frame_submit(sqw_frame* f, uint32_t stages)
{
	if (!f)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	uint32_t launches = 0;

	if ((stages & SQW_STAGE_SURFACES) && f->has_surfaces)
	{
		SurfaceLaunch l{};
		l.jobs = static_cast<const sqw_surface_job*>(f->d_jobs.ptr);
		l.images = ctx->d_images;
		l.lightmaps = ctx->d_lightmaps;
		l.directional_lightmaps = ctx->d_directional_lightmaps;
		l.lights = static_cast<const sqw_light*>(f->d_lights.ptr);
		l.shadow_maps = static_cast<const float*>(f->d_shadow_maps.ptr);
		l.cam_pos[0] = f->cam_pos[0];
		l.cam_pos[1] = f->cam_pos[1];
		l.cam_pos[2] = f->cam_pos[2];
		l.surfaces = view->surfaces.ptr;
		for (const SurfaceGroup& g : f->groups)
		{
			l.bands = static_cast<const SurfaceBand*>(f->d_bands.ptr) + g.band_offset;
			l.num_bands = g.num_bands;
			{
				LaunchTimer timer(ctx, SQW_KERNEL_SURFACE_BUILD);
				SQW_CUDA(ctx, launch_surface_build(g.variant, l, s));
			}
			++launches;
		}
	}

	if ((stages & SQW_STAGE_RASTER) && f->has_draw)
	{
		RasterLaunch l{};
		l.cmds = static_cast<const sqw_draw_cmd*>(f->d_cmds.ptr);
		l.cmd_dev = static_cast<const CmdDev*>(f->d_cmd_dev.ptr);
		l.num_cmds = f->num_cmds;
		l.polygon_points = static_cast<const sqw_polygon_point*>(f->d_polygon_points.ptr);
		l.triangle_points = static_cast<const sqw_triangle_point*>(f->d_triangle_points.ptr);
		l.spans = static_cast<SpanRec*>(view->spans.ptr);
		l.cmd_span_base = static_cast<const uint32_t*>(f->d_cmd_span_base.ptr);
		l.total_spans = f->total_spans;
		l.band_lists = static_cast<uint32_t*>(view->band_lists.ptr);
		l.band_offsets = static_cast<const uint32_t*>(f->d_band_offsets.ptr);
		l.num_bands = f->num_bands;
		l.width = (int32_t)view->width;
		l.height = (int32_t)view->height;
		l.pitch = (int32_t)view->width;
		l.color_buffer = view->d_color;
		l.surfaces_base = view->surfaces.ptr;
		l.color64 = view->color_bits == 64;
		l.clear = f->clear;
		l.clear_color = f->clear_color;
		if (f->total_spans)
		{
			LaunchTimer timer(ctx, SQW_KERNEL_RASTER_SETUP);
			SQW_CUDA(ctx, launch_raster_setup(l, s));
			++launches;
		}
		if (f->num_cmds)
		{
			LaunchTimer timer(ctx, SQW_KERNEL_RASTER_BIN);
			SQW_CUDA(ctx, launch_raster_bin(l, s));
			++launches;
		}
		if (f->num_cmds || f->clear)
		{
			LaunchTimer timer(ctx, SQW_KERNEL_RASTER_TILE);
			SQW_CUDA(ctx, launch_raster_tiles(l, s));
			++launches;
		}
	}

	if ((stages & SQW_STAGE_POST) && f->has_post)
	{
		const sqw_post_params& p = f->post;
		PostLaunch l{};
		l.hdr = static_cast<const uint64_t*>(view->d_color);
		l.width = view->width;
		l.height = view->height;
		l.dst = view->d_ldr;
		// TonemappingFunction::new (postprocessor.rs:1005-1017), ColorVec lanes [B, G, R]
		l.inv_scale_bgr[0] = 1.0f / (p.exposure * p.color_modulate[2]);
		l.inv_scale_bgr[1] = 1.0f / (p.exposure * p.color_modulate[1]);
		l.inv_scale_bgr[2] = 1.0f / (p.exposure * p.color_modulate[0]);
		l.color_sums = view->d_sums;
		SQW_CUDA(ctx, cudaMemsetAsync(view->d_sums, 0, 3 * sizeof(unsigned long long), s));
		if (p.bloom_sigma > 0.0f)
		{
			// perform_bloom (postprocessor.rs:729-789)
			const uint32_t scale_log2 = std::min<uint32_t>(std::max<uint32_t>(p.bloom_buffer_scale_log2, 1), 3);
			const uint32_t scale = 1u << scale_log2;
			l.bloom_scale_log2 = scale_log2;
			l.bloom_w = view->width / scale;
			l.bloom_h = view->height / scale;
			l.bloom0 = view->d_bloom[0];
			l.bloom1 = view->d_bloom[1];
			l.linear = p.linear_bloom_filter != 0;
			if (l.bloom_w == 0 || l.bloom_h == 0)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_submit: view too small for the bloom buffer scale");
			const float sigma_corrected = p.bloom_sigma / (float)scale;
			const int radius = (int)std::min<float>(std::fmax(std::ceil(3.0f * sigma_corrected - 0.5f), 0.0f), 16.0f);
			float kernel[33];
			compute_gaussian_kernel(sigma_corrected, radius, kernel);
			for (int i = 0; i < 33; ++i)
			{
				l.kernel_h[i] = (int32_t)(kernel[i] * 16384.0f);                 // :852
				l.kernel_v[i] = (int32_t)(kernel[i] * p.bloom_scale * 16384.0f); // :906
			}
			l.blur_radius = radius;
			{
				LaunchTimer timer(ctx, SQW_KERNEL_BLOOM_DOWNSCALE);
				SQW_CUDA(ctx, launch_bloom_downscale(l, s));
			}
			++launches;
			if (radius > 0)
			{
				LaunchTimer timer(ctx, SQW_KERNEL_BLOOM_BLUR);
				SQW_CUDA(ctx, launch_bloom_blur(l, s));
				launches += 2;
			}
			{
				LaunchTimer timer(ctx, SQW_KERNEL_TONEMAP_BLOOM);
				SQW_CUDA(ctx, launch_tonemap_bloom(l, s));
			}
			++launches;
			view->sums_divisor = (uint64_t)l.bloom_w * l.bloom_h;
		}
		else
		{
			{
				LaunchTimer timer(ctx, SQW_KERNEL_TONEMAP);
				SQW_CUDA(ctx, launch_tonemap(l, s));
			}
			++launches;
			view->sums_divisor = (uint64_t)view->width * view->height;
		}
	}
	f->stats.num_kernel_launches = launches;
	return SQW_OK;
}

int sqw_profile_begin(sqw_ctx* ctx)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (auto& t : ctx->timed)
	{
		ctx->event_pool.push_back(t.start);
		ctx->event_pool.push_back(t.stop);
	}
	ctx->timed.clear();
	ctx->profiling = true;
	return SQW_OK;
}

int sqw_profile_end(sqw_ctx* ctx, sqw_kernel_times* out)
{
	if (!ctx || !out)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	ctx->profiling = false;
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	std::memset(out, 0, sizeof(*out));
	for (auto& t : ctx->timed)
	{
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess && t.kind >= 0 && t.kind < SQW_NUM_KERNEL_KINDS)
		{
			out->ms[t.kind] += ms;
			out->launches[t.kind] += 1;
		}
		ctx->event_pool.push_back(t.start);
		ctx->event_pool.push_back(t.stop);
	}
	ctx->timed.clear();
	return SQW_OK;
}

int sqw_frame_get_stats(const sqw_frame* f, sqw_frame_stats* out)
{
	if (!f || !out)
		return SQW_ERR_INVALID_ARGUMENT;
	*out = f->stats;
	return SQW_OK;
}

static sqw_frame* immediate_frame(sqw_view* view)
{
	if (!view->immediate)
	{
		sqw_frame* f = nullptr;
		if (sqw_frame_create(view, &f) != SQW_OK)
			return nullptr;
		view->immediate = f;
	}
	return view->immediate;
}

int sqw_build_surfaces(
	sqw_view* view, const sqw_surface_job* jobs, uint32_t num_jobs, const sqw_light* lights, uint32_t num_lights, const float* shadow_maps,
	uint64_t shadow_maps_len, const float cam_pos[3], uint64_t total_surface_pixels)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_frame* f = immediate_frame(view);
	if (!f)
		return SQW_ERR_OUT_OF_MEMORY;
	int rc = sqw_frame_set_surfaces(f, jobs, num_jobs, lights, num_lights, shadow_maps, shadow_maps_len, cam_pos, total_surface_pixels);
	if (rc != SQW_OK)
		return rc;
	return sqw_frame_submit(f, SQW_STAGE_SURFACES);
}

int sqw_draw(
	sqw_view* view, const sqw_draw_cmd* cmds, uint32_t num_cmds, const sqw_polygon_point* polygon_points, uint32_t num_polygon_points,
	const sqw_triangle_point* triangle_points, uint32_t num_triangle_points, const void* clear_color)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_frame* f = immediate_frame(view);
	if (!f)
		return SQW_ERR_OUT_OF_MEMORY;
	int rc = sqw_frame_set_draw(f, cmds, num_cmds, polygon_points, num_polygon_points, triangle_points, num_triangle_points, clear_color);
	if (rc != SQW_OK)
		return rc;
	return sqw_frame_submit(f, SQW_STAGE_RASTER);
}

int sqw_view_get_average_color(sqw_view* view, float avg_color_out[3])
{
	if (!view || !avg_color_out)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	unsigned long long sums[3];
	SQW_CUDA(ctx, cudaMemcpyAsync(sums, view->d_sums, sizeof(sums), cudaMemcpyDeviceToHost, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const float inv = 1.0f / (float)view->sums_divisor; // postprocessor.rs:176-179 / :840-843
	for (int i = 0; i < 3; ++i)
		avg_color_out[i] = (float)sums[i] * inv;
	return SQW_OK;
}

int sqw_postprocess(sqw_view* view, const sqw_post_params* params, uint32_t* dst, size_t dst_pitch, float avg_color_out[3])
{
	if (!view || !params)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_frame* f = immediate_frame(view);
	if (!f)
		return SQW_ERR_OUT_OF_MEMORY;
	int rc = sqw_frame_set_post(f, params);
	if (rc != SQW_OK)
		return rc;
	rc = sqw_frame_submit(f, SQW_STAGE_POST);
	if (rc != SQW_OK)
		return rc;
	if (dst)
	{
		rc = sqw_read_pixels32(view, dst, dst_pitch);
		if (rc != SQW_OK)
			return rc;
	}
	if (avg_color_out)
		return sqw_view_get_average_color(view, avg_color_out);
	return SQW_OK;
}

int sqw_read_pixels(sqw_view* view, void* dst, size_t dst_pitch)
{
	if (!view || !dst || dst_pitch < view->width)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_read_pixels: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t C = (size_t)view->color_bits / 8;
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * C, view->d_color, (size_t)view->width * C, (size_t)view->width * C, view->height, cudaMemcpyDeviceToHost, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return SQW_OK;
}

int sqw_write_pixels(sqw_view* view, const void* src, size_t src_pitch)
{
	if (!view || !src || src_pitch < view->width)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_write_pixels: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t C = (size_t)view->color_bits / 8;
	SQW_CUDA(ctx, cudaMemcpy2DAsync(view->d_color, (size_t)view->width * C, src, src_pitch * C, (size_t)view->width * C, view->height, cudaMemcpyHostToDevice, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return SQW_OK;
}

int sqw_read_pixels32(sqw_view* view, uint32_t* dst, size_t dst_pitch)
{
	if (!view || !dst || dst_pitch < view->width || view->color_bits != 64)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_read_pixels32: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * 4, view->d_ldr, (size_t)view->width * 4, (size_t)view->width * 4, view->height, cudaMemcpyDeviceToHost, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return SQW_OK;
}

int sqw_read_surfaces(sqw_view* view, void* dst, uint64_t offset, uint64_t count)
{
	if (!view || !dst)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	const size_t C = (size_t)view->color_bits / 8;
	if ((offset + count) * C > view->surfaces.capacity)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_read_surfaces: range outside the pool");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(view->surfaces.ptr) + offset * C, count * C, cudaMemcpyDeviceToHost, ctx->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return SQW_OK;
}

int sqw_view_synchronize(sqw_view* view)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(view->ctx, cudaSetDevice(view->ctx->device));
	SQW_CUDA(view->ctx, cudaStreamSynchronize(view->ctx->stream));
	return SQW_OK;
}

} // extern "C"
