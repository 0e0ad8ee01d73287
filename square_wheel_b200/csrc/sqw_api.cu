// C-ABI implementation (include/sqw_cuda.h): contexts, device memory, host-side job/command preparation and the
// kernel launch sequence of one frame.  No CPU fallback: every entry fails with SQW_ERR_NO_DEVICE / SQW_ERR_CUDA
// when the CUDA device is unavailable.
#include "sqw_internal.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace sqw;

// NVTX ranges named after the reference's stage counters (RendererPerformanceCounters, renderer_structs.rs:7-15;
// PostprocessorPerformanceCounters): a timeline tool shows the stages under the names the reference's debug overlay uses.
// Header-only NVTX 3: costs nothing without a tool attached.
#include <nvtx3/nvToolsExt.h>
namespace
{
struct NvtxRange
{
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};
} // namespace

static_assert(sizeof(sqw_surface_job) == 136, "sqw_surface_job layout");
static_assert(sizeof(sqw_light) == 80, "sqw_light layout");
static_assert(sizeof(sqw_draw_cmd) == 64, "sqw_draw_cmd layout");
static_assert(sizeof(sqw_polygon_point) == 8, "sqw_polygon_point layout");
static_assert(sizeof(sqw_triangle_point) == 28, "sqw_triangle_point layout");
static_assert(sizeof(sqw_model_instance) == 296 && sizeof(sqw_model_vertex) == 32, "sqw_model_instance / sqw_model_vertex layout");
static_assert(sizeof(sqw_directional_lightmap_element) == 40, "sqw_directional_lightmap_element layout");
static_assert(sizeof(sqw_post_params) == 32, "sqw_post_params layout");
static_assert(sizeof(sqw_depth_map) == 32, "sqw_depth_map layout");
static_assert(sizeof(sqw_depth_cmd) == 32, "sqw_depth_cmd layout");

namespace
{

thread_local std::string g_create_error;

// Grow-only device buffer.
struct DevBuf
{
	void* ptr = nullptr;
	size_t capacity = 0;
	bool slice = false; // part of a larger allocation owned by someone else (sqw_view::scratch_base): never freed or grown here
	cudaError_t reserve(size_t bytes)
	{
		if (bytes <= capacity)
			return cudaSuccess;
		if (slice)
			return cudaErrorInvalidValue;
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
		if (ptr && !slice)
			cudaFree(ptr);
		ptr = nullptr;
		capacity = 0;
	}
};

// Host->device staging of one frame's lists: the caller's (pageable) arrays are packed into ONE page-locked buffer and sent
// with ONE asynchronous copy; sections are 256-byte aligned inside a single grow-only device allocation.
struct Arena
{
	char* h = nullptr;
	size_t h_capacity = 0;
	DevBuf d;
	size_t used = 0;
	cudaEvent_t copied = nullptr; // recorded after the copy: the staging buffer may be rewritten once it has fired
	bool in_flight = false;

	cudaError_t begin(size_t total_bytes)
	{
		if (in_flight)
		{
			cudaError_t e = cudaEventSynchronize(copied);
			if (e != cudaSuccess)
				return e;
			in_flight = false;
		}
		used = 0;
		if (total_bytes > h_capacity)
		{
			if (h)
				cudaFreeHost(h);
			h = nullptr;
			h_capacity = 0;
			// page-locked allocations are slow and synchronise the device: grow geometrically from a floor of 1 MiB
			const size_t rounded = std::max<size_t>((2 * total_bytes + 4095) & ~size_t(4095), size_t(1) << 20);
			cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&h), rounded, cudaHostAllocDefault);
			if (e != cudaSuccess)
			{
				h = nullptr;
				return e;
			}
			h_capacity = rounded;
		}
		return d.reserve(std::max<size_t>(total_bytes, h_capacity / 2));
	}
	static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
	// returns the device address the section will have after commit()
	void* push(const void* src, size_t bytes)
	{
		void* dev = static_cast<char*>(d.ptr) + used;
		if (bytes)
			std::memcpy(h + used, src, bytes);
		used += padded(bytes);
		return dev;
	}
	cudaError_t commit(cudaStream_t stream)
	{
		if (used == 0)
			return cudaSuccess;
		if (!copied)
		{
			cudaError_t e = cudaEventCreateWithFlags(&copied, cudaEventDisableTiming);
			if (e != cudaSuccess)
				return e;
		}
		cudaError_t e = cudaMemcpyAsync(d.ptr, h, used, cudaMemcpyHostToDevice, stream);
		if (e != cudaSuccess)
			return e;
		in_flight = true;
		return cudaEventRecord(copied, stream);
	}
	void release()
	{
		if (copied)
		{
			cudaEventSynchronize(copied);
			cudaEventDestroy(copied);
		}
		copied = nullptr;
		if (h)
			cudaFreeHost(h);
		h = nullptr;
		h_capacity = 0;
		d.release();
		in_flight = false;
	}
};

struct ImageHost
{
	void* pixels;
	uint32_t width, height;
	int kind;
};

} // namespace

// ---- compressed read-back: decode jobs on a small pool of host threads (readback_decode.cu) ---------------------------------
struct ReadbackJob // one per view: a picture travelling to the caller's buffer
{
	const uint32_t* rle = nullptr;    // page-locked, device-mapped: row r at r * row_capacity words
	const uint32_t* counts = nullptr; // run words per row
	uint32_t* dst = nullptr;
	size_t dst_pitch = 0;
	uint32_t width = 0, height = 0, row_capacity = 0;
	std::atomic<uint32_t> next_row{0}, rows_done{0};
	std::atomic<uint64_t> words{0}; // run / raw words the rows held: what crossed the bus
	int workers_inside = 0; // guarded by the pool mutex
	bool finished = true;   // guarded by the pool mutex: every row decoded and no worker still holds the job
};
struct ReadbackPool
{
	std::vector<std::thread> threads;
	std::deque<ReadbackJob*> queue;
	std::mutex m;
	std::condition_variable work, done;
	bool stop = false;
	static constexpr uint32_t ROWS_PER_CHUNK = 16;

	void start(unsigned n)
	{
		for (unsigned i = 0; i < n; ++i)
			threads.emplace_back([this] { run(); });
	}
	void run()
	{
		std::unique_lock<std::mutex> lock(m);
		for (;;)
		{
			work.wait(lock, [this] { return stop || !queue.empty(); });
			if (stop)
				return;
			ReadbackJob* job = queue.front();
			++job->workers_inside;
			lock.unlock();
			for (;;)
			{
				const uint32_t r0 = job->next_row.fetch_add(ROWS_PER_CHUNK);
				if (r0 >= job->height)
					break;
				const uint32_t r1 = std::min(r0 + ROWS_PER_CHUNK, job->height);
				sqw::decode_picture_rows(job->rle, job->counts, job->row_capacity, job->width, r0, r1, job->dst, job->dst_pitch);
				uint64_t words = 0;
				for (uint32_t r = r0; r < r1; ++r)
					words += job->counts[r] & 0x7FFFFFFFu;
				job->words.fetch_add(words + (r1 - r0));
				job->rows_done.fetch_add(r1 - r0);
			}
			lock.lock();
			if (!queue.empty() && queue.front() == job)
				queue.pop_front(); // every row has an owner: nobody else needs to pick this job up
			--job->workers_inside;
			if (job->workers_inside == 0 && job->rows_done.load() == job->height)
			{
				job->finished = true;
				done.notify_all();
			}
		}
	}
	void push(ReadbackJob* job) // from the stream's host callback: the GPU has written the rows
	{
		{
			std::lock_guard<std::mutex> lock(m);
			queue.push_back(job);
		}
		work.notify_all();
	}
	void wait(ReadbackJob* job)
	{
		std::unique_lock<std::mutex> lock(m);
		done.wait(lock, [job] { return job->finished; });
	}
	void shutdown()
	{
		{
			std::lock_guard<std::mutex> lock(m);
			stop = true;
		}
		work.notify_all();
		for (auto& t : threads)
			t.join();
		threads.clear();
	}
};

struct sqw_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	bool concurrent_views = false; // SQW_CTX_CONCURRENT_VIEWS: every view issues its work on a private stream
	bool fast_surface_numerics = false; // SQW_CTX_FAST_SURFACE_NUMERICS
	bool compressed_readback = false;   // SQW_CTX_COMPRESSED_READBACK
	ReadbackPool* readback_pool = nullptr; // started by the first compressed read-back
	size_t l2_persist_bytes = 0;   // > 0: the views' scratch blocks are marked L2-persisting on their streams (sqw_create)
	cudaEvent_t batch_event = nullptr;
	float* d_lightmaps = nullptr;
	sqw_directional_lightmap_element* d_directional_lightmaps = nullptr;
	uint64_t num_lightmap_elements = 0;
	std::vector<ImageHost> images;
	ImageDev* d_images = nullptr;
	uint32_t max_images = 0;
	std::vector<sqw_view*> views;
	// triangle-model meshes (sqw_upload_model_mesh): host-side descriptors with DEVICE pointers + the device table
	std::vector<sqw::ModelMeshDev> meshes;
	std::vector<void*> mesh_allocations;
	sqw::ModelMeshDev* d_meshes = nullptr;
	size_t d_meshes_capacity = 0;
	// map geometry for sqw_polygons_prepare (sqw_upload_map_geometry)
	sqw_map_polygon* d_map_polygons = nullptr;
	float* d_map_vertices = nullptr;
	uint8_t* d_material_flags = nullptr;
	sqw_map_material* d_map_materials = nullptr; // sqw_upload_map_materials
	float* d_polygon_basis = nullptr;
	uint32_t num_map_polygons = 0, num_map_vertices = 0, num_materials = 0;
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
	cudaStream_t stream;
	LaunchTimer(sqw_ctx* c, int kind, cudaStream_t s) : ctx(c), stream(s)
	{
		if (!ctx->profiling)
			return;
		t.kind = kind;
		t.start = ctx->get_event();
		t.stop = ctx->get_event();
		cudaEventRecord(t.start, stream);
	}
	~LaunchTimer()
	{
		if (!ctx->profiling)
			return;
		cudaEventRecord(t.stop, stream);
		ctx->timed.push_back(t);
	}
};
} // namespace

struct sqw_view
{
	sqw_ctx* ctx = nullptr;
	uint32_t index = 0;
	uint32_t width = 0, height = 0;             // logical size (sqw_view_set_size)
	uint32_t alloc_width = 0, alloc_height = 0; // size the buffers were allocated for
	int color_bits = 0;
	cudaStream_t stream = nullptr; // ctx->stream, or a private one (SQW_CTX_CONCURRENT_VIEWS)
	bool own_stream = false;
	cudaEvent_t done = nullptr; // recorded after every submit: other views that sample this one wait on it
	void* d_color = nullptr;
	uint32_t* d_ldr = nullptr; // post-processed picture (64-bit views)
	// Per-frame scratch of the view — surface pool, span table, band / tile lists, triangle records — lives in ONE allocation
	// (scratch_base, sections in this order) so that a single L2 access-policy window on the view's stream covers everything a
	// frame's kernels hand to each other; the four DevBufs below are non-owning slices of it (view_scratch_reserve).
	char* scratch_base = nullptr;
	size_t scratch_total = 0;
	DevBuf surfaces;
	DevBuf spans;
	DevBuf band_lists;
	// span set-up + binning of a frame run beside its surface building (they are independent; both feed the tile kernel)
	cudaStream_t aux_stream = nullptr;
	cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
	DevBuf tri_recs; // per-triangle set-up records of the draw being submitted (SQW_TRI_REC_BYTES per command)
	uint64_t* d_bloom[2] = {nullptr, nullptr};
	uint8_t* d_tonemap_lut = nullptr;     // 3 x SQW_TONEMAP_LUT_SIZE (64-bit views)
	unsigned long long* d_sums = nullptr; // SQW_SUM_SLOTS x {B, G, R, pad}
	uint64_t sums_divisor = 1;
	sqw_frame* immediate = nullptr;
	// results and scratch of sqw_models_prepare
	DevBuf model_instances, model_vertices, model_triangles, model_counts, model_bones, model_sort;
	uint32_t model_num_instances = 0;
	uint64_t model_num_vertices = 0, model_num_triangles = 0;
	// compressed read-back (sqw_read_pixels32_async with SQW_CTX_COMPRESSED_READBACK)
	uint32_t* h_rle = nullptr; // page-locked, device-mapped: height rows of row_capacity words, then height count words
	size_t h_rle_words = 0;
	ReadbackJob readback_job;
	bool readback_pending = false;
	uint64_t readback_bytes_total = 0, readback_pictures = 0; // completed compressed read-backs of this view
	// inputs and results of sqw_polygons_prepare
	DevBuf polygon_candidates, polygon_clip_planes, polygon_results, polygon_totals;
	uint32_t polygon_num_candidates = 0, polygon_num_clip_sets = 0;
	sqw_polygons_frame polygons_frame{};
	// scratch and results of sqw_polygons_draw
	DevBuf polygon_order, polygon_draw_scratch_cmds, polygon_draw_scratch_points, polygon_draw_offsets, polygon_draw_cmds, polygon_draw_points, polygon_draw_totals;
	uint32_t polygon_num_order = 0;
	bool polygon_draw_valid = false; // the draw lists belong to the records of the last sqw_polygons_prepare
};

struct SurfaceGroup // one kernel launch: all bands without / with dynamic lights, CTA bands first, then warp bands
{
	bool dynamic_lights;
	uint32_t num_cta_bands;
	uint32_t band_offset, num_bands;
};

struct sqw_frame
{
	sqw_view* view = nullptr;
	// surfaces
	Arena surf_arena;
	const sqw_surface_job* d_jobs = nullptr;
	const SurfaceBand* d_bands = nullptr;
	const sqw_light* d_lights = nullptr;
	const float* d_shadow_maps = nullptr;
	std::vector<SurfaceBand> h_bands;
	std::vector<SurfaceGroup> groups;
	float cam_pos[3] = {0, 0, 0};
	bool has_surfaces = false;
	DevBuf d_shadow_dev; // device-only shadow map array (sqw_frame_set_surfaces with shadow_maps == NULL)
	float* d_shadow_maps_rw = nullptr;
	uint64_t shadow_maps_len = 0;
	// shadow map fills
	Arena depth_arena;
	DevBuf d_depth_spans;
	const sqw_depth_cmd* d_depth_cmds = nullptr;
	const DepthCmdDev* d_depth_cmd_dev = nullptr;
	const sqw_polygon_point* d_depth_points = nullptr;
	const sqw_depth_map* d_depth_maps = nullptr;
	const uint2* d_depth_chunks = nullptr;
	const uint2* d_depth_row_chunks = nullptr;
	uint32_t num_depth_chunks = 0, num_depth_row_chunks = 0;
	uint64_t depth_max_end = 0;
	bool has_depth = false;
	std::vector<DepthCmdDev> h_depth_cmd_dev;
	std::vector<uint2> h_depth_chunks, h_depth_row_chunks;
	// draw
	Arena draw_arena;
	const sqw_draw_cmd* d_cmds = nullptr;
	const CmdDev* d_cmd_dev = nullptr;
	const sqw_polygon_point* d_polygon_points = nullptr;
	const sqw_triangle_point* d_triangle_points = nullptr;
	const uint2* d_chunks = nullptr;
	const uint32_t* d_band_offsets = nullptr;
	const uint32_t* d_band_list = nullptr; // ordered command indices per band, band b at [band_offsets[b], band_offsets[b + 1])
	std::vector<uint32_t> h_band_list, h_band_cursor;
	std::vector<uint32_t> sampled_views; // views referenced through SQW_TEX_VIEW
	std::vector<CmdDev> h_cmd_dev;
	std::vector<uint2> h_chunks, h_chunks_tri;
	std::vector<uint32_t> h_band_offsets;
	uint32_t num_cmds = 0, total_spans = 0, num_bands = 0, total_band_entries = 0, num_triangle_cmds = 0, num_general_cmds = 0, num_chunks = 0;
	// sqw_frame_set_draw_from_polygons: the lists live in buffers of the frame, built on the device
	DevBuf dev_cmds, dev_points, dev_cmd_dev, dev_rows, dev_chunk_base, dev_chunks, dev_band_counts, dev_band_offsets, dev_band_list, dev_totals;
	// sqw_frame_set_surfaces_from_polygons
	DevBuf dev_jobs, dev_visible_rank, dev_surf_band_counts, dev_surf_band_base, dev_surf_bands, dev_surf_totals, dev_lights, dev_set_light_offsets, dev_set_lights;
	uint32_t draw_width = 0, draw_height = 0; // view size the command list was prepared for
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
cudaError_t synchronize_all(sqw_ctx* ctx)
{
	cudaError_t e = cudaStreamSynchronize(ctx->stream);
	for (sqw_view* v : ctx->views)
		if (e == cudaSuccess && v && v->stream != ctx->stream)
			e = cudaStreamSynchronize(v->stream);
	return e;
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

// Grow one section of the view's scratch block (0 surfaces, 1 spans, 2 band / tile lists, 3 triangle records).  The caller has
// synchronised the view's stream (a queued frame may still use the old block).  The other sections keep their contents (the
// surface pool of sqw_build_surfaces must survive the sqw_draw that follows).  Re-points the access-policy window of the view's
// own stream at the whole block: the surface kernel's stores, the set-up grid's span records and tile lists all stay in the L2
// carve-out (sqw_create) until the tile kernel has consumed them, whatever the other views in flight stream through L2 meanwhile.
cudaError_t view_scratch_reserve(sqw_view* view, int section, size_t bytes)
{
	DevBuf* const parts[4] = {&view->surfaces, &view->spans, &view->band_lists, &view->tri_recs};
	if (bytes <= parts[section]->capacity)
		return cudaSuccess;
	size_t cap[4], total = 0;
	for (int i = 0; i < 4; ++i)
	{
		cap[i] = parts[i]->capacity;
		if (i == section)
			cap[i] = (bytes + (bytes >> 2) + 255) & ~size_t(255);
		total += cap[i];
	}
	char* block = nullptr;
	cudaError_t e = cudaMalloc(&block, total);
	if (e != cudaSuccess)
		return e;
	size_t off = 0;
	for (int i = 0; i < 4 && e == cudaSuccess; ++i)
	{
		if (i != section && parts[i]->capacity)
			e = cudaMemcpy(block + off, parts[i]->ptr, parts[i]->capacity, cudaMemcpyDeviceToDevice);
		off += cap[i];
	}
	if (e != cudaSuccess)
	{
		cudaFree(block);
		return e;
	}
	if (view->scratch_base)
		cudaFree(view->scratch_base);
	view->scratch_base = block;
	view->scratch_total = total;
	off = 0;
	for (int i = 0; i < 4; ++i)
	{
		parts[i]->ptr = cap[i] ? block + off : nullptr;
		parts[i]->capacity = cap[i];
		parts[i]->slice = true;
		off += cap[i];
	}
	sqw_ctx* ctx = view->ctx;
	if (ctx->l2_persist_bytes && (view->own_stream || ctx->own_stream)) // never touch the attributes of a caller's stream
	{
		int max_window = 0;
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
		cudaStreamAttrValue attr{};
		attr.accessPolicyWindow.base_ptr = block;
		attr.accessPolicyWindow.num_bytes = std::min<size_t>(total, (size_t)std::max(max_window, 0));
		attr.accessPolicyWindow.hitRatio = 1.0f;
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
		e = cudaStreamSetAttribute(view->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
	}
	return e;
}

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

uint32_t variant_bits(const SurfaceVariant& v) // SurfaceBand::variant
{
	if (v.emissive_only)
		return 16;
	return (uint32_t)v.directional | ((uint32_t)v.normal_map << 1) | ((uint32_t)v.specular << 2);
}

} // namespace

extern "C"
{

// The decode job of the view's last compressed read-back, if any: call after the view's stream has been synchronised (the
// stream's host callback has queued it by then).
static void finish_readback(sqw_view* view)
{
	if (view->readback_pending)
	{
		view->ctx->readback_pool->wait(&view->readback_job);
		view->readback_pending = false;
		view->readback_bytes_total += view->readback_job.words.load() * 4;
		++view->readback_pictures;
	}
}

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
	ctx->concurrent_views = (map->flags & SQW_CTX_CONCURRENT_VIEWS) != 0;
	ctx->fast_surface_numerics = (map->flags & SQW_CTX_FAST_SURFACE_NUMERICS) != 0;
	ctx->compressed_readback = (map->flags & SQW_CTX_COMPRESSED_READBACK) != 0;
	{
		// A view's scratch — surface pool, span table, tile lists — is written by one kernel of a frame and read by the next a moment
		// later, while the other views in flight stream ~100 MB of pixels each through the 126 MB L2.  Marking the views' scratch
		// blocks L2-persisting on their streams (view_scratch_reserve) keeps them on chip: 17.7k -> 19.0k frames/s at four 4K views
		// in flight with a 48 MB carve-out (about the four views' working sets; 40-56 MB measured alike, 64 MB and more slow the pixel
		// stores down).  SQW_L2_PERSIST_MB overrides the size, 0 turns it off.  The limit is device-wide.
		const char* mb = std::getenv("SQW_L2_PERSIST_MB");
		int max_persist = 0;
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
		const size_t want = std::min<size_t>((size_t)(mb ? std::atol(mb) : 48) << 20, (size_t)std::max(max_persist, 0));
		if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
			ctx->l2_persist_bytes = want;
		else
			cudaGetLastError(); // not supported here: carry on without it
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
	if (ctx->readback_pool)
	{
		ctx->readback_pool->shutdown();
		delete ctx->readback_pool;
	}
	if (ctx->l2_persist_bytes)
		cudaCtxResetPersistingL2Cache();
	for (auto& img : ctx->images)
		if (img.pixels)
			cudaFree(img.pixels);
	if (ctx->d_images)
		cudaFree(ctx->d_images);
	for (void* p : ctx->mesh_allocations)
		cudaFree(p);
	if (ctx->d_meshes)
		cudaFree(ctx->d_meshes);
	for (void* p : {static_cast<void*>(ctx->d_map_polygons), static_cast<void*>(ctx->d_map_vertices), static_cast<void*>(ctx->d_material_flags), static_cast<void*>(ctx->d_map_materials),
					static_cast<void*>(ctx->d_polygon_basis)})
		if (p)
			cudaFree(p);
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
	if (ctx->batch_event)
		cudaEventDestroy(ctx->batch_event);
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
	SQW_CUDA(ctx, synchronize_all(ctx)); // frames in flight on any view may still sample the old pixels
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
	v->width = v->alloc_width = width;
	v->height = v->alloc_height = height;
	v->color_bits = color_bits;
	const size_t px = (size_t)width * height;
	cudaError_t e = cudaSuccess;
	v->stream = ctx->stream;
	if (ctx->concurrent_views)
	{
		e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
		v->own_stream = e == cudaSuccess;
		if (e != cudaSuccess)
			v->stream = ctx->stream;
	}
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&v->done, cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&v->aux_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&v->aux_fork, cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&v->aux_join, cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaMalloc(&v->d_color, px * (color_bits / 8));
	if (e == cudaSuccess)
		e = cudaMemset(v->d_color, 0, px * (color_bits / 8));
	if (e == cudaSuccess && color_bits == 64)
		e = cudaMalloc(&v->d_ldr, px * 4);
	if (e == cudaSuccess && color_bits == 64)
	{
		// bloom buffers at the largest size (scale 2) of ANY shape sqw_view_set_size accepts: it bounds the area, and
		// (w / 2) * (h / 2) <= w * h / 4 for every shape of that area (1921 x 1081 -> 2076 x 1000 needs more than (1921 / 2) * (1081 / 2))
		const size_t bpx = ((size_t)width * height + 3) / 4 + 1;
		e = cudaMalloc(&v->d_bloom[0], bpx * 8);
		if (e == cudaSuccess)
			e = cudaMalloc(&v->d_bloom[1], bpx * 8);
	}
	if (e == cudaSuccess && color_bits == 64)
		e = cudaMalloc(&v->d_tonemap_lut, 3 * SQW_TONEMAP_LUT_SIZE);
	if (e == cudaSuccess)
		e = cudaMalloc(&v->d_sums, SQW_SUM_WORDS * sizeof(unsigned long long));
	if (e == cudaSuccess)
		e = cudaMemset(v->d_sums, 0, SQW_SUM_WORDS * sizeof(unsigned long long));
	if (e != cudaSuccess)
	{
		int rc = fail_cuda(ctx, e, "sqw_view_create: device allocation");
		v->index = (uint32_t)ctx->views.size();
		ctx->views.push_back(v);
		sqw_view_destroy(v);
		return rc;
	}
	cudaDeviceSynchronize(); // the clears above ran on the legacy stream; private streams do not order against it
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
	cudaStreamSynchronize(view->stream);
	cudaStreamSynchronize(ctx->stream);
	finish_readback(view);
	if (view->h_rle)
		cudaFreeHost(view->h_rle);
	if (view->immediate)
		sqw_frame_destroy(view->immediate);
	if (view->done)
		cudaEventDestroy(view->done);
	if (view->aux_stream)
	{
		cudaStreamSynchronize(view->aux_stream);
		cudaStreamDestroy(view->aux_stream);
	}
	if (view->aux_fork)
		cudaEventDestroy(view->aux_fork);
	if (view->aux_join)
		cudaEventDestroy(view->aux_join);
	if (view->own_stream)
		cudaStreamDestroy(view->stream);
	if (view->d_color)
		cudaFree(view->d_color);
	if (view->d_ldr)
		cudaFree(view->d_ldr);
	if (view->scratch_base)
		cudaFree(view->scratch_base); // the four sections are slices of it
	for (auto& b : view->d_bloom)
		if (b)
			cudaFree(b);
	if (view->d_sums)
		cudaFree(view->d_sums);
	if (view->d_tonemap_lut)
		cudaFree(view->d_tonemap_lut);
	for (DevBuf* b : {&view->model_instances, &view->model_vertices, &view->model_triangles, &view->model_counts, &view->model_bones, &view->model_sort,
					  &view->polygon_candidates, &view->polygon_clip_planes, &view->polygon_results, &view->polygon_totals, &view->polygon_order,
					  &view->polygon_draw_scratch_cmds, &view->polygon_draw_scratch_points, &view->polygon_draw_offsets, &view->polygon_draw_cmds, &view->polygon_draw_points,
					  &view->polygon_draw_totals})
		b->release();
	// keep indices of the other views stable: leave a hole
	for (auto& p : ctx->views)
		if (p == view)
			p = nullptr;
	while (!ctx->views.empty() && ctx->views.back() == nullptr)
		ctx->views.pop_back();
	delete view;
}

int sqw_view_set_size(sqw_view* view, uint32_t width, uint32_t height)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	if (width == 0 || height == 0 || (uint64_t)width * height > (uint64_t)view->alloc_width * view->alloc_height)
		return fail(view->ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_view_set_size: larger than the size the view was created with");
	// frames already queued for this view keep the geometry they were submitted with (kernel arguments are by value)
	view->width = width;
	view->height = height;
	return SQW_OK;
}

int sqw_view_index(const sqw_view* view) { return view ? (int)view->index : SQW_ERR_INVALID_ARGUMENT; }

int sqw_view_mix_texture(sqw_view* view, uint32_t image_id, const int32_t surface_tc_min[2], int blending_mode, const float light[3])
{
	if (!view || !surface_tc_min || !light)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	if (image_id >= ctx->images.size() || ctx->images[image_id].kind != SQW_IMAGE_COLOR32 || blending_mode < 0 || blending_mode > SQW_BLEND_ALPHA_BLEND)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_view_mix_texture: bad overlay image or blending mode");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const ImageHost& img = ctx->images[image_id];
	{
		LaunchTimer timer(ctx, SQW_KERNEL_MIX_TEXTURE, view->stream);
		SQW_CUDA(ctx, launch_mix_texture(view->color_bits == 64, view->d_color, view->width, view->height, surface_tc_min, img.pixels, img.width, img.height, blending_mode, light, view->stream));
	}
	if (ctx->concurrent_views)
		SQW_CUDA(ctx, cudaEventRecord(view->done, view->stream));
	return SQW_OK;
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
	cudaStreamSynchronize(f->view->stream);
	f->surf_arena.release();
	f->draw_arena.release();
	f->depth_arena.release();
	f->d_depth_spans.release();
	f->d_shadow_dev.release();
	for (DevBuf* b : {&f->dev_cmds, &f->dev_points, &f->dev_cmd_dev, &f->dev_rows, &f->dev_chunk_base, &f->dev_chunks, &f->dev_band_counts, &f->dev_band_offsets, &f->dev_band_list, &f->dev_totals,
					  &f->dev_jobs, &f->dev_visible_rank, &f->dev_surf_band_counts, &f->dev_surf_band_base, &f->dev_surf_bands, &f->dev_surf_totals, &f->dev_lights,
					  &f->dev_set_light_offsets, &f->dev_set_lights})
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

	// Every shadow map a light points at must lie inside the frame's shadow map array: the kernel fetches
	// sm[offset + side * size^2 + u + v * size] (light.rs:196-294) without a bound of its own.
	for (uint32_t l = 0; l < num_lights; ++l)
	{
		const sqw_light& li = lights[l];
		if (li.shadow_type == SQW_SHADOW_NONE)
			continue;
		if (li.shadow_type != SQW_SHADOW_CUBE && li.shadow_type != SQW_SHADOW_PROJECTOR)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: unknown shadow type");
		const uint64_t sides = li.shadow_type == SQW_SHADOW_CUBE ? 6 : 1;
		const uint64_t end = (uint64_t)li.shadow_map_offset + sides * (uint64_t)li.shadow_map_size * li.shadow_map_size;
		if (li.shadow_map_size == 0 || end > shadow_maps_len)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces: a light's shadow map lies outside the shadow map array");
	}

	// Warp bands (small surfaces) are sized so that they fill the machine about once: a frame with few small texels (a nested
	// portal view, the usual map at 4K) is latency-bound and wants short bands (two texels per lane like the CTA band of a small
	// surface), a frame with more than a million of them (the stress map) is throughput-bound and wants the prologue amortised.
	uint32_t warp_band_texels = SQW_SURF_WARP_BAND_TEXELS;
	{
		uint64_t small_texels = 0;
		for (uint32_t i = 0; i < num_jobs; ++i)
		{
			const uint64_t texels = (uint64_t)jobs[i].surface_size[0] * jobs[i].surface_size[1];
			if (texels <= SQW_SURF_WARP_JOB_TEXELS)
				small_texels += texels;
		}
		const uint64_t resident_warps = 148ull * 4 * (SQW_SURF_THREADS / 32);
		while (warp_band_texels > SQW_SURF_WARP_BAND_TEXELS_MIN && small_texels < resident_warps * warp_band_texels)
			warp_band_texels /= 2;
	}
	// Validate and group jobs by kernel variant (the reference's const-generic dispatch, surfaces.rs:124-423).
	std::vector<SurfaceBand> buckets[4]; // [0] no dynamic lights, [1] dynamic lights; [2], [3] the same for warp bands (small surfaces)
	uint32_t variants_seen = 0;
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
		if (texels == 0)
			continue;
		const uint32_t vbits = variant_bits(v);
		variants_seen |= 1u << (vbits + (v.dynamic_lights ? 12 : 0));
		const bool warp_bands = texels <= SQW_SURF_WARP_JOB_TEXELS;
		uint32_t band_rows;
		if (warp_bands)
		{
			// rows of a warp band share a pool of line-lightmap elements (the warp-band CTAs of surface_build_kernel)
			band_rows = std::max<uint32_t>(warp_band_texels / j.surface_size[0], 1);
			if (diffuse && j.lightmap_offset != 0)
				band_rows = std::min<uint32_t>(band_rows, std::max<uint32_t>(SQW_SURF_WARP_LM_ELEMENTS / std::min<uint32_t>(j.lightmap_size[0], 19), 1));
			band_rows = std::min<uint32_t>(band_rows, SQW_SURF_MAX_BAND_ROWS);
		}
		else
			band_rows = std::min<uint32_t>(std::max<uint32_t>(SQW_SURF_BAND_TEXELS / j.surface_size[0], 1), SQW_SURF_MAX_BAND_ROWS);
		for (uint32_t row0 = 0; row0 < j.surface_size[1]; row0 += band_rows)
			buckets[(v.dynamic_lights ? 1 : 0) + (warp_bands ? 2 : 0)].push_back(SurfaceBand{i, row0, std::min(band_rows, j.surface_size[1] - row0), vbits});
		st.surface_texels += texels;
		if (emissive)
			st.emissive_texels += texels;
	}
	f->h_bands.clear();
	f->groups.clear();
	for (uint32_t key = 0; key < 2; ++key)
	{
		if (buckets[key].empty() && buckets[key + 2].empty())
			continue;
		SurfaceGroup g{};
		g.dynamic_lights = key == 1;
		g.band_offset = (uint32_t)f->h_bands.size();
		for (uint32_t kind = 0; kind < 2; ++kind) // CTA bands, then warp bands
		{
			std::vector<SurfaceBand>& b = buckets[key + 2 * kind];
			// heaviest variants first so the tail of the grid is made of cheap bands
			std::stable_sort(b.begin(), b.end(), [](const SurfaceBand& x, const SurfaceBand& y) { return (x.variant & 15u) > (y.variant & 15u); });
			f->h_bands.insert(f->h_bands.end(), b.begin(), b.end());
			if (kind == 0)
				g.num_cta_bands = (uint32_t)b.size();
		}
		g.num_bands = (uint32_t)f->h_bands.size() - g.band_offset;
		f->groups.push_back(g);
	}
	st.num_surface_variants = (uint32_t)__builtin_popcount(variants_seen);

	if ((size_t)total_surface_pixels * C + 16 > view->surfaces.capacity)
	{
		// growing frees the old pool: nothing queued on this view may still read it
		SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
		SQW_CUDA(ctx, view_scratch_reserve(view, 0, (size_t)total_surface_pixels * C + 16));
	}
	{
		Arena& a = f->surf_arena;
		const size_t b_jobs = (size_t)num_jobs * sizeof(sqw_surface_job), b_bands = f->h_bands.size() * sizeof(SurfaceBand);
		const size_t b_lights = (size_t)num_lights * sizeof(sqw_light), b_sm = (size_t)shadow_maps_len * sizeof(float);
		const size_t b_sm_upload = shadow_maps ? b_sm : 0; // NULL: the array lives on the device only (filled by the shadow map stage)
		if (b_sm && !shadow_maps && b_sm + 16 > f->d_shadow_dev.capacity)
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
			SQW_CUDA(ctx, f->d_shadow_dev.reserve(b_sm + 16));
		}
		SQW_CUDA(ctx, a.begin(Arena::padded(b_jobs) + Arena::padded(b_bands) + Arena::padded(b_lights) + Arena::padded(b_sm_upload)));
		f->d_jobs = static_cast<const sqw_surface_job*>(a.push(jobs, b_jobs));
		f->d_bands = static_cast<const SurfaceBand*>(a.push(f->h_bands.data(), b_bands));
		f->d_lights = static_cast<const sqw_light*>(a.push(lights, b_lights));
		f->d_shadow_maps_rw = shadow_maps ? static_cast<float*>(a.push(shadow_maps, b_sm)) : static_cast<float*>(f->d_shadow_dev.ptr);
		f->d_shadow_maps = f->d_shadow_maps_rw;
		f->shadow_maps_len = shadow_maps_len;
		SQW_CUDA(ctx, a.commit(view->stream));
	}
	f->cam_pos[0] = cam_pos[0];
	f->cam_pos[1] = cam_pos[1];
	f->cam_pos[2] = cam_pos[2];
	f->has_surfaces = true;
	return SQW_OK;
}

// Layout of sqw_view::band_lists (scratch of the binning CTAs of one draw): [band lists][tile lists][tile descriptors][general tiles].
static size_t bin_tile_lists_offset(uint32_t band_entries)
{
	return ((size_t)band_entries * 4 + 63) & ~(size_t)63;
}
static size_t bin_tile_desc_offset(uint32_t band_entries, uint32_t tiles_x)
{
	return bin_tile_lists_offset(band_entries) + (((size_t)band_entries * tiles_x * 4 + 63) & ~(size_t)63);
}
static size_t bin_general_tiles_offset(uint32_t band_entries, uint32_t tiles_x, uint32_t num_bands)
{
	return bin_tile_desc_offset(band_entries, tiles_x) + (size_t)num_bands * tiles_x * sizeof(uint4);
}
static size_t bin_scratch_bytes(uint32_t band_entries, uint32_t tiles_x, uint32_t num_bands)
{
	// ... [general tiles: count + one entry per tile at most]
	return bin_general_tiles_offset(band_entries, tiles_x, num_bands) + ((size_t)num_bands * tiles_x + SQW_GENERAL_LIST_HEADER) * sizeof(uint32_t) + 64;
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
	f->sampled_views.clear();
	f->h_chunks.clear();
	f->h_chunks_tri.clear();
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
			// Smallest inv_z of the polygon (it is linear, so a vertex has it), lowered by two pixels of gradient for the pixel-centre
			// offsets and the rounding of the span ends, in the rasteriser's 2^29 scale (rasterizer.rs:287-310): when even half of it
			// stays above 2^11 << 8, every per-pixel divisor inv_z >> 8 of the full-perspective loop is >= 2^11 and the opaque tile
			// kernel may use its one-step division.
			uint8_t z_flag = 0;
			if (c.tc_mode == SQW_TC_FULL_PERSPECTIVE)
			{
				double min_iz = 1e300;
				for (uint32_t k = 0; k < c.num_vertices; ++k)
				{
					const sqw_polygon_point& p = polygon_points[c.vertex_offset + k];
					min_iz = std::min(min_iz, (double)c.depth_eq[0] * (p.x / 65536.0) + (double)c.depth_eq[1] * (p.y / 65536.0) + (double)c.depth_eq[2]);
				}
				min_iz -= 2.0 * (std::fabs((double)c.depth_eq[0]) + std::fabs((double)c.depth_eq[1]));
				if (min_iz * 536870912.0 * 0.5 >= 2048.0 * 256.0)
					z_flag = SQW_CMDF_Z_ONE_STEP;
			}
			// polygon textures have the framebuffer's colour type (rasterizer.rs:41)
			if (c.tex_kind == SQW_TEX_SURFACE)
			{
				if ((uint64_t)c.tex_ref + tex_elems > view->surfaces.capacity / C)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: surface reference outside the pool (build surfaces first)");
				d.texture = reinterpret_cast<const void*>((uintptr_t)c.tex_ref * C);
				d.tex_flags = (view->color_bits == 64 ? 1 : 0) | 2 | z_flag;
			}
			else if (c.tex_kind == SQW_TEX_IMAGE)
			{
				if (c.tex_ref >= ctx->images.size())
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad image reference");
				const ImageHost& img = ctx->images[c.tex_ref];
				if (img.kind != (view->color_bits == 64 ? SQW_IMAGE_COLOR64 : SQW_IMAGE_COLOR32) || tex_elems > (uint64_t)img.width * img.height)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: polygon image must match the view colour type and size");
				d.texture = img.pixels;
				d.tex_flags = (view->color_bits == 64 ? 1 : 0) | z_flag;
			}
			else if (c.tex_kind == SQW_TEX_VIEW)
			{
				if (c.tex_ref >= ctx->views.size() || !ctx->views[c.tex_ref] || ctx->views[c.tex_ref] == view)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: bad view reference");
				const sqw_view* tv = ctx->views[c.tex_ref];
				if (tv->color_bits != view->color_bits || tex_elems > (uint64_t)tv->width * tv->height)
					return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: view texture must match colour type and size");
				d.texture = tv->d_color;
				d.tex_flags = (view->color_bits == 64 ? 1 : 0) | z_flag;
				if (std::find(f->sampled_views.begin(), f->sampled_views.end(), c.tex_ref) == f->sampled_views.end())
					f->sampled_views.push_back(c.tex_ref);
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
		// polygon chunks first, triangle chunks after them (h_chunks_tri is appended below): warps resident on an SM at the same
		// time then mostly run the same half of the set-up code, which does not fit the instruction cache as a whole
		for (int32_t r0 = 0; r0 < y1 - y0; r0 += 32)
			(c.kind == SQW_CMD_TRIANGLE ? f->h_chunks_tri : f->h_chunks).push_back(make_uint2(i, (uint32_t)r0));
		span_total += (uint64_t)(y1 - y0);
		if (span_total > 0xFFFFFFFFull)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: too many scan-lines");
		if (y0 < y1)
			for (int32_t b = y0 / SQW_TILE_H; b <= (y1 - 1) / SQW_TILE_H; ++b)
				f->h_band_offsets[(size_t)b + 1]++;
	}
	f->h_chunks.insert(f->h_chunks.end(), f->h_chunks_tri.begin(), f->h_chunks_tri.end());
	f->num_chunks = (uint32_t)f->h_chunks.size();
	for (uint32_t b = 0; b < num_bands; ++b)
		f->h_band_offsets[(size_t)b + 1] += f->h_band_offsets[b];
	f->num_cmds = num_cmds;
	f->num_triangle_cmds = 0;
	f->num_general_cmds = 0; // commands that are not opaque polygons: their tiles go to the general tile kernel
	for (uint32_t i = 0; i < num_cmds; ++i)
	{
		f->num_triangle_cmds += cmds[i].kind == SQW_CMD_TRIANGLE ? 1u : 0u;
		f->num_general_cmds += (cmds[i].kind != SQW_CMD_POLYGON || cmds[i].blend != SQW_BLEND_NONE) ? 1u : 0u;
	}
	f->draw_width = view->width;
	f->draw_height = view->height;
	f->total_spans = (uint32_t)span_total;
	f->num_bands = num_bands;
	f->total_band_entries = f->h_band_offsets[num_bands];
	// The bands' ordered command lists (the binning CTAs used to scan the whole command list per band for them).
	f->h_band_list.resize(f->total_band_entries);
	f->h_band_cursor.assign(f->h_band_offsets.begin(), f->h_band_offsets.end() - 1);
	for (uint32_t i = 0; i < num_cmds; ++i)
	{
		const CmdDev& d = f->h_cmd_dev[i];
		if (d.y0 < d.y1)
			for (int32_t b = d.y0 / SQW_TILE_H; b <= (d.y1 - 1) / SQW_TILE_H; ++b)
				f->h_band_list[f->h_band_cursor[(size_t)b]++] = i;
	}
	f->stats.num_spans = span_total;
	f->stats.num_tiles = num_bands * (uint32_t)((W + SQW_TILE_W - 1) / SQW_TILE_W);

	// Scratch of the binning CTAs: band lists, then the tile lists (tiles_x slots per band entry), then the tile descriptors.
	const uint64_t tiles_x = (uint64_t)((W + SQW_TILE_W - 1) / SQW_TILE_W);
	if ((uint64_t)f->total_band_entries * tiles_x > 0xFFFFFFFFull)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw: too many commands per band");
	const size_t bin_bytes = bin_scratch_bytes(f->total_band_entries, (uint32_t)tiles_x, num_bands);
	const size_t tri_bytes = f->num_triangle_cmds ? (size_t)num_cmds * SQW_TRI_REC_BYTES + 64 : 0;
	if ((size_t)span_total * sizeof(SpanRec) + 64 > view->spans.capacity || bin_bytes > view->band_lists.capacity || tri_bytes > view->tri_recs.capacity)
	{
		SQW_CUDA(ctx, cudaStreamSynchronize(view->stream)); // growing frees scratch a queued frame may still use
		SQW_CUDA(ctx, view_scratch_reserve(view, 1, (size_t)span_total * sizeof(SpanRec) + 64));
		SQW_CUDA(ctx, view_scratch_reserve(view, 2, bin_bytes));
		if (tri_bytes)
			SQW_CUDA(ctx, view_scratch_reserve(view, 3, tri_bytes));
	}
	{
		Arena& a = f->draw_arena;
		const size_t b_cmds = (size_t)num_cmds * sizeof(sqw_draw_cmd), b_dev = (size_t)num_cmds * sizeof(CmdDev);
		const size_t b_pp = (size_t)num_polygon_points * sizeof(sqw_polygon_point), b_tp = (size_t)num_triangle_points * sizeof(sqw_triangle_point);
		const size_t b_base = f->h_chunks.size() * sizeof(uint2), b_off = f->h_band_offsets.size() * 4, b_list = f->h_band_list.size() * 4;
		SQW_CUDA(ctx, a.begin(Arena::padded(b_cmds) + Arena::padded(b_dev) + Arena::padded(b_pp) + Arena::padded(b_tp) + Arena::padded(b_base) + Arena::padded(b_off) + Arena::padded(b_list)));
		f->d_cmds = static_cast<const sqw_draw_cmd*>(a.push(cmds, b_cmds));
		f->d_cmd_dev = static_cast<const CmdDev*>(a.push(f->h_cmd_dev.data(), b_dev));
		f->d_polygon_points = static_cast<const sqw_polygon_point*>(a.push(polygon_points, b_pp));
		f->d_triangle_points = static_cast<const sqw_triangle_point*>(a.push(triangle_points, b_tp));
		f->d_chunks = static_cast<const uint2*>(a.push(f->h_chunks.data(), b_base));
		f->d_band_offsets = static_cast<const uint32_t*>(a.push(f->h_band_offsets.data(), b_off));
		f->d_band_list = static_cast<const uint32_t*>(a.push(f->h_band_list.data(), b_list));
		SQW_CUDA(ctx, a.commit(view->stream));
	}
	f->clear = clear_color != nullptr;
	f->clear_color = 0;
	if (clear_color)
		std::memcpy(&f->clear_color, clear_color, C);
	f->has_draw = true;
	return SQW_OK;
}

static_assert(sizeof(sqw_map_material) == 64, "sqw_map_material layout");

int sqw_upload_map_materials(sqw_ctx* ctx, const sqw_map_material* materials, uint32_t num_materials, const float* polygon_basis, uint32_t num_polygons)
{
	if (!ctx || !materials || !polygon_basis)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_materials: null argument");
	if (!ctx->d_map_polygons || num_materials != ctx->num_materials || num_polygons != ctx->num_map_polygons)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_materials: counts must match the geometry uploaded by sqw_upload_map_geometry");
	for (uint32_t m = 0; m < num_materials; ++m)
		for (int k = 0; k < 5; ++k)
		{
			const sqw_map_material& mat = materials[m];
			if ((mat.flags & SQW_MAT_HAS_DIFFUSE) && (mat.texture_images[k] >= ctx->images.size() || ctx->images[mat.texture_images[k]].kind != SQW_IMAGE_TEXTURE_ELEMENT))
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_materials: bad texture image");
			if ((mat.flags & SQW_MAT_HAS_EMISSIVE) && (mat.emissive_images[k] >= ctx->images.size() || ctx->images[mat.emissive_images[k]].kind != SQW_IMAGE_COLOR32))
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_materials: bad emissive image");
		}
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, synchronize_all(ctx));
	for (void** p : {reinterpret_cast<void**>(&ctx->d_map_materials), reinterpret_cast<void**>(&ctx->d_polygon_basis)})
		if (*p)
		{
			cudaFree(*p);
			*p = nullptr;
		}
	SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_map_materials), (size_t)num_materials * sizeof(sqw_map_material)));
	SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_polygon_basis), (size_t)num_polygons * 48));
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_map_materials, materials, (size_t)num_materials * sizeof(sqw_map_material), cudaMemcpyHostToDevice));
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_polygon_basis, polygon_basis, (size_t)num_polygons * 48, cudaMemcpyHostToDevice));
	return SQW_OK;
}

int sqw_frame_set_surfaces_from_polygons(
	sqw_frame* f, const float cam_pos[3], int use_directional_lightmaps, const sqw_light* lights, uint32_t num_lights, const uint32_t* clip_set_light_offsets,
	const uint16_t* clip_set_lights, uint32_t num_clip_sets, uint64_t shadow_maps_len)
{
	if (!f || !cam_pos || (num_lights && (!lights || !clip_set_light_offsets)))
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	if (!view->polygon_totals.ptr || !ctx->d_map_materials)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: needs sqw_upload_map_materials and a sqw_polygons_prepare on this view");
	if (use_directional_lightmaps ? !ctx->d_directional_lightmaps : !ctx->d_lightmaps)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: lightmap kind not uploaded");
	uint32_t num_set_lights = 0;
	if (num_lights)
	{
		if (num_clip_sets != view->polygon_num_clip_sets)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: one light list per clip set of the last sqw_polygons_prepare is needed");
		// the same checks as sqw_frame_set_surfaces: shadow maps inside the frame's array, light indices in range
		for (uint32_t l = 0; l < num_lights; ++l)
		{
			const sqw_light& li = lights[l];
			if (li.shadow_type == SQW_SHADOW_NONE)
				continue;
			if (li.shadow_type != SQW_SHADOW_CUBE && li.shadow_type != SQW_SHADOW_PROJECTOR)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: unknown shadow type");
			const uint64_t sides = li.shadow_type == SQW_SHADOW_CUBE ? 6 : 1;
			if (li.shadow_map_size == 0 || (uint64_t)li.shadow_map_offset + sides * (uint64_t)li.shadow_map_size * li.shadow_map_size > shadow_maps_len)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: a light's shadow map lies outside the shadow map array");
		}
		for (uint32_t i = 0; i < num_clip_sets; ++i)
			if (clip_set_light_offsets[i] > clip_set_light_offsets[i + 1])
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: light list offsets must not decrease");
		num_set_lights = clip_set_light_offsets[num_clip_sets];
		if (num_set_lights && !clip_set_lights)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: null light list");
		for (uint32_t i = 0; i < num_set_lights; ++i)
			if (clip_set_lights[i] >= num_lights)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_surfaces_from_polygons: light index out of range");
	}
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	const size_t C = (size_t)view->color_bits / 8;
	const uint32_t n = view->polygon_num_candidates;
	// (a queued submission of this frame may still read the old buffers: growing any of them waits for the stream first)
	const size_t need1[8] = {(size_t)n * sizeof(sqw_surface_job) + 16, (size_t)n * 4 + 16, (size_t)n * 16 + 16, (size_t)n * 16 + 16, 64, (size_t)num_lights * sizeof(sqw_light) + 16,
							 (size_t)(num_clip_sets + 1) * 4 + 16, (size_t)num_set_lights * 2 + 16};
	DevBuf* bufs1[8] = {&f->dev_jobs, &f->dev_visible_rank, &f->dev_surf_band_counts, &f->dev_surf_band_base, &f->dev_surf_totals, &f->dev_lights, &f->dev_set_light_offsets, &f->dev_set_lights};
	const size_t b_sm = (size_t)shadow_maps_len * sizeof(float);
	bool grow = num_lights && b_sm + 16 > f->d_shadow_dev.capacity;
	for (int i = 0; i < 8; ++i)
		grow |= need1[i] > bufs1[i]->capacity;
	if (grow)
		SQW_CUDA(ctx, cudaStreamSynchronize(s));
	for (int i = 0; i < 8; ++i)
		SQW_CUDA(ctx, bufs1[i]->reserve(need1[i]));
	if (num_lights && b_sm)
		SQW_CUDA(ctx, f->d_shadow_dev.reserve(b_sm + 16));
	if (num_lights)
	{
		SQW_CUDA(ctx, cudaMemcpyAsync(f->dev_lights.ptr, lights, (size_t)num_lights * sizeof(sqw_light), cudaMemcpyHostToDevice, s));
		SQW_CUDA(ctx, cudaMemcpyAsync(f->dev_set_light_offsets.ptr, clip_set_light_offsets, (size_t)(num_clip_sets + 1) * 4, cudaMemcpyHostToDevice, s));
		if (num_set_lights)
			SQW_CUDA(ctx, cudaMemcpyAsync(f->dev_set_lights.ptr, clip_set_lights, (size_t)num_set_lights * 2, cudaMemcpyHostToDevice, s));
	}
	sqw::SurfaceSetupLaunch l{};
	l.polygons = ctx->d_map_polygons;
	l.materials = ctx->d_map_materials;
	l.polygon_basis = ctx->d_polygon_basis;
	l.candidates = static_cast<const sqw_polygon_candidate*>(view->polygon_candidates.ptr);
	l.prepared = static_cast<const sqw_prepared_polygon*>(view->polygon_results.ptr);
	l.num_candidates = n;
	l.use_directional = use_directional_lightmaps ? 1u : 0u;
	l.lights = num_lights ? static_cast<const sqw_light*>(f->dev_lights.ptr) : nullptr;
	l.num_lights = num_lights;
	l.set_light_offsets = static_cast<const uint32_t*>(f->dev_set_light_offsets.ptr);
	l.set_lights = static_cast<const uint16_t*>(f->dev_set_lights.ptr);
	l.visible_rank = static_cast<uint32_t*>(f->dev_visible_rank.ptr);
	l.jobs = static_cast<sqw_surface_job*>(f->dev_jobs.ptr);
	l.band_counts = static_cast<uint32_t*>(f->dev_surf_band_counts.ptr);
	l.band_base = static_cast<uint32_t*>(f->dev_surf_band_base.ptr);
	l.totals = static_cast<unsigned long long*>(f->dev_surf_totals.ptr);
	SQW_CUDA(ctx, sqw::launch_surface_setup_phase1(l, s));
	unsigned long long totals[8] = {};
	SQW_CUDA(ctx, cudaMemcpyAsync(totals, f->dev_surf_totals.ptr, sizeof(totals), cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s)); // (also: the caller's light lists have left pageable memory)
	const uint32_t num_jobs = (uint32_t)totals[0];
	SQW_CUDA(ctx, sqw::launch_surface_bands_count(l, num_jobs, s));
	SQW_CUDA(ctx, cudaMemcpyAsync(totals, f->dev_surf_totals.ptr, sizeof(totals), cudaMemcpyDeviceToHost, s));
	sqw::PolygonsTotalsDev pt{};
	SQW_CUDA(ctx, cudaMemcpyAsync(&pt, view->polygon_totals.ptr, sizeof(pt), cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	const uint32_t kinds[4] = {(uint32_t)totals[2], (uint32_t)totals[3], (uint32_t)(totals[7] & 0xFFFFFFFFull), (uint32_t)(totals[7] >> 32)};
	const uint32_t num_bands = kinds[0] + kinds[1] + kinds[2] + kinds[3];
	SQW_CUDA(ctx, f->dev_surf_bands.reserve((size_t)num_bands * sizeof(SurfaceBand) + 16));
	l.bands = static_cast<SurfaceBand*>(f->dev_surf_bands.ptr);
	SQW_CUDA(ctx, sqw::launch_surface_setup_phase2(l, num_jobs, s));
	const uint64_t total_surface_pixels = pt.num_surfaces_pixels;
	if ((size_t)total_surface_pixels * C + 16 > view->surfaces.capacity)
	{
		SQW_CUDA(ctx, cudaStreamSynchronize(s));
		SQW_CUDA(ctx, view_scratch_reserve(view, 0, (size_t)total_surface_pixels * C + 16));
	}
	f->d_jobs = static_cast<const sqw_surface_job*>(f->dev_jobs.ptr);
	f->d_bands = static_cast<const SurfaceBand*>(f->dev_surf_bands.ptr);
	f->d_lights = num_lights ? static_cast<const sqw_light*>(f->dev_lights.ptr) : nullptr;
	// the shadow map array lives on the device only: filled by the frame's shadow map stage (sqw_frame_set_shadow_map_draw)
	f->d_shadow_maps_rw = (num_lights && b_sm) ? static_cast<float*>(f->d_shadow_dev.ptr) : nullptr;
	f->d_shadow_maps = f->d_shadow_maps_rw;
	f->shadow_maps_len = num_lights ? shadow_maps_len : 0;
	f->groups.clear();
	for (uint32_t key = 0; key < 2; ++key)
		if (kinds[2 * key] + kinds[2 * key + 1])
		{
			SurfaceGroup g{};
			g.dynamic_lights = key == 1;
			g.num_cta_bands = kinds[2 * key];
			g.band_offset = key == 0 ? 0u : kinds[0] + kinds[1];
			g.num_bands = kinds[2 * key] + kinds[2 * key + 1];
			f->groups.push_back(g);
		}
	f->stats.surface_texels = totals[4];
	f->stats.emissive_texels = totals[5];
	f->stats.lightmap_bytes = totals[6];
	f->stats.num_surface_variants = 0; // (not counted on this path)
	f->cam_pos[0] = cam_pos[0];
	f->cam_pos[1] = cam_pos[1];
	f->cam_pos[2] = cam_pos[2];
	f->has_surfaces = true;
	return SQW_OK;
}

int sqw_frame_set_draw_from_polygons(sqw_frame* f, const void* clear_color)
{
	if (!f)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	if (!view->polygon_draw_totals.ptr || !view->polygon_draw_valid)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw_from_polygons: no sqw_polygons_draw on this view since its last sqw_polygons_prepare");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	const int32_t W = (int32_t)view->width, H = (int32_t)view->height;
	const size_t C = (size_t)view->color_bits / 8;
	const uint32_t num_bands = (uint32_t)((H + SQW_TILE_H - 1) / SQW_TILE_H);
	const uint64_t tiles_x = (uint64_t)((W + SQW_TILE_W - 1) / SQW_TILE_W);
	// the lists of the last sqw_polygons_draw: their sizes, then a copy the frame owns (the view's buffers are reused per frame)
	uint32_t counts[2] = {0, 0};
	SQW_CUDA(ctx, cudaMemcpyAsync(counts, view->polygon_draw_totals.ptr, 8, cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	const uint32_t num_cmds = counts[0], num_points = counts[1];
	const size_t need1[8] = {(size_t)num_cmds * sizeof(sqw_draw_cmd) + 16, (size_t)num_points * sizeof(sqw_polygon_point) + 16, (size_t)num_cmds * sizeof(CmdDev) + 16, (size_t)num_cmds * 4 + 16,
							 (size_t)num_cmds * 4 + 16, (size_t)num_bands * 4 + 16, (size_t)(num_bands + 1) * 4 + 16, 64};
	DevBuf* bufs1[8] = {&f->dev_cmds, &f->dev_points, &f->dev_cmd_dev, &f->dev_rows, &f->dev_chunk_base, &f->dev_band_counts, &f->dev_band_offsets, &f->dev_totals};
	for (int i = 0; i < 8; ++i)
		SQW_CUDA(ctx, bufs1[i]->reserve(need1[i])); // (the stream is idle: synchronised above)
	if (num_cmds)
	{
		SQW_CUDA(ctx, cudaMemcpyAsync(f->dev_cmds.ptr, view->polygon_draw_cmds.ptr, (size_t)num_cmds * sizeof(sqw_draw_cmd), cudaMemcpyDeviceToDevice, s));
		SQW_CUDA(ctx, cudaMemcpyAsync(f->dev_points.ptr, view->polygon_draw_points.ptr, (size_t)num_points * sizeof(sqw_polygon_point), cudaMemcpyDeviceToDevice, s));
	}
	sqw::DrawSetupLaunch l{};
	l.cmds = static_cast<const sqw_draw_cmd*>(f->dev_cmds.ptr);
	l.points = static_cast<const sqw_polygon_point*>(f->dev_points.ptr);
	l.num_cmds = num_cmds;
	l.width = W;
	l.height = H;
	l.num_bands = num_bands;
	l.bytes_per_pixel = (uint32_t)C;
	l.surface_pool_elements = view->surfaces.capacity / C;
	l.cmd_dev = static_cast<CmdDev*>(f->dev_cmd_dev.ptr);
	l.rows = static_cast<uint32_t*>(f->dev_rows.ptr);
	l.chunk_base = static_cast<uint32_t*>(f->dev_chunk_base.ptr);
	l.band_counts = static_cast<uint32_t*>(f->dev_band_counts.ptr);
	l.band_offsets = static_cast<uint32_t*>(f->dev_band_offsets.ptr);
	l.totals = static_cast<uint32_t*>(f->dev_totals.ptr);
	SQW_CUDA(ctx, sqw::launch_draw_setup_phase1(l, s));
	uint32_t totals[8] = {};
	SQW_CUDA(ctx, cudaMemcpyAsync(totals, f->dev_totals.ptr, sizeof(totals), cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	if (totals[4])
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw_from_polygons: the device command list holds a command outside the surface pool (build surfaces first)");
	const uint32_t span_total = totals[0], num_chunks = totals[1], band_entries = totals[2];
	if ((uint64_t)band_entries * tiles_x > 0xFFFFFFFFull)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_draw_from_polygons: too many commands per band");
	SQW_CUDA(ctx, f->dev_chunks.reserve((size_t)num_chunks * sizeof(uint2) + 16));
	SQW_CUDA(ctx, f->dev_band_list.reserve((size_t)band_entries * 4 + 16));
	l.chunks = static_cast<uint2*>(f->dev_chunks.ptr);
	l.band_list = static_cast<uint32_t*>(f->dev_band_list.ptr);
	SQW_CUDA(ctx, sqw::launch_draw_setup_phase2(l, s));
	// the view's scratch, as sqw_frame_set_draw sizes it
	const size_t bin_bytes = bin_scratch_bytes(band_entries, (uint32_t)tiles_x, num_bands);
	if ((size_t)span_total * sizeof(SpanRec) + 64 > view->spans.capacity || bin_bytes > view->band_lists.capacity)
	{
		SQW_CUDA(ctx, cudaStreamSynchronize(s));
		SQW_CUDA(ctx, view_scratch_reserve(view, 1, (size_t)span_total * sizeof(SpanRec) + 64));
		SQW_CUDA(ctx, view_scratch_reserve(view, 2, bin_bytes));
	}
	f->sampled_views.clear();
	f->d_cmds = static_cast<const sqw_draw_cmd*>(f->dev_cmds.ptr);
	f->d_cmd_dev = static_cast<const CmdDev*>(f->dev_cmd_dev.ptr);
	f->d_polygon_points = static_cast<const sqw_polygon_point*>(f->dev_points.ptr);
	f->d_triangle_points = nullptr;
	f->d_chunks = static_cast<const uint2*>(f->dev_chunks.ptr);
	f->d_band_offsets = static_cast<const uint32_t*>(f->dev_band_offsets.ptr);
	f->d_band_list = static_cast<const uint32_t*>(f->dev_band_list.ptr);
	f->num_cmds = num_cmds;
	f->num_triangle_cmds = 0;
	f->num_general_cmds = totals[3];
	f->num_chunks = num_chunks;
	f->draw_width = view->width;
	f->draw_height = view->height;
	f->total_spans = span_total;
	f->num_bands = num_bands;
	f->total_band_entries = band_entries;
	f->stats.num_spans = span_total;
	f->stats.num_tiles = num_bands * (uint32_t)tiles_x;
	f->clear = clear_color != nullptr;
	f->clear_color = 0;
	if (clear_color)
		std::memcpy(&f->clear_color, clear_color, C);
	f->has_draw = true;
	return SQW_OK;
}

int sqw_frame_set_shadow_map_draw(
	sqw_frame* f, const sqw_depth_map* maps, uint32_t num_maps, const sqw_depth_cmd* cmds, uint32_t num_cmds, const sqw_polygon_point* points, uint32_t num_points)
{
	if (!f || (num_maps && !maps) || (num_cmds && (!cmds || !points)))
		return fail(f ? f->view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_shadow_map_draw: null argument");
	sqw_view* view = f->view;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	f->h_depth_cmd_dev.assign(num_cmds, DepthCmdDev{0, 0, 0, 0, 0});
	f->h_depth_chunks.clear();
	f->h_depth_row_chunks.clear();
	std::vector<uint8_t> cmd_seen(num_cmds, 0);
	uint64_t span_total = 0, max_end = 0;
	for (uint32_t m = 0; m < num_maps; ++m)
	{
		const sqw_depth_map& map = maps[m];
		if (map.width == 0 || map.height == 0 || map.width > 16384 || map.height > 16384 || (uint64_t)map.first_cmd + map.num_cmds > num_cmds)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_shadow_map_draw: bad map");
		max_end = std::max<uint64_t>(max_end, (uint64_t)map.offset + (uint64_t)map.width * map.height);
		for (uint32_t r0 = 0; r0 < map.height; r0 += 32)
			f->h_depth_row_chunks.push_back(make_uint2(m, r0));
		for (uint32_t ci = map.first_cmd; ci < map.first_cmd + map.num_cmds; ++ci)
		{
			const sqw_depth_cmd& c = cmds[ci];
			if (cmd_seen[ci] || c.num_vertices < 3 || c.num_vertices > 64 || (uint64_t)c.vertex_offset + c.num_vertices > num_points)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_shadow_map_draw: bad fill call (or one shared by two maps)");
			cmd_seen[ci] = 1;
			int32_t min_y = INT32_MAX, max_y = INT32_MIN;
			for (uint32_t k = 0; k < c.num_vertices; ++k)
			{
				min_y = std::min(min_y, points[c.vertex_offset + k].y);
				max_y = std::max(max_y, points[c.vertex_offset + k].y);
			}
			int32_t y0 = std::max(fixed16_round_to_int_h(min_y), 0);
			int32_t y1 = std::min(fixed16_round_to_int_h(max_y), (int32_t)map.height);
			if (y0 >= y1)
				y0 = y1 = 0;
			DepthCmdDev& d = f->h_depth_cmd_dev[ci];
			d.y0 = (int16_t)y0;
			d.y1 = (int16_t)y1;
			d.span_base = (uint32_t)span_total;
			d.map = m;
			d.depth_test = c.depth_test ? 1u : 0u;
			for (int32_t r0 = 0; r0 < y1 - y0; r0 += 32)
				f->h_depth_chunks.push_back(make_uint2(ci, (uint32_t)r0));
			span_total += (uint64_t)(y1 - y0);
			if (span_total > 0xFFFFFFFFull)
				return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_set_shadow_map_draw: too many scan-lines");
		}
	}
	if ((size_t)span_total * sizeof(DepthSpan) + 64 > f->d_depth_spans.capacity)
	{
		SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
		SQW_CUDA(ctx, f->d_depth_spans.reserve((size_t)span_total * sizeof(DepthSpan) + 64));
	}
	Arena& a = f->depth_arena;
	const size_t b_cmds = (size_t)num_cmds * sizeof(sqw_depth_cmd), b_dev = (size_t)num_cmds * sizeof(DepthCmdDev), b_pts = (size_t)num_points * sizeof(sqw_polygon_point);
	const size_t b_maps = (size_t)num_maps * sizeof(sqw_depth_map), b_ch = f->h_depth_chunks.size() * sizeof(uint2), b_rch = f->h_depth_row_chunks.size() * sizeof(uint2);
	SQW_CUDA(ctx, a.begin(Arena::padded(b_cmds) + Arena::padded(b_dev) + Arena::padded(b_pts) + Arena::padded(b_maps) + Arena::padded(b_ch) + Arena::padded(b_rch)));
	f->d_depth_cmds = static_cast<const sqw_depth_cmd*>(a.push(cmds, b_cmds));
	f->d_depth_cmd_dev = static_cast<const DepthCmdDev*>(a.push(f->h_depth_cmd_dev.data(), b_dev));
	f->d_depth_points = static_cast<const sqw_polygon_point*>(a.push(points, b_pts));
	f->d_depth_maps = static_cast<const sqw_depth_map*>(a.push(maps, b_maps));
	f->d_depth_chunks = static_cast<const uint2*>(a.push(f->h_depth_chunks.data(), b_ch));
	f->d_depth_row_chunks = static_cast<const uint2*>(a.push(f->h_depth_row_chunks.data(), b_rch));
	SQW_CUDA(ctx, a.commit(view->stream));
	f->num_depth_chunks = (uint32_t)f->h_depth_chunks.size();
	f->num_depth_row_chunks = (uint32_t)f->h_depth_row_chunks.size();
	f->depth_max_end = max_end;
	f->has_depth = num_maps != 0;
	return SQW_OK;
}

int sqw_frame_read_shadow_maps(sqw_frame* f, float* dst, uint64_t offset, uint64_t count)
{
	if (!f || !dst)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = f->view->ctx;
	if (!f->d_shadow_maps_rw || offset + count > f->shadow_maps_len)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_read_shadow_maps: range outside the frame's shadow map array");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, cudaMemcpyAsync(dst, f->d_shadow_maps_rw + offset, count * sizeof(float), cudaMemcpyDeviceToHost, f->view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(f->view->stream));
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
	cudaStream_t s = view->stream;
	uint32_t launches = 0;

	const bool do_surfaces = (stages & SQW_STAGE_SURFACES) && f->has_surfaces;
	const bool do_raster = (stages & SQW_STAGE_RASTER) && f->has_draw;
	const bool do_post = (stages & SQW_STAGE_POST) && f->has_post;
	// Tonemapping without bloom is fused into the tile kernel's epilogue (no second pass over the 64-bit buffer).
	const bool fused_tonemap = do_raster && do_post && !(f->post.bloom_sigma > 0.0f);

	float inv_scale_bgr[3] = {0, 0, 0};
	if (do_post)
	{
		// TonemappingFunction::new (postprocessor.rs:1005-1017), ColorVec lanes [B, G, R]
		const sqw_post_params& p = f->post;
		inv_scale_bgr[0] = 1.0f / (p.exposure * p.color_modulate[2]);
		inv_scale_bgr[1] = 1.0f / (p.exposure * p.color_modulate[1]);
		inv_scale_bgr[2] = 1.0f / (p.exposure * p.color_modulate[0]);
	}

	if ((stages & SQW_STAGE_SHADOW_MAPS) && f->has_depth)
	{
		NvtxRange nvtx("shadow_maps_building");
		if (!f->d_shadow_maps_rw || f->depth_max_end > f->shadow_maps_len)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_submit: shadow map fills write outside the frame's shadow map array (set the surfaces first)");
		DepthLaunch l{};
		l.cmds = f->d_depth_cmds;
		l.cmd_dev = f->d_depth_cmd_dev;
		l.points = f->d_depth_points;
		l.maps = f->d_depth_maps;
		l.chunks = f->d_depth_chunks;
		l.num_chunks = f->num_depth_chunks;
		l.row_chunks = f->d_depth_row_chunks;
		l.num_row_chunks = f->num_depth_row_chunks;
		l.spans = static_cast<DepthSpan*>(f->d_depth_spans.ptr);
		l.shadow_maps = f->d_shadow_maps_rw;
		if (l.num_chunks)
		{
			LaunchTimer timer(ctx, SQW_KERNEL_DEPTH_SETUP, s);
			SQW_CUDA(ctx, launch_depth_setup(l, s));
			++launches;
		}
		{
			LaunchTimer timer(ctx, SQW_KERNEL_DEPTH_RESOLVE, s);
			SQW_CUDA(ctx, launch_depth_resolve(l, s));
			++launches;
		}
	}

	if (do_raster && (f->draw_width != view->width || f->draw_height != view->height))
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_frame_submit: the view was resized after the command list was set");
	RasterLaunch rl{};
	const bool raster_launches = do_raster && (f->num_cmds || f->clear || fused_tonemap); // the binning CTAs also write the (empty) tile lists of a clear-only draw
	bool setup_forked = false;
	if (do_raster)
	{
		
		rl.cmds = f->d_cmds;
		rl.cmd_dev = f->d_cmd_dev;
		rl.num_cmds = f->num_cmds;
		rl.polygon_points = f->d_polygon_points;
		rl.triangle_points = f->d_triangle_points;
		rl.tri_recs = view->tri_recs.ptr;
		rl.num_triangle_cmds = f->num_triangle_cmds;
		rl.num_general_cmds = f->num_general_cmds;
		rl.spans = static_cast<SpanRec*>(view->spans.ptr);
		rl.chunks = f->d_chunks;
		rl.num_chunks = f->num_chunks;
		rl.total_spans = f->total_spans;
		{
			const uint32_t tiles_x = (view->width + SQW_TILE_W - 1) / SQW_TILE_W;
			char* const base = static_cast<char*>(view->band_lists.ptr);
			rl.band_lists = f->d_band_list;
			rl.tile_lists = reinterpret_cast<uint32_t*>(base + bin_tile_lists_offset(f->total_band_entries));
			rl.tile_desc = reinterpret_cast<uint4*>(base + bin_tile_desc_offset(f->total_band_entries, tiles_x));
			rl.general_tiles = reinterpret_cast<uint32_t*>(base + bin_general_tiles_offset(f->total_band_entries, tiles_x, f->num_bands));
		}
		rl.band_offsets = f->d_band_offsets;
		rl.num_bands = f->num_bands;
		rl.width = (int32_t)view->width;
		rl.height = (int32_t)view->height;
		rl.pitch = (int32_t)view->width;
		rl.color_buffer = view->d_color;
		rl.surfaces_base = view->surfaces.ptr;
		rl.color64 = view->color_bits == 64;
		rl.clear = f->clear;
		rl.clear_color = f->clear_color;
		if (fused_tonemap)
		{
			rl.tonemap_dst = view->d_ldr;
			rl.inv_scale_bgr[0] = inv_scale_bgr[0];
			rl.inv_scale_bgr[1] = inv_scale_bgr[1];
			rl.inv_scale_bgr[2] = inv_scale_bgr[2];
			rl.color_sums = view->d_sums; // zeroed by the set-up grid
			rl.tonemap_lut = view->d_tonemap_lut;
			view->sums_divisor = (uint64_t)view->width * view->height;
		}
		// Span set-up + binning do not depend on the surfaces: when both stages are submitted together the set-up grid runs on the
		// view's auxiliary stream BESIDE surface building and the tile kernel waits for both.  (Not while profiling: the
		// per-kernel event times are then taken one kernel at a time.)
		if (raster_launches)
		{
			cudaStream_t setup_stream = s;
			// Only for single-stream contexts (one frame at a time: the fork shortens the frame); with SQW_CTX_CONCURRENT_VIEWS other
			// views' kernels already fill the GPU and the extra events only cost.
			static const bool no_fork = std::getenv("SQW_NO_SETUP_FORK") != nullptr; // A/B switch for measurements
			if (do_surfaces && !ctx->profiling && !ctx->concurrent_views && !no_fork && view->aux_stream)
			{
				SQW_CUDA(ctx, cudaEventRecord(view->aux_fork, s));
				SQW_CUDA(ctx, cudaStreamWaitEvent(view->aux_stream, view->aux_fork, 0));
				setup_stream = view->aux_stream;
				setup_forked = true;
			}
			{
				LaunchTimer timer(ctx, SQW_KERNEL_RASTER_SETUP, setup_stream);
				SQW_CUDA(ctx, launch_raster_setup(rl, setup_stream));
			}
			launches += f->num_triangle_cmds ? 2 : 1; // + the per-triangle set-up kernel
			if (setup_forked)
				SQW_CUDA(ctx, cudaEventRecord(view->aux_join, view->aux_stream));
		}
	}

	if (do_surfaces)
	{
		NvtxRange nvtx("surfaces_preparation");
		SurfaceLaunch l{};
		l.jobs = f->d_jobs;
		l.images = ctx->d_images;
		l.lightmaps = ctx->d_lightmaps;
		l.directional_lightmaps = ctx->d_directional_lightmaps;
		l.lights = f->d_lights;
		l.shadow_maps = f->d_shadow_maps;
		l.cam_pos[0] = f->cam_pos[0];
		l.cam_pos[1] = f->cam_pos[1];
		l.cam_pos[2] = f->cam_pos[2];
		l.surfaces = view->surfaces.ptr;
		for (const SurfaceGroup& g : f->groups)
		{
			l.bands = f->d_bands + g.band_offset;
			l.num_bands = g.num_bands;
			l.num_cta_bands = g.num_cta_bands;
			{
				LaunchTimer timer(ctx, SQW_KERNEL_SURFACE_BUILD, s);
				SQW_CUDA(ctx, (ctx->fast_surface_numerics ? launch_surface_build_fast : launch_surface_build)(view->color_bits == 64, g.dynamic_lights, l, s));
			}
			++launches;
		}
	}

	if (do_raster)
	{
		NvtxRange nvtx("rasterization");
		// portal / mirror textures: wait for the views this frame samples (they may run on other streams)
		for (uint32_t vi : f->sampled_views)
			if (vi < ctx->views.size() && ctx->views[vi] && ctx->views[vi]->stream != s)
				SQW_CUDA(ctx, cudaStreamWaitEvent(s, ctx->views[vi]->done, 0));
		if (setup_forked)
			SQW_CUDA(ctx, cudaStreamWaitEvent(s, view->aux_join, 0));
		if (raster_launches)
		{
			LaunchTimer timer(ctx, fused_tonemap ? SQW_KERNEL_RASTER_TILE_TONEMAP : SQW_KERNEL_RASTER_TILE, s);
			SQW_CUDA(ctx, launch_raster_tiles(rl, s));
			launches += f->num_general_cmds ? 2 : 1; // opaque-polygon tile kernel (+ the general one when some tile needs it)
		}
	}

	if (do_post && !fused_tonemap)
	{
		NvtxRange nvtx(f->post.bloom_sigma > 0.0f ? "bloom + tonemapping" : "tonemapping");
		const sqw_post_params& p = f->post;
		PostLaunch l{};
		l.hdr = static_cast<const uint64_t*>(view->d_color);
		l.width = view->width;
		l.height = view->height;
		l.dst = view->d_ldr;
		l.inv_scale_bgr[0] = inv_scale_bgr[0];
		l.inv_scale_bgr[1] = inv_scale_bgr[1];
		l.inv_scale_bgr[2] = inv_scale_bgr[2];
		l.color_sums = view->d_sums;
		SQW_CUDA(ctx, cudaMemsetAsync(view->d_sums, 0, SQW_SUM_WORDS * sizeof(unsigned long long), s));
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
				LaunchTimer timer(ctx, SQW_KERNEL_BLOOM_DOWNSCALE, s);
				SQW_CUDA(ctx, launch_bloom_downscale(l, s));
			}
			++launches;
			if (radius > 0)
			{
				LaunchTimer timer(ctx, SQW_KERNEL_BLOOM_BLUR, s);
				SQW_CUDA(ctx, launch_bloom_blur(l, s));
				launches += 2;
			}
			{
				LaunchTimer timer(ctx, SQW_KERNEL_TONEMAP_BLOOM, s);
				SQW_CUDA(ctx, launch_tonemap_bloom(l, s));
			}
			++launches;
			view->sums_divisor = (uint64_t)l.bloom_w * l.bloom_h;
		}
		else
		{
			{
				LaunchTimer timer(ctx, SQW_KERNEL_TONEMAP, s);
				SQW_CUDA(ctx, launch_tonemap(l, s));
			}
			++launches;
			view->sums_divisor = (uint64_t)view->width * view->height;
		}
	}
	if (ctx->concurrent_views)
		SQW_CUDA(ctx, cudaEventRecord(view->done, s));
	f->stats.num_kernel_launches = launches;
	return SQW_OK;
}

// Fork / join between the context's stream and the private streams of its views (SQW_CTX_CONCURRENT_VIEWS): work
// submitted between the two calls starts after everything queued on the context's stream before sqw_batch_begin, and
// everything queued on the context's stream after sqw_batch_end starts after that work.  This is how a caller that
// owns the context's stream (events, other libraries) brackets a batch of independent views.
int sqw_batch_begin(sqw_ctx* ctx)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	if (!ctx->concurrent_views)
		return SQW_OK;
	if (!ctx->batch_event)
		SQW_CUDA(ctx, cudaEventCreateWithFlags(&ctx->batch_event, cudaEventDisableTiming));
	SQW_CUDA(ctx, cudaEventRecord(ctx->batch_event, ctx->stream));
	for (sqw_view* v : ctx->views)
		if (v && v->stream != ctx->stream)
			SQW_CUDA(ctx, cudaStreamWaitEvent(v->stream, ctx->batch_event, 0));
	return SQW_OK;
}

int sqw_batch_end(sqw_ctx* ctx)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	if (!ctx->concurrent_views)
		return SQW_OK;
	for (sqw_view* v : ctx->views)
		if (v && v->stream != ctx->stream)
		{
			SQW_CUDA(ctx, cudaEventRecord(v->done, v->stream));
			SQW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, v->done, 0));
		}
	return SQW_OK;
}

int sqw_profile_begin(sqw_ctx* ctx)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, synchronize_all(ctx));
	for (auto& t : ctx->timed)
	{
		ctx->event_pool.push_back(t.start);
		ctx->event_pool.push_back(t.stop);
	}
	ctx->timed.clear();
	ctx->profiling = true;
	return SQW_OK;
}

int sqw_upload_model_mesh(sqw_ctx* ctx, const sqw_model_mesh* mesh)
{
	if (!ctx || !mesh)
		return SQW_ERR_INVALID_ARGUMENT;
	if (mesh->kind > SQW_MESH_SKELETON_ANIMATED || mesh->num_vertices == 0 || mesh->num_vertices > 65536 || !mesh->vertices || (mesh->num_triangles && !mesh->triangles) ||
		mesh->num_frames == 0 || (mesh->kind == SQW_MESH_VERTEX_ANIMATED && !mesh->tex_coords) ||
		(mesh->kind == SQW_MESH_SKELETON_ANIMATED && mesh->frame_bones && (mesh->num_bones == 0 || !mesh->bone_parents)))
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_model_mesh: bad mesh description");
	for (uint32_t i = 0; i < mesh->num_triangles * 3; ++i)
		if (mesh->triangles[i] >= mesh->num_vertices)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_model_mesh: triangle refers to a vertex outside the mesh");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	sqw::ModelMeshDev d{};
	d.kind = mesh->kind;
	d.num_vertices = mesh->num_vertices;
	d.num_triangles = mesh->num_triangles;
	d.num_frames = mesh->num_frames;
	d.num_bones = mesh->kind == SQW_MESH_SKELETON_ANIMATED ? mesh->num_bones : 0;
	auto upload = [&](const void* src, size_t bytes, const void** dst) -> cudaError_t {
		*dst = nullptr;
		if (!src || bytes == 0)
			return cudaSuccess;
		void* p = nullptr;
		cudaError_t e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess)
			return e;
		ctx->mesh_allocations.push_back(p);
		*dst = p;
		return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
	};
	const size_t nv = mesh->num_vertices;
	const size_t vbytes = mesh->kind == SQW_MESH_NON_ANIMATED ? nv * 32 : (mesh->kind == SQW_MESH_VERTEX_ANIMATED ? (size_t)mesh->num_frames * nv * 24 : nv * 40);
	SQW_CUDA(ctx, upload(mesh->vertices, vbytes, &d.vertices));
	SQW_CUDA(ctx, upload(mesh->kind == SQW_MESH_VERTEX_ANIMATED ? mesh->tex_coords : nullptr, nv * 8, reinterpret_cast<const void**>(&d.tex_coords)));
	SQW_CUDA(ctx, upload(mesh->triangles, (size_t)mesh->num_triangles * 6, reinterpret_cast<const void**>(&d.triangles)));
	if (mesh->kind == SQW_MESH_SKELETON_ANIMATED && mesh->frame_bones)
	{
		SQW_CUDA(ctx, upload(mesh->bone_parents, (size_t)mesh->num_bones * 4, reinterpret_cast<const void**>(&d.bone_parents)));
		SQW_CUDA(ctx, upload(mesh->frame_bones, (size_t)mesh->num_frames * mesh->num_bones * 64, reinterpret_cast<const void**>(&d.frame_bones)));
	}
	ctx->meshes.push_back(d);
	// the device table is rebuilt (uploads happen at load time)
	if (ctx->meshes.size() > ctx->d_meshes_capacity)
	{
		if (ctx->d_meshes)
			cudaFree(ctx->d_meshes);
		ctx->d_meshes_capacity = std::max<size_t>(64, ctx->meshes.size() * 2);
		SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_meshes), ctx->d_meshes_capacity * sizeof(sqw::ModelMeshDev)));
	}
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_meshes, ctx->meshes.data(), ctx->meshes.size() * sizeof(sqw::ModelMeshDev), cudaMemcpyHostToDevice));
	return (int)ctx->meshes.size() - 1;
}

int sqw_models_prepare(sqw_view* view, const sqw_model_instance* instances, uint32_t num_instances)
{
	if (!view || (num_instances && !instances))
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	std::vector<sqw_model_instance> inst(instances, instances + num_instances);
	uint64_t num_vertices = 0, num_triangles = 0, bones = 0, sort_slots = 0;
	uint32_t max_vertices = 0, max_bones = 0;
	for (sqw_model_instance& in : inst)
	{
		if (in.mesh >= ctx->meshes.size())
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_models_prepare: unknown mesh");
		const sqw::ModelMeshDev& m = ctx->meshes[in.mesh];
		if (in.frames[0] >= m.num_frames || in.frames[1] >= m.num_frames)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_models_prepare: animation frame outside the mesh");
		num_vertices = std::max<uint64_t>(num_vertices, (uint64_t)in.vertex_offset + m.num_vertices);
		num_triangles = std::max<uint64_t>(num_triangles, (uint64_t)in.triangle_offset + m.num_triangles);
		max_vertices = std::max(max_vertices, m.num_vertices);
		in.bone_offset = (uint32_t)bones;
		if (m.kind == SQW_MESH_SKELETON_ANIMATED && m.frame_bones)
		{
			bones += std::min<uint32_t>(m.num_bones, 255u);
			max_bones = std::max(max_bones, m.num_bones);
		}
		uint32_t n = 1;
		while (n < m.num_triangles)
			n <<= 1;
		in.sort_offset = (uint32_t)sort_slots;
		sort_slots += n;
	}
	if (num_vertices > (1ull << 30) || num_triangles > (1ull << 30) || sort_slots > (1ull << 31))
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_models_prepare: too many vertices / triangles");
	cudaStream_t s = view->stream;
	const size_t need[6] = {inst.size() * sizeof(sqw_model_instance) + 16, (size_t)num_vertices * 32 + 16, (size_t)num_triangles * 6 + 16, (size_t)num_instances * 4 + 16,
							(size_t)bones * 25 * 4 + 16, (size_t)sort_slots * 8 + 16};
	DevBuf* bufs[6] = {&view->model_instances, &view->model_vertices, &view->model_triangles, &view->model_counts, &view->model_bones, &view->model_sort};
	for (int i = 0; i < 6; ++i)
		if (need[i] > bufs[i]->capacity)
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(s)); // growing frees the old buffer
			SQW_CUDA(ctx, bufs[i]->reserve(need[i]));
		}
	view->model_num_instances = num_instances;
	view->model_num_vertices = num_vertices;
	view->model_num_triangles = num_triangles;
	if (num_instances == 0)
		return SQW_OK;
	SQW_CUDA(ctx, cudaMemcpyAsync(view->model_instances.ptr, inst.data(), inst.size() * sizeof(sqw_model_instance), cudaMemcpyHostToDevice, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s)); // `inst` is a local staging copy
	sqw::ModelLaunch l{};
	l.instances = static_cast<const sqw_model_instance*>(view->model_instances.ptr);
	l.meshes = ctx->d_meshes;
	l.num_instances = num_instances;
	l.max_vertices = max_vertices;
	l.max_bones = max_bones;
	l.bone_scratch = static_cast<float*>(view->model_bones.ptr);
	l.sort_scratch = static_cast<unsigned long long*>(view->model_sort.ptr);
	l.vertices = static_cast<float*>(view->model_vertices.ptr);
	l.triangles = static_cast<uint16_t*>(view->model_triangles.ptr);
	l.counts = static_cast<uint32_t*>(view->model_counts.ptr);
	{
		NvtxRange nvtx("triangle_models_preparation");
		LaunchTimer timer(ctx, SQW_KERNEL_MODELS_PREPARE, s);
		SQW_CUDA(ctx, sqw::launch_models_prepare(l, s));
	}
	return SQW_OK;
}

int sqw_models_read(sqw_view* view, sqw_model_vertex* vertices, uint64_t num_vertices, uint16_t* triangles, uint64_t num_triangles, uint32_t* counts, uint32_t num_instances)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	if (num_vertices > view->model_num_vertices || num_triangles > view->model_num_triangles || num_instances > view->model_num_instances)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_models_read: more than the last sqw_models_prepare produced");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	if (vertices && num_vertices)
		SQW_CUDA(ctx, cudaMemcpyAsync(vertices, view->model_vertices.ptr, (size_t)num_vertices * 32, cudaMemcpyDeviceToHost, s));
	if (triangles && num_triangles)
		SQW_CUDA(ctx, cudaMemcpyAsync(triangles, view->model_triangles.ptr, (size_t)num_triangles * 6, cudaMemcpyDeviceToHost, s));
	if (counts && num_instances)
		SQW_CUDA(ctx, cudaMemcpyAsync(counts, view->model_counts.ptr, (size_t)num_instances * 4, cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	return SQW_OK;
}

static_assert(sizeof(sqw_map_polygon) == 80 && sizeof(sqw_prepared_polygon) == 64 && sizeof(sqw_polygon_candidate) == 8 && sizeof(sqw_polygons_frame) == 144, "polygon preparation layouts");

int sqw_upload_map_geometry(sqw_ctx* ctx, const sqw_map_polygon* polygons, uint32_t num_polygons, const float* vertices, uint32_t num_vertices, const uint8_t* material_flags, uint32_t num_materials)
{
	if (!ctx || !polygons || !vertices || !material_flags || num_polygons == 0 || num_vertices == 0 || num_materials == 0)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_geometry: null or empty argument");
	for (uint32_t i = 0; i < num_polygons; ++i)
		if (polygons[i].num_vertices < 3 || (uint64_t)polygons[i].first_vertex + polygons[i].num_vertices > num_vertices || polygons[i].texture >= num_materials)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_upload_map_geometry: polygon refers to vertices or a material outside the map");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, synchronize_all(ctx));
	for (void** p : {reinterpret_cast<void**>(&ctx->d_map_polygons), reinterpret_cast<void**>(&ctx->d_map_vertices), reinterpret_cast<void**>(&ctx->d_material_flags)})
		if (*p)
		{
			cudaFree(*p);
			*p = nullptr;
		}
	ctx->num_map_polygons = ctx->num_map_vertices = ctx->num_materials = 0;
	SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_map_polygons), (size_t)num_polygons * sizeof(sqw_map_polygon)));
	SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_map_vertices), (size_t)num_vertices * 12));
	SQW_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_material_flags), (size_t)num_materials));
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_map_polygons, polygons, (size_t)num_polygons * sizeof(sqw_map_polygon), cudaMemcpyHostToDevice));
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_map_vertices, vertices, (size_t)num_vertices * 12, cudaMemcpyHostToDevice));
	SQW_CUDA(ctx, cudaMemcpy(ctx->d_material_flags, material_flags, (size_t)num_materials, cudaMemcpyHostToDevice));
	ctx->num_map_polygons = num_polygons;
	ctx->num_map_vertices = num_vertices;
	ctx->num_materials = num_materials;
	return SQW_OK;
}

int sqw_polygons_prepare(sqw_view* view, const sqw_polygons_frame* frame, const sqw_polygon_candidate* candidates, uint32_t num_candidates, const float* clip_planes, uint32_t num_clip_sets)
{
	if (!view || !frame || (num_candidates && (!candidates || !clip_planes || num_clip_sets == 0)))
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_prepare: null argument");
	sqw_ctx* ctx = view->ctx;
	if (!ctx->d_map_polygons)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_prepare: no map geometry uploaded (sqw_upload_map_geometry)");
	for (uint32_t i = 0; i < num_candidates; ++i)
		if (candidates[i].polygon >= ctx->num_map_polygons || candidates[i].clip_set >= num_clip_sets)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_prepare: candidate refers to a polygon or clip set that does not exist");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	const size_t need[4] = {(size_t)num_candidates * sizeof(sqw_polygon_candidate) + 16, (size_t)num_clip_sets * 96 + 16, (size_t)num_candidates * sizeof(sqw_prepared_polygon) + 16,
							sizeof(sqw::PolygonsTotalsDev)};
	DevBuf* bufs[4] = {&view->polygon_candidates, &view->polygon_clip_planes, &view->polygon_results, &view->polygon_totals};
	for (int i = 0; i < 4; ++i)
		if (need[i] > bufs[i]->capacity)
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(s)); // growing frees the old buffer
			SQW_CUDA(ctx, bufs[i]->reserve(need[i]));
		}
	view->polygon_num_candidates = num_candidates;
	view->polygon_num_clip_sets = num_clip_sets;
	view->polygons_frame = *frame;
	view->polygon_num_order = 0;
	view->polygon_draw_valid = false;
	if (num_candidates)
	{
		SQW_CUDA(ctx, cudaMemcpyAsync(view->polygon_candidates.ptr, candidates, (size_t)num_candidates * sizeof(sqw_polygon_candidate), cudaMemcpyHostToDevice, s));
		SQW_CUDA(ctx, cudaMemcpyAsync(view->polygon_clip_planes.ptr, clip_planes, (size_t)num_clip_sets * 96, cudaMemcpyHostToDevice, s));
	}
	sqw::PolygonsLaunch l{};
	l.polygons = ctx->d_map_polygons;
	l.vertices = ctx->d_map_vertices;
	l.material_flags = ctx->d_material_flags;
	l.num_materials = ctx->num_materials;
	l.frame = *frame;
	l.candidates = static_cast<const sqw_polygon_candidate*>(view->polygon_candidates.ptr);
	l.num_candidates = num_candidates;
	l.clip_planes = static_cast<const float*>(view->polygon_clip_planes.ptr);
	l.out = static_cast<sqw_prepared_polygon*>(view->polygon_results.ptr);
	l.totals = static_cast<sqw::PolygonsTotalsDev*>(view->polygon_totals.ptr);
	{
		NvtxRange nvtx("polygons_preparation");
		LaunchTimer timer(ctx, SQW_KERNEL_POLYGONS_PREPARE, s);
		SQW_CUDA(ctx, sqw::launch_polygons_prepare(l, s));
	}
	// the lists are the caller's (pageable) memory: they must have left it before the call returns
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	return SQW_OK;
}

int sqw_polygons_read(sqw_view* view, sqw_prepared_polygon* out, uint32_t num, sqw_polygons_totals* totals)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	if (num > view->polygon_num_candidates || !view->polygon_totals.ptr)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_read: more than the last sqw_polygons_prepare produced");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	if (out && num)
		SQW_CUDA(ctx, cudaMemcpyAsync(out, view->polygon_results.ptr, (size_t)num * sizeof(sqw_prepared_polygon), cudaMemcpyDeviceToHost, s));
	sqw::PolygonsTotalsDev t{};
	uint32_t sky_material = 0;
	if (totals)
		SQW_CUDA(ctx, cudaMemcpyAsync(&t, view->polygon_totals.ptr, sizeof(t), cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	if (totals)
	{
		std::memset(totals, 0, sizeof(*totals));
		totals->num_visible = t.num_visible;
		totals->num_surfaces_pixels = t.num_surfaces_pixels;
		totals->has_sky = t.first_sky_candidate != 0xFFFFFFFFu ? 1u : 0u;
		if (totals->has_sky)
		{
			std::memcpy(totals->sky_bounds, t.sky_bounds, sizeof(t.sky_bounds));
			// material of the first sky polygon in candidate order: candidate -> polygon -> texture (two small reads)
			sqw_polygon_candidate c{};
			SQW_CUDA(ctx, cudaMemcpy(&c, static_cast<const sqw_polygon_candidate*>(view->polygon_candidates.ptr) + t.first_sky_candidate, sizeof(c), cudaMemcpyDeviceToHost));
			SQW_CUDA(ctx, cudaMemcpy(&sky_material, &ctx->d_map_polygons[c.polygon].texture, 4, cudaMemcpyDeviceToHost));
			totals->sky_material = sky_material;
		}
	}
	return SQW_OK;
}

int sqw_polygons_draw(sqw_view* view, const uint32_t* order, uint32_t num_order)
{
	if (!view || (num_order && !order))
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_draw: null argument");
	sqw_ctx* ctx = view->ctx;
	if (!view->polygon_totals.ptr || !ctx->d_map_polygons)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_draw: no sqw_polygons_prepare on this view yet");
	for (uint32_t i = 0; i < num_order; ++i)
		if (order[i] >= view->polygon_num_candidates)
			return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_draw: order refers to a candidate the last sqw_polygons_prepare did not have");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	const size_t n = num_order;
	const size_t need[7] = {n * 4 + 16, n * sizeof(sqw_draw_cmd) + 16, n * 24 * sizeof(sqw_polygon_point) + 16, n * 8 + 16, n * sizeof(sqw_draw_cmd) + 16,
							n * 24 * sizeof(sqw_polygon_point) + 16, 16};
	DevBuf* bufs[7] = {&view->polygon_order, &view->polygon_draw_scratch_cmds, &view->polygon_draw_scratch_points, &view->polygon_draw_offsets, &view->polygon_draw_cmds,
					   &view->polygon_draw_points, &view->polygon_draw_totals};
	for (int i = 0; i < 7; ++i)
		if (need[i] > bufs[i]->capacity)
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(s));
			SQW_CUDA(ctx, bufs[i]->reserve(need[i]));
		}
	view->polygon_num_order = num_order;
	if (num_order)
		SQW_CUDA(ctx, cudaMemcpyAsync(view->polygon_order.ptr, order, n * 4, cudaMemcpyHostToDevice, s));
	sqw::PolygonsDrawLaunch l{};
	l.polygons = ctx->d_map_polygons;
	l.vertices = ctx->d_map_vertices;
	l.material_flags = ctx->d_material_flags;
	l.num_materials = ctx->num_materials;
	l.frame = view->polygons_frame;
	l.candidates = static_cast<const sqw_polygon_candidate*>(view->polygon_candidates.ptr);
	l.clip_planes = static_cast<const float*>(view->polygon_clip_planes.ptr);
	l.prepared = static_cast<const sqw_prepared_polygon*>(view->polygon_results.ptr);
	l.order = static_cast<const uint32_t*>(view->polygon_order.ptr);
	l.num_order = num_order;
	l.scratch_cmds = static_cast<sqw_draw_cmd*>(view->polygon_draw_scratch_cmds.ptr);
	l.scratch_points = static_cast<sqw_polygon_point*>(view->polygon_draw_scratch_points.ptr);
	l.offsets = static_cast<uint2*>(view->polygon_draw_offsets.ptr);
	l.cmds = static_cast<sqw_draw_cmd*>(view->polygon_draw_cmds.ptr);
	l.points = static_cast<sqw_polygon_point*>(view->polygon_draw_points.ptr);
	l.totals = static_cast<uint32_t*>(view->polygon_draw_totals.ptr);
	{
		NvtxRange nvtx("polygons_draw_setup");
		LaunchTimer timer(ctx, SQW_KERNEL_POLYGONS_PREPARE, s);
		SQW_CUDA(ctx, sqw::launch_polygons_draw(l, s));
	}
	SQW_CUDA(ctx, cudaStreamSynchronize(s)); // `order` is the caller's pageable memory
	view->polygon_draw_valid = true;
	return SQW_OK;
}

int sqw_polygons_draw_read(sqw_view* view, sqw_draw_cmd* cmds, uint32_t cmds_capacity, sqw_polygon_point* points, uint32_t points_capacity, uint32_t counts_out[2])
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	sqw_ctx* ctx = view->ctx;
	if (!view->polygon_draw_totals.ptr)
		return fail(ctx, SQW_ERR_INVALID_ARGUMENT, "sqw_polygons_draw_read: no sqw_polygons_draw on this view yet");
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = view->stream;
	uint32_t totals[2] = {0, 0};
	SQW_CUDA(ctx, cudaMemcpyAsync(totals, view->polygon_draw_totals.ptr, 8, cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	if (counts_out)
	{
		counts_out[0] = totals[0];
		counts_out[1] = totals[1];
	}
	if (cmds && std::min(totals[0], cmds_capacity))
		SQW_CUDA(ctx, cudaMemcpyAsync(cmds, view->polygon_draw_cmds.ptr, (size_t)std::min(totals[0], cmds_capacity) * sizeof(sqw_draw_cmd), cudaMemcpyDeviceToHost, s));
	if (points && std::min(totals[1], points_capacity))
		SQW_CUDA(ctx, cudaMemcpyAsync(points, view->polygon_draw_points.ptr, (size_t)std::min(totals[1], points_capacity) * sizeof(sqw_polygon_point), cudaMemcpyDeviceToHost, s));
	SQW_CUDA(ctx, cudaStreamSynchronize(s));
	return SQW_OK;
}

int sqw_set_fast_surface_numerics(sqw_ctx* ctx, int enable)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	ctx->fast_surface_numerics = enable != 0;
	return SQW_OK;
}

int sqw_self_test(sqw_ctx* ctx)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	unsigned long long* d_mismatches = nullptr;
	SQW_CUDA(ctx, cudaMalloc(&d_mismatches, sizeof(unsigned long long)));
	unsigned long long mismatches = ~0ull;
	cudaError_t e = cudaMemsetAsync(d_mismatches, 0, sizeof(unsigned long long), ctx->stream);
	if (e == cudaSuccess)
		e = launch_z_from_inv_z_test(d_mismatches, ctx->stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(&mismatches, d_mismatches, sizeof(mismatches), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(ctx->stream);
	cudaFree(d_mismatches);
	SQW_CUDA(ctx, e);
	if (mismatches != 0)
		return fail(ctx, SQW_ERR_CUDA, "sqw_self_test: z_from_inv_z differs from the u32 division for " + std::to_string(mismatches) + " divisors");
	return SQW_OK;
}

int sqw_profile_end(sqw_ctx* ctx, sqw_kernel_times* out)
{
	if (!ctx || !out)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	ctx->profiling = false;
	SQW_CUDA(ctx, synchronize_all(ctx));
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
	unsigned long long slots[SQW_SUM_WORDS];
	SQW_CUDA(ctx, cudaMemcpyAsync(slots, view->d_sums, sizeof(slots), cudaMemcpyDeviceToHost, view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
	const float inv = 1.0f / (float)view->sums_divisor; // postprocessor.rs:176-179 / :840-843
	for (int i = 0; i < 3; ++i)
	{
		unsigned long long sum = 0;
		for (int slot = 0; slot < SQW_SUM_SLOTS; ++slot)
			sum += slots[slot * 4 + i];
		avg_color_out[i] = (float)sum * inv;
	}
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
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * C, view->d_color, (size_t)view->width * C, (size_t)view->width * C, view->height, cudaMemcpyDeviceToHost, view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
	return SQW_OK;
}

int sqw_write_pixels(sqw_view* view, const void* src, size_t src_pitch)
{
	if (!view || !src || src_pitch < view->width)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_write_pixels: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t C = (size_t)view->color_bits / 8;
	SQW_CUDA(ctx, cudaMemcpy2DAsync(view->d_color, (size_t)view->width * C, src, src_pitch * C, (size_t)view->width * C, view->height, cudaMemcpyHostToDevice, view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
	return SQW_OK;
}

int sqw_read_pixels32(sqw_view* view, uint32_t* dst, size_t dst_pitch)
{
	if (!view || !dst || dst_pitch < view->width || view->color_bits != 64)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_read_pixels32: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * 4, view->d_ldr, (size_t)view->width * 4, (size_t)view->width * 4, view->height, cudaMemcpyDeviceToHost, view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
	return SQW_OK;
}

int sqw_read_pixels32_async(sqw_view* view, uint32_t* dst, size_t dst_pitch)
{
	if (!view || !dst || dst_pitch < view->width || view->color_bits != 64)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_read_pixels32_async: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->compressed_readback)
	{
		// run-length rows written by the GPU straight into mapped host memory, expanded into dst by the pool (readback_decode.cu)
		if (view->readback_pending) // the previous picture of this view is still travelling: its buffer is about to be reused
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
			finish_readback(view);
		}
		const size_t words = (size_t)view->width * view->height + view->height;
		if (words > view->h_rle_words)
		{
			if (view->h_rle)
				SQW_CUDA(ctx, cudaFreeHost(view->h_rle));
			view->h_rle = nullptr;
			view->h_rle_words = 0;
			SQW_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&view->h_rle), words * 4, cudaHostAllocMapped | cudaHostAllocPortable));
			view->h_rle_words = words;
		}
		if (!ctx->readback_pool)
		{
			unsigned n = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
			if (const char* e = std::getenv("SQW_READBACK_THREADS"))
				n = (unsigned)std::max(1, std::min(64, std::atoi(e)));
			ctx->readback_pool = new ReadbackPool();
			ctx->readback_pool->start(n);
		}
		uint32_t* const counts = view->h_rle + (size_t)view->width * view->height;
		SQW_CUDA(ctx, sqw::launch_picture_rle(view->d_ldr, view->width, view->height, view->width, view->h_rle, view->width, counts, view->stream));
		ReadbackJob& job = view->readback_job;
		job.rle = view->h_rle;
		job.counts = counts;
		job.dst = dst;
		job.dst_pitch = dst_pitch;
		job.width = view->width;
		job.height = view->height;
		job.row_capacity = view->width;
		job.next_row.store(0);
		job.rows_done.store(0);
		job.words.store(0);
		job.workers_inside = 0;
		job.finished = false;
		view->readback_pending = true;
		SQW_CUDA(ctx, cudaLaunchHostFunc(view->stream, [](void* p) {
			sqw_view* v = static_cast<sqw_view*>(p);
			v->ctx->readback_pool->push(&v->readback_job);
		}, view));
		return SQW_OK;
	}
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * 4, view->d_ldr, (size_t)view->width * 4, (size_t)view->width * 4, view->height, cudaMemcpyDeviceToHost, view->stream));
	return SQW_OK;
}

int sqw_view_get_readback_stats(const sqw_view* view, uint64_t out[2])
{
	if (!view || !out)
		return SQW_ERR_INVALID_ARGUMENT;
	out[0] = view->readback_pictures;
	out[1] = view->readback_bytes_total;
	return SQW_OK;
}

int sqw_set_compressed_readback(sqw_ctx* ctx, int enable)
{
	if (!ctx)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	for (sqw_view* v : ctx->views) // pictures on their way keep the mode they were started with
		if (v && v->readback_pending)
		{
			SQW_CUDA(ctx, cudaStreamSynchronize(v->stream));
			finish_readback(v);
		}
	ctx->compressed_readback = enable != 0;
	return SQW_OK;
}

int sqw_read_pixels_async(sqw_view* view, void* dst, size_t dst_pitch)
{
	if (!view || !dst || dst_pitch < view->width)
		return fail(view ? view->ctx : nullptr, SQW_ERR_INVALID_ARGUMENT, "sqw_read_pixels_async: bad argument");
	sqw_ctx* ctx = view->ctx;
	SQW_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t C = (size_t)view->color_bits / 8;
	SQW_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch * C, view->d_color, (size_t)view->width * C, (size_t)view->width * C, view->height, cudaMemcpyDeviceToHost, view->stream));
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
	SQW_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<const char*>(view->surfaces.ptr) + offset * C, count * C, cudaMemcpyDeviceToHost, view->stream));
	SQW_CUDA(ctx, cudaStreamSynchronize(view->stream));
	return SQW_OK;
}

int sqw_view_synchronize(sqw_view* view)
{
	if (!view)
		return SQW_ERR_INVALID_ARGUMENT;
	SQW_CUDA(view->ctx, cudaSetDevice(view->ctx->device));
	SQW_CUDA(view->ctx, cudaStreamSynchronize(view->stream));
	finish_readback(view);
	return SQW_OK;
}

} // extern "C"
