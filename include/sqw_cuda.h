/*
 * sqw_cuda.h — C ABI of the B200 (sm_100a) world-drawing path for Square-Wheel.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own; its public
 * renderer surface is
 *     Renderer::new / prepare_frame::<ColorT> / draw_frame::<ColorT>   (src/square_wheel/renderer.rs:24,61,106)
 *     Postprocessor::get_hdr_buffer / perform_postprocessing           (src/square_wheel/postprocessor.rs:56,68)
 * and those five signatures stay as they are.  What is replaced are four internal seams:
 *
 *   S0 load      texture / lightmap ownership     map_materials_processor.rs:29, partial_renderer.rs:56
 *   S1 surfaces  body of build_polygons_surfaces  partial_renderer.rs:1682-1831 (-> surfaces.rs:66,95,854)
 *   S2 raster    perform_rasterization            partial_renderer.rs:1192-1283 (-> rasterizer.rs:35,731)
 *   S3 post      perform_postprocessing           postprocessor.rs:68-204
 *   readback     `pixels: &mut [ColorT]` argument of draw_frame / perform_postprocessing
 *
 * Conventions
 *   - plain C types only; every struct below is `#[repr(C)]`-compatible and has a fixed size (checked with
 *     static asserts in the implementation).
 *   - the library COPIES whatever it is given; the caller keeps ownership of all host memory.
 *   - every entry returns SQW_OK (0) or a negative error code and never unwinds across the boundary
 *     (the reference aborts on panic, src/Cargo.toml:8); sqw_last_error() gives the text.
 *   - a context may be used from any ONE thread at a time (the reference calls prepare_frame from an
 *     arbitrary rayon worker, host.rs:438-449); every entry does its own cudaSetDevice.
 *   - pixel formats: Color32 = A<<24|R<<16|G<<8|B (common/color.rs:16-24),
 *                    Color64 = A<<48|R<<32|G<<16|B (common/color.rs:139-147).
 *   - numerics follow the PORTABLE branch of fast_math.rs (:370-725): exact 1/x and 1/sqrt, mul+add instead
 *     of fma, truncating float->int.
 */
#ifndef SQW_CUDA_H
#define SQW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQW_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------------------------- */
#define SQW_OK 0
#define SQW_ERR_INVALID_ARGUMENT (-1)
#define SQW_ERR_CUDA (-2)
#define SQW_ERR_OUT_OF_MEMORY (-3)
#define SQW_ERR_NO_DEVICE (-4)

typedef struct sqw_ctx sqw_ctx;   /* one per map replica on one GPU (Renderer::new)                     */
typedef struct sqw_view sqw_view; /* one render target + scratch (a PartialRenderer, partial_renderer.rs:15) */

/* ---- S0: static data ---------------------------------------------------------------------------------- */

/* Image element kinds.  One "image" is one mip of one texture. */
#define SQW_IMAGE_TEXTURE_ELEMENT 0 /* textures.rs:20-27  TextureElement {Color32 diffuse; i8x4 normal.xyz,roughness} 8 B */
#define SQW_IMAGE_COLOR32 1         /* TextureLite = image::Image (common/image.rs:3-8), skybox 32-bit sides   4 B */
#define SQW_IMAGE_COLOR64 2         /* SkyboxSideTexture<Color64> (textures.rs:82-87)                           8 B */

/* bsp_map_compact.rs:142-156.  LightmapElement is [f32;3]; the directional element is 10 floats. */
typedef struct sqw_directional_lightmap_element
{
	float ambient_light[3];
	float light_direction_vector_scaled[3];
	float directional_light_deviation;
	float directional_light_color[3];
} sqw_directional_lightmap_element;

typedef struct sqw_map_desc
{
	const float* lightmaps_data; /* [num_lightmap_elements][3], de-quantised (bsp_map_save_load.rs:208-221) */
	const sqw_directional_lightmap_element* directional_lightmaps_data; /* same count, or NULL */
	uint64_t num_lightmap_elements;
	uint32_t max_images; /* capacity of the image table (0 = default 4096) */
	uint32_t flags;      /* SQW_CTX_* */
} sqw_map_desc;

/* Every view of the context issues its work (uploads, kernels, read-backs) on a private CUDA stream, so frames of
 * different views overlap on the GPU (the view-batch workload; the reference overlaps frame N's prepare
 * with frame N-1's swap the same way, host.rs:438-449).  Work of ONE view stays in submission order.  Without the flag
 * everything runs in order on the context's stream. */
#define SQW_CTX_CONCURRENT_VIEWS 1u
/* Surface building (stage 1) uses the numerics CLASS of the reference's release build instead of its portable branch: hardware
 * reciprocal and reciprocal square root for inv_fast / inv_sqrt_fast and fused multiply-add for f32_mul_add (fast_math.rs:13-22,
 * :729-733; the release build is compiled for haswell, src/build_release.sh:2, where those are rcpss / rsqrtss / vfmadd).  Without
 * the flag every stage is bit-exact against the portable branch (fast_math.rs:370-725).  With it, surfaces differ from both
 * reference builds by no more than those builds differ from each other (north_star's +-1 LSB of the final 8-bit picture);
 * stages 2 and 3 are not affected. */
#define SQW_CTX_FAST_SURFACE_NUMERICS 2u
/* sqw_read_pixels32_async moves the picture as run-length rows instead of raw pixels: the alpha byte of the tonemapped picture is
 * 255 by construction, so a run of equal pixels travels as one word B | G << 8 | R << 16 | (length - 1) << 24, written by the GPU
 * straight into page-locked, device-mapped host memory and expanded into the caller's buffer by a small pool of host threads of the
 * library (SQW_READBACK_THREADS, default min(8, cores / 2)).  Lossless (rows whose alpha is not 255 travel raw); the caller's buffer
 * needs not be page-locked.  At 4K a surface texel covers several pixels, so the 33 MB picture shrinks several times and the PCIe
 * read-back stops being the end-to-end limit.  The picture is complete after sqw_view_synchronize, like the raw asynchronous copy. */
#define SQW_CTX_COMPRESSED_READBACK 4u

/* Create a context on CUDA device `device` and upload the static map data.  Replaces the ownership half of
 * Renderer::new (renderer.rs:24).  `external_stream` may be a cudaStream_t the caller wants all work of this
 * context to be issued on (pass NULL for a private non-blocking stream).
 * Side effect: sets the device-wide cudaLimitPersistingL2CacheSize to 48 MB (environment SQW_L2_PERSIST_MB overrides, 0 = leave
 * it alone) and marks the views' per-frame scratch (surface pool, span table, tile lists: one block per view) L2-persisting on the
 * streams the library owns, so that what one kernel of a frame hands to the next survives in L2 while other views stream their
 * pixels; reset by sqw_destroy. */
int sqw_create(sqw_ctx** out_ctx, int device, const sqw_map_desc* map, void* external_stream);
void sqw_destroy(sqw_ctx* ctx);
const char* sqw_last_error(const sqw_ctx* ctx); /* ctx may be NULL: error of the last failed sqw_create */
int sqw_abi_version(void);

/* Upload one image (one mip).  Returns the image id (>= 0) or a negative error.
 * Replaces texture ownership in MapMaterialsProcessor::new (map_materials_processor.rs:29-167). */
int sqw_upload_image(sqw_ctx* ctx, int kind, uint32_t width, uint32_t height, const void* pixels);
/* Replace the pixels of an existing image (animated textures, map_materials_processor.rs:295-302). */
int sqw_update_image(sqw_ctx* ctx, int image_id, const void* pixels);

/* ---- views (render targets) ---------------------------------------------------------------------------- */

/* color_bits: 32 (Color32 framebuffer, LDR path host.rs:416) or 64 (Color64 HDR buffer, postprocessor.rs:56). */
int sqw_view_create(sqw_ctx* ctx, sqw_view** out_view, uint32_t width, uint32_t height, int color_bits);
void sqw_view_destroy(sqw_view* view);

/* Change the logical size of a view within the size it was created with (portal textures change resolution from frame
 * to frame, partial_renderer.rs:2130-2182, while the reference's pixel pool only ever grows, :2163-2176).  The pitch of the
 * colour buffer becomes `width`.  Contents are undefined afterwards. */
int sqw_view_set_size(sqw_view* view, uint32_t width, uint32_t height);
/* Index of the view in its context: the value to put into sqw_draw_cmd::tex_ref with SQW_TEX_VIEW. */
int sqw_view_index(const sqw_view* view);

/* ---- S1: surfaces --------------------------------------------------------------------------------------- */

#define SQW_SURF_HAS_NORMAL_MAP 1u        /* Texture::has_normal_map        textures.rs:14 */
#define SQW_SURF_HAS_NON_ONE_ROUGHNESS 2u /* Texture::has_non_one_roughness textures.rs:15 */
#define SQW_SURF_IS_METAL 4u              /* Texture::is_metal              textures.rs:16 */
#define SQW_SURF_DIRECTIONAL_LIGHTMAP 8u  /* use_directional_lightmap       partial_renderer.rs:1676 */
#define SQW_SURF_APPLY_DIFFUSE 16u        /* apply_diffuse_layer            partial_renderer.rs:1700 */
#define SQW_SURF_EMISSIVE 32u             /* get_emissive_texture is Some   partial_renderer.rs:1801 */

#define SQW_MAX_POLYGON_LIGHTS 6 /* partial_renderer.rs:1666 */
#define SQW_NO_LIGHTMAP_SCALE 31 /* surfaces.rs:429 */

/* One call of build_surface_{simple,directional}_lightmap (+ mix_surface_with_texture for the emissive layer),
 * i.e. the arguments assembled in partial_renderer.rs:1682-1819.  128 bytes. */
typedef struct sqw_surface_job
{
	float basis_u[3];      /* PolygonBasisVecs (surfaces.rs:7-14), already scaled for the mip and shift-corrected */
	float basis_v[3];      /* (partial_renderer.rs:1736-1740) */
	float basis_start[3];
	float basis_normal[3];
	uint32_t surface_size[2];
	int32_t tc_start[2];            /* surface_tc_min + texture shift >> mip (partial_renderer.rs:1690-1693) */
	uint32_t texture_image;         /* image id of the TextureElement mip */
	uint32_t flags;                 /* SQW_SURF_* */
	uint32_t lightmap_size[2];      /* lightmap.rs:11 */
	uint32_t lightmap_scale_log2;   /* LIGHTMAP_SCALE_LOG2 - mip; ignored when lightmap_offset == 0 */
	uint32_t lightmap_tc_shift[2];  /* partial_renderer.rs:1744-1752 */
	uint32_t lightmap_offset;       /* Polygon::lightmap_data_offset; 0 = no lightmap */
	uint32_t num_lights;            /* <= SQW_MAX_POLYGON_LIGHTS */
	uint16_t lights[SQW_MAX_POLYGON_LIGHTS]; /* indices into the lights array of the call */
	uint32_t emissive_image;        /* COLOR32 image id of the emissive mip (if SQW_SURF_EMISSIVE) */
	float emissive_light[3];        /* emissive_texture.1 */
	uint32_t surface_pixels_offset; /* in pixels of the view's colour type (DrawPolygonData, renderer_structs.rs:86) */
	uint32_t reserved;
} sqw_surface_job;

#define SQW_SHADOW_NONE 0
#define SQW_SHADOW_CUBE 1
#define SQW_SHADOW_PROJECTOR 2

/* DynamicLightWithShadow (light.rs:4-11) with the shadow map flattened to an offset into the call's
 * shadow_maps array.  Cube maps store their six sides consecutively (size*size floats each, side order of
 * light.rs:13-22).  80 bytes. */
typedef struct sqw_light
{
	float position[3];
	float radius;
	float inv_square_radius;
	float color[3];
	uint32_t shadow_type; /* SQW_SHADOW_* */
	uint32_t shadow_map_size;
	uint32_t shadow_map_offset; /* in floats */
	float basis_x[3];           /* ProjectorShadowMap (light.rs:55-62) */
	float basis_y[3];
	float basis_z[3];
} sqw_light;

/* Build all surfaces of one view.  Replaces the body of PartialRenderer::build_polygons_surfaces.
 * Surfaces stay device-resident in the view's surface pool at job.surface_pixels_offset.
 * total_surface_pixels = num_visible_surfaces_pixels (partial_renderer.rs:1449): the pool grows to hold it. */
int sqw_build_surfaces(
	sqw_view* view,
	const sqw_surface_job* jobs,
	uint32_t num_jobs,
	const sqw_light* lights,
	uint32_t num_lights,
	const float* shadow_maps,
	uint64_t shadow_maps_len,
	const float cam_pos[3],
	uint64_t total_surface_pixels);

/* mix_surface_with_texture (surfaces.rs:854-971) applied to the whole colour buffer of a view that serves as a portal /
 * mirror texture: overlay of the portal's glass texture with its blending mode and light (partial_renderer.rs:2212-2271).
 * image_id: SQW_IMAGE_COLOR32 mip of the overlay; surface_tc_min = PortalInfo::tc_min; blending_mode = SQW_BLEND_*
 * (declared below). */
int sqw_view_mix_texture(sqw_view* view, uint32_t image_id, const int32_t surface_tc_min[2], int blending_mode, const float light[3]);

/* ---- shadow maps --------------------------------------------------------------- */
/* The depth rasteriser that fills the shadow maps the surface stage samples: DepthRasterizer::fill_polygon
 * (rasterizer.rs:1295-1474) as driven by DepthRenderer::draw_map (depth_renderer.rs:47-67) from
 * PartialRenderer::build_shadow_maps (partial_renderer.rs:524-590).  The host keeps doing the visibility flood, culling,
 * clipping and the biased depth equation (depth_renderer.rs:214-286) per light / cube side and records the fill calls. */
#define SQW_DEPTH_MAP_CLEAR 1u            /* fill the map with clear_value first (the reference keeps stale texels where nothing is drawn) */
#define SQW_DEPTH_MAP_PROJECTOR_CIRCLE 2u /* make_shadow_map_circle afterwards (light.rs:133-157; square maps) */

typedef struct sqw_depth_map /* 32 bytes */
{
	uint32_t offset; /* in floats into the frame's shadow map array (sqw_light::shadow_map_offset + side * size * size) */
	uint32_t width, height;
	uint32_t first_cmd, num_cmds; /* its commands: consecutive, in submission order */
	uint32_t flags;               /* SQW_DEPTH_MAP_* */
	float clear_value;
	uint32_t reserved;
} sqw_depth_map;

typedef struct sqw_depth_cmd /* one DepthRasterizer::fill_polygon::<DEPTH_TEST> call, 32 bytes */
{
	uint32_t num_vertices;  /* 3..24 */
	uint32_t vertex_offset; /* into polygon_points of the call */
	uint32_t depth_test;    /* 0: overwrite (world polygons, back to front); 1: max with the stored value (submodels) */
	float depth_eq[3];      /* DepthEquation with the bias and scale of depth_renderer.rs:226-237 applied */
	uint32_t reserved[2];
} sqw_depth_cmd;

/* ---- S2: rasterisation ---------------------------------------------------------------------------------- */

#define SQW_CMD_POLYGON 0  /* Rasterizer::fill_polygon  rasterizer.rs:35  */
#define SQW_CMD_TRIANGLE 1 /* Rasterizer::fill_triangle rasterizer.rs:731 */

#define SQW_BLEND_NONE 0 /* common/material.rs:101-116 */
#define SQW_BLEND_AVERAGE 1
#define SQW_BLEND_ADDITIVE 2
#define SQW_BLEND_ALPHA_TEST 3
#define SQW_BLEND_ALPHA_BLEND 4

#define SQW_TC_FULL_PERSPECTIVE 0 /* rasterizer.rs:4-10 */
#define SQW_TC_LINE_Z_CORRECTION 1
#define SQW_TC_AFFINE 2

#define SQW_TEX_SURFACE 0 /* view surface pool, tex_ref = pixel offset (get_polygon_surface_data, partial_renderer.rs:3777) */
#define SQW_TEX_IMAGE 1   /* image table, tex_ref = image id (sky side, model/decal/sprite TextureLite mip) */
#define SQW_TEX_VIEW 2    /* colour buffer of another view of the same ctx, tex_ref = view index (portal textures, partial_renderer.rs:3720) */

/* PolygonPointProjected (rasterizer.rs:1484-1489): 16.16 fixed point. */
typedef struct sqw_polygon_point
{
	int32_t x, y;
} sqw_polygon_point;

/* TrianglePointProjected (rasterizer.rs:1491-1498). 28 bytes. */
typedef struct sqw_triangle_point
{
	int32_t x, y;
	int32_t tc[2];
	int32_t light[3];
} sqw_triangle_point;

/* One recorded Rasterizer::fill_polygon / fill_triangle call, in submission (painter's) order. 64 bytes.
 * The host still does project_and_clip_polygon, mode selection and the z_scale renormalisation
 * (draw_polygon, partial_renderer.rs:4011-4142); the command carries the final arguments. */
typedef struct sqw_draw_cmd
{
	uint8_t kind;     /* SQW_CMD_* */
	uint8_t blend;    /* SQW_BLEND_* */
	uint8_t tc_mode;  /* SQW_TC_* (polygons) */
	uint8_t tex_kind; /* SQW_TEX_* */
	uint32_t num_vertices;  /* polygons: 3..24; triangles: 3 */
	uint32_t vertex_offset; /* element index into polygon_points / triangle_points */
	uint32_t tex_ref;
	int32_t tex_size[2]; /* TextureInfo (rasterizer.rs:1500-1503) */
	float depth_eq[3];   /* DepthEquation {d_inv_z_dx, d_inv_z_dy, k} (equations.rs:5-11) */
	float tc_eq[6];      /* TexCoordEquation {d_tc_dx[2], d_tc_dy[2], k[2]} (equations.rs:14-19) */
	uint32_t reserved;
} sqw_draw_cmd;

/* Draw a command list into the view's colour buffer with the clip rect of the single-threaded reference
 * (whole surface, partial_renderer.rs:1207-1233).  clear_color: NULL = keep the buffer, else the buffer is
 * first filled with this raw pixel value (draw_background, partial_renderer.rs:249-256; 4 or 8 bytes are read). */
int sqw_draw(
	sqw_view* view,
	const sqw_draw_cmd* cmds,
	uint32_t num_cmds,
	const sqw_polygon_point* polygon_points,
	uint32_t num_polygon_points,
	const sqw_triangle_point* triangle_points,
	uint32_t num_triangle_points,
	const void* clear_color);

/* ---- S3: post-processing --------------------------------------------------------------------------------- */

/* Arguments of perform_postprocessing plus the PostprocessorConfig fields it reads (postprocessor_config.rs:5-93). */
typedef struct sqw_post_params
{
	float exposure;          /* Postprocessor::current_exposure (postprocessor.rs:15) */
	float color_modulate[3]; /* FrameInfo::color_modulate */
	float bloom_sigma;       /* 0 disables bloom (postprocessor.rs:86) */
	float bloom_scale;
	uint32_t bloom_buffer_scale_log2; /* clamped to [1,3] (postprocessor.rs:731-734) */
	uint32_t linear_bloom_filter;
} sqw_post_params;

/* Tonemap (+bloom) the 64-bit view into 32-bit pixels.  dst may be NULL (picture stays on the device, see
 * sqw_read_pixels32).  dst_pitch is in pixels (SurfaceInfo::pitch).  avg_color_out = [B,G,R] average colour in
 * ColorVec lane order (postprocessor.rs:176-179 or :840-843) so the host can keep running update_exposure. */
int sqw_postprocess(sqw_view* view, const sqw_post_params* params, uint32_t* dst, size_t dst_pitch, float avg_color_out[3]);

/* ---- readback --------------------------------------------------------------------------------------------- */

/* Copy the view's colour buffer (4 or 8 bytes per pixel) to host; dst_pitch in pixels. Synchronises. */
int sqw_read_pixels(sqw_view* view, void* dst, size_t dst_pitch);
/* Upload host pixels (4 or 8 bytes per pixel) into the view's colour buffer; src_pitch in pixels.  This is the
 * write half of Postprocessor::get_hdr_buffer (postprocessor.rs:56): a host that only replaces S3 renders into its
 * own buffer and hands it over here. */
int sqw_write_pixels(sqw_view* view, const void* src, size_t src_pitch);
/* Copy the post-processed 32-bit picture of a 64-bit view. Synchronises. */
int sqw_read_pixels32(sqw_view* view, uint32_t* dst, size_t dst_pitch);
/* Asynchronous variants: the copy is queued behind the view's rendering and completes by the next
 * sqw_view_synchronize (raw copies also by any other synchronising call on the view).  dst should be page-locked memory, otherwise CUDA stages
 * the copy.  With SQW_CTX_CONCURRENT_VIEWS the read-back of one view overlaps the rendering of the others. */
int sqw_read_pixels_async(sqw_view* view, void* dst, size_t dst_pitch);
int sqw_read_pixels32_async(sqw_view* view, uint32_t* dst, size_t dst_pitch);
/* Switch SQW_CTX_COMPRESSED_READBACK on or off for read-backs started from now on (pictures on their way are completed first). */
int sqw_set_compressed_readback(sqw_ctx* ctx, int enable);
/* out[0] = compressed read-backs of this view completed so far, out[1] = the bytes they moved over the bus (run words + one count
 * word per row). */
int sqw_view_get_readback_stats(const sqw_view* view, uint64_t out[2]);
/* Copy `count` pixels of the surface pool starting at `offset` (both in pixels) — test/inspection hook. */
int sqw_read_surfaces(sqw_view* view, void* dst, uint64_t offset, uint64_t count);
/* Wait for everything issued for this view. */
int sqw_view_synchronize(sqw_view* view);

/* ---- recorded frames (batched views) -------------------------------------------------------- */
/* A recorded frame holds device copies of one view's job list, command list and post params so that it can be
 * replayed without host->device traffic.  This is what the view-batch workload and the device-resident bench
 * number use; sqw_build_surfaces / sqw_draw / sqw_postprocess above are "record + submit one stage". */
typedef struct sqw_frame sqw_frame;

int sqw_frame_create(sqw_view* view, sqw_frame** out_frame);
void sqw_frame_destroy(sqw_frame* frame);
int sqw_frame_set_surfaces(
	sqw_frame* frame,
	const sqw_surface_job* jobs,
	uint32_t num_jobs,
	const sqw_light* lights,
	uint32_t num_lights,
	const float* shadow_maps,
	uint64_t shadow_maps_len,
	const float cam_pos[3],
	uint64_t total_surface_pixels);
int sqw_frame_set_draw(
	sqw_frame* frame,
	const sqw_draw_cmd* cmds,
	uint32_t num_cmds,
	const sqw_polygon_point* polygon_points,
	uint32_t num_polygon_points,
	const sqw_triangle_point* triangle_points,
	uint32_t num_triangle_points,
	const void* clear_color);
int sqw_frame_set_post(sqw_frame* frame, const sqw_post_params* params);
/* Record the shadow-map fills of the frame.  They run (SQW_STAGE_SHADOW_MAPS) before the surface stage and write into the
 * frame's shadow map array: call sqw_frame_set_surfaces with shadow_maps = NULL and shadow_maps_len = the array length to
 * have that array live on the device only (no upload), or with host data to pre-fill it. */
int sqw_frame_set_shadow_map_draw(
	sqw_frame* frame,
	const sqw_depth_map* maps,
	uint32_t num_maps,
	const sqw_depth_cmd* cmds,
	uint32_t num_cmds,
	const sqw_polygon_point* polygon_points,
	uint32_t num_polygon_points);
/* Copy `count` floats of the frame's shadow map array starting at `offset` (test / inspection hook; synchronises). */
int sqw_frame_read_shadow_maps(sqw_frame* frame, float* dst, uint64_t offset, uint64_t count);
/* Launch surfaces -> raster -> post for the recorded frame on the context's stream (asynchronous).
 * stages: bit0 surfaces, bit1 raster, bit2 post. */
#define SQW_STAGE_SURFACES 1u
#define SQW_STAGE_RASTER 2u
#define SQW_STAGE_POST 4u
#define SQW_STAGE_SHADOW_MAPS 8u /* runs first */
#define SQW_STAGE_ALL 15u
int sqw_frame_submit(sqw_frame* frame, uint32_t stages);

/* Fork / join between the context's stream and the private streams of its views (SQW_CTX_CONCURRENT_VIEWS; no-ops
 * otherwise): frames submitted between the two calls start after everything queued on the context's stream before
 * sqw_batch_begin, and work queued on the context's stream after sqw_batch_end starts after those frames.  A caller that
 * passed its own stream to sqw_create brackets a batch of independent views with these (e.g. to time it with events). */
int sqw_batch_begin(sqw_ctx* ctx);
int sqw_batch_end(sqw_ctx* ctx);

/* Statistics of the last submit of this frame (units of the texels/s metric and of the roofline). */
typedef struct sqw_frame_stats
{
	uint64_t surface_texels;      /* S   = sum of surface_size[0]*surface_size[1] */
	uint64_t emissive_texels;     /* S_e */
	uint64_t lightmap_bytes;      /* L   */
	uint64_t num_spans;           /* rows of all commands */
	uint32_t num_kernel_launches; /* kernels launched by one submit */
	uint32_t num_surface_variants;
	uint32_t num_tiles;
	uint32_t reserved;
} sqw_frame_stats;
int sqw_frame_get_stats(const sqw_frame* frame, sqw_frame_stats* out);

/* ---- per-kernel timing (measurement support) ---------------------------------------------------------------- */
/* Between sqw_profile_begin and sqw_profile_end every kernel launch of this context is bracketed by CUDA events on the
 * context's stream (the reference's PerformanceCounter::run_with_measure around each stage, performance_counter.rs:39).
 * sqw_profile_end synchronises and returns the accumulated device time and launch count per kernel kind. */
#define SQW_KERNEL_SURFACE_BUILD 0
#define SQW_KERNEL_RASTER_SETUP 1
#define SQW_KERNEL_RASTER_BIN 2 /* unused since binning runs as tail CTAs of the set-up grid */
#define SQW_KERNEL_RASTER_TILE 3
#define SQW_KERNEL_TONEMAP 4
#define SQW_KERNEL_BLOOM_DOWNSCALE 5
#define SQW_KERNEL_BLOOM_BLUR 6
#define SQW_KERNEL_TONEMAP_BLOOM 7
#define SQW_KERNEL_RASTER_TILE_TONEMAP 8 /* tile kernel with the fused tonemapping epilogue */
#define SQW_KERNEL_MIX_TEXTURE 9
#define SQW_KERNEL_DEPTH_SETUP 10
#define SQW_KERNEL_DEPTH_RESOLVE 11
#define SQW_KERNEL_MODELS_PREPARE 12 /* model_bones + model_vertices + model_triangles kernels of one sqw_models_prepare */
#define SQW_KERNEL_POLYGONS_PREPARE 13 /* the kernels of sqw_polygons_prepare and sqw_polygons_draw */
#define SQW_NUM_KERNEL_KINDS 14
typedef struct sqw_kernel_times
{
	double ms[SQW_NUM_KERNEL_KINDS];
	uint64_t launches[SQW_NUM_KERNEL_KINDS];
} sqw_kernel_times;
int sqw_profile_begin(sqw_ctx* ctx);
int sqw_profile_end(sqw_ctx* ctx, sqw_kernel_times* out);

/* Device self-test of the arithmetic shortcuts the kernels take (synchronises): exhaustive comparison of the reciprocal-based
 * `2^31 / (inv_z >> 8)` of the full-perspective pixel loop (rasterizer.rs:430) with the plain u32 division for every
 * divisor below 2^24, and 2^26 hashed + edge-case operand pairs of the double-precision-assisted fixed16_div (fixed_math.rs:56)
 * against the plain 64-bit division.  SQW_OK, or SQW_ERR_CUDA with the mismatch count in sqw_last_error(). */
/* ---- triangle-model preparation on the GPU ----------------------------------------------------------
 * Replaces the body of PartialRenderer::build_dynamic_models_buffers (partial_renderer.rs:1122-1190): per visible mesh,
 * animate_and_transform_triangle_mesh_vertices (triangle_models_rendering.rs:7-203), reject_triangle_model_back_faces (:327-345)
 * and sort_model_triangles (:347-364).  Meshes are uploaded once (S0); per frame the host hands one sqw_model_instance per
 * visible mesh (matrices, animation point, light cube: per-MODEL scalar work that stays on the host). */
#define SQW_MESH_NON_ANIMATED 0      /* VertexData::NonAnimated      (triangle_model.rs:66) */
#define SQW_MESH_VERTEX_ANIMATED 1   /* VertexData::VertexAnimated   (MD3) */
#define SQW_MESH_SKELETON_ANIMATED 2 /* VertexData::SkeletonAnimated (IQM) */
typedef struct sqw_model_mesh
{
	uint32_t kind; /* SQW_MESH_* */
	uint32_t num_vertices, num_triangles, num_frames, num_bones;
	uint32_t reserved;
	/* NON_ANIMATED: [num_vertices] x {position[3], normal[3], tex_coord[2]} f32 (VertexNonAnimated);
	 * VERTEX_ANIMATED: [num_frames][num_vertices] x {position[3], normal[3]} f32 (VertexAnimatedVertexVariable);
	 * SKELETON_ANIMATED: [num_vertices] x {tex_coord[2], position[3], normal[3] f32, bone_index[4], weight[4] u8} = 40 bytes */
	const void* vertices;
	const float* tex_coords;      /* VERTEX_ANIMATED: [num_vertices][2] (VertexAnimatedVertexConstant), else NULL */
	const uint16_t* triangles;    /* [num_triangles][3] */
	const uint32_t* bone_parents; /* SKELETON_ANIMATED: [num_bones], >= num_bones = no parent; bones sorted in hierarchy order */
	const float* frame_bones;     /* SKELETON_ANIMATED: [num_frames][num_bones][16] column-major 4x4, or NULL (no animation) */
} sqw_model_mesh;
typedef struct sqw_model_instance
{
	uint32_t mesh;      /* id returned by sqw_upload_model_mesh */
	uint32_t frames[2]; /* AnimationPoint (frame_info.rs:71-77) */
	float lerp;
	float model_matrix[16];      /* column-major */
	float model_view_matrix[16]; /* VisibleDynamicMeshInfo::camera_matrices.view_matrix */
	float tc_scale[2];           /* size of the texture mip chosen for the model */
	float tc_shift[2];           /* TriangleModel::tc_shift */
	float light_cube[6][3];      /* ModelLightData (triangle_models_rendering.rs:684-699) */
	float light_vector_scaled[2][3];
	float light_color[2][3];
	uint32_t vertex_offset;   /* first output vertex (in vertices) */
	uint32_t triangle_offset; /* first output triangle; the instance owns num_triangles slots of its mesh */
	uint32_t bone_offset;     /* filled by the library */
	uint32_t sort_offset;     /* filled by the library */
} sqw_model_instance;
typedef struct sqw_model_vertex /* ModelVertex3d (renderer_structs.rs:311-318) */
{
	float pos[3];
	float tc[2];
	float light[3];
} sqw_model_vertex;
/* Upload one mesh; returns its id (>= 0) or an error.  The library copies. */
int sqw_upload_model_mesh(sqw_ctx* ctx, const sqw_model_mesh* mesh);
/* Prepare every instance on the view's stream: transformed, lit vertices at vertex_offset, front-facing triangles sorted back to
 * front (largest z first; equal keys keep the mesh order) at triangle_offset, their count per instance. */
int sqw_models_prepare(sqw_view* view, const sqw_model_instance* instances, uint32_t num_instances);
/* Read the results of the last sqw_models_prepare back (parity tests; a host that still clips and emits the triangles itself).
 * Any pointer may be NULL.  counts: [num_instances]. */
int sqw_models_read(sqw_view* view, sqw_model_vertex* vertices, uint64_t num_vertices, uint16_t* triangles, uint64_t num_triangles, uint32_t* counts, uint32_t num_instances);

/* ---- world-polygon preparation on the GPU (the geometry half of the host front end) ------------------
 * Replaces the loop body of PartialRenderer::prepare_polygons_surfaces (partial_renderer.rs:1327-1360) for the polygons of the
 * visible leafs: prepare_polygon_surface (:1464-1638) — back-face cull, material draw flag, vertex transform, z-near and leaf
 * octagon clipping (renderer_utils.rs:10-58, clipping.rs), sky bounds (:1511-1529), DepthEquation / TexCoordEquation
 * (equations.rs:22-75), calculate_mip (renderer_utils.rs:201-242), the surface rectangle (:1558-1615), the running
 * surface_pixels_offset (:1617-1618, a prefix sum in candidate order) and the shift of the texture-coordinate equation (:1629-1635).
 * The visibility flood stays on the host (map_visibility_calculator.rs): it hands over the (polygon, leaf clip planes) pairs.
 * The map's polygons and vertices are uploaded once (S0). */
typedef struct sqw_map_polygon /* Polygon, bsp_map_compact.rs:58-75, 80 bytes (the .sqwm polygons lump as it is) */
{
	uint32_t first_vertex, num_vertices;
	float plane[4];                 /* Plane {vec, dist} */
	float tex_coord_equation[2][4]; /* two Plane */
	int32_t tex_coord_min[2];
	int32_t tex_coord_max[2];
	uint32_t lightmap_data_offset;
	uint32_t texture; /* material index */
} sqw_map_polygon;
#define SQW_MATERIAL_DRAW 1u   /* Material::draw (material.rs) */
#define SQW_MATERIAL_SKYBOX 2u /* Material::skybox is some: the polygon only widens the sky window */
#define SQW_MATERIAL_BLEND_SHIFT 4 /* bits 4-6: Material::blending_mode as SQW_BLEND_* (sqw_polygons_draw) */
/* vertices: [num_vertices][3] f32 (BSPMap::vertices); material_flags: [num_materials] SQW_MATERIAL_*.  The library copies. */
int sqw_upload_map_geometry(sqw_ctx* ctx, const sqw_map_polygon* polygons, uint32_t num_polygons, const float* vertices, uint32_t num_vertices, const uint8_t* material_flags, uint32_t num_materials);
typedef struct sqw_polygon_candidate
{
	uint32_t polygon;  /* index into the uploaded polygons */
	uint32_t clip_set; /* index into clip_planes: the eight planes of the leaf's ClippingPolygon (clipping_polygon.rs:150-163) */
} sqw_polygon_candidate;
typedef struct sqw_polygons_frame
{
	float planes_matrix[16]; /* CameraMatrices::planes_matrix, column-major (matrix.rs:4-11) */
	float view_matrix[16];   /* CameraMatrices::view_matrix */
	float mip_bias;          /* PartialRenderer::mip_bias (:3786-3804) */
	uint32_t reserved[3];
} sqw_polygons_frame;
#define SQW_PREPARED_VISIBLE 1u /* a surface is built and the polygon drawn (DrawPolygonData::visible_frame == current frame) */
#define SQW_PREPARED_SKY 2u     /* sky polygon that passed culling and clipping: contributed to the sky bounds */
typedef struct sqw_prepared_polygon /* the per-frame half of DrawPolygonData (renderer_structs.rs:75-91), 64 bytes */
{
	uint32_t flags; /* SQW_PREPARED_* */
	uint32_t mip;
	float depth_equation[3];
	float tex_coord_equation[6]; /* d_tc_dx[2], d_tc_dy[2], k[2]; scaled to the mip and shifted to the surface rectangle */
	int32_t surface_tc_min[2];
	uint32_t surface_size[2];
	uint32_t surface_pixels_offset;
} sqw_prepared_polygon;
typedef struct sqw_polygons_totals
{
	uint32_t num_visible;         /* current_frame_visible_polygons.len() */
	uint32_t has_sky;             /* current_sky is some (:1336, :1511-1529) */
	uint64_t num_surfaces_pixels; /* num_visible_surfaces_pixels (:395) */
	float sky_bounds[8];          /* ClippingPolygon: x, y, x_plus_y, x_minus_y as {min, max} */
	uint32_t sky_material;        /* material of the first sky polygon in candidate order */
	uint32_t reserved;
} sqw_polygons_totals;
/* clip_planes: [num_clip_sets][8][3] f32.  Runs on the view's stream; results stay on the device (sqw_polygons_read). */
int sqw_polygons_prepare(sqw_view* view, const sqw_polygons_frame* frame, const sqw_polygon_candidate* candidates, uint32_t num_candidates, const float* clip_planes, uint32_t num_clip_sets);
/* Read the results of the last sqw_polygons_prepare back: out[num] in candidate order (NULL to skip), totals (NULL to skip). */
int sqw_polygons_read(sqw_view* view, sqw_prepared_polygon* out, uint32_t num, sqw_polygons_totals* totals);
/* draw_polygon (partial_renderer.rs:4011-4142) for world polygons prepared by the last sqw_polygons_prepare of this view, in the
 * order the host's BSP walk visits them (draw_tree / draw_leaf, :2450-2565; the walk itself stays on the host): `order` holds
 * candidate indices.  Polygons that are not visible produce nothing.  For each of the others: clipping against the leaf's octagon,
 * 16.16 vertices, the choice between affine / line-z-corrected / full-perspective interpolation (renderer_utils.rs:60-197) and the
 * power-of-two renormalisation of the equations (:4106-4109) -> packed sqw_draw_cmd[] (tex_kind SQW_TEX_SURFACE, tex_ref =
 * surface_pixels_offset, blend = the material's blending mode) and sqw_polygon_point[] in draw order, on the device. */
int sqw_polygons_draw(sqw_view* view, const uint32_t* order, uint32_t num_order);
/* Per-material and per-polygon constants of job building (the closure of build_polygons_surfaces, partial_renderer.rs:1682-1819),
 * uploaded once per map for sqw_frame_set_surfaces_from_polygons. */
#define SQW_MAT_LIGHT 1u                 /* Material::light */
#define SQW_MAT_HAS_DIFFUSE 2u           /* a diffuse texture exists */
#define SQW_MAT_HAS_NORMAL_MAP 4u        /* Texture::has_normal_map (textures.rs:14) */
#define SQW_MAT_HAS_NON_ONE_ROUGHNESS 8u /* Texture::has_non_one_roughness */
#define SQW_MAT_IS_METAL 16u             /* Texture::is_metal */
#define SQW_MAT_HAS_EMISSIVE 32u         /* an emissive layer exists */
typedef struct sqw_map_material /* 64 bytes */
{
	uint32_t flags;              /* SQW_MAT_* */
	uint32_t texture_images[5];  /* image ids of the TextureElement mips */
	uint32_t emissive_images[5]; /* image ids of the emissive mips (Color32) */
	float emissive_light[3];
	int32_t texture_shift[2];    /* MapMaterialsProcessor::get_texture_shift (scrolling); 0 when animations are off */
} sqw_map_material;
/* polygon_basis: [num uploaded polygons][12] f32 = PolygonBasisVecs {u, v, start, normal} for mip 0 (surfaces.rs:7-47), computed once
 * per map by the host like the reference does (partial_renderer.rs:86-100).  num_materials must match sqw_upload_map_geometry. */
int sqw_upload_map_materials(sqw_ctx* ctx, const sqw_map_material* materials, uint32_t num_materials, const float* polygon_basis, uint32_t num_polygons);
/* The surface half of a recorded frame from the records of the last sqw_polygons_prepare of the frame's view: one sqw_surface_job
 * per visible polygon in candidate order (texture shift, basis vectors at the polygon's mip, lightmap shift / size / scale,
 * emissive layer, and the dynamic lights that reach the polygon — the whole closure of build_polygons_surfaces,
 * partial_renderer.rs:1682-1819) and the band lists of the surface kernel, built on the device.  Dynamic lights: `lights` as in
 * sqw_build_surfaces, and for every clip set (= visible leaf) of the prepare call the lights located in that leaf
 * (DynamicObjectsIndex::get_leaf_objects): clip_set_lights[clip_set_light_offsets[s] .. clip_set_light_offsets[s + 1]) are indices into
 * lights, tested in that order by polygon_is_affected_by_light (:3877-3985), at most SQW_MAX_POLYGON_LIGHTS kept (:1705-1734).  The
 * shadow maps live on the device (filled by the frame's shadow map stage); shadow_maps_len is their array's length in floats.
 * num_lights = 0: the static world.  The host reads the totals back to size the surface pool. */
int sqw_frame_set_surfaces_from_polygons(
	sqw_frame* frame, const float cam_pos[3], int use_directional_lightmaps, const sqw_light* lights, uint32_t num_lights, const uint32_t* clip_set_light_offsets,
	const uint16_t* clip_set_lights, uint32_t num_clip_sets, uint64_t shadow_maps_len);
/* The draw half of a recorded frame straight from the device lists of the last sqw_polygons_draw of the frame's view: what
 * sqw_frame_set_draw derives on the host from host lists (row and column ranges, span-table offsets, 32-scan-line chunks, ordered
 * per-band command lists) is derived by set-up kernels; the commands and vertices never visit the host, which only reads four
 * totals back to size the view's scratch.  clear_color as in sqw_draw.  The surfaces the commands reference must have been sized
 * (sqw_frame_set_surfaces / sqw_build_surfaces) before. */
int sqw_frame_set_draw_from_polygons(sqw_frame* frame, const void* clear_color);
/* Read the lists of the last sqw_polygons_draw back (NULL to skip); counts_out = {commands, vertices}. */
int sqw_polygons_draw_read(sqw_view* view, sqw_draw_cmd* cmds, uint32_t cmds_capacity, sqw_polygon_point* points, uint32_t points_capacity, uint32_t counts_out[2]);

int sqw_self_test(sqw_ctx* ctx);

/* Switch SQW_CTX_FAST_SURFACE_NUMERICS on or off for frames submitted from now on (the flag of sqw_create is the initial
 * state).  Lets one context serve a bit-exact consumer and a throughput consumer, and bench.py time both on the same data. */
int sqw_set_fast_surface_numerics(sqw_ctx* ctx, int enable);

/* Average colour produced by the last post stage of this view (synchronises). */
int sqw_view_get_average_color(sqw_view* view, float avg_color_out[3]);

#ifdef __cplusplus
}
#endif
#endif /* SQW_CUDA_H */
