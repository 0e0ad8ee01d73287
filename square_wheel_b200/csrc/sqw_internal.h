// Internal structures shared by the C-ABI implementation (sqw_api.cu) and the kernels.
#pragma once
#include "../../include/sqw_cuda.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqw
{

// Device image table entry (one mip of one texture), filled by sqw_upload_image.
struct ImageDev
{
	const void* pixels;
	uint32_t width, height;
};

// ---- stage 1 ------------------------------------------------------------------------------------------------
// One CTA of surface_build_kernel processes one band: up to SQW_SURF_MAX_BAND_ROWS consecutive rows of one job holding
// about SQW_SURF_BAND_TEXELS texels.
#define SQW_SURF_MAX_BAND_ROWS 32
#ifndef SQW_SURF_BAND_TEXELS
#define SQW_SURF_BAND_TEXELS 2048
#endif
#define SQW_SURF_THREADS 256
// Jobs of at most SQW_SURF_WARP_JOB_TEXELS texels are cut into WARP bands (one warp per band): about
// SQW_SURF_WARP_BAND_TEXELS texels each (fewer, down to SQW_SURF_WARP_BAND_TEXELS_MIN, in frames with few small texels), rows x
// lightmap columns <= SQW_SURF_WARP_LM_ELEMENTS.
#ifndef SQW_SURF_WARP_JOB_TEXELS
#define SQW_SURF_WARP_JOB_TEXELS 1024
#endif
#ifndef SQW_SURF_WARP_BAND_TEXELS
#define SQW_SURF_WARP_BAND_TEXELS 256
#endif
#ifndef SQW_SURF_WARP_BAND_TEXELS_MIN
#define SQW_SURF_WARP_BAND_TEXELS_MIN 64
#endif
#define SQW_SURF_WARP_LM_ELEMENTS 80
struct SurfaceBand
{
	uint32_t job;
	uint32_t row0;
	uint32_t num_rows; // 1..SQW_SURF_MAX_BAND_ROWS, never past the last row of the surface
	uint32_t variant;  // bit0 directional lightmap, bit1 normal map, bits2-3 specular type, 16 = emissive-only
};

// Surface kernel variant = (colour width, directional lightmap, dynamic lights, normal map, specular type);
// mirrors the const-generic dispatch chain of surfaces.rs:124-423.
struct SurfaceVariant
{
	uint8_t color64;
	uint8_t directional;
	uint8_t dynamic_lights;
	uint8_t normal_map;
	uint8_t specular; // 0 none, 1 dielectric, 2 metal
	uint8_t emissive_only; // job without diffuse layer: only mix_surface_with_texture(BlendingMode::None)
};

struct SurfaceLaunch
{
	const sqw_surface_job* jobs;   // device
	const SurfaceBand* bands;      // device, this variant's slice
	uint32_t num_bands;
	uint32_t num_cta_bands;        // the first num_cta_bands bands are CTA bands, the others warp bands (small surfaces)
	const ImageDev* images;        // device
	const float* lightmaps;        // device
	const sqw_directional_lightmap_element* directional_lightmaps; // device
	const sqw_light* lights;       // device
	const float* shadow_maps;      // device
	float cam_pos[3];
	void* surfaces;                // device pool base
};
// One grid for all bands of the launch; `dynamic_lights` selects the kernel built with the dynamic-light loop.
cudaError_t launch_surface_build(bool color64, bool dynamic_lights, const SurfaceLaunch& l, cudaStream_t stream);
// the same kernels built with hardware rcp / rsqrt and FMA (surface_kernels_fast.cu, SQW_CTX_FAST_SURFACE_NUMERICS)
cudaError_t launch_surface_build_fast(bool color64, bool dynamic_lights, const SurfaceLaunch& l, cudaStream_t stream);

cudaError_t launch_mix_texture(
	bool color64, void* surface, uint32_t size_u, uint32_t size_v, const int32_t tc_min[2], const void* texture, uint32_t tex_w, uint32_t tex_h, int blending_mode,
	const float light[3], cudaStream_t stream);

// ---- stage 2 ------------------------------------------------------------------------------------------------
#define SQW_TILE_W 128
#define SQW_TILE_H 8
// Colour sums (average colour for update_exposure) are accumulated with atomics into SQW_SUM_SLOTS separate 32-byte slots
// {B, G, R, pad} chosen by CTA index — thousands of same-address atomics would serialise in one L2 slice — and added
// up on the host when the average is read.
#define SQW_SUM_SLOTS 64
#define SQW_SUM_WORDS (SQW_SUM_SLOTS * 4)
#define SQW_SETUP_THREADS 128
#define SQW_TRI_REC_BYTES 192 /* per-triangle set-up record (both halves), raster_kernels.cu */
#define SQW_TONEMAP_LUT_SIZE 4096 /* channel values below this (16 x LDR white) are tonemapped by table look-up */
#define SQW_TILE_STAGE_HITS 4 /* commands whose span records a warp stages in shared memory at a time (x 8 block rows = 32 lanes) */
#define SQW_RASTER_THREADS 128 /* 4 warps, each owning a 32x8 pixel block of the 128x8 tile */
// Tile classes (tile_desc.w, written by the binning CTAs): which of the two tile kernels renders the tile.
#define SQW_TILE_CLASS_GENERAL 0u /* raster_tile_kernel: any command mix, painter's order per pixel */
#define SQW_TILE_CLASS_OPAQUE 1u  /* raster_tile_opaque_kernel: only polygons with BlendingMode::None */
#define SQW_GENERAL_LIST_HEADER 2u /* words before the first entry of the general-tile list */
#ifndef SQW_OPAQUE_STAGE_HITS
#define SQW_OPAQUE_STAGE_HITS 8
#endif
//   /* commands whose span records a warp of the opaque tile kernel stages at a time */
#ifndef SQW_OPAQUE_WARPS
#define SQW_OPAQUE_WARPS 8
#endif
//        /* warps per CTA of the opaque tile kernel: two 128x8 tiles side by side */
#ifndef SQW_OPAQUE_BLOCKS
#define SQW_OPAQUE_BLOCKS 4 /* adjacent 32x8 blocks a warp of the opaque tile kernel renders from one staged command list */
#endif
#ifndef SQW_OPAQUE_CTAS_PER_SM
#define SQW_OPAQUE_CTAS_PER_SM 2
#endif
#ifndef SQW_TILE_CTAS_PER_SM
#define SQW_TILE_CTAS_PER_SM 5 /* resident CTAs of the persistent tile kernel per SM (register and shared-memory budget) */
#endif

#define SQW_CMDF_Z_ONE_STEP 4u
// Per-command data prepared by the host part of sqw_frame_set_draw (row range, conservative column range,
// span-table offset, resolved texture pointer).  32 bytes.
struct CmdDev
{
	int16_t y0, y1;   // row range [y0, y1) inside the clip rect (rows that can own a span)
	int16_t bx0, bx1; // conservative pixel column range [bx0, bx1)
	uint32_t span_base;
	uint8_t kind, blend, tc_mode, tex_flags; // tex_flags bit0: texture elements are 8 bytes (else 4); bit1: `texture` is a byte offset into the surface pool;
	                                         // bit2 (SQW_CMDF_Z_ONE_STEP): full-perspective polygon whose inv_z >> 8 stays >= 2^11 at every pixel
	const void* texture;
	int32_t tex_width;
	int32_t tex_height;
};
static_assert(sizeof(CmdDev) == 32, "CmdDev must be 32 bytes");

// One scan-line of one command: everything fill_polygon_part* / fill_triangle_part derive per line before the
// pixel loop.  48 bytes, written by raster_setup_kernel, read (warp-uniform) by raster_tile_kernel.
//   polygon, full perspective: v[0] = span_inv_z, v[1] = span_d_inv_z, v[2..3] = span_tc, v[4..5] = span_d_tc   (rasterizer.rs:345-422)
//   polygon, line-z / affine : v[2..3] = tc (16.16), v[4..5] = d_tc                                           (:540-556, :697-703)
//   triangle                 : v[0..1] = line_tc, v[2..3] = d_tc_dx, v[4..6] = line_light, v[7..9] = d_light_dx (:1205-1214)
struct SpanRec
{
	int16_t x0, x1; // pixel span [x0, x1); x0 >= x1 means empty
	int32_t v[11];
};
static_assert(sizeof(SpanRec) == 48, "SpanRec must be 48 bytes");

struct RasterLaunch
{
	const sqw_draw_cmd* cmds; // device copy of the submitted commands
	const CmdDev* cmd_dev;    // device
	uint32_t num_cmds;
	const sqw_polygon_point* polygon_points;
	const sqw_triangle_point* triangle_points;
	void* tri_recs;             // device, num_cmds x SQW_TRI_REC_BYTES: per-triangle set-up records (raster_triangle_setup_kernel)
	uint32_t num_triangle_cmds; // 0: no triangle set-up launch
	SpanRec* spans;                // device, total_spans entries
	const uint2* chunks; // device: {command index, first row relative to the command's y0} per 32 scan-lines of a command
	uint32_t num_chunks;
	uint32_t total_spans;
	const uint32_t* band_lists;   // device, ordered command indices per band (built by sqw_frame_set_draw)
	const uint32_t* band_offsets; // device, num_bands + 1
	uint32_t* tile_lists;         // device, ordered command indices per 128x8 tile: tile (band, col) owns the slots
	                              // [band_offsets[band] * tiles_x + col * N, + N), N = band list length
	uint32_t* general_tiles;      // device, per view (sqw_view::d_general_tiles): [0] number of general tiles (zeroed per frame), entries from [SQW_GENERAL_LIST_HEADER]
	uint4* tile_desc;             // device, {first slot, count, x0 | y0 << 16, -} per tile, written by the binning CTAs
	uint32_t num_bands;
	int32_t width, height, pitch;
	void* color_buffer;
	const void* surfaces_base; // view surface pool (CmdDev::texture is relative to it when flag bit 1 is set)
	int color64;
	int clear;
	uint64_t clear_color;
	// fused tonemapping epilogue (64-bit views, post stage without bloom): 32-bit picture, Reinhard scales, colour sums
	uint32_t* tonemap_dst; // nullptr = not fused
	float inv_scale_bgr[3];
	unsigned long long* color_sums;
	uint8_t* tonemap_lut; // device, 3 x SQW_TONEMAP_LUT_SIZE bytes (B, G, R), rebuilt by the set-up grid of every fused frame
	// number of commands that are not opaque polygons (triangles, blended / alpha-tested polygons): 0 = every tile is of class
	// SQW_TILE_CLASS_OPAQUE and the general tile kernel is not launched
	uint32_t num_general_cmds;
};
// span set-up + binning in one grid
cudaError_t launch_raster_setup(const RasterLaunch& l, cudaStream_t stream);
cudaError_t launch_raster_tiles(const RasterLaunch& l, cudaStream_t stream);
// exhaustive check of z_from_inv_z; *mismatches (device) must be zeroed by the caller
cudaError_t launch_z_from_inv_z_test(unsigned long long* mismatches, cudaStream_t stream);

// ---- triangle models (model_kernels.cu) -----------------------------------------------------------------------------
struct ModelMeshDev // device copy of one sqw_model_mesh
{
	uint32_t kind, num_vertices, num_triangles, num_frames, num_bones, pad;
	const void* vertices;
	const float* tex_coords;
	const uint16_t* triangles;
	const uint32_t* bone_parents;
	const float* frame_bones;
};
struct ModelLaunch
{
	const sqw_model_instance* instances; // device
	const ModelMeshDev* meshes;          // device table
	uint32_t num_instances;
	uint32_t max_vertices; // largest mesh among the instances (grid size)
	uint32_t max_bones;    // 0: no skeleton-animated instance
	float* bone_scratch;   // 25 floats per (instance, bone)
	unsigned long long* sort_scratch;
	float* vertices;       // out: 8 floats per vertex
	uint16_t* triangles;   // out
	uint32_t* counts;      // out
};
cudaError_t launch_models_prepare(const ModelLaunch& l, cudaStream_t stream);

// ---- device-side job and band building (draw_setup_kernels.cu) ---------------------------------------------------------
struct SurfaceSetupLaunch
{
	const sqw_map_polygon* polygons;         // device (uploaded map)
	const sqw_map_material* materials;       // device
	const float* polygon_basis;              // device, 12 floats per polygon
	const sqw_polygon_candidate* candidates; // device, of the last sqw_polygons_prepare
	const sqw_prepared_polygon* prepared;    // device
	uint32_t num_candidates;
	uint32_t use_directional;
	const sqw_light* lights;             // device, or NULL: static world only
	uint32_t num_lights;
	const uint32_t* set_light_offsets;   // device [num clip sets + 1]: the lights of every clip set's leaf ...
	const uint16_t* set_lights;          // ... as indices into lights
	uint32_t* visible_rank;  // scratch [num_candidates]
	sqw_surface_job* jobs;   // out [num visible]
	uint32_t* band_counts;   // scratch [num_candidates][4]: CTA / warp bands of a job without dynamic lights, CTA / warp bands of a job with
	uint32_t* band_base;     // scratch [num_candidates][4]
	SurfaceBand* bands;      // phase 2 out: the four kinds one after the other
	unsigned long long* totals; // [8]: visible, small texels, CTA bands, warp bands (no lights), surface texels, emissive texels, lightmap bytes, CTA | warp << 32 (with lights)
};
cudaError_t launch_surface_setup_phase1(const SurfaceSetupLaunch& l, cudaStream_t stream);
cudaError_t launch_surface_bands_count(const SurfaceSetupLaunch& l, uint32_t num_jobs, cudaStream_t stream);
cudaError_t launch_surface_setup_phase2(const SurfaceSetupLaunch& l, uint32_t num_jobs, cudaStream_t stream);

// ---- device-side frame set-up of the rasteriser (draw_setup_kernels.cu) ----------------------------------------------
struct DrawSetupLaunch
{
	const sqw_draw_cmd* cmds;         // device, packed, draw order
	const sqw_polygon_point* points;  // device
	uint32_t num_cmds;
	int32_t width, height;
	uint32_t num_bands;
	uint32_t bytes_per_pixel;         // 4 or 8
	uint64_t surface_pool_elements;   // capacity of the view's surface pool, in pixels
	CmdDev* cmd_dev;                  // out [num_cmds]
	uint32_t* rows;                   // out [num_cmds]: scan-lines of the command
	uint32_t* chunk_base;             // out [num_cmds]
	uint32_t* band_counts;            // out [num_bands]
	uint32_t* band_offsets;           // out [num_bands + 1]
	uint2* chunks;                    // phase 2 out
	uint32_t* band_list;              // phase 2 out
	uint32_t* totals;                 // [8]: scan-lines, chunks, band list entries, non-opaque commands, error flag
};
cudaError_t launch_draw_setup_phase1(const DrawSetupLaunch& l, cudaStream_t stream);
cudaError_t launch_draw_setup_phase2(const DrawSetupLaunch& l, cudaStream_t stream);

// ---- compressed read-back of the 32-bit picture ----------------------------------------------------------------------
cudaError_t launch_picture_rle(const uint32_t* ldr, uint32_t width, uint32_t height, uint32_t pitch, uint32_t* rows_out, uint32_t row_capacity, uint32_t* row_counts, cudaStream_t stream);
void decode_picture_rows(const uint32_t* rle, const uint32_t* counts, uint32_t row_capacity, uint32_t width, uint32_t row0, uint32_t row1, uint32_t* dst, size_t dst_pitch); // readback_decode.cu

// ---- world-polygon preparation -------------------------------------------------------------------
struct PolygonsTotalsDev
{
	float sky_bounds[8]; // x, y, x_plus_y, x_minus_y as {min, max}
	uint32_t first_sky_candidate;
	uint32_t num_visible;
	unsigned long long num_surfaces_pixels;
};
struct PolygonsLaunch
{
	const sqw_map_polygon* polygons; // device (uploaded map)
	const float* vertices;           // device
	const uint8_t* material_flags;   // device
	uint32_t num_materials;
	sqw_polygons_frame frame;
	const sqw_polygon_candidate* candidates; // device
	uint32_t num_candidates;
	const float* clip_planes; // device, [sets][8][3]
	sqw_prepared_polygon* out; // device
	PolygonsTotalsDev* totals; // device
};
cudaError_t launch_polygons_prepare(const PolygonsLaunch& l, cudaStream_t stream);
struct PolygonsDrawLaunch
{
	const sqw_map_polygon* polygons; // device (uploaded map)
	const float* vertices;
	const uint8_t* material_flags;
	uint32_t num_materials;
	sqw_polygons_frame frame;
	const sqw_polygon_candidate* candidates; // device, of the last sqw_polygons_prepare
	const float* clip_planes;
	const sqw_prepared_polygon* prepared;    // device, results of the last sqw_polygons_prepare
	const uint32_t* order;                   // device: candidate indices in draw order
	uint32_t num_order;
	sqw_draw_cmd* scratch_cmds;        // [num_order]
	sqw_polygon_point* scratch_points; // [num_order][24]
	uint2* offsets;                    // [num_order]
	sqw_draw_cmd* cmds;                // out, packed
	sqw_polygon_point* points;         // out, packed
	uint32_t* totals;                  // out: {commands, vertices}
};
cudaError_t launch_polygons_draw(const PolygonsDrawLaunch& l, cudaStream_t stream);

// ---- shadow maps ----------------------------------------------------------------------------------------------
struct DepthCmdDev // per fill call, prepared by sqw_frame_set_shadow_map_draw. 16 bytes
{
	int16_t y0, y1; // rows [y0, y1) of its map that can own a span
	uint32_t span_base;
	uint32_t map;
	uint32_t depth_test;
};
struct DepthSpan // one scan-line of one fill call. 16 bytes
{
	int16_t x0, x1;   // pixel span [x0, x1)
	float inv_z;      // value written to pixel x0 (rasterizer.rs:1456)
	float d_inv_z_dx; // added after every pixel (:1463)
	uint32_t pad;
};
struct DepthLaunch
{
	const sqw_depth_cmd* cmds;
	const DepthCmdDev* cmd_dev;
	const sqw_polygon_point* points;
	const sqw_depth_map* maps;
	const uint2* chunks; // {command, first row relative to y0} per 32 scan-lines of a command
	uint32_t num_chunks;
	const uint2* row_chunks; // {map, first row} per 32 rows of a map
	uint32_t num_row_chunks;
	DepthSpan* spans;
	float* shadow_maps;
};
cudaError_t launch_depth_setup(const DepthLaunch& l, cudaStream_t stream);
cudaError_t launch_depth_resolve(const DepthLaunch& l, cudaStream_t stream);

// ---- stage 3 ------------------------------------------------------------------------------------------------
struct PostLaunch
{
	const uint64_t* hdr;
	uint32_t width, height;
	uint32_t* dst; // device, pitch = width
	float inv_scale_bgr[3];
	unsigned long long* color_sums; // device, 3 x u64 (B, G, R)
	// bloom
	uint64_t* bloom0;
	uint64_t* bloom1;
	uint32_t bloom_w, bloom_h, bloom_scale_log2;
	int32_t blur_radius;
	int32_t kernel_h[33];
	int32_t kernel_v[33];
	int linear;
};
cudaError_t launch_tonemap(const PostLaunch& l, cudaStream_t stream);
cudaError_t launch_bloom_downscale(const PostLaunch& l, cudaStream_t stream);
cudaError_t launch_bloom_blur(const PostLaunch& l, cudaStream_t stream);
cudaError_t launch_tonemap_bloom(const PostLaunch& l, cudaStream_t stream);

} // namespace sqw
