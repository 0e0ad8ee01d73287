//! FFI binding of `libsqw_cuda.so` (include/sqw_cuda.h) for Square-Wheel @4820f17.
//!
//! Drop this file into `src/square_wheel/sqw_cuda.rs`, add `pub mod sqw_cuda;` to `src/square_wheel/mod.rs` and apply
//! `bindings/seams.patch`.  Every struct is `#[repr(C)]` and its size is asserted at the bottom against the `static_assert`s of
//! `square_wheel_b200/csrc/sqw_api.cu`; constants carry the names of the header.  No Rust toolchain exists in the image this
//! repository is built in: the C side of every entry point is exercised through ctypes (tests/), the Rust side is source a
//! maintainer compiles.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct sqw_ctx { _private: [u8; 0] }
#[repr(C)] pub struct sqw_view { _private: [u8; 0] }
#[repr(C)] pub struct sqw_frame { _private: [u8; 0] }

pub const SQW_OK: c_int = 0;
pub const SQW_ERR_INVALID_ARGUMENT: c_int = -1;
pub const SQW_ERR_CUDA: c_int = -2;
pub const SQW_ERR_OUT_OF_MEMORY: c_int = -3;
pub const SQW_ERR_NO_DEVICE: c_int = -4;

pub const SQW_IMAGE_TEXTURE_ELEMENT: c_int = 0; // textures.rs:20-27
pub const SQW_IMAGE_COLOR32: c_int = 1; // TextureLite / Image, 32-bit sky box sides
pub const SQW_IMAGE_COLOR64: c_int = 2; // SkyboxSideTexture<Color64>

pub const SQW_CTX_CONCURRENT_VIEWS: u32 = 1;
pub const SQW_CTX_FAST_SURFACE_NUMERICS: u32 = 2;

#[repr(C)]
pub struct sqw_map_desc
{
	pub lightmaps_data: *const f32, // BSPMap::lightmaps_data as [n][3] (bsp_map_compact.rs:27)
	pub directional_lightmaps_data: *const crate::common::bsp_map_compact::DirectionalLightmapElement, // already #[repr(C)], 40 bytes
	pub num_lightmap_elements: u64,
	pub max_images: u32,
	pub flags: u32, // SQW_CTX_*
}

pub const SQW_SURF_HAS_NORMAL_MAP: u32 = 1;
pub const SQW_SURF_HAS_NON_ONE_ROUGHNESS: u32 = 2;
pub const SQW_SURF_IS_METAL: u32 = 4;
pub const SQW_SURF_DIRECTIONAL_LIGHTMAP: u32 = 8;
pub const SQW_SURF_APPLY_DIFFUSE: u32 = 16;
pub const SQW_SURF_EMISSIVE: u32 = 32;
pub const SQW_MAX_POLYGON_LIGHTS: usize = 6; // partial_renderer.rs:1666

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_surface_job // 136 bytes: the arguments assembled in partial_renderer.rs:1682-1819
{
	pub basis_u: [f32; 3],
	pub basis_v: [f32; 3],
	pub basis_start: [f32; 3],
	pub basis_normal: [f32; 3],
	pub surface_size: [u32; 2],
	pub tc_start: [i32; 2],
	pub texture_image: u32,
	pub flags: u32,
	pub lightmap_size: [u32; 2],
	pub lightmap_scale_log2: u32,
	pub lightmap_tc_shift: [u32; 2],
	pub lightmap_offset: u32,
	pub num_lights: u32,
	pub lights: [u16; SQW_MAX_POLYGON_LIGHTS],
	pub emissive_image: u32,
	pub emissive_light: [f32; 3],
	pub surface_pixels_offset: u32,
	pub reserved: u32,
}

pub const SQW_SHADOW_NONE: u32 = 0;
pub const SQW_SHADOW_CUBE: u32 = 1;
pub const SQW_SHADOW_PROJECTOR: u32 = 2;

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_light // 80 bytes: DynamicLightWithShadow (light.rs:4-11), shadow map flattened to an offset
{
	pub position: [f32; 3],
	pub radius: f32,
	pub inv_square_radius: f32,
	pub color: [f32; 3],
	pub shadow_type: u32,
	pub shadow_map_size: u32,
	pub shadow_map_offset: u32, // in floats into the shadow map array of the call; cube sides consecutive
	pub basis_x: [f32; 3],      // ProjectorShadowMap (light.rs:55-62)
	pub basis_y: [f32; 3],
	pub basis_z: [f32; 3],
}

pub const SQW_DEPTH_MAP_CLEAR: u32 = 1;
pub const SQW_DEPTH_MAP_PROJECTOR_CIRCLE: u32 = 2;

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_depth_map // 32 bytes
{
	pub offset: u32,
	pub width: u32,
	pub height: u32,
	pub first_cmd: u32,
	pub num_cmds: u32,
	pub flags: u32,
	pub clear_value: f32,
	pub reserved: u32,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_depth_cmd // 32 bytes: one DepthRasterizer::fill_polygon::<DEPTH_TEST> call (rasterizer.rs:1314)
{
	pub num_vertices: u32,
	pub vertex_offset: u32,
	pub depth_test: u32,
	pub depth_eq: [f32; 3],
	pub reserved: [u32; 2],
}

pub const SQW_CMD_POLYGON: u8 = 0; // Rasterizer::fill_polygon  (rasterizer.rs:35)
pub const SQW_CMD_TRIANGLE: u8 = 1; // Rasterizer::fill_triangle (rasterizer.rs:731)
pub const SQW_BLEND_NONE: u8 = 0; // material.rs:101-116, in declaration order
pub const SQW_BLEND_AVERAGE: u8 = 1;
pub const SQW_BLEND_ADDITIVE: u8 = 2;
pub const SQW_BLEND_ALPHA_TEST: u8 = 3;
pub const SQW_BLEND_ALPHA_BLEND: u8 = 4;
pub const SQW_TC_FULL_PERSPECTIVE: u8 = 0; // rasterizer.rs:4-10
pub const SQW_TC_LINE_Z_CORRECTION: u8 = 1;
pub const SQW_TC_AFFINE: u8 = 2;
pub const SQW_TEX_SURFACE: u8 = 0; // tex_ref = DrawPolygonData::surface_pixels_offset
pub const SQW_TEX_IMAGE: u8 = 1; // tex_ref = image id
pub const SQW_TEX_VIEW: u8 = 2; // tex_ref = sqw_view_index of the portal / mirror view

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_polygon_point // PolygonPointProjected (rasterizer.rs:1484-1489)
{
	pub x: i32,
	pub y: i32,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_triangle_point // TrianglePointProjected (rasterizer.rs:1491-1498), 28 bytes
{
	pub x: i32,
	pub y: i32,
	pub tc: [i32; 2],
	pub light: [i32; 3],
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_draw_cmd // 64 bytes: one fill_polygon / fill_triangle call in submission order
{
	pub kind: u8,
	pub blend: u8,
	pub tc_mode: u8,
	pub tex_kind: u8,
	pub num_vertices: u32,
	pub vertex_offset: u32,
	pub tex_ref: u32,
	pub tex_size: [i32; 2],
	pub depth_eq: [f32; 3], // DepthEquation {d_inv_z_dx, d_inv_z_dy, k}
	pub tc_eq: [f32; 6],    // TexCoordEquation {d_tc_dx[2], d_tc_dy[2], k[2]}
	pub reserved: u32,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_post_params
{
	pub exposure: f32,
	pub color_modulate: [f32; 3],
	pub bloom_sigma: f32,
	pub bloom_scale: f32,
	pub bloom_buffer_scale_log2: u32,
	pub linear_bloom_filter: u32,
}

pub const SQW_STAGE_SURFACES: u32 = 1;
pub const SQW_STAGE_RASTER: u32 = 2;
pub const SQW_STAGE_POST: u32 = 4;
pub const SQW_STAGE_SHADOW_MAPS: u32 = 8;
pub const SQW_STAGE_ALL: u32 = 15;

#[repr(C)]
#[derive(Copy, Clone, Default)]
pub struct sqw_frame_stats
{
	pub surface_texels: u64,
	pub emissive_texels: u64,
	pub lightmap_bytes: u64,
	pub num_spans: u64,
	pub num_kernel_launches: u32,
	pub num_surface_variants: u32,
	pub num_tiles: u32,
	pub reserved: u32,
}

pub const SQW_KERNEL_POLYGONS_PREPARE: usize = 13;
pub const SQW_NUM_KERNEL_KINDS: usize = 14;
#[repr(C)]
pub struct sqw_kernel_times
{
	pub ms: [f64; SQW_NUM_KERNEL_KINDS],
	pub launches: [u64; SQW_NUM_KERNEL_KINDS],
}

pub const SQW_MESH_NON_ANIMATED: u32 = 0;
pub const SQW_MESH_VERTEX_ANIMATED: u32 = 1;
pub const SQW_MESH_SKELETON_ANIMATED: u32 = 2;
#[repr(C)]
pub struct sqw_model_mesh
{
	pub kind: u32,
	pub num_vertices: u32,
	pub num_triangles: u32,
	pub num_frames: u32,
	pub num_bones: u32,
	pub reserved: u32,
	pub vertices: *const c_void,    // VertexNonAnimated[] / VertexAnimatedVertexVariable[frames][n] / SkeletonAnimatedVertex[] (triangle_model.rs)
	pub tex_coords: *const f32,     // VertexAnimatedVertexConstant[] (vertex-animated meshes)
	pub triangles: *const u16,      // Triangle = [VertexIndex; 3]
	pub bone_parents: *const u32,
	pub frame_bones: *const f32,    // TriangleModelBoneFrame matrices, column-major
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_model_instance // 296 bytes: one VisibleDynamicMeshInfo + ModelLightData
{
	pub mesh: u32,
	pub frames: [u32; 2],
	pub lerp: f32,
	pub model_matrix: [f32; 16],
	pub model_view_matrix: [f32; 16],
	pub tc_scale: [f32; 2],
	pub tc_shift: [f32; 2],
	pub light_cube: [[f32; 3]; 6],
	pub light_vector_scaled: [[f32; 3]; 2],
	pub light_color: [[f32; 3]; 2],
	pub vertex_offset: u32,
	pub triangle_offset: u32,
	pub bone_offset: u32,
	pub sort_offset: u32,
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_model_vertex // ModelVertex3d (renderer_structs.rs:311-318)
{
	pub pos: [f32; 3],
	pub tc: [f32; 2],
	pub light: [f32; 3],
}

// ---- world-polygon preparation on the GPU (prepare_polygon_surface, partial_renderer.rs:1464-1638) ----
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_map_polygon // = Polygon (bsp_map_compact.rs:58-75): pass BSPMap::polygons.as_ptr() as it is
{
	pub first_vertex: u32,
	pub num_vertices: u32,
	pub plane: [f32; 4],
	pub tex_coord_equation: [[f32; 4]; 2],
	pub tex_coord_min: [i32; 2],
	pub tex_coord_max: [i32; 2],
	pub lightmap_data_offset: u32,
	pub texture: u32,
}
pub const SQW_MATERIAL_DRAW: u8 = 1;
pub const SQW_MATERIAL_SKYBOX: u8 = 2;
pub const SQW_MATERIAL_BLEND_SHIFT: u8 = 4; // bits 4-6: Material::blending_mode as SQW_BLEND_*
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_polygon_candidate
{
	pub polygon: u32,
	pub clip_set: u32,
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_polygons_frame
{
	pub planes_matrix: [f32; 16],
	pub view_matrix: [f32; 16],
	pub mip_bias: f32,
	pub reserved: [u32; 3],
}
pub const SQW_MAT_LIGHT: u32 = 1;
pub const SQW_MAT_HAS_DIFFUSE: u32 = 2;
pub const SQW_MAT_HAS_NORMAL_MAP: u32 = 4;
pub const SQW_MAT_HAS_NON_ONE_ROUGHNESS: u32 = 8;
pub const SQW_MAT_IS_METAL: u32 = 16;
pub const SQW_MAT_HAS_EMISSIVE: u32 = 32;
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_map_material
{
	pub flags: u32,
	pub texture_images: [u32; 5],
	pub emissive_images: [u32; 5],
	pub emissive_light: [f32; 3],
	pub texture_shift: [i32; 2],
}
pub const SQW_PREPARED_VISIBLE: u32 = 1;
pub const SQW_PREPARED_SKY: u32 = 2;
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_prepared_polygon // the per-frame half of DrawPolygonData (renderer_structs.rs:75-91)
{
	pub flags: u32,
	pub mip: u32,
	pub depth_equation: [f32; 3],
	pub tex_coord_equation: [f32; 6],
	pub surface_tc_min: [i32; 2],
	pub surface_size: [u32; 2],
	pub surface_pixels_offset: u32,
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct sqw_polygons_totals
{
	pub num_visible: u32,
	pub has_sky: u32,
	pub num_surfaces_pixels: u64,
	pub sky_bounds: [f32; 8],
	pub sky_material: u32,
	pub reserved: u32,
}

#[link(name = "sqw_cuda")]
extern "C" {
	// S0: context, images (Renderer::new renderer.rs:24, MapMaterialsProcessor::new map_materials_processor.rs:29)
	pub fn sqw_create(out_ctx: *mut *mut sqw_ctx, device: c_int, map: *const sqw_map_desc, external_stream: *mut c_void) -> c_int;
	pub fn sqw_destroy(ctx: *mut sqw_ctx);
	pub fn sqw_last_error(ctx: *const sqw_ctx) -> *const c_char;
	pub fn sqw_abi_version() -> c_int;
	pub fn sqw_upload_image(ctx: *mut sqw_ctx, kind: c_int, width: u32, height: u32, pixels: *const c_void) -> c_int;
	pub fn sqw_update_image(ctx: *mut sqw_ctx, image_id: c_int, pixels: *const c_void) -> c_int;
	pub fn sqw_set_fast_surface_numerics(ctx: *mut sqw_ctx, enable: c_int) -> c_int;
	pub fn sqw_set_compressed_readback(ctx: *mut sqw_ctx, enable: c_int) -> c_int;
	pub fn sqw_view_get_readback_stats(view: *const sqw_view, out: *mut u64) -> c_int;
	pub fn sqw_self_test(ctx: *mut sqw_ctx) -> c_int;
	// views (one per PartialRenderer: root + nested portal renderers, partial_renderer.rs:120-133)
	pub fn sqw_view_create(ctx: *mut sqw_ctx, out_view: *mut *mut sqw_view, width: u32, height: u32, color_bits: c_int) -> c_int;
	pub fn sqw_view_destroy(view: *mut sqw_view);
	pub fn sqw_view_set_size(view: *mut sqw_view, width: u32, height: u32) -> c_int;
	pub fn sqw_view_index(view: *const sqw_view) -> c_int;
	pub fn sqw_view_synchronize(view: *mut sqw_view) -> c_int;
	pub fn sqw_view_get_average_color(view: *mut sqw_view, avg_color_out: *mut f32) -> c_int;
	pub fn sqw_view_mix_texture(view: *mut sqw_view, image_id: u32, surface_tc_min: *const i32, blending_mode: c_int, light: *const f32) -> c_int;
	// S1 surfaces (body of build_polygons_surfaces, partial_renderer.rs:1682-1831)
	pub fn sqw_build_surfaces(
		view: *mut sqw_view, jobs: *const sqw_surface_job, num_jobs: u32, lights: *const sqw_light, num_lights: u32, shadow_maps: *const f32,
		shadow_maps_len: u64, cam_pos: *const f32, total_surface_pixels: u64,
	) -> c_int;
	// S2 rasterisation (perform_rasterization, partial_renderer.rs:1192-1283)
	pub fn sqw_draw(
		view: *mut sqw_view, cmds: *const sqw_draw_cmd, num_cmds: u32, polygon_points: *const sqw_polygon_point, num_polygon_points: u32,
		triangle_points: *const sqw_triangle_point, num_triangle_points: u32, clear_color: *const c_void,
	) -> c_int;
	// S3 post-processing (Postprocessor::perform_postprocessing, postprocessor.rs:68-204)
	pub fn sqw_postprocess(view: *mut sqw_view, params: *const sqw_post_params, dst: *mut u32, dst_pitch: usize, avg_color_out: *mut f32) -> c_int;
	// read-back (the `pixels: &mut [ColorT]` arguments)
	pub fn sqw_read_pixels(view: *mut sqw_view, dst: *mut c_void, dst_pitch: usize) -> c_int;
	pub fn sqw_write_pixels(view: *mut sqw_view, src: *const c_void, src_pitch: usize) -> c_int;
	pub fn sqw_read_pixels32(view: *mut sqw_view, dst: *mut u32, dst_pitch: usize) -> c_int;
	pub fn sqw_read_pixels_async(view: *mut sqw_view, dst: *mut c_void, dst_pitch: usize) -> c_int;
	pub fn sqw_read_pixels32_async(view: *mut sqw_view, dst: *mut u32, dst_pitch: usize) -> c_int;
	pub fn sqw_read_surfaces(view: *mut sqw_view, dst: *mut c_void, offset: u64, count: u64) -> c_int;
	// recorded frames / batched views
	pub fn sqw_frame_create(view: *mut sqw_view, out_frame: *mut *mut sqw_frame) -> c_int;
	pub fn sqw_frame_destroy(frame: *mut sqw_frame);
	pub fn sqw_frame_set_surfaces(
		frame: *mut sqw_frame, jobs: *const sqw_surface_job, num_jobs: u32, lights: *const sqw_light, num_lights: u32, shadow_maps: *const f32,
		shadow_maps_len: u64, cam_pos: *const f32, total_surface_pixels: u64,
	) -> c_int;
	pub fn sqw_frame_set_draw(
		frame: *mut sqw_frame, cmds: *const sqw_draw_cmd, num_cmds: u32, polygon_points: *const sqw_polygon_point, num_polygon_points: u32,
		triangle_points: *const sqw_triangle_point, num_triangle_points: u32, clear_color: *const c_void,
	) -> c_int;
	pub fn sqw_frame_set_post(frame: *mut sqw_frame, params: *const sqw_post_params) -> c_int;
	pub fn sqw_frame_set_shadow_map_draw(
		frame: *mut sqw_frame, maps: *const sqw_depth_map, num_maps: u32, cmds: *const sqw_depth_cmd, num_cmds: u32, polygon_points: *const sqw_polygon_point,
		num_polygon_points: u32,
	) -> c_int;
	pub fn sqw_frame_read_shadow_maps(frame: *mut sqw_frame, dst: *mut f32, offset: u64, count: u64) -> c_int;
	pub fn sqw_frame_submit(frame: *mut sqw_frame, stages: u32) -> c_int;
	pub fn sqw_frame_get_stats(frame: *const sqw_frame, out: *mut sqw_frame_stats) -> c_int;
	pub fn sqw_batch_begin(ctx: *mut sqw_ctx) -> c_int;
	pub fn sqw_batch_end(ctx: *mut sqw_ctx) -> c_int;
	pub fn sqw_profile_begin(ctx: *mut sqw_ctx) -> c_int;
	pub fn sqw_profile_end(ctx: *mut sqw_ctx, out: *mut sqw_kernel_times) -> c_int;
	// triangle-model preparation (build_dynamic_models_buffers, partial_renderer.rs:1122-1190)
	pub fn sqw_upload_model_mesh(ctx: *mut sqw_ctx, mesh: *const sqw_model_mesh) -> c_int;
	pub fn sqw_models_prepare(view: *mut sqw_view, instances: *const sqw_model_instance, num_instances: u32) -> c_int;
	pub fn sqw_models_read(
		view: *mut sqw_view, vertices: *mut sqw_model_vertex, num_vertices: u64, triangles: *mut u16, num_triangles: u64, counts: *mut u32, num_instances: u32,
	) -> c_int;
	// world-polygon preparation (loop body of prepare_polygons_surfaces, partial_renderer.rs:1327-1360)
	pub fn sqw_upload_map_geometry(
		ctx: *mut sqw_ctx, polygons: *const sqw_map_polygon, num_polygons: u32, vertices: *const f32, num_vertices: u32, material_flags: *const u8, num_materials: u32,
	) -> c_int;
	pub fn sqw_polygons_prepare(
		view: *mut sqw_view, frame: *const sqw_polygons_frame, candidates: *const sqw_polygon_candidate, num_candidates: u32, clip_planes: *const f32, num_clip_sets: u32,
	) -> c_int;
	pub fn sqw_polygons_read(view: *mut sqw_view, out: *mut sqw_prepared_polygon, num: u32, totals: *mut sqw_polygons_totals) -> c_int;
	// draw_polygon (partial_renderer.rs:4011-4142) for the prepared polygons in BSP draw order
	pub fn sqw_polygons_draw(view: *mut sqw_view, order: *const u32, num_order: u32) -> c_int;
	pub fn sqw_frame_set_draw_from_polygons(frame: *mut sqw_frame, clear_color: *const c_void) -> c_int;
	pub fn sqw_upload_map_materials(ctx: *mut sqw_ctx, materials: *const sqw_map_material, num_materials: u32, polygon_basis: *const f32, num_polygons: u32) -> c_int;
	pub fn sqw_frame_set_surfaces_from_polygons(
		frame: *mut sqw_frame, cam_pos: *const f32, use_directional_lightmaps: c_int, lights: *const sqw_light, num_lights: u32, clip_set_light_offsets: *const u32,
		clip_set_lights: *const u16, num_clip_sets: u32, shadow_maps_len: u64,
	) -> c_int;
	pub fn sqw_polygons_draw_read(
		view: *mut sqw_view, cmds: *mut sqw_draw_cmd, cmds_capacity: u32, points: *mut sqw_polygon_point, points_capacity: u32, counts_out: *mut u32,
	) -> c_int;
}

/// Turns a status into the abort the reference would have had (`panic = "abort"`, Cargo.toml:8), with the library's message.
pub fn check(ctx: *const sqw_ctx, rc: c_int) -> c_int
{
	if rc < 0
	{
		let msg = unsafe { std::ffi::CStr::from_ptr(sqw_last_error(ctx)) }.to_string_lossy().into_owned();
		panic!("libsqw_cuda: error {}: {}", rc, msg);
	}
	rc
}

// layouts asserted by the C side too (square_wheel_b200/csrc/sqw_api.cu:16-21)
const _: [(); 136] = [(); std::mem::size_of::<sqw_surface_job>()];
const _: [(); 80] = [(); std::mem::size_of::<sqw_light>()];
const _: [(); 64] = [(); std::mem::size_of::<sqw_draw_cmd>()];
const _: [(); 8] = [(); std::mem::size_of::<sqw_polygon_point>()];
const _: [(); 28] = [(); std::mem::size_of::<sqw_triangle_point>()];
const _: [(); 32] = [(); std::mem::size_of::<sqw_depth_map>()];
const _: [(); 32] = [(); std::mem::size_of::<sqw_depth_cmd>()];
const _: [(); 32] = [(); std::mem::size_of::<sqw_post_params>()];
const _: [(); 296] = [(); std::mem::size_of::<sqw_model_instance>()];
const _: [(); 32] = [(); std::mem::size_of::<sqw_model_vertex>()];
const _: [(); 80] = [(); std::mem::size_of::<sqw_map_polygon>()];
const _: [(); 64] = [(); std::mem::size_of::<sqw_prepared_polygon>()];
const _: [(); 144] = [(); std::mem::size_of::<sqw_polygons_frame>()];
