"""ctypes binding of the C-ABI in include/sqw_cuda.h (libsqw_cuda.so) and of the host-side library (libsqw_host.so).

There is no CPU fallback: if libsqw_cuda.so is missing the import of the product path fails loudly with
build instructions, and on a box without a CUDA device ``sqw_create`` returns SQW_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "libsqw_cuda.so")
HOST_LIB_PATH = os.path.join(_HERE, "libsqw_host.so")

# ---- constants (include/sqw_cuda.h) -------------------------------------------------------------------------
SQW_OK = 0
IMAGE_TEXTURE_ELEMENT, IMAGE_COLOR32, IMAGE_COLOR64 = 0, 1, 2
SURF_HAS_NORMAL_MAP, SURF_HAS_NON_ONE_ROUGHNESS, SURF_IS_METAL = 1, 2, 4
SURF_DIRECTIONAL_LIGHTMAP, SURF_APPLY_DIFFUSE, SURF_EMISSIVE = 8, 16, 32
SHADOW_NONE, SHADOW_CUBE, SHADOW_PROJECTOR = 0, 1, 2
CMD_POLYGON, CMD_TRIANGLE = 0, 1
BLEND_NONE, BLEND_AVERAGE, BLEND_ADDITIVE, BLEND_ALPHA_TEST, BLEND_ALPHA_BLEND = range(5)
TC_FULL_PERSPECTIVE, TC_LINE_Z_CORRECTION, TC_AFFINE = 0, 1, 2
TEX_SURFACE, TEX_IMAGE, TEX_VIEW = 0, 1, 2
STAGE_SURFACES, STAGE_RASTER, STAGE_POST, STAGE_SHADOW_MAPS, STAGE_ALL = 1, 2, 4, 8, 15
DEPTH_MAP_CLEAR, DEPTH_MAP_PROJECTOR_CIRCLE = 1, 2
CTX_CONCURRENT_VIEWS = 1
CTX_FAST_SURFACE_NUMERICS = 2
CTX_COMPRESSED_READBACK = 4

# ---- struct layouts as numpy dtypes (sizes asserted against the static_asserts in sqw_api.cu) ---------------
surface_job_dtype = np.dtype(
    [
        ("basis_u", "<f4", 3),
        ("basis_v", "<f4", 3),
        ("basis_start", "<f4", 3),
        ("basis_normal", "<f4", 3),
        ("surface_size", "<u4", 2),
        ("tc_start", "<i4", 2),
        ("texture_image", "<u4"),
        ("flags", "<u4"),
        ("lightmap_size", "<u4", 2),
        ("lightmap_scale_log2", "<u4"),
        ("lightmap_tc_shift", "<u4", 2),
        ("lightmap_offset", "<u4"),
        ("num_lights", "<u4"),
        ("lights", "<u2", 6),
        ("emissive_image", "<u4"),
        ("emissive_light", "<f4", 3),
        ("surface_pixels_offset", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert surface_job_dtype.itemsize == 136

light_dtype = np.dtype(
    [
        ("position", "<f4", 3),
        ("radius", "<f4"),
        ("inv_square_radius", "<f4"),
        ("color", "<f4", 3),
        ("shadow_type", "<u4"),
        ("shadow_map_size", "<u4"),
        ("shadow_map_offset", "<u4"),
        ("basis_x", "<f4", 3),
        ("basis_y", "<f4", 3),
        ("basis_z", "<f4", 3),
    ]
)
assert light_dtype.itemsize == 80

draw_cmd_dtype = np.dtype(
    [
        ("kind", "u1"),
        ("blend", "u1"),
        ("tc_mode", "u1"),
        ("tex_kind", "u1"),
        ("num_vertices", "<u4"),
        ("vertex_offset", "<u4"),
        ("tex_ref", "<u4"),
        ("tex_size", "<i4", 2),
        ("depth_eq", "<f4", 3),
        ("tc_eq", "<f4", 6),
        ("reserved", "<u4"),
    ]
)
assert draw_cmd_dtype.itemsize == 64

polygon_point_dtype = np.dtype([("x", "<i4"), ("y", "<i4")])
triangle_point_dtype = np.dtype([("x", "<i4"), ("y", "<i4"), ("tc", "<i4", 2), ("light", "<i4", 3)])
assert triangle_point_dtype.itemsize == 28

directional_lightmap_dtype = np.dtype(
    [
        ("ambient_light", "<f4", 3),
        ("light_direction_vector_scaled", "<f4", 3),
        ("directional_light_deviation", "<f4"),
        ("directional_light_color", "<f4", 3),
    ]
)
assert directional_lightmap_dtype.itemsize == 40

texture_element_dtype = np.dtype([("diffuse", "<u4"), ("packed_normal_roughness", "<i4")])


depth_map_dtype = np.dtype(
    [("offset", "<u4"), ("width", "<u4"), ("height", "<u4"), ("first_cmd", "<u4"), ("num_cmds", "<u4"), ("flags", "<u4"), ("clear_value", "<f4"), ("reserved", "<u4")]
)
assert depth_map_dtype.itemsize == 32
depth_cmd_dtype = np.dtype([("num_vertices", "<u4"), ("vertex_offset", "<u4"), ("depth_test", "<u4"), ("depth_eq", "<f4", (3,)), ("reserved", "<u4", (2,))])
assert depth_cmd_dtype.itemsize == 32


class ModelMesh(C.Structure):  # sqw_model_mesh
    _fields_ = [
        ("kind", C.c_uint32), ("num_vertices", C.c_uint32), ("num_triangles", C.c_uint32), ("num_frames", C.c_uint32), ("num_bones", C.c_uint32), ("reserved", C.c_uint32),
        ("vertices", C.c_void_p), ("tex_coords", C.c_void_p), ("triangles", C.c_void_p), ("bone_parents", C.c_void_p), ("frame_bones", C.c_void_p),
    ]


MESH_NON_ANIMATED, MESH_VERTEX_ANIMATED, MESH_SKELETON_ANIMATED = 0, 1, 2
model_instance_dtype = np.dtype(
    [
        ("mesh", "<u4"), ("frames", "<u4", 2), ("lerp", "<f4"), ("model_matrix", "<f4", 16), ("model_view_matrix", "<f4", 16), ("tc_scale", "<f4", 2), ("tc_shift", "<f4", 2),
        ("light_cube", "<f4", (6, 3)), ("light_vector_scaled", "<f4", (2, 3)), ("light_color", "<f4", (2, 3)),
        ("vertex_offset", "<u4"), ("triangle_offset", "<u4"), ("bone_offset", "<u4"), ("sort_offset", "<u4"),
    ]
)
assert model_instance_dtype.itemsize == 296
model_vertex_dtype = np.dtype([("pos", "<f4", 3), ("tc", "<f4", 2), ("light", "<f4", 3)])


class MapDesc(C.Structure):
    _fields_ = [
        ("lightmaps_data", C.c_void_p),
        ("directional_lightmaps_data", C.c_void_p),
        ("num_lightmap_elements", C.c_uint64),
        ("max_images", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class PostParams(C.Structure):
    _fields_ = [
        ("exposure", C.c_float),
        ("color_modulate", C.c_float * 3),
        ("bloom_sigma", C.c_float),
        ("bloom_scale", C.c_float),
        ("bloom_buffer_scale_log2", C.c_uint32),
        ("linear_bloom_filter", C.c_uint32),
    ]


assert C.sizeof(PostParams) == 32


class FrameStats(C.Structure):
    _fields_ = [
        ("surface_texels", C.c_uint64),
        ("emissive_texels", C.c_uint64),
        ("lightmap_bytes", C.c_uint64),
        ("num_spans", C.c_uint64),
        ("num_kernel_launches", C.c_uint32),
        ("num_surface_variants", C.c_uint32),
        ("num_tiles", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class KernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 14), ("launches", C.c_uint64 * 14)]


KERNEL_NAMES = ["surface_build", "raster_setup", "raster_bin", "raster_tile", "tonemap", "bloom_downscale", "bloom_blur", "tonemap_bloom", "raster_tile_tonemap", "mix_texture", "depth_setup", "depth_resolve", "models_prepare", "polygons_prepare"]

# world-polygon preparation (sqw_polygons_prepare)
map_polygon_dtype = np.dtype(
    [("first_vertex", "<u4"), ("num_vertices", "<u4"), ("plane", "<f4", 4), ("tex_coord_equation", "<f4", (2, 4)), ("tex_coord_min", "<i4", 2), ("tex_coord_max", "<i4", 2),
     ("lightmap_data_offset", "<u4"), ("texture", "<u4")])
map_material_dtype = np.dtype([("flags", "<u4"), ("texture_images", "<u4", 5), ("emissive_images", "<u4", 5), ("emissive_light", "<f4", 3), ("texture_shift", "<i4", 2)])
assert map_material_dtype.itemsize == 64
polygon_candidate_dtype = np.dtype([("polygon", "<u4"), ("clip_set", "<u4")])
prepared_polygon_dtype = np.dtype(
    [("flags", "<u4"), ("mip", "<u4"), ("depth_equation", "<f4", 3), ("tex_coord_equation", "<f4", 6), ("surface_tc_min", "<i4", 2), ("surface_size", "<u4", 2),
     ("surface_pixels_offset", "<u4")])
assert map_polygon_dtype.itemsize == 80 and prepared_polygon_dtype.itemsize == 64
MATERIAL_DRAW, MATERIAL_SKYBOX = 1, 2
PREPARED_VISIBLE, PREPARED_SKY = 1, 2


class PolygonsFrame(C.Structure):
    _fields_ = [("planes_matrix", C.c_float * 16), ("view_matrix", C.c_float * 16), ("mip_bias", C.c_float), ("reserved", C.c_uint32 * 3)]


class PolygonsTotals(C.Structure):
    _fields_ = [("num_visible", C.c_uint32), ("has_sky", C.c_uint32), ("num_surfaces_pixels", C.c_uint64), ("sky_bounds", C.c_float * 8), ("sky_material", C.c_uint32),
                ("reserved", C.c_uint32)]


assert C.sizeof(PolygonsFrame) == 144

# Every symbol include/sqw_cuda.h declares: (name, restype, argtypes)
_VP = C.c_void_p
CUDA_SYMBOLS = [
    ("sqw_abi_version", C.c_int, []),
    ("sqw_create", C.c_int, [C.POINTER(_VP), C.c_int, C.POINTER(MapDesc), _VP]),
    ("sqw_destroy", None, [_VP]),
    ("sqw_last_error", C.c_char_p, [_VP]),
    ("sqw_upload_image", C.c_int, [_VP, C.c_int, C.c_uint32, C.c_uint32, _VP]),
    ("sqw_update_image", C.c_int, [_VP, C.c_int, _VP]),
    ("sqw_view_create", C.c_int, [_VP, C.POINTER(_VP), C.c_uint32, C.c_uint32, C.c_int]),
    ("sqw_view_destroy", None, [_VP]),
    ("sqw_build_surfaces", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint64, _VP, C.c_uint64]),
    ("sqw_draw", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint32, _VP]),
    ("sqw_postprocess", C.c_int, [_VP, C.POINTER(PostParams), _VP, C.c_size_t, _VP]),
    ("sqw_read_pixels", C.c_int, [_VP, _VP, C.c_size_t]),
    ("sqw_read_pixels32", C.c_int, [_VP, _VP, C.c_size_t]),
    ("sqw_read_pixels_async", C.c_int, [_VP, _VP, C.c_size_t]),
    ("sqw_read_pixels32_async", C.c_int, [_VP, _VP, C.c_size_t]),
    ("sqw_view_set_size", C.c_int, [_VP, C.c_uint32, C.c_uint32]),
    ("sqw_view_index", C.c_int, [_VP]),
    ("sqw_view_mix_texture", C.c_int, [_VP, C.c_uint32, _VP, C.c_int, _VP]),
    ("sqw_frame_set_shadow_map_draw", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint32]),
    ("sqw_frame_read_shadow_maps", C.c_int, [_VP, _VP, C.c_uint64, C.c_uint64]),
    ("sqw_batch_begin", C.c_int, [_VP]),
    ("sqw_batch_end", C.c_int, [_VP]),
    ("sqw_write_pixels", C.c_int, [_VP, _VP, C.c_size_t]),
    ("sqw_read_surfaces", C.c_int, [_VP, _VP, C.c_uint64, C.c_uint64]),
    ("sqw_view_synchronize", C.c_int, [_VP]),
    ("sqw_frame_create", C.c_int, [_VP, C.POINTER(_VP)]),
    ("sqw_frame_destroy", None, [_VP]),
    ("sqw_frame_set_surfaces", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint64, _VP, C.c_uint64]),
    ("sqw_frame_set_draw", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint32, _VP]),
    ("sqw_frame_set_post", C.c_int, [_VP, C.POINTER(PostParams)]),
    ("sqw_frame_submit", C.c_int, [_VP, C.c_uint32]),
    ("sqw_frame_get_stats", C.c_int, [_VP, C.POINTER(FrameStats)]),
    ("sqw_view_get_average_color", C.c_int, [_VP, _VP]),
    ("sqw_profile_begin", C.c_int, [_VP]),
    ("sqw_profile_end", C.c_int, [_VP, C.POINTER(KernelTimes)]),
    ("sqw_self_test", C.c_int, [_VP]),
    ("sqw_set_fast_surface_numerics", C.c_int, [_VP, C.c_int]),
    ("sqw_set_compressed_readback", C.c_int, [_VP, C.c_int]),
    ("sqw_view_get_readback_stats", C.c_int, [_VP, _VP]),
    ("sqw_upload_model_mesh", C.c_int, [_VP, _VP]),
    ("sqw_models_prepare", C.c_int, [_VP, _VP, C.c_uint32]),
    ("sqw_models_read", C.c_int, [_VP, _VP, C.c_uint64, _VP, C.c_uint64, _VP, C.c_uint32]),
    ("sqw_upload_map_geometry", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP, C.c_uint32]),
    ("sqw_polygons_prepare", C.c_int, [_VP, C.POINTER(PolygonsFrame), _VP, C.c_uint32, _VP, C.c_uint32]),
    ("sqw_polygons_read", C.c_int, [_VP, _VP, C.c_uint32, C.POINTER(PolygonsTotals)]),
    ("sqw_polygons_draw", C.c_int, [_VP, _VP, C.c_uint32]),
    ("sqw_polygons_draw_read", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32, _VP]),
    ("sqw_frame_set_draw_from_polygons", C.c_int, [_VP, _VP]),
    ("sqw_upload_map_materials", C.c_int, [_VP, _VP, C.c_uint32, _VP, C.c_uint32]),
    ("sqw_frame_set_surfaces_from_polygons", C.c_int, [_VP, _VP, C.c_int, _VP, C.c_uint32, _VP, _VP, C.c_uint32, C.c_uint64]),
]

_cuda_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load_cuda_library() -> C.CDLL:
    """Load libsqw_cuda.so and bind every entry point.  Fails loudly when the library has not been built."""
    global _cuda_lib
    if _cuda_lib is not None:
        return _cuda_lib
    if not os.path.exists(CUDA_LIB_PATH):
        raise NativeLibraryMissing(
            f"{CUDA_LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). square_wheel_b200 has no CPU fallback for the world-drawing path."
        )
    lib = C.CDLL(CUDA_LIB_PATH)
    for name, restype, argtypes in CUDA_SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _cuda_lib = lib
    return lib


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data_as(C.c_void_p)
    return a
