"""ctypes binding of libsqw_host.so — the C++ host-side mirror of the reference's renderer front end
(square_wheel_b200/host/): synthetic map + materials + textures, camera paths, and the per-frame host work that
"stays host" in SURVEY.md §8 (visibility, surface preparation, BSP traversal) producing job and command lists."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from . import _native as N

_lib = None

FOV_DEFAULT = 0.375 * math.pi  # test_game/game_map.rs:865


class SceneParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint32),
        ("grid", C.c_uint32 * 3),
        ("cell_size", C.c_float * 3),
        ("num_rooms", C.c_uint32),
        ("num_materials", C.c_uint32),
        ("texture_size_log2", C.c_uint32),
        ("max_polygon_size", C.c_float),
        ("material_mode", C.c_uint32),
        ("emissive_fraction", C.c_float),
        ("num_models", C.c_uint32),
        ("model_detail", C.c_uint32),
        ("skybox", C.c_uint32),
        ("sky_size_log2", C.c_uint32),
        ("light_grid", C.c_uint32),
        ("num_submodels", C.c_uint32),
    ]


# Renderer::DynamicMeshExport (host/renderer.hpp)
dynamic_mesh_dtype = np.dtype(
    [
        ("model", "<u4"), ("mesh", "<u4"), ("frames", "<u4", 2), ("lerp", "<f4"), ("model_matrix", "<f4", 16), ("model_view_matrix", "<f4", 16), ("tc_scale", "<f4", 2),
        ("tc_shift", "<f4", 2), ("light_cube", "<f4", (6, 3)), ("light_vector_scaled", "<f4", (2, 3)), ("light_color", "<f4", (2, 3)),
        ("vertices_offset", "<u4"), ("num_vertices", "<u4"), ("triangles_offset", "<u4"), ("num_triangles", "<u4"), ("num_visible_triangles", "<u4"),
    ]
)
assert dynamic_mesh_dtype.itemsize == 304


class SceneInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_nodes", "num_leafs", "num_polygons", "num_portals", "num_vertices", "num_materials", "num_images", "num_waypoints")] + [
        ("num_lightmap_elements", C.c_uint64),
        ("num_models", C.c_uint32),
        ("num_submodels", C.c_uint32),
    ]


class RendererConfig(C.Structure):
    """renderer_config.rs:5-31 (the fields the path reads)."""

    _fields_ = [
        ("use_directional_lightmaps", C.c_uint32),
        ("textures_mip_bias", C.c_float),
        ("clear_background", C.c_uint32),
        ("invert_polygons_order", C.c_uint32),
        ("shadows_quality", C.c_float),
        ("portals_depth", C.c_uint32),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("num_jobs", C.c_uint32),
        ("num_cmds", C.c_uint32),
        ("num_polygon_points", C.c_uint32),
        ("needs_background_clear", C.c_uint32),
        ("total_surface_pixels", C.c_uint64),
        ("num_visible_leafs", C.c_uint32),
        ("num_visible_polygons", C.c_uint32),
        ("mode_counts", C.c_uint32 * 3),
        ("cam_pos", C.c_float * 3),
        ("num_lights", C.c_uint32),
        ("num_depth_maps", C.c_uint32),
        ("num_depth_cmds", C.c_uint32),
        ("num_depth_points", C.c_uint32),
        ("num_triangle_points", C.c_uint32),
        ("shadow_maps_len", C.c_uint64),
    ]


# frame_info.rs:149-168 DynamicLight (rotation = quaternion s, x, y, z and fov only matter for projector shadows)
dynamic_light_dtype = np.dtype(
    [("position", "<f4", (3,)), ("radius", "<f4"), ("color", "<f4", (3,)), ("shadow_type", "<u4"), ("rotation", "<f4", (4,)), ("fov", "<f4")]
)
assert dynamic_light_dtype.itemsize == 52

# frame_info.rs:55-69 ModelEntity (model / skin = indices of the scene's fixtures; rotation = quaternion s, x, y, z)
model_entity_dtype = np.dtype(
    [
        ("position", "<f4", (3,)), ("rotation", "<f4", (4,)), ("frames", "<u4", (2,)), ("lerp", "<f4"), ("model", "<u4"), ("skin", "<u4"), ("blending_mode", "<u4"),
        ("lighting", "<u4"), ("light", "<f4", (3,)), ("grid_light_scale", "<f4"), ("light_position", "<f4", (3,)), ("flags", "<u4"),
    ]
)
assert model_entity_dtype.itemsize == 88
MODEL_LIGHTING_DEFAULT, MODEL_LIGHTING_CONSTANT, MODEL_LIGHTING_ADVANCED = 0, 1, 2

# frame_info.rs:125-147 Sprite (skin = index of the scene's TextureLiteWithMips fixtures; orientation = SpriteOrientation 0..3)
sprite_dtype = np.dtype(
    [("position", "<f4", (3,)), ("angle", "<f4"), ("radius", "<f4"), ("skin", "<u4"), ("blending_mode", "<u4"), ("orientation", "<u4"), ("light_scale", "<f4"), ("light_add", "<f4", (3,))]
)
assert sprite_dtype.itemsize == 48

# frame_info.rs:112-123 Decal: a unit cube placed by position / rotation / scale, projected along its -X axis onto the polygons inside
decal_dtype = np.dtype(
    [("position", "<f4", (3,)), ("rotation", "<f4", (4,)), ("scale", "<f4", (3,)), ("skin", "<u4"), ("blending_mode", "<u4"), ("lightmap_light_scale", "<f4"), ("light_add", "<f4", (3,))]
)
assert decal_dtype.itemsize == 64

# frame_info.rs:13 SubmodelEntityOpt per map submodel: present = 0 hides it; position = where its bbox centre goes
submodel_entity_dtype = np.dtype([("present", "<u4"), ("position", "<f4", (3,)), ("rotation", "<f4", (4,))])
assert submodel_entity_dtype.itemsize == 32

# frame_info.rs:171-205 ViewPortal (+ ViewPortalTexture): view 0 CameraAtPosition / 1 Mirror / 2 ParallaxPortal; up to 8 vertices;
# transform_matrix column-major; texture_skin = index of the scene's TextureLiteWithMips fixtures
MAX_PORTAL_VERTICES = 8
portal_dtype = np.dtype(
    [
        ("view", "<u4"), ("camera_position", "<f4", (3,)), ("camera_rotation", "<f4", (4,)), ("camera_fov", "<f4"), ("transform_matrix", "<f4", (16,)), ("plane", "<f4", (4,)),
        ("tex_coord_equation", "<f4", (2, 4)), ("num_vertices", "<u4"), ("vertices", "<f4", (MAX_PORTAL_VERTICES, 3)), ("blending_mode", "<u4"), ("has_texture", "<u4"),
        ("texture_skin", "<u4"), ("texture_blending_mode", "<u4"), ("texture_light_scale", "<f4"), ("texture_light_add", "<f4", (3,)),
    ]
)
assert portal_dtype.itemsize == 280
PORTAL_VIEW_CAMERA_AT_POSITION, PORTAL_VIEW_MIRROR, PORTAL_VIEW_PARALLAX_PORTAL = 0, 1, 2


class PortalInfo(C.Structure):
    _fields_ = [
        ("resolution", C.c_uint32 * 2), ("tc_min", C.c_int32 * 2), ("mip", C.c_uint32), ("has_texture", C.c_uint32), ("texture_image", C.c_uint32),
        ("texture_blending_mode", C.c_uint32), ("texture_light", C.c_float * 3),
    ]


class PortalView:
    """One visible portal of a host frame: its texture geometry, the nested host frame that fills it (None at the depth limit:
    the texture is black, partial_renderer.rs:2038-2044) and, once rendered, the device view that holds the texture."""

    def __init__(self, index, info, frame):
        self.index, self.frame = index, frame
        self.resolution = (int(info.resolution[0]), int(info.resolution[1]))
        self.tc_min = (int(info.tc_min[0]), int(info.tc_min[1]))
        self.mip = int(info.mip)
        self.overlay = (int(info.texture_image), int(info.texture_blending_mode), tuple(info.texture_light)) if info.has_texture else None
        self.view = None  # device View, set by Renderer


class ModelInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_frames", "num_meshes", "num_vertices", "num_triangles", "kind", "num_bones")] + [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3)]


# Renderer::PolygonsFrameExport (host/renderer.hpp)
polygons_frame_dtype = np.dtype(
    [("planes_matrix", "<f4", 16), ("view_matrix", "<f4", 16), ("mip_bias", "<f4"), ("has_sky", "<u4"), ("sky_material", "<u4"), ("sky_bounds", "<f4", 8), ("num_visible", "<u4"),
     ("num_surfaces_pixels", "<u8")])
assert polygons_frame_dtype.itemsize == 184


def build_host_library():
    subprocess.run(["make", "-C", os.path.join(os.path.dirname(N.HOST_LIB_PATH), "host")], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(N.HOST_LIB_PATH):
            build_host_library()
        l = C.CDLL(N.HOST_LIB_PATH)
        vp = C.c_void_p
        l.sqwh_scene_create.restype = vp
        l.sqwh_scene_create.argtypes = [C.POINTER(SceneParams)]
        l.sqwh_scene_destroy.argtypes = [vp]
        l.sqwh_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
        l.sqwh_scene_image.restype = vp
        l.sqwh_scene_image.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.sqwh_scene_lightmaps.restype = vp
        l.sqwh_scene_lightmaps.argtypes = [vp]
        l.sqwh_scene_directional_lightmaps.restype = vp
        l.sqwh_scene_directional_lightmaps.argtypes = [vp]
        l.sqwh_scene_camera_path.restype = C.c_uint32
        l.sqwh_scene_camera_path.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
        l.sqwh_renderer_create.restype = vp
        l.sqwh_renderer_create.argtypes = [vp, C.POINTER(RendererConfig)]
        l.sqwh_renderer_destroy.argtypes = [vp]
        l.sqwh_renderer_prepare_frame.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_float]
        l.sqwh_renderer_draw_frame.argtypes = [vp, C.c_int]
        l.sqwh_renderer_frame_info.argtypes = [vp, C.POINTER(FrameInfo)]
        l.sqwh_renderer_set_lights.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_dynamic_meshes.restype = C.c_int
        l.sqwh_renderer_dynamic_meshes.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        l.sqwh_renderer_set_models.restype = C.c_int
        l.sqwh_renderer_set_models.argtypes = [vp, vp, vp, C.c_uint32]
        l.sqwh_renderer_set_submodel_entities.restype = C.c_int
        l.sqwh_renderer_set_submodel_entities.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_submodel_center.argtypes = [vp, C.c_uint32, vp]
        l.sqwh_renderer_set_portals.restype = C.c_int
        l.sqwh_renderer_set_portals.argtypes = [vp, vp, vp, C.c_uint32]
        l.sqwh_renderer_num_portals.restype = C.c_uint32
        l.sqwh_renderer_num_portals.argtypes = [vp]
        l.sqwh_renderer_portal_info.restype = C.c_int
        l.sqwh_renderer_portal_info.argtypes = [vp, C.c_uint32, C.POINTER(PortalInfo)]
        l.sqwh_renderer_prepare_portal_view.restype = vp
        l.sqwh_renderer_prepare_portal_view.argtypes = [vp, C.c_uint32, C.c_int]
        l.sqwh_renderer_set_decals.restype = C.c_int
        l.sqwh_renderer_set_decals.argtypes = [vp, vp, vp, C.c_uint32]
        l.sqwh_renderer_set_sprites.restype = C.c_int
        l.sqwh_renderer_set_sprites.argtypes = [vp, vp, vp, C.c_uint32]
        l.sqwh_scene_save_map.restype = C.c_int
        l.sqwh_scene_save_map.argtypes = [vp, C.c_char_p]
        l.sqwh_scene_load_map.restype = C.c_int
        l.sqwh_scene_load_map.argtypes = [vp, C.c_char_p]
        l.sqwh_scene_submodel_classname.restype = C.c_char_p
        l.sqwh_scene_submodel_classname.argtypes = [vp, C.c_uint32]
        l.sqwh_scene_load_compiled_map.restype = C.c_int
        l.sqwh_scene_load_compiled_map.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32]
        l.sqwh_last_map_error.restype = C.c_char_p
        l.sqwh_scene_num_light_grid_samples.restype = C.c_uint32
        l.sqwh_scene_num_light_grid_samples.argtypes = [vp]
        l.sqwh_scene_texture_name.restype = C.c_char_p
        l.sqwh_scene_texture_name.argtypes = [vp, C.c_uint32]
        l.sqwh_scene_model_info.restype = C.c_int
        l.sqwh_scene_model_info.argtypes = [vp, C.c_uint32, C.POINTER(ModelInfo)]
        l.sqwh_scene_load_model_file.restype = C.c_int
        l.sqwh_scene_load_model_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
        l.sqwh_scene_load_model_memory.restype = C.c_int
        l.sqwh_scene_load_model_memory.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        l.sqwh_scene_model_copy.restype = C.c_int64
        l.sqwh_scene_model_copy.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint64]
        for name in (
            "sqwh_renderer_jobs", "sqwh_renderer_cmds", "sqwh_renderer_polygon_points", "sqwh_renderer_lights", "sqwh_renderer_depth_maps", "sqwh_renderer_depth_cmds",
            "sqwh_renderer_depth_points", "sqwh_renderer_triangle_points",
        ):
            getattr(l, name).restype = vp
            getattr(l, name).argtypes = [vp]
        l.sqwh_default_scene_params.argtypes = [C.POINTER(SceneParams)]
        l.sqwh_scene_polygons.restype = vp
        l.sqwh_scene_polygons.argtypes = [vp]
        l.sqwh_scene_vertices.restype = vp
        l.sqwh_scene_vertices.argtypes = [vp]
        l.sqwh_scene_material_flags.restype = C.c_uint32
        l.sqwh_scene_material_flags.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_polygons_prepare_seconds.restype = C.c_double
        l.sqwh_renderer_polygons_prepare_seconds.argtypes = [vp]
        l.sqwh_scene_map_materials.restype = C.c_uint32
        l.sqwh_scene_map_materials.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_polygon_basis.restype = C.c_uint32
        l.sqwh_renderer_polygon_basis.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_clip_set_lights.restype = C.c_uint32
        l.sqwh_renderer_clip_set_lights.argtypes = [vp, vp, vp, vp]
        l.sqwh_renderer_polygon_draw_order.restype = C.c_uint32
        l.sqwh_renderer_polygon_draw_order.argtypes = [vp, vp, C.c_uint32]
        l.sqwh_renderer_prepared_polygons.restype = C.c_int
        l.sqwh_renderer_prepared_polygons.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
        _lib = l
    return _lib


def default_scene_params(**overrides) -> SceneParams:
    p = SceneParams()
    lib().sqwh_default_scene_params(C.byref(p))
    for k, v in overrides.items():
        if k in ("grid", "cell_size"):
            arr = getattr(p, k)
            for i in range(3):
                arr[i] = v[i]
        else:
            setattr(p, k, v)
    return p


def _array_from(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


class Scene:
    """Synthetic stand-in for a compiled + lightmapped .sqwm map with its material textures."""

    def __init__(self, params: SceneParams | None = None, **overrides):
        self.params = params if params is not None else default_scene_params(**overrides)
        self._lib = lib()
        self._handle = self._lib.sqwh_scene_create(C.byref(self.params))
        if not self._handle:
            raise ValueError("bad scene parameters")
        self._refresh()

    def _refresh(self):
        """(Re)read what the scene exposes to Python: counts, lightmaps, images."""
        self.info = SceneInfo()
        self._lib.sqwh_scene_get_info(self._handle, C.byref(self.info))
        n = int(self.info.num_lightmap_elements)
        self.lightmaps = _array_from(self._lib.sqwh_scene_lightmaps(self._handle), n * 3, np.float32).reshape(n, 3)
        self.directional_lightmaps = _array_from(self._lib.sqwh_scene_directional_lightmaps(self._handle), n, N.directional_lightmap_dtype)
        self.images = []  # (kind, ndarray[h, w]) in image-id order
        kind, w, h = C.c_int32(), C.c_uint32(), C.c_uint32()
        for i in range(self.info.num_images):
            p = self._lib.sqwh_scene_image(self._handle, i, C.byref(kind), C.byref(w), C.byref(h))
            dt = {N.IMAGE_TEXTURE_ELEMENT: N.texture_element_dtype, N.IMAGE_COLOR32: np.uint32, N.IMAGE_COLOR64: np.uint64}[kind.value]
            self.images.append((kind.value, _array_from(p, w.value * h.value, dt).reshape(h.value, w.value)))

    MAP_FILE_ERRORS = {1: "i/o error", 2: "not a Square-Wheel map (bad id)", 3: "incompatible map version (expected 11)"}

    def save_map(self, path: str):
        """Write the map as .sqwm v11 (bsp_map_save_load.rs::save_map): lightmaps and light grid quantised like the reference."""
        r = self._lib.sqwh_scene_save_map(self._handle, os.fsencode(path))
        if r != 0:
            raise OSError(f"save_map({path}): {self.MAP_FILE_ERRORS.get(r, r)}")

    def load_map(self, path: str):
        """Replace this scene's map by a .sqwm v11 file (bsp_map_save_load.rs::load_map), e.g. one compiled and lit by the
        reference's map_compiler + lightmapper.  Materials, textures and models of the scene are kept: every polygon's texture
        index must have a material.  Close renderers of this scene before calling."""
        r = self._lib.sqwh_scene_load_map(self._handle, os.fsencode(path))
        if r == 4:
            raise ValueError(f"load_map({path}): {self._lib.sqwh_last_map_error().decode()}")
        if r != 0:
            raise OSError(f"load_map({path}): {self.MAP_FILE_ERRORS.get(r, r)}")
        self._refresh()

    def load_compiled_map(self, path: str, materials_table: str, texture_size_log2: int = 8, seed: int = 1):
        """Replace the whole scene by a compiled + lit .sqwm map (tools/map_compiler.py, or the reference's map_compiler +
        lightmapper): one material per texture name of the map, defined by `materials_table` (the text form of the reference's
        materials/*.json, tools/map_compiler.py::materials_table) with a seeded procedural texture of the kind the definition
        asks for; the camera path runs from info_player_start through the info_teleport_destination entities along the leaf
        portals.  Close renderers of this scene before calling."""
        r = self._lib.sqwh_scene_load_compiled_map(self._handle, os.fsencode(path), materials_table.encode("utf-8"), texture_size_log2, seed)
        if r == 4:
            raise ValueError(f"load_compiled_map({path}): {self._lib.sqwh_last_map_error().decode()}")
        if r != 0:
            raise OSError(f"load_compiled_map({path}): {self.MAP_FILE_ERRORS.get(r, r)}")
        self._refresh()

    def texture_name(self, index: int) -> str:
        n = self._lib.sqwh_scene_texture_name(self._handle, index)
        if n is None:
            raise IndexError(index)
        return n.decode()

    def model_info(self, index: int) -> ModelInfo:
        mi = ModelInfo()
        if self._lib.sqwh_scene_model_info(self._handle, index, C.byref(mi)) != 0:
            raise IndexError(index)
        return mi

    def load_model(self, source) -> int:
        """Append a triangle model from an .md3 / .iqm file (path) or from the bytes of one (triangle_model_md3.rs::load_model_md3,
        triangle_model_iqm.rs::load_model_iqm; recognised by its magic).  Returns the model index for model_entity_dtype["model"].
        Raises ValueError with the loader's message where the reference returns Ok(None) / an io::Error."""
        err = C.create_string_buffer(256)
        if isinstance(source, (bytes, bytearray, memoryview)):
            data = bytes(source)
            r = self._lib.sqwh_scene_load_model_memory(self._handle, data, len(data), err, len(err))
        else:
            r = self._lib.sqwh_scene_load_model_file(self._handle, os.fsencode(source), err, len(err))
        if r < 0:
            raise ValueError(f"load_model: {err.value.decode(errors='replace')}")
        self._lib.sqwh_scene_get_info(self._handle, C.byref(self.info))
        return r

    _MODEL_ARRAYS = {
        "triangles": (0, np.dtype("<u2"), 3), "non_animated": (1, np.dtype("<f4"), 8), "constant": (2, np.dtype("<f4"), 2), "variable": (3, np.dtype("<f4"), 6),
        "skeleton": (4, np.dtype([("tex_coord", "<f4", (2,)), ("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("bone_index", "u1", (4,)), ("weight", "u1", (4,))]), None),
        "frame_bones": (5, np.dtype("<f4"), 16), "frames_bbox": (6, np.dtype("<f4"), 6), "bone_parents": (7, np.dtype("<u4"), None), "tc_shift": (8, np.dtype("<f4"), None),
        "animations": (10, np.dtype([("start_frame", "<u4"), ("num_frames", "<u4"), ("frames_per_second", "<f4"), ("looped", "<u4")]), None),
    }

    def model_arrays(self, index: int, mesh: int = 0) -> dict:
        """The arrays of a model as the loaders / generators left them (frame_bones: [frames * bones, 16] column-major 4x4)."""
        out = {}
        for name, (what, dt, cols) in self._MODEL_ARRAYS.items():
            n = self._lib.sqwh_scene_model_copy(self._handle, index, mesh, what, None, 0)
            if n < 0:
                raise IndexError((index, mesh))
            buf = np.zeros(n, np.uint8)
            self._lib.sqwh_scene_model_copy(self._handle, index, mesh, what, buf.ctypes.data_as(C.c_void_p), n)
            a = buf.view(dt)
            out[name] = a.reshape(-1, cols) if cols else a
        n = self._lib.sqwh_scene_model_copy(self._handle, index, mesh, 9, None, 0)
        buf = np.zeros(max(n, 1), np.uint8)
        self._lib.sqwh_scene_model_copy(self._handle, index, mesh, 9, buf.ctypes.data_as(C.c_void_p), n)
        names = bytes(buf[:n]).split(b"\0")[:-1] if n else []
        out["names"] = [x.decode(errors="replace") for x in names]
        return out

    def camera_path(self, num_frames: int, jitter_seed: int = 0) -> np.ndarray:
        """[num_frames, 5] = position xyz, azimuth, elevation (64-pose Catmull-Rom loop of SURVEY.md §8d)."""
        out = np.zeros((num_frames, 5), np.float32)
        n = self._lib.sqwh_scene_camera_path(self._handle, num_frames, jitter_seed, out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def map_geometry(self):
        """(polygons [map_polygon_dtype: the 80-byte Polygon records], vertices [n, 3] f32, per-material SQW_MATERIAL_* flags) for sqw_upload_map_geometry."""
        info = self.info
        pol = _array_from(self._lib.sqwh_scene_polygons(self._handle), int(info.num_polygons), N.map_polygon_dtype).copy()
        ver = _array_from(self._lib.sqwh_scene_vertices(self._handle), int(info.num_vertices) * 3, np.float32).reshape(-1, 3).copy()
        flags = np.zeros(int(info.num_materials), np.uint8)
        n = self._lib.sqwh_scene_material_flags(self._handle, flags.ctypes.data_as(C.c_void_p), len(flags))
        assert n == len(flags)
        return pol, ver, flags

    def map_materials(self) -> np.ndarray:
        """sqw_map_material records of the scene's materials (map_material_dtype) for sqw_upload_map_materials."""
        out = np.zeros(int(self.info.num_materials), N.map_material_dtype)
        n = self._lib.sqwh_scene_map_materials(self._handle, out.ctypes.data_as(C.c_void_p), len(out))
        assert n == len(out)
        return out

    def upload(self, ctx):
        """Upload every image in image-id order (S0 of SURVEY.md §8b)."""
        for i, (kind, px) in enumerate(self.images):
            got = ctx.upload_image(kind, px)
            assert got == i, "image ids must match the host library's numbering"

    def close(self):
        if self._handle:
            self._lib.sqwh_scene_destroy(self._handle)
            self._handle = None


DEMO2_MAP_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "_ref", "demo2.sqwm")
DEMO2_MATERIALS_PATH = os.path.join(os.path.dirname(DEMO2_MAP_PATH), "demo2_materials.txt")


def demo2_scene(texture_size_log2: int = 8, seed: int = 1):
    """The reference's demo map 2 (BASELINE.json north_star) as a Scene, or None where `__graft_entry__.build_demo_map` has not
    produced it (it compiles /root/reference/demo/maps/demo2.map, which exists only in the build container).  Geometry, entities,
    lights and material definitions are the demo's; the texture images are seeded stand-ins (the demo ships procedural sources)."""
    if not (os.path.exists(DEMO2_MAP_PATH) and os.path.exists(DEMO2_MATERIALS_PATH)):
        return None
    with open(DEMO2_MATERIALS_PATH, "r", encoding="utf-8") as f:
        table = f.read()
    sc = Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    sc.load_compiled_map(DEMO2_MAP_PATH, table, texture_size_log2, seed)
    return sc


class HostFrame:
    """Everything the host produced for one frame: the inputs of S1 and S2."""

    def __init__(
        self, jobs, total_surface_pixels, cam_pos, cmds, polygon_points, needs_clear, info, lights=None, depth_maps=None, depth_cmds=None, depth_points=None, shadow_maps_len=0, triangle_points=None
    ):
        self.jobs, self.total_surface_pixels, self.cam_pos = jobs, total_surface_pixels, cam_pos
        self.cmds, self.polygon_points, self.needs_clear, self.info = cmds, polygon_points, needs_clear, info
        self.triangle_points = triangle_points if triangle_points is not None and len(triangle_points) else None  # SQW_CMD_TRIANGLE vertices (models)
        self.portals = []  # PortalView per VISIBLE portal; polygon commands with tex_kind TEX_VIEW carry the portal index in tex_ref
        # dynamic lights of the frame and the shadow-map fills that feed them (empty without FrameInfo::lights)
        self.lights, self.depth_maps, self.depth_cmds, self.depth_points, self.shadow_maps_len = lights, depth_maps, depth_cmds, depth_points, shadow_maps_len

    def record(self, frame, clear_color=None, post=None):
        """Hand everything to one recorded device frame (sqw_frame_set_*)."""
        has_lights = self.lights is not None and len(self.lights) > 0
        frame.set_surfaces(self.jobs, self.lights if has_lights else None, None, self.cam_pos, self.total_surface_pixels, device_shadow_maps_len=self.shadow_maps_len if has_lights else 0)
        if self.depth_maps is not None:
            frame.set_shadow_map_draw(self.depth_maps, self.depth_cmds, self.depth_points)  # an empty list clears an earlier recording
        frame.set_draw(self.cmds, self.polygon_points, self.triangle_points, clear_color=clear_color)
        if post is not None:
            frame.set_post(post)


class HostRenderer:
    """C++ mirror of Renderer / PartialRenderer front end (renderer.rs:61,106; partial_renderer.rs:156,236)."""

    def __init__(self, scene: Scene, use_directional_lightmaps=True, textures_mip_bias=0.0, clear_background=True, shadows_quality=0.0, portals_depth=1):
        self.scene = scene
        self._lib = lib()
        cfg = RendererConfig(1 if use_directional_lightmaps else 0, textures_mip_bias, 1 if clear_background else 0, 0, shadows_quality, portals_depth)
        self._handle = self._lib.sqwh_renderer_create(scene._handle, C.byref(cfg))

    def build_frame(self, width: int, height: int, pose, fov: float = FOV_DEFAULT, lights=None, models=None, color_bits: int = 64, sprites=None, decals=None, submodels=None, portals=None) -> HostFrame:
        """lights: FrameInfo::lights as an array of dynamic_light_dtype (or None); models: FrameInfo::model_entities as an array of
        model_entity_dtype (or None); sprites / decals: FrameInfo::sprites / ::decals as arrays of sprite_dtype / decal_dtype (or None); submodels: FrameInfo::submodel_entities, one submodel_entity_dtype record per map submodel (None keeps the last ones;
        initially every submodel stands where the map has it); portals: FrameInfo::portals as an array of portal_dtype (or None) — the
        nested views of the visible ones come back as HostFrame.portals; color_bits: colour type of the target (only picks the sky box textures, SkyboxSideTexture<ColorT>)."""
        pose = np.ascontiguousarray(pose, np.float32)
        lt = np.ascontiguousarray(lights, dynamic_light_dtype) if lights is not None else np.zeros(0, dynamic_light_dtype)
        self._lib.sqwh_renderer_set_lights(self._handle, lt.ctypes.data_as(C.c_void_p), len(lt))
        md = np.ascontiguousarray(models, model_entity_dtype) if models is not None else np.zeros(0, model_entity_dtype)
        if self._lib.sqwh_renderer_set_models(self._handle, self.scene._handle, md.ctypes.data_as(C.c_void_p), len(md)) != 0:
            raise ValueError("model entity refers to a missing model, skin or frame")
        if submodels is not None:
            se = np.ascontiguousarray(submodels, submodel_entity_dtype)
            if self._lib.sqwh_renderer_set_submodel_entities(self._handle, se.ctypes.data_as(C.c_void_p), len(se)) != 0:
                raise ValueError("one submodel entity per map submodel is required")
        dc = np.ascontiguousarray(decals, decal_dtype) if decals is not None else np.zeros(0, decal_dtype)
        if self._lib.sqwh_renderer_set_decals(self._handle, self.scene._handle, dc.ctypes.data_as(C.c_void_p), len(dc)) != 0:
            raise ValueError("decal refers to a missing skin or has a bad mode")
        sp = np.ascontiguousarray(sprites, sprite_dtype) if sprites is not None else np.zeros(0, sprite_dtype)
        if self._lib.sqwh_renderer_set_sprites(self._handle, self.scene._handle, sp.ctypes.data_as(C.c_void_p), len(sp)) != 0:
            raise ValueError("sprite refers to a missing skin or has a bad mode / radius")
        pt = np.ascontiguousarray(portals, portal_dtype) if portals is not None else np.zeros(0, portal_dtype)
        if self._lib.sqwh_renderer_set_portals(self._handle, self.scene._handle, pt.ctypes.data_as(C.c_void_p), len(pt)) != 0:
            raise ValueError("bad portal (view kind, 3..8 vertices, blending mode, texture skin)")
        self._lib.sqwh_renderer_prepare_frame(self._handle, width, height, pose.ctypes.data_as(C.c_void_p), fov)
        self._lib.sqwh_renderer_draw_frame(self._handle, color_bits)
        return self._collect(self._handle, color_bits)

    def _collect(self, handle, color_bits: int) -> HostFrame:
        """Copy the lists of the frame a (possibly nested) renderer prepared last, then — portal by portal, because a renderer
        re-uses ONE nested renderer for all of its portals (partial_renderer.rs:2032) — prepare and collect the nested views."""
        h = self._handle
        self._handle = handle
        try:
            hf = self._collect_lists()
        finally:
            self._handle = h
        info = PortalInfo()
        for i in range(self._lib.sqwh_renderer_num_portals(handle)):
            self._lib.sqwh_renderer_portal_info(handle, i, C.byref(info))
            if info.resolution[0] * info.resolution[1] == 0:
                continue
            node = PortalView(i, info, None)
            child = self._lib.sqwh_renderer_prepare_portal_view(handle, i, color_bits)
            if child:
                node.frame = self._collect(child, color_bits)
            hf.portals.append(node)
        return hf

    def _collect_lists(self) -> HostFrame:
        fi = FrameInfo()
        self._lib.sqwh_renderer_frame_info(self._handle, C.byref(fi))
        jobs = _array_from(self._lib.sqwh_renderer_jobs(self._handle), fi.num_jobs, N.surface_job_dtype)
        cmds = _array_from(self._lib.sqwh_renderer_cmds(self._handle), fi.num_cmds, N.draw_cmd_dtype)
        pts = _array_from(self._lib.sqwh_renderer_polygon_points(self._handle), fi.num_polygon_points, N.polygon_point_dtype)
        out_lights = _array_from(self._lib.sqwh_renderer_lights(self._handle), fi.num_lights, N.light_dtype)
        dmaps = _array_from(self._lib.sqwh_renderer_depth_maps(self._handle), fi.num_depth_maps, N.depth_map_dtype)
        dcmds = _array_from(self._lib.sqwh_renderer_depth_cmds(self._handle), fi.num_depth_cmds, N.depth_cmd_dtype)
        dpts = _array_from(self._lib.sqwh_renderer_depth_points(self._handle), fi.num_depth_points, N.polygon_point_dtype)
        tpts = _array_from(self._lib.sqwh_renderer_triangle_points(self._handle), fi.num_triangle_points, N.triangle_point_dtype)
        return HostFrame(
            jobs, int(fi.total_surface_pixels), np.array(list(fi.cam_pos), np.float32), cmds, pts, bool(fi.needs_background_clear), fi, out_lights, dmaps, dcmds, dpts, int(fi.shadow_maps_len), tpts
        )

    def dynamic_meshes(self):
        """build_dynamic_models_buffers (partial_renderer.rs:1122-1190) of the frame built last: (records, vertices, triangles) —
        one dynamic_mesh_dtype record per visible mesh with everything the GPU model-preparation kernels take as input
        (sqw_model_instance fields) and this mirror's own results: lit camera-space vertices and sorted front-facing triangles."""
        counts = (C.c_uint64 * 3)()
        self._lib.sqwh_renderer_dynamic_meshes(self._handle, None, 0, None, 0, None, 0, counts)
        rec = np.zeros(int(counts[0]), dynamic_mesh_dtype)
        v = np.zeros(int(counts[1]), np.dtype([("pos", "<f4", 3), ("tc", "<f4", 2), ("light", "<f4", 3)]))
        t = np.zeros((int(counts[2]), 3), np.uint16)
        self._lib.sqwh_renderer_dynamic_meshes(
            self._handle, rec.ctypes.data_as(C.c_void_p), len(rec), v.ctypes.data_as(C.c_void_p), len(v), t.ctypes.data_as(C.c_void_p), len(t), counts
        )
        return rec, v, t

    def prepared_polygons(self):
        """prepare_polygons_surfaces (partial_renderer.rs:1327-1360) of the frame built last, leaf polygons: (candidates
        [polygon_candidate_dtype], clip planes [sets, 8, 3], this mirror's results [prepared_polygon_dtype], frame record) —
        the inputs of sqw_polygons_prepare and what it must reproduce."""
        counts = (C.c_uint64 * 2)()
        self._lib.sqwh_renderer_prepared_polygons(self._handle, None, None, None, None, counts)
        cand = np.zeros(int(counts[0]), N.polygon_candidate_dtype)
        planes = np.zeros((int(counts[1]), 8, 3), np.float32)
        res = np.zeros(int(counts[0]), N.prepared_polygon_dtype)
        frame = np.zeros(1, polygons_frame_dtype)
        vp = C.c_void_p
        self._lib.sqwh_renderer_prepared_polygons(self._handle, cand.ctypes.data_as(vp), planes.ctypes.data_as(vp), res.ctypes.data_as(vp), frame.ctypes.data_as(vp), counts)
        return cand, planes, res, frame[0]

    def polygon_basis(self) -> np.ndarray:
        """PolygonBasisVecs of every map polygon for mip 0, [num_polygons, 12] f32 {u, v, start, normal} (computed once per map)."""
        n = int(self._lib.sqwh_renderer_polygon_basis(self._handle, None, 0))
        out = np.zeros((n, 12), np.float32)
        self._lib.sqwh_renderer_polygon_basis(self._handle, out.ctypes.data_as(C.c_void_p), n)
        return out

    def clip_set_lights(self):
        """(offsets [sets + 1] uint32, light indices uint16): the dynamic lights located in the leaf of every clip set of the last build_frame."""
        counts = np.zeros(2, np.uint32)
        self._lib.sqwh_renderer_clip_set_lights(self._handle, None, None, counts.ctypes.data_as(C.c_void_p))
        off = np.zeros(int(counts[0]), np.uint32)
        lst = np.zeros(max(int(counts[1]), 1), np.uint16)
        self._lib.sqwh_renderer_clip_set_lights(self._handle, off.ctypes.data_as(C.c_void_p), lst.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p))
        return off, lst[: int(counts[1])]

    def polygon_draw_order(self) -> np.ndarray:
        """Map polygon indices of the leaf polygons the last build_frame drew, in draw order (draw_tree / draw_leaf)."""
        n = int(self._lib.sqwh_renderer_polygon_draw_order(self._handle, None, 0))
        out = np.zeros(n, np.uint32)
        self._lib.sqwh_renderer_polygon_draw_order(self._handle, out.ctypes.data_as(C.c_void_p), n)
        return out

    def polygons_prepare_seconds(self) -> float:
        """Wall time the host mirror spent in prepare_polygons_surfaces during the last build_frame."""
        return float(self._lib.sqwh_renderer_polygons_prepare_seconds(self._handle))

    def default_submodel_entities(self) -> np.ndarray:
        """Every submodel present, at the place the map has it (position = bbox centre, identity rotation)."""
        n = int(self.scene.info.num_submodels)
        e = np.zeros(n, submodel_entity_dtype)
        c = np.zeros(3, np.float32)
        for i in range(n):
            self._lib.sqwh_renderer_submodel_center(self._handle, i, c.ctypes.data_as(C.c_void_p))
            e[i]["present"] = 1
            e[i]["position"] = c
            e[i]["rotation"] = (1.0, 0.0, 0.0, 0.0)
        return e

    def game_submodel_entities(self) -> np.ndarray:
        """default_submodel_entities with the inline models hidden that the reference's game does not draw as such
        (test_game/world_spawn.rs:33-55, :331-434): trigger_* volumes, and func_mirror / func_camera_portal / func_parallax_portal,
        whose polygons become view portals (not set up here: BASELINE configs[1] runs with portals_depth 0)."""
        e = self.default_submodel_entities()
        for i in range(len(e)):
            name = (self._lib.sqwh_scene_submodel_classname(self.scene._handle, i) or b"").decode()
            if name.startswith("trigger") or name in ("func_mirror", "func_camera_portal", "func_parallax_portal"):
                e[i]["present"] = 0
        return e

    def close(self):
        if self._handle:
            self._lib.sqwh_renderer_destroy(self._handle)
            self._handle = None


# draw_background colour (partial_renderer.rs:253-256): ColorVec::from_color_f32x3(&[32.0, 16.0, 8.0]).into()
BACKGROUND_COLOR32 = (32 << 16) | (16 << 8) | 8
BACKGROUND_COLOR64 = (32 << 32) | (16 << 16) | 8
