"""Thin object wrappers over the C-ABI (include/sqw_cuda.h): Context (one map replica on one GPU), View (one render
target), Frame (one recorded frame).  All compute happens inside libsqw_cuda.so; these classes only marshal numpy
arrays into the plain-pointer calls."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as N


class SqwError(RuntimeError):
    pass


def _as(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype) if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"]) else a
    return a


class Context:
    """sqw_ctx: replaces the data-ownership half of Renderer::new (src/square_wheel/renderer.rs:24)."""

    def __init__(
        self, lightmaps: Optional[np.ndarray], directional_lightmaps: Optional[np.ndarray], device: int = 0, stream: int = 0, max_images: int = 0, concurrent_views: bool = False,
        fast_surface_numerics: bool = False, compressed_readback: bool = False,
    ):
        self._lib = N.load_cuda_library()
        self._handle = C.c_void_p()
        lm = _as(lightmaps, np.float32)
        dlm = _as(directional_lightmaps, N.directional_lightmap_dtype)
        n = 0
        if lm is not None:
            n = lm.size // 3
        if dlm is not None:
            if lm is not None and dlm.size != n:
                raise ValueError("lightmap arrays must have the same element count")
            n = dlm.size
        flags = (N.CTX_CONCURRENT_VIEWS if concurrent_views else 0) | (N.CTX_FAST_SURFACE_NUMERICS if fast_surface_numerics else 0) | (N.CTX_COMPRESSED_READBACK if compressed_readback else 0)
        desc = N.MapDesc(N.ptr(lm), N.ptr(dlm), n, max_images, flags)
        rc = self._lib.sqw_create(C.byref(self._handle), device, C.byref(desc), C.c_void_p(stream) if stream else None)
        if rc != N.SQW_OK:
            raise SqwError(f"sqw_create failed ({rc}): {self._lib.sqw_last_error(None).decode()}")
        self.device = device
        self.images = []  # (kind, w, h)

    def _check(self, rc: int, what: str):
        if rc < 0:
            raise SqwError(f"{what} failed ({rc}): {self._lib.sqw_last_error(self._handle).decode()}")
        return rc

    def upload_image(self, kind: int, pixels: np.ndarray) -> int:
        """pixels: 2-D array [height, width] of texture_element_dtype / uint32 / uint64."""
        dt = {N.IMAGE_TEXTURE_ELEMENT: N.texture_element_dtype, N.IMAGE_COLOR32: np.uint32, N.IMAGE_COLOR64: np.uint64}[kind]
        p = _as(pixels, dt)
        h, w = p.shape
        image_id = self._check(self._lib.sqw_upload_image(self._handle, kind, w, h, N.ptr(p)), "sqw_upload_image")
        self.images.append((kind, w, h))
        return image_id

    def update_image(self, image_id: int, pixels: np.ndarray):
        self._check(self._lib.sqw_update_image(self._handle, image_id, N.ptr(np.ascontiguousarray(pixels))), "sqw_update_image")

    def create_view(self, width: int, height: int, color_bits: int) -> "View":
        return View(self, width, height, color_bits)

    def batch_begin(self):
        """Fork the views' private streams from the context's stream (concurrent_views)."""
        self._check(self._lib.sqw_batch_begin(self._handle), "sqw_batch_begin")

    def batch_end(self):
        """Join the views' private streams back into the context's stream."""
        self._check(self._lib.sqw_batch_end(self._handle), "sqw_batch_end")

    def profile_begin(self):
        self._check(self._lib.sqw_profile_begin(self._handle), "sqw_profile_begin")

    def profile_end(self) -> dict:
        """{kernel name: (total ms, launches)} for everything launched since profile_begin."""
        kt = N.KernelTimes()
        self._check(self._lib.sqw_profile_end(self._handle, C.byref(kt)), "sqw_profile_end")
        return {name: (kt.ms[i], int(kt.launches[i])) for i, name in enumerate(N.KERNEL_NAMES)}

    def upload_model_mesh(self, arrays: dict) -> int:
        """One mesh of a triangle model as Scene.model_arrays() returns it (sqw_upload_model_mesh).  Returns the mesh id."""
        tri = np.ascontiguousarray(arrays["triangles"], np.uint16)
        m = N.ModelMesh()
        keep = [tri]
        if len(arrays["non_animated"]):
            v = np.ascontiguousarray(arrays["non_animated"], np.float32)
            m.kind, m.num_vertices, m.num_frames = N.MESH_NON_ANIMATED, len(v), 1
        elif len(arrays["constant"]):
            tc = np.ascontiguousarray(arrays["constant"], np.float32)
            v = np.ascontiguousarray(arrays["variable"], np.float32)
            m.kind, m.num_vertices, m.num_frames = N.MESH_VERTEX_ANIMATED, len(tc), len(v) // len(tc)
            m.tex_coords = tc.ctypes.data
            keep.append(tc)
        else:
            v = np.ascontiguousarray(arrays["skeleton"])
            assert v.dtype.itemsize == 40
            parents = np.ascontiguousarray(arrays["bone_parents"], np.uint32)
            fb = np.ascontiguousarray(arrays["frame_bones"], np.float32)
            m.kind, m.num_vertices, m.num_bones = N.MESH_SKELETON_ANIMATED, len(v), len(parents)
            m.num_frames = max(len(fb) // max(len(parents), 1), 1) if len(fb) else 1
            if len(fb):
                m.bone_parents, m.frame_bones = parents.ctypes.data, fb.ctypes.data
            keep += [parents, fb]
        keep.append(v)
        m.vertices, m.triangles, m.num_triangles = v.ctypes.data, tri.ctypes.data, len(tri)
        return self._check(self._lib.sqw_upload_model_mesh(self._handle, C.byref(m)), "sqw_upload_model_mesh")

    def upload_map_geometry(self, polygons: np.ndarray, vertices: np.ndarray, material_flags: np.ndarray):
        """The map's polygons (80-byte Polygon records), vertices [n, 3] f32 and per-material SQW_MATERIAL_* flags (sqw_upload_map_geometry)."""
        pol = _as(polygons, N.map_polygon_dtype)
        ver = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        fl = np.ascontiguousarray(material_flags, np.uint8)
        self._check(self._lib.sqw_upload_map_geometry(self._handle, N.ptr(pol), len(pol), N.ptr(ver), len(ver), N.ptr(fl), len(fl)), "sqw_upload_map_geometry")

    def upload_map_materials(self, materials: np.ndarray, polygon_basis: np.ndarray):
        """Per-material constants (map_material_dtype) and per-polygon basis vectors [n, 12] f32 of job building (sqw_upload_map_materials)."""
        mat = _as(materials, N.map_material_dtype)
        basis = np.ascontiguousarray(polygon_basis, np.float32).reshape(-1, 12)
        self._check(self._lib.sqw_upload_map_materials(self._handle, N.ptr(mat), len(mat), N.ptr(basis), len(basis)), "sqw_upload_map_materials")

    def set_fast_surface_numerics(self, enable: bool):
        """SQW_CTX_FAST_SURFACE_NUMERICS for the frames submitted from now on (sqw_set_fast_surface_numerics)."""
        self._check(self._lib.sqw_set_fast_surface_numerics(self._handle, 1 if enable else 0), "sqw_set_fast_surface_numerics")

    def set_compressed_readback(self, enable: bool):
        """SQW_CTX_COMPRESSED_READBACK for the read_pixels32_async calls made from now on (sqw_set_compressed_readback)."""
        self._check(self._lib.sqw_set_compressed_readback(self._handle, 1 if enable else 0), "sqw_set_compressed_readback")

    def self_test(self):
        """Exhaustive device check of the arithmetic shortcuts of the kernels (sqw_self_test); raises SqwError on a mismatch."""
        self._check(self._lib.sqw_self_test(self._handle), "sqw_self_test")

    def close(self):
        if self._handle:
            self._lib.sqw_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class View:
    """sqw_view: one render target + per-view scratch (a PartialRenderer, partial_renderer.rs:15-50)."""

    def __init__(self, ctx: Context, width: int, height: int, color_bits: int):
        self.ctx = ctx
        self._lib = ctx._lib
        self._handle = C.c_void_p()
        ctx._check(self._lib.sqw_view_create(ctx._handle, C.byref(self._handle), width, height, color_bits), "sqw_view_create")
        self.width, self.height, self.color_bits = width, height, color_bits
        self.pixel_dtype = np.uint32 if color_bits == 32 else np.uint64

    # ---- immediate-mode stage calls (S1, S2, S3 of SURVEY.md §8b) ----
    def build_surfaces(self, jobs, lights=None, shadow_maps=None, cam_pos=(0.0, 0.0, 0.0), total_surface_pixels: Optional[int] = None):
        jobs = _as(jobs, N.surface_job_dtype)
        lights = _as(lights, N.light_dtype) if lights is not None else np.zeros(0, N.light_dtype)
        sm = _as(shadow_maps, np.float32) if shadow_maps is not None else np.zeros(0, np.float32)
        cp = np.asarray(cam_pos, np.float32)
        if total_surface_pixels is None:
            total_surface_pixels = required_surface_pixels(jobs)
        self.ctx._check(
            self._lib.sqw_build_surfaces(self._handle, N.ptr(jobs), len(jobs), N.ptr(lights), len(lights), N.ptr(sm), sm.size, N.ptr(cp), total_surface_pixels),
            "sqw_build_surfaces",
        )

    def draw(self, cmds, polygon_points=None, triangle_points=None, clear_color: Optional[int] = None):
        cmds = _as(cmds, N.draw_cmd_dtype)
        pp = _as(polygon_points, N.polygon_point_dtype) if polygon_points is not None else np.zeros(0, N.polygon_point_dtype)
        tp = _as(triangle_points, N.triangle_point_dtype) if triangle_points is not None else np.zeros(0, N.triangle_point_dtype)
        cc = np.array([clear_color], np.uint64) if clear_color is not None else None
        self.ctx._check(self._lib.sqw_draw(self._handle, N.ptr(cmds), len(cmds), N.ptr(pp), len(pp), N.ptr(tp), len(tp), N.ptr(cc)), "sqw_draw")

    def postprocess(self, params: N.PostParams, read_back: bool = True):
        """Returns (pixels or None, average colour [B, G, R])."""
        dst = np.empty((self.height, self.width), np.uint32) if read_back else None
        avg = np.zeros(3, np.float32)
        self.ctx._check(self._lib.sqw_postprocess(self._handle, C.byref(params), N.ptr(dst), self.width, N.ptr(avg)), "sqw_postprocess")
        return dst, avg

    def set_size(self, width: int, height: int):
        """Logical size within the allocated one (portal textures, partial_renderer.rs:2130-2182)."""
        self.ctx._check(self._lib.sqw_view_set_size(self._handle, width, height), "sqw_view_set_size")
        self.width, self.height = width, height

    @property
    def index(self) -> int:
        return self.ctx._check(self._lib.sqw_view_index(self._handle), "sqw_view_index")

    def mix_texture(self, image_id: int, surface_tc_min, blending_mode: int, light):
        """mix_surface_with_texture over the whole colour buffer (portal / mirror glass overlay, surfaces.rs:854)."""
        tc = np.asarray(surface_tc_min, np.int32)
        li = np.asarray(light, np.float32)
        self.ctx._check(self._lib.sqw_view_mix_texture(self._handle, image_id, N.ptr(tc), blending_mode, N.ptr(li)), "sqw_view_mix_texture")

    # ---- readback ----
    def read_pixels(self) -> np.ndarray:
        out = np.empty((self.height, self.width), self.pixel_dtype)
        self.ctx._check(self._lib.sqw_read_pixels(self._handle, N.ptr(out), self.width), "sqw_read_pixels")
        return out

    def write_pixels(self, pixels: np.ndarray):
        px = _as(pixels, self.pixel_dtype)
        assert px.shape == (self.height, self.width)
        self.ctx._check(self._lib.sqw_write_pixels(self._handle, N.ptr(px), self.width), "sqw_write_pixels")

    def read_pixels32(self) -> np.ndarray:
        out = np.empty((self.height, self.width), np.uint32)
        self.ctx._check(self._lib.sqw_read_pixels32(self._handle, N.ptr(out), self.width), "sqw_read_pixels32")
        return out

    def read_pixels32_async(self, dst_ptr: int, dst_pitch: int):
        """Queue the copy of the post-processed picture into (page-locked) host memory at dst_ptr; complete after synchronize()."""
        self.ctx._check(self._lib.sqw_read_pixels32_async(self._handle, C.c_void_p(dst_ptr), dst_pitch), "sqw_read_pixels32_async")

    def readback_stats(self):
        """(compressed read-backs completed so far, bytes they moved over the bus) (sqw_view_get_readback_stats)."""
        out = np.zeros(2, np.uint64)
        self.ctx._check(self._lib.sqw_view_get_readback_stats(self._handle, N.ptr(out)), "sqw_view_get_readback_stats")
        return int(out[0]), int(out[1])

    def read_pixels_async(self, dst_ptr: int, dst_pitch: int):
        self.ctx._check(self._lib.sqw_read_pixels_async(self._handle, C.c_void_p(dst_ptr), dst_pitch), "sqw_read_pixels_async")

    def read_surfaces(self, offset: int, count: int) -> np.ndarray:
        out = np.empty(count, self.pixel_dtype)
        self.ctx._check(self._lib.sqw_read_surfaces(self._handle, N.ptr(out), offset, count), "sqw_read_surfaces")
        return out

    def models_prepare(self, instances: np.ndarray):
        """build_dynamic_models_buffers on the GPU (sqw_models_prepare): animation, transform, lighting, back-face rejection, sort."""
        inst = _as(instances, N.model_instance_dtype)
        self.ctx._check(self._lib.sqw_models_prepare(self._handle, N.ptr(inst), len(inst)), "sqw_models_prepare")

    def models_read(self, num_vertices: int, num_triangles: int, num_instances: int):
        """(vertices [model_vertex_dtype], triangles [n, 3] uint16, visible-triangle counts) of the last models_prepare."""
        v = np.zeros(num_vertices, N.model_vertex_dtype)
        t = np.zeros((num_triangles, 3), np.uint16)
        c = np.zeros(num_instances, np.uint32)
        self.ctx._check(self._lib.sqw_models_read(self._handle, N.ptr(v), num_vertices, N.ptr(t), num_triangles, N.ptr(c), num_instances), "sqw_models_read")
        return v, t, c

    def polygons_prepare(self, planes_matrix, view_matrix, mip_bias: float, candidates: np.ndarray, clip_planes: np.ndarray):
        """prepare_polygon_surface for every (polygon, leaf clip planes) candidate on the GPU (sqw_polygons_prepare)."""
        fr = N.PolygonsFrame()
        fr.planes_matrix[:] = [float(x) for x in np.asarray(planes_matrix, np.float32).reshape(16)]
        fr.view_matrix[:] = [float(x) for x in np.asarray(view_matrix, np.float32).reshape(16)]
        fr.mip_bias = float(mip_bias)
        cand = _as(candidates, N.polygon_candidate_dtype)
        planes = np.ascontiguousarray(clip_planes, np.float32).reshape(-1, 8, 3)
        self.ctx._check(self._lib.sqw_polygons_prepare(self._handle, C.byref(fr), N.ptr(cand), len(cand), N.ptr(planes), len(planes)), "sqw_polygons_prepare")

    def polygons_read(self, num: int):
        """(records [prepared_polygon_dtype] in candidate order, totals) of the last polygons_prepare."""
        out = np.zeros(num, N.prepared_polygon_dtype)
        totals = N.PolygonsTotals()
        self.ctx._check(self._lib.sqw_polygons_read(self._handle, N.ptr(out), num, C.byref(totals)), "sqw_polygons_read")
        return out, totals

    def polygons_draw(self, order: np.ndarray):
        """draw_polygon for the prepared polygons in BSP draw order (sqw_polygons_draw); order = candidate indices."""
        o = np.ascontiguousarray(order, np.uint32)
        self.ctx._check(self._lib.sqw_polygons_draw(self._handle, N.ptr(o), len(o)), "sqw_polygons_draw")

    def polygons_draw_read(self):
        """(commands [draw_cmd_dtype], vertices [polygon_point_dtype]) of the last polygons_draw, in draw order."""
        counts = np.zeros(2, np.uint32)
        self.ctx._check(self._lib.sqw_polygons_draw_read(self._handle, None, 0, None, 0, N.ptr(counts)), "sqw_polygons_draw_read")
        cmds = np.zeros(int(counts[0]), N.draw_cmd_dtype)
        pts = np.zeros(int(counts[1]), N.polygon_point_dtype)
        self.ctx._check(self._lib.sqw_polygons_draw_read(self._handle, N.ptr(cmds), len(cmds), N.ptr(pts), len(pts), N.ptr(counts)), "sqw_polygons_draw_read")
        return cmds, pts

    def average_color(self) -> np.ndarray:
        avg = np.zeros(3, np.float32)
        self.ctx._check(self._lib.sqw_view_get_average_color(self._handle, N.ptr(avg)), "sqw_view_get_average_color")
        return avg

    def synchronize(self):
        self.ctx._check(self._lib.sqw_view_synchronize(self._handle), "sqw_view_synchronize")

    def create_frame(self) -> "Frame":
        return Frame(self)

    def close(self):
        if self._handle and self.ctx._handle:
            self._lib.sqw_view_destroy(self._handle)
        self._handle = C.c_void_p()


class Frame:
    """sqw_frame: device copies of one view's job list, command list and post params (SURVEY.md §8e batched views)."""

    def __init__(self, view: View):
        self.view = view
        self._lib = view._lib
        self._handle = C.c_void_p()
        view.ctx._check(self._lib.sqw_frame_create(view._handle, C.byref(self._handle)), "sqw_frame_create")
        self.h2d_bytes = 0

    def set_surfaces(self, jobs, lights=None, shadow_maps=None, cam_pos=(0.0, 0.0, 0.0), total_surface_pixels: Optional[int] = None, device_shadow_maps_len: int = 0):
        """device_shadow_maps_len > 0 (with shadow_maps=None): the shadow map array lives on the device only and is filled by
        the shadow map stage (set_shadow_map_draw)."""
        jobs = _as(jobs, N.surface_job_dtype)
        lights = _as(lights, N.light_dtype) if lights is not None else np.zeros(0, N.light_dtype)
        sm = _as(shadow_maps, np.float32) if shadow_maps is not None else np.zeros(0, np.float32)
        cp = np.asarray(cam_pos, np.float32)
        if total_surface_pixels is None:
            total_surface_pixels = required_surface_pixels(jobs)
        sm_ptr, sm_len = (N.ptr(sm), sm.size) if shadow_maps is not None or device_shadow_maps_len == 0 else (None, device_shadow_maps_len)
        self.view.ctx._check(
            self._lib.sqw_frame_set_surfaces(self._handle, N.ptr(jobs), len(jobs), N.ptr(lights), len(lights), sm_ptr, sm_len, N.ptr(cp), total_surface_pixels),
            "sqw_frame_set_surfaces",
        )
        self.h2d_bytes += jobs.nbytes + lights.nbytes + sm.nbytes

    def set_draw(self, cmds, polygon_points=None, triangle_points=None, clear_color: Optional[int] = None):
        cmds = _as(cmds, N.draw_cmd_dtype)
        pp = _as(polygon_points, N.polygon_point_dtype) if polygon_points is not None else np.zeros(0, N.polygon_point_dtype)
        tp = _as(triangle_points, N.triangle_point_dtype) if triangle_points is not None else np.zeros(0, N.triangle_point_dtype)
        cc = np.array([clear_color], np.uint64) if clear_color is not None else None
        self.view.ctx._check(
            self._lib.sqw_frame_set_draw(self._handle, N.ptr(cmds), len(cmds), N.ptr(pp), len(pp), N.ptr(tp), len(tp), N.ptr(cc)), "sqw_frame_set_draw"
        )
        self.h2d_bytes += cmds.nbytes + pp.nbytes + tp.nbytes

    def set_surfaces_from_polygons(self, cam_pos, use_directional_lightmaps: bool = True, lights=None, clip_set_light_offsets=None, clip_set_lights=None, shadow_maps_len: int = 0):
        """The surface half of the frame from the records of the view's last polygons_prepare (sqw_frame_set_surfaces_from_polygons);
        lights + the per-clip-set light lists select the dynamic lights of every polygon on the device."""
        cam = np.ascontiguousarray(cam_pos, np.float32)
        li = _as(lights, N.light_dtype) if lights is not None and len(lights) else None
        off = np.ascontiguousarray(clip_set_light_offsets, np.uint32) if li is not None else None
        lst = np.ascontiguousarray(clip_set_lights, np.uint16) if li is not None else None
        self.view.ctx._check(
            self._lib.sqw_frame_set_surfaces_from_polygons(
                self._handle, N.ptr(cam), 1 if use_directional_lightmaps else 0, N.ptr(li), len(li) if li is not None else 0, N.ptr(off), N.ptr(lst),
                (len(off) - 1) if off is not None else 0, shadow_maps_len if li is not None else 0),
            "sqw_frame_set_surfaces_from_polygons")

    def set_draw_from_polygons(self, clear_color: Optional[int] = None):
        """The draw half of the frame from the device lists of the view's last polygons_draw (sqw_frame_set_draw_from_polygons)."""
        cc = np.array([clear_color], np.uint64) if clear_color is not None else None
        self.view.ctx._check(self._lib.sqw_frame_set_draw_from_polygons(self._handle, N.ptr(cc)), "sqw_frame_set_draw_from_polygons")

    def set_shadow_map_draw(self, maps, cmds, polygon_points):
        """Record the shadow-map fills (DepthRasterizer::fill_polygon calls, rasterizer.rs:1314) of the frame."""
        maps = _as(maps, N.depth_map_dtype)
        cmds = _as(cmds, N.depth_cmd_dtype)
        pp = _as(polygon_points, N.polygon_point_dtype) if polygon_points is not None else np.zeros(0, N.polygon_point_dtype)
        self.view.ctx._check(
            self._lib.sqw_frame_set_shadow_map_draw(self._handle, N.ptr(maps), len(maps), N.ptr(cmds), len(cmds), N.ptr(pp), len(pp)), "sqw_frame_set_shadow_map_draw"
        )
        self.h2d_bytes += maps.nbytes + cmds.nbytes + pp.nbytes

    def read_shadow_maps(self, offset: int, count: int) -> np.ndarray:
        out = np.empty(count, np.float32)
        self.view.ctx._check(self._lib.sqw_frame_read_shadow_maps(self._handle, N.ptr(out), offset, count), "sqw_frame_read_shadow_maps")
        return out

    def set_post(self, params: N.PostParams):
        self.view.ctx._check(self._lib.sqw_frame_set_post(self._handle, C.byref(params)), "sqw_frame_set_post")

    def submit(self, stages: int = N.STAGE_ALL):
        self.view.ctx._check(self._lib.sqw_frame_submit(self._handle, stages), "sqw_frame_submit")

    def stats(self) -> N.FrameStats:
        st = N.FrameStats()
        self.view.ctx._check(self._lib.sqw_frame_get_stats(self._handle, C.byref(st)), "sqw_frame_get_stats")
        return st

    def close(self):
        if self._handle and self.view._handle and self.view.ctx._handle:
            self._lib.sqw_frame_destroy(self._handle)
        self._handle = C.c_void_p()


def required_surface_pixels(jobs: np.ndarray) -> int:
    if len(jobs) == 0:
        return 0
    sizes = jobs["surface_size"].astype(np.int64)
    return int((jobs["surface_pixels_offset"].astype(np.int64) + sizes[:, 0] * sizes[:, 1]).max())


def make_post_params(exposure=1.0, color_modulate=(1.0, 1.0, 1.0), bloom_sigma=0.0, bloom_scale=0.25, bloom_buffer_scale_log2=1, linear_bloom_filter=False) -> N.PostParams:
    """Defaults follow postprocessor_config.rs:5-93 except bloom_sigma (1.0 there; 0 = tonemap only)."""
    p = N.PostParams()
    p.exposure = exposure
    p.color_modulate = (C.c_float * 3)(*color_modulate)
    p.bloom_sigma = bloom_sigma
    p.bloom_scale = bloom_scale
    p.bloom_buffer_scale_log2 = bloom_buffer_scale_log2
    p.linear_bloom_filter = 1 if linear_bloom_filter else 0
    return p
