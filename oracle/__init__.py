"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes bridge to oracle/liboracle.so, the CPU restatement of the reference's world-drawing path (see oracle_common.h).
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and --impl reference) may import this package;
nothing under square_wheel_b200/ does.  PARITY UNPINNED: the reference has no golden vectors and cannot be built here (tools/rust_golden/ is the emitter a maintainer with
a Rust toolchain runs to pin it; tests/test_reference_golden_cpu.py consumes its output when present).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
# numerics mode -> library (oracle/Makefile): "portable" is NORMATIVE for the bit-exact parity tests; "sse" restates the branch of
# fast_math.rs the reference's release build runs (rcpss / rsqrtss, cvtps2dq rounding, FMA, pavg) and bounds the tolerance tests;
# "sse_native" is the same compiled with -march=native ON THE MACHINE THAT RUNS IT — the CPU baseline bench.py times.
_LIBS = {"portable": "liboracle.so", "sse": "liboracle_sse.so", "sse_native": "liboracle_sse_native.so"}
_loaded = {}
_default = "portable"


class OracleImage(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("kind", C.c_int32), ("pad", C.c_uint32)]


def build(force: bool = False, numerics: str = "portable"):
    """Compile the restatement with the recipe in oracle/Makefile (g++, -ffp-contract=off -fwrapv)."""
    target = _LIBS[numerics]
    path = os.path.join(_HERE, target)
    if force and os.path.exists(path):
        os.remove(path)
    subprocess.run(["make", "-C", _HERE, target], check=True, capture_output=True)
    return path


def set_default_numerics(numerics: str):
    """Which library the module-level functions use when no `numerics=` is given."""
    global _default
    assert numerics in _LIBS
    _default = numerics


def lib(numerics: str = None) -> C.CDLL:
    numerics = numerics or _default
    if numerics not in _loaded:
        path = os.path.join(_HERE, _LIBS[numerics])
        if not os.path.exists(path) or numerics == "sse_native":
            build(numerics=numerics)  # (make is a no-op when the library is newer than its sources)
        l = C.CDLL(path)
        l.oracle_assert_failures.restype = C.c_uint64
        l.oracle_update_exposure.restype = C.c_float
        l.oracle_update_exposure.argtypes = [C.c_float, C.c_void_p] + [C.c_float] * 7
        _loaded[numerics] = l
    return _loaded[numerics]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _image_table(images):
    """images: list of 2-D numpy arrays (texture_element / uint32 / uint64) with their kind."""
    tbl = (OracleImage * max(len(images), 1))()
    keep = []
    for i, (kind, px) in enumerate(images):
        px = np.ascontiguousarray(px)
        keep.append(px)
        tbl[i] = OracleImage(px.ctypes.data, px.shape[1], px.shape[0], kind, 0)
    return tbl, keep


def assert_failures() -> int:
    return int(lib().oracle_assert_failures())


def reset_assert_failures():
    lib().oracle_reset_assert_failures()


def build_surfaces(color_bits, jobs, lights, shadow_maps, cam_pos, images, lightmaps, directional_lightmaps, total_pixels, out=None, num_threads=1, numerics=None):
    """surfaces.rs build_surface_* + emissive mix for every job; returns the surface pool (uint32/uint64 [total_pixels])."""
    dt = np.uint32 if color_bits == 32 else np.uint64
    if out is None:
        out = np.zeros(total_pixels, dt)
    tbl, keep = _image_table(images)
    cam = np.asarray(cam_pos, np.float32)
    jobs = np.ascontiguousarray(jobs)
    lights = np.ascontiguousarray(lights) if lights is not None and len(lights) else None
    sm = np.ascontiguousarray(shadow_maps, np.float32) if shadow_maps is not None and len(shadow_maps) else None
    rc = lib(numerics).oracle_build_surfaces(
        C.c_int(color_bits), _p(jobs), C.c_uint32(len(jobs)), _p(lights), C.c_uint32(0 if lights is None else len(lights)), _p(sm), _p(cam), tbl,
        C.c_uint32(len(images)), _p(lightmaps), _p(directional_lightmaps), _p(out), C.c_int(num_threads),
    )
    assert rc == 0
    return out


def mix_surface_with_texture(color_bits, surface, surface_size, surface_tc_min, texture, blending_mode, light, numerics=None):
    ss = np.asarray(surface_size, np.uint32)
    tc = np.asarray(surface_tc_min, np.int32)
    tex = np.ascontiguousarray(texture, np.uint32)
    ts = np.asarray([tex.shape[1], tex.shape[0]], np.uint32)
    li = np.asarray(light, np.float32)
    lib(numerics).oracle_mix_surface_with_texture(C.c_int(color_bits), _p(ss), _p(tc), _p(tex), _p(ts), C.c_int(blending_mode), _p(li), _p(surface))
    return surface


def draw(color_bits, framebuffer, cmds, polygon_points, triangle_points, surfaces, images, view_textures=(), clip_rect=None, num_threads=1, numerics=None):
    """rasterizer.rs fill_polygon / fill_triangle for every command, in order, into `framebuffer` ([h, w], in place)."""
    h, w = framebuffer.shape
    tbl, keep = _image_table(images)
    vt = (C.c_void_p * max(len(view_textures), 1))()
    for i, t in enumerate(view_textures):
        vt[i] = None if t is None else t.ctypes.data
    cmds = np.ascontiguousarray(cmds)
    pp = np.ascontiguousarray(polygon_points) if polygon_points is not None and len(polygon_points) else None
    tp = np.ascontiguousarray(triangle_points) if triangle_points is not None and len(triangle_points) else None
    cr = None if clip_rect is None else np.asarray(clip_rect, np.int32)
    rc = lib(numerics).oracle_draw(
        C.c_int(color_bits), _p(framebuffer), C.c_int32(w), C.c_int32(h), C.c_int32(w), _p(cmds), C.c_uint32(len(cmds)), _p(pp), _p(tp), _p(surfaces), tbl,
        C.c_uint32(len(images)), vt, _p(cr), C.c_int(num_threads),
    )
    assert rc == 0
    return framebuffer


def draw_shadow_maps(shadow_maps, maps, cmds, polygon_points):
    """DepthRasterizer::fill_polygon for every recorded fill call, map by map, into `shadow_maps` (float32, in place)."""
    maps = np.ascontiguousarray(maps)
    cmds = np.ascontiguousarray(cmds)
    pp = np.ascontiguousarray(polygon_points)
    rc = lib().oracle_draw_shadow_maps(_p(shadow_maps), _p(maps), C.c_uint32(len(maps)), _p(cmds), C.c_uint32(len(cmds)), _p(pp))
    assert rc == 0
    return shadow_maps


def fill(framebuffer, raw_color):
    lib().oracle_fill(C.c_int(32 if framebuffer.dtype == np.uint32 else 64), _p(framebuffer), C.c_uint64(framebuffer.size), C.c_uint64(raw_color))


def postprocess(hdr, params, num_threads=1, want_bloom=False, dst=None, numerics=None):
    """postprocessor.rs perform_postprocessing. Returns (pixels uint32 [h, w], avg colour [B,G,R], bloom buffer or None).
    `dst`: a reusable picture buffer (the reference writes into the window surface it is handed, postprocessor.rs:68)."""
    h, w = hdr.shape
    if dst is None:
        dst = np.zeros((h, w), np.uint32)
    avg = np.zeros(3, np.float32)
    bloom = None
    if want_bloom and params.bloom_sigma > 0:
        s = 1 << min(max(int(params.bloom_buffer_scale_log2), 1), 3)
        bloom = np.zeros((h // s, w // s), np.uint64)
    rc = lib(numerics).oracle_postprocess(_p(np.ascontiguousarray(hdr)), C.c_uint32(w), C.c_uint32(h), C.byref(params), _p(dst), C.c_size_t(w), _p(avg), _p(bloom), C.c_int(num_threads))
    assert rc == 0
    return dst, avg, bloom
