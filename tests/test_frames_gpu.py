"""Whole-frame parity through the reference-facing API (Renderer.prepare_frame / draw_frame, Postprocessor.perform_postprocessing):
host pipeline -> C-ABI -> kernels, against the oracle fed with the same job and command lists.

Covers BASELINE.json configs[0] (1280x720, simple lightmaps, 32-bit LDR) and configs[1] (directional lightmaps + normal
mapping + specular, 64-bit HDR + tonemapping) on the synthetic stand-in map, at sizes the oracle finishes in seconds, plus a
full-size 4K frame checked through a size-independent property."""
import os

import numpy as np
import pytest

from common import assert_pixels_equal
from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S
from square_wheel_b200.device import make_post_params
from square_wheel_b200.renderer import FrameInfo, Postprocessor, PostprocessorConfig, Renderer, RendererConfig, SurfaceInfo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    sc = S.Scene(seed=7, texture_size_log2=7, num_materials=12)
    yield sc
    sc.close()


def _oracle_frame(oracle_lib, sc, hf, W, H, color_bits):
    surf = oracle_lib.build_surfaces(color_bits, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
    fb = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
    if hf.needs_clear:
        oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
    oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
    return surf, fb


def test_config0_ldr_simple_lightmaps(scene, oracle_lib):
    """configs[0]: 1280x720, simple lightmaps, 32-bit LDR, frames of the fixed camera path."""
    W, H = 1280, 720
    r = Renderer(scene, RendererConfig(use_directional_lightmaps=False))
    try:
        path = scene.camera_path(64)
        for k in (0, 13, 29, 47):
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[k]), color_bits=32)
            pixels = np.zeros((H, W), np.uint32)
            r.draw_frame(pixels)
            oracle_lib.reset_assert_failures()
            surf, fb = _oracle_frame(oracle_lib, scene, hf, W, H, 32)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(r._view.read_surfaces(0, hf.total_surface_pixels), surf, f"frame {k} surfaces")
            assert_pixels_equal(pixels, fb, f"frame {k} picture")
    finally:
        r.close()


@pytest.mark.parametrize("bloom_sigma", [0.0, 1.0])
def test_config1_hdr_directional_specular_tonemap(scene, oracle_lib, bloom_sigma):
    """configs[1] at 960x540: directional lightmaps + normal maps + dielectric/metal specular, 64-bit HDR, tonemapping
    (variant B with the default bloom), auto exposure carried from frame to frame like Host::process_frame does."""
    W, H = 960, 540
    r = Renderer(scene, RendererConfig(use_directional_lightmaps=True))
    post = Postprocessor(r, PostprocessorConfig(bloom_sigma=bloom_sigma))
    exposure = 1.0
    try:
        path = scene.camera_path(64)
        for k in (3, 22, 40):
            post.get_hdr_buffer((W, H))
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[k]), color_bits=64)
            view = r.draw_frame(None)
            hdr = view.read_pixels()
            ldr = np.zeros((H, W), np.uint32)
            assert abs(post.current_exposure - exposure) < 1e-6 * exposure
            avg = post.perform_postprocessing(ldr, SurfaceInfo(W, H), (1.0, 1.0, 1.0), 1.0 / 60.0)
            surf, fb = _oracle_frame(oracle_lib, scene, hf, W, H, 64)
            assert_pixels_equal(hdr, fb, f"frame {k} HDR picture")
            p = make_post_params(exposure, (1.0, 1.0, 1.0), bloom_sigma, 0.25, 1, False)
            want_ldr, want_avg, _ = oracle_lib.postprocess(fb, p)
            assert_pixels_equal(ldr, want_ldr, f"frame {k} tonemapped picture")
            np.testing.assert_allclose(avg, want_avg, rtol=1e-5, atol=1e-3)
            exposure = float(
                oracle_lib.lib().oracle_update_exposure(np.float32(exposure), want_avg.ctypes.data, np.float32(1.0 / 60.0), 1.0, 0.25, 1.0, 1.0 / 16.0, 16.0, 2.0)
            )
            assert abs(post.current_exposure - exposure) < 1e-4 * exposure
    finally:
        r.close()


def test_4k_frame_properties(scene, oracle_lib):
    """BASELINE.json full size (3840x2160 HDR).  The oracle needs seconds for a 4K frame, so this checks it once against the
    oracle on a 1/8-height horizontal strip via the clip rect, and checks idempotence of a second submission for the rest."""
    W, H = 3840, 2160
    r = Renderer(scene, RendererConfig(use_directional_lightmaps=True))
    try:
        path = scene.camera_path(64)
        hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[10]), color_bits=64)
        view = r.draw_frame(None)
        first = view.read_pixels()
        r.draw_frame(None)
        second = view.read_pixels()
        assert_pixels_equal(second, first, "re-drawing the same command list is idempotent")
        surf = oracle_lib.build_surfaces(64, hf.jobs, None, None, hf.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf.total_surface_pixels, num_threads=4)
        assert_pixels_equal(view.read_surfaces(0, hf.total_surface_pixels), surf, "4K surfaces")
        # Rows only: a y-only clip rect does not change span end points, so the strip is exactly the single-threaded picture.
        fb = np.zeros((H, W), np.uint64)
        y0, y1 = 1000, 1270
        oracle_lib.draw(64, fb, hf.cmds, hf.polygon_points, None, surf, scene.images, clip_rect=(0, y0, W, y1))
        if hf.needs_clear:
            fb[fb == 0] = S.BACKGROUND_COLOR64
        assert_pixels_equal(first[y0 + 1 : y1], fb[y0 + 1 : y1], "4K strip")
    finally:
        r.close()


def _lights_along_path(path, n, seed):
    r = np.random.default_rng(seed)
    lights = np.zeros(n, S.dynamic_light_dtype)
    for i in range(n):
        lights[i]["position"] = path[(i * 5) % len(path)][:3] + r.uniform(-30.0, 30.0, 3)
        lights[i]["radius"] = (128.0, 256.0, 512.0)[i % 3]
        lights[i]["color"] = r.uniform(4000.0, 60000.0, 3)
        lights[i]["shadow_type"] = (N.SHADOW_NONE, N.SHADOW_CUBE, N.SHADOW_PROJECTOR)[i % 3]
        # projector looking along a random horizontal direction (rotation about z), 0.2*pi opening like the test game's flashlight
        a = r.uniform(0, 2 * np.pi)
        lights[i]["rotation"] = (np.cos(a / 2), 0.0, 0.0, np.sin(a / 2))
        lights[i]["fov"] = 0.2 * np.pi
    return lights


@pytest.mark.parametrize("color_bits", [32, 64])
def test_config2_dynamic_lights_with_gpu_shadow_maps(scene, oracle_lib, color_bits):
    """configs[2] in miniature: 8 dynamic lights (none / cube / projector shadows).  The host records the shadow-map fills
    (DepthRenderer::draw_map per cube side / projector) and the light lists per polygon; shadow maps, surfaces and the picture
    are produced on the GPU and must equal the oracle's, which renders its own shadow maps from the same fill list."""
    W, H = 640, 360
    r = Renderer(scene, RendererConfig(use_directional_lightmaps=True))
    try:
        path = scene.camera_path(64)
        lights = _lights_along_path(path, 8, 3)
        lit_jobs = 0
        for k in (0, 11, 26, 40):
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[k], lights=lights), color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            lit_jobs += int((hf.jobs["num_lights"] > 0).sum())
            sm = np.zeros(max(hf.shadow_maps_len, 1), np.float32)
            if len(hf.depth_maps):
                oracle_lib.draw_shadow_maps(sm, hf.depth_maps, hf.depth_cmds, hf.depth_points)
                got_sm = r._frame.read_shadow_maps(0, hf.shadow_maps_len)
                assert got_sm.view(np.uint32).tolist() == sm[: hf.shadow_maps_len].view(np.uint32).tolist(), f"frame {k} shadow maps"
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, hf.lights, sm, hf.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, None, surf, scene.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(r._view.read_surfaces(0, hf.total_surface_pixels), surf, f"frame {k} surfaces")
            assert_pixels_equal(pixels, fb, f"frame {k} picture")
        assert lit_jobs > 20  # the lights really reach visible polygons
    finally:
        r.close()


def _models_ahead_of_camera(sc, pose, k, light_scale=0.7):
    """One entity per model fixture, fanned out in front of the camera (forward = (-sin az, cos az, 0)), all blend modes and the
    three ModelLighting kinds, animation frames advancing with k."""
    n = int(sc.info.num_models)
    models = np.zeros(n, S.model_entity_dtype)
    az = float(pose[3])
    for i in range(n):
        a = az + (i - 0.5 * (n - 1)) * 0.2
        models[i]["position"] = pose[:3] + (70.0 + 14.0 * i) * np.array([-np.sin(a), np.cos(a), 0.0], np.float32) + np.array([0.0, 0.0, -10.0 + 5.0 * (i % 3)], np.float32)
        models[i]["rotation"] = (np.cos(0.3 * i + 0.1 * k), 0.0, np.sin(0.2 * i) * np.sin(0.3 * i + 0.1 * k), np.cos(0.2 * i) * np.sin(0.3 * i + 0.1 * k))
        nf = sc.model_info(i).num_frames
        models[i]["frames"] = (k % nf, (k + 1) % nf)
        models[i]["lerp"] = (0.4, 0.0, 1.0, 0.75)[i % 4]
        models[i]["model"] = i
        models[i]["skin"] = i
        models[i]["blending_mode"] = (N.BLEND_NONE, N.BLEND_NONE, N.BLEND_ALPHA_TEST, N.BLEND_ALPHA_BLEND, N.BLEND_AVERAGE, N.BLEND_ADDITIVE)[i % 6]
        models[i]["lighting"] = (S.MODEL_LIGHTING_CONSTANT, S.MODEL_LIGHTING_DEFAULT, S.MODEL_LIGHTING_ADVANCED)[i % 3]
        models[i]["light"] = (light_scale, light_scale * 0.8, light_scale * 0.6)
        models[i]["grid_light_scale"] = 0.5
        models[i]["light_position"] = models[i]["position"]
    return models


@pytest.mark.parametrize("color_bits", [32, 64])
def test_config2_triangle_models(oracle_lib, color_bits):
    """configs[2], the model part, in miniature: non-animated, vertex-animated (MD3-class) and skeleton-animated (IQM-class)
    fixtures prepared by the host mirror (animation, light-cube lighting incl. shadow-less dynamic lights, back-face rejection,
    depth sort, clipping by leaf and screen planes) arrive as thousands of small Gouraud triangles interleaved with the world
    polygons in BSP order; the picture must equal the oracle's fed with the same lists."""
    W, H = 640, 360
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_models=8, model_detail=9)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        path = sc.camera_path(64)
        triangles = 0
        for k in (0, 9, 31, 50):
            lights = _lights_along_path(path, 4, 5)
            lights["shadow_type"] = N.SHADOW_NONE
            lights["position"] = path[k][:3] + np.array([[20, 30, 10], [-40, 60, 0], [10, 90, -20], [0, 10, 30]], np.float32)
            models = _models_ahead_of_camera(sc, path[k], k)
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[k], lights=lights, model_entities=models), color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            triangles += int((hf.cmds["kind"] == N.CMD_TRIANGLE).sum())
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, hf.lights, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(pixels, fb, f"frame {k} picture with models")
        assert triangles > 1500  # the models really are on screen
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_models_loaded_from_md3_and_iqm_files(oracle_lib, color_bits):
    """triangle models that come from files (triangle_model_md3.rs / triangle_model_iqm.rs loaders of the host mirror,
    fed with files written by tests/model_files.py) go through the same animation / lighting / clipping / sorting mirror and the GPU
    rasteriser as the seeded fixtures: pictures equal the oracle's, frame pairs and lerp factors varying."""
    import model_files as MF

    W, H = 480, 270
    sc = S.Scene(seed=9, texture_size_log2=6, num_materials=8, num_models=2, model_detail=4)
    rng = np.random.default_rng(41)
    i_md3 = sc.load_model(MF.barrel_md3_file(rng, num_frames=3))
    i_iqm = sc.load_model(MF.tube_iqm_file(rng, rings=9, segments=10, num_joints=4, num_frames=6)[1])
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        path = sc.camera_path(32)
        triangles = 0
        for k in (2, 11, 20):
            pose = path[k]
            fwd = np.array([-np.sin(pose[3]), np.cos(pose[3]), 0.0], np.float32)
            side = np.array([np.cos(pose[3]), np.sin(pose[3]), 0.0], np.float32)
            models = np.zeros(2, S.model_entity_dtype)
            for m, (mi, nf) in enumerate(((i_md3, 3), (i_iqm, 6))):
                models[m]["position"] = pose[:3] + 55.0 * fwd + (18.0 * m - 9.0) * side + np.array([0, 0, -18.0], np.float32)
                models[m]["rotation"] = (np.cos(0.2 * k), 0.0, 0.0, np.sin(0.2 * k))
                models[m]["frames"] = (k % nf, (k + 1) % nf)
                models[m]["lerp"] = (0.3, 1.0, 0.0)[k % 3]
                models[m]["model"] = mi
                models[m]["skin"] = m
                models[m]["blending_mode"] = (N.BLEND_NONE, N.BLEND_ALPHA_BLEND)[(m + k) % 2]
                models[m]["lighting"] = (S.MODEL_LIGHTING_CONSTANT, S.MODEL_LIGHTING_DEFAULT)[m]
                models[m]["light"] = (0.9, 0.8, 0.7)
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(pose, model_entities=models), color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            triangles += int((hf.cmds["kind"] == N.CMD_TRIANGLE).sum())
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, hf.lights, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(pixels, fb, f"frame {k} picture with file models")
        assert triangles > 150  # the loaded models really are on screen
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_config3_skybox_and_blended_models(oracle_lib, color_bits):
    """configs[3] in miniature: sky box seen through the window of the sky-material polygons (six sides, own mip chain, 32- and
    64-bit SkyboxSideTexture, drawn before the BSP walk) plus alpha-tested / alpha-blended / average / additive models on top,
    tonemapped with bloom in the 64-bit case (the 64-bit sky holds a sun disc far above LDR white)."""
    W, H = 640, 360
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, skybox=1, sky_size_log2=7, num_models=6, model_detail=8)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    post = Postprocessor(r, PostprocessorConfig(bloom_sigma=1.0)) if color_bits == 64 else None
    try:
        path = sc.camera_path(64)
        sky_cmds = 0
        for k in (25, 26, 40, 3):
            pose = path[k].copy()
            pose[4] += 0.5  # look up: the sky windows are ceilings
            if post:
                post.get_hdr_buffer((W, H))
            models = _models_ahead_of_camera(sc, pose, k, light_scale=0.45)
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(pose, model_entities=models), color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            view = r.draw_frame(pixels)
            sky_cmds += int(((hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["tex_kind"] == N.TEX_IMAGE)).sum())
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(pixels, fb, f"frame {k} picture with sky")
            if post:
                ldr = np.zeros((H, W), np.uint32)
                exposure = post.current_exposure
                post.perform_postprocessing(ldr, SurfaceInfo(W, H), (1.0, 1.0, 1.0), 1.0 / 60.0)
                p = make_post_params(exposure, (1.0, 1.0, 1.0), 1.0, 0.25, 1, False)
                want_ldr, _, _ = oracle_lib.postprocess(fb, p)
                assert_pixels_equal(ldr, want_ldr, f"frame {k} tonemapped picture with sky")
        assert sky_cmds >= 4  # sky sides really were drawn
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_config3_decals_and_sprites(oracle_lib, color_bits):
    """configs[3] in miniature, second half: decals (the polygons inside a decal box, clipped by its six planes, fan-triangulated
    and subdivided until affine texturing is accurate, lit through the polygon's lightmap) and sprites (four orientations,
    tesselated when close) — all blend modes, interleaved with world polygons and models in the reference's leaf order."""
    from common import scatter_decals, sprites_ahead_of_camera

    W, H = 640, 360
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_models=4, model_detail=6)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        path = sc.camera_path(64)
        triangles = 0
        for k in (0, 12, 33):
            lights = _lights_along_path(path, 3, 9)
            lights["shadow_type"] = N.SHADOW_NONE
            lights["position"] = path[k][:3] + np.array([[20, 30, 10], [-40, 60, 0], [10, 90, -20]], np.float32)
            fi = FrameInfo(
                path[k], lights=lights, model_entities=_models_ahead_of_camera(sc, path[k], k)[:2], sprites=sprites_ahead_of_camera(S, path[k], 8, light_scale=0.5),
                decals=scatter_decals(S, path[k], 48, k),
            )
            hf = r.prepare_frame(SurfaceInfo(W, H), fi, color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            triangles += int((hf.cmds["kind"] == N.CMD_TRIANGLE).sum())
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, hf.lights, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(pixels, fb, f"frame {k} picture with decals and sprites")
        assert triangles > 300
    finally:
        r.close()
        sc.close()


def test_frames_from_a_loaded_sqwm_map(oracle_lib, tmp_path):
    """The data format on the input side of the path: a map written as .sqwm v11 (quantised lightmaps, light grid) and loaded
    back renders bit-exactly against the oracle fed with the loaded lightmaps — models lit through the loaded light grid."""
    W, H = 640, 360
    kw = dict(seed=7, texture_size_log2=6, num_materials=8, light_grid=1, num_models=3, model_detail=8)
    src = S.Scene(**kw)
    path = str(tmp_path / "level.sqwm")
    src.save_map(path)
    src.close()
    sc = S.Scene(**kw)
    sc.load_map(path)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        poses = sc.camera_path(64)
        for k in (5, 30):
            models = _models_ahead_of_camera(sc, poses[k], k)
            models["lighting"] = S.MODEL_LIGHTING_DEFAULT
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(poses[k], model_entities=models), color_bits=64)
            pixels = np.zeros((H, W), np.uint64)
            r.draw_frame(pixels)
            assert hf.triangle_points is not None and hf.triangle_points["light"].max() > 0  # the light grid lights the models
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(64, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), np.uint64)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64)
            oracle_lib.draw(64, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(r._view.read_surfaces(0, hf.total_surface_pixels), surf, f"frame {k} surfaces from the loaded map")
            assert_pixels_equal(pixels, fb, f"frame {k} picture from the loaded map")
    finally:
        r.close()
        sc.close()


@pytest.mark.skipif(not os.path.exists(S.DEMO2_MAP_PATH), reason="assets/_ref/demo2.sqwm is built by __graft_entry__.build() where /root/reference exists")
@pytest.mark.parametrize("color_bits", [32, 64])
def test_frames_of_the_demo_map(oracle_lib, color_bits):
    """BASELINE.json's own map: demo2.map compiled and lit by this repository's map compiler / lightmapper (row f4).  World
    geometry, submodels (doors, details), sky, alpha-tested / blended / twosided world materials as the demo's material table
    defines them — frames along the path through its player start bit-exact against the oracle, surfaces and picture."""
    W, H = 800, 450
    sc = S.demo2_scene(texture_size_log2=6)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        poses = sc.camera_path(64)
        dtype = np.uint32 if color_bits == 32 else np.uint64
        for k in (0, 5, 20, 47):
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(poses[k], submodel_entities=r.host.game_submodel_entities()), color_bits=color_bits)
            pixels = np.zeros((H, W), dtype)
            r.draw_frame(pixels)
            assert len(hf.cmds) > 50
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            got_surf = r._view.read_surfaces(0, hf.total_surface_pixels)
            # Polygons of `light: false` materials without an emissive layer (the demo's mirror / display materials, which the game
            # covers with portal textures) get NO surface layer: neither the reference (partial_renderer.rs:1700-1819) nor the
            # kernels nor the oracle write their texels, which keep what the pool held before.  The oracle draws from the same.
            layers = N.SURF_APPLY_DIFFUSE | N.SURF_EMISSIVE
            unwritten = 0
            for j in hf.jobs[(hf.jobs["flags"] & layers) == 0]:
                o, n = int(j["surface_pixels_offset"]), int(j["surface_size"][0]) * int(j["surface_size"][1])
                surf[o : o + n] = got_surf[o : o + n]
                unwritten += n
            assert unwritten < 0.5 * hf.total_surface_pixels
            fb = np.zeros((H, W), dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(got_surf, surf, f"demo map frame {k} surfaces")
            # frames drawn from inside a leaf are not cleared (partial_renderer.rs:249-256); the few crack pixels between the
            # polygons of a compiled map keep the previous frame's content on the device and are left out
            covered = fb != 0
            assert fb.size - int(covered.sum()) <= 1e-4 * fb.size
            assert_pixels_equal(np.where(covered, pixels, 0), fb, f"demo map frame {k} picture")
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_submodels_moved_rotated_lit_and_shadowing(oracle_lib, color_bits):
    """Inline models (map submodels): surfaces prepared through the submodel's own camera matrices, basis vectors re-derived in
    world space for specular and dynamic lights, polygons clipped by the leaf planes and drawn through the submodel BSP, sorted
    with models and sprites inside a leaf, decals on their polygons, and depth-tested fills in the shadow maps."""
    from common import scatter_decals, sprites_ahead_of_camera

    W, H = 640, 360
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_submodels=24, num_models=2, model_detail=6)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        path = sc.camera_path(64)
        base = r.host.default_submodel_entities()
        submodel_jobs = 0
        for n, k in enumerate((15, 16, 40, 2)):
            ent = base.copy()
            a = 0.3 * (n + 1)
            ent["rotation"][:, 0] = np.cos(a / 2)
            ent["rotation"][:, 3] = np.sin(a / 2)  # about z
            ent["position"][:, 2] += 4.0 * n  # and lifted a little
            ent["present"][n::5] = 0  # some are hidden
            lights = _lights_along_path(path, 6, 21 + n)
            lights["position"] = path[k][:3] + np.array([[30, 20, 10], [-40, 50, 0], [10, 80, -20], [0, 10, 30], [60, -20, 5], [-20, -30, 15]], np.float32)
            fi = FrameInfo(
                path[k], lights=lights, submodel_entities=ent, model_entities=_models_ahead_of_camera(sc, path[k], k), sprites=sprites_ahead_of_camera(S, path[k], 4, num_skins=2),
                decals=scatter_decals(S, path[k], 24, k, num_skins=2),
            )
            hf = r.prepare_frame(SurfaceInfo(W, H), fi, color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            ent_none = base.copy()
            ent_none["present"] = 0
            sm = np.zeros(max(hf.shadow_maps_len, 1), np.float32)
            if len(hf.depth_maps):
                oracle_lib.draw_shadow_maps(sm, hf.depth_maps, hf.depth_cmds, hf.depth_points)
                got_sm = r._frame.read_shadow_maps(0, hf.shadow_maps_len)
                assert got_sm.view(np.uint32).tolist() == sm[: hf.shadow_maps_len].view(np.uint32).tolist(), f"frame {k} shadow maps"
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, hf.lights, sm, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), pixels.dtype)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(r._view.read_surfaces(0, hf.total_surface_pixels), surf, f"frame {k} surfaces with submodels")
            assert_pixels_equal(pixels, fb, f"frame {k} picture with submodels")
            jobs_with = len(hf.jobs)
            hf0 = r.host.build_frame(W, H, path[k], lights=lights, submodels=ent_none, color_bits=color_bits)
            submodel_jobs += jobs_with - len(hf0.jobs)
        assert submodel_jobs >= 8  # submodel polygons really were built and drawn
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits,portals_depth", [(32, 1), (64, 2)])
def test_config3_portals_and_mirrors(oracle_lib, color_bits, portals_depth):
    """configs[3], last part: a mirror with a glass overlay (mix_surface_with_texture on the GPU), a camera portal and a blended
    parallax portal.  Every nested view is its own device render target, built and drawn before the view that samples it
    (SQW_TEX_VIEW), recursively up to portals_depth and black beyond; the whole tree must equal the oracle's."""
    from common import oracle_render_host_frame, wall_portals

    W, H = 640, 360
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_models=2, model_detail=6, skybox=1, sky_size_log2=6)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True, portals_depth=portals_depth, clear_portal_textures=True))
    try:
        path = sc.camera_path(64)
        nested = 0
        for k in (0, 10, 33):
            portals = wall_portals(S, N, path[k])
            lights = _lights_along_path(path, 2, 31)
            lights["shadow_type"] = N.SHADOW_NONE
            lights["position"] = path[k][:3] + np.array([[20, 30, 10], [-30, 60, 0]], np.float32)
            fi = FrameInfo(path[k], lights=lights, portals=portals, model_entities=_models_ahead_of_camera(sc, path[k], k))
            hf = r.prepare_frame(SurfaceInfo(W, H), fi, color_bits=color_bits)
            pixels = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
            r.draw_frame(pixels)
            nested += sum(1 for n in hf.portals if n.frame is not None)
            oracle_lib.reset_assert_failures()
            _, fb = oracle_render_host_frame(oracle_lib, S, sc, hf, W, H, color_bits)
            assert oracle_lib.assert_failures() == 0
            for node in hf.portals:  # the nested render targets themselves
                _, want = (None, np.zeros(node.resolution[::-1], pixels.dtype)) if node.frame is None else oracle_render_host_frame(oracle_lib, S, sc, node.frame, *node.resolution, color_bits)
                if node.overlay is not None:
                    oracle_lib.mix_surface_with_texture(color_bits, want, node.resolution, node.tc_min, sc.images[node.overlay[0]][1], node.overlay[1], node.overlay[2])
                assert_pixels_equal(node.view.read_pixels(), want, f"frame {k} portal {node.index} texture")
            assert_pixels_equal(pixels, fb, f"frame {k} picture with portals")
        assert nested >= 4
    finally:
        r.close()
        sc.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_config3_translucent_and_alpha_tested_world_materials(oracle_lib, color_bits):
    """configs[3], the world part: a third of the map's materials are alpha-blended, alpha-tested (a grid of holes), averaged or
    additive (material.rs:101-116, demo_materials.json alpha_blend / alpha_grid / stained_glass), so the BSP walk interleaves
    blended polygons with opaque ones: tiles of both classes (opaque-polygon kernel / general kernel) share every frame."""
    W, H = 640, 360
    sc = S.Scene(seed=5, texture_size_log2=6, num_materials=24, material_mode=2)
    r = Renderer(sc, RendererConfig(use_directional_lightmaps=True))
    try:
        path = sc.camera_path(64)
        blended = 0
        # The reference clears the picture only when the camera is outside the map (partial_renderer.rs:253): inside it, a blended
        # wall facing the void blends with what the PREVIOUS frame left in the buffer.  The oracle's buffer is carried from frame
        # to frame the same way (both start from zeros).
        fb = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
        for k in (2, 17, 33, 58):
            hf = r.prepare_frame(SurfaceInfo(W, H), FrameInfo(path[k]), color_bits=color_bits)
            pixels = np.zeros((H, W), fb.dtype)
            r.draw_frame(pixels)
            blended += int((hf.cmds["blend"] != 0).sum())
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(color_bits, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            if hf.needs_clear:
                oracle_lib.fill(fb, S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32)
            oracle_lib.draw(color_bits, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert_pixels_equal(r._view.read_surfaces(0, hf.total_surface_pixels), surf, f"frame {k} surfaces")
            assert_pixels_equal(pixels, fb, f"frame {k} picture")
        assert blended > 40  # blended world polygons really are in view
    finally:
        r.close()
        sc.close()
