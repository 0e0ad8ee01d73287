"""Host mirror of the triangle-model path (square_wheel_b200/host/models.cpp, Renderer::draw_mesh): invariants the reference's
code guarantees (triangle_models_rendering.rs, partial_renderer.rs:3525-3717), checked on the command lists, and the oracle
rasterising those lists without tripping its ported debug_assert!s."""
import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S


@pytest.fixture(scope="module")
def scene():
    sc = S.Scene(seed=11, texture_size_log2=6, num_materials=6, num_models=6, model_detail=8)
    yield sc
    sc.close()


def _entities(sc, pose, k, lighting=S.MODEL_LIGHTING_CONSTANT, distance=70.0):
    n = int(sc.info.num_models)
    m = np.zeros(n, S.model_entity_dtype)
    az = float(pose[3])
    for i in range(n):
        a = az + (i - 0.5 * (n - 1)) * 0.2
        m[i]["position"] = pose[:3] + (distance + 14.0 * i) * np.array([-np.sin(a), np.cos(a), 0.0], np.float32)
        m[i]["rotation"] = (np.cos(0.2 * i), 0.0, 0.0, np.sin(0.2 * i))
        nf = sc.model_info(i).num_frames
        m[i]["frames"] = (k % nf, (k + 3) % nf)
        m[i]["lerp"] = 0.3
        m[i]["model"] = i
        m[i]["skin"] = i
        m[i]["lighting"] = lighting
        m[i]["light"] = (0.5, 0.75, 1.0)
    return m


def test_model_fixtures_cover_the_three_vertex_data_kinds(scene):
    kinds = [scene.model_info(i).kind for i in range(scene.info.num_models)]
    assert kinds == [0, 1, 2, 0, 1, 2]  # NonAnimated, VertexAnimated, SkeletonAnimated (triangle_model.rs:64-75)
    for i in range(scene.info.num_models):
        mi = scene.model_info(i)
        assert mi.num_triangles == 2 * 8 * 16 and mi.num_vertices == 9 * 17 and mi.num_vertices <= 65536  # VertexIndex = u16
        assert mi.num_frames == (1 if mi.kind == 0 else 8)
        assert mi.num_bones == (6 if mi.kind == 2 else 0)
        assert all(lo < hi for lo, hi in zip(mi.bbox_min, mi.bbox_max))


def test_bad_entities_are_rejected(scene):
    hr = S.HostRenderer(scene)
    try:
        m = _entities(scene, scene.camera_path(4)[0], 0)
        m[2]["frames"] = (99, 0)
        with pytest.raises(ValueError):
            hr.build_frame(320, 200, scene.camera_path(4)[0], models=m)
        m = _entities(scene, scene.camera_path(4)[0], 0)
        m[1]["model"] = 77
        with pytest.raises(ValueError):
            hr.build_frame(320, 200, scene.camera_path(4)[0], models=m)
    finally:
        hr.close()


def test_triangle_commands_respect_reference_invariants(scene, oracle_lib):
    W, H = 480, 270
    hr = S.HostRenderer(scene)
    try:
        path = scene.camera_path(32)
        total = 0
        for k in (0, 7, 19):
            without = hr.build_frame(W, H, path[k])
            hf = hr.build_frame(W, H, path[k], models=_entities(scene, path[k], k))
            tri = hf.cmds[hf.cmds["kind"] == N.CMD_TRIANGLE]
            poly = hf.cmds[hf.cmds["kind"] == N.CMD_POLYGON]
            # models only ADD commands: the world polygons and their order are untouched
            assert poly.tobytes() == without.cmds.tobytes()
            total += len(tri)
            if len(tri) == 0:
                continue
            assert hf.triangle_points is not None and len(hf.triangle_points) == 3 * len(tri)
            assert np.array_equal(tri["vertex_offset"], 3 * np.arange(len(tri), dtype=np.uint32))
            assert np.all(tri["tex_kind"] == N.TEX_IMAGE)
            # every triangle is clipped to the screen (leaf bounds are inside it) up to the 16.16 rounding of the clippers
            x = hf.triangle_points["x"].astype(np.float64) / 65536.0
            y = hf.triangle_points["y"].astype(np.float64) / 65536.0
            assert x.min() > -1.0 and x.max() < W + 1.0 and y.min() > -1.0 and y.max() < H + 1.0
            # ConstantLight (0.5, 0.75, 1.0) enters BOTH the static and the dynamic light-grid element of get_model_light
            # (triangle_models_rendering.rs:459-462 and :541-546), so the cube sides are 2 * l and every vertex gets
            # 2 * l * (|nx| + |ny| + |nz|) with 1 <= |n|_1 <= sqrt(3) for a unit normal (get_light_cube_light, light.rs:296-348) and the channel ratios survive
            # (away from vertices produced by the clippers' linear interpolation the ratio is exact up to 16.16 truncation)
            light = hf.triangle_points["light"].astype(np.float64) / 65536.0
            assert light[:, 0].min() >= 1.0 * 0.999 and light[:, 0].max() <= 1.0 * 1.7330
            np.testing.assert_allclose(light[:, 1] / light[:, 0], 1.5, atol=2e-3)
            np.testing.assert_allclose(light[:, 2] / light[:, 0], 2.0, atol=2e-3)
            # texture size = the chosen mip of a 64 x 64 skin
            assert set(np.unique(tri["tex_size"][:, 0])) <= {64, 32, 16, 8, 4}
            # the oracle rasterises the list without tripping a ported debug_assert!
            fb = np.zeros((H, W), np.uint32)
            oracle_lib.reset_assert_failures()
            oracle_lib.draw(32, fb, tri, None, hf.triangle_points, None, scene.images)
            assert oracle_lib.assert_failures() == 0
            assert np.count_nonzero(fb) > 500
        assert total > 1000
    finally:
        hr.close()


def test_models_behind_the_camera_or_in_hidden_leafs_draw_nothing(scene):
    hr = S.HostRenderer(scene)
    try:
        path = scene.camera_path(32)
        behind = _entities(scene, path[3], 3, distance=-90.0)
        hf = hr.build_frame(480, 270, path[3], models=behind)
        assert int((hf.cmds["kind"] == N.CMD_TRIANGLE).sum()) == 0
    finally:
        hr.close()


def test_view_models_are_drawn_last_over_the_whole_viewport(scene, oracle_lib):
    """ModelEntityDrawFlags::VIEW_MODEL (frame_info.rs:106): such a model belongs to no leaf (dynamic_objects_index.rs:61), is always
    prepared (partial_renderer.rs:1048) and is drawn by draw_view_models after the tree (:1324, :2284-2309) with no leaf clip planes:
    its triangle commands come after every other command of the frame, also when the model stands inside a wall."""
    hr = S.HostRenderer(scene)
    try:
        path = scene.camera_path(32)
        pose = path[5]
        ents = _entities(scene, pose, 5)
        plain = hr.build_frame(480, 270, pose, models=ents)
        n_plain = int((plain.cmds["kind"] == N.CMD_TRIANGLE).sum())
        assert n_plain > 0
        # the first model becomes the view model: held 40 units ahead of the eye, and left where it stood
        for position in (pose[:3] + 40.0 * np.array([-np.sin(pose[3]), np.cos(pose[3]), 0.0], np.float32), None):
            view = ents.copy()
            view[0]["flags"] = 1
            if position is not None:
                view[0]["position"] = position
            hf = hr.build_frame(480, 270, pose, models=view)
            kinds = hf.cmds["kind"]
            tri = np.flatnonzero(kinds == N.CMD_TRIANGLE)
            assert len(tri) > 0
            # the frame without the view model is a prefix of this one: everything the view model adds comes last
            others = view[1:].copy()
            base = hr.build_frame(480, 270, pose, models=np.concatenate([np.zeros(0, S.model_entity_dtype), others]))
            hf = hr.build_frame(480, 270, pose, models=view)
            assert len(hf.cmds) > len(base.cmds)
            tail = hf.cmds[len(base.cmds) :]
            assert (tail["kind"] == N.CMD_TRIANGLE).all()
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((270, 480), np.uint32)
            oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, hf.triangle_points, surf, scene.images)
            assert oracle_lib.assert_failures() == 0
            fb0 = np.zeros((270, 480), np.uint32)
            oracle_lib.draw(32, fb0, base.cmds, base.polygon_points, base.triangle_points, surf, scene.images)
            assert int((fb != fb0).sum()) > 200  # the view model is on the picture
    finally:
        hr.close()


def test_default_lighting_uses_only_shadowless_dynamic_lights(scene):
    """ModelLighting::Default without a light grid: black without lights, lit by a shadow-less light, and NOT lit by a light
    whose shadow map lives on the device (documented limitation of the host mirror)."""
    hr = S.HostRenderer(scene)
    try:
        path = scene.camera_path(32)
        m = _entities(scene, path[5], 5, lighting=S.MODEL_LIGHTING_DEFAULT)
        hf = hr.build_frame(480, 270, path[5], models=m)
        assert hf.triangle_points is not None and np.all(hf.triangle_points["light"] == 0)
        lights = np.zeros(1, S.dynamic_light_dtype)
        lights[0]["position"] = path[5][:3]
        lights[0]["radius"] = 512.0
        lights[0]["color"] = (40000.0, 30000.0, 20000.0)
        lights[0]["rotation"] = (1.0, 0.0, 0.0, 0.0)
        hf = hr.build_frame(480, 270, path[5], lights=lights, models=m)
        assert hf.triangle_points["light"].max() > 0
        lights[0]["shadow_type"] = N.SHADOW_CUBE
        hf = hr.build_frame(480, 270, path[5], lights=lights, models=m)
        assert np.all(hf.triangle_points["light"] == 0)
    finally:
        hr.close()


def test_sprites_and_decals_emit_clipped_lit_triangles(scene, oracle_lib):
    """Sprites: a quad (2 triangles) or a tesselated grid, constant light = light_add (+ dynamic lights when light_scale > 0),
    texture coordinates inside the chosen mip.  Decals: triangles only on polygons facing the decal, texture coordinates
    clamped to [1/32, size - 1/32] (partial_renderer.rs:3004-3008), light = lightmap sample or light_add."""
    from common import scatter_decals, sprites_ahead_of_camera

    W, H = 480, 270
    hr = S.HostRenderer(scene)
    try:
        path = scene.camera_path(32)
        n_sprite = n_decal = 0
        for k in (0, 7, 19):
            base = hr.build_frame(W, H, path[k])
            hs = hr.build_frame(W, H, path[k], sprites=sprites_ahead_of_camera(S, path[k], 8))
            tri = hs.cmds[hs.cmds["kind"] == N.CMD_TRIANGLE]
            assert hs.cmds[hs.cmds["kind"] == N.CMD_POLYGON].tobytes() == base.cmds.tobytes()
            if len(tri):
                n_sprite += len(tri)
                light = hs.triangle_points["light"].astype(np.float64) / 65536.0
                np.testing.assert_allclose(light, np.broadcast_to([0.8, 0.7, 0.9], light.shape), atol=2e-5)  # light_scale = 0: light_add only
                tc = hs.triangle_points["tc"].astype(np.float64) / 65536.0
                size = np.repeat(tri["tex_size"], 3, axis=0)
                assert np.all(tc >= 1.0 / 16.0 - 1e-3) and np.all(tc <= size - 1.0 / 16.0 + 1e-3)  # tc_reduce (:3311)
            hd = hr.build_frame(W, H, path[k], decals=scatter_decals(S, path[k], 48, k))
            tri = hd.cmds[hd.cmds["kind"] == N.CMD_TRIANGLE]
            assert hd.cmds[hd.cmds["kind"] == N.CMD_POLYGON].tobytes() == base.cmds.tobytes()
            if len(tri):
                n_decal += len(tri)
                tc = hd.triangle_points["tc"].astype(np.float64) / 65536.0
                size = np.repeat(tri["tex_size"], 3, axis=0)
                assert np.all(tc >= 1.0 / 32.0 - 1e-4) and np.all(tc <= size - 1.0 / 32.0 + 1e-4)
                x = hd.triangle_points["x"].astype(np.float64) / 65536.0
                y = hd.triangle_points["y"].astype(np.float64) / 65536.0
                assert x.min() > -1.0 and x.max() < W + 1.0 and y.min() > -1.0 and y.max() < H + 1.0
                assert np.all(hd.triangle_points["light"] >= 0)
                fb = np.zeros((H, W), np.uint32)
                oracle_lib.reset_assert_failures()
                oracle_lib.draw(32, fb, tri, None, hd.triangle_points, None, scene.images)
                assert oracle_lib.assert_failures() == 0
        assert n_sprite >= 16 and n_decal >= 40
        with pytest.raises(ValueError):
            bad = sprites_ahead_of_camera(S, path[0], 2)
            bad[1]["radius"] = 0.0
            hr.build_frame(W, H, path[0], sprites=bad)
    finally:
        hr.close()
