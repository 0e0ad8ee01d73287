"""CPU checks of the host-side mirror (C++ libsqw_host): map structure, visibility, job/command invariants that the
reference guards with debug_assert!s, and the oracle running a whole frame from its output."""
import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S


@pytest.fixture(scope="module")
def scene():
    sc = S.Scene(seed=5, texture_size_log2=6, num_materials=8)
    yield sc
    sc.close()


def test_scene_is_deterministic():
    a = S.Scene(seed=9, texture_size_log2=5, num_materials=4)
    b = S.Scene(seed=9, texture_size_log2=5, num_materials=4)
    c = S.Scene(seed=10, texture_size_log2=5, num_materials=4)
    assert np.array_equal(a.directional_lightmaps, b.directional_lightmaps)
    assert all(np.array_equal(x[1], y[1]) for x, y in zip(a.images, b.images))
    assert a.info.num_polygons != c.info.num_polygons or not np.array_equal(a.lightmaps, c.lightmaps)
    for s in (a, b, c):
        s.close()


def test_jobs_and_commands_respect_reference_invariants(scene):
    r = S.HostRenderer(scene)
    path = scene.camera_path(16)
    total_cmds = 0
    for pose in path:
        hf = r.build_frame(640, 360, pose)
        j = hf.jobs
        assert len(j) > 0 and len(hf.cmds) > 0
        # surface rectangles are a prefix sum (partial_renderer.rs:1617-1618): disjoint and dense
        sizes = j["surface_size"].astype(np.int64)
        offs = j["surface_pixels_offset"].astype(np.int64)
        assert np.array_equal(offs, np.concatenate([[0], np.cumsum(sizes[:, 0] * sizes[:, 1])[:-1]]))
        assert hf.total_surface_pixels == int((sizes[:, 0] * sizes[:, 1]).sum())
        # lightmap fetches stay inside the polygon's lightmap (debug_assert surfaces.rs:482, :541)
        lit = j["lightmap_offset"] != 0
        sc = j["lightmap_scale_log2"].astype(np.int64)
        for axis in (0, 1):
            last = ((sizes[:, axis] - 1 + j["lightmap_tc_shift"][:, axis]) >> sc) + 1
            assert np.all(last[lit] < j["lightmap_size"][:, axis][lit])
        assert np.all(j["lightmap_size"][lit] <= 18)
        # every command references the surface of exactly one job, with the job's size
        assert np.all(hf.cmds["tex_kind"] == N.TEX_SURFACE)
        by_offset = {int(o): tuple(s) for o, s in zip(offs, sizes)}
        for c in hf.cmds:
            assert by_offset[int(c["tex_ref"])] == tuple(int(x) for x in c["tex_size"])
            assert 3 <= c["num_vertices"] <= 24
        # projected vertices are inside the viewport (clipping by the leaf octagon) up to rounding
        x = hf.polygon_points["x"] / 65536.0
        y = hf.polygon_points["y"] / 65536.0
        assert x.min() >= -1 and x.max() <= 641 and y.min() >= -1 and y.max() <= 361
        total_cmds += len(hf.cmds)
    assert total_cmds > 16 * 20
    r.close()


def test_oracle_renders_host_frame_without_gaps(scene, oracle_lib):
    """A closed map seen from inside: every pixel is covered (no background shows through), and no debug assert fires."""
    r = S.HostRenderer(scene)
    W, H = 320, 180
    for pose in scene.camera_path(6):
        hf = r.build_frame(W, H, pose)
        oracle_lib.reset_assert_failures()
        surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf.total_surface_pixels)
        fb = np.full((H, W), 0xDEADBEEF, np.uint32)
        oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, scene.images)
        assert oracle_lib.assert_failures() == 0
        assert np.count_nonzero(fb == 0xDEADBEEF) == 0
    r.close()


def test_skybox_polygons_become_a_window(oracle_lib):
    """Sky-material polygons are not drawn (no surface job, no draw command); instead the sky box sides are the FIRST commands
    of the frame, textured from images of the target's colour type, and the frame still has no gaps."""
    from square_wheel_b200 import _native as N

    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, skybox=1, sky_size_log2=6)
    plain = S.Scene(seed=7, texture_size_log2=6, num_materials=8)
    hr, hr_plain = S.HostRenderer(sc), S.HostRenderer(plain)
    try:
        W, H = 320, 180
        path = sc.camera_path(64)
        seen = 0
        for k in (25, 26, 3):
            pose = path[k].copy()
            pose[4] += 0.5
            hf = hr.build_frame(W, H, pose, color_bits=32)
            hp = hr_plain.build_frame(W, H, pose, color_bits=32)
            sky = hf.cmds["tex_kind"] == N.TEX_IMAGE
            n_sky = int(sky.sum())
            if n_sky == 0:
                continue
            seen += 1
            assert np.all(sky[:n_sky]) and not np.any(sky[n_sky:])  # drawn before the BSP walk (partial_renderer.rs:1296)
            assert len(hf.jobs) < len(hp.jobs) and len(hf.cmds) - n_sky < len(hp.cmds)  # the sky polygons themselves are gone
            for c in hf.cmds[:n_sky]:
                kind, px = sc.images[int(c["tex_ref"])]
                assert kind == N.IMAGE_COLOR32 and px.shape == (int(c["tex_size"][1]), int(c["tex_size"][0]))
            h64 = hr.build_frame(W, H, pose, color_bits=64)
            assert all(sc.images[int(c["tex_ref"])][0] == N.IMAGE_COLOR64 for c in h64.cmds[:n_sky])
            surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((H, W), np.uint32)
            oracle_lib.reset_assert_failures()
            oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert int((fb == 0).sum()) == 0
        assert seen >= 2
    finally:
        hr.close()
        hr_plain.close()
        sc.close()
        plain.close()


def test_submodels_follow_their_entities(oracle_lib):
    """Inline models: hidden entities add nothing; present ones add surface jobs and polygon commands; moving an entity moves
    its picture; depth-tested fills appear in the shadow-map lists; the oracle draws the frame without gaps or asserts."""
    from square_wheel_b200 import _native as N

    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_submodels=24)
    plain = S.Scene(seed=7, texture_size_log2=6, num_materials=8)
    hr, hp = S.HostRenderer(sc), S.HostRenderer(plain)
    try:
        assert sc.info.num_submodels == 24 and sc.info.num_polygons > plain.info.num_polygons
        W, H = 320, 180
        pose = sc.camera_path(64)[15]
        ent = hr.default_submodel_entities()
        hidden = ent.copy()
        hidden["present"] = 0
        f_hidden = hr.build_frame(W, H, pose, submodels=hidden)
        f_plain = hp.build_frame(W, H, pose)
        assert len(f_hidden.jobs) == len(f_plain.jobs) and len(f_hidden.cmds) == len(f_plain.cmds)
        f_shown = hr.build_frame(W, H, pose, submodels=ent)
        assert len(f_shown.jobs) > len(f_hidden.jobs) and len(f_shown.cmds) > len(f_hidden.cmds)
        moved = ent.copy()
        moved["position"][:, 0] += 9.0
        f_moved = hr.build_frame(W, H, pose, submodels=moved)
        assert f_moved.polygon_points.tobytes() != f_shown.polygon_points.tobytes()
        with pytest.raises(ValueError):
            hr.build_frame(W, H, pose, submodels=ent[:3])
        # shadow maps: submodel polygons are the depth-tested fills (depth_renderer.rs:64-66)
        lights = np.zeros(1, S.dynamic_light_dtype)
        lights[0]["position"] = pose[:3]
        lights[0]["radius"] = 512.0
        lights[0]["color"] = (30000.0, 30000.0, 30000.0)
        lights[0]["shadow_type"] = N.SHADOW_CUBE
        lights[0]["rotation"] = (1.0, 0.0, 0.0, 0.0)
        f_lit = hr.build_frame(W, H, pose, lights=lights, submodels=ent)
        assert len(f_lit.depth_cmds) and int(f_lit.depth_cmds["depth_test"].sum()) > 0
        f_lit_hidden = hr.build_frame(W, H, pose, lights=lights, submodels=hidden)
        assert int(f_lit_hidden.depth_cmds["depth_test"].sum()) == 0
        surf = oracle_lib.build_surfaces(32, f_shown.jobs, None, None, f_shown.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, f_shown.total_surface_pixels)
        fb = np.zeros((H, W), np.uint32)
        oracle_lib.reset_assert_failures()
        oracle_lib.draw(32, fb, f_shown.cmds, f_shown.polygon_points, None, surf, sc.images)
        assert oracle_lib.assert_failures() == 0 and int((fb == 0).sum()) == 0
    finally:
        hr.close()
        hp.close()
        sc.close()
        plain.close()


def test_portals_and_mirrors_nest_host_frames(oracle_lib):
    """FrameInfo::portals: every visible portal gets a texture rectangle, a nested host frame prepared with the portal's own
    camera matrices (mirror: reflected eye, visibility started at the mirror's leafs), and a TEX_VIEW polygon command in the
    outer frame; past portals_depth the nested frame is missing (black texture).  The oracle renders the whole tree."""
    from common import oracle_render_host_frame, wall_portals
    from square_wheel_b200 import _native as N

    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_models=2, model_detail=6)
    try:
        W, H = 320, 180
        pose = sc.camera_path(64)[0]
        portals = wall_portals(S, N, pose)
        for depth in (0, 1, 2):
            hr = S.HostRenderer(sc, portals_depth=depth)
            try:
                plain = hr.build_frame(W, H, pose, color_bits=32)
                hf = hr.build_frame(W, H, pose, portals=portals, color_bits=32)
                assert len(plain.portals) == 0 and int((plain.cmds["tex_kind"] == N.TEX_VIEW).sum()) == 0
                assert len(hf.portals) >= 2
                view_cmds = hf.cmds[hf.cmds["tex_kind"] == N.TEX_VIEW]
                assert sorted(set(view_cmds["tex_ref"].tolist())) == sorted(n.index for n in hf.portals)
                for node in hf.portals:
                    assert 0 < node.resolution[0] <= 95 and 0 < node.resolution[1] <= 80  # 60 x 50 units at 1.5 texels per unit, mip >= 0
                    cmd = view_cmds[view_cmds["tex_ref"] == node.index][0]
                    assert tuple(cmd["tex_size"]) == node.resolution
                    assert (node.frame is None) == (depth == 0)
                    if node.frame is not None:
                        assert len(node.frame.cmds) > 0
                        # nested views see the portals too; at depth 1 their own nested frames are past the limit
                        assert all((inner.frame is None) == (depth == 1) for inner in node.frame.portals)
                    assert (node.overlay is not None) == (node.index == 0)
                oracle_lib.reset_assert_failures()
                _, fb = oracle_render_host_frame(oracle_lib, S, sc, hf, W, H, 32)
                assert oracle_lib.assert_failures() == 0
                if depth == 0:
                    black = fb.copy()
                else:
                    assert int((fb != black).sum()) > 500  # the nested views really show something
            finally:
                hr.close()
    finally:
        sc.close()
