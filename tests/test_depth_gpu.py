"""Shadow-map depth rasteriser parity: depth_setup_kernel + depth_resolve_kernel vs the CPU restatement
of DepthRasterizer::fill_polygon (rasterizer.rs:1295-1474), through the C-ABI (sqw_frame_set_shadow_map_draw), bit-exact —
including the reference's non-associative running float sums — and the surface stage sampling maps that never left the device."""
import numpy as np
import pytest

import synth
from common import Assets, assert_pixels_equal
from square_wheel_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def assets():
    return Assets(seed=21)


def make_depth_scene(r, sizes, polys_per_map=40, tested_per_map=8):
    """Several maps, each with world fills (no depth test) followed by depth-tested fills, like DepthRenderer::draw_map."""
    maps = np.zeros(len(sizes), N.depth_map_dtype)
    cmds_all, pts_all = [], []
    offset = 0
    for m, (w, h, flags) in enumerate(sizes):
        n = polys_per_map + tested_per_map
        cmds, pts = synth.make_polygon_scene(r, w, h, n, [(16, 16)] * n, modes=(0,))
        dc = np.zeros(n, N.depth_cmd_dtype)
        dc["num_vertices"] = cmds["num_vertices"]
        dc["vertex_offset"] = cmds["vertex_offset"] + sum(len(p) for p in pts_all)
        dc["depth_eq"] = cmds["depth_eq"]
        dc["depth_test"][polys_per_map:] = 1
        maps[m]["offset"] = offset
        maps[m]["width"], maps[m]["height"] = w, h
        maps[m]["first_cmd"] = sum(len(c) for c in cmds_all)
        maps[m]["num_cmds"] = n
        maps[m]["flags"] = flags
        maps[m]["clear_value"] = 0.0
        offset += w * h
        cmds_all.append(dc)
        pts_all.append(pts)
    return maps, np.concatenate(cmds_all), np.concatenate(pts_all), offset


def test_shadow_maps_bit_exact(assets, oracle_lib):
    ctx = assets.device_context()
    r = synth.rng(5)
    sizes = [(64, 64, N.DEPTH_MAP_CLEAR), (256, 256, N.DEPTH_MAP_CLEAR | N.DEPTH_MAP_PROJECTOR_CIRCLE), (128, 128, 0), (33, 70, N.DEPTH_MAP_CLEAR), (1024, 1024, N.DEPTH_MAP_CLEAR)]
    maps, cmds, pts, total = make_depth_scene(r, sizes)
    stale = r.uniform(0.0, 0.01, total).astype(np.float32)  # maps without the CLEAR flag keep what the array held
    view = ctx.create_view(16, 16, 32)
    frame = view.create_frame()
    frame.set_surfaces(np.zeros(0, N.surface_job_dtype), None, stale, (0, 0, 0), 0)
    frame.set_shadow_map_draw(maps, cmds, pts)
    frame.submit(N.STAGE_SHADOW_MAPS)
    got = frame.read_shadow_maps(0, total)
    want = oracle_lib.draw_shadow_maps(stale.copy(), maps, cmds, pts)
    assert got.view(np.uint32).tolist() == want.view(np.uint32).tolist()
    assert np.count_nonzero(want != stale) > total // 4  # the fills really covered the maps
    frame.close()
    view.close()


def test_surfaces_sample_device_resident_shadow_maps(assets, oracle_lib):
    """Lights whose shadow maps are produced by the shadow map stage on the device (never uploaded) light the surfaces exactly
    like the oracle fed with the oracle's own maps."""
    ctx = assets.device_context()
    r = synth.rng(6)
    lights, sm_random = synth.make_lights(r, 6)
    sizes, light_of_map = [], []
    for i, l in enumerate(lights):
        s = int(l["shadow_map_size"])
        if l["shadow_type"] == N.SHADOW_CUBE:
            for side in range(6):
                sizes.append((s, s, N.DEPTH_MAP_CLEAR))
                light_of_map.append((i, side))
        elif l["shadow_type"] == N.SHADOW_PROJECTOR:
            sizes.append((s, s, N.DEPTH_MAP_CLEAR | N.DEPTH_MAP_PROJECTOR_CIRCLE))
            light_of_map.append((i, 0))
    maps, cmds, pts, total = make_depth_scene(r, sizes, polys_per_map=12, tested_per_map=3)
    for m, (i, side) in enumerate(light_of_map):
        s = int(lights[i]["shadow_map_size"])
        maps[m]["offset"] = int(lights[i]["shadow_map_offset"]) + side * s * s
    assert total == len(sm_random)
    # depth values of make_polygon_scene are z_scale-normalised (~2^-5); bring the lights' compare depths into the same range
    jobs, total_px, _ = synth.make_surface_jobs(r, 24, assets.texture_ids, assets.texture_flags, directional=True, lights=len(lights), max_size=48)
    jobs["flags"] |= N.SURF_APPLY_DIFFUSE
    cam = np.array([5.0, -3.0, 40.0], np.float32)
    view = ctx.create_view(16, 16, 64)
    frame = view.create_frame()
    frame.set_surfaces(jobs, lights, None, cam, total_px, device_shadow_maps_len=total)
    frame.set_shadow_map_draw(maps, cmds, pts)
    frame.submit(N.STAGE_SHADOW_MAPS | N.STAGE_SURFACES)
    got = view.read_surfaces(0, total_px)
    sm = oracle_lib.draw_shadow_maps(np.zeros(total, np.float32), maps, cmds, pts)
    want = oracle_lib.build_surfaces(64, jobs, lights, sm, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, total_px)
    assert_pixels_equal(got, want, "surfaces lit through device-rendered shadow maps")
    assert frame.read_shadow_maps(0, total).view(np.uint32).tolist() == sm.view(np.uint32).tolist()
    frame.close()
    view.close()
