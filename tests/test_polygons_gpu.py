"""World-polygon preparation on the GPU (the geometry half of the host front end): sqw_polygons_prepare against
the host mirror of PartialRenderer::prepare_polygons_surfaces / prepare_polygon_surface (partial_renderer.rs:1327-1360, :1464-1638).

The visibility flood stays on the host; it hands over the (polygon, leaf clip planes) candidates.  Kernel and mirror evaluate the same
float expressions in the same order without fused multiply-add, so every record — visibility, mip, depth and texture-coordinate
equations, surface rectangle, surface_pixels_offset — the frame totals and the sky window are compared BIT FOR BIT."""
import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S
from square_wheel_b200.device import Context, SqwError

pytestmark = pytest.mark.gpu


def _compare(view, hr, what):
    cand, planes, want, fr = hr.prepared_polygons()
    assert len(cand) > 20, what
    view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
    got, totals = view.polygons_read(len(cand))
    assert got["flags"].tolist() == want["flags"].tolist(), f"{what}: visibility / sky flags"
    for field in ("mip", "surface_tc_min", "surface_size", "surface_pixels_offset"):
        assert np.array_equal(got[field], want[field]), f"{what}: {field}"
    for field in ("depth_equation", "tex_coord_equation"):
        a, b = np.ascontiguousarray(got[field]).view(np.uint32), np.ascontiguousarray(want[field]).view(np.uint32)
        assert np.array_equal(a, b), f"{what}: {field} differs in {int((a != b).any(axis=1).sum())} records"
    assert totals.num_visible == int(fr["num_visible"]) == int((want["flags"] == N.PREPARED_VISIBLE).sum())
    assert totals.num_surfaces_pixels == int(fr["num_surfaces_pixels"])
    assert totals.has_sky == int(fr["has_sky"])
    if totals.has_sky:
        assert totals.sky_material == int(fr["sky_material"])
        assert np.array_equal(np.array(list(totals.sky_bounds), np.float32).view(np.uint32), np.ascontiguousarray(fr["sky_bounds"]).view(np.uint32)), f"{what}: sky window"
    return int(totals.num_visible), int((want["flags"] == N.PREPARED_SKY).sum()), want


def _compare_draw(view, hr, hf, what):
    """sqw_polygons_draw against the commands the mirror's draw_polygon recorded for the same polygons (after _compare)."""
    cand, _, _, _ = hr.prepared_polygons()
    index_of = {int(p): i for i, p in enumerate(cand["polygon"])}
    order = np.array([index_of[int(p)] for p in hr.polygon_draw_order()], np.uint32)
    view.polygons_draw(order)
    got_cmds, got_pts = view.polygons_draw_read()
    sel = (hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["tex_kind"] == N.TEX_SURFACE)
    want_cmds = hf.cmds[sel]
    assert len(got_cmds) == len(want_cmds) == len(order) > 10, what
    for field in ("kind", "blend", "tc_mode", "tex_kind", "num_vertices", "tex_ref", "tex_size"):
        assert np.array_equal(got_cmds[field], want_cmds[field]), f"{what}: {field}"
    for field in ("depth_eq", "tc_eq"):
        a, b = np.ascontiguousarray(got_cmds[field]).view(np.uint32), np.ascontiguousarray(want_cmds[field]).view(np.uint32)
        assert np.array_equal(a, b), f"{what}: {field} differs in {int((a != b).any(axis=1).sum())} commands"
    # vertices: the mirror's list also holds the vertices of other commands (sky box), so compare per command
    assert got_cmds["vertex_offset"].tolist() == np.concatenate(([0], np.cumsum(got_cmds["num_vertices"])[:-1])).tolist()
    for g, w in zip(got_cmds, want_cmds):
        a = got_pts[int(g["vertex_offset"]) : int(g["vertex_offset"]) + int(g["num_vertices"])]
        b = hf.polygon_points[int(w["vertex_offset"]) : int(w["vertex_offset"]) + int(w["num_vertices"])]
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), f"{what}: vertices"
    return set(want_cmds["tc_mode"].tolist()), set(want_cmds["blend"].tolist())


@pytest.mark.parametrize("mip_bias", [0.0, -1.0, 0.6])
def test_polygons_prepare_matches_the_host_mirror_bit_for_bit(mip_bias):
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, skybox=1, sky_size_log2=6, material_mode=2)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    hr = S.HostRenderer(sc, textures_mip_bias=mip_bias)
    modes, blends = set(), set()
    try:
        ctx.upload_map_geometry(*sc.map_geometry())
        view = ctx.create_view(64, 64, 64)
        path = sc.camera_path(64)
        visible = sky = 0
        mips = set()
        for k, (w, h) in zip((0, 5, 17, 30, 44, 59), ((640, 360), (1280, 720), (3840, 2160), (333, 777), (1920, 1080), (3840, 2160))):
            hf = hr.build_frame(w, h, path[k])
            v, s, want = _compare(view, hr, f"pose {k} at {w}x{h}, mip bias {mip_bias}")
            m, b = _compare_draw(view, hr, hf, f"draw, pose {k} at {w}x{h}, mip bias {mip_bias}")
            modes |= m
            blends |= b
            visible += v
            sky += s
            mips |= set(want["mip"][want["flags"] == N.PREPARED_VISIBLE].tolist())
        assert visible > 300 and sky > 0 and len(mips) >= 2  # culled, clipped, sky and several mips all occur
        assert len(modes) >= 2 and len(blends) >= 2  # several interpolation modes and blended world materials
    finally:
        hr.close()
        ctx.close()
        sc.close()


def test_polygons_prepare_on_the_bench_scene_and_bad_arguments():
    """The 4K bench map (seed 2, 13k polygons): every pose of a strided camera path; then the argument checks of the C-ABI."""
    sc = S.Scene(seed=2, grid=(24, 24, 5), cell_size=(160.0, 160.0, 128.0), num_rooms=26, num_materials=24, texture_size_log2=5, max_polygon_size=64.0, material_mode=0, emissive_fraction=0.15)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    hr = S.HostRenderer(sc)
    try:
        view = ctx.create_view(64, 64, 64)
        path = sc.camera_path(64)
        hr.build_frame(3840, 2160, path[0])
        cand, planes, want, fr = hr.prepared_polygons()
        with pytest.raises(SqwError):  # no geometry uploaded yet
            view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], 0.0, cand, planes)
        ctx.upload_map_geometry(*sc.map_geometry())
        modes = set()
        for k in range(0, 64, 7):
            hf = hr.build_frame(3840, 2160, path[k])
            _compare(view, hr, f"bench scene pose {k}")
            modes |= _compare_draw(view, hr, hf, f"draw, bench scene pose {k}")[0]
        assert modes == {N.TC_FULL_PERSPECTIVE, N.TC_LINE_Z_CORRECTION, N.TC_AFFINE} or len(modes) >= 2
        bad = cand.copy()
        bad["polygon"][3] = int(sc.info.num_polygons)
        with pytest.raises(SqwError):
            view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], 0.0, bad, planes)
        bad = cand.copy()
        bad["clip_set"][0] = len(planes)
        with pytest.raises(SqwError):
            view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], 0.0, bad, planes)
    finally:
        hr.close()
        ctx.close()
        sc.close()


@pytest.mark.parametrize("material_mode, size", [(0, (1280, 720)), (2, (1918, 1078)), (0, (3840, 2160))])
def test_frame_drawn_from_device_command_lists(material_mode, size):
    """The chain of SURVEY 8 f3 for world polygons: candidates and draw order from the host's visibility flood and BSP walk, geometry
    (sqw_polygons_prepare / _draw) and the rasteriser's frame set-up (sqw_frame_set_draw_from_polygons: CmdDev, span offsets, chunk and
    band lists) on the device — the command list never visits the host.  The picture must equal, bit for bit, the one of the same
    frame recorded from the host mirror's lists (which the frame tests pin against the oracle)."""
    from square_wheel_b200.device import make_post_params

    W, H = size
    sc = S.Scene(seed=11, texture_size_log2=6, num_materials=8, material_mode=material_mode)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    hr = S.HostRenderer(sc)
    try:
        sc.upload(ctx)
        ctx.upload_map_geometry(*sc.map_geometry())
        ctx.upload_map_materials(sc.map_materials(), hr.polygon_basis())
        view = ctx.create_view(W, H, 64)
        post = make_post_params(exposure=0.9, bloom_sigma=0.0)
        path = sc.camera_path(64)
        blended = 0
        for k in (2, 23, 41):
            hf = hr.build_frame(W, H, path[k])
            clear = S.BACKGROUND_COLOR64 if hf.needs_clear else None
            frame = view.create_frame()
            hf.record(frame, clear_color=clear, post=post)
            # blended polygons read what lies below them: both pictures start from the same buffer
            view.write_pixels(np.zeros((H, W), np.uint64))
            frame.submit(N.STAGE_ALL)
            want_hdr, want_ldr, want_avg = view.read_pixels(), view.read_pixels32(), view.average_color()
            frame.close()
            # the same frame with the draw half built on the device
            cand, planes, want, fr = hr.prepared_polygons()
            view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
            frame = view.create_frame()
            with pytest.raises(SqwError):  # the draw lists of an earlier prepare are not this frame's
                frame.set_draw_from_polygons(clear_color=clear)
            frame.close()
            index_of = {int(p): i for i, p in enumerate(cand["polygon"])}
            view.polygons_draw(np.array([index_of[int(p)] for p in hr.polygon_draw_order()], np.uint32))
            frame = view.create_frame()
            frame.set_surfaces(hf.jobs, None, None, hf.cam_pos, hf.total_surface_pixels)
            frame.set_draw_from_polygons(clear_color=clear)
            frame.set_post(post)
            view.write_pixels(np.zeros((H, W), np.uint64))
            frame.submit(N.STAGE_ALL)
            got_hdr, got_ldr, got_avg = view.read_pixels(), view.read_pixels32(), view.average_color()
            frame.close()
            assert len(hf.cmds) == int(((hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["tex_kind"] == N.TEX_SURFACE)).sum()) > 10  # a world-only frame
            assert np.array_equal(got_hdr, want_hdr), f"pose {k}: 64-bit picture differs in {int((got_hdr != want_hdr).sum())} pixels"
            assert np.array_equal(got_ldr, want_ldr), f"pose {k}: tonemapped picture"
            assert np.array_equal(got_avg, want_avg)
            # and with the surface half from the device too: jobs and band lists built from the prepared records — the host's part
            # of the frame is the visibility flood, the BSP walk and two small index lists
            frame = view.create_frame()
            frame.set_surfaces_from_polygons(hf.cam_pos, use_directional_lightmaps=True)
            frame.set_draw_from_polygons(clear_color=clear)
            frame.set_post(post)
            view.write_pixels(np.zeros((H, W), np.uint64))
            frame.submit(N.STAGE_ALL)
            got_hdr, got_ldr = view.read_pixels(), view.read_pixels32()
            st = frame.stats()
            frame.close()
            assert np.array_equal(got_hdr, want_hdr), f"pose {k}, device jobs: 64-bit picture differs in {int((got_hdr != want_hdr).sum())} pixels"
            assert np.array_equal(got_ldr, want_ldr), f"pose {k}, device jobs: tonemapped picture"
            assert st.surface_texels == hf.total_surface_pixels
            blended += int((hf.cmds["blend"] != 0).sum())
        assert (blended > 0) == (material_mode == 2)
    finally:
        hr.close()
        ctx.close()
        sc.close()


def test_lit_frame_with_device_side_light_selection():
    """Dynamic lights on the device chain: the per-polygon light lists (polygon_is_affected_by_light, partial_renderer.rs:3877-3985: sphere
    against the polygon's plane and texture-coordinate rectangle; projector pyramid against its corners; at most six per polygon, in the
    order of the leaf's list) are selected by the job kernel from the lights of every visible leaf.  Shadow maps are rendered on the GPU
    from the host's fill list as in every lit frame.  Picture and light counts must equal the host-recorded frame's."""
    from square_wheel_b200.device import make_post_params
    from test_frames_gpu import _lights_along_path

    W, H = 1280, 720
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    hr = S.HostRenderer(sc)
    try:
        sc.upload(ctx)
        ctx.upload_map_geometry(*sc.map_geometry())
        ctx.upload_map_materials(sc.map_materials(), hr.polygon_basis())
        view = ctx.create_view(W, H, 64)
        post = make_post_params(exposure=1.0, bloom_sigma=0.0)
        path = sc.camera_path(64)
        lit = 0
        for k in (4, 30, 52):
            lights = _lights_along_path(path, 9, 40 + k)
            lights["position"] = path[k][:3] + np.random.default_rng(k).uniform(-120.0, 120.0, (9, 3)) * (1.0, 1.0, 0.3)
            hf = hr.build_frame(W, H, path[k], lights=lights)
            assert len(hf.lights) == 9 and (hf.jobs["num_lights"] > 0).sum() > 5
            lit += int((hf.jobs["num_lights"] > 0).sum())
            clear = S.BACKGROUND_COLOR64 if hf.needs_clear else None
            frame = view.create_frame()
            hf.record(frame, clear_color=clear, post=post)
            view.write_pixels(np.zeros((H, W), np.uint64))
            frame.submit(N.STAGE_ALL)
            want_hdr, want_ldr = view.read_pixels(), view.read_pixels32()
            frame.close()
            cand, planes, want, fr = hr.prepared_polygons()
            view.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
            index_of = {int(p): i for i, p in enumerate(cand["polygon"])}
            view.polygons_draw(np.array([index_of[int(p)] for p in hr.polygon_draw_order()], np.uint32))
            offsets, indices = hr.clip_set_lights()
            assert len(offsets) == len(planes) + 1
            frame = view.create_frame()
            with pytest.raises(SqwError):  # one light list per clip set
                frame.set_surfaces_from_polygons(hf.cam_pos, True, lights=hf.lights, clip_set_light_offsets=offsets[:-1], clip_set_lights=indices, shadow_maps_len=hf.shadow_maps_len)
            frame.set_surfaces_from_polygons(hf.cam_pos, True, lights=hf.lights, clip_set_light_offsets=offsets, clip_set_lights=indices, shadow_maps_len=hf.shadow_maps_len)
            frame.set_shadow_map_draw(hf.depth_maps, hf.depth_cmds, hf.depth_points)
            frame.set_draw_from_polygons(clear_color=clear)
            frame.set_post(post)
            view.write_pixels(np.zeros((H, W), np.uint64))
            frame.submit(N.STAGE_ALL)
            got_hdr, got_ldr = view.read_pixels(), view.read_pixels32()
            frame.close()
            assert np.array_equal(got_hdr, want_hdr), f"pose {k}: 64-bit picture differs in {int((got_hdr != want_hdr).sum())} pixels"
            assert np.array_equal(got_ldr, want_ldr), f"pose {k}: tonemapped picture"
        assert lit > 40
    finally:
        hr.close()
        ctx.close()
        sc.close()
