"""Stage 2 parity: raster_setup/bin/tile kernels vs the CPU restatement of rasterizer.rs, through sqw_draw."""
import numpy as np
import pytest

import synth
from common import Assets, assert_pixels_equal
from square_wheel_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def assets():
    return Assets(seed=21)


def _surface_pool(r, color_bits, tex_sizes):
    """A fake surface pool: random pixels for every command's texture, packed back to back."""
    offs, total = [], 0
    for (w, h) in tex_sizes:
        offs.append(total)
        total += w * h
    if color_bits == 32:
        pool = r.integers(0, 2**32, total, dtype=np.uint64).astype(np.uint32)
    else:
        pool = r.integers(0, 2**63, total, dtype=np.uint64) & np.uint64(0x00FF0FFF0FFF0FFF)
    return pool, offs




@pytest.mark.parametrize("color_bits", [32, 64])
@pytest.mark.parametrize("mode", [N.TC_FULL_PERSPECTIVE, N.TC_LINE_Z_CORRECTION, N.TC_AFFINE])
def test_polygons_bit_exact(assets, oracle_lib, color_bits, mode):
    """Random convex polygons (3..8 vertices, partly off screen, thin ones included), all five blend modes, one tc mode per
    case, odd resolution (tile tails)."""
    r = synth.rng(300 + color_bits + mode)
    ctx = assets.device_context()
    W, H = 333, 201
    n = 120
    # polygon textures must have the framebuffer colour type: use COLOR32 / COLOR64 images
    ids = assets.lite_ids if color_bits == 32 else assets.sky64_ids
    tex = [ids[int(r.integers(0, len(ids)))] for _ in range(n)]
    cmds, pts = synth.make_polygon_scene(r, W, H, n, [assets.image_size(t) for t in tex], modes=(mode,), blends=(0, 1, 2, 3, 4))
    cmds["tex_kind"] = N.TEX_IMAGE
    cmds["tex_ref"] = tex
    clear = 0x0012345600654321 if color_bits == 64 else 0x80402010
    view = ctx.create_view(W, H, color_bits)
    try:
        view.draw(cmds, pts, None, clear_color=clear)
        got = view.read_pixels()
    finally:
        view.close()
    fb = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
    oracle_lib.fill(fb, clear)
    oracle_lib.reset_assert_failures()
    oracle_lib.draw(color_bits, fb, cmds, pts, None, None, assets.images)
    assert oracle_lib.assert_failures() == 0
    assert_pixels_equal(got, fb, f"polygons c{color_bits} mode{mode}")


@pytest.mark.parametrize("color_bits", [32, 64])
def test_triangles_bit_exact(assets, oracle_lib, color_bits):
    """Gouraud-lit affine triangles with Color32 textures into 32- and 64-bit targets, all blend modes, degenerate edges,
    out-of-range texture coordinates (correction block rasterizer.rs:1018-1192)."""
    r = synth.rng(400 + color_bits)
    ctx = assets.device_context()
    W, H = 257, 130
    n = 400
    tex = [assets.lite_ids[int(r.integers(0, len(assets.lite_ids)))] for _ in range(n)]
    cmds, pts = synth.make_triangle_scene(r, W, H, n, [assets.image_size(t) for t in tex], blends=(0, 1, 2, 3, 4), max_extent=80.0)
    cmds["tex_ref"] = tex
    view = ctx.create_view(W, H, color_bits)
    try:
        view.draw(cmds, None, pts, clear_color=0)
        got = view.read_pixels()
    finally:
        view.close()
    fb = np.zeros((H, W), np.uint32 if color_bits == 32 else np.uint64)
    oracle_lib.reset_assert_failures()
    oracle_lib.draw(color_bits, fb, cmds, None, pts, None, assets.images)
    assert_pixels_equal(got, fb, f"triangles c{color_bits}")


def test_mixed_order_preserved(assets, oracle_lib):
    """Polygons and triangles interleaved: painter's order must be the submission order for every pixel."""
    r = synth.rng(77)
    ctx = assets.device_context()
    W, H = 640, 360
    n = 300
    ptex = [assets.lite_ids[int(r.integers(0, len(assets.lite_ids)))] for _ in range(n)]
    pc, pp = synth.make_polygon_scene(r, W, H, n, [assets.image_size(t) for t in ptex], blends=(0, 0, 0, 1, 2, 3, 4))
    pc["tex_kind"] = N.TEX_IMAGE
    pc["tex_ref"] = ptex
    ttex = [assets.lite_ids[int(r.integers(0, len(assets.lite_ids)))] for _ in range(n)]
    tc, tp = synth.make_triangle_scene(r, W, H, n, [assets.image_size(t) for t in ttex], blends=(0, 4), max_extent=120.0)
    tc["tex_ref"] = ttex
    cmds = np.zeros(2 * n, N.draw_cmd_dtype)
    cmds[0::2] = pc
    cmds[1::2] = tc
    view = ctx.create_view(W, H, 32)
    try:
        view.draw(cmds, pp, tp, clear_color=0xFF000000)
        got = view.read_pixels()
        # second submission without clear: blends read what the first one left behind
        view.draw(cmds[: n // 2], pp, tp, clear_color=None)
        got2 = view.read_pixels()
    finally:
        view.close()
    fb = np.zeros((H, W), np.uint32)
    oracle_lib.fill(fb, 0xFF000000)
    oracle_lib.draw(32, fb, cmds, pp, tp, None, assets.images)
    assert_pixels_equal(got, fb, "mixed")
    oracle_lib.draw(32, fb, cmds[: n // 2], pp, tp, None, assets.images)
    assert_pixels_equal(got2, fb, "mixed, second pass without clear")


def test_surface_textures_and_empty_list(assets, oracle_lib):
    """Polygons textured from the view's surface pool (the normal world-polygon case) + an empty command list."""
    r = synth.rng(91)
    ctx = assets.device_context()
    W, H = 320, 200
    jobs, total, _ = synth.make_surface_jobs(r, 40, assets.texture_ids, assets.texture_flags, directional=True, max_size=64)
    jobs["flags"] |= N.SURF_APPLY_DIFFUSE
    cam = np.zeros(3, np.float32)
    sizes = [tuple(int(x) for x in j["surface_size"]) for j in jobs]
    cmds, pts = synth.make_polygon_scene(r, W, H, len(jobs), sizes, blends=(0,))
    cmds["tex_kind"] = N.TEX_SURFACE
    cmds["tex_ref"] = jobs["surface_pixels_offset"]
    view = ctx.create_view(W, H, 64)
    try:
        view.draw(np.zeros(0, N.draw_cmd_dtype), None, None, clear_color=7)
        assert np.all(view.read_pixels() == 7)
        view.build_surfaces(jobs, None, None, cam, total)
        view.draw(cmds, pts, None, clear_color=0)
        got = view.read_pixels()
    finally:
        view.close()
    surf = oracle_lib.build_surfaces(64, jobs, None, None, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, total)
    fb = np.zeros((H, W), np.uint64)
    oracle_lib.draw(64, fb, cmds, pts, None, surf, assets.images)
    assert_pixels_equal(got, fb, "surface-textured polygons")


def test_crowded_bands(assets, oracle_lib):
    """Thousands of small triangles on a small target: more than 1024 commands per 8-row band (the binning CTAs spill their
    band list to global memory) and tile lists far longer than one 32-entry chunk; order must still be the submission order."""
    r = synth.rng(5150)
    ctx = assets.device_context()
    W, H = 300, 40
    n = 3000
    tex = [assets.lite_ids[int(r.integers(0, len(assets.lite_ids)))] for _ in range(n)]
    cmds, pts = synth.make_triangle_scene(r, W, H, n, [assets.image_size(t) for t in tex], blends=(0, 0, 1, 2, 3, 4), max_extent=60.0)
    cmds["tex_ref"] = tex
    view = ctx.create_view(W, H, 64)
    try:
        view.draw(cmds, None, pts, clear_color=0)
        got = view.read_pixels()
    finally:
        view.close()
    fb = np.zeros((H, W), np.uint64)
    oracle_lib.draw(64, fb, cmds, None, pts, None, assets.images)
    assert_pixels_equal(got, fb, "crowded bands")


def test_device_self_test(assets):
    """sqw_self_test: the reciprocal-based z = 2^31 / (inv_z >> 8) of the full-perspective pixel loop equals the u32 division
    of rasterizer.rs:430 for every divisor below 2^24 (checked exhaustively on the device)."""
    assets.device_context().self_test()


@pytest.mark.parametrize("color_bits", [32, 64])
@pytest.mark.parametrize("clear", [True, False])
def test_opaque_polygon_tiles_crowded(assets, oracle_lib, color_bits, clear):
    """Hundreds of small polygons drawn with BlendingMode::None on a small target: every tile is of the opaque class
    (raster_tile_opaque_kernel: last covering command wins), tile lists are longer than one 32-entry piece and blocks are hit by
    more commands than one stage holds (the chunked path), three tc modes mixed, odd size.  Without clear the uncovered pixels
    keep what the buffer held."""
    r = synth.rng(8800 + color_bits + (1 if clear else 0))
    ctx = assets.device_context()
    W, H = 301, 45
    n = 700
    if color_bits == 32:
        ids = [assets.lite_ids[2]] * 3 + [assets.lite_ids[0]]  # mostly the 8x8 texture: small polygons
    else:
        ids = [assets.sky64_ids[1]] * 3 + [assets.sky64_ids[0]]
    tex = [ids[int(r.integers(0, len(ids)))] for _ in range(n)]
    cmds, pts = synth.make_polygon_scene(r, W, H, n, [assets.image_size(t) for t in tex], blends=(0,))
    cmds["tex_kind"] = N.TEX_IMAGE
    cmds["tex_ref"] = tex
    dtype = np.uint32 if color_bits == 32 else np.uint64
    before = r.integers(1, 2**31, (H, W), dtype=np.uint64).astype(dtype)
    clear_color = 0x0012345600654321 if color_bits == 64 else 0x80402010
    view = ctx.create_view(W, H, color_bits)
    try:
        view.write_pixels(before)
        view.draw(cmds, pts, None, clear_color=clear_color if clear else None)
        got = view.read_pixels()
    finally:
        view.close()
    fb = before.copy()
    if clear:
        oracle_lib.fill(fb, clear_color)
    oracle_lib.reset_assert_failures()
    oracle_lib.draw(color_bits, fb, cmds, pts, None, None, assets.images)
    assert oracle_lib.assert_failures() == 0
    # the case really is crowded: some 128x8 tile sees more than 32 commands and some 32x8 block more than 8
    x = pts["x"].astype(np.float64) / 65536.0
    y = pts["y"].astype(np.float64) / 65536.0
    blocks = np.zeros(((H + 7) // 8, (W + 31) // 32), np.int32)
    tiles = np.zeros(((H + 7) // 8, (W + 127) // 128), np.int32)
    for c in cmds:
        o, k = int(c["vertex_offset"]), int(c["num_vertices"])
        x0, x1 = max(int(np.floor(x[o : o + k].min())), 0), min(int(np.ceil(x[o : o + k].max())), W - 1)
        y0, y1 = max(int(np.floor(y[o : o + k].min())), 0), min(int(np.ceil(y[o : o + k].max())), H - 1)
        if x1 >= x0 and y1 >= y0:
            blocks[y0 // 8 : y1 // 8 + 1, x0 // 32 : x1 // 32 + 1] += 1
            tiles[y0 // 8 : y1 // 8 + 1, x0 // 128 : x1 // 128 + 1] += 1
    assert tiles.max() > 40 and blocks.max() > 12
    assert_pixels_equal(got, fb, f"crowded opaque tiles c{color_bits} clear={clear}")
