"""CPU checks of the oracle (the restatement of surfaces.rs / rasterizer.rs / postprocessor.rs).

The reference ships no golden vectors, so the restatement is pinned two ways:
  * properties that follow from the reference's formulas independently of the C++ code (closed forms in numpy),
  * committed known-answer digests (tests/golden/oracle_kat.json, made by tests/golden/make_golden.py) so that the
    oracle built on another box or with another compiler produces identical bits.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import synth
from common import Assets, assert_pixels_equal, channels32, channels64
from square_wheel_b200 import _native as N
from square_wheel_b200.device import make_post_params

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def assets():
    return Assets(seed=41, num_lightmap_elements=4000, tex_sizes=(32, 16))


def _flat_job(assets, tex_index, su, sv, flags=0, lm_offset=0):
    j = np.zeros(1, N.surface_job_dtype)
    j["basis_u"], j["basis_v"], j["basis_start"], j["basis_normal"] = (0.5, 0, 0), (0, 0.5, 0), (0, 0, 0), (0, 0, 1)
    j["surface_size"] = (su, sv)
    j["texture_image"] = assets.texture_ids[tex_index]
    j["flags"] = assets.texture_flags[tex_index] | N.SURF_APPLY_DIFFUSE | flags
    j["lightmap_size"] = (18, 18)
    j["lightmap_scale_log2"] = 4
    j["lightmap_offset"] = lm_offset
    return j


def test_surface_without_lightmap_is_black(oracle_lib, assets):
    """lightmap_data.is_empty() and no dynamic lights: total light is zero (surfaces.rs:534-535, :833)."""
    j = _flat_job(assets, 0, 20, 9)
    out = oracle_lib.build_surfaces(32, j, None, None, (0, 0, 0), assets.images, assets.lightmaps, assets.directional_lightmaps, 180)
    # alpha lane: diffuse alpha (255) * 0 = 0 as well
    assert np.all(out == 0)


def test_surface_constant_lightmap_closed_form(oracle_lib, assets):
    """Constant simple lightmap L, no normal map, no specular: texel = min(diffuse * L, 255) truncated (surfaces.rs:551-554,
    :833-842, fast_math.rs:472-484); alpha = diffuse alpha * 1."""
    lm = assets.lightmaps.copy()
    lm[:] = (0.5, 1.0, 3.0)  # [R, G, B]
    j = _flat_job(assets, 0, 31, 17, lm_offset=1)
    j["tc_start"] = (-5, 40)
    out = oracle_lib.build_surfaces(32, j, None, None, (0, 0, 0), assets.images, lm, assets.directional_lightmaps, 31 * 17).reshape(17, 31)
    tex = assets.images[assets.texture_ids[0]][1]
    th, tw = tex.shape
    vv = (np.arange(17) + 40) % th
    uu = (np.arange(31) - 5) % tw
    d = channels32(tex["diffuse"][np.ix_(vv, uu)]).astype(np.float32)  # lanes B, G, R, A
    light = np.array([3.0, 1.0, 0.5, 1.0], np.float32)
    want = np.minimum(d * light, np.float32(255.0)).astype(np.uint32)
    got = channels32(out)
    assert np.array_equal(got, want)
    # the 64-bit surface keeps values above 255
    out64 = oracle_lib.build_surfaces(64, j, None, None, (0, 0, 0), assets.images, lm, assets.directional_lightmaps, 31 * 17).reshape(17, 31)
    want64 = np.minimum(d * light, np.float32(65535.0)).astype(np.uint64)
    assert np.array_equal(channels64(out64), want64)


def test_lightmap_bilinear_weights(oracle_lib, assets):
    """k = ((v + shift) & mask) / scale + 0.5 / scale and mix(l1, l0, k) (surfaces.rs:483-502, :544-549) on a lightmap that is
    linear in u: the interpolated light must be linear in the texel index."""
    lm = np.zeros_like(assets.lightmaps)
    lw = 18
    ramp = np.arange(lw, dtype=np.float32) / 4.0
    for v in range(18):
        lm[1 + v * lw : 1 + (v + 1) * lw, :] = ramp[:, None]
    j = _flat_job(assets, 0, 64, 3, lm_offset=1)
    j["lightmap_scale_log2"] = 2
    j["lightmap_tc_shift"] = (1, 0)
    # white texture so the surface value is the light itself
    imgs = list(assets.images)
    tex = imgs[assets.texture_ids[0]][1].copy()
    tex["diffuse"] = 0xFFFFFFFF
    imgs[assets.texture_ids[0]] = (N.IMAGE_TEXTURE_ELEMENT, tex)
    out = oracle_lib.build_surfaces(64, j, None, None, (0, 0, 0), imgs, lm, assets.directional_lightmaps, 64 * 3).reshape(3, 64)
    u = np.arange(64) + 1
    k = ((u & 3).astype(np.float32) * np.float32(0.25)) + np.float32(0.125)
    l0 = ramp[u >> 2]
    l1 = ramp[(u >> 2) + 1]
    light = l1 * k + l0 * (np.float32(1.0) - k)
    want = np.minimum(np.float32(255.0) * light, np.float32(65535.0)).astype(np.uint64)
    assert np.array_equal(channels64(out[1])[:, 0], want)


def test_polygon_coverage_matches_pixel_centre_rule(oracle_lib, assets):
    """Coverage: a pixel is filled iff its centre is inside the polygon (rasterizer.rs:316-331 round-to-centre rule).  Checked
    against an analytic half-plane test for pixels whose centre is not within 1/256 px of an edge."""
    r = synth.rng(8)
    W, H = 200, 150
    tex_id = assets.lite_ids[0]
    white = np.full((32, 32), 0xFFFFFFFF, np.uint32)
    imgs = list(assets.images)
    imgs[tex_id] = (N.IMAGE_COLOR32, white)
    for trial in range(30):
        cmds, pts = synth.make_polygon_scene(r, W, H, 1, [(32, 32)], modes=(N.TC_AFFINE,), blends=(0,), allow_offscreen=False)
        cmds["tex_kind"], cmds["tex_ref"] = N.TEX_IMAGE, tex_id
        fb = np.zeros((H, W), np.uint32)
        oracle_lib.draw(32, fb, cmds, pts, None, None, imgs)
        px = pts["x"].astype(np.float64) / 65536.0
        py = pts["y"].astype(np.float64) / 65536.0
        yy, xx = np.mgrid[0:H, 0:W]
        cx, cy = xx + 0.5, yy + 0.5
        inside = np.ones((H, W), bool)
        near = np.zeros((H, W), bool)
        n = len(px)
        for i in range(n):
            x0, y0, x1, y1 = px[i], py[i], px[(i + 1) % n], py[(i + 1) % n]
            cross = (x1 - x0) * (cy - y0) - (y1 - y0) * (cx - x0)
            length = np.hypot(x1 - x0, y1 - y0)
            inside &= cross > 0
            near |= np.abs(cross) < length * 0.02
        covered = fb != 0
        assert np.array_equal(covered[~near], inside[~near]), f"trial {trial}"


def test_blend_modes_closed_form(oracle_lib, assets):
    """write_into_framebuffer (rasterizer.rs:1511-1536) against numpy formulas from common/color.rs and abstract_color.rs."""
    r = synth.rng(15)
    W, H = 64, 64
    tex_id = assets.lite_ids[2]  # 8x8
    tex = assets.images[tex_id][1]
    quad = np.array([(0, 0), (W << 16, 0), (W << 16, H << 16), (0, H << 16)], N.polygon_point_dtype)
    for blend in range(5):
        cmds = np.zeros(1, N.draw_cmd_dtype)
        c = cmds[0]
        c["kind"], c["blend"], c["tc_mode"], c["tex_kind"], c["tex_ref"] = N.CMD_POLYGON, blend, N.TC_AFFINE, N.TEX_IMAGE, tex_id
        c["num_vertices"], c["tex_size"] = 4, (8, 8)
        c["depth_eq"] = (0, 0, 1)
        c["tc_eq"] = (0, 0, 0, 0, 3.5, 2.5)  # constant texel (3, 2)
        dst = r.integers(0, 2**32, (H, W), dtype=np.uint64).astype(np.uint32)
        fb = dst.copy()
        oracle_lib.draw(32, fb, cmds, quad, None, None, assets.images)
        t = np.uint32(tex[2, 3])
        d, s = channels32(dst), channels32(np.full((H, W), t))
        if blend == N.BLEND_NONE:
            want = s
        elif blend == N.BLEND_AVERAGE:
            want = (d + s) >> 1
        elif blend == N.BLEND_ADDITIVE:
            want = np.minimum(d + s, 255)
        elif blend == N.BLEND_ALPHA_TEST:
            want = s if (int(t) >> 24) > 0x7F else d
        else:
            a = int(t) >> 24
            want = (d * (255 - a) + s * a) >> 8
        assert np.array_equal(channels32(fb), want), f"blend {blend}"


def test_tonemap_closed_form(oracle_lib):
    """Reinhard c / (c/255 + 1/(exposure*modulate)) (postprocessor.rs:1005-1023), alpha forced to 255 by min(255.0)."""
    r = synth.rng(4)
    hdr = synth.make_hdr_image(r, 50, 20)
    p = make_post_params(exposure=0.8, color_modulate=(1.1, 1.0, 0.7))
    out, avg, _ = oracle_lib.postprocess(hdr, p)
    c = channels64(hdr).astype(np.float32)
    inv_scale = np.array([1.0 / (np.float32(0.8) * np.float32(0.7)), 1.0 / (np.float32(0.8) * np.float32(1.0)), 1.0 / (np.float32(0.8) * np.float32(1.1))], np.float32)
    mapped = c[..., :3] / (c[..., :3] * np.float32(1.0 / 255.0) + inv_scale)
    want = np.minimum(mapped, np.float32(255.0)).astype(np.uint32)
    got = channels32(out)
    assert np.array_equal(got[..., :3], want)
    assert np.all(got[..., 3] == 255)
    np.testing.assert_allclose(avg, c[..., :3].reshape(-1, 3).mean(0), rtol=1e-5)


def test_multithreaded_oracle_matches_single_thread_where_defined(oracle_lib, assets):
    """Surfaces and tonemapping do not depend on the thread count; rasterisation does (span end points
    follow the per-thread clip rect) but coverage must be identical."""
    r = synth.rng(23)
    jobs, total, _ = synth.make_surface_jobs(r, 60, assets.texture_ids, assets.texture_flags, directional=True, max_size=48)
    a = oracle_lib.build_surfaces(64, jobs, None, None, (1, 2, 3), assets.images, assets.lightmaps, assets.directional_lightmaps, total, num_threads=1)
    b = oracle_lib.build_surfaces(64, jobs, None, None, (1, 2, 3), assets.images, assets.lightmaps, assets.directional_lightmaps, total, num_threads=4)
    assert_pixels_equal(a, b, "surfaces 1 vs 4 threads")
    hdr = synth.make_hdr_image(r, 97, 41)
    p = make_post_params(exposure=1.0, bloom_sigma=3.0)
    o1, _, _ = oracle_lib.postprocess(hdr, p, num_threads=1)
    o4, _, _ = oracle_lib.postprocess(hdr, p, num_threads=4)
    assert_pixels_equal(o1, o4, "bloom tonemap 1 vs 4 threads")
    W, H = 160, 120
    ids = [assets.lite_ids[0]] * 40
    cmds, pts = synth.make_polygon_scene(r, W, H, 40, [assets.image_size(i) for i in ids], blends=(0,))
    cmds["tex_kind"], cmds["tex_ref"] = N.TEX_IMAGE, ids
    f1 = np.zeros((H, W), np.uint32)
    f4 = np.zeros((H, W), np.uint32)
    oracle_lib.draw(32, f1, cmds, pts, None, None, assets.images, num_threads=1)
    oracle_lib.draw(32, f4, cmds, pts, None, None, assets.images, num_threads=4)
    assert np.array_equal(f1 != 0, f4 != 0)


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kat_cases(oracle_lib):
    """Known-answer cases shared by make_golden.py and the test below."""
    out = {}
    assets = Assets(seed=1234, num_lightmap_elements=30000, tex_sizes=(32, 16))
    for color_bits in (32, 64):
        for directional in (False, True):
            r = synth.rng(9000 + color_bits + directional)
            lights, sm = synth.make_lights(r, 6)
            jobs, total, _ = synth.make_surface_jobs(r, 96, assets.texture_ids, assets.texture_flags, emissive_images=assets.lite_ids, directional=directional, lights=6, max_size=40)
            s = oracle_lib.build_surfaces(color_bits, jobs, lights, sm, (5, 6, 7), assets.images, assets.lightmaps, assets.directional_lightmaps, total)
            out[f"surfaces_c{color_bits}_dir{int(directional)}"] = _digest(s)
        for mode in (0, 1, 2):
            r = synth.rng(9100 + color_bits + mode)
            ids = assets.lite_ids if color_bits == 32 else assets.sky64_ids
            tex = [ids[int(r.integers(0, len(ids)))] for _ in range(60)]
            cmds, pts = synth.make_polygon_scene(r, 211, 133, 60, [assets.image_size(t) for t in tex], modes=(mode,), blends=(0, 1, 2, 3, 4))
            cmds["tex_kind"], cmds["tex_ref"] = N.TEX_IMAGE, tex
            fb = np.zeros((133, 211), np.uint32 if color_bits == 32 else np.uint64)
            oracle_lib.draw(color_bits, fb, cmds, pts, None, None, assets.images)
            out[f"polygons_c{color_bits}_mode{mode}"] = _digest(fb)
        r = synth.rng(9200 + color_bits)
        tex = [assets.lite_ids[int(r.integers(0, 3))] for _ in range(200)]
        cmds, pts = synth.make_triangle_scene(r, 211, 133, 200, [assets.image_size(t) for t in tex], blends=(0, 1, 2, 3, 4), max_extent=60.0)
        cmds["tex_ref"] = tex
        fb = np.zeros((133, 211), np.uint32 if color_bits == 32 else np.uint64)
        oracle_lib.draw(color_bits, fb, cmds, None, pts, None, assets.images)
        out[f"triangles_c{color_bits}"] = _digest(fb)
    r = synth.rng(9300)
    hdr = synth.make_hdr_image(r, 131, 75)
    for name, p in (
        ("tonemap", make_post_params(exposure=0.9)),
        ("bloom_nearest_s2", make_post_params(bloom_sigma=3.0, bloom_buffer_scale_log2=1)),
        ("bloom_linear_s4", make_post_params(bloom_sigma=5.0, bloom_buffer_scale_log2=2, linear_bloom_filter=True)),
        ("bloom_linear_s8", make_post_params(bloom_sigma=9.0, bloom_buffer_scale_log2=3, linear_bloom_filter=True)),
    ):
        o, _, _ = oracle_lib.postprocess(hdr, p)
        out[f"post_{name}"] = _digest(o)
    return out


def test_known_answer_digests(oracle_lib):
    path = os.path.join(HERE, "golden", "oracle_kat.json")
    golden = json.load(open(path))
    oracle_lib.reset_assert_failures()
    got = kat_cases(oracle_lib)
    assert got == golden["digests"], {k: (got[k], golden["digests"].get(k)) for k in got if got[k] != golden["digests"].get(k)}


def test_depth_rasterizer_closed_forms(oracle_lib):
    """DepthRasterizer (rasterizer.rs:1295-1474): a screen-filling polygon with a constant depth equation writes that constant
    everywhere; DEPTH_TEST keeps the larger inverse depth; the projector circle mask is symmetric and leaves the centre alone."""
    from square_wheel_b200 import _native as N

    w = h = 32
    pts = np.array([(-65536, -65536), ((w + 1) << 16, -65536), ((w + 1) << 16, (h + 1) << 16), (-65536, (h + 1) << 16)], N.polygon_point_dtype)
    maps = np.zeros(1, N.depth_map_dtype)
    maps[0] = (0, w, h, 0, 2, N.DEPTH_MAP_CLEAR | N.DEPTH_MAP_PROJECTOR_CIRCLE, 0.0, 0)
    cmds = np.zeros(2, N.depth_cmd_dtype)
    cmds["num_vertices"] = 4
    cmds[0]["depth_eq"] = (0.0, 0.0, 0.25)
    cmds[1]["depth_eq"] = (1.0 / 64.0, 0.0, 0.0)  # inv_z = (x + 0.5) / 64: larger than 0.25 for x >= 16
    cmds[1]["depth_test"] = 1
    out = oracle_lib.draw_shadow_maps(np.full(w * h, -1.0, np.float32), maps, cmds, pts).reshape(h, w)
    row = out[h // 2]
    assert np.all(row[:16] == 0.25)
    assert np.allclose(row[16:], (np.arange(16, 32) + 0.5) / 64.0, rtol=1e-6)
    assert out[0, 0] == np.float32(1.0e24) and out[0, w - 1] == np.float32(1.0e24) and out[h - 1, 0] == np.float32(1.0e24)
    assert np.array_equal(out == np.float32(1.0e24), (out == np.float32(1.0e24))[:, ::-1])


# ---- the second numerics mode: the branch of fast_math.rs the reference's release build runs (liboracle_sse.so) ----------------
def test_sse_tonemapping_is_the_release_builds_formula(oracle_lib):
    """numerics="sse": perform_tonemapping with the release build's operations (fast_math.rs:123-146, :208-213): ONE fused
    multiply-add c * (1/255) + 1/(exposure * modulate), an IEEE division, min(255), cvtps2dq (round to nearest EVEN).  Checked
    against that formula evaluated in numpy (products of a 16-bit integer and an f32 are exact in f64, so the f64 sum rounded to
    f32 is the fused result up to a double rounding that random data does not hit more than a handful of times)."""
    r = synth.rng(4321)
    hdr = synth.make_hdr_image(r, 257, 63)
    exposure, modulate = 0.7, (1.0, 0.9, 1.2)
    p = make_post_params(exposure=exposure, color_modulate=modulate)
    got, avg, _ = oracle_lib.postprocess(hdr, p, numerics="sse")
    c = channels64(hdr)[..., :3].astype(np.float64)  # B, G, R
    inv_scale = np.array([np.float32(1.0) / (np.float32(exposure) * np.float32(m)) for m in modulate[::-1]], np.float32)
    den = (c * np.float64(np.float32(1.0) / np.float32(255.0)) + inv_scale.astype(np.float64)).astype(np.float32)
    mapped = np.minimum((c.astype(np.float32) / den).astype(np.float32), np.float32(255.0))
    want = np.rint(mapped.astype(np.float64)).astype(np.int64)  # rint = round half to even, like cvtps2dq
    d = np.abs(channels32(got)[..., :3] - want)
    assert d.max() <= 1 and (d != 0).mean() < 1e-4
    assert (channels32(got)[..., 3] == 255).all()  # alpha lane: x / 0 -> inf or NaN -> min_ps(., 255) = 255
    # the portable branch truncates instead: the two builds differ by at most one LSB of the picture
    port, avg_p, _ = oracle_lib.postprocess(hdr, p)
    assert np.abs(channels32(got)[..., :3] - channels32(port)[..., :3]).max() <= 1
    np.testing.assert_allclose(avg, avg_p, rtol=1e-6)


@pytest.mark.parametrize("directional", [False, True])
def test_sse_surfaces_stay_within_the_spread_of_the_two_builds(oracle_lib, directional):
    """What north_star's tolerance means in the reference's own terms: the release build (rcpss / rsqrtss with 1.5 * 2^-12 relative
    error, FMA, rounding conversions) against the portable build on static surfaces — a few HDR units (255 = LDR white), a few
    1e-3 relative on bright texels.  This is the bound the fast GPU surface kernels are tested against."""
    a = Assets(seed=11, num_lightmap_elements=60000)
    r = synth.rng(77 + directional)
    jobs, total, need = synth.make_surface_jobs(r, 200, a.texture_ids, a.texture_flags, emissive_images=a.lite_ids, directional=directional, lights=0, max_size=96)
    assert need <= len(a.lightmaps)
    cam = np.array([10.0, 20.0, 30.0], np.float32)
    p = channels64(oracle_lib.build_surfaces(64, jobs, None, None, cam, a.images, a.lightmaps, a.directional_lightmaps, total))[..., :3]
    q = channels64(oracle_lib.build_surfaces(64, jobs, None, None, cam, a.images, a.lightmaps, a.directional_lightmaps, total, numerics="sse"))[..., :3]
    d = np.abs(p - q)
    assert d.max() <= 4
    assert (d > 1).mean() < 0.05
    big = p >= 1024
    assert (d[big] <= 3e-3 * p[big]).all()
    assert (p != q).mean() > 0.3  # and it IS a different build: most texels round differently


def test_sse_average_blend_rounds_up(oracle_lib):
    """BlendingMode::Average in the release build is pavgb / pavgw, (a + b + 1) >> 1 per channel (fast_math.rs:43-63); the
    portable branch rounds down (color.rs:46-49).  One polygon over a constant background shows the difference."""
    a = Assets(seed=5, num_lightmap_elements=16, tex_sizes=(16,))
    W, H = 64, 32
    tex = np.zeros((8, 8), np.uint64)
    tex[:] = (np.uint64(255) << np.uint64(48)) | (np.uint64(7) << np.uint64(32)) | (np.uint64(4) << np.uint64(16)) | np.uint64(1)
    images = list(a.images) + [(N.IMAGE_COLOR64, tex)]
    cmds = np.zeros(1, N.draw_cmd_dtype)
    cmds[0]["kind"], cmds[0]["blend"], cmds[0]["tc_mode"], cmds[0]["num_vertices"] = N.CMD_POLYGON, N.BLEND_AVERAGE, N.TC_AFFINE, 4
    cmds[0]["tex_size"], cmds[0]["tex_kind"], cmds[0]["tex_ref"] = (8, 8), N.TEX_IMAGE, len(images) - 1
    cmds[0]["depth_eq"] = (0.0, 0.0, 1.0)
    cmds[0]["tc_eq"] = (0.1, 0.0, 0.0, 0.1, 0.0, 0.0)
    pts = np.array([(0, 0), (W << 16, 0), (W << 16, H << 16), (0, H << 16)], N.polygon_point_dtype)
    bg = (np.uint64(9) << np.uint64(32)) | (np.uint64(2) << np.uint64(16)) | np.uint64(4)
    out = {}
    for mode in ("portable", "sse"):
        fb = np.full((H, W), bg, np.uint64)
        oracle_lib.draw(64, fb, cmds, pts, None, None, images, numerics=mode)
        out[mode] = channels64(fb)[H // 2, W // 2, :3].tolist()
    assert out["portable"] == [(1 + 4) // 2, (4 + 2) // 2, (7 + 9) // 2]
    assert out["sse"] == [(1 + 4 + 1) // 2, (4 + 2 + 1) // 2, (7 + 9 + 1) // 2]
