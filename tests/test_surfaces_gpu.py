"""Stage 1 parity: surface_build_kernel vs the CPU restatement of surfaces.rs, through the C-ABI (sqw_build_surfaces)."""
import numpy as np
import pytest

import synth
from common import Assets, assert_pixels_equal
from square_wheel_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def assets():
    return Assets(seed=11)


@pytest.mark.parametrize("color_bits", [32, 64])
@pytest.mark.parametrize("directional", [False, True])
@pytest.mark.parametrize("with_lights", [False, True])
def test_surface_variants_bit_exact(assets, oracle_lib, color_bits, directional, with_lights):
    """Every (colour width x lightmap kind x dynamic lights) combination; normal map / specular type / lightmap scale /
    no-lightmap / emissive vary per job inside each call (2 x 3 x 6 x emissive).  Mirrors the 288 monomorphs of
    surfaces.rs:124-423."""
    r = synth.rng(100 + color_bits + 7 * directional + 13 * with_lights)
    ctx = assets.device_context()
    lights, sm = synth.make_lights(r, 9) if with_lights else (np.zeros(0, N.light_dtype), np.zeros(0, np.float32))
    jobs, total, lm_needed = synth.make_surface_jobs(
        r, 160, assets.texture_ids, assets.texture_flags, emissive_images=assets.lite_ids, directional=directional, lights=len(lights), max_size=96
    )
    assert lm_needed <= len(assets.lightmaps)
    cam = r.uniform(-300, 300, 3).astype(np.float32)
    view = ctx.create_view(64, 64, color_bits)
    try:
        view.build_surfaces(jobs, lights, sm, cam, total)
        got = view.read_surfaces(0, total)
    finally:
        view.close()
    oracle_lib.reset_assert_failures()
    want = oracle_lib.build_surfaces(color_bits, jobs, lights, sm, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, total)
    assert oracle_lib.assert_failures() == 0
    # jobs with neither diffuse nor emissive layer leave their texels untouched in both implementations (zeros here)
    assert_pixels_equal(got, want, f"surfaces c{color_bits} dir{directional} lights{with_lights}")


@pytest.mark.parametrize("num_jobs, band", [(300, 64), (5000, 128), (11000, 256)])
def test_small_surfaces_on_warp_bands(oracle_lib, num_jobs, band):
    """Surfaces of at most 1024 texels are built one WARP per band (eight bands per CTA, csrc/surface_kernels.cu); the band size
    follows the number of small texels in the frame (64 / 128 / 256 texels: sqw_frame_set_surfaces).  All three sizes, with
    dynamic lights on a third of the jobs so that both grids carry warp bands."""
    from common import Assets

    a = Assets(seed=31, num_lightmap_elements=1600000, tex_sizes=(32,))
    r = synth.rng(500 + num_jobs)
    ctx = a.device_context()
    try:
        lights, sm = synth.make_lights(r, 5)
        jobs, total, lm_needed = synth.make_surface_jobs(r, num_jobs, a.texture_ids, a.texture_flags, emissive_images=a.lite_ids, directional=True, lights=len(lights), max_size=32)
        assert lm_needed <= len(a.lightmaps)
        jobs["num_lights"][::3] = 0
        small = int((jobs["surface_size"].astype(np.int64).prod(axis=1)).sum())
        resident_warps = 148 * 4 * 8
        expect = 256
        while expect > 64 and small < resident_warps * expect:
            expect //= 2
        assert expect == band, f"{small} small texels give {expect}-texel bands: adjust the job count of this case"
        cam = r.uniform(-300, 300, 3).astype(np.float32)
        view = ctx.create_view(64, 64, 64)
        try:
            view.build_surfaces(jobs, lights, sm, cam, total)
            got = view.read_surfaces(0, total)
        finally:
            view.close()
        oracle_lib.reset_assert_failures()
        want = oracle_lib.build_surfaces(64, jobs, lights, sm, cam, a.images, a.lightmaps, a.directional_lightmaps, total)
        assert oracle_lib.assert_failures() == 0
        assert_pixels_equal(got, want, f"{num_jobs} small surfaces, {band}-texel warp bands")
    finally:
        ctx.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_non_power_of_two_textures(oracle_lib, color_bits):
    """Texture and emissive images whose sizes are not powers of two take the general rem_euclid wrap (surfaces.rs:507-510, :843-847;
    an out-of-line integer division in the kernel) instead of the mask shortcut; negative and huge tc_start included."""
    from common import Assets

    a = Assets(seed=23, num_lightmap_elements=20000, tex_sizes=(24, 40))
    r = synth.rng(77)
    for (w, h) in ((20, 12), (7, 33)):
        a.lite_ids.append(len(a.images))
        a.images.append((N.IMAGE_COLOR32, synth.make_lite_texture(r, w, h)))
    ctx = a.device_context()
    try:
        jobs, total, lm_needed = synth.make_surface_jobs(r, 96, a.texture_ids, a.texture_flags, emissive_images=a.lite_ids[-2:], directional=True, lights=0, max_size=80)
        assert lm_needed <= len(a.lightmaps)
        jobs["tc_start"][::3] = r.integers(-100000, -1, (len(jobs[::3]), 2))
        jobs["tc_start"][1::3] = r.integers(1 << 20, 1 << 24, (len(jobs[1::3]), 2))
        cam = r.uniform(-300, 300, 3).astype(np.float32)
        view = ctx.create_view(64, 64, color_bits)
        try:
            view.build_surfaces(jobs, None, None, cam, total)
            got = view.read_surfaces(0, total)
        finally:
            view.close()
        oracle_lib.reset_assert_failures()
        want = oracle_lib.build_surfaces(color_bits, jobs, None, None, cam, a.images, a.lightmaps, a.directional_lightmaps, total)
        assert oracle_lib.assert_failures() == 0
        assert_pixels_equal(got, want, f"non power-of-two textures c{color_bits}")
    finally:
        ctx.close()


def test_large_surfaces_and_wrap(assets, oracle_lib):
    """Maximum lit surface size (272 x 272 at mip 0, lightmap 18 x 18), negative tc_start, texture smaller than surface."""
    r = synth.rng(5)
    ctx = assets.device_context()
    jobs = np.zeros(3, N.surface_job_dtype)
    off = 0
    for i, (su, sv, scale, tc) in enumerate([(272, 272, 4, (-1000, -77)), (1, 1, 0, (5, 5)), (33, 136, 3, (123456, -98765))]):
        j = jobs[i]
        j["basis_u"], j["basis_v"], j["basis_start"], j["basis_normal"] = (0.5, 0, 0), (0, 0.5, 0), (10, 20, 30), (0, 0, 1)
        j["surface_size"] = (su, sv)
        j["tc_start"] = tc
        j["texture_image"] = assets.texture_ids[-2]  # 16x16 normal-mapped dielectric
        j["flags"] = assets.texture_flags[-2] | N.SURF_APPLY_DIFFUSE | N.SURF_DIRECTIONAL_LIGHTMAP
        j["lightmap_size"] = (18, 18)
        j["lightmap_scale_log2"] = scale
        j["lightmap_offset"] = 1 + i * 400
        j["surface_pixels_offset"] = off
        off += su * sv
    cam = np.array([100.0, -50.0, 200.0], np.float32)
    view = ctx.create_view(16, 16, 64)
    try:
        view.build_surfaces(jobs, None, None, cam, off)
        got = view.read_surfaces(0, off)
    finally:
        view.close()
    want = oracle_lib.build_surfaces(64, jobs, None, None, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, off)
    assert_pixels_equal(got, want, "large surfaces")


def test_empty_job_list(assets):
    ctx = assets.device_context()
    view = ctx.create_view(16, 16, 32)
    try:
        view.build_surfaces(np.zeros(0, N.surface_job_dtype), None, None, (0, 0, 0), 0)
        view.synchronize()
    finally:
        view.close()


def test_bad_jobs_are_rejected(assets):
    """Error behaviour: the ABI returns an error code instead of reading out of bounds (reference: debug_assert / UB)."""
    from square_wheel_b200.device import SqwError

    ctx = assets.device_context()
    view = ctx.create_view(16, 16, 32)
    try:
        j = np.zeros(1, N.surface_job_dtype)
        j["surface_size"] = (64, 64)
        j["flags"] = N.SURF_APPLY_DIFFUSE
        j["texture_image"] = assets.texture_ids[0]
        j["lightmap_size"] = (2, 2)
        j["lightmap_scale_log2"] = 4
        j["lightmap_offset"] = 1  # 64 texels at scale 16 need 5 lightmap texels per axis
        with pytest.raises(SqwError):
            view.build_surfaces(j, None, None, (0, 0, 0), 64 * 64)
        j["lightmap_offset"] = 0
        j["texture_image"] = assets.lite_ids[0]  # wrong image kind
        with pytest.raises(SqwError):
            view.build_surfaces(j, None, None, (0, 0, 0), 64 * 64)
    finally:
        view.close()


@pytest.mark.parametrize("color_bits", [32, 64])
def test_view_mix_texture_all_blending_modes(assets, oracle_lib, color_bits):
    """mix_surface_with_texture (surfaces.rs:854-971) as the portal / mirror glass overlay (partial_renderer.rs:2261-2271): every
    blending mode, wrapped negative texture coordinates, a view resized below its allocated size, bit-exact."""
    ctx = assets.device_context()
    r = synth.rng(77 + color_bits)
    view = ctx.create_view(200, 120, color_bits)
    dt = np.uint32 if color_bits == 32 else np.uint64
    for k, (w, h) in enumerate([(200, 120), (131, 57), (64, 64), (17, 9), (1, 1)]):
        view.set_size(w, h)
        for mode in range(5):
            if color_bits == 32:
                base = r.integers(0, 2**32, size=(h, w), dtype=np.uint64).astype(np.uint32)
            else:
                base = (r.integers(0, 4096, size=(h, w, 4), dtype=np.uint64) << (np.arange(4, dtype=np.uint64) * 16)).sum(axis=2).astype(np.uint64)
            view.write_pixels(base)
            tc_min = [int(r.integers(-300, 300)), int(r.integers(-300, 300))]
            light = r.uniform(0.0, 3.0, 3).astype(np.float32)
            image = assets.lite_ids[(k + mode) % len(assets.lite_ids)]
            view.mix_texture(image, tc_min, mode, light)
            got = view.read_pixels()
            want = base.copy().reshape(-1)
            oracle_lib.mix_surface_with_texture(color_bits, want, [w, h], tc_min, assets.images[image][1], mode, light)
            assert_pixels_equal(got, want.reshape(h, w).astype(dt), f"mix {w}x{h} mode {mode}")
    with pytest.raises(Exception):
        view.set_size(201, 120)
    view.close()


def test_shadow_map_references_are_validated(assets):
    """A light whose shadow map lies outside the frame's shadow map array is refused on the host (the kernel fetches
    sm[offset + side * size^2 + u + v * size] unchecked, surface_kernels.cu cube / projector fetch): truncated array, zero size,
    offset past the end, shadow type set with no array at all, unknown shadow type."""
    from square_wheel_b200.device import SqwError

    r = synth.rng(515)
    ctx = assets.device_context()
    lights, sm = synth.make_lights(r, 6)
    jobs, total, _ = synth.make_surface_jobs(r, 8, assets.texture_ids, assets.texture_flags, directional=True, lights=len(lights), max_size=32)
    cam = np.zeros(3, np.float32)
    view = ctx.create_view(32, 32, 64)
    try:
        view.build_surfaces(jobs, lights, sm, cam, total)  # the valid set passes
        with pytest.raises(SqwError):
            view.build_surfaces(jobs, lights, sm[:-1], cam, total)
        with pytest.raises(SqwError):
            view.build_surfaces(jobs, lights, None, cam, total)
        for field, value in (("shadow_map_size", 0), ("shadow_map_offset", len(sm)), ("shadow_map_size", 1 << 15), ("shadow_type", 7)):
            bad = lights.copy()
            bad[1][field] = value  # light 1 has a cube shadow map
            with pytest.raises(SqwError):
                view.build_surfaces(jobs, bad, sm, cam, total)
        view.build_surfaces(jobs, lights, sm, cam, total)  # and the view still works afterwards
    finally:
        view.close()


def test_view_reshape_keeps_bloom_inside_its_buffers(assets, oracle_lib):
    """sqw_view_set_size bounds the AREA: 1921 x 1081 reshaped to 2076 x 1000 needs more half-resolution bloom texels than
    (1921 / 2) * (1081 / 2); the buffers are sized for any shape of the allocated area (checked small: 61 x 33 -> 66 x 30)."""
    from square_wheel_b200.device import make_post_params

    r = synth.rng(616)
    ctx = assets.device_context()
    view = ctx.create_view(61, 33, 64)
    try:
        assert (66 // 2) * (30 // 2) > (61 // 2) * (33 // 2) and 66 * 30 <= 61 * 33
        view.set_size(66, 30)
        hdr = synth.make_hdr_image(r, 66, 30)
        view.write_pixels(hdr)
        p = make_post_params(exposure=1.1, bloom_sigma=3.0, bloom_buffer_scale_log2=1)
        got, _ = view.postprocess(p)
        want, _, _ = oracle_lib.postprocess(hdr, p)
        assert_pixels_equal(got, want, "bloom on a reshaped view")
    finally:
        view.close()


def _channels64(x):
    return np.stack([(x >> np.uint64(s)) & np.uint64(0xFFFF) for s in (0, 16, 32)], -1).astype(np.int64)


@pytest.mark.parametrize("directional", [False, True])
@pytest.mark.parametrize("with_lights", [False, True])
def test_fast_surface_numerics_within_the_reference_builds_spread(assets, oracle_lib, directional, with_lights):
    """SQW_CTX_FAST_SURFACE_NUMERICS: hardware reciprocal / reciprocal square root and fused multiply-add, the numerics CLASS of
    the reference's release build (fast_math.rs:6-368).  The tolerance is north_star's, made concrete by the reference itself:
    its two builds (oracle numerics "portable" = fast_math.rs:370-725 and "sse" = :6-368, rcpss / rsqrtss / FMA / cvtps2dq)
    differ from each other by up to 3 HDR units (255 = LDR white) and 2e-3 relative on static surfaces, more where an
    approximate reciprocal flips a shadow-map comparison.  The fast kernels must stay within

        |gpu - portable| <= 2 units or 1e-3 relative           for every channel of surfaces without dynamic lights,
        the same for all but 1e-3 of the channels               with dynamic lights (shadow-map comparisons flip in both builds too),

    and must be no further from the SSE build than the portable build is.  Every variant (normal map x specular kind x lightmap
    scale x emissive) appears in the job list."""
    from square_wheel_b200.device import Context

    r = synth.rng(900 + 7 * directional + 13 * with_lights)
    ctx = Context(assets.lightmaps, assets.directional_lightmaps, device=0, fast_surface_numerics=True)
    try:
        for i, (kind, px) in enumerate(assets.images):
            assert ctx.upload_image(kind, px) == i
        lights, sm = synth.make_lights(r, 9) if with_lights else (np.zeros(0, N.light_dtype), np.zeros(0, np.float32))
        jobs, total, _ = synth.make_surface_jobs(
            r, 200, assets.texture_ids, assets.texture_flags, emissive_images=assets.lite_ids, directional=directional, lights=len(lights), max_size=96
        )
        cam = r.uniform(-300, 300, 3).astype(np.float32)
        view = ctx.create_view(64, 64, 64)
        view.build_surfaces(jobs, lights, sm, cam, total)
        got = _channels64(view.read_surfaces(0, total))
        view.close()
    finally:
        ctx.close()
    portable = _channels64(oracle_lib.build_surfaces(64, jobs, lights, sm, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, total))
    sse = _channels64(oracle_lib.build_surfaces(64, jobs, lights, sm, cam, assets.images, assets.lightmaps, assets.directional_lightmaps, total, numerics="sse"))
    d = np.abs(got - portable)
    ok = (d <= 2) | (d <= 1e-3 * portable)
    bad = 1.0 - ok.mean()
    assert bad <= (1e-3 if with_lights else 0.0), f"{bad:.2e} of the channels outside 2 units / 1e-3 (max {d.max()})"
    assert (d == 0).mean() > 0.9  # and most texels are simply equal: only rcp / rsqrt / FMA roundings differ
    # distance to the release-build restatement: not worse than the portable build's own distance to it
    d_sse, d_ref = np.abs(got - sse), np.abs(portable - sse)
    assert d_sse.mean() <= d_ref.mean() * 1.05 + 1e-3
    assert (d_sse > 3).mean() <= (d_ref > 3).mean() + (1e-3 if with_lights else 1e-5)
