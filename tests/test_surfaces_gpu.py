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
