"""Stage 3 parity: tonemap / bloom kernels vs the CPU restatement of postprocessor.rs, through sqw_postprocess."""
import numpy as np
import pytest

import synth
from common import Assets, assert_pixels_equal
from square_wheel_b200 import _native as N
from square_wheel_b200.device import make_post_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def assets():
    return Assets(seed=31, num_lightmap_elements=16, tex_sizes=(16,))


def _put_hdr(ctx, assets, hdr):
    """Hand a host-rendered Color64 picture to the device (the write half of get_hdr_buffer, postprocessor.rs:56)."""
    h, w = hdr.shape
    view = ctx.create_view(w, h, 64)
    view.write_pixels(hdr)
    return view


@pytest.mark.parametrize("size", [(64, 48), (333, 77), (1, 1), (257, 129)])
def test_tonemap_bit_exact(assets, oracle_lib, size):
    r = synth.rng(500 + size[0])
    ctx = assets.device_context()
    w, h = size
    hdr = synth.make_hdr_image(r, w, h)
    view = _put_hdr(ctx, assets, hdr)
    try:
        assert_pixels_equal(view.read_pixels(), hdr, "write_pixels / read_pixels round trip")
        p = make_post_params(exposure=0.7, color_modulate=(1.0, 0.9, 1.2))
        got, avg = view.postprocess(p)
    finally:
        view.close()
    want, want_avg, _ = oracle_lib.postprocess(hdr, p)
    assert_pixels_equal(got, want, f"tonemap {size}")
    # average colour: exact integer sum on the GPU vs f32 row sums in the reference -> relative 1e-5
    np.testing.assert_allclose(avg, want_avg, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("scale_log2", [1, 2, 3])
@pytest.mark.parametrize("linear", [False, True])
@pytest.mark.parametrize("size", [(128, 96), (131, 75)])
def test_bloom_bit_exact(assets, oracle_lib, scale_log2, linear, size):
    """Bloom chain at even and odd sizes (leftover rows/columns, postprocessor.rs:383-461) for every buffer scale and
    both upsampling filters."""
    r = synth.rng(600 + scale_log2 + 10 * linear + size[0])
    ctx = assets.device_context()
    w, h = size
    hdr = synth.make_hdr_image(r, w, h)
    view = _put_hdr(ctx, assets, hdr)
    try:
        p = make_post_params(exposure=1.3, bloom_sigma=float(r.uniform(1.0, 9.0)), bloom_scale=0.25, bloom_buffer_scale_log2=scale_log2, linear_bloom_filter=linear)
        got, avg = view.postprocess(p)
    finally:
        view.close()
    want, want_avg, _ = oracle_lib.postprocess(hdr, p)
    assert_pixels_equal(got, want, f"bloom s{scale_log2} linear{linear} {size}")
    np.testing.assert_allclose(avg, want_avg, rtol=1e-5, atol=1e-3)
