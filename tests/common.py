"""Shared helpers of the parity tests: one seeded asset set usable by both the oracle and the CUDA library."""
from __future__ import annotations

import numpy as np

import synth
from square_wheel_b200 import _native as N


class Assets:
    """Textures (every has_normal_map x roughness x metal combination, several sizes), lite textures, sky images, lightmaps."""

    def __init__(self, seed=1, num_lightmap_elements=20000, tex_sizes=(64, 32, 16)):
        r = synth.rng(seed)
        self.images = []  # (kind, ndarray) in image-id order
        self.texture_ids, self.texture_flags = [], []
        for size in tex_sizes:
            for nm in (False, True):
                for rough, metal in ((False, False), (True, False), (True, True)):
                    flags = (N.SURF_HAS_NORMAL_MAP if nm else 0) | (N.SURF_HAS_NON_ONE_ROUGHNESS if rough else 0) | (N.SURF_IS_METAL if metal else 0)
                    self.texture_ids.append(len(self.images))
                    self.texture_flags.append(flags)
                    self.images.append((N.IMAGE_TEXTURE_ELEMENT, synth.make_texture(r, size, nm, rough)))
        self.lite_ids = []
        for (w, h) in ((32, 32), (64, 16), (8, 8)):
            self.lite_ids.append(len(self.images))
            self.images.append((N.IMAGE_COLOR32, synth.make_lite_texture(r, w, h)))
        self.sky64_ids = []
        for (w, h) in ((64, 64), (32, 32)):
            self.sky64_ids.append(len(self.images))
            self.images.append((N.IMAGE_COLOR64, synth.make_color64_texture(r, w, h)))
        self.lightmaps, self.directional_lightmaps = synth.make_lightmaps(r, num_lightmap_elements)
        self.ctx = None

    def image_size(self, image_id):
        px = self.images[image_id][1]
        return px.shape[1], px.shape[0]

    def device_context(self, stream=0):
        from square_wheel_b200.device import Context

        if self.ctx is None:
            self.ctx = Context(self.lightmaps, self.directional_lightmaps, device=0, stream=stream)
            for i, (kind, px) in enumerate(self.images):
                assert self.ctx.upload_image(kind, px) == i
        return self.ctx


def channels32(a):
    a = a.astype(np.uint32)
    return np.stack([(a >> s) & 0xFF for s in (0, 8, 16, 24)], -1).astype(np.int64)


def channels64(a):
    a = a.astype(np.uint64)
    return np.stack([(a >> np.uint64(s)) & np.uint64(0xFFFF) for s in (0, 16, 32, 48)], -1).astype(np.int64)


def assert_pixels_equal(got, want, what=""):
    """Bit-exact comparison with a useful message (the bar for integer work; the float stages are bit-exact too
    because the CUDA path uses the portable reference numerics, which is stricter than north_star's +-1 LSB / 1e-3)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = np.flatnonzero(got.reshape(-1) != want.reshape(-1))
    if bad.size:
        i = int(bad[0])
        raise AssertionError(f"{what}: {bad.size} of {got.size} elements differ; first at {i}: got {int(got.reshape(-1)[i]):#x} want {int(want.reshape(-1)[i]):#x}")
