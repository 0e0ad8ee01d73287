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


# ---- FrameInfo fixtures for the host mirror (decals, sprites) ---------------------------------------------------
def decal_rotation_for_normal(axis):
    """Quaternion (s, x, y, z) turning +X onto -axis, so that the decal's -X axis — the direction its texture is projected
    AGAINST, compared with the polygon normal in draw_polygon_decals (partial_renderer.rs:2838-2846) — equals `axis`."""
    d = -np.asarray(axis, np.float64)
    x = np.array([1.0, 0.0, 0.0])
    c = float(np.dot(x, d))
    if c > 0.9999:
        return (1.0, 0.0, 0.0, 0.0)
    if c < -0.9999:
        return (0.0, 0.0, 0.0, 1.0)
    v = np.cross(x, d)
    s = np.sqrt((1.0 + c) * 2.0)
    return (s / 2.0, v[0] / s, v[1] / s, v[2] / s)


def scatter_decals(scene_module, pose, n, seed, num_skins=4):
    """n decal boxes around the camera, projected onto floors, ceilings and the four wall directions, all blend modes, lit by
    the polygon's lightmap (even indices) or by a constant (odd indices)."""
    r = np.random.default_rng(seed)
    dc = np.zeros(n, scene_module.decal_dtype)
    axes = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    for i in range(n):
        dc[i]["position"] = pose[:3] + r.uniform(-160.0, 160.0, 3) * np.array([1.0, 1.0, 0.3])
        dc[i]["rotation"] = decal_rotation_for_normal(axes[i % 6])
        dc[i]["scale"] = (200.0, r.uniform(15.0, 45.0), r.uniform(15.0, 45.0))
        dc[i]["skin"] = i % num_skins
        dc[i]["blending_mode"] = (0, 4, 3, 4, 1, 2)[i % 6]
        dc[i]["lightmap_light_scale"] = 1.0 if i % 2 == 0 else 0.0
        dc[i]["light_add"] = (0.0, 0.0, 0.0) if i % 2 == 0 else (0.9, 0.8, 0.7)
    return dc


def sprites_ahead_of_camera(scene_module, pose, n, num_skins=4, light_scale=0.0):
    """n billboards fanned out in front of the camera (forward = (-sin az, cos az, 0)): the four SpriteOrientation kinds, with
    and without in-plane rotation, all blend modes; near ones get tesselated (partial_renderer.rs:896-909)."""
    sp = np.zeros(n, scene_module.sprite_dtype)
    az = float(pose[3])
    for i in range(n):
        a = az + (i - 0.5 * (n - 1)) * 0.12
        sp[i]["position"] = pose[:3] + (18.0 + 11.0 * i) * np.array([-np.sin(a), np.cos(a), 0.0]) + np.array([0.0, 0.0, (i % 3 - 1) * 8.0])
        sp[i]["radius"] = 10.0 + 2.0 * i
        sp[i]["skin"] = i % num_skins
        sp[i]["orientation"] = i % 4
        sp[i]["angle"] = 0.3 * i if i % 2 else 0.0
        sp[i]["blending_mode"] = (0, 3, 4, 1, 2, 3, 4, 0)[i % 8]
        sp[i]["light_scale"] = light_scale
        sp[i]["light_add"] = (0.8, 0.7, 0.9)
    return sp
