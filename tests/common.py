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


def oracle_render_host_frame(oracle_lib, scene_module, sc, hf, width, height, color_bits, shadow_maps=None):
    """The oracle's picture of a host frame, nested portal / mirror views included (build_view_portals_textures,
    partial_renderer.rs:2008-2282): every visible portal's nested frame is rendered first (black at the depth limit), its overlay
    texture mixed in, and the frame's TEX_VIEW commands sample those pictures.  Returns (surfaces, framebuffer).
    `hf.cmds` may already carry device view indices in tex_ref (Renderer patches them): the mapping is rebuilt from hf.portals."""
    dtype = np.uint32 if color_bits == 32 else np.uint64
    view_textures = []
    cmds = hf.cmds.copy()
    for node in hf.portals:
        w, h = node.resolution
        if node.frame is None:
            tex = np.zeros((h, w), dtype)
        else:
            _, tex = oracle_render_host_frame(oracle_lib, scene_module, sc, node.frame, w, h, color_bits)
        if node.overlay is not None:
            image, blend, light = node.overlay
            oracle_lib.mix_surface_with_texture(color_bits, tex, (w, h), node.tc_min, sc.images[image][1], blend, light)
        ref = node.view.index if node.view is not None else node.index
        while len(view_textures) <= ref:
            view_textures.append(None)
        view_textures[ref] = tex
    lights = hf.lights if hf.lights is not None and len(hf.lights) else None
    sm = shadow_maps
    if lights is not None and sm is None:
        sm = np.zeros(max(hf.shadow_maps_len, 1), np.float32)
        if hf.depth_maps is not None and len(hf.depth_maps):
            oracle_lib.draw_shadow_maps(sm, hf.depth_maps, hf.depth_cmds, hf.depth_points)
    surf = oracle_lib.build_surfaces(color_bits, hf.jobs, lights, sm, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
    fb = np.zeros((height, width), dtype)
    if hf.needs_clear:
        oracle_lib.fill(fb, scene_module.BACKGROUND_COLOR64 if color_bits == 64 else scene_module.BACKGROUND_COLOR32)
    oracle_lib.draw(color_bits, fb, cmds, hf.polygon_points, hf.triangle_points, surf, sc.images, view_textures=view_textures)
    return surf, fb


def wall_portals(scene_module, N, pose, overlay_skin=0, texels_per_unit=1.5):
    """FrameInfo::portals fixture: three 60 x 50 unit quads on a vertical plane 90 units ahead of the camera, facing it — a
    Mirror with an Average-blended overlay texture (a 'glass' material), a CameraAtPosition view and a ParallaxPortal that looks
    into the space 140 units further ahead.  The tex-coord bases have the orientation the game gives them
    (world_spawn.rs:372-381 / :407-415): mirrors (u x v) . n >= 0, parallax portals <= 0."""
    az = float(pose[3])
    fwd = np.array([-np.sin(az), np.cos(az), 0.0])
    right = np.array([np.cos(az), np.sin(az), 0.0])
    up = np.array([0.0, 0.0, 1.0])
    p = np.zeros(3, scene_module.portal_dtype)
    kinds = (scene_module.PORTAL_VIEW_MIRROR, scene_module.PORTAL_VIEW_CAMERA_AT_POSITION, scene_module.PORTAL_VIEW_PARALLAX_PORTAL)
    for i, (off, kind) in enumerate(zip((-62.0, 0.0, 62.0), kinds)):
        normal = -fwd
        center = pose[:3] + fwd * 90.0 + right * off + up * (4.0 * i)
        r = np.cross(up, normal)
        r /= np.linalg.norm(r)
        half_w, half_h = 30.0, 25.0
        p[i]["view"] = kind
        p[i]["plane"][:3] = normal
        p[i]["plane"][3] = float(np.dot(normal, center))
        tu, tv = r * texels_per_unit, -up * texels_per_unit
        p[i]["tex_coord_equation"][0][:3] = tu
        p[i]["tex_coord_equation"][0][3] = -float(np.dot(tu, center)) + half_w * texels_per_unit
        p[i]["tex_coord_equation"][1][:3] = tv
        p[i]["tex_coord_equation"][1][3] = -float(np.dot(tv, center)) + half_h * texels_per_unit
        if kind == scene_module.PORTAL_VIEW_MIRROR:
            p[i]["tex_coord_equation"][0] *= -1.0
        p[i]["num_vertices"] = 4
        p[i]["vertices"][:4] = [center - r * half_w - up * half_h, center - r * half_w + up * half_h, center + r * half_w + up * half_h, center + r * half_w - up * half_h]
        p[i]["blending_mode"] = N.BLEND_NONE if i != 2 else N.BLEND_AVERAGE
        p[i]["camera_position"] = pose[:3] + fwd * 20.0
        a = az + np.pi / 2 + 0.5
        p[i]["camera_rotation"] = (np.cos(a / 2), 0.0, 0.0, np.sin(a / 2))
        p[i]["camera_fov"] = 1.2
        shift = np.eye(4, dtype=np.float32)
        shift[:3, 3] = fwd * 140.0
        p[i]["transform_matrix"] = shift.T.reshape(16)  # column-major
        p[i]["has_texture"] = 1 if i == 0 else 0
        p[i]["texture_skin"] = overlay_skin
        p[i]["texture_blending_mode"] = N.BLEND_AVERAGE
        p[i]["texture_light_add"] = (0.8, 0.7, 0.9)
    return p
