"""Seeded synthetic inputs for the parity tests: textures, lightmaps, surface jobs, polygon / triangle commands.

Value distributions follow SURVEY.md §8(d): value-noise albedo, normals from a height field, roughness in
[1/126, 1], HDR lightmaps log-uniform 0.05..8 with a few texels up to 64, deviation in [0, 1].
"""
from __future__ import annotations

import numpy as np

from square_wheel_b200 import _native as N


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


# ---- textures ---------------------------------------------------------------------------------------------------
def pack_normal_roughness(normal, roughness):
    """PackedNormalRoughness::pack (textures.rs:34-40) via pack_f32x4_into_bytes (fast_math.rs:418)."""
    v = np.concatenate([normal, roughness[..., None]], axis=-1).astype(np.float32) * np.float32(127.0)
    b = np.clip(v, -127.0, 127.0).astype(np.int32) & 0xFF
    packed = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)
    return packed.astype(np.uint32).view(np.int32)


def make_texture(r, size=64, normal_map=True, rough=True):
    """TextureElement image [size, size] (textures.rs:20-27)."""
    h = w = size
    base = r.random((h // 8 + 1, w // 8 + 1, 3))
    albedo = np.kron(base, np.ones((8, 8, 1)))[:h, :w] * 0.7 + r.random((h, w, 3)) * 0.3
    rgb = (albedo * 255).astype(np.uint32)
    diffuse = (np.uint32(255) << 24) | (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2]
    if normal_map:
        height = r.random((h, w)).astype(np.float32)
        gx = np.roll(height, -1, 1) - height
        gy = np.roll(height, -1, 0) - height
        n = np.stack([-gx, -gy, np.full_like(gx, 0.6)], -1)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
    else:
        n = np.zeros((h, w, 3), np.float32)
        n[..., 2] = 1.0
    roughness = (r.random((h, w)) * (1.0 - 1.0 / 126.0) + 1.0 / 126.0).astype(np.float32) if rough else np.ones((h, w), np.float32)
    tex = np.zeros((h, w), N.texture_element_dtype)
    tex["diffuse"] = diffuse
    tex["packed_normal_roughness"] = pack_normal_roughness(n.astype(np.float32), roughness)
    return tex


def make_lite_texture(r, w=32, h=32, alpha=True):
    """Color32 image (TextureLite)."""
    px = r.integers(0, 256, (h, w, 4), dtype=np.uint32)
    if not alpha:
        px[..., 3] = 255
    return (px[..., 3] << 24) | (px[..., 0] << 16) | (px[..., 1] << 8) | px[..., 2]


def make_color64_texture(r, w=32, h=32, hdr_max=4096):
    px = r.integers(0, hdr_max, (h, w, 4), dtype=np.uint64)
    px[..., 3] = r.integers(0, 256, (h, w), dtype=np.uint64)
    return (px[..., 3] << np.uint64(48)) | (px[..., 0] << np.uint64(32)) | (px[..., 1] << np.uint64(16)) | px[..., 2]


# ---- lightmaps --------------------------------------------------------------------------------------------------
def make_lightmaps(r, n):
    """Returns (simple [n,3] f32, directional [n]) sharing offsets like the two arrays of BSPMap."""
    def hdr(shape):
        v = np.exp(r.uniform(np.log(0.05), np.log(8.0), shape))
        hot = r.random(shape) < 0.01
        return np.where(hot, v * 8.0, v).astype(np.float32)

    simple = hdr((n, 3))
    d = np.zeros(n, N.directional_lightmap_dtype)
    d["ambient_light"] = hdr((n, 3)) * 0.3
    vec = r.normal(size=(n, 3)).astype(np.float32)
    vec[:, 2] = np.abs(vec[:, 2]) + 0.2
    vec /= np.linalg.norm(vec, axis=1, keepdims=True)
    d["light_direction_vector_scaled"] = vec * r.uniform(0.05, 1.0, (n, 1)).astype(np.float32)
    d["directional_light_deviation"] = r.random(n).astype(np.float32)
    d["directional_light_color"] = hdr((n, 3))
    return simple, d


# ---- surface jobs -------------------------------------------------------------------------------------------------
def make_surface_jobs(r, n, texture_images, texture_flags, num_lightmap_elements_hint=None, emissive_images=(), directional=None, lights=0, max_size=None):
    """Random but valid jobs.  texture_images: image ids of TextureElement mips; texture_flags: SQW_SURF_* bits per image.
    Returns (jobs, total_pixels, lightmap_elements_needed)."""
    jobs = np.zeros(n, N.surface_job_dtype)
    offset = 0
    lm_cursor = 1  # offset 0 means "no lightmap" (bsp_map_compact.rs:73)
    for i in range(n):
        j = jobs[i]
        scale = int(r.integers(0, 5))
        lm_w, lm_h = int(r.integers(2, 19)), int(r.integers(2, 19))
        shift = r.integers(0, 1 << scale, 2)
        max_u = ((lm_w - 1) << scale) - int(shift[0])
        max_v = ((lm_h - 1) << scale) - int(shift[1])
        cap = max_size or 272
        su = int(r.integers(1, min(max_u, cap) + 1))
        sv = int(r.integers(1, min(max_v, cap) + 1))
        has_lm = r.random() > 0.1
        t = int(r.integers(0, len(texture_images)))
        flags = int(texture_flags[t]) | N.SURF_APPLY_DIFFUSE
        use_dir = (r.random() < 0.5) if directional is None else directional
        if use_dir:
            flags |= N.SURF_DIRECTIONAL_LIGHTMAP
        if len(emissive_images) and r.random() < 0.3:
            flags |= N.SURF_EMISSIVE
            j["emissive_image"] = emissive_images[int(r.integers(0, len(emissive_images)))]
            j["emissive_light"] = r.uniform(0.0, 6.0, 3)
            if r.random() < 0.3:
                flags &= ~N.SURF_APPLY_DIFFUSE
        normal = r.normal(size=3)
        normal /= np.linalg.norm(normal)
        a = np.cross(normal, [0.3, 0.5, 0.8])
        a /= np.linalg.norm(a)
        b = np.cross(normal, a)
        mip_scale = float(1 << (4 - scale))
        j["basis_u"] = a * 0.5 * mip_scale
        j["basis_v"] = b * 0.5 * mip_scale
        j["basis_start"] = r.uniform(-512, 512, 3)
        j["basis_normal"] = normal
        j["surface_size"] = (su, sv)
        j["tc_start"] = r.integers(-600, 600, 2)
        j["texture_image"] = texture_images[t]
        j["flags"] = flags
        j["lightmap_size"] = (lm_w, lm_h)
        j["lightmap_scale_log2"] = scale
        j["lightmap_tc_shift"] = shift
        if has_lm:
            j["lightmap_offset"] = lm_cursor
            lm_cursor += lm_w * lm_h
        if lights:
            nl = int(r.integers(0, 7))
            nl = min(nl, lights)
            j["num_lights"] = nl
            j["lights"][:nl] = r.choice(lights, nl, replace=False)
        j["surface_pixels_offset"] = offset
        offset += su * sv
    return jobs, offset, lm_cursor


def make_lights(r, n, center_spread=600.0):
    """DynamicLightWithShadow set (light.rs:4-11): none / cube / projector shadow maps with random depth content."""
    lights = np.zeros(n, N.light_dtype)
    maps = []
    cursor = 0
    for i in range(n):
        l = lights[i]
        l["position"] = r.uniform(-center_spread, center_spread, 3)
        radius = float(r.choice([128.0, 512.0, 2048.0]))
        l["radius"] = radius
        l["inv_square_radius"] = np.float32(1.0) / (np.float32(radius) * np.float32(radius))
        l["color"] = r.uniform(1024.0, 65536.0, 3)
        kind = i % 3
        l["shadow_type"] = kind
        if kind == N.SHADOW_CUBE:
            size = int(r.choice([16, 64]))
            l["shadow_map_size"] = size
            l["shadow_map_offset"] = cursor
            data = (1.0 / r.uniform(8.0, 900.0, 6 * size * size)).astype(np.float32)
            maps.append(data)
            cursor += data.size
        elif kind == N.SHADOW_PROJECTOR:
            size = int(r.choice([32, 128]))
            l["shadow_map_size"] = size
            l["shadow_map_offset"] = cursor
            data = (1.0 / r.uniform(8.0, 900.0, size * size)).astype(np.float32)
            maps.append(data)
            cursor += data.size
            z = r.normal(size=3)
            z /= np.linalg.norm(z)
            x = np.cross(z, [0.1, 0.9, 0.3])
            x /= np.linalg.norm(x)
            y = np.cross(z, x)
            l["basis_x"], l["basis_y"], l["basis_z"] = x * 2.0, y * 2.0, z
    sm = np.concatenate(maps) if maps else np.zeros(0, np.float32)
    return lights, sm


# ---- draw commands --------------------------------------------------------------------------------------------------
def plane_equations(p0, u_axis, v_axis, tc_scale, tc_offset, focal, cx, cy):
    """Screen-space DepthEquation / TexCoordEquation of a textured plane in camera space (x right, y down, z forward),
    projected with sx = focal*X/Z + cx.  Same algebra as equations.rs:22-75."""
    n = np.cross(u_axis, v_axis)
    d = float(np.dot(n, p0))
    depth = np.array([n[0] / focal, n[1] / focal, n[2] - n[0] * cx / focal - n[1] * cy / focal]) / d
    tc = np.zeros(6)
    for i, axis in enumerate((u_axis, v_axis)):
        a = axis * tc_scale
        base = np.array([a[0] / focal, a[1] / focal, a[2] - a[0] * cx / focal - a[1] * cy / focal])
        k = tc_offset[i] - float(np.dot(a, p0))
        e = base + k * depth
        tc[i], tc[2 + i], tc[4 + i] = e
    return depth.astype(np.float32), tc.astype(np.float32)


def make_polygon_scene(r, width, height, n, tex_sizes, modes=(0, 1, 2), blends=(0,), allow_offscreen=True):
    """n textured convex polygons.  Returns (cmds without tex refs filled, polygon_points). tex_sizes[i] = (w, h) of the texture
    the i-th command will use (caller fills tex_kind/tex_ref)."""
    cmds = np.zeros(n, N.draw_cmd_dtype)
    pts = []
    focal = width * 0.6
    cx, cy = width * 0.5, height * 0.5
    for i in range(n):
        tw, th = tex_sizes[i]
        for _ in range(100):
            # plane in camera space
            z0 = float(np.exp(r.uniform(np.log(40.0), np.log(3000.0))))
            p0 = np.array([r.uniform(-0.6, 0.6) * z0, r.uniform(-0.4, 0.4) * z0, z0])
            nrm = r.normal(size=3)
            nrm[2] = -abs(nrm[2]) - 0.15
            nrm /= np.linalg.norm(nrm)
            ua = np.cross(nrm, [0.0, 1.0, 0.1])
            ua /= np.linalg.norm(ua)
            va = np.cross(nrm, ua)
            tc_scale = float(np.exp(r.uniform(np.log(0.05), np.log(4.0))))  # texels per unit
            ext_u, ext_v = tw / tc_scale, th / tc_scale
            # convex polygon in (s, t) inside the texture rectangle
            k = int(r.integers(3, 9))
            ang = np.sort(r.uniform(0, 2 * np.pi, k))
            rad = r.uniform(0.35, 0.5)
            s = (0.5 + rad * np.cos(ang)) * ext_u
            t = (0.5 + rad * np.sin(ang)) * ext_v
            p = p0[None, :] + s[:, None] * ua[None, :] + t[:, None] * va[None, :]
            if np.any(p[:, 2] < 4.0):
                continue
            sx = focal * p[:, 0] / p[:, 2] + cx
            sy = focal * p[:, 1] / p[:, 2] + cy
            lim = 3.0 if allow_offscreen else 1.0
            if np.any(np.abs(sx - cx) > lim * width) or np.any(np.abs(sy - cy) > lim * height):
                continue
            area2 = np.sum(sx * np.roll(sy, -1) - np.roll(sx, -1) * sy)
            if abs(area2) < 8.0:
                continue
            if area2 < 0:  # make it clockwise on a y-down screen
                sx, sy = sx[::-1], sy[::-1]
            break
        else:
            raise RuntimeError("could not place polygon")
        depth, tc = plane_equations(p0, ua, va, tc_scale, (0.0, 0.0), focal, cx, cy)
        mode = int(modes[int(r.integers(0, len(modes)))])
        if mode != N.TC_AFFINE:
            # z_scale renormalisation of draw_polygon (partial_renderer.rs:4106-4109)
            inv_z = depth[0] * sx + depth[1] * sy + depth[2]
            max_inv_z = max(float(inv_z.max()), 1.0 / (1 << 20))
            z_scale = np.float32(2.0 ** (-5.0 - np.ceil(np.log2(max_inv_z))))
            depth = (depth * z_scale).astype(np.float32)
            tc = (tc * z_scale).astype(np.float32)
        c = cmds[i]
        c["kind"] = N.CMD_POLYGON
        c["blend"] = int(blends[int(r.integers(0, len(blends)))])
        c["tc_mode"] = mode
        c["num_vertices"] = len(sx)
        c["vertex_offset"] = len(pts)
        c["tex_size"] = (tw, th)
        c["depth_eq"] = depth
        c["tc_eq"] = tc
        for x, y in zip(sx, sy):
            pts.append((int(np.float32(x) * np.float32(65536.0)), int(np.float32(y) * np.float32(65536.0))))
    return cmds, np.array(pts, N.polygon_point_dtype)


def make_triangle_scene(r, width, height, n, tex_sizes, blends=(0,), max_extent=200.0):
    """n Gouraud-lit affine triangles (TrianglePointProjected, rasterizer.rs:1491-1498)."""
    cmds = np.zeros(n, N.draw_cmd_dtype)
    pts = np.zeros(n * 3, N.triangle_point_dtype)
    for i in range(n):
        tw, th = tex_sizes[i]
        c = np.array([r.uniform(-0.1, 1.1) * width, r.uniform(-0.1, 1.1) * height])
        ext = float(np.exp(r.uniform(np.log(0.3), np.log(max_extent))))
        v = c[None, :] + r.uniform(-ext, ext, (3, 2))
        if r.random() < 0.1:
            v[1, 1] = v[0, 1]  # flat edge: exercises the dy clamps (rasterizer.rs:797-799)
        for k in range(3):
            p = pts[i * 3 + k]
            p["x"] = int(v[k, 0] * 65536.0)
            p["y"] = int(v[k, 1] * 65536.0)
            # slightly out-of-range texture coordinates exercise the correction block (:1018-1138)
            p["tc"] = (int(r.uniform(-0.05, 1.05) * tw * 65536.0), int(r.uniform(-0.05, 1.05) * th * 65536.0))
            p["light"] = (r.uniform(0.0, 3.0, 3) * 65536.0).astype(np.int64)
        cm = cmds[i]
        cm["kind"] = N.CMD_TRIANGLE
        cm["blend"] = int(blends[int(r.integers(0, len(blends)))])
        cm["num_vertices"] = 3
        cm["vertex_offset"] = i * 3
        cm["tex_size"] = (tw, th)
        cm["tex_kind"] = N.TEX_IMAGE
    return cmds, pts


def make_hdr_image(r, width, height):
    """Color64 picture with HDR content: mostly 0..600, some texels up to 65535."""
    v = np.exp(r.uniform(np.log(1.0), np.log(600.0), (height, width, 3)))
    hot = r.random((height, width, 1)) < 0.01
    v = np.where(hot, v * 100.0, v).clip(0, 65535).astype(np.uint64)
    a = r.integers(0, 256, (height, width), dtype=np.uint64)
    return (a << np.uint64(48)) | (v[..., 0] << np.uint64(32)) | (v[..., 1] << np.uint64(16)) | v[..., 2]
