#!/usr/bin/env python
"""Inputs for the REFERENCE golden-vector emitter (tools/rust_golden): small seeded cases of the three stages in one
little-endian binary file, tests/golden/reference_inputs.bin.  A maintainer with a Rust toolchain runs

    cd tools/rust_golden && cargo run --release -- ../../tests/golden/reference_inputs.bin ../../tests/golden/reference_outputs.bin

against the unmodified reference (path dependency on its crate) and commits reference_outputs.bin;
tests/test_reference_golden_cpu.py then pins the oracle against it bit for bit (in the numerics mode the emitter was built
with: it records cfg!(target_feature = "sse4.1") / "fma").  Until then the oracle stays "parity unpinned".

Layout (all little-endian; u = u32, i = i32, f = f32, q = u64):
  'SQWG' u:version=1
  u:n_surfaces, each: u color_bits, u directional, 12f basis (u, v, start, normal), 2u surface_size, 2i tc_min, 2u tex_size,
     u tex_flags (1 normal map, 2 non-one roughness, 4 metal), q[tex_w * tex_h] TextureElement raw, 2u lm_size, u lm_scale_log2,
     2u lm_tc_shift, u has_lightmap, f[lm_w * lm_h * (10 if directional else 3)], 3f cam_pos
  u:n_draws, each: u color_bits, 2u size, q clear, u n_cmds, each cmd: u kind, u blend, u tc_mode, u n_vertices,
     polygon: (2i) * n_vertices, 3f depth_eq, 6f tc_eq (d_tc_dx[2], d_tc_dy[2], k[2]), 2i tex_size, q[w * h] texels of the target's colour type
     triangle: (7i) * 3 (x, y, tc[2], light[3]), 2i tex_size, q[w * h] Color32 texels
  u:n_posts, each: 2u size, f exposure (always 1: Postprocessor::new starts there), 3f color_modulate (R, G, B), f bloom_sigma,
     f bloom_scale, u bloom_buffer_scale_log2, u linear_bloom_filter, q[w * h] Color64 pixels
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import synth  # noqa: E402
from common import Assets  # noqa: E402
from square_wheel_b200 import _native as N  # noqa: E402


def build_cases():
    """The cases as Python structures (also what tests/test_reference_golden_cpu.py feeds the oracle with)."""
    a = Assets(seed=77, num_lightmap_elements=4000, tex_sizes=(16,))
    r = synth.rng(7700)
    surfaces = []
    for color_bits in (32, 64):
        for directional in (False, True):
            jobs, total, need = synth.make_surface_jobs(r, 18, a.texture_ids, a.texture_flags, directional=directional, lights=0, max_size=24)
            assert need <= len(a.lightmaps)
            cam = r.uniform(-200, 200, 3).astype(np.float32)
            surfaces.append(dict(color_bits=color_bits, directional=directional, jobs=jobs, total=total, cam=cam))
    draws = []
    for color_bits in (32, 64):
        W, H = 96, 64
        ids = a.lite_ids if color_bits == 32 else a.sky64_ids
        tex = [ids[int(r.integers(0, len(ids)))] for _ in range(24)]
        pc, pp = synth.make_polygon_scene(r, W, H, 24, [a.image_size(t) for t in tex], blends=(0, 1, 2, 3, 4))
        pc["tex_kind"], pc["tex_ref"] = N.TEX_IMAGE, tex
        ttex = [a.lite_ids[int(r.integers(0, len(a.lite_ids)))] for _ in range(24)]
        tc, tp = synth.make_triangle_scene(r, W, H, 24, [a.image_size(t) for t in ttex], blends=(0, 1, 2, 3, 4), max_extent=50.0)
        tc["tex_ref"] = ttex
        cmds = np.concatenate([pc, tc])
        draws.append(dict(color_bits=color_bits, size=(W, H), clear=0x0011223300445566 if color_bits == 64 else 0x80112233, cmds=cmds, pp=pp, tp=tp))
    posts = []
    for (w, h, sigma, scale_log2, linear) in ((64, 48, 0.0, 1, 0), (65, 47, 0.0, 1, 0), (64, 48, 2.0, 1, 0), (66, 50, 3.0, 2, 1)):
        posts.append(dict(size=(w, h), modulate=(1.0, 0.9, 1.2), bloom_sigma=sigma, bloom_scale=0.25, scale_log2=scale_log2, linear=linear, hdr=synth.make_hdr_image(r, w, h)))
    return a, surfaces, draws, posts


def write(path):
    a, surfaces, draws, posts = build_cases()
    out = bytearray()
    out += b"SQWG" + struct.pack("<I", 1)
    n_surf = sum(len(s["jobs"]) for s in surfaces)
    out += struct.pack("<I", n_surf)
    for s in surfaces:
        for j in s["jobs"]:
            tex = a.images[int(j["texture_image"])][1]
            flags = int(j["flags"])
            tflags = (1 if flags & N.SURF_HAS_NORMAL_MAP else 0) | (2 if flags & N.SURF_HAS_NON_ONE_ROUGHNESS else 0) | (4 if flags & N.SURF_IS_METAL else 0)
            out += struct.pack("<II", s["color_bits"], 1 if s["directional"] else 0)
            out += np.concatenate([j["basis_u"], j["basis_v"], j["basis_start"], j["basis_normal"]]).astype("<f4").tobytes()
            out += struct.pack("<IIii", *[int(x) for x in j["surface_size"]], *[int(x) for x in j["tc_start"]])
            out += struct.pack("<III", tex.shape[1], tex.shape[0], tflags)
            out += np.ascontiguousarray(tex).view("<u8").tobytes()
            lw, lh = int(j["lightmap_size"][0]), int(j["lightmap_size"][1])
            has_lm = int(j["lightmap_offset"]) != 0
            out += struct.pack("<IIIIII", lw, lh, int(j["lightmap_scale_log2"]), int(j["lightmap_tc_shift"][0]), int(j["lightmap_tc_shift"][1]), 1 if has_lm else 0)
            off = int(j["lightmap_offset"])
            if s["directional"]:
                lm = a.directional_lightmaps[off : off + lw * lh].view("<f4") if has_lm else np.zeros(0, "<f4")
            else:
                lm = a.lightmaps.reshape(-1, 3)[off : off + lw * lh].astype("<f4").reshape(-1) if has_lm else np.zeros(0, "<f4")
            if not has_lm:
                lw = lh = 0
            out += np.ascontiguousarray(lm).tobytes()
            out += s["cam"].astype("<f4").tobytes()
    out += struct.pack("<I", len(draws))
    for d in draws:
        out += struct.pack("<IIIQI", d["color_bits"], d["size"][0], d["size"][1], d["clear"], len(d["cmds"]))
        for c in d["cmds"]:
            kind, nv, vo = int(c["kind"]), int(c["num_vertices"]), int(c["vertex_offset"])
            out += struct.pack("<IIII", kind, int(c["blend"]), int(c["tc_mode"]), nv)
            img = a.images[int(c["tex_ref"])][1]
            if kind == N.CMD_POLYGON:
                for p in d["pp"][vo : vo + nv]:
                    out += struct.pack("<ii", int(p["x"]), int(p["y"]))
                out += np.asarray(c["depth_eq"], "<f4").tobytes() + np.asarray(c["tc_eq"], "<f4").tobytes()
            else:
                for p in d["tp"][vo : vo + 3]:
                    out += struct.pack("<7i", int(p["x"]), int(p["y"]), int(p["tc"][0]), int(p["tc"][1]), *[int(x) for x in p["light"]])
            out += struct.pack("<ii", int(c["tex_size"][0]), int(c["tex_size"][1]))
            out += np.ascontiguousarray(img).astype("<u8").tobytes()
    out += struct.pack("<I", len(posts))
    for p in posts:
        out += struct.pack("<IIffffffII", p["size"][0], p["size"][1], 1.0, *p["modulate"], p["bloom_sigma"], p["bloom_scale"], p["scale_log2"], p["linear"])
        out += np.ascontiguousarray(p["hdr"]).astype("<u8").tobytes()
    with open(path, "wb") as f:
        f.write(out)
    return len(out)


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_inputs.bin")
    print(path, write(path), "bytes")
