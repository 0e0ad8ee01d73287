#!/usr/bin/env python
"""Device-resident throughput and per-kernel times of the other BASELINE.json configs (bench.py measures configs[1]).

    python tools/bench_configs.py [--frames 32] [--out profiles/rNN_configs.json]

configs[0]  1280x720, simple lightmaps, 32-bit LDR (no post stage)
configs[2]  synthetic stress: finely split map (tens of thousands of polygons in the map), mip bias -1, 16 dynamic lights
            (12 plain, 2 cube, 2 projector shadow maps rendered on the GPU), 16 triangle models of ~2.9k triangles, 4K HDR
configs[3]  sky box + 32 decals + 32 sprites + blended models, 4K HDR, tonemapping with the default bloom
configs[4]  fly-through views at 1920x1080 with the configs[1] settings (the view-batch unit; bench.py --gpus N shards it)

Same method as bench.py: recorded frames (inputs resident in HBM), 4 views in flight, CUDA events on the launching stream for
the whole run, then the same frames again one at a time with sqw_profile events around every launch.  One GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def lights_for(scene_mod, N, path, k, n):
    r = np.random.default_rng(1000 + k)
    lights = np.zeros(n, scene_mod.dynamic_light_dtype)
    for i in range(n):
        lights[i]["position"] = path[(k + i * 3) % len(path)][:3] + r.uniform(-40.0, 40.0, 3)
        lights[i]["radius"] = (128.0, 256.0, 512.0)[i % 3]
        lights[i]["color"] = r.uniform(4000.0, 60000.0, 3)
        lights[i]["shadow_type"] = N.SHADOW_NONE if i < n - 4 else (N.SHADOW_CUBE if i < n - 2 else N.SHADOW_PROJECTOR)
        a = r.uniform(0, 2 * np.pi)
        lights[i]["rotation"] = (np.cos(a / 2), 0.0, 0.0, np.sin(a / 2))
        lights[i]["fov"] = 0.2 * np.pi
    return lights


def models_for(scene_mod, sc, pose, k, n):
    m = np.zeros(n, scene_mod.model_entity_dtype)
    az = float(pose[3])
    num_models = int(sc.info.num_models)
    for i in range(n):
        a = az + (i - 0.5 * (n - 1)) * 0.11
        m[i]["position"] = pose[:3] + (60.0 + 17.0 * (i % 8)) * np.array([-np.sin(a), np.cos(a), 0.0]) + np.array([0.0, 0.0, -10.0 + 6.0 * (i % 4)])
        m[i]["rotation"] = (np.cos(0.3 * i), 0.0, 0.0, np.sin(0.3 * i))
        mi = i % num_models
        nf = sc.model_info(mi).num_frames
        m[i]["frames"] = (k % nf, (k + 1) % nf)
        m[i]["lerp"] = 0.4
        m[i]["model"] = mi
        m[i]["skin"] = mi
        m[i]["blending_mode"] = (0, 0, 3, 4)[i % 4]
        m[i]["lighting"] = (scene_mod.MODEL_LIGHTING_CONSTANT, scene_mod.MODEL_LIGHTING_ADVANCED)[i % 2]
        m[i]["light"] = (0.5, 0.45, 0.4)
        m[i]["grid_light_scale"] = 0.5
    return m


def portals_for(scene_mod, N, pose, num_skins):
    """BASELINE configs[3]: 2 mirrors + 1 camera portal + 1 parallax portal on a wall ahead of the camera (the fixture of the parity
    tests plus a second mirror above the first)."""
    from common import wall_portals

    p3 = wall_portals(scene_mod, N, pose, overlay_skin=0)
    p = np.zeros(4, scene_mod.portal_dtype)
    p[:3] = p3
    p[3] = p3[0]
    up = np.array([0.0, 0.0, 56.0], np.float32)
    p[3]["vertices"][:4] += up
    p[3]["plane"][3] += float(np.dot(p3[0]["plane"][:3], up))
    p[3]["tex_coord_equation"][1][3] -= float(np.dot(p3[0]["tex_coord_equation"][1][:3], up))
    p[3]["has_texture"] = 0
    return p


class FrameTree:
    """A recorded frame with its nested portal / mirror frames (build_view_portals_textures, partial_renderer.rs:2008-2282): every
    visible portal owns a render target; submit() replays nested frames first (depth-first), mixes the glass overlays in and then
    submits the frame that samples them.  Nothing synchronises with the host."""

    def __init__(self, ctx, view, hf, color_bits, clear_color, post, S, N):
        self.view, self.N = view, N
        self.children = []  # (FrameTree of the nested view, overlay or None, tc_min)
        self.owned_views = []
        for node in hf.portals:
            w, h = node.resolution
            v = ctx.create_view(w, h, color_bits)
            self.owned_views.append(v)
            if node.frame is None:
                child = None  # past the depth limit: black
            else:
                cc = (S.BACKGROUND_COLOR64 if color_bits == 64 else S.BACKGROUND_COLOR32) if node.frame.needs_clear else 0
                child = FrameTree(ctx, v, node.frame, color_bits, cc, None, S, N)
            self.children.append((child, v, node.overlay, node.tc_min))
            sel = (hf.cmds["tex_kind"] == N.TEX_VIEW) & (hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["tex_ref"] == node.index)
            hf.cmds["tex_ref"][sel] = v.index
        self.frame = view.create_frame()
        hf.record(self.frame, clear_color=clear_color, post=post)
        self.empty = np.zeros(0, N.draw_cmd_dtype)

    def submit(self, stages):
        N = self.N
        for child, v, overlay, tc_min in self.children:
            if child is None:
                v.draw(self.empty, None, None, clear_color=0)
            else:
                child.submit(N.STAGE_SHADOW_MAPS | N.STAGE_SURFACES | N.STAGE_RASTER)
            if overlay is not None:
                image, blend, light = overlay
                v.mix_texture(image, tc_min, blend, light)
        self.frame.submit(stages)

    def stats(self):
        return self.frame.stats()

    def num_nested(self):
        return sum(1 + (c.num_nested() if c is not None else 0) for c, _, _, _ in self.children)

    def close(self):
        for child, v, _, _ in self.children:
            if child is not None:
                child.close()
        self.frame.close()
        for v in self.owned_views:
            v.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--repeats", type=int, default=4)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring of the config name to run alone (for ncu captures)")
    args = ap.parse_args()

    import torch

    from common import scatter_decals, sprites_ahead_of_camera
    from square_wheel_b200 import _native as N
    from square_wheel_b200 import scene as S
    from square_wheel_b200.device import Context, make_post_params

    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(stream)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    base_scene = dict(seed=2, grid=(24, 24, 5), cell_size=(160.0, 160.0, 128.0), num_rooms=26, num_materials=24, texture_size_log2=8, max_polygon_size=64.0, material_mode=0, emissive_fraction=0.15)
    configs = {
        # the reference's own demo map 2 (assets/_ref/demo2.sqwm, compiled by __graft_entry__.build(); skipped where it is absent)
        "configs[0] demo map 2, 1280x720 simple lightmaps LDR": dict(scene="demo2", size=(1280, 720), bits=32, directional=False, post=None),
        "configs[1] variant B: demo map 2, 4K HDR + tonemap with the default bloom": dict(scene="demo2", size=(3840, 2160), bits=64, directional=True, post=dict(bloom_sigma=1.0)),
        "configs[0] 1280x720 simple lightmaps LDR": dict(scene=base_scene, size=(1280, 720), bits=32, directional=False, post=None),
        # ~20k visible polygons per frame on average (7.5-unit polygons: 0.75 M polygons in the map), mip bias -1 => mip 0
        "configs[2] stress 4K HDR: ~20k visible polygons, mip bias -1, 16 lights, 16 models": dict(
            scene=dict(base_scene, max_polygon_size=7.5, num_models=6, model_detail=27), size=(3840, 2160), bits=64, directional=True, post=dict(), mip_bias=-1.0, lights=16, models=16
        ),
        # translucent / alpha-tested / average / additive WORLD materials (material_mode 2), 32 decals, 32 sprites, sky box (6 x 512^2),
        # 2 mirrors + camera portal + parallax portal at portals_depth 2, blended models, 4K HDR (variant with the default bloom below)
        "configs[3] translucent + alpha-tested world, decals, sprites, sky, portals + mirrors (depth 2), 4K HDR + tonemap": dict(
            scene=dict(base_scene, material_mode=2, skybox=1, sky_size_log2=9, num_models=6, model_detail=12), size=(3840, 2160), bits=64, directional=True, post=dict(), models=6, sprites=32, decals=32,
            look_up=0.35, portals=True, portals_depth=2,
        ),
        "configs[3] same without portals, with the default bloom": dict(
            scene=dict(base_scene, material_mode=2, skybox=1, sky_size_log2=9, num_models=6, model_detail=12), size=(3840, 2160), bits=64, directional=True, post=dict(bloom_sigma=1.0), models=6, sprites=32, decals=32,
            look_up=0.35,
        ),
        "configs[4] view-batch unit: 1920x1080 HDR + tonemap": dict(scene=base_scene, size=(1920, 1080), bits=64, directional=True, post=dict()),
        # the stress map's static world alone (no lights, no models): what the device-side front-end chain is for
        "front-end chain: stress map, static world only, 4K HDR": dict(scene=dict(base_scene, max_polygon_size=7.5), size=(3840, 2160), bits=64, directional=True, post=dict(), mip_bias=-1.0, frames=8),
    }
    results = {}
    for name, cfg in configs.items():
        if args.only and args.only not in name:
            continue
        W, H = cfg["size"]
        C = cfg["bits"] // 8
        sc = S.demo2_scene(texture_size_log2=8) if cfg["scene"] == "demo2" else S.Scene(**cfg["scene"])
        if sc is None:
            continue
        hr = S.HostRenderer(sc, use_directional_lightmaps=cfg["directional"], textures_mip_bias=cfg.get("mip_bias", 0.0), portals_depth=cfg.get("portals_depth", 1))
        path = sc.camera_path(64)
        if cfg["scene"] == "demo2":  # inline models as the reference's game shows them (bench.prepare_host_frames)
            hr.build_frame(64, 64, path[0], submodels=hr.game_submodel_entities())
        frames = []
        geometry_frames = []
        draw_orders = []
        build_s = []
        num_frames = min(args.frames, cfg.get("frames", args.frames))
        for i in range(num_frames):
            k = (i * 64) // num_frames
            pose = path[k].copy()
            pose[4] += cfg.get("look_up", 0.0)
            kw = {}
            if cfg.get("lights"):
                kw["lights"] = lights_for(S, N, path, k, cfg["lights"])
            if cfg.get("models"):
                kw["models"] = models_for(S, sc, pose, k, cfg["models"])
            if cfg.get("sprites"):
                kw["sprites"] = sprites_ahead_of_camera(S, pose, cfg["sprites"], num_skins=int(sc.info.num_models))
            if cfg.get("decals"):
                kw["decals"] = scatter_decals(S, pose, cfg["decals"], k, num_skins=int(sc.info.num_models))
            if cfg.get("portals"):
                kw["portals"] = portals_for(S, N, pose, int(sc.info.num_models))
            t_build = time.perf_counter()
            frames.append(hr.build_frame(W, H, pose, color_bits=cfg["bits"], **kw))
            build_s.append(time.perf_counter() - t_build)
            if i % 4 == 0:  # world-polygon preparation of this pose: the mirror's lists, results and its wall time
                geometry_frames.append(hr.prepared_polygons() + (hr.polygons_prepare_seconds(),))
                index_of = {int(p): ci for ci, p in enumerate(geometry_frames[-1][0]["polygon"])}
                order = hr.polygon_draw_order()
                draw_orders.append(np.array([index_of[int(p)] for p in order], np.uint32) if all(int(p) in index_of for p in order) else None)
        hr.close()
        ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0, stream=stream.cuda_stream, concurrent_views=True)
        sc.upload(ctx)
        views = [ctx.create_view(W, H, cfg["bits"]) for _ in range(4)]
        post = make_post_params(exposure=1.0, **cfg["post"]) if cfg["post"] is not None else None
        recorded = []
        for i, hf in enumerate(frames):
            clear = None
            if hf.needs_clear:
                clear = S.BACKGROUND_COLOR64 if cfg["bits"] == 64 else S.BACKGROUND_COLOR32
            recorded.append(FrameTree(ctx, views[i % 4], hf, cfg["bits"], clear, post, S, N))
        for v in views:
            v.synchronize()

        def run(serial=False):
            ctx.batch_begin()
            for f in recorded:
                f.submit(N.STAGE_ALL)
                if serial:
                    torch.cuda.synchronize()
            ctx.batch_end()

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.repeats):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        n = args.repeats * len(recorded)
        ctx.profile_begin()
        run(serial=True)
        kt = ctx.profile_end()
        # the geometry half of the host front end on the GPU (sqw_polygons_prepare) beside the host mirror, same candidates
        ctx.upload_map_geometry(*sc.map_geometry())
        gv = views[0]
        for cand, planes, want, fr, _ in geometry_frames[:2]:
            gv.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
        torch.cuda.synchronize()
        ctx.profile_begin()
        t0 = time.perf_counter()
        for cand, planes, want, fr, _ in geometry_frames:
            gv.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
        call_s = time.perf_counter() - t0
        gkt = ctx.profile_end()
        # the whole chain for frames of static world geometry: prepare + draw + device-side job / band / frame set-up, against the
        # host-list path (sqw_frame_set_surfaces + sqw_frame_set_draw of the mirror's lists), wall clock per frame
        chain = None
        world_only = all(len(frames[i * 4].cmds) == int(((frames[i * 4].cmds["kind"] == N.CMD_POLYGON) & (frames[i * 4].cmds["tex_kind"] == N.TEX_SURFACE)).sum()) and (frames[i * 4].lights is None or len(frames[i * 4].lights) == 0) for i in range(len(geometry_frames)))
        if world_only and draw_orders and all(o is not None for o in draw_orders):
            hr2 = S.HostRenderer(sc, use_directional_lightmaps=cfg["directional"], textures_mip_bias=cfg.get("mip_bias", 0.0))
            ctx.upload_map_materials(sc.map_materials(), hr2.polygon_basis())
            hr2.close()
            fr_dev, fr_host = gv.create_frame(), gv.create_frame()
            t_dev = t_host = 0.0
            for rep in range(3):
                for gi, (cand, planes, want, fr, _) in enumerate(geometry_frames):
                    hf = frames[gi * 4]
                    t0 = time.perf_counter()
                    gv.polygons_prepare(fr["planes_matrix"], fr["view_matrix"], float(fr["mip_bias"]), cand, planes)
                    gv.polygons_draw(draw_orders[gi])
                    fr_dev.set_surfaces_from_polygons(hf.cam_pos, use_directional_lightmaps=cfg["directional"])
                    fr_dev.set_draw_from_polygons(clear_color=None)
                    gv.synchronize()
                    t1 = time.perf_counter()
                    fr_host.set_surfaces(hf.jobs, None, None, hf.cam_pos, hf.total_surface_pixels)
                    fr_host.set_draw(hf.cmds, hf.polygon_points, None, clear_color=None)
                    gv.synchronize()
                    t2 = time.perf_counter()
                    if rep:
                        t_dev += t1 - t0
                        t_host += t2 - t1
            fr_dev.close()
            fr_host.close()
            chain = {
                "device_chain_us_per_frame": t_dev * 1e6 / (2 * len(geometry_frames)),
                "host_lists_us_per_frame": t_host * 1e6 / (2 * len(geometry_frames)),
                "note": "device chain = sqw_polygons_prepare + sqw_polygons_draw + sqw_frame_set_surfaces_from_polygons + sqw_frame_set_draw_from_polygons from the candidate and draw-order lists (wall clock, five small synchronising read-backs of totals); host lists = sqw_frame_set_surfaces + sqw_frame_set_draw of the mirror's job / command / vertex lists (the mirror's own time to BUILD those lists is host_mirror_us_per_frame for the geometry alone, 0.3 ms per 4K frame for the whole front end)",
            }
        got, totals = gv.polygons_read(len(geometry_frames[-1][0]))
        want = geometry_frames[-1][2]
        geometry = {
            "candidates_per_frame": float(np.mean([len(g[0]) for g in geometry_frames])),
            "visible_polygons_per_frame": float(np.mean([int(g[3]["num_visible"]) for g in geometry_frames])),
            "host_mirror_us_per_frame": 1e6 * float(np.mean([g[4] for g in geometry_frames])),
            "host_mirror_whole_front_end_us_per_frame": 1e6 * float(np.mean(build_s)),
            "gpu_kernels_us_per_frame": gkt["polygons_prepare"][0] * 1e3 / len(geometry_frames),
            "gpu_call_us_per_frame": call_s * 1e6 / len(geometry_frames),
            "bit_exact_vs_host_mirror": bool(np.array_equal(got.view(np.uint8), want.view(np.uint8))),
            "chain": chain,
            "note": "prepare_polygons_surfaces for the leaf polygons (partial_renderer.rs:1327-1360): host mirror = one thread, wall clock; GPU kernels = polygons_prepare + "
                    "offsets by CUDA events; GPU call = sqw_polygons_prepare wall clock including the upload of the candidate list from pageable memory and the stream synchronisation",
        }
        st = [f.stats() for f in recorded]
        S_tex = sum(s.surface_texels for s in st) / len(st)
        Se = sum(s.emissive_texels for s in st) / len(st)
        L = sum(s.lightmap_bytes for s in st) / len(st)
        P = W * H
        b1 = S_tex * (8 + C) + Se * (4 + 2 * C) + L
        b2 = P * C + S_tex * C + (4 * P if post is not None and not cfg["post"].get("bloom_sigma") else 0)  # P C + S C (+ the picture when tonemapping is fused)
        kernels = {k: {"us_per_frame": v[0] * 1e3 / len(recorded), "launches_per_frame": v[1] / len(recorded)} for k, v in kt.items() if v[1]}
        for kname, b in (("surface_build", b1), ("raster_tile", b2), ("raster_tile_tonemap", b2)):
            if kname in kernels:
                gbs = b / (kernels[kname]["us_per_frame"] * 1e-6) / 1e9
                kernels[kname].update(algorithmic_bytes_per_frame=b, achieved_gbs=gbs, frac_of_measured_peak=gbs / peak)
        res = {
            "resolution": [W, H], "color_bits": cfg["bits"], "frames_per_sec": n / (ms * 1e-3), "us_per_frame": ms * 1e3 / n, "frames": len(recorded),
            "surface_texels_per_frame": S_tex, "surface_texels_per_sec": S_tex * n / (ms * 1e-3),
            "draw_commands_per_frame": float(np.mean([len(hf.cmds) for hf in frames])),
            "triangle_commands_per_frame": float(np.mean([(hf.cmds["kind"] == N.CMD_TRIANGLE).sum() for hf in frames])),
            "surface_jobs_per_frame": float(np.mean([len(hf.jobs) for hf in frames])),
            "spans_per_frame": float(np.mean([s.num_spans for s in st])),
            "map_polygons": int(sc.info.num_polygons),
            "nested_portal_views_per_frame": float(np.mean([f.num_nested() for f in recorded])),
            "blended_world_commands_per_frame": float(np.mean([((hf.cmds["kind"] == N.CMD_POLYGON) & (hf.cmds["blend"] != 0)).sum() for hf in frames])),
            "kernels": kernels,
            "front_end_geometry": geometry,
        }
        results[name] = res
        print(name, json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items() if k not in ("kernels", "front_end_geometry")}), flush=True)
        print("    front-end geometry:", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in geometry.items() if k != "note"}, flush=True)
        for kname, kv in kernels.items():
            print("   ", kname, {a: round(b, 3) if isinstance(b, float) else b for a, b in kv.items()}, flush=True)
        for f in recorded:
            f.close()
        for v in views:
            v.close()
        ctx.close()
        sc.close()
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"hbm_peak_gbs": peak, "method": __doc__.split("Same method")[1].strip(), "configs": results}, fh, indent=1)


if __name__ == "__main__":
    main()
