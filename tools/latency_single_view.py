#!/usr/bin/env python
"""Frame latency of ONE view on a single-stream context (the interactive case: Renderer.prepare_frame / draw_frame one frame at a
time), 4K HDR workload of bench.py.  Set SQW_NO_SETUP_FORK=1 to run the span set-up after surface building instead of beside it."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S
from square_wheel_b200.device import Context, make_post_params

sc = S.Scene(**bench.SCENE_KW)
poses = sc.camera_path(64)
frames, _ = bench.prepare_host_frames(sc, poses)
stream = torch.cuda.Stream(device=0)
torch.cuda.set_stream(stream)
ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0, stream=stream.cuda_stream, concurrent_views=False)
sc.upload(ctx)
view = ctx.create_view(bench.WIDTH, bench.HEIGHT, 64)
post = make_post_params(exposure=1.0)
rec = []
ms = max(hf.total_surface_pixels for hf in frames)
for hf in frames:
    f = view.create_frame()
    f.set_surfaces(hf.jobs, None, None, hf.cam_pos, ms)
    f.set_draw(hf.cmds, hf.polygon_points, None, clear_color=S.BACKGROUND_COLOR64 if hf.needs_clear else None)
    f.set_post(post)
    rec.append(f)
view.synchronize()
for _ in range(3):
    for f in rec:
        f.submit(N.STAGE_ALL)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(4):
    for f in rec:
        f.submit(N.STAGE_ALL)
e1.record(stream)
torch.cuda.synchronize()
print(json.dumps({"fork": os.environ.get("SQW_NO_SETUP_FORK") is None, "us_per_frame": e0.elapsed_time(e1) * 1e3 / (4 * len(rec))}))
