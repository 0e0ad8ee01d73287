import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from square_wheel_b200 import _native as N
from square_wheel_b200.device import Context, make_post_params
ctx = Context(np.zeros((4,3),np.float32), None)
W,H=3840,2160
view = ctx.create_view(W,H,64)
f = view.create_frame()
f.set_draw(np.zeros(0,N.draw_cmd_dtype), None, None, clear_color=(32<<32)|(16<<16)|8)
f.set_post(make_post_params(exposure=1.0))
for it in range(3):
    ctx.profile_begin()
    for _ in range(20): f.submit(N.STAGE_RASTER|N.STAGE_POST); view.synchronize()
    kt = ctx.profile_end()
    print({k:(round(v[0]/max(v[1],1)*1000,1), v[1]) for k,v in kt.items() if v[1]})
# raster only (no tonemap)
f2 = view.create_frame()
f2.set_draw(np.zeros(0,N.draw_cmd_dtype), None, None, clear_color=(32<<32)|(16<<16)|8)
ctx.profile_begin()
for _ in range(20): f2.submit(N.STAGE_RASTER); view.synchronize()
kt = ctx.profile_end()
print('raster only', {k:(round(v[0]/max(v[1],1)*1000,1), v[1]) for k,v in kt.items() if v[1]})
# reference point: plain device fills of the same sizes (what the memory system gives a write-only stream)
import torch
for mb in (66.4, 99.5, 256.0):
    n = int(mb * 1e6 // 8)
    t = torch.empty(n, dtype=torch.int64, device="cuda")
    for _ in range(3): t.fill_(7)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): t.fill_(7)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    print(f"torch fill {mb} MB: {us:.1f} us = {mb*1e6/us/1e6:.0f} GB/s")
