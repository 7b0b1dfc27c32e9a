"""Small fixed workload for ncu: renders `--iters` synthetic frames at `--res` with one engine.
    python tools/profile_render.py --res 512 --iters 2 --precision bf16
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posegen_b200 import synthetic as syn          # noqa: E402
from posegen_b200.engine import Engine             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
eng = Engine()
eng.load_checkpoint(syn.synthetic_raycaster_state(0, alpha_gain=400.))
dev = eng.device
for i in range(a.iters):
    f = syn.synthetic_frame(100 + i, a.res, a.res)
    rb = torch.as_tensor(syn.ray_batch(f.rays_o, f.rays_d), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.render(rb, torch.as_tensor(f.pose.skts, device=dev), torch.as_tensor(f.pose.cyl, device=dev),
                     nanfill_chunk=4096, precision=a.precision, return_alpha=False)
    e1.record()
    torch.cuda.synchronize()
    eng.check_status()
    print(f"iter {i}: {rb.shape[0]} rays {e0.elapsed_time(e1):.3f} ms acc_mean {float(out['acc_map'].mean()):.4f}")
