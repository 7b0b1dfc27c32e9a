#!/usr/bin/env python
"""CPU baseline of the training step (BASELINE.json configs[3]): forward + backward of the reference's training loss
through the ORACLE (autograd over oracle/render_oracle.py) on the host cores, on a bounded sample of the batch
tools/train_step_bench.py times on the GPU.  Lives under tests/ because the oracle is test infrastructure.

    python tests/train_step_cpu_baseline.py [--rays 768]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import render_oracle as orc                     # noqa: E402
from posegen_b200 import synthetic as syn                   # noqa: E402
from train_step_bench import make_batch                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=768)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    ckpt = syn.synthetic_raycaster_state(4, alpha_gain=40.0)
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    for net in nets:
        for v in net.values():
            v.requires_grad_(True)
    rb, sk, cy = make_batch(0)
    m = a.rays
    tgt = torch.rand(m, 3)
    t0 = time.perf_counter()
    r = orc.render_rays(torch.as_tensor(rb[:m]), torch.as_tensor(sk[:m]), torch.as_tensor(cy[:m]), nets, emb)
    loss = ((r["rgb_map"] + (1 - r["acc_map"][:, None]) - tgt) ** 2).mean() + ((r["rgb0"] + (1 - r["acc0"][:, None]) - tgt) ** 2).mean()
    loss.backward()
    sec = time.perf_counter() - t0
    print(json.dumps({"metric": "train_rays_per_sec_cpu", "value": m / sec, "cores": os.cpu_count(), "kind": "port",
                      "sample": f"{m} rays of the 3072-ray batch, fwd+bwd (autograd through the oracle), {sec:.1f} s"}))


if __name__ == "__main__":
    main()
