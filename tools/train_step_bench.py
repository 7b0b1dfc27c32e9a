#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3] / SURVEY.md §8d config 4): 3072-ray batches drawn from 256
synthetic poses x 12 rays (per-ray skts, like configs/h36m/h36m_prot2.txt:34-35), forward + backward through
posegen_b200.RayCaster in train mode, one NCCL all-reduce of the gradients, Adam step, with the config's
training-time randomness (perturb = 1, raw_noise_std = 1, configs/surreal/surreal.txt; `--deterministic` for the
parity setting).  Auxiliary to bench.py; prints one JSON line on rank 0.

    python tools/train_step_bench.py [--steps 20] [--graph] [--deterministic]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posegen_b200 import dist as pdist, synthetic as syn                         # noqa: E402
from posegen_b200.raycaster import raycaster_from_checkpoint                     # noqa: E402
from posegen_b200.train import allreduce_gradients                               # noqa: E402


def make_batch(seed, n_poses=256, rays_per_pose=12, res=512):
    rng = np.random.RandomState(seed)
    rb, sk, cy = [], [], []
    for p in range(n_poses):
        frame = syn.synthetic_frame(1000 + (seed * n_poses + p) % 64, res, res)      # 64 distinct poses are enough
        b = syn.ray_batch(frame.rays_o, frame.rays_d)
        sel = rng.randint(0, b.shape[0], rays_per_pose)
        rb.append(b[sel]); sk.append(np.repeat(frame.pose.skts[None], rays_per_pose, 0)); cy.append(np.repeat(frame.pose.cyl[None], rays_per_pose, 0))
    return np.concatenate(rb).astype(np.float32), np.concatenate(sk).astype(np.float32), np.concatenate(cy).astype(np.float32)


def run(rank, world, local, steps=20, warmup=3, deterministic=False, graph=True, profile=None):
    """One leg: returns the JSON-able result dict (same on every rank).  The process group (world > 1) must already be
    initialised; CUDA device `local`."""
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ckpt = syn.synthetic_raycaster_state(4, alpha_gain=40.0)
    rc = raycaster_from_checkpoint(ckpt, device=dev, precision="bf16")
    rc.train()
    opt = torch.optim.Adam([p for p in rc.parameters() if p.requires_grad], lr=5e-4, fused=True, capturable=graph)
    rb, sk, cy = make_batch(rank)
    n = rb.shape[0]
    rbt, skt, cyt = (torch.as_tensor(x, device=dev) for x in (rb, sk, cy))
    tgt = torch.rand(n, 3, device=dev)
    perturb, noise = (0., 0.) if deterministic else (1., 1.)

    def loss_fn(ret, t):
        return ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()

    def step():
        opt.zero_grad(set_to_none=True)
        ret = rc(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=skt, cyls=cyt, bones=None, cams=None,
                 perturb=perturb, raw_noise_std=noise)
        loss = loss_fn(ret, tgt)
        loss.backward()
        allreduce_gradients(rc.parameters())
        opt.step()
        return loss

    if graph:
        from posegen_b200.train import GraphedTrainStep
        graphed = GraphedTrainStep(rc, opt, loss_fn, {"ray_batch": rbt, "skts": skt, "cyls": cyt, "target": tgt},
                                   perturb=perturb, raw_noise_std=noise)

        def step():                                       # noqa: F811  a new batch is copied into the static inputs every step
            return graphed(ray_batch=rbt, skts=skt, cyls=cyt, target=tgt)
    eng = rc.engine(dev) if not graph else None
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = pdist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    # the gradient all-reduce alone (the step's only collective), same bucket, timed back to back
    ar_ms = 0.0
    if world > 1:
        for _ in range(3):
            allreduce_gradients(rc.parameters())
        torch.cuda.synchronize()
        torch.distributed.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            allreduce_gradients(rc.parameters())
        a1.record()
        torch.cuda.synchronize()
        ar_ms = pdist.max_over_ranks(a0.elapsed_time(a1), dev) / 10
    flop_per_ray = 3 * 248_205_312                        # forward + dX + dW, algorithmic nn.Linear MACs x 2
    tflops_per_gpu = n * flop_per_ray / (ms * 1e-3) / 1e12
    line = {"metric": "train_steps_per_sec", "value": 1e3 / ms, "ms_per_step": ms, "n_gpus": world, "rays_per_step": n * world,
            "rays_per_sec": n * world / ms * 1e3, "final_loss": float(loss.detach()), "allreduce_ms": ar_ms,
            "allreduce_elems": int(sum(p.numel() for p in rc.parameters() if p.requires_grad)),
            "tflops_per_gpu": tflops_per_gpu, "flop_per_ray": flop_per_ray,
            "config": "3072 rays/GPU from 256 poses x 12 rays, coarse+fine, fwd+bwd+allreduce+Adam, "
                      + ("perturb=0, raw_noise_std=0" if deterministic else "perturb=1, raw_noise_std=1") + ", bf16 tensor-core forward" + (", CUDA graph" if graph else "")}
    if profile and rank == 0:
        from torch.profiler import profile as tprofile, ProfilerActivity
        with tprofile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        with open(profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
    del eng
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--deterministic", action="store_true", help="perturb = 0, raw_noise_std = 0")
    ap.add_argument("--graph", action="store_true", help="capture the step into a CUDA graph (posegen_b200.train.GraphedTrainStep)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of 3 steps to this file (after the timed run)")
    a = ap.parse_args()
    rank, world, local = pdist.env_rank_world()
    pdist.init_process_group("nccl" if world > 1 else None)
    line = run(rank, world, local, a.steps, a.warmup, a.deterministic, a.graph, a.profile)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
