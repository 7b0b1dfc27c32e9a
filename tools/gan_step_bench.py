#!/usr/bin/env python
"""GAN-step benchmark (BASELINE.json configs[4] / SURVEY.md §8d config 5): pose generator -> differentiable FK ->
512x512 differentiable render (grad w.r.t. the bone transforms) -> crop / resize / normalise -> random-init HMR ->
MPJPE loss -> backward to the generator, data-parallel over poses with one NCCL all-reduce of the generator gradients.

The pose generator and HMR are outside the render path (SURVEY.md §8 scope): they are plain torch stand-ins of the
reference modules' shapes (BAGenerator run_gan.py:818-897: noise 32 -> 256 -> 2 residual stages -> 24 x (axis, angle);
HMR run_gan.py:1255-1369: ResNet-50 trunk + 3-iteration regressor of 24 6-D rotations, zero mean parameters), random
init, eval-mode normalisation.  What is measured is posegen_b200.gan: the frozen NeRF's forward, the chunked
recompute + input-gradient backward, and the image hand-off.  Prints one JSON line on rank 0.

    python tools/gan_step_bench.py [--poses-per-gpu 2] [--steps 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gan_step_bench.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posegen_b200 import dist as pdist, fk, gan, synthetic as syn                # noqa: E402
from posegen_b200.raycaster import raycaster_from_checkpoint                     # noqa: E402
from posegen_b200.train import allreduce_gradients                               # noqa: E402


class PoseGeneratorStandIn(nn.Module):
    """BAGenerator's shape (run_gan.py:818-897): noise -> axis-angle per joint, root rotation scaled by 2 pi."""

    def __init__(self, noise=32, width=256, stages=2):
        super().__init__()
        self.noise = noise
        self.w1 = nn.Linear(noise, width)
        self.stages = nn.ModuleList([nn.Sequential(nn.Linear(width, width), nn.LeakyReLU(), nn.Linear(width, width), nn.LeakyReLU())
                                     for _ in range(stages)])
        self.w2 = nn.Linear(width, 24 * 4)

    def forward(self, n, device, generator=None):
        y = F.leaky_relu(self.w1(torch.randn(n, self.noise, device=device, generator=generator)))
        for st in self.stages:
            y = y + st(y)
        y = self.w2(y).view(n, 24, 4)
        axis = y[..., :3] / torch.linalg.norm(y[..., :3], dim=-1, keepdim=True)
        out = axis * y[..., 3:4]
        scale = torch.ones(24, 1, device=device)
        scale[0] = 6.28
        return out * scale


def rot6d_to_rotmat(x):
    """run_gan.py:1188-1205."""
    x = x.view(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-1)


class HmrStandIn(nn.Module):
    """ResNet-50 trunk + the iterative regressor of run_gan.py:1273-1358 (pose only), random init."""

    def __init__(self):
        super().__init__()
        import torchvision
        trunk = torchvision.models.resnet50(weights=None)
        trunk.fc = nn.Identity()
        self.trunk = trunk
        self.fc1 = nn.Linear(2048 + 144 + 13, 1024)
        self.fc2 = nn.Linear(1024, 1024)
        self.decpose = nn.Linear(1024, 144)
        nn.init.xavier_uniform_(self.decpose.weight, gain=0.01)
        self.register_buffer("init_pose", torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24)[None])

    def forward(self, x, n_iter=3):
        xf = self.trunk(x)
        pose = self.init_pose.expand(x.shape[0], -1)
        rest = torch.zeros(x.shape[0], 13, device=x.device)
        for _ in range(n_iter):
            h = self.fc2(self.fc1(torch.cat([xf, pose, rest], 1)))
            pose = pose + self.decpose(h)
        return rot6d_to_rotmat(pose).view(-1, 24, 3, 3)


def joints_from_rotmats(R, rest):
    """FK on rotation matrices (get_smpl_l2ws_torch(axis_to_matrix=False), skeleton_utils.py:379-463) -> joints [B,24,3]."""
    B = R.shape[0]
    out_R, out_t = [], []
    for i, p in enumerate(syn.SMPL_PARENTS):
        off = rest[i] if i == 0 else rest[i] - rest[p]
        if i == 0:
            out_R.append(R[:, 0]); out_t.append(off.expand(B, 3))
        else:
            out_R.append(out_R[p] @ R[:, i]); out_t.append(out_t[p] + (out_R[p] @ off[:, None]).squeeze(-1))
    return torch.stack(out_t, 1)


def run(rank, world, local, steps=3, warmup=1, poses_per_gpu=2, res=512, chunk=16384, profile=None):
    """One leg (process group already initialised for world > 1): returns the result dict on every rank."""
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(1234 + rank)
    rc = raycaster_from_checkpoint(syn.synthetic_raycaster_state(4, alpha_gain=40.0), device=dev, precision="bf16")
    rc.eval()
    for p in rc.parameters():
        p.requires_grad_(False)                                   # run_gan.py:159-160: the NeRF is frozen
    gen = PoseGeneratorStandIn().to(dev)
    hmr = HmrStandIn().to(dev).eval()
    for p in hmr.parameters():
        p.requires_grad_(False)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-4)
    rest_np = (syn.SMPL_REST_POSE * syn.BODY_SCALE).astype(np.float32)
    rest = torch.as_tensor(rest_np, device=dev)
    H = W = res
    focal = 1000.0 * res / 512
    c2w = syn.run_gan_c2w()
    crop = tuple(int(round(v * res / 512)) for v in (100, 100, 412, 412))
    eng = rc.engine(dev)
    gan.resize_operator(eng, crop[2] - crop[0], 224)              # one-time probe of the resize operator
    stats = {}

    def step():
        opt.zero_grad(set_to_none=True)
        bones = gen(poses_per_gpu, dev)
        frames, kps = gan.render_pose_images(rc, bones, rest_np, c2w, H, W, focal, chunk=chunk)
        x = torch.stack([gan.hmr_input(eng, f, crop=crop) for f in frames])
        pred = joints_from_rotmats(hmr(x), rest)
        tgt = kps.float()
        sel = [1, 2, 4, 5, 7, 8, 12, 15, 16, 17, 18, 19, 20, 21]                       # run_gan.py:2097-2098
        loss = torch.norm((pred - pred[:, :1])[:, sel] - (tgt - tgt[:, :1])[:, sel], dim=-1).mean()   # mpjpe, run_gan.py:1458-1464
        loss.backward()
        allreduce_gradients(gen.parameters())
        gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in gen.parameters()))
        nn.utils.clip_grad_norm_(gen.parameters(), max_norm=1)                         # run_gan.py:2106
        opt.step()
        stats["loss"], stats["gnorm"] = loss.detach(), gnorm.detach()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    eng.check_status()
    ms = pdist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    n_img = poses_per_gpu * world
    line = {"metric": "gan_steps_per_sec", "value": 1e3 / ms, "ms_per_step": ms, "n_gpus": world, "images_per_step": n_img,
            "images_per_sec": n_img / ms * 1e3, "ms_per_image_per_gpu": ms / poses_per_gpu, "loss": float(stats["loss"]),
            "generator_grad_norm": float(stats["gnorm"]), "gpu_launches_per_step_rank0": int((eng.launch_count - l0) // steps),
            "config": f"{poses_per_gpu} poses/GPU, {res}x{res} bbox rays, frozen A-NeRF (bf16 tcgen05) forward + "
                      f"masks-only dump, backward to skts over the rays the crop reads (chunk {chunk} rays), device FK forward/backward, "
                      "crop/resize 224, ResNet-50 HMR stand-in, MPJPE, Adam on the generator, one all-reduce of the generator gradients"}
    if profile and rank == 0:
        from torch.profiler import profile as tprofile, ProfilerActivity
        with tprofile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        with open(profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=50, max_name_column_width=70))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--poses-per-gpu", type=int, default=2)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of one step to this file")
    a = ap.parse_args()
    rank, world, local = pdist.env_rank_world()
    pdist.init_process_group("nccl" if world > 1 else None)
    line = run(rank, world, local, a.steps, a.warmup, a.poses_per_gpu, a.res, a.chunk, a.profile)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
