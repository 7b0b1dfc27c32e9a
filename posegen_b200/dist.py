"""Sharding of render jobs across ranks.

Every ray is independent given (weights, pose, camera), so the render path shards by
image/pose with no data-path collective: rank r renders jobs r, r+W, r+2W, ... with the
weights replicated.  The only exchange is an optional final gather of finished frames.
One process per GPU (torchrun); `gloo` on CPU for tests, `nccl` on the GPU box.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n_jobs: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: job i belongs to rank i % world (config 3: pose_idx % world_size)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_jobs, world))


def shard_counts(n_jobs: int, world: int) -> List[int]:
    return [len(range(r, n_jobs, world)) for r in range(world)]


def gather_frames(local_frames: torch.Tensor, n_jobs: int, rank: int, world: int) -> torch.Tensor:
    """Final gather: local_frames [n_local, ...] (round-robin shard) -> [n_jobs, ...] on every rank,
    restored to job order.  Ranks with one frame fewer pad with zeros for the collective."""
    if world == 1:
        return local_frames
    per = (n_jobs + world - 1) // world
    pad = per - local_frames.shape[0]
    if pad:
        local_frames = torch.cat([local_frames, local_frames.new_zeros((pad,) + tuple(local_frames.shape[1:]))], 0)
    out = local_frames.new_empty((world * per,) + tuple(local_frames.shape[1:]))
    dist.all_gather_into_tensor(out, local_frames.contiguous())
    out = out.view(world, per, *local_frames.shape[1:])
    # job i = rank (i % world), slot (i // world)
    idx = torch.arange(n_jobs, device=out.device)
    return out[idx % world, idx // world]


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
