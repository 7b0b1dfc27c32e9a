"""CPU: the N>1 sharding / gather path with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from posegen_b200 import dist as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_jobs, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = pdist.init_process_group("gloo")
    mine = pdist.shard_indices(n_jobs, r, w)
    frames = torch.stack([torch.full((4, 4, 3), float(i)) for i in mine]) if mine else torch.zeros(0, 4, 4, 3)
    full = pdist.gather_frames(frames, n_jobs, r, w)
    expect = torch.stack([torch.full((4, 4, 3), float(i)) for i in range(n_jobs)])
    good = torch.equal(full, expect)
    good &= pdist.max_over_ranks(float(r + 1)) == float(w)
    ok[rank] = int(good)
    dist.destroy_process_group()


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in pdist.shard_indices(256, r, world))
        assert seen == list(range(256))
        assert pdist.shard_counts(256, world) == [256 // world] * world
    assert pdist.shard_counts(5, 2) == [3, 2]


def test_gather_frames_world2_gloo():
    port = _free_port()
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, 5, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def _grad_worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    pdist.init_process_group("gloo")
    from posegen_b200.train import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    net[1].bias.grad = None                                   # parameters without a gradient are skipped
    allreduce_gradients(net.parameters(), average=True)
    good = net[1].bias.grad is None
    for i, p in enumerate(list(net.parameters())[:3]):
        good &= bool(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))))     # mean of ranks' (1, 2) * (i + 1)
    # gradients that are views of one arena buffer (what the training backward hands out): one in-place all-reduce,
    # the views see the result, untouched gaps of the arena stay out of the parameters
    arena = torch.full((200,), 1000.0)
    o = 3
    for i, p in enumerate(net.parameters()):
        v = arena[o:o + p.numel()].view(p.shape)
        v.fill_(float(rank + 1) * (i + 2))
        p.grad = v
        o += p.numel() + 2
    allreduce_gradients(net.parameters(), average=True)
    for i, p in enumerate(net.parameters()):
        good &= bool(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 2))))
        good &= p.grad.untyped_storage().data_ptr() == arena.untyped_storage().data_ptr()
    good &= float(arena[0]) == 1000.0                          # outside [lo, hi) nothing was reduced
    ok[rank] = int(good)
    dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo():
    """The training step's only collective: one all-reduce of the flattened gradient bucket."""
    port = _free_port()
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_grad_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
