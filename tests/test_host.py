"""CPU: host logic, C-ABI surface, layout tables (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from posegen_b200 import _lib, raycaster as rcmod, synthetic as syn
from tests import parity_util as pu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(lib_path):
    header = open(os.path.join(ROOT, "include", "posegen_b200.h")).read()
    declared = sorted(set(re.findall(r"PGN_API\s+[\w\s\*]+?\b(pgn_\w+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTED)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by {lib_path}"
    assert lib.pgn_abi_version() == 2


def test_no_cpu_fallback_without_gpu(lib_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.Config(24, 64, 16, 7, 4, 8, 256, 4, 0)
    h = ctypes.c_void_p()
    rc = lib.pgn_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == _lib.PGN_E_CUDA
    assert b"no CPU fallback" in lib.pgn_last_error()
    with pytest.raises(RuntimeError):
        from posegen_b200.engine import Engine
        Engine()


def test_create_rejects_other_architectures(lib_path):
    lib = _lib.load()
    cfg = _lib.Config(24, 128, 16, 7, 4, 8, 256, 4, 0)
    h = ctypes.c_void_p()
    assert lib.pgn_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.PGN_E_INVALID


def test_linspace_tables_match_torch():
    # pgn_linspace01 (pgn_common.cuh) restated: symmetric evaluation like torch.linspace on CPU
    def lin(n):
        step = np.float32(1.0) / np.float32(n - 1)
        return np.array([step * np.float32(i) if i < n // 2 else np.float32(1.0 - float(step) * (n - 1 - i))  # fused
                         for i in range(n)], dtype=np.float32)
    assert np.array_equal(lin(64), torch.linspace(0., 1., 64).numpy())
    assert np.array_equal(lin(16), torch.linspace(0., 1., 16).numpy())


def test_bf16_k_permutations_are_bijections():
    src = r'''
#include <cstdio>
#include "pgn_bf16_layout.h"
int main() {
  for (int k = 0; k < 480; ++k) printf("%d ", pgn_xperm_refcol(k));
  printf("\n");
  for (int q = 0; q < 768; ++q) printf("%d ", pgn_dperm_refcol(q));
  printf("\n%zu\n", pgn_wstream_elems());
  int ks = 0; for (int L = 0; L < 9; ++L) ks += pgn_layer_ksteps(L);
  printf("%d\n", ks);
}'''
    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "t.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(tmp, "t")
        subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "posegen_b200", "csrc"), cpp, "-o", exe], check=True)
        lines = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    x = [int(v) for v in lines[0].split()]
    d = [int(v) for v in lines[1].split()]
    assert sorted(v for v in x if v >= 0) == list(range(432)) and x.count(-1) == 48
    assert sorted(v for v in d if v >= 0) == list(range(432, 1080)) and d.count(-1) == 120
    assert int(lines[2]) == 870400      # bf16 elements per net = executed tensor MACs per sample (zero pads + bias K-steps)
    assert int(lines[3]) == 245


def test_raycaster_drop_in_surface():
    kw_train, kw_test, start, grad_vars, optim, ckpt = rcmod.create_raycaster(rcmod.surreal_args(), {"skel_type": None})
    rc = kw_test["ray_caster"]
    assert sum(p.numel() for p in rc.parameters()) == 1728568
    assert list(rc.state_dict().keys()) == ["network_fn_state_dict", "network_fine_state_dict", "embed_state_dict",
                                            "embedbones_state_dict", "embeddirs_state_dict"]
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    assert set(rc.state_dict()["embed_state_dict"].keys()) == {"cutoff_dist", "tau"}
    state = syn.synthetic_raycaster_state(3, alpha_gain=400.)
    rc.load_state_dict(state)
    got = rc.state_dict()["network_fine_state_dict"]["pts_linears.5.weight"].numpy()
    assert np.array_equal(got, state["network_fine_state_dict"]["pts_linears.5.weight"])
    # tau schedule of core/cutoff_embedder.py:181-183
    rc.update_embed_fns(250000, rcmod.surreal_args())
    assert abs(rc.embed_fn.get_tau() - 200.0) < 1e-3
    rc.update_embed_fns(10 ** 7, rcmod.surreal_args())
    assert rc.embed_fn.get_tau() == 2000.0
    with pytest.raises(RuntimeError):
        rc.eval()(torch.zeros(4, 11), N_samples=64, N_importance=16, kp_batch=None,
                  skts=torch.zeros(4, 24, 4, 4), cyls=torch.zeros(4, 5))      # CPU tensors: no fallback


def test_create_raycaster_dataparallel_contract_and_checkpoint_reload(tmp_path):
    """core/raycasters.py:125-142,157,172 + core/cutoff_embedder.py:227-238: the training kwargs carry an nn.DataParallel
    (`.module` is dereferenced by core/trainer.py:267,272,506), the test kwargs the bare module; the newest *.tar of
    basedir/expname is reloaded with global_step and the optimizer state unless no_reload / finetune say otherwise."""
    kw_train, kw_test, start, grad_vars, optim, ckpt = rcmod.create_raycaster(rcmod.surreal_args(), {"skel_type": None})
    assert isinstance(kw_train["ray_caster"], torch.nn.DataParallel)
    assert kw_train["ray_caster"].module is kw_test["ray_caster"] and isinstance(kw_test["ray_caster"], rcmod.RayCaster)
    assert start == 0 and ckpt is None
    rc = kw_test["ray_caster"]
    # what Trainer.save_nerf writes (core/trainer.py:486-509)
    for p in grad_vars:
        p.grad = torch.full_like(p, 1e-3)
    optim.step()
    rc.update_embed_fns(125000, rcmod.surreal_args())
    exp = tmp_path / "exp"
    exp.mkdir()
    torch.save({"global_step": 100, "optimizer_state_dict": optim.state_dict(), "poseopt_layer_state_dict": None,
                **kw_train["ray_caster"].module.state_dict()}, exp / "000100.tar")
    torch.save({"global_step": 7, **rc.state_dict()}, exp / "000007.tar")            # older: must not be picked
    torch.save({"global_step": 999}, exp / "pose_000999.tar")                          # pose-only files are skipped
    args = rcmod.surreal_args(no_reload=False, basedir=str(tmp_path), expname="exp")
    _, kw2, start2, gv2, optim2, ckpt2 = rcmod.create_raycaster(args, {"skel_type": None})
    rc2 = kw2["ray_caster"]
    assert start2 == 100 and ckpt2["global_step"] == 100
    for a, b in zip(rc.parameters(), rc2.parameters()):
        assert torch.equal(a, b)
    assert abs(rc2.embed_fn.get_tau() - rc.embed_fn.get_tau()) < 1e-4 and rc2.embed_fn.get_tau() > 20.
    st = optim2.state[gv2[0]]
    assert int(st["step"]) == 1 and torch.equal(st["exp_avg"], optim.state[grad_vars[0]]["exp_avg"])
    # finetune: weights yes, step counter and optimizer state no
    _, _, start3, gv3, optim3, _ = rcmod.create_raycaster(rcmod.surreal_args(no_reload=False, finetune=True, basedir=str(tmp_path),
                                                                          expname="exp"), {"skel_type": None})
    assert start3 == 0 and len(optim3.state) == 0
    # explicit ft_path wins over the directory listing
    _, _, start4, _, _, _ = rcmod.create_raycaster(rcmod.surreal_args(no_reload=False, ft_path=str(exp / "000007.tar"),
                                                                      basedir=str(tmp_path), expname="exp"), {"skel_type": None})
    assert start4 == 7


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        rcmod.create_raycaster(rcmod.surreal_args(N_samples=128), {"skel_type": None})
    with pytest.raises(NotImplementedError):
        rcmod.NeRF(D=4)


def test_synthetic_geometry_properties():
    pose = syn.synthetic_pose(7)
    ident = pose.skts.astype(np.float64) @ pose.l2ws.astype(np.float64)
    assert np.abs(ident - np.eye(4)).max() < 1e-5
    frame = syn.synthetic_frame(7, 64, 64)
    assert frame.rays_o.shape == frame.rays_d.shape == (len(frame.valid_idx), 3)
    assert 0.5 < len(frame.valid_idx) / (64 * 64) <= 1.0
    full = syn.synthetic_frame(7, 64, 64, full_frame=True)
    assert len(full.valid_idx) == 64 * 64


def test_hmr_input_oracle_properties():
    """The skimage.resize restatement: identity when the crop already has the output size, constants stay constant."""
    import numpy as np
    from oracle import next_oracle as nxt
    rng = np.random.RandomState(0)
    img = rng.rand(64, 64, 3).astype(np.float32)
    out = nxt.hmr_input(img, crop=(8, 8, 40, 40), out_res=32, quantize_u8=False)
    want = (np.transpose(img[8:40, 8:40], (2, 0, 1)) - np.array([0.485, 0.456, 0.406], np.float32)[:, None, None]) / \
        np.array([0.485, 0.456, 0.406], np.float32)[:, None, None]
    assert np.abs(out - want).max() <= 1e-6
    flat = np.full((80, 80, 3), 1.0, np.float32)                      # white background frame
    out = nxt.hmr_input(flat, crop=(0, 0, 80, 80), out_res=24)
    assert np.abs(out - (1.0 - np.array([0.485, 0.456, 0.406], np.float32))[:, None, None] /
                  np.array([0.485, 0.456, 0.406], np.float32)[:, None, None]).max() <= 1e-5


def test_mlp_backward_matches_autograd_on_cpu():
    """Host logic of the training step (posegen_b200/train.py): the GEMM weight / input gradients of one NeRF MLP,
    fed with an activation dump in the kernel's layout (trunk layers tile-blocked, view layer row-major) built from a torch forward and a torch
    restatement of the fused delta pass (`pgn_mlp_delta`), against autograd through the oracle's nerf_forward."""
    import numpy as np
    import torch
    from oracle import render_oracle as orc
    from posegen_b200 import synthetic as syn
    from posegen_b200.train import mlp_backward, PARAM_ORDER
    torch.manual_seed(0)
    m = 192
    net = {k: torch.as_tensor(v) for k, v in syn.synthetic_nerf_state(5).items()}
    for v in net.values():
        v.requires_grad_(True)
    enc = (torch.randn(m, 1080) * 0.5).requires_grad_(True)
    raw = orc.nerf_forward(enc, net)
    d_raw = torch.randn(m, 4)
    (raw * d_raw).sum().backward()
    # activation dump: post-ReLU activations per layer, bf16, layers 0-7 [rows (padded), 256] then the view layer [rows, 128]
    with torch.no_grad():
        x_p, x_v = enc[:, :432], enc[:, 432:]
        h, acts = x_p, []
        for l in range(8):
            h = torch.relu(torch.nn.functional.linear(h, net[f"pts_linears.{l}.weight"], net[f"pts_linears.{l}.bias"]))
            acts.append(h)
            if l == 4:
                h = torch.cat([x_p, h], -1)
        feat = torch.nn.functional.linear(acts[7], net["feature_linear.weight"], net["feature_linear.bias"])
        g = torch.relu(torch.nn.functional.linear(torch.cat([feat, x_v], -1), net["views_linears.0.weight"], net["views_linears.0.bias"]))
        rows = 256
        dump = torch.zeros((rows * 2304,), dtype=torch.bfloat16)            # 2176 activations + 128 (mask bits, unused here) per row
        from posegen_b200.train import to_tile_blocked
        for l in range(8):                                                    # trunk layers: tile-blocked like the kernel's stores
            full = torch.zeros((rows, 256), dtype=torch.bfloat16)
            full[:m] = acts[l].to(torch.bfloat16)
            dump[l * rows * 256:(l + 1) * rows * 256] = to_tile_blocked(full)
        dump[8 * rows * 256:rows * 2176].view(rows, 128)[:m] = g.to(torch.bfloat16)

    def fuse(dh, act, rs, wr, has_input, want_wsum):          # what pgn_mlp_delta computes, in torch
        pre = dh.float() if has_input else torch.zeros(dh.shape)
        if rs is not None:
            pre = pre + rs @ wr.float()
        z = torch.where(act > 0, pre, torch.zeros(())) if act is not None else pre
        dh.copy_(z.to(torch.bfloat16))
        return z.sum(0), (rs.t() @ act.float() if (want_wsum and rs is not None) else None)

    params = {k: v.detach() for k, v in net.items()}
    got = mlp_backward(params, enc.detach().to(torch.bfloat16), dump, d_raw, fuse, want_input_grad=True)
    flat, ref = [], []
    for k in PARAM_ORDER:
        assert got[k].reshape(net[k].shape).shape == net[k].grad.shape
        flat.append(got[k].reshape(-1).double()); ref.append(net[k].grad.reshape(-1).double())
    a, b = torch.cat(flat), torch.cat(ref)
    assert float((a - b).norm() / b.norm()) <= 2e-2                      # bf16 activations / deltas / weights
    ge, gr = torch.cat([got["_g_xp"], got["_g_d"]], 1).double(), enc.grad.double()
    assert float((ge - gr).norm() / gr.norm()) <= 2e-2


def test_torch_fk_matches_numpy_helpers_and_is_differentiable():
    import numpy as np
    import torch
    from posegen_b200 import fk, synthetic as syn
    rng = np.random.RandomState(1)
    bones = (rng.randn(5, 24, 3) * 0.3).astype(np.float64)
    rest = (syn.SMPL_REST_POSE * syn.BODY_SCALE).astype(np.float64)
    b = torch.tensor(bones, requires_grad=True)
    skts, kps = fk.smpl_skts(b, torch.tensor(rest))
    ref = np.stack([syn.smpl_local_to_world(x, rest) for x in bones])
    assert np.abs(skts.detach().numpy() - syn.rigid_inverse(ref)).max() <= 1e-10
    assert np.abs(kps.detach().numpy() - ref[:, :, :3, 3]).max() <= 1e-10
    (skts[..., :3, :] ** 2).sum().backward()
    assert torch.isfinite(b.grad).all() and float(b.grad.abs().max()) > 0
    assert torch.autograd.gradcheck(lambda x: fk.smpl_skts(x, torch.tensor(rest))[0][..., :3, :], (b[:1].detach().requires_grad_(True),), atol=1e-6)


def test_chain_weight_stream_layout():
    """train.chain_wstream: the 60 fills (two K=16 steps each) pgn_mlp_delta_chain streams,
    [K/32][2 N halves][2 K-steps][2][128][8] per weight, in the order fold layer, W_7 .. W_1 (include/posegen_b200.h)."""
    import torch
    from posegen_b200 import synthetic as syn
    from posegen_b200.train import chain_wstream
    P = {k: torch.as_tensor(v) for k, v in syn.synthetic_nerf_state(3).items()}
    ws = chain_wstream(P)
    assert ws.dtype == torch.bfloat16 and ws.numel() == 120 * 4096
    fold = (P["views_linears.0.weight"][:, :256] @ P["feature_linear.weight"]).t().to(torch.bfloat16)      # [256,128]
    mats = [fold] + [(P[f"pts_linears.{l}.weight"][:, 432:] if l == 5 else P[f"pts_linears.{l}.weight"]).t().to(torch.bfloat16)
                     for l in range(7, 0, -1)]
    off = 0
    for w in mats:
        K = w.shape[1]
        fills = ws[off:off + 256 * K].view(K // 32, 2, 2, 2, 128, 8)
        for (f, h, g, kc, n, e) in ((0, 0, 0, 0, 0, 0), (K // 32 - 1, 1, 1, 1, 127, 7), (1, 0, 1, 0, 17, 3), (2, 1, 0, 1, 72, 5)):
            assert fills[f, h, g, kc, n, e] == w[h * 128 + n, (2 * f + g) * 16 + kc * 8 + e]
        off += 256 * K
    assert off == ws.numel()


def test_activation_dump_views():
    """train.act_layer / act_masks address the flat dump buffer of pgn_render_forward_train as documented."""
    import torch
    from posegen_b200.train import ACT_ROW_ELEMS, act_layer, act_masks
    rows, m = 1024, 1000
    buf = torch.arange(rows * ACT_ROW_ELEMS, dtype=torch.float32).to(torch.bfloat16)
    assert act_layer(buf, 0, m).shape == (m, 256) and act_layer(buf, 8, m).shape == (m, 128)
    assert act_layer(buf, 8, m).data_ptr() == buf.data_ptr() + 8 * rows * 256 * 2
    # trunk layers are tile-blocked: element (r, c) of layer l sits at ((r // 128) * 32 + c // 8) * 1024 + (r % 128) * 8 + c % 8
    from posegen_b200.train import to_tile_blocked
    x = torch.randn(rows, 256).to(torch.bfloat16)
    buf[3 * rows * 256:4 * rows * 256] = to_tile_blocked(x)
    assert torch.equal(act_layer(buf, 3, m), x[:m])
    r, c = 517, 203
    assert buf[3 * rows * 256 + ((r // 128) * 32 + c // 8) * 1024 + (r % 128) * 8 + c % 8] == x[r, c]
    mask, r = act_masks(buf)
    assert r == rows and mask.data_ptr() == buf.data_ptr() + rows * 4352 and mask.numel() * 2 == rows * 256
