"""GPU: the training step (BASELINE.json configs[3], reference core/trainer.py:232-370) — fused bf16 forward with
activation dump, hand-written compositing backward, GEMM weight gradients — against torch autograd through the
oracle (the op-for-op restatement of the reference) on the same rays, weights and loss."""
import numpy as np
import pytest
import torch

from oracle import render_oracle as orc
from posegen_b200 import synthetic as syn
from posegen_b200.raycaster import raycaster_from_checkpoint
from posegen_b200.train import PARAM_ORDER, allreduce_gradients

pytestmark = pytest.mark.gpu


def _oracle_loss(rb, sk, cy, nets, emb, tgt, bg=1.0, rand=None, cams=None):
    """render_rays with the reference's detach of the importance samples (core/utils/ray_utils.py:286) and the
    trainer's loss: MSE(rgb_map + (1 - acc) bg, tgt) + MSE(rgb0 + (1 - acc0) bg, tgt)  (core/trainer.py:355-370)."""
    rays_o, rays_d = rb[:, 0:3], rb[:, 3:6]
    near, far = orc.near_far_in_cylinder(rays_o, rays_d, cy, rb[:, 6:7], rb[:, 7:8])
    rand = rand or {}
    z = orc.coarse_z_vals(near, far, 64, t_rand=rand.get("t_rand"))
    enc = orc.encode(rays_o[:, None] + rays_d[:, None] * z[:, :, None], rays_d, sk, emb)
    n = rb.shape[0]
    raw0 = orc.nerf_forward(enc.reshape(-1, 1080), nets[0], frame_code=orc.frame_codes(nets[0], cams, 64, n, training=True)).reshape(-1, 64, 4)
    r0 = orc.raw2outputs(raw0, z, rays_d, noise=rand.get("noise0"))
    z_all, _, _, _, _ = orc.importance_z_vals(z, r0["weights"].detach(), 16, u=rand.get("u_is"))
    enc_f = orc.encode(rays_o[:, None] + rays_d[:, None] * z_all[:, :, None], rays_d, sk, emb)
    raw = orc.nerf_forward(enc_f.reshape(-1, 1080), nets[1], frame_code=orc.frame_codes(nets[1], cams, 80, n, training=True)).reshape(-1, 80, 4)
    r = orc.raw2outputs(raw, z_all, rays_d, noise=rand.get("noise"))
    loss = ((r["rgb_map"] + (1 - r["acc_map"][:, None]) * bg - tgt) ** 2).mean() + \
        ((r0["rgb_map"] + (1 - r0["acc_map"][:, None]) * bg - tgt) ** 2).mean()
    return loss, r, r0


@pytest.fixture(scope="module")
def train_case():
    frame = syn.synthetic_frame(3, 64, 64)
    ckpt = syn.synthetic_raycaster_state(4, alpha_gain=40.0)           # semi-transparent volume: smooth gradients
    n = 1536
    rb = syn.ray_batch(frame.rays_o, frame.rays_d)[:n]
    rng = np.random.RandomState(0)
    tgt = rng.rand(n, 3).astype(np.float32)
    return frame, ckpt, rb, tgt


def test_training_step_gradients_match_oracle_autograd(engine, train_case):
    frame, ckpt, rb, tgt = train_case
    n = rb.shape[0]
    dev = torch.device("cuda")
    # ---- oracle: fp32 autograd on the host
    nets = orc.nets_from_ckpt(ckpt)
    emb = orc.embed_params_from_ckpt(ckpt)
    for net in nets:
        for v in net.values():
            v.requires_grad_(True)
    sk = torch.as_tensor(frame.pose.skts)[None].expand(n, -1, -1, -1)
    cy = torch.as_tensor(frame.pose.cyl)[None].expand(n, -1)
    loss_ref, r_ref, r0_ref = _oracle_loss(torch.as_tensor(rb), sk, cy, nets, emb, torch.as_tensor(tgt))
    loss_ref.backward()
    # ---- ours: RayCaster in train mode
    rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
    rc.train()
    ret = rc(torch.as_tensor(rb, device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk.to(dev).contiguous(),
             cyls=cy.to(dev).contiguous(), bones=None, cams=None, subject_idxs=None, perturb=0., raw_noise_std=0.)
    t = torch.as_tensor(tgt, device=dev)
    loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    engine.check_status()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-3 * max(1.0, abs(float(loss_ref.detach())))
    flat, flat_ref = [], []
    for net, ref in ((rc.network, nets[0]), (rc.network_fine, nets[1])):
        pd = dict(net.named_parameters())
        for k in PARAM_ORDER:
            g, gr = pd[k].grad.detach().cpu().double(), ref[k].grad.double()
            assert g.shape == gr.shape and torch.isfinite(g).all(), k
            flat.append(g.reshape(-1)); flat_ref.append(gr.reshape(-1))
            if float(gr.norm()) > 1e-7:
                rel = float((g - gr).norm() / gr.norm())
                assert rel <= 6e-2, (k, rel)                     # bf16 forward activations and deltas
    g, gr = torch.cat(flat), torch.cat(flat_ref)
    cos = float((g @ gr) / (g.norm() * gr.norm()))
    assert cos >= 0.999, cos
    assert float((g - gr).norm() / gr.norm()) <= 3e-2


@pytest.mark.parametrize("fused", [False, True])
def test_training_step_reduces_the_loss(engine, train_case, fused):
    """A few Adam steps through the drop-in RayCaster (weights are re-packed after every optimizer.step; torch's fused
    Adam updates parameters without bumping their version counters, which must not leave the packed copy stale)."""
    frame, ckpt, rb, tgt = train_case
    n = 1024
    dev = torch.device("cuda")
    rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
    rc.train()
    opt = torch.optim.Adam([p for p in rc.parameters() if p.requires_grad], lr=5e-4, fused=fused)
    rbt = torch.as_tensor(rb[:n], device=dev)
    sk = torch.as_tensor(frame.pose.skts, device=dev)
    cy = torch.as_tensor(frame.pose.cyl, device=dev)
    t = torch.full((n, 3), 0.25, device=dev)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        ret = rc(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
        loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
        loss.backward()
        allreduce_gradients(rc.parameters())          # no-op for world size 1
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.7 * losses[0], losses
    # the packed copy is the parameters as the last optimizer.step left them: same render as a fresh module
    rc.eval()
    with torch.no_grad():
        a = rc(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
        rc2 = raycaster_from_checkpoint({k: ({kk: vv.detach().cpu() for kk, vv in v.items()} if isinstance(v, dict) else v)
                                         for k, v in rc.state_dict().items()}, device="cuda", precision="bf16")
        rc2.eval()
        b = rc2(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
    assert torch.equal(a["rgb_map"], b["rgb_map"]) and torch.equal(a["acc_map"], b["acc_map"])


def test_pose_gradient_matches_oracle_autograd(engine, train_case):
    """BASELINE.json configs[4]: differentiable render, gradient w.r.t. the pose / bone transforms (skts)."""
    frame, ckpt, rb, tgt = train_case
    n = 768
    dev = torch.device("cuda")
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    sk_ref = torch.as_tensor(frame.pose.skts)[None].repeat(n, 1, 1, 1).requires_grad_(True)       # per-ray leaf, like the reference
    cy = torch.as_tensor(frame.pose.cyl)[None].expand(n, -1)
    loss_ref, _, _ = _oracle_loss(torch.as_tensor(rb[:n]), sk_ref, cy, nets, emb, torch.as_tensor(tgt[:n]))
    loss_ref.backward()
    rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
    rc.train()
    sk = torch.as_tensor(frame.pose.skts, device=dev)[None].repeat(n, 1, 1, 1).requires_grad_(True)
    ret = rc(torch.as_tensor(rb[:n], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy.to(dev).contiguous(),
             bones=None, cams=None, perturb=0., raw_noise_std=0.)
    t = torch.as_tensor(tgt[:n], device=dev)
    loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    engine.check_status()
    g, gr = sk.grad.cpu().double(), sk_ref.grad.double()
    assert g.shape == gr.shape and torch.isfinite(g).all()
    assert float(g[:, :, 3].abs().max()) == 0.0                       # the homogeneous row carries no gradient
    rel = float((g - gr).norm() / gr.norm())
    cos = float((g.reshape(-1) @ gr.reshape(-1)) / (g.norm() * gr.norm()))
    assert cos >= 0.995 and rel <= 0.1, (cos, rel)
    # a pose shared by all rays ([24,4,4]) receives the sum over rays
    sk1 = torch.as_tensor(frame.pose.skts, device=dev).clone().requires_grad_(True)
    ret = rc(torch.as_tensor(rb[:n], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk1, cyls=torch.as_tensor(frame.pose.cyl, device=dev),
             bones=None, cams=None, perturb=0., raw_noise_std=0.)
    loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
    loss.backward()
    assert float((sk1.grad.cpu().double() - gr.sum(0)).norm() / gr.sum(0).norm()) <= 0.1


def test_training_step_with_sampling_noise_matches_oracle(engine, train_case):
    """The reference's training-time randomness (perturb = 1: stratified jitter + random importance quantiles;
    raw_noise_std: density noise) with the SAME random numbers fed to the oracle and to the kernel."""
    frame, ckpt, rb, tgt = train_case
    n = 1024
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(123)
    rand = {"t_rand": torch.rand((n, 64), generator=g), "u_is": torch.rand((n, 16), generator=g),
            "noise0": torch.randn((n, 64), generator=g) * 0.3, "noise": torch.randn((n, 80), generator=g) * 0.3}
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    for net in nets:
        for v in net.values():
            v.requires_grad_(True)
    sk = torch.as_tensor(frame.pose.skts)[None].expand(n, -1, -1, -1)
    cy = torch.as_tensor(frame.pose.cyl)[None].expand(n, -1)
    loss_ref, r_ref, r0_ref = _oracle_loss(torch.as_tensor(rb[:n]), sk, cy, nets, emb, torch.as_tensor(tgt[:n]), rand=rand)
    loss_ref.backward()
    rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
    rc.train()
    ret = rc(torch.as_tensor(rb[:n], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk.to(dev).contiguous(),
             cyls=cy.to(dev).contiguous(), bones=None, cams=None, perturb=1.0, raw_noise_std=1.0,
             train_random={k: v.to(dev).contiguous() for k, v in rand.items()})
    t = torch.as_tensor(tgt[:n], device=dev)
    loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    engine.check_status()
    # forward: the rays whose last (delta = 1e10) sample sits on the ReLU edge can flip between the bf16 and fp32 nets
    for k, ref in (("rgb0", r0_ref["rgb_map"]), ("acc0", r0_ref["acc_map"]), ("rgb_map", r_ref["rgb_map"]), ("acc_map", r_ref["acc_map"])):
        err = (ret[k].detach().cpu() - ref.detach()).abs()
        err = err.reshape(n, -1).max(1).values
        assert float((err <= 2e-2).float().mean()) >= 0.97, k
    flat, flat_ref = [], []
    for net, ref in ((rc.network, nets[0]), (rc.network_fine, nets[1])):
        pd = dict(net.named_parameters())
        for k in PARAM_ORDER:
            flat.append(pd[k].grad.detach().cpu().double().reshape(-1)); flat_ref.append(ref[k].grad.double().reshape(-1))
    a, b = torch.cat(flat), torch.cat(flat_ref)
    assert torch.isfinite(a).all()
    cos = float((a @ b) / (a.norm() * b.norm()))
    assert cos >= 0.99, cos
    # without explicit numbers the drop-in draws its own: two calls differ, the result is finite
    r1 = rc(torch.as_tensor(rb[:256], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk[:256].to(dev).contiguous(),
            cyls=cy[:256].to(dev).contiguous(), perturb=1.0, raw_noise_std=1.0)
    r2 = rc(torch.as_tensor(rb[:256], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=sk[:256].to(dev).contiguous(),
            cyls=cy[:256].to(dev).contiguous(), perturb=1.0, raw_noise_std=1.0)
    assert torch.isfinite(r1["rgb_map"]).all() and not torch.equal(r1["rgb_map"], r2["rgb_map"])


def test_mlp_delta_kernel_matches_torch(engine):
    """pgn_mlp_delta (ReLU backward + bias gradient + skinny heads in one pass) against the same arithmetic in torch,
    every supported shape, ragged row counts (incl. rows that are no multiple of the block tile)."""
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for m, cols, nrs, has_in, masked in ((1000, 256, 0, True, True), (4097, 256, 1, True, True), (777, 128, 3, False, True),
                                         (513, 128, 0, True, False), (1, 256, 1, True, True)):
        dh = torch.randn((m, cols), device=dev, generator=g).to(torch.bfloat16)
        act = torch.relu(torch.randn((m, cols), device=dev, generator=g)).to(torch.bfloat16) if masked else None
        raw = torch.randn((m, 4), device=dev, generator=g)
        rs = {0: None, 1: raw[:, 3:4], 3: raw[:, :3]}[nrs]
        wr = torch.randn((nrs, cols), device=dev, generator=g) if nrs else None
        pre = dh.float() if has_in else torch.zeros((m, cols), device=dev)
        if nrs:
            pre = pre + rs @ wr
        want = torch.where(act > 0, pre, torch.zeros((), device=dev)) if masked else pre
        got = dh.clone()
        colsum, wsum = engine.mlp_delta(got, act, rs, wr, has_input=has_in, want_wsum=nrs > 0)
        engine.check_status()
        if nrs == 0:
            assert torch.equal(got, want.to(torch.bfloat16)), (m, cols, nrs)
        else:       # the kernel forms dh + rs @ wr with FMAs: within one bf16 rounding of the two-step torch result
            assert bool(((got.float() - want).abs() <= 2.0 ** -7 * want.abs() + 1e-6).all()), (m, cols, nrs)
            assert torch.equal(got == 0, want.to(torch.bfloat16) == 0) or masked
        ref = want.double().sum(0)
        assert float((colsum.double() - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())) + 1e-3
        if nrs:
            wref = rs.double().t() @ act.double()
            assert float((wsum.double() - wref).abs().max()) <= 1e-4 * max(1.0, float(wref.abs().max())) + 1e-3


def test_encode_bf16_is_the_rounded_fp32_encoding(engine, train_case):
    """pgn_encode_bf16 against pgn_encode (the parity-checked fp32 encoding) within one bf16 rounding plus the fast-math
    error of the double-angle recurrences the tensor-core tier uses (<= 2e-4), incl. a ragged last row block."""
    frame, ckpt, rb, tgt = train_case
    engine.load_checkpoint(ckpt)
    dev = torch.device("cuda")
    n = 37
    rbt = torch.as_tensor(rb[:n], device=dev)
    sk = torch.as_tensor(frame.pose.skts, device=dev)
    cy = torch.as_tensor(frame.pose.cyl, device=dev)
    z = torch.rand((n, 5), device=dev).sort(-1).values * 2 + 3
    a = engine.encode(rbt, sk, cy, z)
    b = engine.encode_bf16(rbt, sk, cy, z)
    assert b.dtype == torch.bfloat16 and b.shape == a.shape
    assert bool(((a - b.float()).abs() <= 2.0 ** -8 * a.abs() + 3e-4).all())


def test_forward_dump_masks_match_the_activations(engine, train_case):
    """The 1-bit ReLU masks the training forward dumps behind the activations are [activation > 0] of the same dump."""
    from posegen_b200.train import act_layer, act_masks
    frame, ckpt, rb, tgt = train_case
    engine.load_checkpoint(ckpt)
    dev = torch.device("cuda")
    n = 100                                                        # ragged: not a multiple of the 8-ray groups
    ret, acts = engine.render_train(torch.as_tensor(rb[:n], device=dev), torch.as_tensor(frame.pose.skts, device=dev),
                                    torch.as_tensor(frame.pose.cyl, device=dev))
    engine.check_status()
    for key, s in (("c", 64), ("f", 80)):
        m = n * s
        mask, rows = act_masks(acts[key])
        bits = mask.view(torch.int32).view(8, 8, rows)[:, :, :m].permute(0, 2, 1)   # word planes -> 8 words of 32 columns per row
        got = ((bits[..., None] >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(8, m, 256).bool()
        for l in range(8):
            assert torch.equal(got[l], act_layer(acts[key], l, m) > 0), (key, l)


def test_delta_chain_kernel_matches_layerwise_reference(engine):
    """pgn_mlp_delta_chain (eight tcgen05 layers per 512-row block of a CTA pair, masks from bits, TMA-stored deltas, bias
    column sums) against the same chain in torch with bf16 rounding at the same places (fp32 accumulation, bf16 deltas);
    ragged row counts, incl. blocks whose second tile / peer CTA holds no valid row."""
    from posegen_b200.train import chain_wstream
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    P = {k: torch.as_tensor(v, device=dev) for k, v in syn.synthetic_nerf_state(7).items()}
    for m in (100, 256, 1000, 37 * 1024 + 5):
        rows = ((m + 1279) // 1280) * 1280
        dG = (torch.randn((m, 128), device=dev, generator=g) * 0.1).to(torch.bfloat16)
        d_raw = torch.randn((m, 4), device=dev, generator=g) * 0.1
        act = torch.rand((8, rows, 256), device=dev, generator=g) > 0.4
        words = (act.view(8, rows, 8, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(-1)
        mask = words.to(torch.int32)                                                   # wraps bit 31 into the sign
        mask = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).permute(0, 2, 1).contiguous()   # word planes [8][8][rows]
        dz, colsum = engine.mlp_delta_chain(dG, d_raw.contiguous(), mask, rows, chain_wstream(P),
                                            P["alpha_linear.weight"].reshape(-1).float().contiguous())
        engine.check_status()
        bf = torch.bfloat16
        fold = (P["views_linears.0.weight"][:, :256] @ P["feature_linear.weight"]).to(bf).float()
        pre = dG.float() @ fold + d_raw[:, 3:4] * P["alpha_linear.weight"].float()
        for l in range(7, -1, -1):
            z = torch.where(act[l, :m], pre, torch.zeros((), device=dev))
            zb = z.to(bf)
            err = (dz[l].float() - zb.float()).abs()
            tol = 2.0 ** -7 * zb.float().abs() + 1e-5                                   # one bf16 rounding (accumulation order)
            assert bool((err <= tol).all()), (m, l, float(err.max()))
            ref = dz[l].double().sum(0)                # the bias gradient is the column sum of the bf16 deltas it wrote
            assert float((colsum[l].double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), (m, l)
            assert float((colsum[l].double() - z.double().sum(0)).abs().max()) <= 1e-2 * max(1.0, float(ref.abs().max())), (m, l)
            if l > 0:
                W = P[f"pts_linears.{l}.weight"][:, 432:] if l == 5 else P[f"pts_linears.{l}.weight"]
                pre = dz[l].float() @ W.to(bf).float()                                  # continue from the kernel's own bf16 deltas
        # layer mask: only the requested deltas / sums are produced, bit-identical to the full run
        dz2, cs2 = engine.mlp_delta_chain(dG, d_raw.contiguous(), mask, rows, chain_wstream(P),
                                          P["alpha_linear.weight"].reshape(-1).float().contiguous(), layer_mask=0x21)
        engine.check_status()
        assert torch.equal(dz2[0], dz[0]) and torch.equal(dz2[5], dz[5])
        # the weight stream the library packs at upload time (pgn_mlp_delta_chain_net) against the torch-packed one: same
        # weights, the fold W_v[:, :256] W_f summed in another order (a bf16 rounding apart in a few elements)
        engine.upload_net(0, {k: v for k, v in P.items()})
        dz3, cs3 = engine.mlp_delta_chain_net(0, dG, d_raw.contiguous(), mask, rows)
        engine.check_status()
        for l in range(8):
            rel = float((dz3[l].float() - dz[l].float()).norm() / dz[l].float().norm().clamp_min(1e-12))
            assert rel <= 1e-2, (m, l, rel)
        assert float(cs2[[1, 2, 3, 4, 6, 7]].abs().max()) == 0.0 and float((cs2[0] - colsum[0]).abs().max()) <= 1e-3 * max(1.0, float(colsum[0].abs().max()))


def test_graphed_training_step_follows_the_eager_one(engine, train_case):
    """GraphedTrainStep (whole step captured into a CUDA graph, weights re-packed inside the graph) against the same
    steps issued eagerly from the same initial weights: same loss trajectory."""
    from posegen_b200.train import GraphedTrainStep
    frame, ckpt, rb, tgt = train_case
    n, dev = 1024, torch.device("cuda")
    rbt = torch.as_tensor(rb[:n], device=dev)
    sk = torch.as_tensor(np.repeat(frame.pose.skts[None], n, 0), device=dev)
    cy = torch.as_tensor(np.repeat(frame.pose.cyl[None], n, 0), device=dev)
    t = torch.full((n, 3), 0.25, device=dev)

    def loss_fn(ret, tg):
        return ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - tg) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - tg) ** 2).mean()

    losses = {}
    for mode in ("eager", "graph"):
        rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
        rc.train()
        opt = torch.optim.Adam([p for p in rc.parameters() if p.requires_grad], lr=5e-4, fused=True, capturable=True)
        if mode == "eager":
            for _ in range(6):
                opt.zero_grad(set_to_none=True)
                ret = rc(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
                loss = loss_fn(ret, t)
                loss.backward()
                opt.step()
            losses[mode] = float(loss.detach())
        else:
            g = GraphedTrainStep(rc, opt, loss_fn, {"ray_batch": rbt, "skts": sk, "cyls": cy, "target": t}, warmup=3,
                                 perturb=0., raw_noise_std=0.)
            for _ in range(3):
                loss = g(ray_batch=rbt, skts=sk, cyls=cy, target=t)
            losses[mode] = float(loss)
    assert abs(losses["eager"] - losses["graph"]) <= 2e-3 * max(1.0, abs(losses["eager"])), losses


def test_graph_replay_survives_a_larger_eager_render(engine, train_case):
    """the per-call near/far scratch must not be a cached, growable buffer - a captured step would keep the
    old pointer.  Capture a step, render a much larger batch eagerly through the SAME module, empty the allocator
    cache, replay: the replayed losses continue the trajectory of an undisturbed twin."""
    from posegen_b200.train import GraphedTrainStep
    frame, ckpt, rb, tgt = train_case
    n, dev = 512, torch.device("cuda")
    rbt = torch.as_tensor(rb[:n], device=dev)
    sk = torch.as_tensor(np.repeat(frame.pose.skts[None], n, 0), device=dev)
    cy = torch.as_tensor(np.repeat(frame.pose.cyl[None], n, 0), device=dev)
    t = torch.full((n, 3), 0.25, device=dev)

    def loss_fn(ret, tg):
        return ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - tg) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - tg) ** 2).mean()

    big = torch.as_tensor(np.tile(rb, (200, 1)), device=dev)             # 307k rays: a 2.4 MB near/far scratch
    out = {}
    for disturb in (False, True):
        rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
        rc.train()
        opt = torch.optim.Adam([p for p in rc.parameters() if p.requires_grad], lr=5e-4, fused=True, capturable=True)
        g = GraphedTrainStep(rc, opt, loss_fn, {"ray_batch": rbt, "skts": sk, "cyls": cy, "target": t}, warmup=3, perturb=0., raw_noise_std=0.)
        g(ray_batch=rbt, skts=sk, cyls=cy, target=t)
        if disturb:
            rc.eval()
            with torch.no_grad():
                r = rc(big, N_samples=64, N_importance=16, kp_batch=None, skts=torch.as_tensor(frame.pose.skts, device=dev),
                       cyls=torch.as_tensor(frame.pose.cyl, device=dev), bones=None, cams=None)
            assert torch.isfinite(r["rgb_map"]).all()
            rc.train()
            del r
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            junk = [torch.full((1 << 20,), float("nan"), device=dev) for _ in range(8)]      # recycle freed blocks with poison
        losses = [float(g(ray_batch=rbt, skts=sk, cyls=cy, target=t)) for _ in range(3)]
        torch.cuda.synchronize()
        rc.engine(dev).check_status()
        out[disturb] = losses
    assert np.isfinite(out[True]).all()
    assert np.allclose(out[True], out[False], rtol=2e-3, atol=1e-5), out


@pytest.mark.parametrize("m,Ma,Nb,n_ctas", [(4096, 256, 256, 7), (1000, 256, 176, 3), (4097, 128, 136, 148), (33, 128, 256, 5), (20000, 256, 256, 148)])
def test_wgrad_split_k_kernel_matches_torch(engine, m, Ma, Nb, n_ctas):
    """pgn_debug_wgrad: out[Ma,Nb] = A[:, :Ma]^T B[:, :Nb] over all rows, bf16 operands (strided row-major views, as the
    activation dump / the [m,1080] network input are), fp32 accumulation in TMEM, split-K over n_ctas CTAs."""
    g = torch.Generator(device="cuda").manual_seed(m + Ma + Nb)
    A_full = (torch.randn((m, 256), device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B_full = (torch.randn((m, 1080), device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    A = A_full[:, :Ma]
    B = B_full[:, 432:432 + Nb]
    got = engine.debug_wgrad(A, B, Ma, Nb, n_ctas=n_ctas)
    torch.cuda.synchronize()
    engine.check_status()
    want = A.double().t() @ B.double()
    err = float((got.double() - want).abs().max())
    scale = float(want.abs().max())
    assert err <= 2e-4 * max(1.0, scale) + 1e-3 * (m ** 0.5) * 1e-3, (err, scale)
    # the same product with B in the tile-blocked layout of the activation dump (one 64 KB bulk copy per 128-row tile,
    # SWIZZLE_NONE descriptor)
    if Nb == 256:
        from posegen_b200.train import to_tile_blocked
        got_tb = engine.debug_wgrad(A, to_tile_blocked(B.contiguous()), Ma, Nb, n_ctas=n_ctas, b_tile_blocked=True)
        torch.cuda.synchronize()
        engine.check_status()
        assert float((got_tb.double() - want).abs().max()) <= 2e-4 * max(1.0, scale) + 1e-3 * (m ** 0.5) * 1e-3
    # accumulates into `out` (split-K partials are added): a second call doubles the result
    got2 = engine.debug_wgrad(A, B, Ma, Nb, n_ctas=n_ctas, out=got.clone())
    assert float((got2.double() - 2 * want).abs().max()) <= 4e-4 * max(1.0, scale) + 2e-6 * m ** 0.5


def test_fused_weight_gradients_match_the_gemm_formulation(engine, train_case):
    """pgn_mlp_weight_grads (split-K tcgen05 + the feature/view fold + the alpha head) against the same gradients formed
    by plain matrix products (posegen_b200.train.mlp_backward without the kernel) on the same deltas."""
    import posegen_b200.train as tr
    frame, ckpt, rb, tgt = train_case
    n, dev = 1024, torch.device("cuda")
    rbt = torch.as_tensor(rb[:n], device=dev)
    sk = torch.as_tensor(frame.pose.skts, device=dev)
    cy = torch.as_tensor(frame.pose.cyl, device=dev)
    t = torch.as_tensor(tgt[:n], device=dev)
    grads = {}
    for use in (True, False):
        tr.USE_WGRAD_KERNEL = use
        try:
            rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
            rc.train()
            ret = rc(rbt, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
            loss = ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()
            loss.backward()
            torch.cuda.synchronize()
            rc.engine(dev).check_status()
            grads[use] = {(i, k): p.grad.detach().double().clone() for i, net in enumerate((rc.network, rc.network_fine)) for k, p in net.named_parameters()}
        finally:
            tr.USE_WGRAD_KERNEL = True
    for key, ref in grads[False].items():
        got = grads[True][key]
        assert got.shape == ref.shape and torch.isfinite(got).all(), key
        if float(ref.norm()) > 1e-9:
            rel = float((got - ref).norm() / ref.norm())
            assert rel <= 5e-3, (key, rel)           # same bf16 operands, fp32 accumulation in a different order


@pytest.mark.parametrize("m", [128, 1000, 40000])
def test_input_grad_kernel_matches_matrix_products(engine, m):
    """pgn_mlp_input_grads (tcgen05: K-major SWIZZLE_128B deltas x MN-major weights, five <= 256-column jobs per 128-row
    tile) against the three matrix products it replaces, on random deltas and the uploaded synthetic weights."""
    ckpt = syn.synthetic_raycaster_state(4, alpha_gain=40.0)
    engine.load_checkpoint(ckpt)
    g = torch.Generator(device="cuda").manual_seed(m)
    dz = (torch.randn((8, m, 256), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    dG = (torch.randn((m, 128), device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    g_xp, g_d = engine.mlp_input_grads(1, dz, dG)
    torch.cuda.synchronize()
    engine.check_status()
    net = {k: torch.as_tensor(v, device="cuda") for k, v in ckpt["network_fine_state_dict"].items()}
    bf = lambda w: w.to(torch.bfloat16).double()                                   # noqa: E731  (the kernel reads bf16 weights)
    want_xp = dz[5].double() @ bf(net["pts_linears.5.weight"][:, :432]) + dz[0].double() @ bf(net["pts_linears.0.weight"])
    want_d = dG.double() @ bf(net["views_linears.0.weight"][:, 256:904])
    for got, want in ((g_xp, want_xp), (g_d, want_d)):
        assert got.shape == want.shape
        err = float((got.double() - want).abs().max())
        assert err <= 8e-3 * max(1.0, float(want.abs().max())), err               # one bf16 rounding of the result
    # the tile-blocked form (what the training / GAN backward hands to pgn_encode_backward_bf16): the same numbers in
    # [ceil(m/128)][cols/8][128][8], rows beyond m zero
    t_xp, t_d = engine.mlp_input_grads(1, dz, dG, tile_blocked=True)
    torch.cuda.synchronize()
    engine.check_status()
    for tb, rm, cols in ((t_xp, g_xp, 432), (t_d, g_d, 648)):
        full = engine.from_tile_blocked(tb, tb.numel() // cols, cols)
        assert torch.equal(full[:m], rm)
        assert float(full[m:].abs().max()) == 0.0 if full.shape[0] > m else True


def test_training_gradients_share_one_arena(engine, train_case):
    """Every weight / bias gradient of both nets is a view of ONE buffer after backward (the weight-gradient kernel writes
    into it, the small tensors are copied behind): the data-parallel all-reduce is a single in-place collective."""
    frame, ckpt, rb, tgt = train_case
    n, dev = 512, torch.device("cuda")
    rc = raycaster_from_checkpoint(ckpt, device="cuda", precision="bf16")
    rc.train()
    ret = rc(torch.as_tensor(rb[:n], device=dev), N_samples=64, N_importance=16, kp_batch=None, skts=torch.as_tensor(frame.pose.skts, device=dev),
             cyls=torch.as_tensor(frame.pose.cyl, device=dev), bones=None, cams=None, perturb=0., raw_noise_std=0.)
    t = torch.as_tensor(tgt[:n], device=dev)
    (((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - t) ** 2).mean() + ((ret["rgb0"] + (1 - ret["acc0"][:, None]) - t) ** 2).mean()).backward()
    ps = [p for p in rc.parameters() if p.grad is not None]
    assert len(ps) == 48
    assert len({p.grad.untyped_storage().data_ptr() for p in ps}) == 1
    spans = sorted((p.grad.storage_offset(), p.grad.storage_offset() + p.grad.numel()) for p in ps)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))            # disjoint views


def test_ray_group_sizes_agree(engine, train_case):
    """The fused kernel cuts a batch into ray groups of 8, 6 or 4 (pgn_render_bf16.cu tile_of; chosen from the ray count):
    3,072 rays run as groups of 6, the same rays in three launches of 1,024 as groups of 4.  Per-sample network outputs,
    sample positions and the activation dump are identical, the composited maps agree up to the association order of the
    fine-pass compositing across tile boundaries."""
    from posegen_b200.train import act_layer, act_masks
    frame, ckpt, rb, tgt = train_case
    engine.load_checkpoint(ckpt)
    dev = torch.device("cuda")
    big = syn.synthetic_frame(5, 128, 128)
    rays = torch.as_tensor(syn.ray_batch(big.rays_o, big.rays_d), device=dev)
    assert rays.shape[0] >= 2 * 3072
    rays = rays[::rays.shape[0] // 3072][:3072].contiguous()                     # spread over the whole bbox
    sk, cy = torch.as_tensor(big.pose.skts, device=dev), torch.as_tensor(big.pose.cyl, device=dev)
    # inference kernels
    # (near / far of rays that miss the cylinder are filled from their chunk's statistics: same 1,024-ray chunks in both)
    full = engine.render(rays, sk, cy, nanfill_chunk=1024, precision="bf16", return_alpha=False)
    parts = [engine.render(rays[i:i + 1024], sk, cy, nanfill_chunk=1024, precision="bf16", return_alpha=False) for i in range(0, 3072, 1024)]
    engine.check_status()
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "disp_map"):
        got = torch.cat([p[k] for p in parts])
        assert float((full[k] - got).abs().max()) <= 2e-5, k
    assert float(full["acc_map"].max()) > 0.5                                    # the rays do hit the body
    # training kernels: dumps row for row
    ret, acts = engine.render_train(rays, sk, cy, nanfill_chunk=1024)
    prt = [engine.render_train(rays[i:i + 1024], sk, cy, nanfill_chunk=1024) for i in range(0, 3072, 1024)]
    engine.check_status()
    for k in ("raw", "raw0", "z_fine"):
        assert torch.equal(ret[k], torch.cat([p[0][k] for p in prt])), k
    for key, s in (("c", 64), ("f", 80)):
        for l in (0, 4, 7, 8):
            want = torch.cat([act_layer(p[1][key], l, 1024 * s) for p in prt])
            assert torch.equal(act_layer(acts[key], l, 3072 * s), want), (key, l)
        mask, rows = act_masks(acts[key])
        bits = mask.view(torch.int32).view(8, 8, rows)[:, :, :3072 * s]
        want = torch.cat([act_masks(p[1][key])[0].view(torch.int32).view(8, 8, -1)[:, :, :1024 * s] for p in prt], dim=2)
        assert torch.equal(bits, want), key
