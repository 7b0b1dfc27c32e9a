"""GPU: the differentiable frame render of the GAN step (BASELINE.json configs[4], posegen_b200/gan.py): chunked
recompute backward to the bone transforms against the training path (itself checked against oracle autograd in
test_gpu_train.py), the white-background composite against pgn_compose_frame, and the adjoint of the HMR hand-off."""
import numpy as np
import pytest
import torch

from posegen_b200 import gan, synthetic as syn
from posegen_b200.raycaster import raycaster_from_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rc():
    m = raycaster_from_checkpoint(syn.synthetic_raycaster_state(4, alpha_gain=40.0), device="cuda", precision="bf16")
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def test_frame_render_pose_gradient_matches_training_path(rc):
    frame = syn.synthetic_frame(3, 64, 64)
    dev = torch.device("cuda")
    rb = torch.as_tensor(syn.ray_batch(frame.rays_o, frame.rays_d), device=dev)
    n = rb.shape[0]
    cy = torch.as_tensor(frame.pose.cyl, dtype=torch.float32, device=dev)
    g = torch.Generator(device="cuda").manual_seed(0)
    w_rgb = torch.randn((n, 3), device=dev, generator=g)
    w_acc = torch.randn((n,), device=dev, generator=g)
    dead = torch.rand((n,), device=dev, generator=g) < 0.3            # rays without upstream gradient are skipped
    w_rgb[dead] = 0
    w_acc[dead] = 0
    # (a) frame path: eval-mode forward, chunked recompute in the backward
    rc.eval()
    sk_a = torch.as_tensor(frame.pose.skts, dtype=torch.float32, device=dev).requires_grad_(True)
    rgb, acc = gan.render_frame(rc, rb, sk_a, cy, chunk=1000)
    ((rgb * w_rgb).sum() + (acc * w_acc).sum()).backward()
    # (b) training path on the same rays (activation dump of the whole batch, per-ray skts gradient summed)
    rc.train()
    sk_b = torch.as_tensor(frame.pose.skts, dtype=torch.float32, device=dev).requires_grad_(True)
    ret = rc(rb, N_samples=64, N_importance=16, kp_batch=None, skts=sk_b, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
    ((ret["rgb_map"] * w_rgb).sum() + (ret["acc_map"] * w_acc).sum()).backward()
    rc.eval()
    assert torch.equal(rgb.detach(), ret["rgb_map"].detach()) and torch.equal(acc.detach(), ret["acc_map"].detach())
    ga, gb = sk_a.grad.double(), sk_b.grad.double()
    assert torch.isfinite(ga).all() and float(ga.abs().max()) > 0
    assert float((ga - gb).norm() / gb.norm()) <= 2e-3, float((ga - gb).norm() / gb.norm())
    assert float(ga[:, 3].abs().max()) == 0.0                            # bottom rows carry no gradient
    # (c) frame path with the live-ray hint: dumps kept in the forward (the budget holds one of the three chunks, the
    # others fall back to the recompute), same image, same gradient
    rc.eval()
    sk_c = torch.as_tensor(frame.pose.skts, dtype=torch.float32, device=dev).requires_grad_(True)
    one_chunk = 1000 * 80 * 4608
    rgb_c, acc_c = gan.render_frame(rc, rb, sk_c, cy, chunk=1000, live=~dead, dump_budget_bytes=one_chunk + 1024)
    ((rgb_c * w_rgb).sum() + (acc_c * w_acc).sum()).backward()
    # a ray that hits the cylinder renders to the same value whatever batch it is in, up to the association order of its
    # fine-pass compositing (a ray's 80 samples straddle 128-row tiles differently at another position in its group of 8);
    # the bbox corners that miss the cylinder take the mean near/far of their batch (the reference's chunk-level
    # nan-mean, ray_utils.py:328-342), which depends on how the frame is split
    o_xz, d_xz = rb[:, [0, 2]], rb[:, [3, 5]]
    to_c = cy[None, :2] - o_xz
    dist = (to_c[:, 0] * d_xz[:, 1] - to_c[:, 1] * d_xz[:, 0]).abs() / d_xz.norm(dim=-1)
    hit = dist < cy[2] * (1 - 1e-4)
    assert float(hit.float().mean()) > 0.3
    assert float((rgb_c.detach() - rgb.detach())[hit].abs().max()) <= 1e-5
    assert float((acc_c.detach() - acc.detach())[hit].abs().max()) <= 1e-5
    assert float((rgb_c.detach() - rgb.detach()).abs().max()) <= 1e-1
    gc = sk_c.grad.double()
    assert float((gc - ga).norm() / ga.norm()) <= 5e-2, float((gc - ga).norm() / ga.norm())


def test_compose_white_matches_the_device_composite(rc):
    eng = rc.engine(torch.device("cuda"))
    H, W, x0, y0, x1, y1 = 40, 48, 5, 7, 33, 29
    n = (x1 - x0) * (y1 - y0)
    rgb = torch.rand((n, 3), device="cuda")
    acc = torch.rand((n,), device="cuda")
    assert torch.equal(gan.compose_white(rgb, acc, H, W, x0, y0, x1, y1, 1.0), eng.compose_frame(H, W, x0, y0, x1, y1, rgb, acc, 1.0))


def test_hmr_input_backward_is_the_adjoint(rc):
    eng = rc.engine(torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.rand((96, 96, 3), device="cuda", generator=g)
    v = torch.randn((96, 96, 3), device="cuda", generator=g)
    crop, R = (10, 12, 74, 76), 24
    x = img.clone().requires_grad_(True)
    out = gan.hmr_input(eng, x, crop=crop, out_res=R, quantize_u8=False)
    gout = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(gout)
    jv = gan.hmr_input(eng, img + v, crop=crop, out_res=R, quantize_u8=False) - gan.hmr_input(eng, img, crop=crop, out_res=R, quantize_u8=False)
    lhs, rhs = float((gout * jv).sum()), float((x.grad * v).sum())
    assert abs(lhs - rhs) <= 2e-3 * max(1.0, abs(lhs)), (lhs, rhs)
    outside = x.grad.clone()
    outside[crop[1]:crop[3], crop[0]:crop[2]] = 0
    assert float(outside.abs().max()) == 0.0


def test_pose_images_backpropagate_to_the_bones(rc):
    dev = torch.device("cuda")
    rng = np.random.RandomState(5)
    bones = torch.tensor(rng.randn(1, 24, 3) * 0.3, dtype=torch.float32, device=dev, requires_grad=True)
    rest = torch.as_tensor(syn.SMPL_REST_POSE * syn.BODY_SCALE, dtype=torch.float32, device=dev)
    frames, kps = gan.render_pose_images(rc, bones, rest, syn.run_gan_c2w(), 64, 64, 125.0, chunk=2048)
    assert frames.shape == (1, 64, 64, 3) and kps.shape == (1, 24, 3)
    eng = rc.engine(dev)
    x = gan.hmr_input(eng, frames[0], crop=(12, 12, 52, 52), out_res=28)
    (x ** 2).sum().backward()
    assert torch.isfinite(bones.grad).all() and float(bones.grad.abs().max()) > 0
