"""GPU: the differentiable frame render of the GAN step (BASELINE.json configs[4], posegen_b200/gan.py): masks-only
forward and input-gradient backward to the bone transforms against the training path (itself checked against oracle autograd in
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
    # (a) frame path: eval-mode forward that keeps the ReLU masks, chunked input-gradient backward over the live rays
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
    assert float((ga - gb).norm() / gb.norm()) <= 5e-3, float((ga - gb).norm() / gb.norm())
    assert float(ga[:, 3].abs().max()) == 0.0                            # bottom rows carry no gradient


def test_masks_only_forward_matches_the_full_dump(rc):
    """pgn_render_forward_masks keeps exactly the ReLU masks of the full activation dump (trunk bits, view layer = [g > 0])
    and renders the same values."""
    from posegen_b200.train import act_layer, act_masks
    frame = syn.synthetic_frame(3, 64, 64)
    dev = torch.device("cuda")
    n = 203
    rb = torch.as_tensor(syn.ray_batch(frame.rays_o, frame.rays_d)[:n], device=dev)
    sk = torch.as_tensor(frame.pose.skts, dtype=torch.float32, device=dev)
    cy = torch.as_tensor(frame.pose.cyl, dtype=torch.float32, device=dev)
    eng = rc.engine(dev)
    ret_m, (trunk, view) = eng.render_masks(rb, sk, cy)
    ret_d, acts = eng.render_train(rb, sk, cy, dump_coarse=False)
    eng.check_status()
    m = n * 80
    # per-sample network outputs and sample positions are identical; the composited maps agree up to the association order
    # of the fine-pass compositing (the two forwards cut this small batch into groups of 8 and of 4 rays)
    for k in ("raw", "z_fine"):
        assert torch.equal(ret_m[k], ret_d[k]), k
    for k in ("rgb_map", "acc_map"):
        assert float((ret_m[k] - ret_d[k]).abs().max()) <= 1e-5, k
    full, rows = act_masks(acts["f"])
    assert torch.equal(trunk[:, :, :m], full.view(torch.int32).view(8, 8, rows)[:, :, :m])
    bits = ((view[:, :m].t()[:, :, None] >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(m, 128).bool()
    assert torch.equal(bits, act_layer(acts["f"], 8, m) > 0)


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


@pytest.mark.gpu
def test_gather_ray_rows_matches_index_select(engine):
    """pgn_gather_ray_rows (the GAN backward's ray selection) against torch.index_select, single- and multi-plane."""
    dev = engine.device
    g = torch.Generator(device="cpu").manual_seed(3)
    idx = torch.randperm(5000, generator=g)[:1234].to(dev)
    z = torch.randn((5000, 80), generator=g).to(dev)
    assert torch.equal(engine.gather_ray_rows(z, idx), z.index_select(0, idx))
    raw = torch.randn((5000, 80, 4), generator=g).to(dev)
    assert torch.equal(engine.gather_ray_rows(raw, idx), raw.index_select(0, idx))
    planes = torch.randint(-2**31, 2**31 - 1, (8, 5000, 640), generator=g, dtype=torch.int32).to(dev)
    assert torch.equal(engine.gather_ray_rows(planes, idx, n_planes=8), planes.index_select(1, idx))
    assert engine.gather_ray_rows(z, idx[:0]).shape == (0, 80)
    with pytest.raises(Exception):
        engine.gather_ray_rows(torch.zeros((10, 11), device=dev), idx[:3] % 10)          # 44-byte rows: not a multiple of 16
    bad = engine.gather_ray_rows(z, torch.tensor([1, 5000, 7], device=dev))               # 5000 is out of range
    torch.cuda.synchronize()
    assert torch.equal(bad[0], z[1]) and torch.equal(bad[2], z[7]) and float(bad[1].abs().max()) == 0.0
    with pytest.raises(Exception, match="out of range"):
        engine.check_status()
    engine.check_status()                                                                  # (the latch is cleared by the report)
