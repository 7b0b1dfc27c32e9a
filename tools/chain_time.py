"""Time pgn_mlp_delta_chain alone on a fine-pass-sized batch (245,760 rows): python tools/chain_time.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from posegen_b200 import synthetic as syn
from posegen_b200.engine import Engine
from posegen_b200.train import chain_wstream
eng = Engine(); dev = torch.device('cuda')
P = {k: torch.as_tensor(v, device=dev) for k, v in syn.synthetic_nerf_state(7).items()}
m = 3072 * 80; rows = m
dG = (torch.randn((m, 128), device=dev) * 0.1).to(torch.bfloat16)
d_raw = torch.randn((m, 4), device=dev)
mask = torch.randint(-2**31, 2**31 - 1, (8, 8, rows), device=dev, dtype=torch.int32)
ws = chain_wstream(P); wa = P["alpha_linear.weight"].reshape(-1).float().contiguous()
for lm in (0xFF, 0x21, 0x00):
    for _ in range(3): eng.mlp_delta_chain(dG, d_raw, mask, rows, ws, wa, layer_mask=lm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.mlp_delta_chain(dG, d_raw, mask, rows, ws, wa, layer_mask=lm)
    e1.record(); torch.cuda.synchronize()
    print(sys.argv[1:], "layer_mask", hex(lm), "chain us per call:", round(e0.elapsed_time(e1) * 100, 1), "rows", m)
eng.check_status()
