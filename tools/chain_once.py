"""One pgn_mlp_delta_chain call on a fine-pass-sized batch (for ncu captures): python tools/chain_once.py [layer_mask]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from posegen_b200 import synthetic as syn
from posegen_b200.engine import Engine
from posegen_b200.train import chain_wstream
eng = Engine(); dev = torch.device('cuda')
P = {k: torch.as_tensor(v, device=dev) for k, v in syn.synthetic_nerf_state(7).items()}
m = 3072 * 80; rows = m
lm = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0xFF
dG = (torch.randn((m, 128), device=dev) * 0.1).to(torch.bfloat16)
d_raw = torch.randn((m, 4), device=dev)
mask = torch.randint(-2**31, 2**31 - 1, (8, 8, rows), device=dev, dtype=torch.int32)
ws = chain_wstream(P); wa = P["alpha_linear.weight"].reshape(-1).float().contiguous()
for _ in range(2): eng.mlp_delta_chain(dG, d_raw, mask, rows, ws, wa, layer_mask=lm)
torch.cuda.synchronize()
eng.check_status()
