"""Role-level cycle counters of pgn_delta_chain_kernel (CTA 0).  Needs a library built with -DPGN_CHAIN_PROF
(add it to NVCC_FLAGS in posegen_b200/build.py, `python -m posegen_b200.build --force`): python tools/chain_prof.py"""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from posegen_b200 import synthetic as syn, _lib
from posegen_b200.engine import Engine
from posegen_b200.train import chain_wstream
eng = Engine(); dev = torch.device('cuda')
P = {k: torch.as_tensor(v, device=dev) for k, v in syn.synthetic_nerf_state(7).items()}
m = 3072 * 80; rows = m
dG = (torch.randn((m, 128), device=dev) * 0.1).to(torch.bfloat16)
d_raw = torch.randn((m, 4), device=dev)
mask = torch.randint(-2**31, 2**31 - 1, (8, 8, rows), device=dev, dtype=torch.int32)
ws = chain_wstream(P); wa = P["alpha_linear.weight"].reshape(-1).float().contiguous()
for _ in range(3): eng.mlp_delta_chain(dG, d_raw, mask, rows, ws, wa)
torch.cuda.synchronize()
buf = (C.c_uint64 * 16)()
eng.lib.pgn_debug_chain_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lm = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0xFF
e0.record(); eng.mlp_delta_chain(dG, d_raw, mask, rows, ws, wa, layer_mask=lm); e1.record(); torch.cuda.synchronize()
eng.lib.pgn_debug_chain_prof(buf, 0)
names = ["issuer wait act_ready", "issuer wait w_full", "store-grp wait cs_ready", "store-grp work", "epi wait acc_full", "epi wait cs_done", "epi drain", "issuer total", "epi total", "store-grp total"]
print("kernel us", e0.elapsed_time(e1) * 1e3, "block-layers per CTA", (m // 256 + 147) // 148 * 8)
for n, v in zip(names, buf): print(f"{n:26s} {v:12d} cycles")
