import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posegen_b200.engine import Engine
eng = Engine()
torch.manual_seed(0)
for (K, N) in [(16, 256), (64, 256), (256, 256), (144, 256), (112, 128), (256, 128)]:
    A = torch.randn(256, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    D = eng.debug_umma_gemm(A, B, 2)
    torch.cuda.synchronize()
    try:
        eng.check_status(); st = "ok"
    except Exception as e:
        st = str(e)
    err = (D - ref).abs()
    print(f"pair probe K={K} N={N}: max_err {float(err.max()):.3e} (rows0-127 {float(err[:128].max()):.3e}, rows128-255 {float(err[128:].max()):.3e}) ref_max {float(ref.abs().max()):.2f} {st}", flush=True)
