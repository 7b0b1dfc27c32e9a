import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posegen_b200.engine import Engine
eng = Engine()
torch.manual_seed(0)
for variant, rows, name in ((4, 128, "1cta M=128"), (6, 256, "2cta M=256")):
    for (K, N) in [(256, 256), (256, 128), (128, 256)]:
        A = torch.randn(rows, K, device="cuda"); B = torch.randn(N, K, device="cuda")
        for it in range(3):
            D = eng.debug_umma_gemm(A, B, variant)
            torch.cuda.synchronize()
        eng.check_status()
        n_mma = 8 * K // 16
        issue, total = float(D[rows, 0]), float(D[rows, 1])
        print(f"{name} K={K} N={N}: {n_mma} MMAs issue {issue:.0f} cyc ({issue/n_mma:.1f}/MMA), until complete {total:.0f} cyc ({total/n_mma:.1f}/MMA)", flush=True)
