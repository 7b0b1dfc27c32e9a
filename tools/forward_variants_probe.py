"""Diagnostic: device time of the fused forward on ONE training-shaped batch (3,072 rays from 256 poses x 12 rays) in its
variants - inference kernel vs training kernel (activation + mask dump), per-ray poses vs one shared pose - to separate
the cost of the dump from the cost of the small batch.  python tools/forward_variants_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posegen_b200 import synthetic as syn                      # noqa: E402
from posegen_b200.engine import Engine                         # noqa: E402
from tools.train_step_bench import make_batch                  # noqa: E402

eng = Engine()
eng.load_checkpoint(syn.synthetic_raycaster_state(4, alpha_gain=40.0))
dev = eng.device
rb, sk, cy = (torch.as_tensor(x, device=dev) for x in make_batch(0))


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n in (3072, 2048, 1024, 6144):
    r, s, c = rb[:n].contiguous() if n <= rb.shape[0] else rb.repeat(2, 1), sk[:n].contiguous() if n <= sk.shape[0] else sk.repeat(2, 1, 1, 1), \
        cy[:n].contiguous() if n <= cy.shape[0] else cy.repeat(2, 1)
    t_inf = timed(lambda: eng.render(r, s, c, nanfill_chunk=n, return_alpha=False))
    t_inf_shared = timed(lambda: eng.render(r, s[0].contiguous(), c[0].contiguous(), nanfill_chunk=n, return_alpha=False))
    t_train = timed(lambda: eng.render_train(r, s, c, nanfill_chunk=n))
    t_train_f = timed(lambda: eng.render_train(r, s, c, nanfill_chunk=n, dump_coarse=False))
    print(f"n={n}: inference {t_inf:.3f} ms (shared pose {t_inf_shared:.3f}), training forward {t_train:.3f} ms (fine dump only {t_train_f:.3f})", flush=True)
eng.check_status()
