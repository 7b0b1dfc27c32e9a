import torch
x = torch.empty(2 * 1024**3 // 4, device='cuda', dtype=torch.float32)
y = torch.empty_like(x)
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b = x.numel() * 4
print("fill_ GB/s", b / t(lambda: x.fill_(1.0)) / 1e6)
print("zero_ GB/s", b / t(lambda: x.zero_()) / 1e6)
print("copy_ GB/s (r+w)", 2 * b / t(lambda: y.copy_(x)) / 1e6)
print("sum GB/s (read)", b / t(lambda: x.sum()) / 1e6)
