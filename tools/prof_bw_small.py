"""How much HBM bandwidth does a single short kernel reach?  (context for the GN roofline: ~200 MB per launch)"""
import torch

def t(fn, reps=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e-3

for mb in (50, 100, 200, 400, 800, 2000):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    s = t(lambda: y.copy_(x))
    r = t(lambda: x.sum())
    print("%5d MB: copy %.1f us = %.0f GB/s (r+w) | sum %.1f us = %.0f GB/s (read)" % (mb, s * 1e6, 2 * 4 * n / s / 1e9, r * 1e6, 4 * n / r / 1e9))
