"""How fast can the SMs write into L2? (fill_ / copy_ of L2-sized tensors, CUDA events, torch kernels.)
The GEMM epilogues of the wide-output shapes (q/k/v: 19.7 MB, FF1: 26 MB per launch) are bounded by this number."""
import torch

def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

for mb in (2, 5, 10, 20, 40, 80, 160, 640):
    x = torch.empty(mb * 1024 * 1024 // 2, device="cuda", dtype=torch.bfloat16)
    y = torch.empty_like(x)
    us_fill = t(lambda: x.fill_(1.0))
    us_copy = t(lambda: y.copy_(x))
    print("%4d MB: fill %7.1f us = %6.2f TB/s written | copy %7.1f us = %6.2f TB/s read + %6.2f TB/s written" %
          (mb, us_fill, mb * 1.048576 / us_fill, us_copy, mb * 1.048576 / us_copy, mb * 1.048576 / us_copy))
