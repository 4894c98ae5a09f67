"""Per-CTA clock64 timeline of the fused feed-forward kernel (profiling aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _estimator as E  # noqa: E402
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402

L = E._lib()
M = int(os.environ.get("PROF_M", "6400"))
bwd = int(os.environ.get("PROF_BWD", "0"))
act = int(os.environ.get("PROF_ACT", "0"))
dt = torch.bfloat16
x = torch.randn(M, 256, device="cuda").to(dt)
w1 = (torch.randn(1024, 256, device="cuda") * 0.08).to(dt)
w2 = (torch.randn(256, 1024, device="cuda") * 0.05).to(dt)
b1, b2 = torch.randn(1024, device="cuda"), torch.randn(256, device="cuda")
res = torch.randn(M, 256, device="cuda")
out = torch.empty(M, 256, device="cuda")
pre = torch.randn(M, 1024, device="cuda").to(dt)
dx = torch.empty(M, 256, device="cuda", dtype=dt)
dbg = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
L.cvflow_debug_mlp_stamps(C.c_void_p(dbg.data_ptr()))
for _ in range(3):
    dbg.zero_()
    if bwd:
        L.cvflow_mlp_backward(x.data_ptr(), w1.data_ptr(), pre.data_ptr(), w2.data_ptr(), dx.data_ptr(), M, N.dtype_code(dt), 0, E._stream())
    else:
        L.cvflow_mlp_forward(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), res.data_ptr(), out.data_ptr(),
                             pre.data_ptr(), M, N.dtype_code(dt), act, E._stream())
torch.cuda.synchronize()
L.cvflow_debug_mlp_stamps(None)
t = dbg.view(-1, 64).cpu().double()
t = t[t[:, 0] > 0]
rel = lambda k: float((t[:, k] - t[:, 0]).mean())
print("M=%d %s: %d CTAs; clocks since CTA start (mean over CTAs)" % (M, "bwd" if bwd else "fwd", t.shape[0]))
print("  X tile landed %.0f" % rel(1))
for s in range(8):
    print("  chunk %d: G1 issued %6.0f | E1 start %6.0f end %6.0f (%5.0f) | MMA saw P ready %6.0f" %
          (s, rel(2 + 2 * s), rel(20 + 2 * s), rel(21 + 2 * s), rel(21 + 2 * s) - rel(20 + 2 * s), rel(3 + 2 * s)))
print("  acc2 ready %.0f, final epilogue done %.0f" % (rel(40), rel(41)))
