"""clock64 timeline of the dK/dV kernel row thread 0. The stamps are not compiled into the shipped kernel: re-apply them
(see the git history of attention_bwd.cu around the double-buffering change) before running this."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _estimator as E  # noqa: E402
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402

L_ = E._lib()
B, L = int(os.environ.get("PROF_B", "32")), int(os.environ.get("PROF_L", "200"))
dt = torch.bfloat16
qkv = torch.randn(B, L, 1600, device="cuda").to(dt)
mask = torch.ones(B, L, device="cuda")
o = torch.empty(B, L, 512, device="cuda", dtype=dt)
lse = torch.empty(B, 8, L, device="cuda")
dout = torch.randn(B, L, 512, device="cuda").to(dt)
dqkv = torch.empty(B, L, 1536, device="cuda", dtype=dt)
delta = torch.empty(B, 8, L, device="cuda")
kmax = torch.zeros(L_.cvflow_attention_scratch_ints(B, L), dtype=torch.int32, device="cuda")
code = N.dtype_code(dt)
N.check(L_.cvflow_attention_forward(qkv.data_ptr(), 1600, B, L, code, mask.data_ptr(), kmax.data_ptr(), 0, o.data_ptr(), lse.data_ptr(), E._stream()))
dbg = torch.zeros(148 * 32, device="cuda", dtype=torch.int64)
for it in range(3):
    if it == 2:
        L_.cvflow_debug_attention_stamps(C.c_void_p(dbg.data_ptr()))
    N.check(L_.cvflow_attention_backward(qkv.data_ptr(), 1600, B, L, code, mask.data_ptr(), kmax.data_ptr(), 0, o.data_ptr(), lse.data_ptr(),
                                         dout.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), E._stream()))
torch.cuda.synchronize()
L_.cvflow_debug_attention_stamps(None)
t = dbg.view(-1, 32).cpu().double()
t = t[t[:, 0] > 0]
n = int((t[0] > 0).sum())
print("B=%d L=%d: %d CTAs, %d stamps; clocks since the row thread entered its loop (mean over CTAs)" % (B, L, t.shape[0], n))
print("  " + " ".join("%6.0f" % float((t[:, k] - t[:, 0]).mean()) for k in range(n)))
print("  per block: [stats staged+bar] [S^T/dP^T ready] [P^T/dS^T written]; then [accumulators ready] ... [loop exit]")
