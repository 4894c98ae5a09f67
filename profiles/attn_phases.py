"""Phase timeline of the attention forward kernel (in-kernel globaltimer stamps written by row thread 0 of
every persistent CTA for its first two items), standalone launch through the C ABI."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _estimator as E  # noqa: E402
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402

L_ = E._lib()


def hot(ms=300):
    """Sustained load right before a measurement so the SM clock is at its boost level, as inside a training step."""
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while True:
        for _ in range(10):
            a @ a
        e1.record()
        e1.synchronize()
        if e0.elapsed_time(e1) > ms:
            break

B, L = int(os.environ.get("PROF_B", "32")), int(os.environ.get("PROF_L", "200"))
dt = torch.bfloat16
dbg = torch.zeros(148 * 32, device="cuda", dtype=torch.int64)
L_.cvflow_debug_attention_stamps(C.c_void_p(dbg.data_ptr()))
qkv = torch.randn(B, L, 1600, device="cuda").to(dt)
mask = torch.ones(B, L, device="cuda")
o = torch.empty(B, L, 512, device="cuda", dtype=dt)
lse = torch.empty(B, 8, L, device="cuda")
kmax = torch.zeros(E._lib().cvflow_attention_scratch_ints(B, L), dtype=torch.int32, device="cuda")
for _ in range(3):
    dbg.zero_()
    N.check(L_.cvflow_attention_forward(qkv.data_ptr(), 1600, B, L, N.dtype_code(dt), mask.data_ptr(), kmax.data_ptr(), 0,
                                        o.data_ptr(), lse.data_ptr(), E._stream()))
torch.cuda.synchronize()
L_.cvflow_debug_attention_stamps(None)
t = dbg.view(-1, 32).cpu().double()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = {1: "prologue (barriers, TMEM alloc)", 2: "pdl wait", 17: "MMA0: Q landed (item 0)", 18: "MMA0: K/V landed (item 0)",
         3: "S ready (item 0)", 4: "pass 1 (max) done", 5: "pass 2 (exp, P stored)", 6: "O' ready [PV MMA]", 7: "O' read", 8: "output stored",
         9: "S ready (item 1)", 10: "pass 1 done", 11: "pass 2 done", 12: "O' ready", 13: "O' read", 14: "output stored", 15: "exit"}
print("SM clock during the kernel: %.0f MHz" % float(((t[:, 31] - t[:, 30]) / (t[:, 15] - t[:, 0])).mean() * 1e3))
print("B=%d L=%d: %d CTAs, kernel span %.1f us, CTA lifetime mean %.2f us" %
      (B, L, t.shape[0], float(t[:, 15].max() - t0) / 1e3, float((t[:, 15] - t[:, 0]).mean()) / 1e3))
for k in (1, 2, 17, 18, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15):
    sel = t[:, k] > 0
    if sel.any():
        print("  %-34s at %6.2f us (n=%d)" % (names[k], float((t[sel, k] - t[sel, 0]).mean()) / 1e3, int(sel.sum())))

