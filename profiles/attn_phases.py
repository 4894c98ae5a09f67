"""Phase timeline of the attention forward kernel (in-kernel globaltimer stamps), tiny estimator."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _estimator as E  # noqa: E402
from tests.helpers import build_estimator  # noqa: E402

L = E._lib()
B, T = 32, int(os.environ.get("PROF_T", "200"))
est, _, _ = build_estimator(1, 0)
est = est.cuda()
est.cvflow_dtype = torch.bfloat16
dbg = torch.zeros(B * 8 * 8 * 16, device="cuda", dtype=torch.int64)
L.cvflow_debug_attention_stamps(C.c_void_p(dbg.data_ptr()))
x = torch.randn(B, 80, T, device="cuda")
with torch.no_grad():
    for _ in range(3):
        est(x, torch.ones(B, 1, T, device="cuda"), x, torch.rand(B, device="cuda"), torch.randn(B, 80, device="cuda"), x)
torch.cuda.synchronize()
# the last attention launched at length T wrote the buffer last (up_blocks.1): analyse it
n_cta = ((T + 127) // 128) * 8 * B
t = dbg.view(-1, 16)[:n_cta].cpu().double()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = {1: "prologue (barriers, TMEM alloc)", 2: "pdl wait", 3: "S(0) ready [TMA Q,K + MMA]", 4: "row max exchanged", 5: "P(0) written",
         6: "O'(0) ready [PV MMA]", 7: "S(1) ready", 8: "row max exchanged", 9: "P(1) written", 10: "O'(1) ready", 11: "output stored", 12: "exit"}
print("T=%d: %d CTAs, kernel span %.1f us, CTA lifetime mean %.2f us" % (T, t.shape[0], float(t[:, 12].max() - t0) / 1e3,
                                                                       float((t[:, 12] - t[:, 0]).mean()) / 1e3))
prev = 0
for k in range(1, 13):
    if (t[:, k] > 0).all():
        print("  %-34s +%.2f us  (at %.2f us)" % (names[k], float((t[:, k] - t[:, prev]).mean()) / 1e3, float((t[:, k] - t[:, 0]).mean()) / 1e3))
        prev = k
