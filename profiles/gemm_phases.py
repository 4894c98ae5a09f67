"""Per-CTA phase timeline of the GEMM kernel from in-kernel globaltimer stamps (profiling aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402
from tests.test_gemm_gpu import _desc  # noqa: E402

dt = torch.bfloat16
L = N.lib()
names = ["start", "grid-dependency wait passed", "first stage landed (MMA warp)", "last tile's MMAs issued", "first accumulator ready (epilogue)",
         "last epilogue done", "exit"]
for (M, Nn, K, mode) in [(6400, 1536, 256, "h16"), (6400, 256, 512, "resid"), (6400, 1024, 256, "gelu"), (6400, 256, 1024, "resid"), (6400, 1024, 256, "mulgrad"), (6400, 256, 1024, "h16"), (6400, 512, 256, "h16"), (6400, 256, 1536, "h16"), (6400, 64, 1536, "h16"), (12800, 1536, 256, "h16"), (12800, 256, 1024, "resid"),
                        (6400, 256, 512, "resid+ln"), (6400, 256, 1024, "resid+ln")]:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    W = (torch.randn(Nn, K, device="cuda") * 0.1).to(dt)
    out = torch.randn(M, Nn, device="cuda") if mode.startswith("resid") else torch.empty(M, Nn, device="cuda", dtype=dt)
    kw = {}
    if mode.startswith("resid"):
        kw = dict(resid=out, ldr=Nn, bias=torch.randn(Nn, device="cuda"))
    if mode == "resid+ln":      # LayerNorm fused into the row-owning 128x256 epilogue
        kw.update(ln_gamma=torch.ones(Nn, device="cuda"), ln_beta=torch.zeros(Nn, device="cuda"),
                  aux_out=torch.empty(M, Nn, device="cuda", dtype=dt), ld_aux=Nn)
    if mode == "gelu":
        kw = dict(act=N.ACT_GELU_TANH, aux_out=torch.empty(M, Nn, device="cuda", dtype=dt), ld_aux=Nn, bias=torch.randn(Nn, device="cuda"))
    if mode == "mulgrad":
        kw = dict(act=N.ACT_MUL_GELU_TANH_GRAD, mul_src=torch.randn(M, Nn, device="cuda").to(dt), ld_aux=Nn)
    dbg = torch.zeros(4096 * 16, device="cuda", dtype=torch.int64)
    d = _desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, dtype=dt, **kw)
    d.dbg = dbg.data_ptr()
    for _ in range(3):
        L.cvflow_gemm(C.byref(d), N.current_stream())
    torch.cuda.synchronize()
    dbg.zero_()
    L.cvflow_gemm(C.byref(d), N.current_stream())
    torch.cuda.synchronize()
    t = dbg.view(-1, 16).cpu()
    t = t[t[:, 0] > 0].double()
    MHZ = 1965.0                                  # SM clock under load on these boxes (bench.py's clocks line)
    t0 = t[:, 0].min()
    start_us = (t[:, 0] - t0) / 1e3               # CTA start, us after the first CTA (globaltimer)
    cyc = lambda k: (t[:, k] - t[:, 14]) / MHZ     # us since the CTA's own start (cycle counter)
    end_us = start_us + cyc(6)
    span = float(end_us.max())
    fl = 2.0 * M * Nn * K
    print("== M=%d N=%d K=%d %s: %d persistent CTAs, kernel span %.1f us (%.0f TFLOP/s), CTA start spread %.2f us, "
          "CTA lifetime mean %.2f / max %.2f us" % (M, Nn, K, mode, t.shape[0], span, fl / span / 1e6, float(start_us.max()),
                                                   float(cyc(6).mean()), float(cyc(6).max())))
    # mean / max over CTAs of every stamp, us after the CTA's own start; CTAs grouped by the number of tiles they ran
    for nt in sorted(set(t[:, 7].tolist())):
        sel = t[:, 7] == nt
        print("   %3d CTAs with %d tile(s): " % (int(sel.sum()), int(nt)) + " | ".join(
            "%s %.2f/%.2f" % (names[k].split(" (")[0], float(cyc(k)[sel].mean()), float(cyc(k)[sel].max())) for k in range(1, 7)))
        e = {8: "chunk0 tmem-ld done", 9: "chunk0 math done", 10: "chunk0 stored", 11: "chunk1 tmem-ld done", 12: "chunk1 stored",
             13: "tile0 acc released"}
        print("        first tile's epilogue (warp 2), us after its accumulator was ready: " + " | ".join(
            "%s %.2f" % (e[k], float(((t[:, k] - t[:, 4]) / MHZ)[sel & (t[:, k] > 0)].mean())) for k in sorted(e)
            if (sel & (t[:, k] > 0)).any()))
