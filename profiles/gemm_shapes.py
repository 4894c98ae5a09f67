"""Time the tcgen05 GEMM engine on the estimator's GEMM shapes (isolated launches, CUDA events,
L2-cold between reps by cycling through several operand copies) next to torch.matmul (cuBLAS)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402
from tests.test_gemm_gpu import _desc  # noqa: E402

dt = torch.bfloat16
L = N.lib()


def bench(fn, reps=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def run(name, M, Nn, K, mode):
    ncopy = 6
    As = [(torch.randn(M, K, device="cuda") * 0.5).to(dt) for _ in range(ncopy)]
    W = (torch.randn(Nn, K, device="cuda") * 0.1).to(dt)
    bias = torch.randn(Nn, device="cuda")
    kw = {}
    if mode == "h16":
        outs = [torch.empty(M, Nn, device="cuda", dtype=dt) for _ in range(ncopy)]
    elif mode == "resid":
        outs = [torch.randn(M, Nn, device="cuda") for _ in range(ncopy)]
    elif mode == "gelu":
        outs = [torch.empty(M, Nn, device="cuda", dtype=dt) for _ in range(ncopy)]
        aux = [torch.empty(M, Nn, device="cuda", dtype=dt) for _ in range(ncopy)]
    elif mode == "mulgrad":
        outs = [torch.empty(M, Nn, device="cuda", dtype=dt) for _ in range(ncopy)]
        aux = [torch.randn(M, Nn, device="cuda").to(dt) for _ in range(ncopy)]
    descs = []
    for i in range(ncopy):
        kw = dict(bias=bias)
        if mode == "resid":
            kw.update(resid=outs[i], ldr=Nn)
        if mode == "gelu":
            kw.update(act=N.ACT_GELU_TANH, aux_out=aux[i], ld_aux=Nn)
        if mode == "mulgrad":
            kw = dict(act=N.ACT_MUL_GELU_TANH_GRAD, mul_src=aux[i], ld_aux=Nn)
        descs.append(_desc(As[i], W, outs[i], segs=[(0, 0, 0, K // 64)], R=M, dtype=dt, **kw))
    st = N.current_stream()
    us = bench(lambda i: L.cvflow_gemm(C.byref(descs[i % ncopy]), st))
    us_t = bench(lambda i: torch.matmul(As[i % ncopy], W.t()))
    fl = 2.0 * M * Nn * K
    print("%-28s M=%6d N=%5d K=%5d  cvflow %7.1f us %6.0f TF/s | cuBLAS(plain) %7.1f us %6.0f TF/s" %
          (name, M, Nn, K, us, fl / us / 1e6, us_t, fl / us_t / 1e6))


for M in (6400, 12800):
    run("qkv fwd", M, 1536, 256, "h16")
    run("out-proj fwd (+resid)", M, 256, 512, "resid")
    run("ff1 fwd (gelu+stash)", M, 1024, 256, "gelu")
    run("ff2 fwd (+resid)", M, 256, 1024, "resid")
    run("ff2 dgrad (*gelu')", M, 1024, 256, "mulgrad")
    run("ff1 dgrad", M, 256, 1024, "h16")
    run("out-proj dgrad", M, 512, 256, "h16")
    run("qkv dgrad", M, 256, 1536, "h16")
    run("lora u / v", M, 64, 1536, "h16")


def run_mlp(M):
    """fused FeedForward (one launch) next to the two engine GEMMs it replaces"""
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E
    Lm = E._lib()
    ncopy = 6
    xs = [torch.randn(M, 256, device="cuda").to(dt) for _ in range(ncopy)]
    w1 = (torch.randn(1024, 256, device="cuda") * 0.08).to(dt)
    w2 = (torch.randn(256, 1024, device="cuda") * 0.05).to(dt)
    b1, b2 = torch.randn(1024, device="cuda"), torch.randn(256, device="cuda")
    res = [torch.randn(M, 256, device="cuda") for _ in range(ncopy)]
    outs = [torch.empty(M, 256, device="cuda") for _ in range(ncopy)]
    pres = [torch.empty(M, 1024, device="cuda", dtype=dt) for _ in range(ncopy)]
    dxs = [torch.empty(M, 256, device="cuda", dtype=dt) for _ in range(ncopy)]
    st = E._stream()
    code = N.dtype_code(dt)

    def fwd(i):
        j = i % ncopy
        Lm.cvflow_mlp_forward(xs[j].data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), res[j].data_ptr(),
                              outs[j].data_ptr(), pres[j].data_ptr(), M, code, 0, st)

    def bwd(i):
        j = i % ncopy
        Lm.cvflow_mlp_backward(xs[j].data_ptr(), w1.data_ptr(), pres[j].data_ptr(), w2.data_ptr(), dxs[j].data_ptr(), M, code, 0, st)

    for j in range(ncopy):
        fwd(j)
    fl = 2.0 * M * 256 * 1024 * 2
    uf, ub = bench(fwd), bench(bwd)
    print("fused mlp M=%6d  fwd %7.1f us %6.0f TF/s | bwd %7.1f us %6.0f TF/s" % (M, uf, fl / uf / 1e6, ub, fl / ub / 1e6))


for M in (700, 1400, 6400, 12800, 24000):
    run_mlp(M)
