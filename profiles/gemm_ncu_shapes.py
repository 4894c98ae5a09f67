"""One launch of the GEMM engine per distinct (M, N, K, epilogue) shape of the 32 x 400 training step, for
`ncu --set full -k regex:gemm_tc` (the per-shape counter table of profiles/r02_gemm_shape_table.txt):

  python profiles/gemm_ncu_shapes.py                      # plain run first (must exit 0)
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -o gpurun_out/r02_gemm_shapes \
      python profiles/gemm_ncu_shapes.py

Every shape is launched twice (warm-up, then the launch to read): the table takes the second launch of each pair.
Operands are fresh tensors per shape (L2 holds them after the warm-up launch, as inside the step).
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_lora_finetune_framework_b200 import _native as N  # noqa: E402
from tests.test_gemm_gpu import _desc  # noqa: E402

dt = torch.bfloat16
L = N.lib()

SHAPES = [  # name, N, K, epilogue mode
    ("qkv_fwd(+lora u)", 1600, 256, "h16"),
    ("out_proj_fwd(+resid)", 256, 512, "resid"),
    ("ff1_fwd(gelu+stash)", 1024, 256, "gelu"),
    ("ff2_fwd(+resid)", 256, 1024, "resid"),
    ("ff2_dgrad(*gelu')", 1024, 256, "mulgrad"),
    ("ff1_dgrad", 256, 1024, "h16"),
    ("out_proj_dgrad", 512, 256, "h16"),
    ("qkv_dgrad(+lora v)", 320, 1536, "h16"),
    ("conv_k3(256)", 256, 768, "h16"),
]


def launch(name, M, Nn, K, mode):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
    W = (torch.randn(Nn, K, device="cuda") * 0.1).to(dt)
    bias = torch.randn(Nn, device="cuda")
    kw = dict(bias=bias)
    if mode == "resid":
        out = torch.randn(M, Nn, device="cuda")
        kw.update(resid=out, ldr=Nn)
    else:
        out = torch.empty(M, Nn, device="cuda", dtype=dt)
    if mode == "gelu":
        aux = torch.empty(M, Nn, device="cuda", dtype=dt)
        kw.update(act=N.ACT_GELU_TANH, aux_out=aux, ld_aux=Nn)
    if mode == "mulgrad":
        aux = torch.randn(M, Nn, device="cuda").to(dt)
        kw = dict(act=N.ACT_MUL_GELU_TANH_GRAD, mul_src=aux, ld_aux=Nn)
    d = _desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, dtype=dt, **kw)
    st = N.current_stream()
    for _ in range(2):
        N.check(L.cvflow_gemm(C.byref(d), st), "cvflow_gemm")
    torch.cuda.synchronize()
    print("%-24s M=%6d N=%5d K=%5d" % (name, M, Nn, K))


for M in (6400, 12800):
    for name, Nn, K, mode in SHAPES:
        launch(name, M, Nn, K, mode)
