"""Time the attn1 kernels on their own (C ABI, CUDA events, rotating operand copies so every launch
reads L2-cold data) at the shapes of the BASELINE configs."""
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

dt = torch.bfloat16


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def run(B, L, ragged, ldq=1600):
    ncopy = 1 if os.environ.get("ATTN_NCU") else 6      # ncu backs device memory up to the host for its replay passes: keep it small
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(int(0.6 * L) + 1, L + 1, (B,), generator=g) if ragged else torch.full((B,), L)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).float().cuda()
    qkv = [(torch.randn(B, L, ldq, device="cuda")).to(dt) for _ in range(ncopy)]
    dout = [(torch.randn(B, L, 512, device="cuda") * mask[:, :, None]).to(dt) for _ in range(ncopy)]
    o = [torch.empty(B, L, 512, device="cuda", dtype=dt) for _ in range(ncopy)]
    lse = [torch.empty(B, 8, L, device="cuda") for _ in range(ncopy)]
    dqkv = torch.empty(B, L, 1536, device="cuda", dtype=dt)
    delta = torch.empty(B, 8, L, device="cuda")
    kmax = torch.zeros(E._lib().cvflow_attention_scratch_ints(B, L), dtype=torch.int32, device="cuda")
    st = E._stream()
    code = N.dtype_code(dt)

    def fwd(i):
        j = i % ncopy
        N.check(L_.cvflow_attention_forward(qkv[j].data_ptr(), ldq, B, L, code, mask.data_ptr(), kmax.data_ptr(), 0,
                                            o[j].data_ptr(), lse[j].data_ptr(), st))

    def bwd(i):
        j = i % ncopy
        N.check(L_.cvflow_attention_backward(qkv[j].data_ptr(), ldq, B, L, code, mask.data_ptr(), kmax.data_ptr(), 0,
                                             o[j].data_ptr(), lse[j].data_ptr(), dout[j].data_ptr(), delta.data_ptr(),
                                             dqkv.data_ptr(), st))

    if os.environ.get("ATTN_NCU"):      # profiling mode: one forward + one backward launch set per shape, in the order printed
        fwd(0); bwd(0)
        torch.cuda.synchronize()
        print("ncu shape: B=%d L=%d %s (8 heads x 64, bf16): attn_kinfo, attn_fwd, attn_kinfo, attn_bwd_dq, attn_bwd_dkv" %
              (B, L, "ragged" if ragged else "full"))
        return
    for j in range(ncopy):
        fwd(j)
    tf, tb = timeit(fwd), timeit(bwd)
    valid = float((lens.double() ** 2).sum()) * 8 * 64
    print("B=%3d L=%5d %-6s fwd %7.1f us (%5.0f TF/s valid)  bwd %7.1f us (%5.0f TF/s valid)  [includes the kmax launch]" %
          (B, L, "ragged" if ragged else "full", tf, 4 * valid / tf / 1e6, tb, 10 * valid / tb / 1e6))


if __name__ == "__main__" and os.environ.get("ATTN_NCU"):
    for B, L in ((32, 200), (32, 400), (16, 1500)):
        run(B, L, True)
    sys.exit(0)

if __name__ == "__main__":
    for ragged in (False, True):
        run(32, 200, ragged)
        run(32, 400, ragged)
    run(2, 350, False)
    run(2, 700, False)
    run(16, 750, True)
    run(16, 1500, True)
