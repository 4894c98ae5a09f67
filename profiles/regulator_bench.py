"""Time the path-input kernels (csrc/regulator.cu) through the C ABI with CUDA events and put them next to their
rooflines, and next to the same work in eager PyTorch (the reference's code path on this GPU).

    python profiles/regulator_bench.py > profiles/r02_regulator_bench.txt

Length regulator, 32 utterances, 232 tokens -> 400 frames, 80 channels, fp32:
  per k=3 layer 2 * 12800 * 80 * 240 = 0.492 GFLOP (forward; the input-gradient layer is the same size),
  bytes per layer (algorithmic): read 12800 x 80 x 4 + write the same = 8.2 MB.
FMA-pipe peak: 148 SMs x 128 lanes x 2 FLOP x SM clock (taken from nvidia-smi under load)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import flow_oracle as O  # noqa: E402  (weights only: synth_regulator_state_dict)


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3      # us


def main():
    from cosyvoice_lora_finetune_framework_b200 import _path_inputs as PI
    from cosyvoice_lora_finetune_framework_b200.encoder import InterpolateRegulator
    import torch.nn.functional as F
    dev = "cuda"
    B, n_src, T = 32, 232, 400
    reg = InterpolateRegulator(channels=80, sampling_ratios=(1, 1, 1, 1), out_channels=80, groups=1)
    sd = O.synth_regulator_state_dict({k: tuple(v.shape) for k, v in reg.state_dict().items()}, 1234)
    reg.load_state_dict(sd)
    for p in reg.parameters():
        p.requires_grad_(False)
    reg = reg.to(dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, n_src, 80, generator=g).to(dev)
    lens = torch.randint(241, 401, (B,), generator=g)
    lens[0] = T
    R = torch.randn(B, T, 80, generator=g).to(dev)
    lens_d = lens.to(dev)

    def fwd_only():
        with torch.no_grad():
            PI.regulate(reg, x, T, lens=lens_d)

    xg = x.clone().requires_grad_(True)

    def fwd_bwd():
        xg.grad = None
        out = PI.regulate(reg, xg, T, lens=lens_d)
        out.backward(R)

    def torch_fwd_bwd():
        xg.grad = None
        keep = (torch.arange(T, device=dev)[None, :] < lens_d[:, None]).float().unsqueeze(-1)
        h = F.interpolate(xg.transpose(1, 2).contiguous(), size=T, mode='linear')
        out = reg.model(h).transpose(1, 2).contiguous() * keep
        out.backward(R)

    clk = 1.92e9
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip()
        clk = float(q) * 1e6
    except Exception:
        pass
    fma_peak = 148 * 128 * 2 * clk / 1e12
    def graphed(fn):          # device time without the host-side launch path (ctypes + autograd.Function + allocator)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay

    t_f_host, t_fb_host = timeit(fwd_only), timeit(fwd_bwd)
    t_f, t_fb = timeit(graphed(fwd_only)), timeit(graphed(fwd_bwd))
    torch.backends.cudnn.allow_tf32 = False
    t_ref = timeit(torch_fwd_bwd, 20)
    fl_f = (4 * 2 * B * T * 80 * 240 + 2 * B * T * 80 * 80) / 1e12
    print("length regulator %d x (%d tokens -> %d frames), fp32, FMA-pipe peak %.1f TFLOP/s at %.0f MHz" % (B, n_src, T, fma_peak, clk / 1e6))
    print("  forward (5 launches, graph replay):                 %7.1f us   %5.2f TFLOP/s = %.2f of the FMA peak" % (t_f, fl_f / (t_f * 1e-6), fl_f / (t_f * 1e-6) / fma_peak))
    print("  forward + backward (5 + 6 launches, graph):  %7.1f us   %5.2f TFLOP/s = %.2f of the FMA peak" %
          (t_fb, 2 * fl_f / (t_fb * 1e-6), 2 * fl_f / (t_fb * 1e-6) / fma_peak))
    print("  (the same launched eagerly from Python, host-bound: forward %.1f us, forward + backward %.1f us)" % (t_f_host, t_fb_host))
    print("  eager PyTorch fp32 (cuDNN / ATen), forward + backward: %7.1f us  (this path, eager: %.1fx)" % (t_ref, t_ref / t_fb_host))


if __name__ == "__main__":
    main()
