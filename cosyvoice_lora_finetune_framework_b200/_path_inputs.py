"""Host side of the path-input kernels (csrc/regulator.cu; include/cvflow.h, last section): the length regulator with
its autograd, the speaker affine layer and the conditioning pack, as called by encoder.InterpolateRegulator and
flow_model.MaskedDiffWithXvec when their tensors live on a CUDA device.

Reference: InterpolateRegulator modules.py:800-837, MaskedDiffWithXvec.forward flow_model.py:248-400.
The kernels take the module's FROZEN parameters (LoRA fine-tuning never touches them: lora.py injects only nn.Linear /
1x1 convs named in target_modules, config.py:207-216); a trainable regulator or speaker layer is refused loudly."""
import ctypes as C

import torch

from . import _native as N

_protos_done = False


class RegulatorWeights(C.Structure):
    _fields_ = [("wf", C.c_void_p * 5), ("wb", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("gamma", C.c_void_p * 4),
                ("beta", C.c_void_p * 4)]


class RegulatorIO(C.Structure):
    _fields_ = [("src", C.c_void_p), ("B", C.c_int32), ("n_src", C.c_int32), ("T", C.c_int32), ("n_seg", C.c_int32),
                ("seg", (C.c_int32 * 4) * 4), ("lens", C.c_void_p), ("blind", C.c_void_p), ("out", C.c_void_p),
                ("channel_major", C.c_int32), ("saved", C.c_void_p)]


def _lib():
    global _protos_done
    L = N.lib()
    if not _protos_done:
        vp, i32, i64, f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.cvflow_regulator_saved_floats.argtypes = [i32, i32]
        L.cvflow_regulator_saved_floats.restype = i64
        L.cvflow_regulator_scratch_floats.argtypes = [i32, i32]
        L.cvflow_regulator_scratch_floats.restype = i64
        L.cvflow_regulator_forward.argtypes = [C.POINTER(RegulatorWeights), C.POINTER(RegulatorIO), vp]
        L.cvflow_regulator_backward.argtypes = [C.POINTER(RegulatorWeights), C.POINTER(RegulatorIO), vp, vp, vp, vp]
        L.cvflow_path_inputs_pack.argtypes = [vp, vp, i32, vp, f, f, f, vp, vp, vp, i32, i32, vp]
        L.cvflow_spk_affine.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
        _protos_done = True
    return L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _conv_image(w):
    """Conv1d weight [co][ci][taps] -> kernel image [ci][tap][16][6] (5 output channels per group, padded to 6)."""
    co, ci, taps = w.shape
    img = torch.zeros(ci, taps, 16, 6, device=w.device, dtype=torch.float32)
    img[:, :, :, :5] = w.detach().float().permute(1, 2, 0).reshape(ci, taps, 16, 5)
    return img.contiguous()


class RegulatorImages:
    """Kernel-side images of InterpolateRegulator.model's parameters, rebuilt when a parameter changes."""

    def __init__(self, module):
        convs = [m for m in module.model if isinstance(m, torch.nn.Conv1d)]
        norms = [m for m in module.model if isinstance(m, torch.nn.GroupNorm)]
        ok = (len(convs) == 5 and len(norms) == 4 and all(c.weight.shape == (80, 80, 3) and c.padding == (1,) for c in convs[:4])
              and convs[4].weight.shape == (80, 80, 1) and all(n.num_groups == 1 and n.num_channels == 80 for n in norms)
              and all(c.bias is not None for c in convs))
        if not ok:
            raise NotImplementedError("the CUDA length regulator implements the CosyVoice-300M configuration: "
                                      "InterpolateRegulator(channels=80, sampling_ratios=(1,1,1,1), groups=1)")
        self.convs, self.norms = convs, norms
        self.key = None
        self.refresh()

    def _params(self):
        return [p for m in self.convs + self.norms for p in m.parameters()]

    def refresh(self):
        ps = self._params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in ps):      # inference (no_grad / inference_mode) is fine
            raise RuntimeError("the CUDA length regulator computes input gradients only: its parameters must be frozen "
                               "(requires_grad=False), as they are under the reference's LoRA fine-tuning")
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key == self.key:
            return
        f32 = lambda t: t.detach().float().contiguous()
        self.wf = [_conv_image(c.weight) for c in self.convs]
        # dgrad: dx[t][ci] = sum_{co,k} w[co][ci][k] dy[t - k + pad][co] = a convolution with w'[ci][co][k'] = w[co][ci][taps-1-k']
        self.wb = [_conv_image(c.weight.detach().permute(1, 0, 2).flip(2)) for c in self.convs]
        self.bias = [f32(c.bias) for c in self.convs]
        self.gamma = [f32(n.weight) for n in self.norms]
        self.beta = [f32(n.bias) for n in self.norms]
        w = RegulatorWeights()
        for i in range(5):
            w.wf[i], w.wb[i], w.bias[i] = self.wf[i].data_ptr(), self.wb[i].data_ptr(), self.bias[i].data_ptr()
        for i in range(4):
            w.gamma[i], w.beta[i] = self.gamma[i].data_ptr(), self.beta[i].data_ptr()
        self.c = w
        self.key = key


def images_of(module):
    im = module.__dict__.get("_cvflow_images")
    if im is None or im.wf[0].device != next(module.parameters()).device:
        im = RegulatorImages(module)
        module.__dict__["_cvflow_images"] = im
    else:
        im.refresh()
    return im


def _io(src, T, segs, lens, blind, out, channel_major, saved):
    io = RegulatorIO()
    io.src, io.B, io.n_src, io.T, io.n_seg = src.data_ptr(), src.shape[0], src.shape[1], T, len(segs)
    for i, s in enumerate(segs):
        for j in range(4):
            io.seg[i][j] = int(s[j])
    io.lens = lens.data_ptr() if lens is not None else None
    io.blind = blind.data_ptr() if blind is not None else None
    io.out = out.data_ptr() if out is not None else None
    io.channel_major = 1 if channel_major else 0
    io.saved = saved.data_ptr()
    return io


class _RegulatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, im, T, segs, lens, blind, channel_major):
        L = _lib()
        B, n_src, Cc = src.shape
        src = src.contiguous().float()
        saved = torch.empty(L.cvflow_regulator_saved_floats(B, T), device=src.device, dtype=torch.float32)
        out = torch.empty((B, 80, T) if channel_major else (B, T, 80), device=src.device, dtype=torch.float32)
        io = _io(src, T, segs, lens, blind, out, channel_major, saved)
        N.check(L.cvflow_regulator_forward(C.byref(im.c), C.byref(io), _stream()), "cvflow_regulator_forward")
        ctx.im, ctx.meta = im, (T, segs, channel_major)
        ctx.save_for_backward(src, saved, lens, blind)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib()
        src, saved, lens, blind = ctx.saved_tensors
        T, segs, channel_major = ctx.meta
        B = src.shape[0]
        dout = dout.contiguous().float()
        dsrc = torch.empty_like(src)
        scratch = torch.empty(L.cvflow_regulator_scratch_floats(B, T), device=src.device, dtype=torch.float32)
        io = _io(src, T, segs, lens, blind, None, channel_major, saved)
        N.check(L.cvflow_regulator_backward(C.byref(ctx.im.c), C.byref(io), dout.data_ptr(), dsrc.data_ptr(),
                                            scratch.data_ptr(), _stream()), "cvflow_regulator_backward")
        return dsrc, None, None, None, None, None, None


def _i32(x, device):
    if x is None:
        return None
    if not torch.is_tensor(x):
        x = torch.tensor(list(x), dtype=torch.int32)
    return x.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


def regulate(module, x, T, lens=None, blind=None, channel_major=False, segs=None):
    """x [B][n_src][80] -> regulated [B][T][80] (or [B][80][T]); differentiable with respect to x."""
    if x.shape[2] != 80:
        raise NotImplementedError("the CUDA length regulator is built for 80 channels")
    im = images_of(module)
    segs = tuple(tuple(int(v) for v in s) for s in (segs or ((0, x.shape[1], 0, T),)))
    return _RegulatorFn.apply(x, im, int(T), segs, _i32(lens, x.device), _i32(blind, x.device), bool(channel_major))


def inference_segments(n_prompt, n_target, mel_len1, mel_len2, input_frame_rate=50):
    """The piecewise interpolation of InterpolateRegulator.inference (modules.py:826-836) as a segment table over the
    concatenated [prompt | target] token sequence."""
    segs, dst = [], 0
    if n_prompt != 0:
        segs.append((0, n_prompt, 0, mel_len1))
        dst = mel_len1
    if n_target > 40:
        edge = int(20 / input_frame_rate * 22050 / 256)
        mid = mel_len2 - edge * 2
        for s0, sn, dn in ((n_prompt, 20, edge), (n_prompt + 20, n_target - 40, mid), (n_prompt + n_target - 20, 20, edge)):
            segs.append((s0, sn, dst, dn))
            dst += dn
    else:
        segs.append((n_prompt, n_target, dst, mel_len2))
    return segs


def spk_affine(layer, embedding):
    """Linear(F.normalize(embedding, dim=1)) with the layer's frozen parameters (flow_model.py:297-298)."""
    if torch.is_grad_enabled() and (layer.weight.requires_grad or embedding.requires_grad):
        raise RuntimeError("the CUDA speaker affine kernel is forward-only: its parameters / input must not require grad")
    e = embedding.contiguous().float()
    w = layer.weight.detach().float().contiguous()
    b = layer.bias.detach().float().contiguous() if layer.bias is not None else None
    out = torch.empty(e.shape[0], w.shape[0], device=e.device, dtype=torch.float32)
    N.check(_lib().cvflow_spk_affine(e.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                     e.shape[0], e.shape[1], w.shape[0], _stream()), "cvflow_spk_affine")
    return out


def pack_inputs(feat, cross, desc, mel_mean, mel_std, silence):
    """feat raw log-mel [B][T][80] (+ optional cross-sample mel) and desc [B][4] int32 (host or device) ->
    x1 [B][80][T] normalised, cond [B][80][T], mask [B][1][T]."""
    B, T, _ = feat.shape
    dev = feat.device
    feat = feat.contiguous().float()
    cross = cross.contiguous().float() if cross is not None else None
    desc = _i32(desc, dev)
    x1 = torch.empty(B, 80, T, device=dev, dtype=torch.float32)
    cond = torch.empty_like(x1)
    mask = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
    N.check(_lib().cvflow_path_inputs_pack(feat.data_ptr(), cross.data_ptr() if cross is not None else None,
                                           cross.shape[1] if cross is not None else 0, desc.data_ptr(), float(mel_mean),
                                           float(mel_std), float(silence), x1.data_ptr(), cond.data_ptr(), mask.data_ptr(),
                                           B, T, _stream()), "cvflow_path_inputs_pack")
    return x1, cond, mask
