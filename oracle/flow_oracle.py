"""CPU oracle of the flow hot path -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32, functional restatement of the reference's algorithm for the path this repo
accelerates. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; the product (cosyvoice_lora_finetune_framework_b200) never does.

Parity pin: tests/golden/*.pt were produced by the *real* reference
(/root/reference/cosyvoice_flow_finetune: flow_model.py, modules.py, lora.py) on seeded inputs and
weights by tests/golden/make_golden.py; tests/test_oracle.py checks this restatement against them
(and the docstring known-answers of reference utils.py:28-33), so the oracle is pinned, not
"parity unpinned".

Every function cites the reference lines it restates (paths relative to
/root/reference/cosyvoice_flow_finetune/). Weights arrive as a flat {name: tensor} dict in the
reference's state_dict key layout, with or without LoRA wrapping
(`<p>.original_layer.weight`, `<p>.lora_A`, `<p>.lora_B`  vs  `<p>.weight`).
"""
import math
import zlib
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# integer / mask helpers                                                       utils.py:20-41,103-109
# ----------------------------------------------------------------------------------------------
def make_pad_mask(lengths: Tensor, max_len: int = 0) -> Tensor:
    n = max_len if max_len > 0 else int(lengths.max())
    return torch.arange(n, dtype=torch.int64, device=lengths.device)[None, :] >= lengths.to(torch.int64)[:, None]


def mask_to_bias(mask: Tensor, dtype=torch.float32) -> Tensor:
    return (1.0 - mask.to(dtype)) * -1.0e10


def isolation_bias(seq_len: int, p: int) -> Tensor:
    """modules.py:844-879"""
    m = torch.zeros(1, seq_len, seq_len)
    if 0 < p < seq_len:
        m[:, p:, :p] = float("-inf")
        m[:, :p, p:] = float("-inf")
    return m


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
def sinusoidal_embedding(t: Tensor, dim: int = 320, scale: float = 1000.0) -> Tensor:
    """modules.py:27-42 -- [sin | cos] of scale * t * exp(-i ln(1e4)/(half-1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=t.device).float() * -(math.log(10000) / (half - 1)))
    arg = scale * t[:, None] * freq[None, :]
    return torch.cat([arg.sin(), arg.cos()], dim=-1)


def _linear(P, prefix: str, x: Tensor, lora_scaling: Dict[str, float]) -> Tensor:
    """nn.Linear or LoRALinear (lora.py:64-76). nn.Dropout's draw on the LoRA input is supplied by the caller as an
    optional `<prefix>.lora_dropout_mask` entry of P (keep / (1 - p), broadcastable to x); absent = no dropout.
    For TIMING runs (bench.py's CPU baseline at the reference's default lora_dropout = 0.05) a float entry
    P["__lora_dropout_p__"] makes every LoRA layer draw its own mask with F.dropout, like the reference's nn.Dropout."""
    if prefix + ".lora_A" in P:
        w, b = P[prefix + ".original_layer.weight"], P.get(prefix + ".original_layer.bias")
        y = F.linear(x, w, b)
        dm = P.get(prefix + ".lora_dropout_mask")
        xl = x * dm if dm is not None else x
        if dm is None and P.get("__lora_dropout_p__"):
            xl = F.dropout(x, float(P["__lora_dropout_p__"]), training=True)
        low = F.linear(F.linear(xl, P[prefix + ".lora_A"]), P[prefix + ".lora_B"])
        return y + low * lora_scaling[prefix]
    return F.linear(x, P[prefix + ".weight"], P.get(prefix + ".bias"))


def block1d(P, prefix: str, x: Tensor, mask: Tensor, groups: int = 8) -> Tensor:
    """modules.py:60-73 -- Mish(GN(Conv1d_k3(x*mask))) * mask; GN statistics over the padded T."""
    h = F.conv1d(x * mask, P[prefix + ".block.0.weight"], P[prefix + ".block.0.bias"], padding=1)
    h = F.group_norm(h, groups, P[prefix + ".block.1.weight"], P[prefix + ".block.1.bias"], eps=1e-5)
    return F.mish(h) * mask


def resnet_block(P, prefix: str, x: Tensor, mask: Tensor, temb: Tensor) -> Tensor:
    """modules.py:76-94 -- the residual branch is NOT re-masked."""
    h = block1d(P, prefix + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), P[prefix + ".mlp.1.weight"], P[prefix + ".mlp.1.bias"])[:, :, None]
    h = block1d(P, prefix + ".block2", h, mask)
    return h + F.conv1d(x * mask, P[prefix + ".res_conv.weight"], P[prefix + ".res_conv.bias"])


def transformer_block(P, prefix: str, h: Tensor, bias: Tensor, lora_scaling, heads: int = 8,
                      gelu_approximate: str = "tanh") -> Tensor:
    """modules.py:349-375 (block), :253-293 (attention), :192-224,127-139 (feed-forward)."""
    b, n, c = h.shape
    xn = F.layer_norm(h, (c,), P[prefix + ".norm1.weight"], P[prefix + ".norm1.bias"], eps=1e-5)
    q = _linear(P, prefix + ".attn1.to_q", xn, lora_scaling)
    k = _linear(P, prefix + ".attn1.to_k", xn, lora_scaling)
    v = _linear(P, prefix + ".attn1.to_v", xn, lora_scaling)
    d = q.shape[-1] // heads
    q, k, v = (z.view(b, n, heads, d).transpose(1, 2) for z in (q, k, v))
    sim = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5) + bias[:, None]
    o = torch.matmul(sim.softmax(dim=-1), v).transpose(1, 2).reshape(b, n, heads * d)
    h = h + _linear(P, prefix + ".attn1.to_out.0", o, lora_scaling)
    xn = F.layer_norm(h, (c,), P[prefix + ".norm3.weight"], P[prefix + ".norm3.bias"], eps=1e-5)
    f = F.gelu(_linear(P, prefix + ".ff.net.0.proj", xn, lora_scaling), approximate=gelu_approximate)
    return h + _linear(P, prefix + ".ff.net.2", f, lora_scaling)


def _count(P, fmt: str) -> int:
    n = 0
    while any(k.startswith(fmt % n) for k in P):
        n += 1
    return n


def _stage(P, prefix, x, mask, temb, full_T, iso_len, iso_enabled, lora_scaling, gelu_approximate):
    """One resnet + its transformer stack (modules.py:1023-1046 / 1054-1075 / 1081-1101)."""
    x = resnet_block(P, prefix + ".0", x, mask, temb)
    h = x.transpose(1, 2).contiguous()
    L = h.shape[1]
    bias = mask_to_bias(mask.bool().expand(-1, L, -1))
    if iso_enabled and iso_len > 0:
        p = max(1, int(iso_len * (L / full_T)))
        if p < L:
            bias = bias + isolation_bias(L, p).to(bias.device)
    for j in range(_count(P, prefix + ".1.%d.")):
        h = transformer_block(P, "%s.1.%d" % (prefix, j), h, bias, lora_scaling,
                              gelu_approximate=gelu_approximate)
    return h.transpose(1, 2).contiguous()


def estimator_forward(P: Dict[str, Tensor], x: Tensor, mask: Tensor, mu: Tensor, t: Tensor,
                      spks: Optional[Tensor] = None, cond: Optional[Tensor] = None, *,
                      isolation_len: int = 0, isolation_enabled: bool = True,
                      lora_scaling: Optional[Dict[str, float]] = None,
                      gelu_approximate: str = "tanh") -> Tensor:
    """ConditionalDecoder.forward, modules.py:998-1106. P uses keys relative to the estimator."""
    lora_scaling = lora_scaling or {}
    temb = sinusoidal_embedding(t, 320).to(t.dtype)
    temb = F.linear(temb, P["time_mlp.linear_1.weight"], P["time_mlp.linear_1.bias"])
    temb = F.linear(F.silu(temb), P["time_mlp.linear_2.weight"], P["time_mlp.linear_2.bias"])
    T = x.shape[-1]
    parts = [x, mu]
    if spks is not None:
        parts.append(spks[:, :, None].expand(-1, -1, T))
    if cond is not None:
        parts.append(cond)
    x = torch.cat(parts, dim=1)
    args = (T, isolation_len, isolation_enabled, lora_scaling, gelu_approximate)

    hiddens, masks = [], [mask]
    n_down = _count(P, "down_blocks.%d.")
    for i in range(n_down):
        m = masks[-1]
        x = _stage(P, "down_blocks.%d" % i, x, m, temb, *args)
        hiddens.append(x)
        if "down_blocks.%d.2.conv.weight" % i in P:      # Downsample1D: k3 s2 p1
            x = F.conv1d(x * m, P["down_blocks.%d.2.conv.weight" % i], P["down_blocks.%d.2.conv.bias" % i],
                         stride=2, padding=1)
        else:                                             # last stage: plain k3 p1
            x = F.conv1d(x * m, P["down_blocks.%d.2.weight" % i], P["down_blocks.%d.2.bias" % i], padding=1)
        masks.append(m[:, :, ::2])
    masks = masks[:-1]
    m_mid = masks[-1]
    for j in range(_count(P, "mid_blocks.%d.")):
        x = _stage(P, "mid_blocks.%d" % j, x, m_mid, temb, *args)
    for i in range(_count(P, "up_blocks.%d.")):
        m = masks.pop()
        skip = hiddens.pop()
        x = torch.cat([x[:, :, : skip.shape[-1]], skip], dim=1)
        x = _stage(P, "up_blocks.%d" % i, x, m, temb, *args)
        if "up_blocks.%d.2.conv.weight" % i in P:         # Upsample1D: ConvTranspose1d k4 s2 p1
            x = F.conv_transpose1d(x * m, P["up_blocks.%d.2.conv.weight" % i], P["up_blocks.%d.2.conv.bias" % i],
                                   stride=2, padding=1)
        else:
            x = F.conv1d(x * m, P["up_blocks.%d.2.weight" % i], P["up_blocks.%d.2.bias" % i], padding=1)
    x = block1d(P, "final_block", x, m)
    out = F.conv1d(x * m, P["final_proj.weight"], P["final_proj.bias"])
    return out * mask


# ----------------------------------------------------------------------------------------------
# conditional flow matching                                                flow_model.py:50-204
# ----------------------------------------------------------------------------------------------
SIGMA_MIN = 1e-6
PI_HALF = 0.5 * 3.14159265359  # the reference's literal (flow_model.py:90,148)


def cfm_loss_weights(mask: Tensor, prompt_lens: Optional[List[int]], boundary_frames: int = 25,
                     boundary_weight: float = 5.0, boundary_enabled: bool = True) -> Tensor:
    """flow_model.py:180-194 -- prompt frames weigh 0, the next `boundary_frames` weigh 5."""
    w = mask.clone()
    if prompt_lens is not None:
        for i, p in enumerate(prompt_lens):
            if p > 0:
                w[i, :, :p] = 0
                if boundary_enabled:
                    w[i, :, p:min(p + boundary_frames, w.shape[2])] = boundary_weight
    return w


def cfm_compute_loss(P, x1: Tensor, mask: Tensor, mu: Tensor, spks: Tensor, cond: Tensor,
                     prompt_lens: Optional[List[int]], t_rand: Tensor, z: Tensor, cfg_rand: Tensor, *,
                     sigma_min: float = SIGMA_MIN, training_cfg_rate: float = 0.2,
                     lora_scaling=None, boundary_frames: int = 25, boundary_weight: float = 5.0,
                     gelu_approximate: str = "tanh"):
    """ConditionalCFM.compute_loss, flow_model.py:127-204, with the three random draws
    (rand([B,1,1]), randn_like(x1), rand(B)) supplied by the caller. Returns (loss, y, pred)."""
    b = mu.shape[0]
    t = 1 - torch.cos(t_rand * PI_HALF)
    y = (1 - (1 - sigma_min) * t) * z + t * x1
    u = x1 - (1 - sigma_min) * z
    if training_cfg_rate > 0:
        keep = cfg_rand > training_cfg_rate
        mu = mu * keep.view(-1, 1, 1)
        spks = spks * keep.view(-1, 1)
        cond = cond * keep.view(-1, 1, 1)
    iso = max(prompt_lens) if prompt_lens else 0
    pred = estimator_forward(P, y, mask, mu, t.view(b), spks, cond, isolation_len=iso,
                             lora_scaling=lora_scaling, gelu_approximate=gelu_approximate)
    w = cfm_loss_weights(mask, prompt_lens, boundary_frames, boundary_weight)
    diff = (pred - u) * w
    denom = w.sum() * u.shape[1]
    loss = (diff ** 2).sum() / denom if denom > 0 else torch.zeros((), requires_grad=True, device=diff.device)
    return loss, y, pred


def cfm_t_span(n_timesteps: int) -> Tensor:
    """flow_model.py:88-90"""
    return 1 - torch.cos(torch.linspace(0, 1, n_timesteps + 1) * PI_HALF)


def cfm_solve_euler(P, x: Tensor, t_span: Tensor, mu: Tensor, mask: Tensor, spks: Tensor, cond: Tensor, *,
                    inference_cfg_rate: float = 0.7, lora_scaling=None, gelu_approximate: str = "tanh") -> Tensor:
    """ConditionalCFM.solve_euler, flow_model.py:94-125: batch-2 packing (row 0 conditional,
    row 1 = zeros for mu/spks/cond), CFG combine, accumulated t and re-derived dt."""
    t = t_span[0:1]
    dt = t_span[1] - t_span[0]
    T = x.shape[2]
    for step in range(1, len(t_span)):
        x_in = x.expand(2, -1, -1)
        mask_in = mask.expand(2, -1, -1)
        mu_in = torch.cat([mu, torch.zeros_like(mu)], 0)
        spks_in = torch.cat([spks, torch.zeros_like(spks)], 0)
        cond_in = torch.cat([cond, torch.zeros_like(cond)], 0)
        t_in = t.expand(2)
        d = estimator_forward(P, x_in, mask_in, mu_in, t_in, spks_in, cond_in, lora_scaling=lora_scaling,
                              gelu_approximate=gelu_approximate)
        d = (1.0 + inference_cfg_rate) * d[0:1] - inference_cfg_rate * d[1:2]
        x = x + dt * d
        t = t + dt
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t
    return x.float()


def cfm_forward(P, mu: Tensor, mask: Tensor, n_timesteps: int, z: Tensor, spks: Tensor, cond: Tensor, *,
                temperature: float = 1.0, prompt_len: int = 0, cache: Optional[Tensor] = None, **kw):
    """ConditionalCFM.forward, flow_model.py:74-92, with z = randn_like(mu) supplied by the caller."""
    z = z * temperature
    mu = mu.clone()
    if cache is not None and cache.shape[2] != 0:
        n = cache.shape[2]
        z[:, :, :n] = cache[:, :, :, 0]
        mu[:, :, :n] = cache[:, :, :, 1]
    if prompt_len > 0:
        z_cache = torch.cat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2)
        mu_cache = torch.cat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)
    else:
        z_cache, mu_cache = z[:, :, -34:], mu[:, :, -34:]
    new_cache = torch.stack([z_cache, mu_cache], dim=-1)
    mel = cfm_solve_euler(P, z, cfm_t_span(n_timesteps), mu, mask, spks, cond, **kw)
    return mel, new_cache


# ----------------------------------------------------------------------------------------------
# the inputs of the path: length regulator, speaker affine, conditioning    modules.py:800-837, flow_model.py:266-387
# ----------------------------------------------------------------------------------------------
def interp_linear_taps(n_in: int, n_out: int):
    """Index arithmetic of F.interpolate(mode='linear', align_corners=False) (ATen upsample_linear1d:
    area_pixel_compute_source_index, fp32 throughout): for each of the n_out frames the two source indices and their
    weights, as numpy arrays (i0, i1, w0, w1). Pinned bit-for-bit against torch itself in tests/test_oracle.py."""
    import numpy as np
    f32 = np.float32
    scale = f32(n_in) / f32(n_out)
    j = np.arange(n_out, dtype=np.float32)
    # ATen evaluates scale * (j + 0.5) - 0.5 as ONE fused multiply-add (CPU and CUDA builds alike); float64 holds the
    # product of two float32 and the subtraction exactly, so rounding it once to float32 is that FMA
    s = (float(scale) * (j + f32(0.5)).astype(np.float64) - 0.5).astype(np.float32)
    s = np.where(s < 0, f32(0), s).astype(np.float32)
    i0 = np.minimum(s.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    w1 = np.clip(s - i0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    w0 = (f32(1) - w1).astype(np.float32)
    return i0, i1, w0, w1


def interp_linear(x: Tensor, n_out: int) -> Tensor:
    """x [B][n_in][C] -> [B][n_out][C] with the taps above (== F.interpolate on the transposed tensor)."""
    i0, i1, w0, w1 = interp_linear_taps(x.shape[1], n_out)
    i0, i1 = torch.from_numpy(i0), torch.from_numpy(i1)
    w0, w1 = torch.from_numpy(w0).to(x)[None, :, None], torch.from_numpy(w1).to(x)[None, :, None]
    return w0 * x[:, i0] + w1 * x[:, i1]


def regulator_stack(P, prefix: str, x: Tensor) -> Tensor:
    """InterpolateRegulator.model (modules.py:806-818): x [B][T][80] -> [B][T][80]."""
    h = x.transpose(1, 2)
    n = 0
    while f"{prefix}model.{3 * n}.weight" in P and P[f"{prefix}model.{3 * n}.weight"].shape[-1] == 3:
        h = F.conv1d(h, P[f"{prefix}model.{3 * n}.weight"], P[f"{prefix}model.{3 * n}.bias"], padding=1)
        h = F.group_norm(h, 1, P[f"{prefix}model.{3 * n + 1}.weight"], P[f"{prefix}model.{3 * n + 1}.bias"], eps=1e-5)
        h = F.mish(h)
        n += 1
    h = F.conv1d(h, P[f"{prefix}model.{3 * n}.weight"], P[f"{prefix}model.{3 * n}.bias"])
    return h.transpose(1, 2)


def regulator_forward(P, prefix: str, x: Tensor, ylens: Tensor) -> Tensor:
    """InterpolateRegulator.forward (modules.py:820-824)."""
    keep = (~make_pad_mask(ylens)).to(x).unsqueeze(-1)
    return regulator_stack(P, prefix, interp_linear(x, int(ylens.max()))) * keep


def regulator_inference(P, prefix: str, x1: Tensor, x2: Tensor, mel_len1: int, mel_len2: int, input_frame_rate: int = 50) -> Tensor:
    """InterpolateRegulator.inference (modules.py:826-837): the target is stretched in three pieces."""
    if x2.shape[1] > 40:
        edge = int(20 / input_frame_rate * 22050 / 256)
        x2 = torch.cat([interp_linear(x2[:, :20], edge), interp_linear(x2[:, 20:-20], mel_len2 - edge * 2),
                        interp_linear(x2[:, -20:], edge)], dim=1)
    else:
        x2 = interp_linear(x2, mel_len2)
    x = torch.cat([interp_linear(x1, mel_len1), x2], dim=1) if x1.shape[1] != 0 else x2
    return regulator_stack(P, prefix, x)


def speaker_affine(W: Tensor, b: Tensor, embedding: Tensor) -> Tensor:
    """flow_model.py:297-298."""
    return F.linear(F.normalize(embedding, dim=1), W, b)


def path_inputs_pack(feat: Tensor, cross: Optional[Tensor], desc, mel_mean: float, mel_std: float, silence: float):
    """Mel normalisation, conditioning and mask of flow_model.py:266-269, 319-387 from per-utterance descriptors
    desc[b] = (len, prompt frames, silence-gap frames, prompt from cross): x1, cond [B][80][T], mask [B][1][T]."""
    B, T, _ = feat.shape
    f = (feat - mel_mean) / mel_std
    c = (cross - mel_mean) / mel_std if cross is not None else None
    cond = torch.zeros_like(f)
    mask = torch.zeros(B, 1, T)
    for i, (n, p, gap, use_cross) in enumerate(desc):
        if p > 0:
            cond[i, :p] = (c if use_cross else f)[i, :p]
        if gap > 0:
            cond[i, p:p + gap] = silence
        mask[i, 0, :n] = 1.0
    return f.transpose(1, 2).contiguous(), cond.transpose(1, 2).contiguous(), mask


# ----------------------------------------------------------------------------------------------
# seeded synthetic weights shared by the golden generator, the tests and the bench
# ----------------------------------------------------------------------------------------------
def synth_tensor(name: str, shape, seed: int) -> Tensor:
    """Deterministic values keyed by parameter *name* (independent of construction order, so the
    reference model here and our model on the GPU box get bit-identical weights)."""
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "lora_A":
        bound = 1.0 / math.sqrt(shape[1])
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if leaf == "lora_B":
        return torch.randn(shape, generator=g) * 0.01
    if leaf == "bias":
        if ".norm" in name or ".block.1." in name:
            return torch.randn(shape, generator=g) * 0.1
        return torch.randn(shape, generator=g) * 0.02
    if leaf == "weight" and len(shape) == 1:
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    if ".2.conv.weight" in name and name.startswith("up_blocks"):
        fan_in = shape[0] * shape[2] // 2   # ConvTranspose1d [Cin, Cout, k], stride 2
    return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)


def synth_state_dict(spec: Dict[str, tuple], seed: int = 1234) -> Dict[str, Tensor]:
    return {k: synth_tensor(k, shp, seed) for k, shp in spec.items()}


def synth_regulator_state_dict(spec: Dict[str, tuple], seed: int = 1234) -> Dict[str, Tensor]:
    """synth_state_dict for an InterpolateRegulator: the GroupNorm scales are moved to 1 + 0.1 w so the stack stays alive."""
    sd = synth_state_dict(spec, seed)
    for k in sd:
        if k.endswith("weight") and sd[k].dim() == 1:
            sd[k] = 1.0 + 0.1 * sd[k]
    return sd
