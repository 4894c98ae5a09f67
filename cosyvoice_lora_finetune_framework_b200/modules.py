"""Estimator module tree of the flow decoder (API mirror of the reference's modules.py:20-375,
844-1106) whose compute runs in the sm_100a CUDA library.

The classes keep the reference's names, constructor signatures, construction order and
parameter names, so that
  * CosyVoice-300M `flow.pt` / reference state-dicts load with strict=True,
  * `lora.apply_lora_to_model` finds the same children (to_q / to_k / to_v), and
  * the same torch seed yields the same random-init weights as the reference.

Only `ConditionalDecoder.forward` computes: it hands the whole U-Net (16 ResnetBlock1D, 64
BasicTransformerBlock at the CosyVoice-300M config) to `cvflow_estimator_forward/backward`
(include/cvflow.h). The sub-modules are parameter holders; calling their `forward` raises,
because there is deliberately no PyTorch/CPU fallback for this path.
"""
import math

import torch
import torch.nn as nn

from . import _estimator


def _fused(name):
    raise RuntimeError(
        "%s is fused into the cvflow CUDA estimator; call ConditionalDecoder.forward "
        "(no PyTorch fallback exists for the flow hot path)" % name)


class SinusoidalPosEmb(nn.Module):
    """[sin(s*t*f_i) | cos(s*t*f_i)], f_i = exp(-i*ln(1e4)/(dim/2-1)), s = 1000 (modules.py:27-42)."""

    def __init__(self, dim):
        super().__init__()
        assert dim % 2 == 0, "SinusoidalPosEmb requires dim to be even"
        self.dim = dim

    def forward(self, x, scale=1000):
        _fused("SinusoidalPosEmb")


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu"):
        super().__init__()
        if act_fn != "silu":
            raise NotImplementedError("cvflow estimator implements the SiLU time MLP only")
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        _fused("TimestepEmbedding")


class Block1D(nn.Module):
    """Conv1d(k=3,p=1) -> GroupNorm(groups) -> Mish on x*mask, output *mask (modules.py:60-73)."""

    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.block = nn.Sequential(nn.Conv1d(dim, dim_out, 3, padding=1), nn.GroupNorm(groups, dim_out),
                                   nn.Mish())

    def forward(self, x, mask):
        _fused("Block1D")


class ResnetBlock1D(nn.Module):
    def __init__(self, dim, dim_out, time_emb_dim, groups=8):
        super().__init__()
        self.mlp = nn.Sequential(nn.Mish(), nn.Linear(time_emb_dim, dim_out))
        self.block1 = Block1D(dim, dim_out, groups=groups)
        self.block2 = Block1D(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv1d(dim, dim_out, 1)

    def forward(self, x, mask, t):
        _fused("ResnetBlock1D")


class Downsample1D(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv1d(dim, dim, 3, 2, 1)

    def forward(self, x):
        _fused("Downsample1D")


class Upsample1D(nn.Module):
    def __init__(self, dim, use_conv_transpose=True):
        super().__init__()
        if not use_conv_transpose:
            raise NotImplementedError("cvflow estimator implements the ConvTranspose1d upsampler only")
        self.conv = nn.ConvTranspose1d(dim, dim, 4, 2, 1)

    def forward(self, x):
        _fused("Upsample1D")


class GELU(nn.Module):
    """Linear + GELU; `approximate` is honoured by the GEMM epilogue ("tanh" default, modules.py:132)."""

    def __init__(self, dim_in, dim_out, approximate="tanh"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, x):
        _fused("GELU")


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu"):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim_out or dim
        if activation_fn in ("gelu", "gelu-approximate"):
            act = GELU(dim, inner, approximate="tanh")
        elif activation_fn in ("geglu", "snakebeta", "snake"):
            raise NotImplementedError(
                "cvflow estimator implements activation_fn='gelu' (the CosyVoice-300M setting, "
                "reference flow_model.py:698); got %r" % activation_fn)
        else:
            act = GELU(dim, inner)
        self.net = nn.ModuleList([act, nn.Dropout(dropout), nn.Linear(inner, dim_out)])

    def forward(self, x):
        _fused("FeedForward")


class Attention(nn.Module):
    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False, cross_attention_dim=None,
                 upcast_attention=False):
        super().__init__()
        if cross_attention_dim is not None:
            raise NotImplementedError("cvflow estimator implements self-attention (attn1) only")
        inner = dim_head * heads
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = upcast_attention
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        _fused("Attention")


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, activation_fn="snakebeta",
                 cross_attention_dim=None, attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, upcast_attention=False):
        super().__init__()
        if cross_attention_dim is not None or double_self_attention or only_cross_attention:
            raise NotImplementedError("cvflow estimator has no cross-attention (attn2) path")
        self.only_cross_attention = False
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                               dropout=dropout, bias=attention_bias, upcast_attention=upcast_attention)
        self.norm2 = None
        self.attn2 = None
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)

    def forward(self, hidden_states, attention_mask=None, **kwargs):
        _fused("BasicTransformerBlock")


def create_prompt_isolation_mask(seq_len: int, prompt_len: int, device, dtype=torch.float32) -> torch.Tensor:
    """(1,1,L,L) additive bias: -inf between the prompt block [0,p) and the target block [p,L)
    in both directions, zeros when p<=0 or p>=L (modules.py:844-879). The CUDA attention kernel
    applies the same rule from the single integer p; this dense form serves tests and callers."""
    bias = torch.zeros(1, 1, seq_len, seq_len, device=device, dtype=dtype)
    if 0 < prompt_len < seq_len:
        bias[:, :, prompt_len:, :prompt_len] = float('-inf')
        bias[:, :, :prompt_len, prompt_len:] = float('-inf')
    return bias


class ConditionalDecoder(nn.Module):
    """U-Net1D estimator v(x_t, t | mu, spks, cond) with prompt isolation (modules.py:886-1106)."""

    def __init__(self, in_channels, out_channels, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                 n_blocks=1, num_mid_blocks=2, num_heads=4, act_fn="snakebeta"):
        super().__init__()
        channels = tuple(channels)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.prompt_isolation_enabled = True
        self.prompt_isolation_len = 0

        self.time_embeddings = SinusoidalPosEmb(in_channels)
        time_embed_dim = channels[0] * 4
        self.time_mlp = TimestepEmbedding(in_channels=in_channels, time_embed_dim=time_embed_dim, act_fn="silu")

        def stage(dim_in, dim_out, tail):
            resnet = ResnetBlock1D(dim=dim_in, dim_out=dim_out, time_emb_dim=time_embed_dim)
            blocks = nn.ModuleList([
                BasicTransformerBlock(dim=dim_out, num_attention_heads=num_heads,
                                      attention_head_dim=attention_head_dim, dropout=dropout,
                                      activation_fn=act_fn)
                for _ in range(n_blocks)])
            parts = [resnet, blocks]
            if tail is not None:
                parts.append(tail(dim_out))
            return nn.ModuleList(parts)

        plain = lambda c: nn.Conv1d(c, c, 3, padding=1)
        self.down_blocks = nn.ModuleList([])
        self.mid_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])
        width = in_channels
        for i, c in enumerate(channels):
            last = i == len(channels) - 1
            self.down_blocks.append(stage(width, c, plain if last else Downsample1D))
            width = c
        for _ in range(num_mid_blocks):
            self.mid_blocks.append(stage(channels[-1], channels[-1], None))
        up = channels[::-1] + (channels[0],)
        for i in range(len(up) - 1):
            last = i == len(up) - 2
            self.up_blocks.append(stage(up[i] * 2, up[i + 1], plain if last else Upsample1D))
        self.final_block = Block1D(up[-1], up[-1])
        self.final_proj = nn.Conv1d(up[-1], self.out_channels, 1)
        self._initialize_weights()
        self._cvflow = None  # native handle + bound weights, created lazily on first CUDA call

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.GroupNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    # -- native binding ---------------------------------------------------------------------
    def _cvflow_invalidate(self):
        """Drop the bound 16-bit weight images (after load_state_dict / LoRA inject / merge)."""
        self._cvflow = None

    def _apply(self, fn, *args, **kwargs):
        self._cvflow = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._cvflow = None
        return super().load_state_dict(*args, **kwargs)

    def forward(self, x, mask, mu, t, spks=None, cond=None):
        """x, mu, cond: (B, 80, T); mask: (B, 1, T) in {0,1}; t: (B,); spks: (B, 80) -> (B, 80, T)."""
        return _estimator.estimator_forward(self, x, mask, mu, t, spks, cond)
