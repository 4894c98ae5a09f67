"""Host-side producers of the prepared tensors (`mu`) that feed the CUDA flow path: the token
Conformer encoder and the interpolating length regulator of CosyVoice-300M's flow model.

These are callers of the accelerated path (SURVEY.md section 8 a12 / 8f "next"), kept in plain
PyTorch with the reference's class names, constructor signatures, construction order and parameter
names (reference modules.py:382-837), so `flow.pt` loads strictly and seeded random init matches.
Only the CosyVoice-300M flow configuration is implemented: pre-norm blocks with relative-position
attention and a feed-forward, no macaron branch, no convolution module (flow_model.py:663-677).
"""
import math
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils import make_pad_mask


class RelPositionalEncoding(nn.Module):
    """ESPnet-style relative positional table: positions +(L-1) .. 0 .. -(L-1)."""

    def __init__(self, d_model, dropout_rate=0.0, max_len=5000):
        super().__init__()
        self.d_model = d_model
        self.dropout = nn.Dropout(p=dropout_rate)
        self.pe = None
        self.extend_pe(torch.tensor(0.0).expand(1, max_len))

    def extend_pe(self, x):
        n = x.size(1)
        if self.pe is not None and self.pe.size(1) >= 2 * n - 1:
            if self.pe.dtype != x.dtype or self.pe.device != x.device:
                self.pe = self.pe.to(dtype=x.dtype, device=x.device)
            return
        pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / self.d_model))
        plus = torch.zeros(n, self.d_model)
        minus = torch.zeros(n, self.d_model)
        plus[:, 0::2], plus[:, 1::2] = torch.sin(pos * freq), torch.cos(pos * freq)
        minus[:, 0::2], minus[:, 1::2] = torch.sin(-1 * pos * freq), torch.cos(-1 * pos * freq)
        table = torch.cat([torch.flip(plus, [0]).unsqueeze(0), minus[1:].unsqueeze(0)], dim=1)
        self.pe = table.to(device=x.device, dtype=x.dtype)

    def forward(self, x):
        self.extend_pe(x)
        mid, n = self.pe.size(1) // 2, x.size(1)
        return self.dropout(x), self.dropout(self.pe[:, mid - n + 1: mid + n])

    def position_encoding(self, offset, size):
        self.extend_pe(torch.tensor(0.0).expand(1, offset + size))
        mid = self.pe.size(1) // 2
        return self.pe[:, mid - offset - size + 1: mid - offset + size]


class LinearNoSubsampling(nn.Module):
    def __init__(self, idim, odim, dropout_rate, pos_enc):
        super().__init__()
        self.out = nn.Sequential(nn.Linear(idim, odim), nn.LayerNorm(odim, eps=1e-5), nn.Dropout(dropout_rate))
        self.pos_enc = pos_enc
        self.right_context = 0
        self.subsampling_rate = 1

    def forward(self, x, x_mask, offset=0):
        x, pos_emb = self.pos_enc(self.out(x))
        return x, pos_emb, x_mask


class RelPositionMultiHeadedAttention(nn.Module):
    """Multi-head attention with the (u, v)-biased relative-position score of Transformer-XL."""

    def __init__(self, n_head, n_feat, dropout_rate, key_bias=True):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        self.linear_q = nn.Linear(n_feat, n_feat, bias=key_bias)
        self.linear_k = nn.Linear(n_feat, n_feat, bias=key_bias)
        self.linear_v = nn.Linear(n_feat, n_feat, bias=key_bias)
        self.linear_out = nn.Linear(n_feat, n_feat, bias=key_bias)
        self.linear_pos = nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = nn.Parameter(torch.Tensor(self.h, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.Tensor(self.h, self.d_k))
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)
        self.dropout = nn.Dropout(p=dropout_rate)

    @staticmethod
    def rel_shift(x):
        """(b, h, t, 2t-1) scores indexed by relative offset -> (b, h, t, t) indexed by key."""
        b, h, t, w = x.shape
        padded = torch.cat([x.new_zeros(b, h, t, 1), x], dim=-1).view(b, h, w + 1, t)
        return padded[:, :, 1:].view_as(x)[:, :, :, : w // 2 + 1]

    def _heads(self, lin, x):
        return lin(x).view(x.size(0), -1, self.h, self.d_k)

    def forward(self, query, key, value, mask, pos_emb, cache=torch.zeros((0, 0, 0, 0))):
        q = self._heads(self.linear_q, query)                       # (b, t1, h, d)
        k = self._heads(self.linear_k, key).transpose(1, 2)         # (b, h, t2, d)
        v = self._heads(self.linear_v, value).transpose(1, 2)
        if cache.size(0) > 0:
            k_old, v_old = torch.split(cache, cache.size(-1) // 2, dim=-1)
            k, v = torch.cat([k_old, k], dim=2), torch.cat([v_old, v], dim=2)
        new_cache = torch.cat((k, v), dim=-1)
        p = self._heads(self.linear_pos, pos_emb).transpose(1, 2)   # (1, h, 2t-1, d)
        ac = torch.matmul((q + self.pos_bias_u).transpose(1, 2), k.transpose(-2, -1))
        bd = torch.matmul((q + self.pos_bias_v).transpose(1, 2), p.transpose(-2, -1))
        if ac.shape != bd.shape:
            bd = self.rel_shift(bd)
        scores = (ac + bd) / math.sqrt(self.d_k)
        if mask.size(2) > 0:
            blocked = mask.unsqueeze(1).eq(0)[:, :, :, : scores.size(-1)]
            attn = torch.softmax(scores.masked_fill(blocked, -float('inf')), dim=-1).masked_fill(blocked, 0.0)
        else:
            attn = torch.softmax(scores, dim=-1)
        ctx = torch.matmul(self.dropout(attn), v).transpose(1, 2).contiguous().view(query.size(0), -1, self.h * self.d_k)
        return self.linear_out(ctx), new_cache


class PositionwiseFeedForward(nn.Module):
    def __init__(self, idim, hidden_units, dropout_rate, activation):
        super().__init__()
        self.w_1 = nn.Linear(idim, hidden_units)
        self.activation = activation
        self.dropout = nn.Dropout(dropout_rate)
        self.w_2 = nn.Linear(hidden_units, idim)

    def forward(self, xs):
        return self.w_2(self.dropout(self.activation(self.w_1(xs))))


class ConformerEncoderLayer(nn.Module):
    def __init__(self, size, self_attn, feed_forward, feed_forward_macaron, conv_module, dropout_rate,
                 normalize_before=True):
        super().__init__()
        if feed_forward_macaron is not None or conv_module is not None:
            raise NotImplementedError("macaron / convolution branches are not part of the CosyVoice-300M flow encoder")
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.feed_forward_macaron = None
        self.conv_module = None
        self.norm_ff = nn.LayerNorm(size, eps=1e-5)
        self.norm_mha = nn.LayerNorm(size, eps=1e-5)
        self.ff_scale = 1.0
        self.dropout = nn.Dropout(dropout_rate)
        self.normalize_before = normalize_before

    def forward(self, x, mask, pos_emb, mask_pad=None, att_cache=torch.zeros((0, 0, 0, 0)), cnn_cache=None):
        y = self.norm_mha(x) if self.normalize_before else x
        att, new_att_cache = self.self_attn(y, y, y, mask, pos_emb, cache=att_cache)
        x = x + self.dropout(att)
        y = self.norm_ff(x) if self.normalize_before else x
        x = x + self.ff_scale * self.dropout(self.feed_forward(y))
        return x, mask, new_att_cache, torch.zeros((0, 0, 0), dtype=x.dtype, device=x.device)


class ConformerEncoder(nn.Module):
    def __init__(self, input_size: int, output_size: int = 256, attention_heads: int = 4, linear_units: int = 2048,
                 num_blocks: int = 6, dropout_rate: float = 0.1, positional_dropout_rate: float = 0.1,
                 attention_dropout_rate: float = 0.0, normalize_before: bool = True, cnn_module_kernel: int = 15,
                 use_cnn_module: bool = True, macaron_style: bool = True, causal: bool = False):
        super().__init__()
        if use_cnn_module or macaron_style:
            raise NotImplementedError("only use_cnn_module=False, macaron_style=False (CosyVoice-300M flow) is built")
        self._output_size = output_size
        self.embed = LinearNoSubsampling(input_size, output_size, dropout_rate,
                                         RelPositionalEncoding(output_size, positional_dropout_rate))
        self.normalize_before = normalize_before
        self.after_norm = nn.LayerNorm(output_size, eps=1e-5)
        act = nn.SiLU()
        self.encoders = nn.ModuleList([
            ConformerEncoderLayer(output_size,
                                  RelPositionMultiHeadedAttention(attention_heads, output_size, attention_dropout_rate),
                                  PositionwiseFeedForward(output_size, linear_units, dropout_rate, act),
                                  None, None, dropout_rate, normalize_before)
            for _ in range(num_blocks)])

    def output_size(self) -> int:
        return self._output_size

    def forward(self, xs, xs_lens, decoding_chunk_size=0, num_decoding_left_chunks=-1):
        masks = ~make_pad_mask(xs_lens, xs.size(1)).unsqueeze(1)           # (b, 1, t)
        xs, pos_emb, masks = self.embed(xs, masks)
        empty = masks.sum(dim=-1, keepdim=True) == 0                       # fully padded rows attend everywhere
        attn_masks = masks | empty                                         # (no host synchronisation: graph-capturable)
        for layer in self.encoders:
            xs, attn_masks, _, _ = layer(xs, attn_masks, pos_emb, masks)
        if self.normalize_before:
            xs = self.after_norm(xs)
        return xs, masks


class InterpolateRegulator(nn.Module):
    """Token-rate -> mel-rate: linear interpolation to the target length, then a small conv stack
    (reference modules.py:800-837). Length arithmetic is integer and exact."""

    def __init__(self, channels: int, sampling_ratios: Tuple, out_channels: int = None, groups: int = 1):
        super().__init__()
        self.sampling_ratios = sampling_ratios
        out_channels = out_channels or channels
        layers = []
        for _ in sampling_ratios:
            layers += [nn.Conv1d(channels, channels, 3, 1, 1), nn.GroupNorm(groups, channels), nn.Mish()]
        layers.append(nn.Conv1d(channels, out_channels, 1, 1))
        self.model = nn.Sequential(*layers)

    def forward(self, x, ylens=None):
        if x.is_cuda:      # csrc/regulator.cu through the C ABI (five fused conv launches); CPU tensors: host logic / tests
            from . import _path_inputs as PI
            return PI.regulate(self, x, int(ylens.max()), lens=ylens), ylens
        keep = (~make_pad_mask(ylens)).to(x).unsqueeze(-1)
        x = F.interpolate(x.transpose(1, 2).contiguous(), size=ylens.max(), mode='linear')
        return self.model(x).transpose(1, 2).contiguous() * keep, ylens

    def inference(self, x1, x2, mel_len1, mel_len2, input_frame_rate=50):
        if x2.is_cuda:
            from . import _path_inputs as PI
            segs = PI.inference_segments(x1.shape[1], x2.shape[1], mel_len1, mel_len2, input_frame_rate)
            x = torch.concat([x1, x2], dim=1) if x1.shape[1] != 0 else x2
            return PI.regulate(self, x, mel_len1 + mel_len2, segs=segs), mel_len1 + mel_len2
        up = lambda t, n: F.interpolate(t.transpose(1, 2).contiguous(), size=n, mode='linear')
        edge = int(20 / input_frame_rate * 22050 / 256)
        if x2.shape[1] > 40:
            x2 = torch.concat([up(x2[:, :20], edge), up(x2[:, 20:-20], mel_len2 - edge * 2), up(x2[:, -20:], edge)], dim=2)
        else:
            x2 = up(x2, mel_len2)
        x = torch.concat([up(x1, mel_len1), x2], dim=2) if x1.shape[1] != 0 else x2
        return self.model(x).transpose(1, 2).contiguous(), mel_len1 + mel_len2
