"""Helpers with the reference's names and semantics (reference: transformer/common.py).

Only `impute`, `mask_reduce` and the initialisers are used by host-side code (loss, init); the
position table and the attention masks are produced inside the CUDA kernels from indices and
lengths, so `get_sinusoid_encoding_table` / `attention_bias` exist for API compatibility and
for callers that drive `MultiheadAttention` directly.
"""
import math

import numpy as np
import torch

from tts_b200.engine import sinusoid_table


def get_sinusoid_encoding_table(length, channels, min_timescale=1, max_timescale=1e4):
    """[length, channels] fp32 table, sin half then cos half (reference common.py:4-29)."""
    if min_timescale != 1 or max_timescale != 1e4:
        half = channels // 2
        inc = math.log(float(max_timescale) / float(min_timescale)) / (half - 1)
        inv = min_timescale * np.exp(np.arange(half) * -inc)
        ang = np.arange(length)[:, None] * inv[None, :]
        tab = np.concatenate([np.sin(ang), np.cos(ang)], axis=1)
        tab = np.pad(tab, [[0, 0], [0, channels % 2]])
        return torch.FloatTensor(tab)
    return sinusoid_table(length, channels)


def attention_bias(inputs, mode, inf=-1e20):
    """Additive bias with the reference's shapes: "causal" -> [1,1,T,T] from an int T,
    "masking" -> [B,1,1,T] from a bool mask (reference common.py:32-48)."""
    if mode == "causal":
        n = int(inputs)
        out = torch.ones(n, n).triu_(1).mul_(inf).view(1, 1, n, n)
    elif mode == "masking":
        out = ((~inputs.bool()).float() * inf)[:, None, None, :]
    else:
        raise ValueError("Unknown mode %s" % mode)
    return out


def impute(x, lengths, channels_last=True):
    """Zero every position at or beyond its sequence length (reference common.py:51-70)."""
    n = x.shape[1] if channels_last else x.shape[-1]
    keep = torch.arange(n, device=lengths.device)[None, :] < lengths[:, None]
    shape = [x.shape[0]] + [1] * (x.dim() - 1)
    shape[1 if channels_last else -1] = n
    return x * keep.view(shape)


def mask_reduce(loss, lengths, per_sample=False):
    """Mean of a [B,T] tensor over valid positions, overall or per sample (reference common.py:73-88)."""
    kept = impute(loss, lengths)
    return kept.sum(-1) / lengths if per_sample else kept.sum() / lengths.sum()


def truncated_normal(tensor, mean=0, std=0.5):
    """tf.random.truncated_normal look-alike: of 8 normal draws per element keep the first that
    falls within two standard deviations (reference common.py:90-105; the draw order matters for
    seed-for-seed identical weights)."""
    with torch.no_grad():
        draws = tensor.new_empty(tuple(tensor.shape) + (8,)).normal_(mean=mean, std=std)
        inside = (draws < 2 * std) & (draws > -2 * std)
        pick = inside.max(-1, keepdim=True)[1]
        return draws.gather(-1, pick).squeeze(-1)


def variance_scaling_initializer(tensor, factor=2.0):
    """Fan-average variance scaling on top of truncated_normal, std = sqrt(1.3*factor/n)
    (reference common.py:108-124; receptive-field size multiplies both fans)."""
    field = 1
    for d in tensor.shape[2:]:
        field *= d
    n = (tensor.shape[1] * field + tensor.shape[0] * field) / 2
    return truncated_normal(tensor, std=np.sqrt(1.3 * factor / n))
