"""MultiheadAttention with the reference's constructor, parameters and outputs
(reference: transformer/attention.py), computed by the sm_100a attention / GEMM kernels."""
import torch
from torch import nn

from tts_b200 import _native as N
from tts_b200 import ops


def split_heads(x, num_heads):
    """[B,T,C] -> [B,H,T,C/H] view (reference attention.py:6-15).  The kernels address heads by
    stride and never materialise this; kept for API compatibility."""
    assert x.shape[-1] % num_heads == 0, str(x.shape)
    b, t, c = x.shape
    return x.view(b, t, num_heads, c // num_heads).transpose(1, 2)


def combine_heads(x):
    """[B,H,T,c] -> [B,T,H*c] (reference attention.py:18-26)."""
    b, h, t, c = x.shape
    return x.transpose(1, 2).reshape(b, t, h * c)


def _mask_from_bias(bias, batch, tq, tk):
    """Recover (causal, key_len) from an additive bias built by common.attention_bias."""
    if bias is None:
        return False, None
    if bias.dim() == 4 and bias.shape[0] == 1 and tuple(bias.shape[-2:]) == (tq, tk) and tq == tk:
        want = torch.ones(tq, tk, device=bias.device).triu_(1) * -1e20
        if torch.equal(bias.view(tq, tk).float(), want):
            return True, None
    if bias.dim() == 4 and bias.shape[0] == batch and tuple(bias.shape[1:3]) == (1, 1) and bias.shape[-1] == tk:
        open_ = bias.view(batch, tk) == 0
        key_len = open_.sum(-1).to(torch.int32)
        prefix = torch.arange(tk, device=bias.device)[None, :] < key_len[:, None]
        if torch.equal(open_, prefix):
            return False, key_len
    raise NotImplementedError("tts_b200 attention supports the causal and key-padding biases of "
                              "common.attention_bias, got a bias of shape %s" % (tuple(bias.shape),))


class MultiheadAttention(nn.Module):
    def __init__(self, key_size, value_size, is_self_attention, num_heads, dropout_rate=0.1):
        super().__init__()
        assert key_size % num_heads == 0, "key_size=%d, num_heads=%d" % (key_size, num_heads)
        assert value_size % num_heads == 0, "value_size=%d, num_heads=%d" % (value_size, num_heads)
        if is_self_attention:
            self.qkv_transform = nn.Linear(key_size, 2 * key_size + value_size, bias=False)
        else:
            self.q_transform = nn.Linear(key_size, key_size, bias=False)
            self.kv_transform = nn.Linear(key_size, key_size + value_size, bias=False)
        self.output_transform = nn.Linear(key_size, key_size, bias=False)
        self.attn_dropout = nn.Dropout(dropout_rate)
        self.num_heads, self.key_size, self.value_size = num_heads, key_size, value_size
        self.is_self_attention = is_self_attention

    def forward(self, queries, memories, bias):
        """-> {"outputs": [B,Tq,C], "align": [B,H,Tk,Tq]} (reference attention.py:94-122)."""
        if torch.is_grad_enabled() and self.training and (
                queries.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("tts_b200: MultiheadAttention has no backward yet; call under torch.no_grad()")
        if self.key_size != self.value_size:
            raise NotImplementedError("tts_b200: key_size != value_size is not supported")
        dev = queries.device
        b, tq, c = queries.shape
        h, dh = self.num_heads, self.key_size // self.num_heads
        q2d = N.f32c(queries).view(b * tq, c)
        if memories is None:
            tk = tq
            qkv = ops.linear(q2d, self.qkv_transform.weight)
            base = qkv.data_ptr()
            qp, kp, vp, ldq, ldk = base, base + 4 * c, base + 8 * c, 3 * c, 3 * c
            keep = qkv
        else:
            tk = memories.shape[1]
            qb = ops.linear(q2d, self.q_transform.weight)
            kv = ops.linear(N.f32c(memories).view(b * tk, c), self.kv_transform.weight)
            qp, kp, vp, ldq, ldk = qb.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * c, c, 2 * c
            keep = (qb, kv)
        causal, key_len = _mask_from_bias(bias, b, tq, tk)
        ctx, align = ops.attention(qp, ldq, kp, ldk, vp, ldk, b, h, tq, tk, dh, causal, key_len, True, dev)
        del keep
        out = ops.linear(ctx, self.output_transform.weight).view(b, tq, c)
        return {"outputs": out, "align": align.transpose(2, 3)}
