"""Tensor-level wrappers over the C ABI (one Python function per kernel entry point).

Every function takes CUDA fp32 tensors, launches on torch's current stream and returns torch
tensors that own the output memory.  No function here computes anything in torch.
"""
import ctypes as C
import os

import torch

from . import _native as N

LN_EPS = 1e-6  # transformer/modules.py:36,88
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2


def _i32(t):
    """Lengths arrive as int64 from the feeder (dataloader.py:499-500) or int32 from eval_batch
    (synthesize.py:23); the kernels read int32."""
    if t is None:
        return None
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def linear(x, w, bias=None, act=ACT_NONE, residual=None, alpha=1.0, out=None, row_len=None,
           rows_per_batch=0):
    """out[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias) + residual   (nn.Linear call sites)."""
    lib = N.load()
    M, K = x.shape
    Nn = w.shape[0]
    assert w.shape[1] == K and x.is_contiguous() and w.is_contiguous()
    if out is None:
        out = torch.empty((M, Nn), device=x.device, dtype=torch.float32)
    epi = N.GemmEpilogue()
    epi.alpha = alpha
    epi.bias = N.ptr(bias)
    epi.act = act
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
        epi.residual = N.ptr(residual)
        epi.ldr = Nn
    if row_len is not None:
        epi.row_len = N.ptr(row_len)
        epi.rows_per_batch = rows_per_batch
    N.check(lib.tts_gemm_nt(N.ptr(x), K, N.ptr(w), K, N.ptr(out), out.stride(0), M, Nn, K, C.byref(epi),
                            N.stream_ptr(x.device)), "gemm_nt")
    return out


def cross_kv(memory2d, w_kv, batch, seq, n_heads, out_k, out_v):
    """K/V of the encoder memory in cache layout [B][H][S][dh] (attention.py:66-68 + split_heads)."""
    lib = N.load()
    D = memory2d.shape[1]
    epi = N.GemmEpilogue()
    epi.alpha = 1.0
    epi.rows_per_batch = seq
    epi.head_dim = D // n_heads
    epi.n_heads = n_heads
    epi.head_rows = seq
    epi.out_v = N.ptr(out_v)
    N.check(lib.tts_gemm_nt(N.ptr(memory2d), D, N.ptr(w_kv), D, N.ptr(out_k), 0, batch * seq, 2 * D, D,
                            C.byref(epi), N.stream_ptr(memory2d.device)), "gemm_nt(cross_kv)")


def conv5(xpad, w_packed, scale, shift, act, row_len, batch, frames, out, out_padded, residual=None):
    """One Postnet layer: Conv1d(k=5, pad=2) + folded BatchNorm + tanh + length mask, as a GEMM over
    the zero-padded channels-last buffer xpad [B][T+4][Cin] (tacotron.py:83-88)."""
    lib = N.load()
    cin = xpad.shape[-1]
    cout = w_packed.shape[0]
    assert w_packed.shape[1] == 5 * cin and xpad.is_contiguous()
    epi = N.GemmEpilogue()
    epi.alpha = 1.0
    epi.scale, epi.shift = N.ptr(scale), N.ptr(shift)
    epi.act = act
    epi.row_len = N.ptr(row_len)
    epi.rows_per_batch = frames + 4
    epi.valid_rows = frames
    epi.out_rows_per_batch = frames + 4 if out_padded else frames
    epi.out_row_offset = 2 if out_padded else 0
    if residual is not None:  # indexed by OUTPUT row, so only meaningful for the unpadded last layer
        assert not out_padded and residual.is_contiguous()
        epi.residual = N.ptr(residual)
        epi.ldr = cout
    M = batch * (frames + 4) - 4
    N.check(lib.tts_gemm_nt(N.ptr(xpad), cin, N.ptr(w_packed), 5 * cin, N.ptr(out), cout, M, cout, 5 * cin,
                            C.byref(epi), N.stream_ptr(xpad.device)), "gemm_nt(conv5)")
    return out


def layernorm(x, gamma, beta, row_len=None, rows_per_batch=0, out=None):
    lib = N.load()
    rows, ch = x.shape
    if out is None:
        out = torch.empty_like(x)
    N.check(lib.tts_layernorm(N.ptr(x), N.ptr(out), N.ptr(gamma), N.ptr(beta), rows, ch, LN_EPS, N.ptr(row_len),
                              rows_per_batch, N.stream_ptr(x.device)), "layernorm")
    return out


def embed_pe(ids, lengths, table, pe, pe_scale, batch, seq):
    """ids [B,S] int64 + embedding table, or ids=None and table = already embedded rows [B*S,C]."""
    lib = N.load()
    ch = table.shape[1]
    assert table.is_contiguous() and (ids is None or ids.is_contiguous())
    out = torch.empty((batch * seq, ch), device=table.device, dtype=torch.float32)
    N.check(lib.tts_embed_pe(N.ptr(ids), N.ptr(lengths), N.ptr(table), N.ptr(pe), N.ptr(pe_scale),
                             N.ptr(out), batch, seq, ch, table.shape[0], N.stream_ptr(table.device)), "embed_pe")
    return out


def shift_pe(pre, lengths, pe, pe_scale, batch, frames):
    lib = N.load()
    ch = pre.shape[-1]
    out = torch.empty((batch * frames, ch), device=pre.device, dtype=torch.float32)
    N.check(lib.tts_shift_pe(N.ptr(pre), N.ptr(lengths), N.ptr(pe), N.ptr(pe_scale), N.ptr(out), batch, frames, ch,
                             N.stream_ptr(pre.device)), "shift_pe")
    return out


def pad_rows(x, lengths, batch, frames, pad=2):
    lib = N.load()
    ch = x.shape[-1]
    out = torch.empty((batch, frames + 2 * pad, ch), device=x.device, dtype=torch.float32)
    N.check(lib.tts_pad_rows(N.ptr(x), N.ptr(lengths), N.ptr(out), batch, frames, ch, pad,
                             N.stream_ptr(x.device)), "pad_rows")
    return out


def cond_embed(mem, col_offset, w2, b2, *, vec=None, w1=None, ids=None):
    lib = N.load()
    B, S, width = mem.shape
    emb = w2.shape[0]
    vec_dim = 0 if vec is None else vec.shape[1]
    N.check(lib.tts_cond_embed(N.ptr(vec), vec_dim, N.ptr(ids), N.ptr(w1), N.ptr(w2), N.ptr(b2), emb, N.ptr(mem), B,
                               S, width, col_offset, N.stream_ptr(mem.device)), "cond_embed")


def attention(q, ldq, k, ldk, v, ldv, batch, n_heads, tq, tk, head_dim, causal, key_len, want_align, device):
    """q/k/v are (data_ptr, row stride) views into packed projection buffers; returns
    (ctx [B*Tq, H*dh], align [B,H,Tq,Tk] or None)."""
    lib = N.load()
    ctx = torch.empty((batch * tq, n_heads * head_dim), device=device, dtype=torch.float32)
    align = torch.empty((batch, n_heads, tq, tk), device=device, dtype=torch.float32) if want_align else None
    N.check(lib.tts_attention(q, ldq, k, ldk, v, ldv, N.ptr(ctx), N.ptr(align), batch, n_heads, tq, tk, head_dim,
                              float(head_dim) ** -0.5, 1 if causal else 0, N.ptr(key_len),
                              N.stream_ptr(device)), "attention")
    return ctx, align


# ---------------------------------------------------------------------------------------------------------------------
# training path (bf16 operands, fp32 accumulation): include/tts_b200.h "teacher-forced TRAINING path"
# ---------------------------------------------------------------------------------------------------------------------
_GEMM_LOG = os.environ.get("TTS_GEMM_LOG")


def gemm_bf16(a, b, *, a_mn=False, b_mn=False, m=None, n=None, k=None, out=None, out_dtype=torch.bfloat16, bias=None,
              act=ACT_NONE, alpha=1.0, residual=None, drop_p=0.0, seed=0, rng_stream=0, gate=None, gate_scale=1.0,
              taps=1, split_k=1, row_len=None, rows_per_batch=0, valid_rows=0, out_rows_per_batch=0, out_row_offset=0,
              out_rows=None, a_rows=0):
    """C[M,N] = epilogue(A . B^T) on the tcgen05 bf16 kernel.  a / b are 2-D bf16 tensors (inner stride 1): K-major
    operands are [rows][K], MN-major ones ([a_mn] / [b_mn]) are [K][rows].  Returns the output tensor."""
    lib = N.load()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(-1) == 1 and b.stride(-1) == 1
    M = m if m is not None else (a.shape[1] if a_mn else a.shape[0])
    Nn = n if n is not None else (b.shape[1] if b_mn else b.shape[0])
    K = k if k is not None else (a.shape[0] if a_mn else a.shape[1])
    if out is None:
        rows = out_rows if out_rows is not None else M
        if split_k > 1:
            out = torch.zeros((rows, Nn), device=a.device, dtype=torch.float32)
        else:
            out = torch.empty((rows, Nn), device=a.device, dtype=out_dtype)
    g = N.GemmBf16()
    g.A, g.lda, g.a_mn_major, g.a_rows = a.data_ptr(), a.stride(0), 1 if a_mn else 0, a_rows
    g.B, g.ldb, g.b_mn_major = b.data_ptr(), b.stride(0), 1 if b_mn else 0
    g.C, g.ldc, g.out_bf16 = out.data_ptr(), out.stride(0), 1 if out.dtype == torch.bfloat16 else 0
    g.M, g.N, g.K, g.taps, g.split_k = M, Nn, K, taps, split_k
    g.bias, g.act, g.alpha = N.ptr(bias), act, alpha
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    g.drop_p, g.seed, g.rng_stream = drop_p, seed, rng_stream
    if gate is not None:
        assert gate.dtype == torch.bfloat16 and gate.stride(-1) == 1
        g.gate, g.ldg, g.gate_scale = gate.data_ptr(), gate.stride(0), gate_scale
    g.row_len, g.rows_per_batch, g.valid_rows = N.ptr(row_len), rows_per_batch, valid_rows
    g.out_rows_per_batch, g.out_row_offset = out_rows_per_batch, out_row_offset
    if _GEMM_LOG:   # diagnostics: one line per call, to be joined with an ncu launch list by order
        with open(_GEMM_LOG, "a") as f:
            f.write("M=%d N=%d K=%d a_mn=%d b_mn=%d taps=%d split=%d out=%s res=%d drop=%g gate=%d act=%d\n" % (
                M, Nn, K, a_mn, b_mn, taps, split_k, "bf16" if g.out_bf16 else "f32", residual is not None, drop_p,
                gate is not None, act))
    N.check(lib.tts_gemm_bf16(C.byref(g), N.stream_ptr(a.device)), "gemm_bf16")
    return out
