"""Tensor-level wrappers over the training-path entry points of the C ABI (include/tts_b200.h, "teacher-forced
TRAINING path").  Like ops.py: CUDA tensors in, kernels on torch's current stream, no arithmetic in torch."""
import ctypes as C
import os

import torch

from . import _native as N
from .ops import ACT_NONE, ACT_RELU, LN_EPS, _i32, gemm_bf16  # noqa: F401

BF16 = torch.bfloat16
BN_EPS, BN_MOMENTUM = 1e-5, 0.1   # torch.nn.BatchNorm1d defaults (tacotron.py:79)
_ATTN_DETERMINISTIC = os.environ.get("TTS_ATTN_DETERMINISTIC", "0") not in ("", "0")
_ATTN_KEEP_MASK = os.environ.get("TTS_ATTN_KEEP_MASK", "1") not in ("", "0")   # 0: backward regenerates the Philox bits


def _s(t):
    return N.stream_ptr(t.device)


def _scratch(dev, n_floats, cache={}):
    key = (dev.type, dev.index)
    cur = cache.get(key)
    if cur is None or cur.numel() < n_floats:
        cur = torch.empty((int(n_floats),), device=dev, dtype=torch.float32)
        cache[key] = cur
    return cur


# ---- LayerNorm -----------------------------------------------------------------------------------------------------
def ln_fwd(x, gamma, beta, row_len=None, rows_per_batch=0):
    """-> (y bf16 [R,C], mean [R], rstd [R])."""
    R, Cc = x.shape
    y = torch.empty((R, Cc), device=x.device, dtype=BF16)
    mean = torch.empty((R,), device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    N.check(N.load().tts_ln_fwd_train(x.data_ptr(), y.data_ptr(), Cc, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                      rstd.data_ptr(), R, Cc, LN_EPS, N.ptr(row_len), rows_per_batch, _s(x)), "ln_fwd_train")
    return y, mean, rstd


def ln_bwd(dy, x, mean, rstd, gamma, dres=None, row_len=None, rows_per_batch=0, cast_drop=None):
    """dy bf16 [R,C] -> (dx fp32 [R,C] (+ dres), dgamma [C], dbeta [C]).  cast_drop = (p, seed, stream): a fourth result
    dyb = dropout_cast(dx, p, seed, stream) (bf16) comes out of the same kernel."""
    R, Cc = x.shape
    lib = N.load()
    dx = torch.empty_like(x)
    dg = torch.empty((Cc,), device=x.device, dtype=torch.float32)
    db = torch.empty_like(dg)
    scratch = _scratch(x.device, lib.tts_ln_bwd_scratch_floats(Cc))
    dyb = torch.empty((R, Cc), device=x.device, dtype=BF16) if cast_drop is not None else None
    p, seed, stream = cast_drop if cast_drop is not None else (0.0, 0, 0)
    N.check(lib.tts_ln_bwd_train(dy.data_ptr(), dy.stride(0), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                 N.ptr(dres), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), scratch.data_ptr(), R, Cc,
                                 N.ptr(row_len), rows_per_batch, N.ptr(dyb), Cc, p, seed, stream, _s(x)), "ln_bwd_train")
    if cast_drop is not None:
        return dx, dg, db, dyb
    return dx, dg, db


def dropout_cast(src, drop_p=0.0, seed=0, stream=0, row_len=None, rows_per_batch=0, out=None):
    """fp32 [R,C] -> bf16 [R,C]: backward of an epilogue dropout (or a plain cast when drop_p == 0)."""
    R, Cc = src.shape
    if out is None:
        out = torch.empty((R, Cc), device=src.device, dtype=BF16)
    N.check(N.load().tts_dropout_cast(src.data_ptr(), src.stride(0), out.data_ptr(), out.stride(0), R, Cc, drop_p, seed, stream,
                                      N.ptr(row_len), rows_per_batch, _s(src)), "dropout_cast")
    return out


class MultiTable:
    """Device table for the multi-tensor kernels (one launch over many tensors)."""

    def __init__(self, dev):
        self.dev = dev
        self.chunk = int(N.load().tts_multi_chunk_elems())
        self.table = None
        self.n_entries = 0
        self.n_chunks = 0

    def _upload(self, rows, words_per_entry):
        """Pointer table -> device without blocking the host: a pageable `.to(device)` waits for everything queued on the
        stream (with freshly allocated gradients the optimizer's table is rebuilt every step, and that wait serialised
        host and GPU: 48 ms of the 57 ms `FusedAdam.step` spent on the host, tests/tools_train_hostprof.py).  Two pinned
        staging buffers alternate; a buffer is reused only after the copy that read it has completed."""
        import numpy as np
        arr = np.array(rows, dtype=np.int64).reshape(-1, words_per_entry)
        n = arr.size
        if not hasattr(self, "_stage"):
            self._stage, self._stage_ev, self._stage_i = [None, None], [None, None], 0
        i = self._stage_i
        self._stage_i ^= 1
        if self._stage[i] is None or self._stage[i].numel() < n:
            self._stage[i] = torch.empty((max(n, 2048),), dtype=torch.int64).pin_memory()
            self._stage_ev[i] = torch.cuda.Event()
        else:
            self._stage_ev[i].synchronize()
        self._stage[i][:n].copy_(torch.from_numpy(arr).view(-1))
        if self.table is None or self.table.numel() != n:
            self.table = torch.empty(arr.shape, dtype=torch.int64, device=self.dev)
        self.table.view(-1).copy_(self._stage[i][:n], non_blocking=True)
        self._stage_ev[i].record(torch.cuda.current_stream(self.dev))
        self.n_entries = arr.shape[0]

    def build_cast(self, pairs):
        """pairs: [(src fp32, dst bf16)] contiguous tensors."""
        rows, chunk0 = [], 0
        for src, dst in pairs:
            assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
            rows.append([src.data_ptr(), dst.data_ptr(), src.numel(), chunk0])
            chunk0 += (src.numel() + self.chunk - 1) // self.chunk
        self.n_chunks = chunk0
        self._upload(rows, 4)

    def build_opt(self, entries):
        """entries: [(p, g, m, v, decay)]; g / m / v may be None for the L2 (sum of squares) table."""
        rows, chunk0 = [], 0
        for p, g, m, v, decay in entries:
            assert p.is_contiguous()
            rows.append([p.data_ptr(), 0 if g is None else g.data_ptr(), 0 if m is None else m.data_ptr(),
                         0 if v is None else v.data_ptr(), p.numel(), chunk0, int(bool(decay))])
            chunk0 += (p.numel() + self.chunk - 1) // self.chunk
        self.n_chunks = chunk0
        # struct OptEntry: 4 pointers, n, first_chunk, {int decay; int pad} packed in one 8-byte word
        self._upload(rows, 7)


def multi_cast(table):
    N.check(N.load().tts_multi_cast_bf16(table.table.data_ptr(), table.n_entries, table.n_chunks, N.stream_ptr(table.dev)),
            "multi_cast_bf16")


def sumsq_multi(table, out):
    N.check(N.load().tts_sumsq_multi(table.table.data_ptr(), table.n_entries, table.n_chunks, out.data_ptr(),
                                     N.stream_ptr(table.dev)), "sumsq_multi")


def adam_multi(table, lr, beta1, beta2, eps, step, reg_weight=0.0, grad_scale=1.0):
    N.check(N.load().tts_adam_multi(table.table.data_ptr(), table.n_entries, table.n_chunks, lr, beta1, beta2, eps, step,
                                    reg_weight, grad_scale, N.stream_ptr(table.dev)), "adam_multi")


# ---- prologues -----------------------------------------------------------------------------------------------------
def embed_fwd(ids, lengths, embed, pe, pe_scale, B, S, drop_p, seed, stream):
    Cc = embed.shape[1]
    out = torch.empty((B * S, Cc), device=embed.device, dtype=torch.float32)
    N.check(N.load().tts_embed_train_fwd(ids.data_ptr(), lengths.data_ptr(), embed.data_ptr(), pe.data_ptr(), pe_scale.data_ptr(),
                                         out.data_ptr(), B, S, Cc, embed.shape[0], drop_p, seed, stream, _s(embed)), "embed_train_fwd")
    return out


def embed_bwd(dx, ids, lengths, pe, vocab, B, S, drop_p, seed, stream):
    Cc = dx.shape[1]
    d_embed = torch.zeros((vocab, Cc), device=dx.device, dtype=torch.float32)
    d_scale = torch.zeros((), device=dx.device, dtype=torch.float32)
    N.check(N.load().tts_embed_train_bwd(dx.data_ptr(), ids.data_ptr(), lengths.data_ptr(), pe.data_ptr(), d_embed.data_ptr(),
                                         d_scale.data_ptr(), B, S, Cc, vocab, drop_p, seed, stream, _s(dx)), "embed_train_bwd")
    return d_embed, d_scale


def shift_pe_fwd(pre, lengths, pe, pe_scale, B, T, drop_p, seed, stream):
    Cc = pre.shape[1]
    out = torch.empty((B * T, Cc), device=pre.device, dtype=torch.float32)
    N.check(N.load().tts_shift_pe_train_fwd(pre.data_ptr(), lengths.data_ptr(), pe.data_ptr(), pe_scale.data_ptr(), out.data_ptr(),
                                            B, T, Cc, drop_p, seed, stream, _s(pre)), "shift_pe_train_fwd")
    return out


def shift_pe_bwd(dx, lengths, pe, B, T, drop_p, seed, stream):
    Cc = dx.shape[1]
    dpre = torch.empty((B * T, Cc), device=dx.device, dtype=BF16)
    d_scale = torch.zeros((), device=dx.device, dtype=torch.float32)
    N.check(N.load().tts_shift_pe_train_bwd(dx.data_ptr(), lengths.data_ptr(), pe.data_ptr(), dpre.data_ptr(), d_scale.data_ptr(),
                                            B, T, Cc, drop_p, seed, stream, _s(dx)), "shift_pe_train_bwd")
    return dpre, d_scale


# ---- reductions ----------------------------------------------------------------------------------------------------
def colsum(x, row_weight=None):
    """out[c] = sum_r w[r] x[r][c] for a bf16 [R,C] tensor."""
    R, Cc = x.shape
    out = torch.zeros((Cc,), device=x.device, dtype=torch.float32)
    N.check(N.load().tts_colsum_bf16(x.data_ptr(), x.stride(0), N.ptr(row_weight), out.data_ptr(), R, Cc, _s(x)), "colsum_bf16")
    return out


def sum_f32(x):
    out = torch.zeros((), device=x.device, dtype=torch.float32)
    N.check(N.load().tts_sum_f32(x.data_ptr(), x.numel(), out.data_ptr(), _s(x)), "sum_f32")
    return out


def rowdot(x, w, bias, row_len, rows_per_batch):
    R, K = x.shape
    out = torch.empty((R,), device=x.device, dtype=torch.float32)
    N.check(N.load().tts_rowdot_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), N.ptr(bias), N.ptr(row_len), rows_per_batch,
                                     out.data_ptr(), R, K, _s(x)), "rowdot_bf16")
    return out


# ---- attention -----------------------------------------------------------------------------------------------------
def _attn_struct(q, k, v, out, lse, B, H, Tq, Tk, dh, causal, key_len, drop_p, seed, stream):
    a = N.AttnTrain()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
    a.out, a.ldo, a.lse = out.data_ptr(), out.stride(0), lse.data_ptr()
    a.batch, a.n_heads, a.tq, a.tk, a.head_dim, a.causal = B, H, Tq, Tk, dh, 1 if causal else 0
    a.key_len = N.ptr(key_len)
    a.drop_p, a.seed, a.rng_stream = drop_p, seed, stream
    return a


def attn_fwd(q, k, v, B, H, Tq, Tk, dh, causal, key_len, drop_p=0.0, seed=0, stream=0, keep_mask=True):
    """q [B*Tq, >=H*dh], k / v [B*Tk, ...] bf16 2-D views (row stride = ld) -> (ctx bf16 [B*Tq, H*dh], lse [B,H,Tq]).
    keep_mask (with drop_p > 0): the forward kernel also stores its dropout keep bits, 1 bit per attention weight, and
    the single-pass backward reads them instead of running Philox again; the words travel as `lse.keep_mask`."""
    ctx = torch.empty((B * Tq, H * dh), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, Tq), device=q.device, dtype=torch.float32)
    a = _attn_struct(q, k, v, ctx, lse, B, H, Tq, Tk, dh, causal, key_len, drop_p, seed, stream)
    if keep_mask and drop_p > 0.0 and _ATTN_KEEP_MASK:
        lse.keep_mask = torch.empty((B * H, N.load().tts_attn_keep_words(Tk), Tq), device=q.device, dtype=torch.int32)
        a.keep_mask = lse.keep_mask.data_ptr()
    N.check(N.load().tts_attn_train_fwd(C.byref(a), _s(q)), "attn_train_fwd")
    return ctx, lse


def attn_bwd(q, k, v, ctx, lse, d_ctx, dq, dk, dv, B, H, Tq, Tk, dh, causal, key_len, drop_p=0.0, seed=0, stream=0,
             deterministic=None):
    """Writes dq / dk / dv (bf16 2-D views with their own row strides).  Default: the single-pass kernel (dQ summed with fp32
    atomics into a scratch buffer); deterministic=True (or TTS_ATTN_DETERMINISTIC=1): the two-kernel path."""
    a = _attn_struct(q, k, v, ctx, lse, B, H, Tq, Tk, dh, causal, key_len, drop_p, seed, stream)
    delta = torch.empty((B, H, Tq), device=q.device, dtype=torch.float32)
    if deterministic is None:
        deterministic = _ATTN_DETERMINISTIC
    if not deterministic:
        a.dq_acc = _scratch(q.device, B * Tq * H * dh).data_ptr()
        km = getattr(lse, "keep_mask", None)
        if km is not None and drop_p > 0.0:
            a.keep_mask = km.data_ptr()
    a.d_out, a.lddo, a.delta = d_ctx.data_ptr(), d_ctx.stride(0), delta.data_ptr()
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.lddq, a.lddk, a.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    N.check(N.load().tts_attn_train_bwd(C.byref(a), _s(q)), "attn_train_bwd")


# ---- Postnet BatchNorm ---------------------------------------------------------------------------------------------
def bn_fwd(z, gamma, beta, running_mean, running_var, num_batches, act_tanh, drop_p, seed, stream, lengths, B, T,
           out_pad=None, out_f32=None, residual=None):
    Cc = z.shape[1]
    lib = N.load()
    mean = torch.empty((Cc,), device=z.device, dtype=torch.float32)
    invstd = torch.empty_like(mean)
    scratch = _scratch(z.device, lib.tts_bn_scratch_floats(Cc))
    N.check(lib.tts_bn_train_fwd(z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                 N.ptr(running_mean), N.ptr(running_var), N.ptr(num_batches), BN_MOMENTUM, BN_EPS,
                                 1 if act_tanh else 0, drop_p, seed, stream, N.ptr(lengths), B, T, Cc, N.ptr(out_pad),
                                 N.ptr(out_f32), N.ptr(residual), scratch.data_ptr(), _s(z)), "bn_train_fwd")
    return mean, invstd


def bn_bwd(z, dout, gamma, beta, mean, invstd, act_tanh, drop_p, seed, stream, lengths, mask_rows, B, T, dz_pad):
    Cc = z.shape[1]
    lib = N.load()
    dg = torch.empty((Cc,), device=z.device, dtype=torch.float32)
    db = torch.empty_like(dg)
    scratch = _scratch(z.device, lib.tts_bn_scratch_floats(Cc))
    N.check(lib.tts_bn_train_bwd(z.data_ptr(), dout.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                 1 if act_tanh else 0, drop_p, seed, stream, N.ptr(lengths), 1 if mask_rows else 0, B, T, Cc,
                                 dz_pad.data_ptr(), dg.data_ptr(), db.data_ptr(), scratch.data_ptr(), _s(z)), "bn_train_bwd")
    return dg, db


def pad_cast(x, lengths, B, T, out=None, only_pads=False):
    """[B*T, C] fp32 -> masked zero-padded bf16 [B, T+4, C]; only_pads: just zero the 4 pad rows of an existing buffer."""
    Cc = out.shape[-1] if x is None else x.shape[-1]
    if out is None:
        out = torch.empty((B, T + 4, Cc), device=x.device, dtype=BF16)
    N.check(N.load().tts_pad_cast_bf16(N.ptr(x), N.ptr(lengths), out.data_ptr(), B, T, Cc, 1 if only_pads else 0, _s(out)), "pad_cast_bf16")
    return out


# ---- loss ----------------------------------------------------------------------------------------------------------
def loss_fwd(mel_bef, mel_aft, stop, targets, lengths, total_len, want_grads=True):
    """-> (sums3 [4] = {sum mse_bef, sum mse_aft, sum bce, 0}, aft_per_sample [B], d_bef, d_aft, d_stop)."""
    B, T, M = targets.shape
    dev = targets.device
    sums = torch.empty((4,), device=dev, dtype=torch.float32)
    aft_b = torch.empty((B,), device=dev, dtype=torch.float32)
    d_bef = torch.empty_like(mel_bef) if want_grads else None
    d_aft = torch.empty_like(mel_aft) if want_grads else None
    d_stop = torch.empty_like(stop) if want_grads else None
    N.check(N.load().tts_loss_train(mel_bef.data_ptr(), mel_aft.data_ptr(), stop.data_ptr(), targets.data_ptr(), lengths.data_ptr(),
                                    total_len.data_ptr(), B, T, M, 5.0, sums.data_ptr(), aft_b.data_ptr(), N.ptr(d_bef), N.ptr(d_aft),
                                    N.ptr(d_stop), _s(targets)), "loss_train")
    return sums, aft_b, d_bef, d_aft, d_stop
