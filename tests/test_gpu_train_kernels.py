"""GPU unit tests of the training-path kernels (bf16 tcgen05 GEMM, LayerNorm / attention / BatchNorm / loss forward and
backward, fused Adam) against plain PyTorch fp32 references of the same op, through the C ABI.

Tolerances: operands are rounded to bf16 before both the kernel and the reference see them, so what is compared is the
kernel's fp32-accumulated product against an fp64 product of the SAME bf16 values: max-abs <= 2e-3 * sqrt(K / 64) on
O(1) outputs (fp32 accumulation order + one bf16 output rounding where the output is bf16)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from tts_b200 import ops as _ops
    return _ops


def _bf(t):
    return t.to(DEV).to(torch.bfloat16)


def _err(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def _status():
    from tts_b200 import _native
    return _native.load().tts_gemm_bf16_status()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (300, 200, 80), (1000, 768, 768), (4099, 2304, 768),
                                   (515, 80, 768), (64, 1, 768), (2048, 3072, 768), (2048, 768, 3072), (33, 48, 32)])
def test_gemm_bf16_forward_layout(ops, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a, b = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    want = a.double() @ b.double().t()
    got32 = ops.gemm_bf16(a, b, out_dtype=torch.float32)
    assert _err(got32, want) < 2e-4 * max(1.0, (K / 64) ** 0.5), (M, N, K)
    got16 = ops.gemm_bf16(a, b)
    assert got16.dtype == torch.bfloat16 and _err(got16, want) < 3e-2
    assert _status() == 0


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 768, 2304), (777, 80, 256), (4099, 768, 3072), (130, 32, 48)])
def test_gemm_bf16_dgrad_layout(ops, M, N, K):
    """dX[M,N] = dY[M,K] . W[K,N]: the weight is read as the MN-major B operand, no transposed copy."""
    g = torch.Generator().manual_seed(M * 3 + N)
    dy, w = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(K, N, generator=g) / K ** 0.5)
    got = ops.gemm_bf16(dy, w, b_mn=True, out_dtype=torch.float32)
    assert _err(got, dy.double() @ w.double()) < 2e-4 * max(1.0, (K / 64) ** 0.5)
    assert _status() == 0


@pytest.mark.parametrize("R,N,K,split", [(256, 128, 128, 1), (1000, 768, 768, 4), (5000, 2304, 768, 5), (333, 80, 768, 2),
                                         (4099, 48, 32, 3), (2000, 256, 80, 1)])
def test_gemm_bf16_wgrad_layout(ops, R, N, K, split):
    """dW[N,K] = dY[R,N]^T . X[R,K]: both operands MN-major (contraction over the rows), split-K fp32 reduction."""
    g = torch.Generator().manual_seed(R + N)
    dy, x = _bf(torch.randn(R, N, generator=g)), _bf(torch.randn(R, K, generator=g) / R ** 0.5)
    got = ops.gemm_bf16(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=split)
    assert got.shape == (N, K)
    assert _err(got, dy.double().t() @ x.double()) < 2e-4 * max(1.0, (R / 64) ** 0.5)
    assert _status() == 0


def test_gemm_bf16_epilogues(ops):
    g = torch.Generator().manual_seed(9)
    M, N, K = 333, 256, 128
    a, b = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    bias, res = torch.randn(N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    acc = a.double() @ b.double().t()
    got = ops.gemm_bf16(a, b, out_dtype=torch.float32, bias=bias, act=ops.ACT_RELU, residual=res, alpha=0.5)
    assert _err(got, torch.relu(0.5 * acc + bias.double()) + res.double()) < 5e-4
    gate = _bf(torch.randn(M, N, generator=g))
    got = ops.gemm_bf16(a, b, out_dtype=torch.float32, gate=gate, gate_scale=2.0)
    assert _err(got, acc * (gate.double() > 0) * 2.0) < 1e-3
    # dropout: keep rate, scaling, determinism in (seed, stream), independence of the tiling
    p = 0.25
    d1 = ops.gemm_bf16(a, b, out_dtype=torch.float32, drop_p=p, seed=1234, rng_stream=7)
    d2 = ops.gemm_bf16(a, b, out_dtype=torch.float32, drop_p=p, seed=1234, rng_stream=7)
    d3 = ops.gemm_bf16(a, b, out_dtype=torch.float32, drop_p=p, seed=1234, rng_stream=8)
    assert torch.equal(d1, d2) and not torch.equal(d1, d3)
    kept = d1 != 0
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    assert _err(d1[kept], (acc / (1 - p))[kept.cpu()]) < 1e-3
    # length mask (impute): rows at or beyond the length are zero
    lens = torch.tensor([3, 37, 0], dtype=torch.int32, device=DEV)
    got = ops.gemm_bf16(a, b, out_dtype=torch.float32, row_len=lens, rows_per_batch=111)
    mask = (torch.arange(111)[None, :] < lens.cpu()[:, None]).reshape(-1, 1)
    assert _err(got, acc * mask.to(DEV)) < 5e-4
    assert _status() == 0


@pytest.mark.parametrize("B,T,cin,cout", [(2, 37, 80, 48), (3, 200, 48, 80), (2, 130, 512, 512)])
def test_gemm_bf16_conv5_taps(ops, B, T, cin, cout):
    """Conv1d(k=5, pad=2) as 5 accumulated taps over the zero-padded channels-last buffer [B][T+4][cin]."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 5, generator=g) / (5 * cin) ** 0.5).to(torch.bfloat16)
    xpad = torch.zeros(B, T + 4, cin, dtype=torch.bfloat16)
    xpad[:, 2:T + 2] = x
    wp = w.permute(0, 2, 1).reshape(cout, 5 * cin).contiguous()
    want = F.conv1d(x.double().transpose(1, 2), w.double(), padding=2).transpose(1, 2)
    M = B * (T + 4) - 4
    got = ops.gemm_bf16(xpad.to(DEV).view(-1, cin), wp.to(DEV), m=M, k=cin, taps=5, out_dtype=torch.float32,
                        rows_per_batch=T + 4, valid_rows=T, out_rows_per_batch=T, out_rows=B * T, a_rows=B * (T + 4))
    assert _err(got.view(B, T, cout), want) < 2e-3
    assert _status() == 0


# ---------------------------------------------------------------------------------------------------------------------
# Philox4x32-7 restated in numpy: pins csrc/philox.cuh and lets the references apply the SAME dropout masks
# ---------------------------------------------------------------------------------------------------------------------
def _philox(seed, idx, stream):
    import numpy as np
    idx = np.asarray(idx, dtype=np.uint64)
    M32 = np.uint64(0xffffffff)
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64(seed >> 32)
    c0, c1 = idx & M32, idx >> np.uint64(32)
    c2 = np.full_like(idx, stream, dtype=np.uint64)
    c3 = np.full_like(idx, 0x5eed, dtype=np.uint64)
    for _ in range(7):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M32, (k1 + np.uint64(0xBB67AE85)) & M32
    return np.stack([c0, c1, c2, c3], axis=-1)


def _thresh(p):
    return min(int(p * 4294967296.0), 0xffffffff)


def keep_linear(seed, stream, rows, cols, p):
    """keep mask [rows, cols] of a row-major tensor (GEMM epilogues, element-wise kernels)."""
    import numpy as np
    e = np.arange(rows * cols, dtype=np.uint64)
    w = _philox(seed, e >> np.uint64(3), stream)   # eight elements per call, 16-bit lanes (philox.cuh)
    lane = (e & np.uint64(7)).astype(np.int64)
    word = np.take_along_axis(w, (lane >> 1)[:, None], axis=1)[:, 0]
    r = np.where(lane & 1, word >> np.uint64(16), word & np.uint64(0xffff))
    return torch.from_numpy((r >= np.uint64(min(int(p * 65536.0), 65535))).reshape(rows, cols))


def keep_attn(seed, stream, BH, Tq, Tk, p):
    """keep mask [BH, Tq, Tk] of the attention weights (philox.cuh attn_dropout_index / attn_dropout_half: 16 bits each)."""
    import numpy as np
    u = np.uint64
    nI, nJ = (Tq + 15) // 16, (Tk + 15) // 16
    bh, i, j = np.meshgrid(np.arange(BH, dtype=u), np.arange(Tq, dtype=u), np.arange(Tk, dtype=u), indexing="ij")
    idx = (((bh * u(nI) + (i >> u(4))) * u(nJ) + (j >> u(4))) << u(5)) + (i & u(7)) * u(4) + ((j & u(7)) >> u(1))
    half = (u(4) * ((i >> u(3)) & u(1)) + u(2) * ((j >> u(3)) & u(1)) + (j & u(1))).astype(np.int64).reshape(-1)
    w = _philox(seed, idx.reshape(-1), stream)
    word = np.take_along_axis(w, (half >> 1)[:, None], axis=1)[:, 0]
    r = np.where(half & 1, word >> u(16), word & u(0xffff))
    return torch.from_numpy((r >= u(min(int(p * 65536.0), 65535))).reshape(BH, Tq, Tk))


def test_dropout_mask_is_the_documented_philox_function(ops):
    from tts_b200 import train_ops as TO
    R, Cc, p = 37, 64, 0.3
    ones = torch.ones(R, Cc, device=DEV)
    got = TO.dropout_cast(ones, p, seed=0x1234567890abcdef, stream=11).float().cpu()
    keep = keep_linear(0x1234567890abcdef, 11, R, Cc, p)
    assert torch.equal(got != 0, keep)
    assert abs(float(got.max()) - 1 / (1 - p)) < 1e-2
    # the GEMM epilogue drops exactly the same elements for the same (seed, stream)
    a, b = _bf(torch.ones(R, 64)), _bf(torch.eye(Cc, 64))
    d = ops.gemm_bf16(a, b, out_dtype=torch.float32, drop_p=p, seed=0x1234567890abcdef, rng_stream=11).cpu()
    assert torch.equal(d != 0, keep)


def _attn_ref(q, k, v, B, H, Tq, Tk, dh, causal, key_len, keep, p):
    """fp32 torch reference on the bf16-rounded operands; q/k/v [B*T, H*dh] leaves requiring grad."""
    qh = q.view(B, Tq, H, dh).transpose(1, 2)
    kh = k.view(B, Tk, H, dh).transpose(1, 2)
    vh = v.view(B, Tk, H, dh).transpose(1, 2)
    logits = (qh * dh ** -0.5) @ kh.transpose(2, 3)
    if causal:
        logits = logits + torch.triu(torch.ones(Tq, Tk, dtype=logits.dtype), 1)[None, None] * -1e20
    if key_len is not None:
        logits = logits + ((torch.arange(Tk)[None, :] >= key_len[:, None]).to(logits.dtype) * -1e20)[:, None, None, :]
    w = torch.softmax(logits, -1)
    if keep is not None:
        w = w * keep.view(B, H, Tq, Tk).to(w.dtype) / (1 - p)
    return (w @ vh).transpose(1, 2).reshape(B * Tq, H * dh)


@pytest.mark.parametrize("dh,H,B,Tq,Tk,mode,p", [(96, 2, 2, 150, 150, "causal", 0.0), (64, 2, 3, 70, 70, "keylen", 0.0),
                                                 (96, 3, 2, 130, 50, "keylen", 0.0), (32, 2, 2, 33, 33, "causal", 0.0),
                                                 (96, 2, 2, 100, 100, "causal", 0.1), (64, 2, 2, 90, 41, "keylen", 0.25),
                                                 # several 128-key blocks / 64-query blocks of the tcgen05 backward, ragged edges
                                                 (96, 2, 2, 333, 333, "causal", 0.1), (96, 2, 2, 300, 258, "keylen", 0.1),
                                                 (96, 1, 1, 1000, 1000, "causal", 0.1), (96, 2, 1, 257, 129, "keylen", 0.0)])
def test_attention_train_fwd_bwd(ops, dh, H, B, Tq, Tk, mode, p):
    from tts_b200 import train_ops as TO
    g = torch.Generator().manual_seed(dh + Tq + Tk)
    D = H * dh
    # q/k/v live in one packed projection buffer like the model's (row stride 3D), rounded to bf16
    qkv_q = torch.randn(B * Tq, D, generator=g).to(torch.bfloat16)
    k_ = torch.randn(B * Tk, D, generator=g).to(torch.bfloat16)
    v_ = torch.randn(B * Tk, D, generator=g).to(torch.bfloat16)
    d_ctx = torch.randn(B * Tq, D, generator=g).to(torch.bfloat16)
    causal = mode == "causal"
    key_len = None if causal else torch.randint(1, Tk + 1, (B,), generator=g, dtype=torch.int32)
    if key_len is not None:
        key_len[0] = Tk
    seed, stream = 99, 5
    keep = keep_attn(seed, stream, B * H, Tq, Tk, p) if p > 0 else None
    q, k, v = (t.double().requires_grad_() for t in (qkv_q, k_, v_))
    want = _attn_ref(q, k, v, B, H, Tq, Tk, dh, causal, key_len, keep, p)
    want.backward(d_ctx.double())
    qd, kd, vd = qkv_q.to(DEV), k_.to(DEV), v_.to(DEV)
    kl = None if key_len is None else key_len.to(DEV)
    ctx, lse = TO.attn_fwd(qd, kd, vd, B, H, Tq, Tk, dh, causal, kl, p, seed, stream)
    assert _err(ctx, want) < 3e-2, "forward"
    dq, dk, dv = (torch.empty_like(t) for t in (qd, kd, vd))
    TO.attn_bwd(qd, kd, vd, ctx, lse, d_ctx.to(DEV), dq, dk, dv, B, H, Tq, Tk, dh, causal, kl, p, seed, stream)
    from tts_b200 import _native
    assert _native.load().tts_attn_tc_status() == 0, "a barrier wait of the tcgen05 backward kernel timed out"
    for name, got, ref in (("dq", dq, q.grad), ("dk", dk, k.grad), ("dv", dv, v.grad)):
        scale = float(ref.abs().max())
        assert _err(got, ref) < 3e-2 * max(scale, 1.0), (name, _err(got, ref), scale)
    if dh == 96:   # the tcgen05 kernel against the mma.sync two-kernel path on the same inputs
        dq2, dk2, dv2 = (torch.empty_like(t) for t in (qd, kd, vd))
        TO.attn_bwd(qd, kd, vd, ctx, lse, d_ctx.to(DEV), dq2, dk2, dv2, B, H, Tq, Tk, dh, causal, kl, p, seed, stream, deterministic=True)
        for name, got, ref in (("dq", dq, dq2), ("dk", dk, dk2), ("dv", dv, dv2)):
            assert _err(got, ref.float().cpu()) < 2e-2 * max(float(ref.float().abs().max()), 1.0), ("tc vs mma.sync", name)


def test_layernorm_train_fwd_bwd(ops):
    from tts_b200 import train_ops as TO
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    for (R, Cc) in ((37, 768), (300, 512), (9, 64), (2500, 128)):
        x = (torch.randn(R, Cc, generator=g) * 2 + 0.5)
        gam, bet = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
        dy = torch.randn(R, Cc, generator=g).to(torch.bfloat16)
        dres = torch.randn(R, Cc, generator=g)
        xr, gr, br = x.double().requires_grad_(), gam.double().requires_grad_(), bet.double().requires_grad_()
        y = F.layer_norm(xr, (Cc,), gr, br, 1e-6)
        y.backward(dy.double())
        yk, mean, rstd = TO.ln_fwd(x.to(DEV), gam.to(DEV), bet.to(DEV))
        assert _err(yk, y) < 4e-2
        dx, dg, db = TO.ln_bwd(dy.to(DEV), x.to(DEV), mean, rstd, gam.to(DEV), dres=dres.to(DEV))
        assert _err(dx, xr.grad + dres.double()) < 1e-4 * max(1.0, float(xr.grad.abs().max()))
        assert _err(dg, gr.grad) < 2e-3 * max(1.0, float(gr.grad.abs().max()) / 10)
        assert _err(db, br.grad) < 2e-3 * max(1.0, float(br.grad.abs().max()) / 10)
        # the fused bf16 output = the separate dropout-cast kernel on dx (same Philox mask), with and without dropout
        for pd in (0.0, 0.3):
            dx2, _, _, dyb = TO.ln_bwd(dy.to(DEV), x.to(DEV), mean, rstd, gam.to(DEV), dres=dres.to(DEV), cast_drop=(pd, 77, 9))
            assert torch.equal(dx2, dx)
            assert torch.equal(dyb, TO.dropout_cast(dx, pd, 77, 9))
    # row mask (final LayerNorm + impute)
    lens = torch.tensor([2, 5], dtype=torch.int32, device=DEV)
    x = torch.randn(10, 64, generator=g).to(DEV)
    yk, _, _ = TO.ln_fwd(x, torch.ones(64, device=DEV), torch.zeros(64, device=DEV), row_len=lens, rows_per_batch=5)
    assert float(yk[2:5].abs().max()) == 0 and float(yk[:2].abs().max()) > 0 and float(yk[5:].abs().min()) >= 0


def test_prologues_train_fwd_bwd(ops):
    from tts_b200 import train_ops as TO
    from tts_b200.engine import sinusoid_table
    g = torch.Generator().manual_seed(2)
    B, S, Cc, V = 3, 17, 64, 50
    ids = torch.randint(0, V, (B, S), generator=g)
    lens = torch.tensor([17, 5, 9], dtype=torch.int32)
    emb = torch.randn(V, Cc, generator=g)
    pe = sinusoid_table(32, Cc)
    sc = torch.tensor(0.7)
    dx = torch.randn(B * S, Cc, generator=g)
    p, seed, stream = 0.2, 5, 3
    keep = keep_linear(seed, stream, B * S, Cc, p).double() / (1 - p)
    er, sr = emb.double().requires_grad_(), sc.double().requires_grad_()
    mask = (torch.arange(S)[None, :] < lens[:, None]).double()[..., None]
    ref = ((er[ids] * mask + pe[:S].double() * sr).view(B * S, Cc)) * keep
    ref.backward(dx.double())
    out = TO.embed_fwd(ids.to(DEV), lens.to(DEV), emb.to(DEV), pe.to(DEV), sc.to(DEV), B, S, p, seed, stream)
    assert _err(out, ref) < 1e-5
    de, ds = TO.embed_bwd(dx.to(DEV), ids.to(DEV), lens.to(DEV), pe.to(DEV), V, B, S, p, seed, stream)
    assert _err(de, er.grad) < 1e-4 and _err(ds, sr.grad) < 1e-3
    # decoder prologue
    T = 11
    tl = torch.tensor([11, 4, 7], dtype=torch.int32)
    pre = torch.randn(B * T, Cc, generator=g)
    dx = torch.randn(B * T, Cc, generator=g)
    keep = keep_linear(seed, stream, B * T, Cc, p).double() / (1 - p)
    pr, sr = pre.double().requires_grad_(), sc.double().requires_grad_()
    m = (torch.arange(T)[None, :] < tl[:, None]).double()[..., None]
    xs = pr.view(B, T, Cc) * m
    xs = torch.cat([torch.zeros(B, 1, Cc, dtype=torch.double), xs[:, :-1]], 1) + pe[:T].double() * sr
    ref = xs.view(B * T, Cc) * keep
    ref.backward(dx.double())
    out = TO.shift_pe_fwd(pre.to(DEV), tl.to(DEV), pe.to(DEV), sc.to(DEV), B, T, p, seed, stream)
    assert _err(out, ref) < 1e-5
    dpre, ds = TO.shift_pe_bwd(dx.to(DEV), tl.to(DEV), pe.to(DEV), B, T, p, seed, stream)
    assert _err(dpre, pr.grad) < 3e-2 and _err(ds, sr.grad) < 1e-3


@pytest.mark.parametrize("last", [False, True])
def test_batchnorm_train_fwd_bwd(ops, last):
    from tts_b200 import train_ops as TO
    g = torch.Generator().manual_seed(3)
    B, T, Cc = 3, 29, 48
    z = torch.randn(B * T, Cc, generator=g) * 1.5 + 0.3
    gam, bet = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1
    lens = torch.tensor([29, 10, 1], dtype=torch.int32)
    dout = torch.randn(B * T, Cc, generator=g)
    res = torch.randn(B * T, Cc, generator=g)
    p, seed, stream = 0.5, 8, 2
    keep = keep_linear(seed, stream, B * T, Cc, p).double() / (1 - p)
    zr, gr, br = z.double().requires_grad_(), gam.double().requires_grad_(), bet.double().requires_grad_()
    mean, var = zr.mean(0), zr.var(0, unbiased=False)
    y = (zr - mean) / torch.sqrt(var + 1e-5) * gr + br
    mask = (torch.arange(T)[None, :] < lens[:, None]).double().reshape(B * T, 1)
    if last:
        ref = y * keep + res.double()
        ref.backward(dout.double())
    else:
        ref = torch.tanh(y) * keep * mask
        ref.backward(dout.double())
    rm, rv, nb = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    zd = z.to(DEV)
    if last:
        out = torch.empty(B * T, Cc, device=DEV)
        m_, is_ = TO.bn_fwd(zd, gam.to(DEV), bet.to(DEV), rm, rv, nb, False, p, seed, stream, None, B, T, out_f32=out,
                            residual=res.to(DEV))
        assert _err(out, ref) < 1e-4
    else:
        out = torch.full((B, T + 4, Cc), 7.0, device=DEV, dtype=torch.bfloat16)
        TO.pad_cast(None, None, B, T, out=out, only_pads=True)
        m_, is_ = TO.bn_fwd(zd, gam.to(DEV), bet.to(DEV), rm, rv, nb, True, p, seed, stream, lens.to(DEV), B, T, out_pad=out)
        assert _err(out[:, 2:T + 2].reshape(B * T, Cc), ref) < 2e-2
        assert float(out[:, :2].abs().max()) == 0 and float(out[:, T + 2:].abs().max()) == 0
    assert _err(rm, 0.1 * mean) < 1e-5 and _err(rv, 0.9 + 0.1 * zr.var(0, unbiased=True)) < 1e-4 and int(nb) == 1
    dz = torch.zeros(B, T + 4, Cc, device=DEV, dtype=torch.bfloat16)
    dg, db = TO.bn_bwd(zd, dout.to(DEV), gam.to(DEV), bet.to(DEV), m_, is_, not last, p, seed, stream, lens.to(DEV), not last, B, T, dz)
    assert _err(dz[:, 2:T + 2].reshape(B * T, Cc), zr.grad) < 2e-2 * max(1.0, float(zr.grad.abs().max()))
    assert _err(dg, gr.grad) < 1e-3 * max(1.0, float(gr.grad.abs().max()))
    assert _err(db, br.grad) < 1e-3 * max(1.0, float(br.grad.abs().max()))


def test_loss_and_reductions(ops):
    from tts_b200 import train_ops as TO
    from oracle import tts_oracle as O
    g = torch.Generator().manual_seed(4)
    B, T, M = 4, 23, 80
    lens = torch.tensor([23, 7, 1, 15], dtype=torch.int32)
    tgt, bef, aft = (torch.randn(B, T, M, generator=g) for _ in range(3))
    stop = torch.randn(B, T, generator=g) * 2
    br, ar, sr = bef.double().requires_grad_(), aft.double().requires_grad_(), stop.double().requires_grad_()
    cfg = O.ModelConfig.tiny()
    want = O.compute_loss({}, cfg, tgt.double(), lens, {"mel_bef": br, "mel_aft": ar, "stop_logits": sr})
    (want["bef_loss"] + want["aft_loss"] + want["stop_loss"]).backward()
    total = lens.sum().to(torch.int32).to(DEV)
    sums, aft_b, d_bef, d_aft, d_stop = TO.loss_fwd(bef.to(DEV), aft.to(DEV), stop.to(DEV), tgt.to(DEV), lens.to(DEV), total)
    n = float(lens.sum())
    assert abs(float(sums[0]) / n - float(want["bef_loss"])) < 1e-5
    assert abs(float(sums[1]) / n - float(want["aft_loss"])) < 1e-5
    assert abs(float(sums[2]) / n - float(want["stop_loss"])) < 1e-5
    assert _err(aft_b.cpu() / lens, want["aft_losses"]) < 1e-5
    assert _err(d_bef, br.grad) < 1e-7 and _err(d_aft, ar.grad) < 1e-7 and _err(d_stop, sr.grad) < 1e-7
    # reductions
    x = torch.randn(1000, 257, generator=g).to(torch.bfloat16)
    w = torch.randn(1000, generator=g)
    assert _err(TO.colsum(x.to(DEV)), x.double().sum(0)) < 2e-3
    assert _err(TO.colsum(x.to(DEV), w.to(DEV)), (x.double() * w.double()[:, None]).sum(0)) < 2e-3
    assert abs(float(TO.sum_f32(w.to(DEV))) - float(w.double().sum())) < 1e-3
    xr = torch.randn(50, 768, generator=g).to(torch.bfloat16)
    wv, bias = torch.randn(768, generator=g) / 27, torch.tensor([0.3])
    rl = torch.tensor([10, 25], dtype=torch.int32)
    got = TO.rowdot(xr.to(DEV), wv.to(DEV), bias.to(DEV), rl.to(DEV), 25)
    ref = (xr.double() @ wv.double() + 0.3) * (torch.arange(25)[None, :] < rl[:, None]).double().reshape(-1)
    assert _err(got, ref) < 1e-4


def test_multi_tensor_cast_l2_adam(ops):
    from tts_b200 import train_ops as TO
    g = torch.Generator().manual_seed(5)
    shapes = [(300, 70), (5,), (8192,), (1, 9001), (64, 64, 5)]
    ps = [torch.randn(*s, generator=g).to(DEV) for s in shapes]
    dst = [torch.empty(p.shape, device=DEV, dtype=torch.bfloat16) for p in ps]
    tab = TO.MultiTable(torch.device(DEV))
    tab.build_cast(list(zip(ps, dst)))
    TO.multi_cast(tab)
    for p, d in zip(ps, dst):
        assert torch.equal(d, p.to(torch.bfloat16))
    decay = [True, False, True, False, True]
    l2 = TO.MultiTable(torch.device(DEV))
    l2.build_opt([(p, None, None, None, dc) for p, dc in zip(ps, decay)])
    out = torch.empty((), device=DEV)
    TO.sumsq_multi(l2, out)
    want = sum(float((p.double() ** 2).sum()) for p, dc in zip(ps, decay) if dc)
    assert abs(float(out) - want) / want < 1e-5
    # Adam: two steps against torch.optim.Adam with the L2 term as an explicit gradient
    ref = [p.clone().requires_grad_() for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-3, eps=5e-8)
    mine = [p.clone() for p in ps]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    reg = 0.01
    for step in (1, 2):
        grads = [torch.randn(p.shape, generator=g).to(DEV) for p in ps]
        for r, gr, dc in zip(ref, grads, decay):
            r.grad = gr + (reg * r.detach() if dc else 0)
        opt.step()
        t = TO.MultiTable(torch.device(DEV))
        t.build_opt(list(zip(mine, grads, ms, vs, decay)))
        TO.adam_multi(t, 1e-3, 0.9, 0.999, 5e-8, step, reg_weight=reg)
        torch.cuda.synchronize()
    for r, m in zip(ref, mine):
        assert _err(r, m) < 1e-6
