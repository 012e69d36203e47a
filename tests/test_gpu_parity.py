"""GPU parity tests: every CUDA entry point and the whole decode / teacher-forced paths against
the CPU oracle (oracle/tts_oracle.py) and the golden vectors produced by the reference itself.
All calls go through the C ABI (ctypes) — there is no other implementation to fall back to.

Tolerances (fp32 everywhere): per-kernel <= 1e-4 max-abs on O(1) values; whole-model mel frames
<= 1e-3 max-abs (the north-star bound); generated lengths / stop indices bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tts_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
DEV = "cuda:0"
KERNEL_TOL = 1e-4
MEL_TOL = 1e-3


def _dev(t):
    return t.to(DEV)


def _err(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from tts_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def tiny_engine(tiny_params, ops):
    from tts_b200.engine import TtsEngine
    cfg, params = tiny_params
    return TtsEngine.from_state_dict(params, cfg, DEV)


@pytest.fixture(scope="module")
def full_engine(full_params, ops):
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    return TtsEngine.from_state_dict(params, cfg, DEV)


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 1, 16), (7, 5, 80), (33, 81, 768), (130, 200, 256), (300, 1536, 512),
                                   (2100, 768, 3072), (64, 64, 4), (129, 257, 36)])
def test_gemm_nt_plain(ops, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    got = ops.linear(_dev(a), _dev(w))
    assert _err(got, a @ w.t()) < KERNEL_TOL


def test_gemm_nt_epilogues(ops):
    g = torch.Generator().manual_seed(3)
    M, N, K = 70, 96, 128
    a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    got = ops.linear(_dev(a), _dev(w), bias=_dev(bias), act=ops.ACT_RELU, residual=_dev(res), alpha=0.5)
    want = torch.relu(0.5 * (a @ w.t()) + bias) + res
    assert _err(got, want) < KERNEL_TOL
    lens = torch.tensor([3, 10, 0, 7, 10, 1, 9], dtype=torch.int32)   # 7 batches of 10 rows
    got = ops.linear(_dev(a), _dev(w), row_len=_dev(lens), rows_per_batch=10)
    mask = (torch.arange(10)[None, :] < lens[:, None]).reshape(-1, 1)
    assert _err(got, (a @ w.t()) * mask) < KERNEL_TOL


def test_gemm_cross_kv_head_split(ops):
    g = torch.Generator().manual_seed(4)
    B, S, D, H = 3, 11, 128, 2
    mem, w = torch.randn(B * S, D, generator=g), torch.randn(2 * D, D, generator=g) / D ** 0.5
    ok = torch.zeros(B, H, S, D // H, device=DEV)
    ov = torch.zeros_like(ok)
    ops.cross_kv(_dev(mem), _dev(w), B, S, H, ok, ov)
    kv = (mem @ w.t()).view(B, S, 2, H, D // H)
    assert _err(ok, kv[:, :, 0].permute(0, 2, 1, 3)) < KERNEL_TOL
    assert _err(ov, kv[:, :, 1].permute(0, 2, 1, 3)) < KERNEL_TOL


@pytest.mark.parametrize("B,T,cin,cout,last", [(2, 9, 80, 48, False), (3, 33, 48, 80, True), (1, 1, 16, 16, False)])
def test_conv5_as_gemm(ops, B, T, cin, cout, last):
    g = torch.Generator().manual_seed(B + T)
    x, w = torch.randn(B, T, cin, generator=g), torch.randn(cout, cin, 5, generator=g) / (5 * cin) ** 0.5
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    lens = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = T
    mask = (torch.arange(T)[None, :] < lens[:, None])[:, :, None].float()
    y = F.conv1d((x * mask).transpose(1, 2), w, padding=2).transpose(1, 2) * scale + shift
    xpad = ops.pad_rows(_dev(x), _dev(lens), B, T)
    assert _err(xpad[:, 2:T + 2], x * mask) == 0 and float(xpad[:, :2].abs().max()) == 0
    wp = _dev(w.permute(0, 2, 1).reshape(cout, -1).contiguous())
    if last:
        res = torch.randn(B, T, cout, generator=g)
        out = torch.empty(B, T, cout, device=DEV)
        ops.conv5(xpad, wp, _dev(scale), _dev(shift), ops.ACT_NONE, None, B, T, out, False, residual=_dev(res))
        assert _err(out, y + res) < KERNEL_TOL
    else:
        out = torch.zeros(B, T + 4, cout, device=DEV)
        ops.conv5(xpad, wp, _dev(scale), _dev(shift), ops.ACT_TANH, _dev(lens), B, T, out, True)
        assert _err(out[:, 2:T + 2], torch.tanh(y) * mask) < KERNEL_TOL
        assert float(out[:, :2].abs().max()) == 0 and float(out[:, T + 2:].abs().max()) == 0


def test_layernorm_and_prologues(ops):
    g = torch.Generator().manual_seed(5)
    rows, C = 37, 768
    x, gam, bet = torch.randn(rows, C, generator=g) * 3 + 1, torch.randn(C, generator=g), torch.randn(C, generator=g)
    got = ops.layernorm(_dev(x), _dev(gam), _dev(bet))
    assert _err(got, F.layer_norm(x, (C,), gam, bet, 1e-6)) < KERNEL_TOL
    # embed + mask + PE, and shift-right + mask + PE
    from tts_b200.engine import sinusoid_table
    B, S, V, E = 3, 13, 50, 64
    ids = torch.randint(0, V, (B, S), generator=g)
    lens = torch.tensor([13, 5, 1], dtype=torch.int32)
    table, scale = torch.randn(V, E, generator=g), torch.tensor(1.3)
    pe = sinusoid_table(32, E)
    assert _err(pe[:S], O.sinusoid_table(S, E)) == 0
    got = ops.embed_pe(_dev(ids), _dev(lens), _dev(table), _dev(pe), _dev(scale), B, S)
    want = table[ids] * (torch.arange(S)[None, :] < lens[:, None])[..., None] + pe[:S] * scale
    assert _err(got.view(B, S, E), want) < 1e-6
    pre = torch.randn(B * S, E, generator=g)
    got = ops.shift_pe(_dev(pre), _dev(lens), _dev(pe), _dev(scale), B, S)
    p3 = pre.view(B, S, E) * (torch.arange(S)[None, :] < lens[:, None])[..., None]
    want = torch.cat([torch.zeros(B, 1, E), p3[:, :-1]], 1) + pe[:S] * scale
    assert _err(got.view(B, S, E), want) < 1e-6


@pytest.mark.parametrize("dh,H,B,Tq,Tk,mode", [(32, 2, 2, 5, 5, "causal"), (64, 2, 3, 70, 70, "causal"),
                                               (96, 8, 2, 130, 41, "keys"), (64, 4, 2, 33, 33, "keys"),
                                               (96, 2, 1, 1, 200, "none"), (32, 1, 2, 65, 64, "keys")])
def test_attention_full_sequence(ops, dh, H, B, Tq, Tk, mode):
    g = torch.Generator().manual_seed(dh + Tq)
    C = H * dh
    q, k, v = (torch.randn(B, t, C, generator=g) for t in (Tq, Tk, Tk))
    klen = None
    bias = None
    if mode == "causal":
        bias = torch.triu(torch.ones(Tq, Tk), 1)[None, None] * O.NEG_BIAS
    elif mode == "keys":
        klen = torch.randint(1, Tk + 1, (B,), generator=g, dtype=torch.int32)
        klen[0] = Tk
        if B > 1:
            klen[1] = 0   # fully masked row: the reference's softmax over all -1e20 is uniform
        bias = ((torch.arange(Tk)[None, :] >= klen[:, None]).float() * O.NEG_BIAS)[:, None, None, :]
    qh, kh, vh = (O._heads(t, H) for t in (q, k, v))
    logits = (qh * dh ** -0.5) @ kh.transpose(2, 3)
    if bias is not None:
        logits = logits + bias
    wts = torch.softmax(logits, -1)
    want = (wts @ vh).transpose(1, 2).reshape(B, Tq, C)
    qd, kd, vd = _dev(q), _dev(k), _dev(v)
    ctx, align = ops.attention(qd.data_ptr(), C, kd.data_ptr(), C, vd.data_ptr(), C, B, H, Tq, Tk, dh,
                               mode == "causal", None if klen is None else _dev(klen), True, torch.device(DEV))
    assert _err(ctx.view(B, Tq, C), want) < KERNEL_TOL
    assert _err(align, wts) < 1e-5


# ---------------------------------------------------------------------------------------------
# model paths vs oracle and golden (tiny model: exercises every mask / ragged edge quickly)
# ---------------------------------------------------------------------------------------------
def test_tiny_encoder_and_teacher_forced_forward(tiny_engine, tiny_params, golden_dir):
    cfg, params = tiny_params
    z = np.load(os.path.join(golden_dir, "tiny_forward_loss_grad.npz"))
    batch = O.synth_batch(cfg, batch=3, text_len=20, n_frames=30, seed=6, ragged=True)
    want = O.tacotron_forward(params, cfg, batch)
    got = tiny_engine.forward(batch)
    assert _err(got["memory"], want["memory"]) < KERNEL_TOL
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert _err(got[k], want[k]) < KERNEL_TOL, k
        assert _err(got[k], torch.from_numpy(z[k])) < MEL_TOL, k          # the reference's own output
    for kind in ("self", "encdec"):
        for a, b in zip(got["alignments"][kind], want["alignments"][kind]):
            assert a.shape == b.shape and _err(a, b) < 1e-5


@pytest.mark.parametrize("impl", [1, 2, 4])
def test_tiny_autoregressive_vs_golden(tiny_engine, tiny_params, golden_dir, impl):
    cfg, params = tiny_params
    z = np.load(os.path.join(golden_dir, "tiny_ar.npz"))
    from tts_b200.engine import TtsEngine
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([float(z["stop_bias"])])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=5, text_len=24, n_frames=4, seed=7, ragged=True)
    T = int(z["max_frames"])
    got = eng.generate(batch, max_frames=T, record_align="all", chunk=7, impl=impl)
    want = O.eval_batch_cached(p, cfg, batch, T)
    assert got["generated_lengths"].cpu().tolist() == z["generated_lengths"].tolist()     # bit-exact stop indices
    assert got["generated_lengths"].dtype == torch.int32
    assert _err(got["stop_logits"], want["stop_logits"]) < 2e-4
    assert _err(got["mel_pre"], torch.from_numpy(z["mel_pre"])) < MEL_TOL
    assert _err(got["mel_aft"], torch.from_numpy(z["mel_aft"])) < MEL_TOL
    assert _err(got["mel_pre"], want["mel_pre"]) < 2e-4
    # cache contents: the K/V rows written step by step equal the oracle's
    sess = got["session"]
    t = want["self_k"][0].shape[2]
    for l in range(cfg.n_decoder_layer):
        assert _err(sess.self_k[l][:, :, :t], want["self_k"][l]) < 2e-4
        assert _err(sess.cross_v[l], want["cross_v"][l]) < KERNEL_TOL
    # attention rows recorded per step are the softmax rows of a teacher-forced pass over the result
    tf = O.decoder_forward(p, cfg, want["memory"], batch["input_lengths"],
                           torch.cat([want["mel_pre"], torch.zeros(5, 1, cfg.num_mels)], 1)[:, :T],
                           want["generated_lengths"], leave_one=True)
    # (only rows of still-running samples are comparable: the loop's lengths change over time)
    b = int(np.argmax(z["generated_lengths"]))
    a_got = got["alignments"]["encdec"][-1][b].cpu()
    assert _err(a_got, tf[2]["encdec"][-1][b]) < 1e-4


# ---------------------------------------------------------------------------------------------
# full-size model
# ---------------------------------------------------------------------------------------------
def test_cfg1_forward_vs_reference_golden(full_engine, full_params, golden_dir):
    """BASELINE config 0: B=1, 120-byte text -> 400 frames, teacher forced."""
    cfg, _ = full_params
    z = np.load(os.path.join(golden_dir, "cfg1_forward.npz"))
    batch = O.synth_batch(cfg, batch=1, text_len=122, n_frames=400, seed=1)
    got = full_engine.forward(batch)
    errs = {k: _err(got[k], torch.from_numpy(z[k])) for k in ("mel_bef", "mel_aft", "stop_logits")}
    print("cfg1 max-abs vs reference:", errs)
    assert max(errs.values()) < MEL_TOL
    a = got["alignments"]
    assert _err(a["self"][0][0, 0, ::8, ::8], torch.from_numpy(z["align_self_l0_h0"])) < 1e-5
    assert _err(a["encdec"][5][0, 7, :, ::8], torch.from_numpy(z["align_encdec_l5_h7"])) < 1e-5


def test_full_ragged_forward_vs_reference_golden(full_engine, full_params, golden_dir):
    cfg, _ = full_params
    z = np.load(os.path.join(golden_dir, "full_ragged_forward.npz"))
    batch = O.synth_batch(cfg, batch=3, text_len=40, n_frames=64, seed=2, ragged=True)
    got = full_engine.forward(batch, want_align=False)
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert _err(got[k], torch.from_numpy(z[k])) < MEL_TOL, k


@pytest.mark.parametrize("impl", [1, 2, 4])
def test_full_autoregressive_vs_reference_golden(full_params, golden_dir, ops, impl):
    """The reference's own eval_batch output (staggered stops) on the full-size model."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    z = np.load(os.path.join(golden_dir, "full_ar.npz"))
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([float(z["stop_bias"])])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=4, text_len=48, n_frames=4, seed=3, ragged=True)
    got = eng.generate(batch, max_frames=int(z["max_frames"]), record_align="all", chunk=16, impl=impl)
    assert got["generated_lengths"].cpu().tolist() == z["generated_lengths"].tolist()
    e1, e2 = _err(got["mel_pre"], torch.from_numpy(z["mel_pre"])), _err(got["mel_aft"], torch.from_numpy(z["mel_aft"]))
    print("full AR max-abs vs reference: mel_pre %.2e mel_aft %.2e (reference stop margin %.3f)" % (e1, e2, float(z["stop_margin"])))
    assert e1 < MEL_TOL and e2 < MEL_TOL
    T = got["mel_pre"].shape[1]
    assert _err(got["alignments"]["encdec"][5][:, :, :, T - 1], torch.from_numpy(z["align_encdec_l5"])) < 1e-4
    assert _err(got["alignments"]["self"][0][:, :, :, T - 1], torch.from_numpy(z["align_self_l0"])) < 1e-4


@pytest.mark.parametrize("B,split_note", [(1, "split-KV over 18 CTAs per head"), (3, "split-KV"), (32, "one CTA per head"),
                                          (40, "two row blocks, second one partial"), (64, "two full row blocks")])
@pytest.mark.parametrize("impl", [1, 4, 5])
def test_full_decode_steps_vs_oracle_batches(full_engine, full_params, B, split_note, impl):
    """Decode at several batch sizes (different split-KV factors), 24 steps, vs the cached oracle."""
    if impl == 5 and B <= 16:
        pytest.skip("impl 5 (two CTAs per SM, one row group each) serves batches of more than one row group")
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    from tts_b200.engine import TtsEngine
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=B, text_len=37, n_frames=4, seed=11, ragged=B > 1)
    want = O.eval_batch_cached(p, cfg, batch, 24)
    got = eng.generate(batch, max_frames=24, record_align="encdec", chunk=24, impl=impl)
    assert got["generated_lengths"].cpu().tolist() == want["generated_lengths"].tolist() == [25] * B
    assert _err(got["mel_pre"], want["mel_pre"]) < 2e-4
    assert _err(got["mel_aft"], want["mel_aft"]) < 2e-4


def test_decode_properties_at_baseline_shape(full_params, ops):
    """BASELINE config 1 shape (B=32, S=258): size-independent properties — determinism, batch
    permutation equivariance, impl agreement, zero frames after a stop, session reuse."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-5.0])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=32, text_len=258, n_frames=4, seed=21)
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    T = 48
    a = eng.generate(batch, max_frames=T, record_align="none", memory=mem)
    b = eng.generate(batch, max_frames=T, record_align="none", memory=mem, session=a["session"])   # reuse buffers
    assert torch.equal(a["mel_pre"], b["mel_pre"]) and torch.equal(a["generated_lengths"], b["generated_lengths"])
    for other in (1, 5):   # the per-phase kernels and the two-CTAs-per-SM experiment agree with the default (impl 4)
        c = eng.generate(batch, max_frames=T, record_align="none", memory=mem, impl=other)
        assert _err(a["mel_pre"], c["mel_pre"]) < 1e-4 and torch.equal(a["generated_lengths"], c["generated_lengths"])
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(0))
    pb = {k: (v[perm] if torch.is_tensor(v) else v) for k, v in batch.items()}
    d = eng.generate(pb, max_frames=T, record_align="none", memory=mem[perm.to(DEV)].contiguous())
    assert _err(d["mel_pre"], a["mel_pre"][perm.to(DEV)]) < 1e-4
    assert d["generated_lengths"].cpu().tolist() == a["generated_lengths"].cpu()[perm].tolist()
    lens = a["generated_lengths"].cpu()
    assert len(set(lens.tolist())) > 1, "expected staggered stops at this bias"
    for i in range(32):   # frames at or beyond the frozen length are exactly zero (modules.py:144)
        assert float(a["mel_pre"][i, int(lens[i]):].abs().max() if int(lens[i]) < a["mel_pre"].shape[1] else 0.0) == 0.0
    # and one oracle comparison at this shape, on the first 6 steps
    want = O.eval_batch_cached(p, cfg, batch, 6)
    assert _err(a["mel_pre"][:, :6], want["mel_pre"][:, :6]) < 2e-4


@pytest.mark.parametrize("rows", [8, 11])
def test_pipelined_group_size_override(full_params, ops, rows):
    """TTS_GROUP_ROWS splits the batch differently (4 groups of 8 with split K/V streams and combine phases; 3 ragged
    groups of 11/11/10): same frames as the oracle."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=32, text_len=70, n_frames=4, seed=8, ragged=True)
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    want = O.eval_batch_cached(p, cfg, batch, 20)
    ref = eng.generate(batch, max_frames=20, record_align="encdec", memory=mem, chunk=20, impl=1)
    assert _err(ref["mel_pre"], want["mel_pre"]) < 2e-4
    old = os.environ.get("TTS_GROUP_ROWS")
    os.environ["TTS_GROUP_ROWS"] = str(rows)
    try:
        got = eng.generate(batch, max_frames=20, record_align="encdec", memory=mem, chunk=20, impl=4)
        got5 = eng.generate(batch, max_frames=20, record_align="encdec", memory=mem, chunk=20, impl=5)
    finally:
        if old is None:
            del os.environ["TTS_GROUP_ROWS"]
        else:
            os.environ["TTS_GROUP_ROWS"] = old
    assert _err(got["mel_pre"], want["mel_pre"]) < 2e-4
    assert _err(got["mel_aft"], want["mel_aft"]) < 2e-4
    assert _err(got["alignments"]["encdec"][5], ref["alignments"]["encdec"][5]) < 1e-5
    assert _err(got5["mel_pre"], want["mel_pre"]) < 2e-4     # impl 5: groups of 8 / 11 rows spread over the two CTA slots
    assert _err(got5["alignments"]["encdec"][5], ref["alignments"]["encdec"][5]) < 1e-5


def test_tcgen05_gemm_vs_float64(ops):
    """csrc/gemm_tc.cu (tcgen05.mma kind::tf32, 3-term split, TMEM accumulator) against a float64 product, with every
    epilogue feature the dense path uses and ragged M / N tails; the FFMA2 kernel on the same inputs as a second check."""
    from tts_b200 import _native
    lib = _native.load()
    prev = lib.tts_gemm_use_tensor_cores(1)
    try:
        g = torch.Generator().manual_seed(4)
        for (M, N, K) in ((256, 128, 64), (1000, 200, 96), (4128, 512, 512), (2000, 1536, 768)):
            x = torch.randn(M, K, generator=g).to(DEV)
            w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
            b = torch.randn(N, generator=g).to(DEV)
            res = torch.randn(M, N, generator=g).to(DEV)
            lib.tts_gemm_use_tensor_cores(1)
            y = ops.linear(x, w, bias=b, act=ops.ACT_RELU, residual=res, alpha=0.5)
            lib.tts_gemm_use_tensor_cores(0)
            y0 = ops.linear(x, w, bias=b, act=ops.ACT_RELU, residual=res, alpha=0.5)
            want = torch.relu(0.5 * (x.double() @ w.double().t()) + b.double()) + res.double()
            assert _err(y.double(), want) < 2e-5, (M, N, K)
            assert _err(y, y0) < 2e-5, (M, N, K)
    finally:
        lib.tts_gemm_use_tensor_cores(prev)


def _resume_state(sess, lengths, finished, t, n_layers):
    """Oracle `resume` dict from a CUDA session's state at step t (K/V caches, last frame)."""
    return {"t": t, "self_k": [sess.self_k[l][:, :, :t].cpu() for l in range(n_layers)],
            "self_v": [sess.self_v[l][:, :, :t].cpu() for l in range(n_layers)],
            "prev": sess.frames[:, t - 1].cpu(), "lengths": lengths, "finished": finished}


def test_headline_config_full_horizon_vs_oracle(full_params, ops):
    """BASELINE configs[1] over its WHOLE horizon: B=32, S=258, 1000 frames, stop disabled, default kernel vs the
    cached oracle (fp32 CPU) at every frame.  North-star tolerance 1e-3 max-abs on mel frames across 1000 dependent
    steps (SURVEY.md §7 hard part 4); the error at t=100/500/999 is printed."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=32, text_len=258, n_frames=4, seed=1)
    T = 1000
    got = eng.generate(batch, max_frames=T, record_align="none", chunk=50)
    want = O.eval_batch_cached(p, cfg, batch, T)
    assert got["generated_lengths"].cpu().tolist() == want["generated_lengths"].tolist() == [T + 1] * 32
    per_t = (got["mel_pre"].cpu().double() - want["mel_pre"].double()).abs().amax(dim=(0, 2))
    aft = _err(got["mel_aft"], want["mel_aft"])
    print("headline horizon: max|mel_pre diff| t=100 %.2e t=500 %.2e t=999 %.2e overall %.2e; mel_aft %.2e"
          % (per_t[100], per_t[500], per_t[999], per_t.max(), aft))
    assert float(per_t.max()) < MEL_TOL and aft < MEL_TOL
    assert _err(got["stop_logits"], want["stop_logits"]) < MEL_TOL
    for l in (0, cfg.n_decoder_layer - 1):   # the K/V cache itself after 1000 appended rows
        assert _err(got["session"].self_k[l][:, :, :T], want["self_k"][l]) < MEL_TOL


def test_headline_config_staggered_stops_vs_oracle(full_params, ops):
    """Same shape with live stop decisions (bias -5): generated lengths bit-exact against the oracle over up to 400
    frames, exact zeros after each stop, and the oracle's stop margin (smallest |logit| on a live frame) reported."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-5.0])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=32, text_len=258, n_frames=4, seed=21)
    T = 400
    got = eng.generate(batch, max_frames=T, record_align="none", chunk=50)
    want = O.eval_batch_cached(p, cfg, batch, T)
    lg, ln = want["stop_logits"], want["generated_lengths"]
    live = torch.arange(lg.shape[1])[None, :] < ln[:, None]
    print("staggered stops: lengths", sorted(set(ln.tolist())), "oracle stop margin %.4f" % float(lg.abs()[live].min()))
    assert got["generated_lengths"].cpu().tolist() == ln.tolist()
    assert got["mel_pre"].shape == want["mel_pre"].shape
    assert _err(got["mel_pre"], want["mel_pre"]) < MEL_TOL and _err(got["mel_aft"], want["mel_aft"]) < MEL_TOL
    # early-exit compaction (generate's default without alignments): the live rows moved into smaller sessions
    assert len(got["compactions"]) >= 1 and got["compactions"][-1][1] <= 16, got["compactions"]
    flat = eng.generate(batch, max_frames=T, record_align="none", chunk=50, compact=False)
    assert flat["compactions"] == [] and flat["generated_lengths"].cpu().tolist() == ln.tolist()
    assert _err(flat["mel_pre"], got["mel_pre"]) < 1e-4 and _err(flat["stop_logits"], got["stop_logits"]) < 1e-3
    print("compaction points (step, live rows):", got["compactions"])


def test_generate_compaction_reused_session(full_params, ops):
    """Compaction with a caller-owned session that is reused across utterances: frames of rows that finished before a
    compacted stretch are zero even when the buffers held an earlier utterance; B=40 (three row groups -> two -> one)."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-5.0])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    T = 120
    sess = eng.new_session(40, 64, T, "none")
    first = O.synth_batch(cfg, batch=40, text_len=64, n_frames=4, seed=5)
    eng.generate(first, max_frames=T, record_align="none", chunk=20, session=sess, compact=False)   # dirty the buffers
    batch = O.synth_batch(cfg, batch=40, text_len=64, n_frames=4, seed=6)
    a = eng.generate(batch, max_frames=T, record_align="none", chunk=20, session=sess, compact=True)
    a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.items()}
    b = eng.generate(batch, max_frames=T, record_align="none", chunk=20, compact=False)
    assert a["generated_lengths"].cpu().tolist() == b["generated_lengths"].cpu().tolist()
    assert len(set(b["generated_lengths"].cpu().tolist())) > 2, "expected staggered stops"
    assert len(a["compactions"]) >= 1, a["compactions"]
    assert a["mel_pre"].shape == b["mel_pre"].shape
    assert _err(a["mel_pre"], b["mel_pre"]) < 1e-4 and _err(a["mel_aft"], b["mel_aft"]) < 1e-4
    lens = b["generated_lengths"].cpu()
    for i in range(40):
        if int(lens[i]) < a["mel_pre"].shape[1]:
            assert float(a["mel_pre"][i, int(lens[i]):].abs().max()) == 0.0


def test_long_reference_golden_640_frames(full_params, golden_dir, ops):
    """The REAL reference's eval_batch (uncached O(T^2) loop, tests/golden/make_golden.py long) over 640 frames at
    B=2: the K/V-cached CUDA path against the reference itself, not only against the oracle port."""
    from tts_b200.engine import TtsEngine
    path = os.path.join(golden_dir, "full_ar_long.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/full_ar_long.npz not generated")
    cfg, params = full_params
    z = np.load(path)
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=2, text_len=64, n_frames=4, seed=9, ragged=True)
    got = eng.generate(batch, max_frames=int(z["max_frames"]), record_align="none", chunk=50)
    assert got["generated_lengths"].cpu().tolist() == z["generated_lengths"].tolist()
    e1, e2 = _err(got["mel_pre"], torch.from_numpy(z["mel_pre"])), _err(got["mel_aft"], torch.from_numpy(z["mel_aft"]))
    print("640-frame reference golden: mel_pre %.2e mel_aft %.2e" % (e1, e2))
    assert e1 < MEL_TOL and e2 < MEL_TOL


def test_cfg5_long_form_2000_frames_vs_oracle(full_params, ops):
    """BASELINE configs[4]: T=2000.  B=2 (split K/V streams) is compared with the oracle at every frame; B=128
    (eight row groups, 250-tile streams) is compared on the first 50 frames and, by resuming the oracle from the
    CUDA session's own state at t=1950, on the last 50 frames."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    cfg5 = O.ModelConfig(max_generation_frames=2000)
    eng = TtsEngine.from_state_dict(p, cfg5, DEV)
    T = 2000
    small = O.synth_batch(cfg5, batch=2, text_len=258, n_frames=4, seed=31)
    got = eng.generate(small, max_frames=T, record_align="none", chunk=50)
    want = O.eval_batch_cached(p, cfg5, small, T)
    per_t = (got["mel_pre"].cpu().double() - want["mel_pre"].double()).abs().amax(dim=(0, 2))
    print("cfg5 B=2: max|mel_pre diff| t=500 %.2e t=1000 %.2e t=1999 %.2e" % (per_t[500], per_t[1000], per_t[1999]))
    assert float(per_t.max()) < MEL_TOL and _err(got["mel_aft"], want["mel_aft"]) < MEL_TOL
    del got, want
    wide = O.synth_batch(cfg5, batch=128, text_len=258, n_frames=4, seed=32, ragged=True)
    sess = eng.new_session(128, 258, T, "none")
    mem = eng.encode(wide["inputs"], wide["input_lengths"], wide["input_spk_ids"], wide["input_language_vecs"])
    sess.begin(mem, wide["input_lengths"].to(DEV))
    sess.step(50)
    head = O.eval_batch_cached(p, cfg5, wide, 50, memory=mem.cpu())
    assert _err(sess.frames[:, :50], head["mel_pre"]) < 2e-4
    for _ in range(38):
        sess.step(50)            # t = 1950
    torch.cuda.synchronize()
    state = _resume_state(sess, sess.lengths.cpu(), sess.finished.cpu().bool(), 1950, cfg5.n_decoder_layer)
    sess.step(50)                # t = 2000
    tail = O.eval_batch_cached(p, cfg5, wide, T, memory=mem.cpu(), resume=state)
    e = _err(sess.frames[:, 1950:2000], tail["mel_pre"])
    print("cfg5 B=128: last 50 frames (oracle resumed from the CUDA state at t=1950) max-abs %.2e" % e)
    assert e < 2e-4
    assert sess.lengths.cpu().tolist() == tail["generated_lengths"].tolist() == [T + 1] * 128
    with pytest.raises(RuntimeError):   # stepping past t_max is refused, on the host ...
        sess.step(1)


def test_decode_contract_frames_are_inputs_and_bounds(full_params, ops):
    """ABI contract (include/tts_b200.h): step t reads frame t-1 from st->frames, whoever wrote it.  A frame perturbed
    between single-step launches must change the next frame exactly as it does for the per-phase kernels; and the
    kernels refuse to step past t_max even when the host-side guard is bypassed (n_unfinished = -2, nothing written)."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=5, text_len=33, n_frames=4, seed=12, ragged=True)
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    outs = []
    for impl in (4, 1):
        sess = eng.new_session(5, 33, 8, "none")
        sess.begin(mem, batch["input_lengths"].to(DEV))
        sess.step(3, impl=impl)
        sess.frames[:, 2] += 0.5 * torch.arange(5, device=DEV)[:, None]      # the caller edits frame t-1 = 2
        sess.step(1, impl=impl)
        sess.step(2, impl=4 if impl == 1 else 1)                             # and switches implementation mid-session
        outs.append(sess.frames[:, :6].clone())
    assert _err(outs[0], outs[1]) < 2e-4
    plain = eng.generate(batch, max_frames=6, record_align="none", memory=mem)["mel_pre"]
    assert _err(outs[0][:, 3], plain[:, 3]) > 1e-2                            # the edit did change frame 3
    # device-side bound: bypass the Python guard
    for impl in (4, 1):
        sess = eng.new_session(5, 33, 4, "none")
        sess.begin(mem, batch["input_lengths"].to(DEV))
        sess.step(4, impl=impl)
        guard = torch.full((64,), 7.0, device=DEV)
        sess.t = 0                                                           # pretend the host lost count
        before = sess.frames.clone()
        sess.step(2, impl=impl)
        torch.cuda.synchronize()
        assert int(sess.counters[1].item()) == -2 and int(sess.counters[0].item()) == 4
        assert torch.equal(before, sess.frames) and float(guard.min()) == 7.0
    # the same two contracts on the two-CTAs-per-SM kernel (impl 5, three row groups over two slots)
    wide = O.synth_batch(cfg, batch=40, text_len=33, n_frames=4, seed=13, ragged=True)
    wmem = eng.encode(wide["inputs"], wide["input_lengths"], wide["input_spk_ids"], wide["input_language_vecs"])
    outs = []
    for impl in (5, 1):
        sess = eng.new_session(40, 33, 8, "none")
        sess.begin(wmem, wide["input_lengths"].to(DEV))
        sess.step(3, impl=impl)
        sess.frames[:, 2] += 0.01 * torch.arange(40, device=DEV)[:, None]
        sess.step(1, impl=impl)
        sess.step(2, impl=1 if impl == 5 else 5)
        outs.append(sess.frames[:, :6].clone())
    assert _err(outs[0], outs[1]) < 2e-4
    sess = eng.new_session(40, 33, 4, "none")
    sess.begin(wmem, wide["input_lengths"].to(DEV))
    sess.step(4, impl=5)
    sess.t = 0
    before = sess.frames.clone()
    sess.step(2, impl=5)
    torch.cuda.synchronize()
    assert int(sess.counters[1].item()) <= -(1 << 29) and int(sess.counters[0].item()) == 4
    assert torch.equal(before, sess.frames)


# ---------------------------------------------------------------------------------------------
# the drop-in nn.Module API, driven the way the reference's synthesize.py / train.py drive it
# ---------------------------------------------------------------------------------------------
def _eval_loop_like_synthesize(model, data, max_frames, num_mels):
    """The loop of the reference's synthesize.eval_batch (synthesize.py:17-72), restated here because the
    reference checkout does not exist on the GPU box: same calls, same bookkeeping, same outputs."""
    with torch.no_grad():
        device = data["inputs"].device
        n = data["inputs"].shape[0]
        target_lengths = torch.ones([n], dtype=torch.int32, device=device)
        finished = torch.zeros([n], dtype=torch.bool, device=device)
        mels = torch.zeros([n, 0, num_mels], dtype=torch.float32, device=device)
        enc = model.encoder(data["inputs"], data["input_lengths"], data["input_spk_ids"], data["input_language_vecs"])
        n_calls = 0
        while not torch.all(finished) and mels.shape[1] < max_frames:
            dec_in = torch.cat([mels, torch.zeros([n, 1, num_mels], device=device)], dim=1)
            mel_bef, stop_logits, align = model.decoder(enc, data["input_lengths"], dec_in, target_lengths,
                                                        leave_one=True)
            stop = stop_logits[:, -1] > 0
            mels = torch.cat([mels, mel_bef[:, -1:]], dim=1)
            finished = torch.logical_or(finished, stop)
            target_lengths = torch.where(finished, target_lengths, target_lengths + 1)
            n_calls += 1
        mel_aft = mels + model.postnet(mels, target_lengths)
        return {"mel_pre": mels, "mel_aft": mel_aft, "alignments": align, "generated_lengths": target_lengths,
                "n_calls": n_calls}


def test_module_api_unchanged_synthesis_loop(tiny_params, golden_dir, ops):
    from tts_b200.config import hparams_from
    from transformer import tacotron
    cfg, params = tiny_params
    z = np.load(os.path.join(golden_dir, "tiny_ar.npz"))
    hp = hparams_from(cfg)
    hp.max_generation_frames = int(z["max_frames"])
    m = tacotron.Tacotron(hp)
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([float(z["stop_bias"])])
    m.load_state_dict(p, strict=True)
    m.to(DEV).eval()
    m.decoder.record_alignments = "all"   # the reference's full structure (default: 'encdec' only, 'self' list empty)
    batch = {k: (_dev(v) if torch.is_tensor(v) else v)
             for k, v in O.synth_batch(cfg, batch=5, text_len=24, n_frames=4, seed=7, ragged=True).items()}
    out = _eval_loop_like_synthesize(m, batch, hp.max_generation_frames, cfg.num_mels)
    assert out["generated_lengths"].cpu().tolist() == z["generated_lengths"].tolist()
    assert _err(out["mel_pre"], torch.from_numpy(z["mel_pre"])) < MEL_TOL
    assert _err(out["mel_aft"], torch.from_numpy(z["mel_aft"])) < MEL_TOL
    T = out["mel_pre"].shape[1]
    assert out["alignments"]["self"][0].shape == (5, cfg.n_attention_head, T, T)
    assert out["alignments"]["encdec"][0].shape == (5, cfg.n_attention_head, 24, T)
    # the incremental path really was used: one cached step per call
    assert m.decoder._inc is not None and m.decoder._inc["sess"].t == out["n_calls"]
    # a second utterance batch through the same model restarts the cache and reproduces the result
    out2 = _eval_loop_like_synthesize(m, batch, hp.max_generation_frames, cfg.num_mels)
    assert torch.equal(out2["mel_pre"], out["mel_pre"])
    # teacher-forced call through the same module (train.py:171 signature), eval mode
    tb = {k: (_dev(v) if torch.is_tensor(v) else v)
          for k, v in O.synth_batch(cfg, batch=3, text_len=20, n_frames=30, seed=6, ragged=True).items()}
    with torch.no_grad():
        o = m(**tb)
    zf = np.load(os.path.join(golden_dir, "tiny_forward_loss_grad.npz"))
    p0 = dict(params)
    want = O.tacotron_forward(p, cfg, {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in tb.items()})
    for k in ("mel_bef", "mel_aft", "stop_logits"):
        assert _err(o[k], want[k]) < KERNEL_TOL, k
    losses = tacotron.compute_loss(m, tb["mel_targets"], tb["target_lengths"], o, hp)
    assert set(losses) == {"loss", "bef_loss", "aft_loss", "aft_losses", "mse_loss", "l2", "stop_loss"}
    del p0, zf


def test_module_api_submodules(tiny_params, ops):
    """MultiheadAttention / FFNLayer / DecoderPrenet / Postnet / TransformerEncoder called directly."""
    from tts_b200.config import hparams_from
    from transformer import attention, common, modules, tacotron
    cfg, params = tiny_params
    hp = hparams_from(cfg)
    g = torch.Generator().manual_seed(9)
    mha = attention.MultiheadAttention(128, 128, True, 2).to(DEV).eval()
    x = torch.randn(2, 9, 128, generator=g)
    sd = {k: v.cpu() for k, v in mha.state_dict().items()}
    bias = common.attention_bias(9, "causal")
    with torch.no_grad():
        out = mha(_dev(x), None, _dev(bias))
    want, al = O.attention({"a." + k: v for k, v in sd.items()}, "a", x, None, bias, 2)
    assert _err(out["outputs"], want) < KERNEL_TOL and _err(out["align"], al) < 1e-5
    cross = attention.MultiheadAttention(128, 128, False, 2).to(DEV).eval()
    mem = torch.randn(2, 6, 128, generator=g)
    mask = torch.arange(6)[None, :] < torch.tensor([6, 2])[:, None]
    bias = common.attention_bias(mask, "masking")
    with torch.no_grad():
        out = cross(_dev(x), _dev(mem), _dev(bias))
    sd = {"a." + k: v.cpu() for k, v in cross.state_dict().items()}
    want, al = O.attention(sd, "a", x, mem, bias, 2)
    assert _err(out["outputs"], want) < KERNEL_TOL and _err(out["align"], al) < 1e-5
    ffn = modules.FFNLayer(128, 512, 128).to(DEV).eval()
    with torch.no_grad():
        y = ffn(_dev(x))
    sd = {"f." + k: v.cpu() for k, v in ffn.state_dict().items()}
    assert _err(y, O.ffn(sd, "f", x)) < KERNEL_TOL
    post = tacotron.Postnet(hp)
    post.load_state_dict({k[len("postnet."):]: v for k, v in params.items() if k.startswith("postnet.")})
    post.to(DEV).eval()
    mel = torch.randn(3, 17, 80, generator=g)
    lens = torch.tensor([17, 4, 9])
    with torch.no_grad():
        r = post(_dev(mel), _dev(lens))
    assert _err(r, O.postnet_forward(params, cfg, mel, lens)) < KERNEL_TOL
    enc = tacotron.Encoder(hp)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in params.items() if k.startswith("encoder.")})
    enc.to(DEV).eval()
    b = O.synth_batch(cfg, batch=3, text_len=15, n_frames=4, seed=2, ragged=True)
    with torch.no_grad():
        memo = enc(_dev(b["inputs"]), _dev(b["input_lengths"]), _dev(b["input_spk_ids"]), _dev(b["input_language_vecs"]))
        emb = enc.encoder(enc.embed.weight[_dev(b["inputs"])], _dev(b["input_lengths"]))
    want = O.encoder_forward(params, cfg, b["inputs"], b["input_lengths"], b["input_spk_ids"], b["input_language_vecs"])
    assert _err(memo, want) < KERNEL_TOL
    assert _err(emb, want[:, :, :cfg.encoder_hidden]) < KERNEL_TOL


# ---------------------------------------------------------------------------------------------
# decoder.train() at synthesis time (eval.py:116-117): Philox dropout inside the decode kernel
# ---------------------------------------------------------------------------------------------
def _philox_np(seed, idx, stream):
    idx = np.asarray(idx, dtype=np.uint64)
    M32 = np.uint64(0xffffffff)
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64(seed >> 32)
    c0, c1 = idx & M32, idx >> np.uint64(32)
    c2, c3 = np.full_like(idx, stream), np.full_like(idx, 0x5eed)
    for _ in range(7):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & M32, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & M32
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M32, (k1 + np.uint64(0xBB67AE85)) & M32
    return np.stack([c0, c1, c2, c3], axis=-1)


def _keep(seed, site, e, p):
    e = np.asarray(e, dtype=np.uint64)
    w = _philox_np(seed, (e >> np.uint64(2)).reshape(-1), site)
    r = np.take_along_axis(w, (e.reshape(-1) & np.uint64(3)).astype(np.int64)[:, None], axis=1)[:, 0]
    return torch.from_numpy((r >= np.uint64(min(int(p * 4294967296.0), 0xffffffff))).reshape(e.shape))


def test_decode_dropout_matches_oracle_with_the_same_masks(tiny_params, ops):
    """`m.eval(); m.decoder.train()` semantics (SURVEY §8b "Modes"): the decode kernel's Philox dropout against the oracle
    applying the SAME masks (csrc/philox.cuh restated in numpy) at every site the reference drops - placement, scaling and
    the pre-dropout softmax normaliser are all pinned, not just a keep rate.  Also: keep rate, seed dependence, eval mode."""
    from tts_b200.engine import TtsEngine
    cfg, params = tiny_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    batch = O.synth_batch(cfg, batch=5, text_len=24, n_frames=4, seed=7, ragged=True)
    B, H, T = 5, cfg.n_attention_head, 12
    pd, pt, seed = 0.5, 0.1, 0x1234abcd5678
    names = {"self_w": 0, "self_out": 1, "cross_w": 2, "cross_out": 3, "ffn_hid": 4, "ffn_out": 5}
    rates = []

    def drop(site, layer, t, x):
        if site in ("pre0", "pre1", "dec_in"):
            sid, rate = {"pre0": 1, "pre1": 2, "dec_in": 3}[site], (pt if site == "dec_in" else pd)
        else:
            sid, rate = 16 + 8 * layer + names[site], pt
        if site.endswith("_w"):      # [B,H,1,Tk]: element = ((t * B * H + bh) << 16) + key
            Tk = x.shape[-1]
            bh = np.arange(B * H, dtype=np.uint64)[:, None]
            e = ((np.uint64(t) * np.uint64(B * H) + bh) << np.uint64(16)) + np.arange(Tk, dtype=np.uint64)[None, :]
            keep = _keep(seed, sid, e, rate).view(B, H, 1, Tk)
        else:                        # [B,N]: element = (t * B + b) * N + n
            Nn = x.shape[-1]
            e = (np.uint64(t) * np.uint64(B) + np.arange(B, dtype=np.uint64)[:, None]) * np.uint64(Nn) + np.arange(Nn, dtype=np.uint64)[None, :]
            keep = _keep(seed, sid, e, rate)
        rates.append((rate, float(keep.float().mean()), keep.numel()))
        return x * keep.to(x.dtype) / (1 - rate)

    want = O.eval_batch_cached(p, cfg, batch, T, dropout=drop)
    got = eng.generate(batch, max_frames=T, record_align="encdec", chunk=5, dropout=(pd, pt, seed))
    e = _err(got["mel_pre"], want["mel_pre"])
    print("decode with dropout vs oracle with the same masks: max-abs %.2e" % e)
    assert e < 2e-4 and _err(got["mel_aft"], want["mel_aft"]) < 2e-4
    big = [(r, k) for r, k, n in rates if n >= 600]
    assert big and all(abs(k - (1 - r)) < 0.08 for r, k in big)
    plain = eng.generate(batch, max_frames=T, record_align="encdec", chunk=5)
    other = eng.generate(batch, max_frames=T, record_align="encdec", chunk=5, dropout=(pd, pt, seed + 1))
    assert _err(plain["mel_pre"], got["mel_pre"]) > 1e-2 and _err(other["mel_pre"], got["mel_pre"]) > 1e-2
    assert _err(plain["mel_pre"], O.eval_batch_cached(p, cfg, batch, T)["mel_pre"]) < 2e-4
    with pytest.raises(RuntimeError):     # only the pipelined kernel implements it: no silent deterministic fallback
        eng.generate(batch, max_frames=T, record_align="none", chunk=5, impl=1, dropout=(pd, pt, seed))
    # through the module API: m.eval(); m.decoder.train() gives live dropout in the incremental decode path
    from tts_b200.config import hparams_from
    from transformer import tacotron
    hp = hparams_from(cfg)
    hp.max_generation_frames = T
    m = tacotron.Tacotron(hp)
    m.load_state_dict(p, strict=True)
    m.to(DEV).eval()
    dbatch = {k: (_dev(v) if torch.is_tensor(v) else v) for k, v in batch.items()}
    det = _eval_loop_like_synthesize(m, dbatch, T, cfg.num_mels)["mel_pre"]
    m.decoder.train()
    a = _eval_loop_like_synthesize(m, dbatch, T, cfg.num_mels)["mel_pre"]
    b = _eval_loop_like_synthesize(m, dbatch, T, cfg.num_mels)["mel_pre"]
    assert _err(a, det) > 1e-2 and _err(a, b) > 1e-2 and bool(torch.isfinite(a).all())


def test_cfg4_multilingual_mixed_batch_share(full_params, ops):
    """BASELINE configs[3]: 38-language / 128-speaker mixed ragged batch, global B=128 = 16 per GPU.  One GPU's share
    (16 rows, text lengths spanning 32..258 in ONE batch, distinct languages and speakers per row): encoder memory,
    conditioning columns, 48 decoded frames and the teacher-forced forward against the oracle at the ragged edges."""
    from tts_b200.engine import TtsEngine
    cfg, params = full_params
    p = dict(params)
    p["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(p, cfg, DEV)
    g = torch.Generator().manual_seed(44)
    B, S = 16, 258
    lens = torch.randint(32, 259, (B,), generator=g)
    lens[0], lens[1] = 32, 258                                   # both edges in the same batch
    batch = O.synth_batch(cfg, batch=B, text_len=S, n_frames=4, seed=45)
    batch["input_lengths"] = lens
    for b in range(B):
        batch["inputs"][b, int(lens[b]) - 1] = 1                 # eos at the ragged end, pad after it
        batch["inputs"][b, int(lens[b]):] = 0
    batch["input_spk_ids"] = (torch.arange(B) * 37 + 5) % 572   # distinct speakers
    lang = torch.zeros(B, cfg.max_num_language)
    lang[torch.arange(B), (torch.arange(B) * 7) % 38] = 1.0      # distinct languages out of 38
    batch["input_language_vecs"] = lang
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    want_mem = O.encoder_forward(p, cfg, batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    live = (torch.arange(S)[None, :] < lens[:, None])[:, :, None]
    assert _err(mem.cpu() * live, want_mem * live) < 2e-4        # valid positions (padded query rows are computed too,
    assert _err(mem, want_mem) < 2e-3                            # the reference masks only the keys: modules.py:50-52)
    assert _err(mem[:, :, cfg.encoder_hidden:], want_mem[:, :, cfg.encoder_hidden:]) < 1e-5   # speaker | language columns
    want = O.eval_batch_cached(p, cfg, batch, 48)
    got = eng.generate(batch, max_frames=48, record_align="encdec", chunk=48)
    assert _err(got["mel_pre"], want["mel_pre"]) < 2e-4 and _err(got["mel_aft"], want["mel_aft"]) < 2e-4
    a = got["alignments"]["encdec"][5].cpu()                     # no attention mass beyond each row's text length
    for b in (0, 1, 7):
        if int(lens[b]) < S:
            assert float(a[b, :, int(lens[b]):, :].abs().max()) == 0.0
        assert abs(float(a[b, 0, :, 3].sum()) - 1.0) < 1e-4
