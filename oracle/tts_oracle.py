"""CPU oracle for the Transformer-TTS mel path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, state-dict driven restatement (plain torch ops on the
CPU, fp32 by default, fp64 on request) of the reference algorithm of
mutiann/few-shot-transformer-tts for the hot path named in BASELINE.json.  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``few-shot-transformer-tts_b200/``) never does.

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
§4), so this restatement is pinned against outputs of the reference itself,
generated in the build container by ``tests/golden/make_golden.py`` (which
imports ``/root/reference``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks it on every run.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  All functions run in inference semantics (dropout = 0),
which is the only mode in which the reference is deterministic (SURVEY.md,
fact 5); ``postnet`` additionally supports batch-statistics BatchNorm for the
teacher-forced training parity tests.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

NEG_BIAS = -1e20  # transformer/common.py:32  (attention_bias "inf")
LN_EPS = 1e-6     # transformer/modules.py:36,88
BN_EPS = 1e-5     # torch.nn.BatchNorm1d default, transformer/tacotron.py:79


# The model-shape config and the seeded synthetic weight / batch generators live in the package (they are workload
# definitions, not model arithmetic) so that bench.py's product arm does not import this file; re-exported here.
try:
    from tts_b200.synthetic import (ModelConfig, param_shapes, buffer_shapes, _truncated_normal, synth_params,  # noqa: F401
                                    params_checksum, synth_batch)
except ImportError:  # oracle imported without the package directory on sys.path
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                      "few-shot-transformer-tts_b200"))
    from tts_b200.synthetic import (ModelConfig, param_shapes, buffer_shapes, _truncated_normal, synth_params,  # noqa: F401
                                    params_checksum, synth_batch)


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def sinusoid_table(length: int, channels: int, dtype=torch.float32) -> torch.Tensor:
    """[sin | cos] halves, float64 math then cast (transformer/common.py:4-29)."""
    half = channels // 2
    step = np.log(1e4 / 1.0) / (half - 1)
    inv = np.exp(np.arange(half) * -step)
    ang = np.arange(length)[:, None] * inv[None, :]
    tab = np.concatenate([np.sin(ang), np.cos(ang)], axis=1)
    if channels % 2:
        tab = np.pad(tab, [[0, 0], [0, 1]])
    return torch.from_numpy(tab).to(dtype)


def _length_mask(lengths: torch.Tensor, n: int) -> torch.Tensor:
    """[B, n] bool, True where position < length (transformer/common.py:51-70)."""
    return torch.arange(n, device=lengths.device)[None, :] < lengths[:, None]


def _ln(x, params: Params, prefix: str):
    return F.layer_norm(x, (x.shape[-1],), params[prefix + ".weight"].to(x.dtype),
                        params[prefix + ".bias"].to(x.dtype), LN_EPS)


def _heads(x, n_heads):  # [B,T,C] -> [B,H,T,C/H]      (transformer/attention.py:6-15)
    B, T, C = x.shape
    return x.view(B, T, n_heads, C // n_heads).transpose(1, 2)


def attention(params: Params, prefix: str, queries, memories, bias, n_heads: int):
    """Bias-free multi-head scaled dot-product attention (transformer/attention.py:53-122).
    Returns (outputs [B,Tq,C], align [B,H,Tk,Tq])."""
    dt = queries.dtype
    C = queries.shape[-1]
    if memories is None:
        qkv = queries @ params[prefix + ".qkv_transform.weight"].to(dt).t()
        q, k, v = qkv.split([C, C, C], dim=-1)                      # attention.py:63-64
    else:
        q = queries @ params[prefix + ".q_transform.weight"].to(dt).t()
        kv = memories @ params[prefix + ".kv_transform.weight"].to(dt).t()
        k, v = kv.split([kv.shape[-1] // 2] * 2, dim=-1)            # attention.py:66-68
    q, k, v = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    q = q * (C // n_heads) ** -0.5                                  # attention.py:113-114
    logits = q @ k.transpose(2, 3)
    if bias is not None:
        logits = logits + bias.to(dt)                               # attention.py:84-85
    w = torch.softmax(logits, dim=-1)
    ctx = w @ v
    B, H, Tq, dh = ctx.shape
    ctx = ctx.transpose(1, 2).reshape(B, Tq, H * dh)                # attention.py:18-26
    out = ctx @ params[prefix + ".output_transform.weight"].to(dt).t()
    return out, w.transpose(2, 3)                                   # attention.py:88


def ffn(params: Params, prefix: str, x):
    """Linear -> ReLU -> Linear, no biases (transformer/modules.py:8-20)."""
    h = torch.relu(x @ params[prefix + ".input_layer.weight"].to(x.dtype).t())
    return h @ params[prefix + ".output_layer.weight"].to(x.dtype).t()


def prenet(params: Params, x):
    """80 -> 256 -> 256 -> D (transformer/tacotron.py:55-65), dropout off."""
    dt = x.dtype
    p = "decoder.prenet."
    x = torch.relu(x @ params[p + "dense0.weight"].to(dt).t() + params[p + "dense0.bias"].to(dt))
    x = torch.relu(x @ params[p + "dense1.weight"].to(dt).t() + params[p + "dense1.bias"].to(dt))
    return x @ params[p + "dense_final.weight"].to(dt).t()


# --------------------------------------------------------------------------- #
# encoder / decoder / postnet / full model
# --------------------------------------------------------------------------- #
def encoder_forward(params: Params, cfg: ModelConfig, inputs, input_lengths,
                    spk_ids=None, lang_vecs=None, dtype=torch.float32):
    """transformer/tacotron.py:33-44 + transformer/modules.py:49-69."""
    x = params["encoder.embed.weight"].to(dtype)[inputs]                       # tacotron.py:34
    B, S, E = x.shape
    mask = _length_mask(input_lengths, S)
    x = x * mask[..., None]                                                    # modules.py:50-51
    bias = ((~mask).to(dtype) * NEG_BIAS)[:, None, None, :]                    # common.py:44-46
    x = x + sinusoid_table(S, E, dtype).to(x.device) * params["encoder.encoder.pe_scale"].to(dtype)
    p = "encoder.encoder."
    for i in range(cfg.n_encoder_layer):
        y, _ = attention(params, f"{p}self_attentions.{i}", _ln(x, params, f"{p}attn_layer_norms.{i}"),
                         None, bias, cfg.n_attention_head)
        x = x + y
        x = x + ffn(params, f"{p}ffn_layers.{i}", _ln(x, params, f"{p}ffn_layer_norms.{i}"))
    out = _ln(x, params, p + "output_layer_norm")
    if cfg.multi_speaker:                                                      # tacotron.py:27-31,36-39
        e = params["encoder.speaker_embed.weight"].to(dtype)[spk_ids]
        e = e @ params["encoder.speaker_layer.weight"].to(dtype).t() + params["encoder.speaker_layer.bias"].to(dtype)
        e = F.softsign(e)
        out = torch.cat([out, e[:, None, :].expand(B, S, -1)], dim=-1)
    if cfg.multi_lingual:                                                      # tacotron.py:21-25,40-43
        e = lang_vecs.to(dtype) @ params["encoder.language_embed.weight"].to(dtype).t()
        e = e @ params["encoder.language_layer.weight"].to(dtype).t() + params["encoder.language_layer.bias"].to(dtype)
        e = F.softsign(e)
        out = torch.cat([out, e[:, None, :].expand(B, S, -1)], dim=-1)
    return out


def decoder_forward(params: Params, cfg: ModelConfig, memory, input_lengths, targets, target_lengths,
                    leave_one: bool = False):
    """Teacher-forced (full-sequence) decoder: transformer/tacotron.py:107-116 +
    transformer/modules.py:108-145.  Returns (mels [B,T,M], stop_logits [B,T], align dict)."""
    dt = memory.dtype
    B, T, _ = targets.shape
    S = memory.shape[1]
    x = prenet(params, targets.to(dt))
    if leave_one:                                                              # tacotron.py:109-110
        x = torch.cat([x[:, :-1], torch.zeros_like(x[:, -1:])], dim=1)
    tmask = _length_mask(target_lengths, T)
    x = x * tmask[..., None]                                                   # modules.py:114
    x = torch.cat([torch.zeros_like(x[:, :1]), x[:, :-1]], dim=1)              # modules.py:115-116
    x = x + sinusoid_table(T, x.shape[-1], dt).to(x.device) * params["decoder.decoder.pe_scale"].to(dt)
    enc_bias = ((~_length_mask(input_lengths, S)).to(dt) * NEG_BIAS)[:, None, None, :]
    causal = torch.triu(torch.ones(T, T, dtype=dt, device=x.device), diagonal=1) * NEG_BIAS     # common.py:41-43
    causal = causal[None, None]
    p = "decoder.decoder."
    self_align, cross_align = [], []
    for i in range(cfg.n_decoder_layer):
        y, a = attention(params, f"{p}self_attentions.{i}", _ln(x, params, f"{p}attn_layer_norms.{i}"),
                         None, causal, cfg.n_attention_head)
        self_align.append(a)
        x = x + y
        y, a = attention(params, f"{p}encdec_attentions.{i}", _ln(x, params, f"{p}encdec_layer_norms.{i}"),
                         memory, enc_bias, cfg.n_attention_head)
        cross_align.append(a)
        x = x + y
        x = x + ffn(params, f"{p}ffn_layers.{i}", _ln(x, params, f"{p}ffn_layer_norms.{i}"))
    x = _ln(x, params, p + "output_layer_norm") * tmask[..., None]             # modules.py:142-144
    mels = (x @ params["decoder.mel_net.weight"].to(dt).t()) * tmask[..., None]
    # the stop head reads *detached* features (tacotron.py:114)
    stop = (x.detach() @ params["decoder.stop_net.weight"].to(dt).t() + params["decoder.stop_net.bias"].to(dt))
    stop = stop.squeeze(-1) * tmask
    return mels, stop, {"self": self_align, "encdec": cross_align}


def postnet_forward(params: Params, cfg: ModelConfig, mels, lengths, batch_stats: bool = False):
    """5 x (mask -> Conv1d k5 pad2 -> BatchNorm1d -> tanh except last), returns the
    residual in [B,T,M] (transformer/tacotron.py:81-90).  ``batch_stats`` selects the
    train-mode statistics (over all B x T positions, padding included)."""
    dt = mels.dtype
    x = mels.transpose(1, 2)
    T = x.shape[-1]
    m = _length_mask(lengths, T)[:, None, :].to(dt)
    n = cfg.n_postnet_layer
    for i in range(n):
        x = F.conv1d(x * m, params[f"postnet.conv_layers.{i}.weight"].to(dt), padding=2)
        b = f"postnet.batchnorm_layers.{i}."
        if batch_stats:
            mean = x.mean(dim=(0, 2))
            var = x.var(dim=(0, 2), unbiased=False)
        else:
            mean, var = params[b + "running_mean"].to(dt), params[b + "running_var"].to(dt)
        x = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
        x = x * params[b + "weight"].to(dt)[None, :, None] + params[b + "bias"].to(dt)[None, :, None]
        if i != n - 1:
            x = torch.tanh(x)
    return x.transpose(1, 2)


def tacotron_forward(params: Params, cfg: ModelConfig, batch, dtype=torch.float32,
                     batch_stats: bool = False):
    """transformer/tacotron.py:126-133."""
    mem = encoder_forward(params, cfg, batch["inputs"], batch["input_lengths"],
                          batch.get("input_spk_ids"), batch.get("input_language_vecs"), dtype)
    mel_bef, stop, align = decoder_forward(params, cfg, mem, batch["input_lengths"],
                                           batch["mel_targets"].to(dtype), batch["target_lengths"])
    mel_aft = mel_bef + postnet_forward(params, cfg, mel_bef, batch["target_lengths"], batch_stats)
    return {"mel_bef": mel_bef, "mel_aft": mel_aft, "stop_logits": stop, "alignments": align,
            "memory": mem}


def l2_names(params: Params) -> List[str]:
    """The weight tensors that enter the L2 term, selected by name exactly as
    transformer/tacotron.py:144-146 does."""
    return [n for n in params
            if "weight" in n and "layer_norm" not in n and "batchnorm" not in n
            and "encoder.speaker_embed" not in n and "encoder.embed" not in n
            and params[n].dtype.is_floating_point]


def compute_loss(params: Params, cfg: ModelConfig, mel_targets, target_lengths, outputs):
    """transformer/tacotron.py:136-158 (masked MSE before/after, BCE(pos_weight 5) on the
    one-hot stop target at length-1, L2 on the name-selected weights)."""
    dt = outputs["mel_bef"].dtype
    T = mel_targets.shape[1]
    m = _length_mask(target_lengths, T).to(dt)
    n_valid = target_lengths.sum().to(dt)

    def masked_mean(per_frame):                                                # common.py:73-88
        return (per_frame * m).sum() / n_valid

    bef = ((outputs["mel_bef"] - mel_targets) ** 2).mean(-1)
    aft = ((outputs["mel_aft"] - mel_targets) ** 2).mean(-1)
    aft_each = (aft * m).sum(-1) / target_lengths.to(dt)
    l2 = cfg.reg_weight * sum((params[n].to(dt) ** 2).sum() / 2 for n in l2_names(params))
    dev = outputs["mel_bef"].device
    stop_target = (torch.arange(T, device=dev)[None, :] == (target_lengths.to(dev)[:, None] - 1)).to(dt)
    ce = F.binary_cross_entropy_with_logits(outputs["stop_logits"], stop_target, reduction="none",
                                            pos_weight=torch.tensor([5.0], dtype=dt, device=dev))
    bef_l, aft_l, ce_l = masked_mean(bef), masked_mean(aft), masked_mean(ce)
    return {"loss": bef_l + aft_l + l2 + ce_l, "bef_loss": bef_l, "aft_loss": aft_l,
            "aft_losses": aft_each, "mse_loss": (bef_l + aft_l) / 2, "l2": l2, "stop_loss": ce_l}


# --------------------------------------------------------------------------- #
# autoregressive synthesis
# --------------------------------------------------------------------------- #
def eval_batch_uncached(params: Params, cfg: ModelConfig, batch, max_frames: Optional[int] = None,
                        dtype=torch.float32, return_align: bool = False):
    """The reference's AR loop, restated faithfully *including its O(T^2) recompute*:
    every step re-runs the whole decoder over all frames so far and keeps the last
    frame (synthesize.py:17-72).  This is the CPU baseline algorithm."""
    max_frames = cfg.max_generation_frames if max_frames is None else max_frames
    B = batch["inputs"].shape[0]
    lengths = torch.ones(B, dtype=torch.int32)                                # synthesize.py:23
    finished = torch.zeros(B, dtype=torch.bool)
    mels = torch.zeros(B, 0, cfg.num_mels, dtype=dtype)
    mem = encoder_forward(params, cfg, batch["inputs"], batch["input_lengths"],
                          batch.get("input_spk_ids"), batch.get("input_language_vecs"), dtype)
    align = None
    while not bool(finished.all()) and mels.shape[1] < max_frames:           # synthesize.py:35
        dec_in = torch.cat([mels, torch.zeros(B, 1, cfg.num_mels, dtype=dtype)], dim=1)
        mel_bef, stop, align = decoder_forward(params, cfg, mem, batch["input_lengths"], dec_in,
                                               lengths, leave_one=True)
        fire = stop[:, -1] > 0                                                # synthesize.py:42
        mels = torch.cat([mels, mel_bef[:, -1:]], dim=1)
        finished = finished | fire
        lengths = torch.where(finished, lengths, lengths + 1)                 # synthesize.py:44-45
    mel_aft = mels + postnet_forward(params, cfg, mels, lengths)              # synthesize.py:56
    out = {"mel_pre": mels, "mel_aft": mel_aft, "generated_lengths": lengths, "memory": mem}
    if return_align:
        out["alignments"] = align
    return out


def eval_batch_cached(params: Params, cfg: ModelConfig, batch, max_frames: Optional[int] = None,
                      dtype=torch.float32, return_trace: bool = False, resume: Optional[dict] = None,
                      memory: Optional[torch.Tensor] = None, dropout=None):
    """Mathematically identical K/V-cached restatement of the same loop (SURVEY.md
    Appendix A): one decoder *row* per step; self K/V appended to a (preallocated) cache, cross
    K/V of the encoder memory computed once.  Used by tests to localise per-step kernel bugs and
    to reach long horizons on the CPU in reasonable time; validated against
    ``eval_batch_uncached`` in tests/test_oracle_golden.py and against the reference's own
    640-frame eval_batch output (tests/golden/full_ar_long.npz).

    ``resume`` = {"t": t0, "self_k": [L x [B,H,t0,dh]], "self_v": ..., "prev": [B,M],
    "lengths": [B] int32, "finished": [B] bool} continues a decode from step t0 with the given
    state (tests use it to check late steps of very long decodes without replaying all of them);
    ``max_frames`` is then the absolute step to stop at.

    ``dropout(site, layer, t, x)`` (optional) applies a caller-defined dropout mask at every place the reference's
    ``decoder.train()`` mode does (eval.py:116-117): sites "pre0" / "pre1" (tacotron.py:58,62), "dec_in"
    (modules.py:120), "self_w" / "cross_w" (attention.py:89, after the softmax), "self_out" / "cross_out" / "ffn_out"
    (modules.py:132,138,141) and "ffn_hid" (modules.py:18).  Tests pass the masks of csrc/philox.cuh."""
    max_frames = cfg.max_generation_frames if max_frames is None else max_frames
    drop = dropout if dropout is not None else (lambda site, layer, t, x: x)
    B = batch["inputs"].shape[0]
    Hn, L, D, M = cfg.n_attention_head, cfg.n_decoder_layer, cfg.decoder_hidden, cfg.num_mels
    mem = memory if memory is not None else encoder_forward(
        params, cfg, batch["inputs"], batch["input_lengths"], batch.get("input_spk_ids"),
        batch.get("input_language_vecs"), dtype)
    S = mem.shape[1]
    p = "decoder.decoder."
    dh = D // Hn
    cross_k, cross_v = [], []
    for l in range(L):
        kv = mem @ params[f"{p}encdec_attentions.{l}.kv_transform.weight"].to(dtype).t()
        k, v = kv.split([D, D], dim=-1)
        cross_k.append(_heads(k, Hn).contiguous())
        cross_v.append(_heads(v, Hn).contiguous())
    key_bias = ((~_length_mask(batch["input_lengths"], S)).to(dtype) * NEG_BIAS)[:, None, None, :]
    dev = mem.device   # CPU in every test; bench.py's gpu_eager_baseline runs the same code on cuda (cuBLAS fp32)
    pe = sinusoid_table(max_frames, D, dtype).to(dev) * params[p + "pe_scale"].to(dtype)
    kbuf = [torch.zeros(B, Hn, max_frames, dh, dtype=dtype, device=dev) for _ in range(L)]
    vbuf = [torch.zeros(B, Hn, max_frames, dh, dtype=dtype, device=dev) for _ in range(L)]
    lengths = torch.ones(B, dtype=torch.int32, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    prev = torch.zeros(B, M, dtype=dtype, device=dev)
    t = 0
    if resume is not None:
        t = int(resume["t"])
        for l in range(L):
            kbuf[l][:, :, :t] = resume["self_k"][l].to(dtype)
            vbuf[l][:, :, :t] = resume["self_v"][l].to(dtype)
        lengths = resume["lengths"].to(torch.int32).clone()
        finished = resume["finished"].to(torch.bool).clone()
        prev = resume["prev"].to(dtype).clone()
    t_first = t
    frames, logits, trace = [], [], []
    while not bool(finished.all()) and t < max_frames:
        if t == 0:
            x = torch.zeros(B, D, dtype=dtype, device=dev)
        elif dropout is None:
            x = prenet(params, prev) * ((t - 1) < lengths)[:, None].to(dtype)
        else:                                                                  # tacotron.py:55-65 with its dropouts
            pp = "decoder.prenet."
            h0 = drop("pre0", 0, t, torch.relu(prev @ params[pp + "dense0.weight"].to(dtype).t() + params[pp + "dense0.bias"].to(dtype)))
            h1 = drop("pre1", 0, t, torch.relu(h0 @ params[pp + "dense1.weight"].to(dtype).t() + params[pp + "dense1.bias"].to(dtype)))
            x = (h1 @ params[pp + "dense_final.weight"].to(dtype).t()) * ((t - 1) < lengths)[:, None].to(dtype)
        x = drop("dec_in", 0, t, x + pe[t])                                    # modules.py:117-120
        for l in range(L):
            h = _ln(x, params, f"{p}attn_layer_norms.{l}")
            qkv = h @ params[f"{p}self_attentions.{l}.qkv_transform.weight"].to(dtype).t()
            q, k, v = qkv.split([D, D, D], dim=-1)
            kbuf[l][:, :, t] = k.view(B, Hn, dh)
            vbuf[l][:, :, t] = v.view(B, Hn, dh)
            q = q.view(B, Hn, 1, dh) * dh ** -0.5
            w = drop("self_w", l, t, torch.softmax(q @ kbuf[l][:, :, :t + 1].transpose(2, 3), dim=-1))
            a = (w @ vbuf[l][:, :, :t + 1]).reshape(B, D)
            x = x + drop("self_out", l, t, a @ params[f"{p}self_attentions.{l}.output_transform.weight"].to(dtype).t())
            h = _ln(x, params, f"{p}encdec_layer_norms.{l}")
            q = (h @ params[f"{p}encdec_attentions.{l}.q_transform.weight"].to(dtype).t())
            q = q.view(B, Hn, 1, dh) * dh ** -0.5
            w = drop("cross_w", l, t, torch.softmax(q @ cross_k[l].transpose(2, 3) + key_bias, dim=-1))
            a = (w @ cross_v[l]).reshape(B, D)
            x = x + drop("cross_out", l, t, a @ params[f"{p}encdec_attentions.{l}.output_transform.weight"].to(dtype).t())
            hid = torch.relu(_ln(x, params, f"{p}ffn_layer_norms.{l}") @ params[f"{p}ffn_layers.{l}.input_layer.weight"].to(dtype).t())
            hid = drop("ffn_hid", l, t, hid)                                    # modules.py:14-19
            x = x + drop("ffn_out", l, t, hid @ params[f"{p}ffn_layers.{l}.output_layer.weight"].to(dtype).t())
        live = (t < lengths)[:, None].to(dtype)
        o = _ln(x, params, p + "output_layer_norm") * live
        mel = (o @ params["decoder.mel_net.weight"].to(dtype).t()) * live
        logit = (o @ params["decoder.stop_net.weight"].to(dtype).t()
                 + params["decoder.stop_net.bias"].to(dtype)).squeeze(-1) * live.squeeze(-1)
        frames.append(mel)
        logits.append(logit)
        if return_trace:
            trace.append({"x": x.clone(), "o": o.clone()})
        prev = mel
        finished = finished | (logit > 0)
        lengths = torch.where(finished, lengths, lengths + 1)
        t += 1
    mels = torch.stack(frames, dim=1) if frames else torch.zeros(B, 0, M, dtype=dtype, device=dev)
    out = {"mel_pre": mels, "generated_lengths": lengths, "memory": mem, "first_step": t_first,
           "stop_logits": torch.stack(logits, dim=1) if logits else torch.zeros(B, 0, dtype=dtype),
           "self_k": [k[:, :, :t] for k in kbuf], "self_v": [v[:, :, :t] for v in vbuf],
           "cross_k": cross_k, "cross_v": cross_v}
    if resume is None:   # mel_aft needs every frame from 0 on (synthesize.py:56)
        out["mel_aft"] = mels + postnet_forward(params, cfg, mels, lengths)
    if return_trace:
        out["trace"] = trace
    return out


def learning_rate_factor(step: int, warmup_steps=50000, decay_rate=1e-2, decay_step=550000,
                         min_lr=1e-5, max_lr=1e-3) -> float:
    """transformer/tacotron.py:176-179."""
    s = max(step - warmup_steps, 0)
    return max(min_lr / max_lr, decay_rate ** (s / decay_step))
