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


@dataclasses.dataclass
class ModelConfig:
    """Model-shaping hyper-parameters (hyperparams.py:4,19,24-35,52-61)."""
    num_mels: int = 80
    vocab_size: int = 6000
    embed_size: int = 512
    encoder_hidden: int = 512
    decoder_hidden: int = 768
    n_encoder_layer: int = 6
    n_decoder_layer: int = 6
    n_attention_head: int = 8
    prenet_hidden: int = 256
    postnet_hidden: int = 512
    n_postnet_layer: int = 5
    multi_speaker: bool = True
    max_num_speaker: int = 1000
    speaker_embedding_size: int = 128
    multi_lingual: bool = True
    max_num_language: int = 100
    language_embedding_size: int = 128
    max_generation_frames: int = 1100
    reg_weight: float = 5e-9

    @property
    def memory_width(self) -> int:
        """Width of the encoder memory seen by the decoder (tacotron.py:96-100)."""
        w = self.encoder_hidden
        if self.multi_speaker:
            w += self.speaker_embedding_size
        if self.multi_lingual:
            w += self.language_embedding_size
        return w

    @classmethod
    def from_hparams(cls, hp) -> "ModelConfig":
        names = [f.name for f in dataclasses.fields(cls)]
        return cls(**{n: getattr(hp, n) for n in names})

    @classmethod
    def tiny(cls) -> "ModelConfig":
        """A small model used by fast CPU tests (same structure, small widths)."""
        return cls(vocab_size=300, embed_size=64, encoder_hidden=64, decoder_hidden=128,
                   n_encoder_layer=2, n_decoder_layer=2, n_attention_head=2,
                   prenet_hidden=32, postnet_hidden=48, n_postnet_layer=3,
                   max_num_speaker=20, speaker_embedding_size=32,
                   max_num_language=12, language_embedding_size=32,
                   max_generation_frames=64)


# --------------------------------------------------------------------------- #
# parameter schema + deterministic synthetic weights
# --------------------------------------------------------------------------- #
def param_shapes(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every *parameter* in reference state_dict order
    (transformer/tacotron.py:8-124, modules.py:23-106).
    BatchNorm buffers are listed by ``buffer_shapes``."""
    E, D, M = cfg.encoder_hidden, cfg.decoder_hidden, cfg.num_mels
    out: List[Tuple[str, Tuple[int, ...]]] = []
    out.append(("encoder.embed.weight", (cfg.vocab_size, cfg.embed_size)))
    if cfg.multi_speaker:
        s = cfg.speaker_embedding_size
        out += [("encoder.speaker_embed.weight", (cfg.max_num_speaker, s)),
                ("encoder.speaker_layer.weight", (s, s)), ("encoder.speaker_layer.bias", (s,))]
    if cfg.multi_lingual:
        g = cfg.language_embedding_size
        out += [("encoder.language_embed.weight", (g, cfg.max_num_language)),
                ("encoder.language_layer.weight", (g, g)), ("encoder.language_layer.bias", (g,))]
    p = "encoder.encoder."
    out.append((p + "pe_scale", ()))
    for group in ("self_attentions", "attn_layer_norms", "ffn_layers", "ffn_layer_norms"):
        for i in range(cfg.n_encoder_layer):
            c = cfg.embed_size if i == 0 else E
            if group == "self_attentions":
                out += [(f"{p}{group}.{i}.qkv_transform.weight", (3 * c, c)),
                        (f"{p}{group}.{i}.output_transform.weight", (c, c))]
            elif group == "attn_layer_norms":
                out += [(f"{p}{group}.{i}.weight", (c,)), (f"{p}{group}.{i}.bias", (c,))]
            elif group == "ffn_layers":
                out += [(f"{p}{group}.{i}.input_layer.weight", (4 * E, E)),
                        (f"{p}{group}.{i}.output_layer.weight", (E, 4 * E))]
            else:
                out += [(f"{p}{group}.{i}.weight", (E,)), (f"{p}{group}.{i}.bias", (E,))]
    out += [(p + "output_layer_norm.weight", (E,)), (p + "output_layer_norm.bias", (E,))]

    H = cfg.prenet_hidden
    out += [("decoder.prenet.dense0.weight", (H, M)), ("decoder.prenet.dense0.bias", (H,)),
            ("decoder.prenet.dense1.weight", (H, H)), ("decoder.prenet.dense1.bias", (H,)),
            ("decoder.prenet.dense_final.weight", (D, H))]
    p = "decoder.decoder."
    out.append((p + "pe_scale", ()))
    W = cfg.memory_width
    for group in ("self_attentions", "attn_layer_norms", "encdec_attentions", "encdec_layer_norms",
                  "ffn_layers", "ffn_layer_norms"):
        for i in range(cfg.n_decoder_layer):
            c = W if i == 0 else D
            if group == "self_attentions":
                out += [(f"{p}{group}.{i}.qkv_transform.weight", (3 * c, c)),
                        (f"{p}{group}.{i}.output_transform.weight", (c, c))]
            elif group in ("attn_layer_norms", "encdec_layer_norms"):
                out += [(f"{p}{group}.{i}.weight", (c,)), (f"{p}{group}.{i}.bias", (c,))]
            elif group == "encdec_attentions":
                out += [(f"{p}{group}.{i}.q_transform.weight", (D, D)),
                        (f"{p}{group}.{i}.kv_transform.weight", (2 * D, D)),
                        (f"{p}{group}.{i}.output_transform.weight", (D, D))]
            elif group == "ffn_layers":
                out += [(f"{p}{group}.{i}.input_layer.weight", (4 * D, D)),
                        (f"{p}{group}.{i}.output_layer.weight", (D, 4 * D))]
            else:
                out += [(f"{p}{group}.{i}.weight", (D,)), (f"{p}{group}.{i}.bias", (D,))]
    out += [(p + "output_layer_norm.weight", (D,)), (p + "output_layer_norm.bias", (D,))]
    out += [("decoder.mel_net.weight", (M, D)),
            ("decoder.stop_net.weight", (1, D)), ("decoder.stop_net.bias", (1,))]
    for i in range(cfg.n_postnet_layer):
        cin = M if i == 0 else cfg.postnet_hidden
        cout = M if i == cfg.n_postnet_layer - 1 else cfg.postnet_hidden
        out.append((f"postnet.conv_layers.{i}.weight", (cout, cin, 5)))
    for i in range(cfg.n_postnet_layer):
        cout = M if i == cfg.n_postnet_layer - 1 else cfg.postnet_hidden
        out += [(f"postnet.batchnorm_layers.{i}.weight", (cout,)),
                (f"postnet.batchnorm_layers.{i}.bias", (cout,))]
    return out


def buffer_shapes(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    out = []
    for i in range(cfg.n_postnet_layer):
        cout = cfg.num_mels if i == cfg.n_postnet_layer - 1 else cfg.postnet_hidden
        out += [(f"postnet.batchnorm_layers.{i}.running_mean", (cout,)),
                (f"postnet.batchnorm_layers.{i}.running_var", (cout,)),
                (f"postnet.batchnorm_layers.{i}.num_batches_tracked", ())]
    return out


def _truncated_normal(shape, std: float, gen: torch.Generator) -> torch.Tensor:
    """Draw 8 candidates per element, keep the first inside (-2 std, 2 std)
    (transformer/common.py:90-105)."""
    cand = torch.empty(tuple(shape) + (8,), dtype=torch.float32).normal_(0.0, std, generator=gen)
    ok = (cand < 2 * std) & (cand > -2 * std)
    first = ok.to(torch.uint8).max(-1, keepdim=True)[1]
    return cand.gather(-1, first).squeeze(-1)


def synth_params(cfg: ModelConfig, seed: int = 0, randomize_norm: bool = True) -> Params:
    """Deterministic synthetic weights following the *distributions* of
    ``initialize_variables`` (transformer/tacotron.py:161-173; fan-average variance
    scaling, transformer/common.py:108-124) drawn from a private generator, so the same
    state dict can be rebuilt from the seed on any box with the same torch build.

    ``randomize_norm`` additionally perturbs LayerNorm/BatchNorm affine parameters,
    BatchNorm running statistics, the pe_scales and the biases (all identity / zero at
    reference init) so that parity tests exercise them.
    """
    gen = torch.Generator().manual_seed(seed)
    out: Params = {}
    for name, shape in param_shapes(cfg):
        is_norm = "layer_norm" in name or "batchnorm" in name
        if name == "encoder.embed.weight":
            t = torch.empty(shape).normal_(0.0, 1.0, generator=gen)
        elif name in ("encoder.speaker_embed.weight", "encoder.language_embed.weight"):
            t = _truncated_normal(shape, 0.5, gen)
        elif "weight" in name and not is_norm:
            fan_in, fan_out = shape[1], shape[0]
            for d in shape[2:]:
                fan_in *= d
                fan_out *= d
            t = _truncated_normal(shape, math.sqrt(2.6 / ((fan_in + fan_out) / 2.0)), gen)
        elif name.endswith("pe_scale"):
            t = torch.tensor(1.0)
            if randomize_norm:
                t = t + 0.1 * torch.empty(()).normal_(generator=gen)
        elif is_norm and name.endswith("weight"):
            t = torch.ones(shape)
            if randomize_norm:
                t = t + 0.1 * torch.empty(shape).normal_(generator=gen)
        else:  # biases
            t = torch.zeros(shape)
            if randomize_norm:
                t = 0.05 * torch.empty(shape).normal_(generator=gen)
        out[name] = t.contiguous()
    for name, shape in buffer_shapes(cfg):
        if name.endswith("running_mean"):
            t = torch.zeros(shape)
            if randomize_norm:
                t = 0.1 * torch.empty(shape).normal_(generator=gen)
        elif name.endswith("running_var"):
            t = torch.ones(shape)
            if randomize_norm:
                t = t + 0.2 * torch.empty(shape).uniform_(-1.0, 1.0, generator=gen)
        else:
            t = torch.tensor(0, dtype=torch.long)
        out[name] = t
    return out


def params_checksum(params: Params) -> float:
    """Order-sensitive scalar fingerprint of a state dict (float64)."""
    acc = 0.0
    for i, (k, v) in enumerate(sorted(params.items())):
        if v.dtype.is_floating_point:
            v64 = v.double().flatten()
            w = torch.arange(1, v64.numel() + 1, dtype=torch.float64) % 97 + 1.0
            acc += float((v64 * w).sum()) * (1.0 + (i % 13))
    return acc


def synth_batch(cfg: ModelConfig, batch: int, text_len: int, n_frames: int, seed: int = 1,
                ragged: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded synthetic byte-text / mel pairs of a named shape (token ids
    follow utils/text.py:13-19 — 0 pad, 1 eos, 2 sos, bytes; batch-dict layout follows
    dataloader.py:419-439,498-508)."""
    g = torch.Generator().manual_seed(seed)
    S, T = text_len, n_frames
    if ragged and batch > 1:
        in_len = torch.randint(max(3, S // 3), S + 1, (batch,), generator=g)
        tg_len = torch.randint(max(2, T // 3), T + 1, (batch,), generator=g)
        in_len[0], tg_len[-1] = S, T
    else:
        in_len = torch.full((batch,), S, dtype=torch.long)
        tg_len = torch.full((batch,), T, dtype=torch.long)
    hi = min(256, cfg.vocab_size)
    ids = torch.randint(3, hi, (batch, S), generator=g)
    ids[:, 0] = 2
    pos = torch.arange(S)[None, :]
    ids = torch.where(pos == (in_len[:, None] - 1), torch.ones_like(ids), ids)
    ids = torch.where(pos < in_len[:, None], ids, torch.zeros_like(ids))
    mel = torch.empty(batch, T, cfg.num_mels).normal_(generator=g).clamp_(-4.0, 4.0)
    mel = mel * (torch.arange(T)[None, :, None] < tg_len[:, None, None])
    spk = torch.arange(batch) % min(572, cfg.max_num_speaker)
    lang = torch.zeros(batch, cfg.max_num_language)
    lang[torch.arange(batch), torch.arange(batch) % min(38, cfg.max_num_language)] = 1.0
    return {"inputs": ids, "input_lengths": in_len, "mel_targets": mel.contiguous(),
            "target_lengths": tg_len, "input_spk_ids": spk, "input_language_vecs": lang,
            "names": ["synth_%d" % i for i in range(batch)]}


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
    return torch.arange(n)[None, :] < lengths[:, None]


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
    x = x + sinusoid_table(S, E, dtype) * params["encoder.encoder.pe_scale"].to(dtype)
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
    x = x + sinusoid_table(T, x.shape[-1], dt) * params["decoder.decoder.pe_scale"].to(dt)
    enc_bias = ((~_length_mask(input_lengths, S)).to(dt) * NEG_BIAS)[:, None, None, :]
    causal = torch.triu(torch.ones(T, T, dtype=dt), diagonal=1) * NEG_BIAS     # common.py:41-43
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
    stop_target = (torch.arange(T)[None, :] == (target_lengths[:, None] - 1)).to(dt)
    ce = F.binary_cross_entropy_with_logits(outputs["stop_logits"], stop_target, reduction="none",
                                            pos_weight=torch.tensor([5.0], dtype=dt))
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
                      dtype=torch.float32, return_trace: bool = False):
    """Mathematically identical K/V-cached restatement of the same loop (SURVEY.md
    Appendix A): one decoder *row* per step; self K/V appended to a cache, cross K/V of
    the encoder memory computed once.  Used by tests to localise per-step kernel bugs and
    to reach long horizons on the CPU in reasonable time; validated against
    ``eval_batch_uncached`` in tests/test_oracle_golden.py."""
    max_frames = cfg.max_generation_frames if max_frames is None else max_frames
    B = batch["inputs"].shape[0]
    Hn, L, D, M = cfg.n_attention_head, cfg.n_decoder_layer, cfg.decoder_hidden, cfg.num_mels
    mem = encoder_forward(params, cfg, batch["inputs"], batch["input_lengths"],
                          batch.get("input_spk_ids"), batch.get("input_language_vecs"), dtype)
    S = mem.shape[1]
    p = "decoder.decoder."
    dh = D // Hn
    cross_k, cross_v = [], []
    for l in range(L):
        kv = mem @ params[f"{p}encdec_attentions.{l}.kv_transform.weight"].to(dtype).t()
        k, v = kv.split([D, D], dim=-1)
        cross_k.append(_heads(k, Hn))
        cross_v.append(_heads(v, Hn))
    key_bias = ((~_length_mask(batch["input_lengths"], S)).to(dtype) * NEG_BIAS)[:, None, None, :]
    pe = sinusoid_table(max_frames, D, dtype) * params[p + "pe_scale"].to(dtype)
    self_k = [torch.zeros(B, Hn, 0, dh, dtype=dtype) for _ in range(L)]
    self_v = [torch.zeros(B, Hn, 0, dh, dtype=dtype) for _ in range(L)]
    lengths = torch.ones(B, dtype=torch.int32)
    finished = torch.zeros(B, dtype=torch.bool)
    prev = torch.zeros(B, M, dtype=dtype)
    frames, logits, trace = [], [], []
    t = 0
    while not bool(finished.all()) and t < max_frames:
        if t == 0:
            x = torch.zeros(B, D, dtype=dtype)
        else:
            x = prenet(params, prev) * ((t - 1) < lengths)[:, None].to(dtype)
        x = x + pe[t]
        for l in range(L):
            h = _ln(x, params, f"{p}attn_layer_norms.{l}")
            qkv = h @ params[f"{p}self_attentions.{l}.qkv_transform.weight"].to(dtype).t()
            q, k, v = qkv.split([D, D, D], dim=-1)
            self_k[l] = torch.cat([self_k[l], k.view(B, Hn, 1, dh)], dim=2)
            self_v[l] = torch.cat([self_v[l], v.view(B, Hn, 1, dh)], dim=2)
            q = q.view(B, Hn, 1, dh) * dh ** -0.5
            w = torch.softmax(q @ self_k[l].transpose(2, 3), dim=-1)
            a = (w @ self_v[l]).reshape(B, D)
            x = x + a @ params[f"{p}self_attentions.{l}.output_transform.weight"].to(dtype).t()
            h = _ln(x, params, f"{p}encdec_layer_norms.{l}")
            q = (h @ params[f"{p}encdec_attentions.{l}.q_transform.weight"].to(dtype).t())
            q = q.view(B, Hn, 1, dh) * dh ** -0.5
            w = torch.softmax(q @ cross_k[l].transpose(2, 3) + key_bias, dim=-1)
            a = (w @ cross_v[l]).reshape(B, D)
            x = x + a @ params[f"{p}encdec_attentions.{l}.output_transform.weight"].to(dtype).t()
            x = x + ffn(params, f"{p}ffn_layers.{l}", _ln(x, params, f"{p}ffn_layer_norms.{l}"))
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
    mels = torch.stack(frames, dim=1) if frames else torch.zeros(B, 0, M, dtype=dtype)
    mel_aft = mels + postnet_forward(params, cfg, mels, lengths)
    out = {"mel_pre": mels, "mel_aft": mel_aft, "generated_lengths": lengths, "memory": mem,
           "stop_logits": torch.stack(logits, dim=1) if logits else torch.zeros(B, 0, dtype=dtype),
           "self_k": self_k, "self_v": self_v, "cross_k": cross_k, "cross_v": cross_v}
    if return_trace:
        out["trace"] = trace
    return out


def learning_rate_factor(step: int, warmup_steps=50000, decay_rate=1e-2, decay_step=550000,
                         min_lr=1e-5, max_lr=1e-3) -> float:
    """transformer/tacotron.py:176-179."""
    s = max(step - warmup_steps, 0)
    return max(min_lr / max_lr, decay_rate ** (s / decay_step))
