"""Synthetic workloads of the mel path: the model-shape config, deterministic random weights with the reference's
state-dict schema, and seeded byte-text / mel batches.  Used by bench.py, the diagnostics under
tests/ and - re-exported - by the CPU oracle, so that the product arm of the benchmark never imports `oracle/`.
No model arithmetic lives here."""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Tuple

import torch

Params = Dict[str, torch.Tensor]


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
