"""FFNLayer / TransformerEncoder / TransformerDecoder with the reference's constructors,
parameter names and forward signatures (reference: transformer/modules.py).  The modules only
own parameters; the arithmetic runs in the sm_100a kernels through tts_b200.engine."""
import weakref

import torch
from torch import nn

from tts_b200 import _native as N
from tts_b200 import ops
from tts_b200.engine import TtsEngine
from transformer.attention import MultiheadAttention
from transformer.common import *  # noqa: F401,F403  (the reference re-exports these names too)

_ENGINES = weakref.WeakKeyDictionary()


def engine_for(module, prefix, hparams):
    """The TtsEngine serving `module`, over its live parameters/buffers renamed to the reference's
    full state-dict names (`prefix` + local name).  Rebuilt when the parameters move device."""
    tensors = dict(module.named_parameters())
    tensors.update(dict(module.named_buffers()))
    first = next(iter(tensors.values()))
    if not first.is_cuda:
        raise RuntimeError("tts_b200: the model is on %s; this implementation has no CPU path - move it to a CUDA "
                           "device (model.to('cuda'))" % first.device)
    hit = _ENGINES.get(module)
    if hit is not None and hit[1] == first.device and all(hit[0].w.get(prefix + k) is v for k, v in tensors.items()):
        return hit[0]
    eng = TtsEngine({prefix + k: v for k, v in tensors.items()}, hparams, first.device)
    _ENGINES[module] = (eng, first.device)
    return eng


_TRAIN_ENGINES = weakref.WeakKeyDictionary()


def train_engine_for(module, prefix, hparams):
    """The TrainEngine (bf16 tensor-core forward + backward) serving `module`; same keying as engine_for."""
    from tts_b200.engine_train import TrainEngine
    tensors = dict(module.named_parameters())
    tensors.update(dict(module.named_buffers()))
    first = next(iter(tensors.values()))
    if not first.is_cuda:
        raise RuntimeError("tts_b200: the model is on %s; this implementation has no CPU path - move it to a CUDA "
                           "device (model.to('cuda'))" % first.device)
    hit = _TRAIN_ENGINES.get(module)
    if hit is not None and hit[1] == first.device and all(hit[0].w.get(prefix + k) is v for k, v in tensors.items()):
        return hit[0]
    eng = TrainEngine({prefix + k: v for k, v in tensors.items()}, hparams, first.device)
    _TRAIN_ENGINES[module] = (eng, first.device)
    return eng


def needs_backward(module, *inputs):
    """True when autograd is recording and something this call touches requires a gradient - in train() AND in eval()
    mode (fine-tuning with frozen dropout must not silently return detached outputs).  Calls under torch.no_grad()
    (synthesize.eval_batch, also with its `decoder.train()`) are inference calls."""
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for p in module.parameters()) or any(
        torch.is_tensor(t) and t.requires_grad for t in inputs)


def no_backward(module, what, *inputs):
    """The differentiable entry points are Tacotron / Encoder / Decoder / Postnet (tts_b200.autograd); the smaller
    building blocks are inference-only when called on their own."""
    if needs_backward(module, *inputs):
        raise NotImplementedError(
            "tts_b200: %s called on its own has no backward pass (the differentiable entry points are Tacotron, Encoder, "
            "Decoder and Postnet); call it under torch.no_grad() or freeze its parameters" % what)


def param_names(module, prefix):
    return [prefix + n for n, _ in module.named_parameters()]


class FFNLayer(nn.Module):
    def __init__(self, input_size, hidden_size, output_size, dropout_rate=0.1):
        super().__init__()
        self.input_layer = nn.Linear(input_size, hidden_size, bias=False)
        self.dropout = nn.Dropout(dropout_rate)
        self.output_layer = nn.Linear(hidden_size, output_size, bias=False)

    def forward(self, inputs):
        """W2 . relu(W1 . x), bias-free (reference modules.py:14-20)."""
        no_backward(self, "FFNLayer", inputs)
        x = N.f32c(inputs)
        flat = x.reshape(-1, x.shape[-1])
        hid = ops.linear(flat, self.input_layer.weight, act=ops.ACT_RELU)
        out = ops.linear(hid, self.output_layer.weight)
        return out.view(*x.shape[:-1], out.shape[-1])


def _stack_lists(self, n_layers, in_size_first, hidden, heads, rate, with_cross):
    """Registers the per-layer ModuleLists in the reference's order (so state_dict() lists the
    same keys in the same order) and fills them layer by layer like the reference's loop."""
    self.self_attentions = nn.ModuleList()
    self.attn_layer_norms = nn.ModuleList()
    if with_cross:
        self.encdec_attentions = nn.ModuleList()
        self.encdec_layer_norms = nn.ModuleList()
    self.ffn_layers = nn.ModuleList()
    self.ffn_layer_norms = nn.ModuleList()
    self.pe_scale = nn.Parameter(torch.tensor(1.0))
    self.dropout = nn.Dropout(rate)
    for i in range(n_layers):
        width = in_size_first if i == 0 else hidden
        self.attn_layer_norms.append(nn.LayerNorm(width, eps=1e-6))
        self.self_attentions.append(MultiheadAttention(width, width, True, heads, dropout_rate=rate))
        if with_cross:
            self.encdec_layer_norms.append(nn.LayerNorm(width, eps=1e-6))
            self.encdec_attentions.append(MultiheadAttention(hidden, hidden, False, heads, dropout_rate=rate))
        self.ffn_layer_norms.append(nn.LayerNorm(hidden, eps=1e-6))
        self.ffn_layers.append(FFNLayer(hidden, hidden * 4, hidden, dropout_rate=rate))
    self.output_layer_norm = nn.LayerNorm(hidden, eps=1e-6)


class TransformerEncoder(nn.Module):
    def __init__(self, input_size, hparams):
        super().__init__()
        self.hparams = hparams
        _stack_lists(self, hparams.n_encoder_layer, input_size, hparams.encoder_hidden, hparams.n_attention_head,
                     hparams.transformer_dropout_rate, with_cross=False)

    def forward(self, inputs, input_lengths):
        """inputs: embedded text [B,S,E] -> [B,S,E] (reference modules.py:58-69)."""
        no_backward(self, "TransformerEncoder", inputs)
        eng = engine_for(self, "encoder.encoder.", self.hparams)
        if self.training and self.dropout.p > 0:
            eng.warn_dropout("TransformerEncoder")
        b, s, e = inputs.shape
        y = eng.encoder_stack(None, N.f32c(inputs).view(b * s, e), input_lengths, b, s)
        return y.view(b, s, -1)


class TransformerDecoder(nn.Module):
    def __init__(self, input_size, hparams):
        super().__init__()
        self.hparams = hparams
        _stack_lists(self, hparams.n_decoder_layer, input_size, hparams.decoder_hidden, hparams.n_attention_head,
                     hparams.transformer_dropout_rate, with_cross=True)

    def forward(self, inputs, targets, input_lengths, target_lengths):
        """inputs: encoder memory [B,S,D]; targets: prenet outputs [B,T,D] ->
        (outputs [B,T,D], {'self': [...], 'encdec': [...]})  (reference modules.py:123-145)."""
        no_backward(self, "TransformerDecoder", inputs, targets)
        eng = engine_for(self, "decoder.decoder.", self.hparams)
        if self.training and self.dropout.p > 0:
            eng.warn_dropout("TransformerDecoder")
        b, t, d = targets.shape
        o, _, align = eng.decoder_stack(inputs, N.f32c(targets).view(b * t, d), input_lengths, target_lengths, b, t)
        return o.view(b, t, d), align
