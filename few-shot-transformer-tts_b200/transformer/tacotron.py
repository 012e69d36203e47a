"""Tacotron / Encoder / Decoder / DecoderPrenet / Postnet, compute_loss, initialize_variables and
learning_rate_schedule with the reference's API and state-dict schema (reference:
transformer/tacotron.py), running on the sm_100a kernels of libtts_b200.so.

The autoregressive loop of the reference's synthesize.eval_batch calls
``decoder(enc_outputs, input_lengths, all_frames_so_far + 1, target_lengths, leave_one=True)``
once per frame and re-runs the whole decoder each time.  ``Decoder.forward`` recognises that
call pattern and runs ONE K/V-cached step instead, so the unchanged loop costs O(T) instead of
O(T^2) decoder rows while returning tensors of the shapes the loop expects.
"""
import torch
from torch import nn
from torch.nn import functional as F

from tts_b200 import _native as N
from tts_b200 import ops
from transformer.common import impute, mask_reduce, truncated_normal, variance_scaling_initializer
from tts_b200 import autograd as AG
from transformer.modules import (TransformerDecoder, TransformerEncoder, engine_for, needs_backward, no_backward,
                                 param_names, train_engine_for)


class Encoder(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.embed = nn.Embedding(hparams.vocab_size, hparams.embed_size)
        if hparams.multi_speaker:
            self.speaker_embed = nn.Embedding(hparams.max_num_speaker, hparams.speaker_embedding_size)
            self.speaker_layer = nn.Linear(hparams.speaker_embedding_size, hparams.speaker_embedding_size)
        if hparams.multi_lingual:
            self.language_embed = nn.Linear(hparams.max_num_language, hparams.language_embedding_size, bias=False)
            self.language_layer = nn.Linear(hparams.language_embedding_size, hparams.language_embedding_size)
        self.encoder = TransformerEncoder(hparams.embed_size, hparams)

    def _cond(self, width, w2, b2, **src):
        eng = engine_for(self, "encoder.", self.hparams)
        batch = (src["ids"] if src.get("ids") is not None else src["vec"]).shape[0]
        mem = torch.empty((batch, 1, width), device=eng.device, dtype=torch.float32)
        ops.cond_embed(mem, 0, w2, b2, **src)
        return mem[:, 0]

    def get_language_embed(self, x):
        """softsign(language_layer(language_embed(one_hot)))  (reference tacotron.py:21-25)."""
        return self._cond(self.hparams.language_embedding_size, self.language_layer.weight, self.language_layer.bias,
                          vec=N.f32c(x), w1=self.language_embed.weight)

    def get_speaker_embed(self, x):
        """softsign(speaker_layer(speaker_embed(id)))  (reference tacotron.py:27-31)."""
        return self._cond(self.hparams.speaker_embedding_size, self.speaker_layer.weight, self.speaker_layer.bias,
                          ids=x.long().contiguous(), w1=self.speaker_embed.weight)

    def forward(self, inputs, input_lengths, input_spk_ids=None, input_language_vecs=None):
        """[B,S] token ids -> encoder memory [B,S,encoder_hidden(+spk)(+lang)] (reference tacotron.py:33-44)."""
        if needs_backward(self):   # training step: bf16 tensor-core forward, hand-written backward behind autograd
            eng = train_engine_for(self, "encoder.", self.hparams)
            return AG.EncoderFn.apply(eng, param_names(self, "encoder."), self.training, inputs, input_lengths, input_spk_ids,
                                      input_language_vecs, *self.parameters())
        eng = engine_for(self, "encoder.", self.hparams)
        if self.training and self.hparams.transformer_dropout_rate > 0:
            eng.warn_dropout("Encoder")
        return eng.encode(inputs, input_lengths, input_spk_ids, input_language_vecs)


class DecoderPrenet(nn.Module):
    def __init__(self, in_size, hidden_size, out_size, dropout_rate):
        super().__init__()
        self.dense0 = nn.Linear(in_size, hidden_size)
        self.dense1 = nn.Linear(hidden_size, hidden_size)
        self.dense_final = nn.Linear(hidden_size, out_size, bias=False)
        self.dropout = nn.Dropout(dropout_rate)

    def forward(self, x):
        """relu(dense0) -> relu(dense1) -> dense_final (reference tacotron.py:55-65; dropout is not applied)."""
        no_backward(self, "DecoderPrenet", x)
        x = N.f32c(x)
        flat = x.reshape(-1, x.shape[-1])
        h = ops.linear(flat, self.dense0.weight, bias=self.dense0.bias, act=ops.ACT_RELU)
        h = ops.linear(h, self.dense1.weight, bias=self.dense1.bias, act=ops.ACT_RELU)
        out = ops.linear(h, self.dense_final.weight)
        return out.view(*x.shape[:-1], out.shape[-1])


class Postnet(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.conv_layers = nn.ModuleList()
        self.batchnorm_layers = nn.ModuleList()
        self.dropout = nn.Dropout(hparams.decoder_dropout_rate)
        hidden, n = hparams.postnet_hidden, hparams.n_postnet_layer
        for i in range(n):
            cin = hparams.num_mels if i == 0 else hidden
            cout = hparams.num_mels if i == n - 1 else hidden
            self.conv_layers.append(nn.Conv1d(cin, cout, 5, stride=1, padding=2, bias=False))
            self.batchnorm_layers.append(nn.BatchNorm1d(cout))

    def forward(self, inputs, input_lengths):
        """[B,T,M] -> residual [B,T,M]: 5 x (impute, conv k5, BatchNorm, tanh, dropout) (reference tacotron.py:81-90)."""
        return self._run(inputs, input_lengths, add_input=False)

    def _run(self, inputs, input_lengths, add_input):
        if self.training or needs_backward(self, inputs):
            # batch-statistics BatchNorm + dropout (train() mode), differentiable when autograd is recording
            if not self.training:
                raise NotImplementedError("tts_b200: a differentiable Postnet with frozen (eval-mode) BatchNorm statistics "
                                          "is not built; call postnet.train() or run under torch.no_grad()")
            eng = train_engine_for(self, "postnet.", self.hparams)
            if needs_backward(self, inputs):
                return AG.PostnetFn.apply(eng, param_names(self, "postnet."), True, add_input, inputs, input_lengths,
                                          *self.parameters())
            with torch.no_grad():
                return eng.postnet_fwd(inputs, input_lengths, True, add_input)[0]
        return engine_for(self, "postnet.", self.hparams).postnet(inputs, input_lengths, add_input=add_input)


class Decoder(nn.Module):
    #: "all" | "encdec" | "none": which attention maps the cached decode records.  The reference returns both
    #: lists, but its callers only consume 'encdec' (synthesize.py:90-92) and merely iterate over 'self'
    #: (synthesize.py:59-61); the self maps cost L*B*H*t_max^2 floats of HBM (7.4 GB at B=32, t_max=1100; 98 GB at
    #: B=128, T=2000), so the default records 'encdec' only and returns an empty 'self' list.  Set to "all" to get
    #: the reference's full structure.
    record_alignments = "encdec"

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        in_size = hparams.encoder_hidden
        if hparams.multi_speaker:
            in_size += hparams.speaker_embedding_size
        if hparams.multi_lingual:
            in_size += hparams.language_embedding_size
        if in_size != hparams.decoder_hidden:
            raise ValueError("encoder_hidden (+speaker +language embedding) = %d must equal decoder_hidden = %d"
                             % (in_size, hparams.decoder_hidden))
        self.prenet = DecoderPrenet(hparams.num_mels, hparams.prenet_hidden, hparams.decoder_hidden,
                                    hparams.decoder_dropout_rate)
        self.decoder = TransformerDecoder(in_size, hparams)
        self.mel_net = nn.Linear(hparams.decoder_hidden, hparams.num_mels, bias=False)
        self.stop_net = nn.Linear(hparams.decoder_hidden, 1)
        self._inc = None  # incremental (K/V-cached) decode session

    # -- K/V-cached incremental path ------------------------------------------------------------------
    def _incremental(self, eng, encoder_outputs, input_lengths, targets, target_lengths):
        """Serve one call of the synthesize.py loop (synthesize.py:37-41) from the cache, or return None."""
        B, T, _ = targets.shape
        inc = self._inc
        key = (encoder_outputs.data_ptr(), tuple(encoder_outputs.shape), encoder_outputs._version)
        if T == 1:
            t_max = max(int(self.hparams.max_generation_frames), 1)
            S = encoder_outputs.shape[1]
            if (inc is None or inc["engine"] is not eng or inc["sess"].batch != B or inc["sess"].mem_len != S
                    or inc["sess"].t_max != t_max or inc["record"] != self.record_alignments):
                inc = {"engine": eng, "sess": eng.new_session(B, S, t_max, self.record_alignments),
                       "record": self.record_alignments}
                self._inc = inc
            # `m.eval(); m.decoder.train()` (eval.py:116-117): live dropout in the decoder, eval everywhere else
            hp = self.hparams
            drop = None
            if self.training and (hp.decoder_dropout_rate > 0 or hp.transformer_dropout_rate > 0):
                drop = (hp.decoder_dropout_rate, hp.transformer_dropout_rate, eng.next_dropout_seed())
            inc["sess"].begin(encoder_outputs, input_lengths, dropout=drop)
            inc["key"] = key
        elif (inc is None or inc["engine"] is not eng or inc.get("key") != key or inc["sess"].batch != B
              or inc["sess"].t + 1 != T or T > inc["sess"].t_max):
            return None
        sess = inc["sess"]
        t = T - 1
        if t > 0:  # the frame the caller appended after the previous step (synthesize.py:43)
            sess.frames[:, t - 1].copy_(targets[:, t - 1])
        sess.lengths.copy_(target_lengths)  # the caller owns the lengths/finished bookkeeping (synthesize.py:44-45)
        sess.step(1, update_state=False)
        return sess.frames[:, :T], sess.stop_logits[:, :T], sess.alignments(T)

    def forward(self, encoder_outputs, input_lengths, targets, target_lengths, leave_one=False):
        """-> (mels [B,T,M], stop_logits [B,T], {'self': [...], 'encdec': [...]})  (reference tacotron.py:107-116)."""
        if needs_backward(self, encoder_outputs):
            # training step.  The attention maps are not materialised (flash attention; train.py never reads them):
            # the alignment lists are returned empty.
            eng = train_engine_for(self, "decoder.", self.hparams)
            mels, stop = AG.DecoderFn.apply(eng, param_names(self, "decoder."), self.training, leave_one, encoder_outputs,
                                            input_lengths, targets, target_lengths, *self.parameters())
            return mels, stop, {"self": [], "encdec": []}
        # inside an utterance (T > 1) the cached session's engine serves the call without re-validating the module's
        # parameter set: one frame of synthesize.py's loop costs a kernel launch, not a walk over 100 tensors
        inc = self._inc
        mid = leave_one and inc is not None and targets.shape[1] > 1
        eng = inc["engine"] if mid else engine_for(self, "decoder.", self.hparams)
        if leave_one:
            out = self._incremental(eng, encoder_outputs, input_lengths, targets, target_lengths)
            if out is not None:
                return out
            if mid:
                eng = engine_for(self, "decoder.", self.hparams)
        if self.training and (self.hparams.decoder_dropout_rate > 0 or self.hparams.transformer_dropout_rate > 0):
            eng.warn_dropout("Decoder")
        return eng.decode_teacher_forced(encoder_outputs, input_lengths, targets, target_lengths, leave_one=leave_one)


class Tacotron(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.encoder = Encoder(hparams)
        self.decoder = Decoder(hparams)
        self.postnet = Postnet(hparams)

    def forward(self, inputs, input_lengths, mel_targets, target_lengths, input_spk_ids, input_language_vecs,
                **kwargs):
        """Teacher-forced pass (reference tacotron.py:126-133)."""
        enc_outputs = self.encoder(inputs, input_lengths, input_spk_ids, input_language_vecs)
        mel_bef, stop_logits, alignments = self.decoder(enc_outputs, input_lengths, mel_targets, target_lengths)
        # mel_aft = mel_bef + postnet(mel_bef): the add is fused into the last Postnet layer's epilogue
        mel_aft = self.postnet._run(mel_bef, target_lengths, add_input=True)
        return {"mel_bef": mel_bef, "mel_aft": mel_aft, "stop_logits": stop_logits, "alignments": alignments}


def _l2_selected(model):
    """The tensors the reference's L2 term covers, selected by name exactly as tacotron.py:144-146."""
    return [p for n, p in model.named_parameters()
            if "weight" in n and "layer_norm" not in n and "batchnorm" not in n
            and "encoder.speaker_embed" not in n and "encoder.embed" not in n]


_L2_TABLES = {}


def compute_loss(model, mel_targets, target_lengths, outputs, hparams):
    """The 7-key loss dict of the reference (tacotron.py:136-158): length-masked MSE before/after the
    Postnet, per-sample after-loss, stop BCE (pos_weight 5) and L2 on the name-selected weights.
    On CUDA the masked losses and their gradients come from ONE fused kernel (tts_loss_train) and the L2 term from one
    multi-tensor launch; both are differentiable through tts_b200.autograd.  `hparams.l2_in_optimizer = True` (not a
    reference key) reports the L2 value without a graph, for optimizers that apply reg_weight * W themselves
    (tts_b200.optim.FusedAdam: exactly the same gradient)."""
    if not outputs["mel_bef"].is_cuda:
        raise RuntimeError("tts_b200: compute_loss expects CUDA tensors (no CPU path)")
    bef_loss, aft_loss, stop_loss, aft_losses = AG.LossFn.apply(outputs["mel_bef"], outputs["mel_aft"], outputs["stop_logits"],
                                                                mel_targets, target_lengths)
    decayed = _l2_selected(model)
    key = tuple(p.data_ptr() for p in decayed)
    hit = _L2_TABLES.get(id(model))
    if hit is None or hit[0] != key:
        from tts_b200 import train_ops as TO
        tab = TO.MultiTable(decayed[0].device)
        tab.build_opt([(p.detach(), None, None, None, True) for p in decayed])
        hit = (key, tab)
        _L2_TABLES[id(model)] = hit
    if getattr(hparams, "l2_in_optimizer", False):
        with torch.no_grad():
            l2 = AG.L2Fn.apply(float(hparams.reg_weight), hit[1], *decayed)
    else:
        l2 = AG.L2Fn.apply(float(hparams.reg_weight), hit[1], *decayed)
    return {"loss": bef_loss + aft_loss + l2 + stop_loss, "bef_loss": bef_loss, "aft_loss": aft_loss,
            "aft_losses": aft_losses, "mse_loss": (bef_loss + aft_loss) / 2, "l2": l2, "stop_loss": stop_loss}


def initialize_variables(model):
    """Seed-for-seed identical to the reference (tacotron.py:161-173): walks state_dict() in order and
    draws N(0,1) for the token embedding, truncated normal (std 0.5) for the speaker / language
    embeddings, fan-average variance scaling for every other non-norm weight, zeros for biases."""
    fresh = {}
    for name, tensor in model.state_dict().items():
        norm = "layer_norm" in name or "batchnorm" in name
        if name == "encoder.embed.weight":
            fresh[name] = torch.normal(mean=0, std=1, size=tensor.shape).to(tensor.device)
        elif name in ("encoder.speaker_embed.weight", "encoder.language_embed.weight"):
            fresh[name] = truncated_normal(tensor, mean=0, std=0.5).to(tensor.device)
        elif "weight" in name and not norm:
            fresh[name] = variance_scaling_initializer(tensor).to(tensor.device)
        elif "bias" in name:
            fresh[name] = torch.zeros_like(tensor)
    model.load_state_dict(fresh, strict=False)


def learning_rate_schedule(global_step, hp):
    """LambdaLR factor: 1 during warm-up, then exponential decay floored at min_lr/max_lr
    (reference tacotron.py:176-179)."""
    past = max(global_step - hp.warmup_steps, 0)
    return max(hp.min_lr / hp.max_lr, hp.lr_decay_rate ** (past / hp.lr_decay_step))
