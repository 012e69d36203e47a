"""Host-side engine of the mel path: owns nothing but references to weights (CUDA tensors keyed
by the reference's state-dict names), derived weight caches, and decode sessions (K/V caches,
frame buffers).  All arithmetic happens in libtts_b200.so.

The `transformer/` package (the drop-in mirror of the reference API) is a thin nn.Module shell
around this engine; tests and bench.py can also drive it directly from a state dict.
"""
import ctypes as C
import math
import warnings

import numpy as np
import torch

from . import _native as N
from . import ops


def sinusoid_table(length: int, channels: int) -> torch.Tensor:
    """[sin | cos] position table, float64 math then fp32 (transformer/common.py:4-29).
    Built ONCE per (length, channels) and kept on the device; the reference rebuilds it in numpy
    and copies it host-to-device on every forward."""
    half = channels // 2
    inc = math.log(1e4) / (half - 1)
    inv = np.exp(np.arange(half) * -inc)
    ang = np.arange(length)[:, None] * inv[None, :]
    tab = np.concatenate([np.sin(ang), np.cos(ang)], axis=1)
    if channels % 2:
        tab = np.pad(tab, [[0, 0], [0, 1]])
    return torch.from_numpy(tab.astype(np.float32))


def pack_rows(wt, const=None, rowsum=None, ksplit=1):
    """Packed operand of the pipelined decode kernel (include/tts_b200.h, pk_*): [ksplit][N][K/ksplit + 16] with
    row = weights | additive constant | LayerNorm row sum | zeros.  Within every 16-float chunk of the weights,
    float 4t+i holds k = t + 4i, so that lane t of an MMA quad reads its four B-operand values (two k8-steps)
    with one 128-bit load (pipelined.cu: mma_tiles)."""
    with torch.no_grad():
        n, k = wt.shape
        kc = k // ksplit
        assert kc * ksplit == k and kc % 16 == 0, (k, ksplit)
        out = torch.zeros((ksplit, n, kc + 16), device=wt.device, dtype=torch.float32)
        rows = wt.detach().view(n, ksplit, kc).permute(1, 0, 2)
        out[:, :, :kc] = rows.reshape(ksplit, n, kc // 16, 4, 4).transpose(3, 4).reshape(ksplit, n, kc)
        if const is not None:
            out[0, :, kc] = const.detach().view(-1)
        if rowsum is not None:
            out[0, :, kc + 1] = rowsum.view(-1)
    return out


class DecodeSession:
    """Device state of one autoregressive batch: self/cross K/V caches, frames, lengths.
    Layout and meaning of every buffer: include/tts_b200.h (TtsDecodeState)."""

    def __init__(self, engine, batch, mem_len, t_max, record_align="encdec"):
        cfg, dev = engine.cfg, engine.device
        L, D, H, M = cfg.n_decoder_layer, cfg.decoder_hidden, cfg.n_attention_head, cfg.num_mels
        dh = D // H
        self.engine, self.batch, self.mem_len, self.t_max = engine, batch, mem_len, t_max
        f32 = dict(device=dev, dtype=torch.float32)
        self.self_k = torch.empty((L, batch, H, t_max, dh), **f32)
        self.self_v = torch.empty((L, batch, H, t_max, dh), **f32)
        self.cross_k = torch.empty((L, batch, H, mem_len, dh), **f32)
        self.cross_v = torch.empty((L, batch, H, mem_len, dh), **f32)
        self.lengths = torch.ones((batch,), device=dev, dtype=torch.int32)
        self.finished = torch.zeros((batch,), device=dev, dtype=torch.uint8)
        self.frames = torch.zeros((batch, t_max, M), **f32)
        self.stop_logits = torch.zeros((batch, t_max), **f32)
        self.align_self = torch.zeros((L, batch, H, t_max, t_max), **f32) if record_align == "all" else None
        self.align_cross = (torch.zeros((L, batch, H, t_max, mem_len), **f32)
                            if record_align in ("all", "encdec") else None)
        self.counters = torch.zeros((2,), device=dev, dtype=torch.int32)  # [step, n_unfinished]
        nbytes = N.load().tts_decode_scratch_bytes(C.byref(engine.decoder_weights()), batch, mem_len, t_max)
        self.scratch = torch.empty((nbytes // 4 + 64,), **f32)
        self.memory = None
        self.input_lengths = None
        self.t = 0
        self._st = None
        self._dw = None
        self.dropout = (0.0, 0.0, 0)   # (prenet rate, transformer rate, seed): decoder.train() at synthesis time

    def state(self):
        st = N.DecodeState()
        st.batch, st.mem_len, st.t_max = self.batch, self.mem_len, self.t_max
        st.memory = N.ptr(self.memory)
        st.input_lengths = N.ptr(self.input_lengths)
        for name in ("self_k", "self_v", "cross_k", "cross_v", "lengths", "finished", "frames", "stop_logits",
                     "align_self", "align_cross", "scratch"):
            setattr(st, name, N.ptr(getattr(self, name)))
        st.step_counter = self.counters.data_ptr()
        st.n_unfinished = self.counters.data_ptr() + 4
        st.drop_p_prenet, st.drop_p_transformer, st.drop_seed = self.dropout
        return st

    def begin(self, memory, input_lengths, dropout=None):
        """Once per utterance batch: cross K/V from the encoder memory, reset lengths/finished/t.
        dropout = (prenet rate, transformer rate, seed) decodes with the reference's `decoder.train()` semantics
        (eval.py:116-117); None / zeros = deterministic."""
        self.dropout = (0.0, 0.0, 0) if dropout is None else (float(dropout[0]), float(dropout[1]), int(dropout[2]))
        self.memory = N.f32c(memory)
        self.input_lengths = ops._i32(input_lengths)
        assert self.memory.shape == (self.batch, self.mem_len, self.engine.cfg.decoder_hidden), self.memory.shape
        self._st = self.state()
        # the weight struct is validated against the live parameters (pointer + version of every decoder tensor) once per
        # utterance, not once per step: parameters do not change while an utterance is being decoded
        self._dw = self.engine.decoder_weights()
        N.check(N.load().tts_decode_begin(C.byref(self._dw), C.byref(self._st),
                                          N.stream_ptr(self.engine.device)), "decode_begin")
        self.t = 0

    def step(self, n_steps=1, update_state=True, impl=0):
        if self.t + n_steps > self.t_max:   # the kernels also clamp against the DEVICE step counter and report -2
            raise RuntimeError("tts_b200: decode session overflow (t=%d + %d steps > t_max=%d)"
                               % (self.t, n_steps, self.t_max))
        dw = self._dw if getattr(self, "_dw", None) is not None else self.engine.decoder_weights()
        N.check(N.load().tts_decode_steps(C.byref(dw), C.byref(self._st), n_steps, None, 0,
                                          1 if update_state else 0, impl, N.stream_ptr(self.engine.device)),
                "decode_steps")
        self.t += n_steps

    def phase_profile(self, n_phases=None):
        """[n_phases, 8] SM-clock stamps of CTA 0 for the last step.  Fused kernel (impl 3): 0 start, 1 compute
        done, 2 barrier passed; GEMM phases also 3 activations staged, 4 LayerNorm done, 5 weights landed, 6 FMA +
        reduction done (last pass), 7 epilogue done (last pass).  Pipelined kernel (impl 4): for row group
        g in {0, 1}: 3g wait for the group's inputs begins, 3g+1 inputs ready, 3g+2 group done and handed to the
        signaler; GEMM phases, group 0: 6 fragments loaded, 7 weights landed, 8 products done, 9 cross-warp buffer
        complete, 10 epilogue done (rows of 16 stamps)."""
        n = n_phases if n_phases is not None else 3 + 8 * self.engine.cfg.n_decoder_layer + 1
        stride = 8 if n_phases is None else 16
        buf = (C.c_int64 * (stride * n))()
        N.check(N.load().tts_decode_profile(C.byref(self.engine.decoder_weights()), C.byref(self._st), buf, stride * n),
                "decode_profile")
        return np.array(list(buf), dtype=np.int64).reshape(n, stride)

    def alignments(self, t):
        """Views shaped like the reference's (attention.py:88): [B,H,T_kv,T_q] per layer."""
        out = {"self": [], "encdec": []}
        L = self.engine.cfg.n_decoder_layer
        for l in range(L):
            if self.align_self is not None:
                out["self"].append(self.align_self[l][:, :, :t, :t].transpose(2, 3))
            if self.align_cross is not None:
                out["encdec"].append(self.align_cross[l][:, :, :t, :].transpose(2, 3))
        return out


class TtsEngine:
    """Forward paths of the model over a dict of CUDA weights (reference state-dict names)."""

    def __init__(self, weights, cfg, device):
        self.w, self.cfg, self.device = weights, cfg, torch.device(device)
        N.load()
        mw = cfg.encoder_hidden
        mw += cfg.speaker_embedding_size if cfg.multi_speaker else 0
        mw += cfg.language_embedding_size if cfg.multi_lingual else 0
        if mw != cfg.decoder_hidden:
            # the reference only works when these agree (layer 0 of the decoder is built with
            # the memory width but fed the prenet output of width decoder_hidden)
            raise ValueError("encoder_hidden + speaker + language embedding (%d) must equal decoder_hidden (%d)"
                             % (mw, cfg.decoder_hidden))
        self._pe = {}
        self._derived = {}
        self._dec_w = None
        self._dec_key = None
        self._warned_dropout = False

    # ---- construction helpers --------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, params, cfg, device="cuda:0"):
        dev = torch.device(device)
        w = {k: (v.to(dev).float().contiguous() if v.dtype.is_floating_point else v.to(dev))
             for k, v in params.items()}
        return cls(w, cfg, dev)

    def pe(self, length, channels):
        key = channels
        cur = self._pe.get(key)
        if cur is None or cur.shape[0] < length:
            n = max(length, 256)
            n = 1 << (n - 1).bit_length()
            cur = sinusoid_table(n, channels).to(self.device)
            self._pe[key] = cur
            self._dec_w = None
        return cur

    def _versions(self, names):
        return tuple((self.w[n].data_ptr(), self.w[n]._version) for n in names)

    def decoder_weights(self, t_max=0):
        """The TtsDecoderWeights struct (device pointers borrowed from the live parameters)."""
        cfg = self.cfg
        pe = self.pe(max(t_max, 1), cfg.decoder_hidden)
        names = [n for n in self.w if n.startswith("decoder.")]
        key = (self._versions(names), pe.data_ptr())
        if self._dec_w is not None and self._dec_key == key:
            return self._dec_w
        w, dw = self.w, N.DecoderWeights()
        dw.n_layers, dw.d_model, dw.n_heads = cfg.n_decoder_layer, cfg.decoder_hidden, cfg.n_attention_head
        dw.d_ffn, dw.n_mels, dw.prenet_hidden = 4 * cfg.decoder_hidden, cfg.num_mels, cfg.prenet_hidden
        g = lambda n: self._c(w[n]).data_ptr()
        dw.prenet_w0, dw.prenet_b0 = g("decoder.prenet.dense0.weight"), g("decoder.prenet.dense0.bias")
        dw.prenet_w1, dw.prenet_b1 = g("decoder.prenet.dense1.weight"), g("decoder.prenet.dense1.bias")
        dw.prenet_w2 = g("decoder.prenet.dense_final.weight")
        dw.pe_scale, dw.pe_table = g("decoder.decoder.pe_scale"), pe.data_ptr()
        dw.ln_out_g, dw.ln_out_b = g("decoder.decoder.output_layer_norm.weight"), g("decoder.decoder.output_layer_norm.bias")
        dw.w_mel, dw.w_stop, dw.b_stop = g("decoder.mel_net.weight"), g("decoder.stop_net.weight"), g("decoder.stop_net.bias")
        p = "decoder.decoder."
        for l in range(cfg.n_decoder_layer):
            lw = dw.layer[l]
            lw.ln_self_g, lw.ln_self_b = g(f"{p}attn_layer_norms.{l}.weight"), g(f"{p}attn_layer_norms.{l}.bias")
            lw.w_qkv = g(f"{p}self_attentions.{l}.qkv_transform.weight")
            lw.w_self_out = g(f"{p}self_attentions.{l}.output_transform.weight")
            lw.ln_cross_g, lw.ln_cross_b = g(f"{p}encdec_layer_norms.{l}.weight"), g(f"{p}encdec_layer_norms.{l}.bias")
            lw.w_cross_q = g(f"{p}encdec_attentions.{l}.q_transform.weight")
            lw.w_cross_kv = g(f"{p}encdec_attentions.{l}.kv_transform.weight")
            lw.w_cross_out = g(f"{p}encdec_attentions.{l}.output_transform.weight")
            lw.ln_ffn_g, lw.ln_ffn_b = g(f"{p}ffn_layer_norms.{l}.weight"), g(f"{p}ffn_layer_norms.{l}.bias")
            lw.w_ffn_in = g(f"{p}ffn_layers.{l}.input_layer.weight")
            lw.w_ffn_out = g(f"{p}ffn_layers.{l}.output_layer.weight")
        # LayerNorm folded into the projection that follows it (packed operands of the fused kernel)
        keep = []

        def fold(wname, gname, bname):
            with torch.no_grad():
                wt, gam, bet = w[wname].detach(), w[gname].detach(), w[bname].detach()
                wf = (wt * gam[None, :]).contiguous()
                cf = ops.linear(bet[None, :].contiguous(), wt.contiguous()).view(-1)
                sf = wf.double().sum(dim=1).float().contiguous()   # row sums: LayerNorm applied after the product
            keep.extend([wf, cf, sf])
            return wf, cf, sf

        def pack(wt, const=None, rowsum=None, ksplit=1):
            out = pack_rows(wt, const, rowsum, ksplit)
            keep.append(out)
            return out.data_ptr()

        ksplit = (dw.d_ffn + 767) // 768   # K split of FFN-out in the pipelined kernel (pipelined.cu: ksplit_for)
        if dw.d_ffn % ksplit:
            ksplit = 1
        for l in range(cfg.n_decoder_layer):
            lw = dw.layer[l]
            wf, cf, sf = fold(f"{p}self_attentions.{l}.qkv_transform.weight", f"{p}attn_layer_norms.{l}.weight",
                              f"{p}attn_layer_norms.{l}.bias")
            lw.w_qkv_ln, lw.c_qkv_ln, lw.pk_qkv = wf.data_ptr(), cf.data_ptr(), pack(wf, cf, sf)
            wf, cf, sf = fold(f"{p}encdec_attentions.{l}.q_transform.weight", f"{p}encdec_layer_norms.{l}.weight",
                              f"{p}encdec_layer_norms.{l}.bias")
            lw.w_cross_q_ln, lw.c_cross_q_ln, lw.pk_cross_q = wf.data_ptr(), cf.data_ptr(), pack(wf, cf, sf)
            wf, cf, sf = fold(f"{p}ffn_layers.{l}.input_layer.weight", f"{p}ffn_layer_norms.{l}.weight",
                              f"{p}ffn_layer_norms.{l}.bias")
            lw.w_ffn_in_ln, lw.c_ffn_in_ln, lw.pk_ffn_in = wf.data_ptr(), cf.data_ptr(), pack(wf, cf, sf)
            lw.pk_self_out = pack(w[f"{p}self_attentions.{l}.output_transform.weight"])
            lw.pk_cross_out = pack(w[f"{p}encdec_attentions.{l}.output_transform.weight"])
            lw.pk_ffn_out = pack(w[f"{p}ffn_layers.{l}.output_layer.weight"], ksplit=ksplit)
        wm, cm, sm_ = fold("decoder.mel_net.weight", p + "output_layer_norm.weight", p + "output_layer_norm.bias")
        ws, cs, ss_ = fold("decoder.stop_net.weight", p + "output_layer_norm.weight", p + "output_layer_norm.bias")
        cout = torch.cat([cm, cs]).contiguous()
        sout = torch.cat([sm_, ss_]).contiguous()
        keep.extend([cout, sout])
        dw.w_mel_ln, dw.w_stop_ln, dw.c_out_ln = wm.data_ptr(), ws.data_ptr(), cout.data_ptr()
        # rows M+1.. of pk_final: the first prenet layer applied to the frame being emitted, folded through mel_net
        # (pipelined.cu: the final phase then produces p0 of the NEXT step and steps t > 0 skip that prenet phase)
        with torch.no_grad():
            w0, b0 = w["decoder.prenet.dense0.weight"].detach(), w["decoder.prenet.dense0.bias"].detach()
            wf = ops.linear(w0.contiguous(), wm.t().contiguous())                   # [P, D] = W0 @ W_mel_ln
            cf = ops.linear(cm[None, :].contiguous(), w0.contiguous()).view(-1) + b0  # W0 c_mel + b0
            sf = wf.double().sum(dim=1).float()
        keep.extend([wf, cf, sf])
        dw.pk_final = pack(torch.cat([wm, ws, wf], 0), torch.cat([cout, cf]), torch.cat([sout, sf]))
        dw.pk_pre0 = pack(w["decoder.prenet.dense0.weight"], w["decoder.prenet.dense0.bias"])
        dw.pk_pre1 = pack(w["decoder.prenet.dense1.weight"], w["decoder.prenet.dense1.bias"])
        dw.pk_pre2 = pack(w["decoder.prenet.dense_final.weight"])
        dw.pk_ksplit = ksplit
        self._dec_keep = keep  # owns the packed tensors for as long as the struct is cached
        self._dec_w, self._dec_key = dw, key
        return dw

    @staticmethod
    def _c(t):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("tts_b200: weights must be contiguous fp32 CUDA tensors (no CPU fallback)")
        return t

    def _postnet_packed(self):
        """Conv weights repacked [Cout][5][Cin] and eval-mode BatchNorm folded into a per-channel
        scale/shift (tacotron.py:85-86); rebuilt whenever a source tensor changes."""
        cfg, w = self.cfg, self.w
        names = [n for n in w if n.startswith("postnet.")]
        key = self._versions(names)
        hit = self._derived.get("postnet")
        if hit is not None and hit[0] == key:
            return hit[1]
        layers = []
        with torch.no_grad():
            for i in range(cfg.n_postnet_layer):
                cw = w[f"postnet.conv_layers.{i}.weight"]
                b = f"postnet.batchnorm_layers.{i}."
                scale = w[b + "weight"] / torch.sqrt(w[b + "running_var"] + 1e-5)
                shift = w[b + "bias"] - w[b + "running_mean"] * scale
                layers.append((cw.permute(0, 2, 1).reshape(cw.shape[0], -1).contiguous(), scale.contiguous(),
                               shift.contiguous()))
        self._derived["postnet"] = (key, layers)
        return layers

    def warn_dropout(self, what):
        if not self._warned_dropout:
            warnings.warn("tts_b200: %s is in train() mode under no_grad on a full-sequence (non-incremental) call; this path "
                          "does not apply dropout (the K/V-cached decode and the training step do)" % what)
            self._warned_dropout = True

    def next_dropout_seed(self):
        self._drop_calls = getattr(self, "_drop_calls", 0) + 1
        return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + self._drop_calls * 0xD1B54A32D192ED03) & 0x7fffffffffffffff

    # ---- encoder (tacotron.py:33-44, modules.py:49-69) ------------------------------------------
    def encoder_stack(self, ids, embedded, input_lengths, batch, seq):
        """TransformerEncoder.forward: mask + PE, 6 x (LN, self-attention, LN, FFN), final LN.
        Either token `ids` [B,S] (embedding fused into the prologue) or `embedded` [B*S,E]."""
        cfg, w = self.cfg, self.w
        B, S, E, H = batch, seq, cfg.encoder_hidden, cfg.n_attention_head
        lens = ops._i32(input_lengths.to(self.device))
        p = "encoder.encoder."
        table = self._c(w["encoder.embed.weight"]) if ids is not None else embedded
        x = ops.embed_pe(ids, lens, table, self.pe(S, cfg.embed_size), w[p + "pe_scale"], B, S)
        dh = E // H
        for l in range(cfg.n_encoder_layer):
            h = ops.layernorm(x, w[f"{p}attn_layer_norms.{l}.weight"], w[f"{p}attn_layer_norms.{l}.bias"])
            qkv = ops.linear(h, w[f"{p}self_attentions.{l}.qkv_transform.weight"])
            base = qkv.data_ptr()
            ctx, _ = ops.attention(base, 3 * E, base + 4 * E, 3 * E, base + 8 * E, 3 * E, B, H, S, S, dh, False, lens,
                                   False, self.device)
            x = ops.linear(ctx, w[f"{p}self_attentions.{l}.output_transform.weight"], residual=x)
            h = ops.layernorm(x, w[f"{p}ffn_layer_norms.{l}.weight"], w[f"{p}ffn_layer_norms.{l}.bias"])
            hid = ops.linear(h, w[f"{p}ffn_layers.{l}.input_layer.weight"], act=ops.ACT_RELU)
            x = ops.linear(hid, w[f"{p}ffn_layers.{l}.output_layer.weight"], residual=x)
        return ops.layernorm(x, w[p + "output_layer_norm.weight"], w[p + "output_layer_norm.bias"])

    def encode(self, inputs, input_lengths, spk_ids=None, lang_vecs=None):
        cfg, w = self.cfg, self.w
        B, S = inputs.shape
        E = cfg.encoder_hidden
        if cfg.embed_size != E:
            raise ValueError("embed_size (%d) must equal encoder_hidden (%d)" % (cfg.embed_size, E))
        y = self.encoder_stack(inputs.to(self.device).long().contiguous(), None, input_lengths, B, S)
        width = cfg.decoder_hidden
        mem = torch.empty((B, S, width), device=self.device, dtype=torch.float32)
        mem[:, :, :E].copy_(y.view(B, S, E))  # strided device copy (the torch.cat of tacotron.py:39,43)
        off = E
        if cfg.multi_speaker:
            ops.cond_embed(mem, off, w["encoder.speaker_layer.weight"], w["encoder.speaker_layer.bias"],
                           w1=w["encoder.speaker_embed.weight"], ids=spk_ids.to(self.device).long().contiguous())
            off += cfg.speaker_embedding_size
        if cfg.multi_lingual:
            ops.cond_embed(mem, off, w["encoder.language_layer.weight"], w["encoder.language_layer.bias"],
                           vec=N.f32c(lang_vecs.to(self.device)), w1=w["encoder.language_embed.weight"])
        return mem

    # ---- teacher-forced decoder (tacotron.py:107-116, modules.py:108-145) ------------------------
    def prenet(self, targets2d):
        w, pp = self.w, "decoder.prenet."
        h = ops.linear(targets2d, w[pp + "dense0.weight"], bias=w[pp + "dense0.bias"], act=ops.ACT_RELU)
        h = ops.linear(h, w[pp + "dense1.weight"], bias=w[pp + "dense1.bias"], act=ops.ACT_RELU)
        return ops.linear(h, w[pp + "dense_final.weight"])

    def decoder_stack(self, memory, pre2d, input_lengths, target_lengths, batch, frames, want_align=True):
        """TransformerDecoder.forward over full sequences: impute + shift + PE, 6 x (causal self-attention,
        cross-attention, FFN), final LN, impute.  Returns (o [B*T,D], tg_len int32, align dict or None)."""
        cfg, w = self.cfg, self.w
        B, T = batch, frames
        S, D, H = memory.shape[1], cfg.decoder_hidden, cfg.n_attention_head
        dh = D // H
        in_len, tg_len = ops._i32(input_lengths.to(self.device)), ops._i32(target_lengths.to(self.device))
        mem2d = N.f32c(memory).view(B * S, D)
        p = "decoder.decoder."
        x = ops.shift_pe(pre2d, tg_len, self.pe(T, D), w[p + "pe_scale"], B, T)
        self_align, cross_align = [], []
        for l in range(cfg.n_decoder_layer):
            h = ops.layernorm(x, w[f"{p}attn_layer_norms.{l}.weight"], w[f"{p}attn_layer_norms.{l}.bias"])
            qkv = ops.linear(h, w[f"{p}self_attentions.{l}.qkv_transform.weight"])
            base = qkv.data_ptr()
            ctx, al = ops.attention(base, 3 * D, base + 4 * D, 3 * D, base + 8 * D, 3 * D, B, H, T, T, dh, True, None,
                                    want_align, self.device)
            self_align.append(al)
            x = ops.linear(ctx, w[f"{p}self_attentions.{l}.output_transform.weight"], residual=x)
            h = ops.layernorm(x, w[f"{p}encdec_layer_norms.{l}.weight"], w[f"{p}encdec_layer_norms.{l}.bias"])
            q = ops.linear(h, w[f"{p}encdec_attentions.{l}.q_transform.weight"])
            kv = ops.linear(mem2d, w[f"{p}encdec_attentions.{l}.kv_transform.weight"])
            kb = kv.data_ptr()
            ctx, al = ops.attention(q.data_ptr(), D, kb, 2 * D, kb + 4 * D, 2 * D, B, H, T, S, dh, False, in_len,
                                    want_align, self.device)
            cross_align.append(al)
            x = ops.linear(ctx, w[f"{p}encdec_attentions.{l}.output_transform.weight"], residual=x)
            h = ops.layernorm(x, w[f"{p}ffn_layer_norms.{l}.weight"], w[f"{p}ffn_layer_norms.{l}.bias"])
            hid = ops.linear(h, w[f"{p}ffn_layers.{l}.input_layer.weight"], act=ops.ACT_RELU)
            x = ops.linear(hid, w[f"{p}ffn_layers.{l}.output_layer.weight"], residual=x)
        o = ops.layernorm(x, w[p + "output_layer_norm.weight"], w[p + "output_layer_norm.bias"], row_len=tg_len,
                          rows_per_batch=T)
        align = None
        if want_align:  # reference layout [B,H,T_kv,T_q] (attention.py:88) as transposed views
            align = {"self": [a.transpose(2, 3) for a in self_align],
                     "encdec": [a.transpose(2, 3) for a in cross_align]}
        return o, tg_len, align

    def decode_teacher_forced(self, memory, input_lengths, targets, target_lengths, leave_one=False,
                              want_align=True):
        w = self.w
        B, T, M = targets.shape
        pre = self.prenet(N.f32c(targets).view(B * T, M))
        # leave_one zeroes the last prenet row, which the shift-right then drops: a no-op (tacotron.py:109-110)
        o, tg_len, align = self.decoder_stack(memory, pre, input_lengths, target_lengths, B, T, want_align)
        mels = ops.linear(o, w["decoder.mel_net.weight"], row_len=tg_len, rows_per_batch=T).view(B, T, M)
        stop = ops.linear(o, w["decoder.stop_net.weight"], bias=w["decoder.stop_net.bias"], row_len=tg_len,
                          rows_per_batch=T).view(B, T)
        return mels, stop, align

    # ---- postnet (tacotron.py:81-90), eval-mode BatchNorm ----------------------------------------
    def postnet(self, mels, lengths, add_input=False):
        """Returns the residual [B,T,M]; with add_input the last layer's epilogue adds `mels`
        (mel_aft = mel_bef + postnet(mel_bef), tacotron.py:131-132 / synthesize.py:56)."""
        cfg = self.cfg
        B, T, M = mels.shape
        lens = ops._i32(lengths.to(self.device))
        layers = self._postnet_packed()
        mels = N.f32c(mels)
        x = ops.pad_rows(mels, lens, B, T, pad=2)
        n = len(layers)
        for i, (wp, scale, shift) in enumerate(layers):
            last = i == n - 1
            cout = wp.shape[0]
            if last:
                out = torch.empty((B, T, cout), device=self.device, dtype=torch.float32)
            else:
                out = torch.zeros((B, T + 4, cout), device=self.device, dtype=torch.float32)
            # rows at or beyond the length are zeroed because the NEXT layer imputes its input
            # (tacotron.py:84); the last layer's output is not masked.
            ops.conv5(x, wp, scale, shift, ops.ACT_NONE if last else ops.ACT_TANH, None if last else lens, B, T, out,
                      out_padded=not last, residual=mels if (last and add_input) else None)
            x = out
        return x

    # ---- whole model, teacher forced (tacotron.py:126-133) ---------------------------------------
    def forward(self, batch, want_align=True):
        mem = self.encode(batch["inputs"], batch["input_lengths"], batch.get("input_spk_ids"),
                          batch.get("input_language_vecs"))
        tg = batch["mel_targets"].to(self.device)
        mel_bef, stop, align = self.decode_teacher_forced(mem, batch["input_lengths"], tg, batch["target_lengths"],
                                                          want_align=want_align)
        mel_aft = self.postnet(mel_bef, batch["target_lengths"], add_input=True)
        return {"mel_bef": mel_bef, "mel_aft": mel_aft, "stop_logits": stop, "alignments": align, "memory": mem}

    # ---- autoregressive synthesis (synthesize.py:17-72 as one call) ------------------------------
    def new_session(self, batch, mem_len, t_max, record_align="encdec"):
        self.decoder_weights(t_max)  # make sure the PE table covers t_max before pointers are taken
        return DecodeSession(self, batch, mem_len, t_max, record_align)

    def generate(self, batch, max_frames=None, record_align="encdec", chunk=32, impl=0, session=None,
                 memory=None, dropout=None, compact=None, waveform=False):
        """The whole of synthesize.eval_batch (synthesize.py:17-72) as one call: encoder, K/V-cached decode loop with the
        stop bookkeeping on the device (one 4-byte D2H poll per `chunk` steps), Postnet.

        compact: early-exit compaction of finished samples.  A finished sample's frames are
        zero from its stop on (modules.py:144) but its K/V streams would still be read on every step; when enough
        samples have finished to save a whole 16-row group of the decode kernel, the live rows (K/V caches, lengths,
        last frame) are gathered into a smaller session between two launches and decoding goes on there; frames and
        stop logits are scattered back at the end.  Per-row results do not depend on the batch composition (the
        K/V streams of a batch of <= 16 rows are split across SMs, so sums may differ in the last bits).  Default
        (None): on when no attention rows are recorded and dropout is off - alignment rows of a finished sample
        beyond its stop would stay zero, and the dropout masks are indexed by batch row.

        waveform: also run the stage after the path (synthesize.py:82: mel2wav of every utterance's mel_aft[:length]) as one
        batched Griffin-Lim call on the device (tts_b200/vocoder.py): out["wav"] [B, 200 (T - 1)] and out["wav_lengths"]."""
        cfg = self.cfg
        max_frames = cfg.max_generation_frames if max_frames is None else max_frames
        if memory is None:
            memory = self.encode(batch["inputs"], batch["input_lengths"], batch.get("input_spk_ids"),
                                 batch.get("input_language_vecs"))
        B, S, _ = memory.shape
        sess = session if session is not None else self.new_session(B, S, max_frames, record_align)
        sess.begin(memory, batch["input_lengths"].to(self.device), dropout=dropout)
        if compact is None:
            compact = sess.align_self is None and sess.align_cross is None and dropout is None and impl in (0, 4)
        if compact and (sess.align_self is not None or sess.align_cross is not None):
            raise ValueError("tts_b200: generate(compact=True) needs record_align='none'")
        top = sess                      # the caller-visible session: receives all frames / lengths at the end
        rows = None                     # original batch row of every row of `sess` (None: identity)
        parts = []                      # (session, rows, t_from, t_to) of every compacted stretch
        t_from = 0
        compactions = []
        done = 0
        all_finished = False
        while done < max_frames:
            n = min(chunk, max_frames - done)
            sess.step(n, update_state=True, impl=impl)
            done += n
            left = int(sess.counters[1].item())  # one 4-byte D2H per chunk instead of one per frame
            if left == -2 or left <= -(1 << 29):
                raise RuntimeError("tts_b200: decode stepped past the session's t_max (device step counter)")
            if left < 0:
                raise RuntimeError("tts_b200: the decode kernel reported a grid-barrier timeout")
            if left == 0:
                all_finished = True
                break
            if compact and done < max_frames and (left + 15) // 16 < (sess.batch + 15) // 16:
                parts.append((sess, rows, t_from, done))
                sess, rows = self._compact_session(sess, rows, left, done)
                t_from = done
                compactions.append((done, sess.batch))
        if sess is not top:             # scatter the compacted stretches back into the caller-visible session
            parts.append((sess, rows, t_from, done))
            # rows that had finished before a stretch were not decoded in it: their frames there are zero (modules.py:144)
            top.frames[:, parts[1][2]:done].zero_()
            top.stop_logits[:, parts[1][2]:done].zero_()
            for ps, pr, a, b in parts[1:]:
                top.frames[pr, a:b] = ps.frames[:, a:b]
                top.stop_logits[pr, a:b] = ps.stop_logits[:, a:b]
                top.lengths[pr] = ps.lengths
                top.finished[pr] = ps.finished
            top.t = done
        lengths = top.lengths.clone()
        # eval_batch stops right after the step in which the last sample fired (synthesize.py:35)
        t_gen = int(lengths.max().item()) if all_finished else done
        t_gen = min(t_gen, max_frames)
        mels = top.frames[:, :t_gen].contiguous()
        mel_aft = self.postnet(mels, lengths, add_input=True)
        out = {"mel_pre": mels, "mel_aft": mel_aft, "generated_lengths": lengths, "memory": memory,
               "stop_logits": top.stop_logits[:, :t_gen], "alignments": top.alignments(t_gen), "session": top,
               "compactions": compactions}
        if waveform:
            from .vocoder import GriffinLim
            if getattr(self, "_vocoder", None) is None:
                self._vocoder = GriffinLim(self.device)
            lens_host = [int(v) for v in lengths.tolist()]
            # an utterance shorter than one reflection of the STFT pad (7 frames) is converted with its zero frames up to 7
            wav, _ = self._vocoder(mel_aft, [min(max(n, 7), t_gen) for n in lens_host])
            out["wav"] = wav
            out["wav_lengths"] = [200 * max(n - 1, 0) for n in lens_host]
        return out

    def _compact_session(self, sess, rows, n_live, t):
        """Gather the unfinished rows of `sess` (decoded up to step t) into a session of n_live rows."""
        dev = self.device
        live = torch.nonzero(sess.finished == 0).view(-1)
        assert live.numel() == n_live, (live.numel(), n_live)
        cache = self.__dict__.setdefault("_compact_sessions", {})
        key = (n_live, sess.mem_len, sess.t_max)
        small = cache.get(key)
        if small is None or small is sess:
            small = DecodeSession(self, n_live, sess.mem_len, sess.t_max, "none")
            cache[key] = small
        for name in ("self_k", "self_v"):      # [L, B, H, t_max, dh]: only the t rows written so far
            getattr(small, name)[:, :, :, :t].copy_(getattr(sess, name)[:, :, :, :t].index_select(1, live))
        for name in ("cross_k", "cross_v"):
            getattr(small, name).copy_(getattr(sess, name).index_select(1, live))
        small.lengths.copy_(sess.lengths.index_select(0, live))
        small.finished.zero_()
        # the first step of a launch reads frame t-1 from the frames buffer (include/tts_b200.h); earlier frames stay in
        # the sessions that produced them
        small.frames[:, t - 1].copy_(sess.frames[:, t - 1].index_select(0, live))
        small.memory = sess.memory.index_select(0, live)
        small.input_lengths = sess.input_lengths.index_select(0, live)
        small.dropout = sess.dropout
        small.counters.copy_(torch.tensor([t, n_live], dtype=torch.int32), non_blocking=False)
        small.t = t
        small._st = small.state()
        small._dw = sess._dw
        new_rows = live if rows is None else rows.index_select(0, live)
        return small, new_rows
