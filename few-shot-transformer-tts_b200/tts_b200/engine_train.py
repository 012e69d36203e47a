"""Teacher-forced TRAINING step of the mel path (train.py:171-174: forward, loss, backward) on the sm_100a kernels:
bf16 tcgen05 GEMMs with fp32 accumulation (forward, dgrad, wgrad), flash attention forward / backward, LayerNorm /
BatchNorm / prologue / loss kernels (csrc/gemm_bf16.cu, attn_train.cu, train.cu).  fp32 master weights stay the
module's own nn.Parameters; bf16 operand copies are refreshed in one multi-tensor launch when a parameter changed.

Three forward/backward pairs mirror the three sub-modules of the reference's Tacotron (tacotron.py:126-133) so that the
transformer/ package can wrap each in an autograd.Function: encoder, decoder, postnet.  Every *_fwd returns the API
outputs plus a `saved` dict; every *_bwd takes the output gradients and returns a {state-dict name: gradient} dict
(plus the input gradient where one exists).  No arithmetic happens in torch except O(B x 128) conditioning-vector
glue (speaker / language embeddings) and layout glue on weights (conv weight packing)."""
import torch

from . import _native as N
from . import ops
from . import train_ops as TO
from .engine import sinusoid_table

BF16 = torch.bfloat16


def _split_for(contraction_rows, out_tiles):
    """split-K factor of a weight-gradient GEMM: enough (tile, split) units to fill 148 SMs about twice."""
    if contraction_rows < 4096:
        return 1
    want = max(1, (296 + out_tiles - 1) // out_tiles)
    return int(min(want, 16, max(1, contraction_rows // 2048)))


class TrainEngine:
    def __init__(self, weights, cfg, device):
        self.w, self.cfg, self.device = weights, cfg, torch.device(device)
        N.load()
        self._pe = {}
        self._wb = {}            # name -> bf16 copy
        self._wb_key = None
        self._cast_table = None
        self._conv = {}
        self._step = 0
        self.base_seed = None

    # ---- helpers ------------------------------------------------------------------------------------------------------
    def pe(self, length, channels):
        cur = self._pe.get(channels)
        if cur is None or cur.shape[0] < length:
            n = 1 << (max(length, 256) - 1).bit_length()
            cur = sinusoid_table(n, channels).to(self.device)
            self._pe[channels] = cur
        return cur

    def next_seed(self):
        """One 64-bit seed per forward; every dropout site adds its own stream id (philox.cuh)."""
        if self.base_seed is None:
            self.base_seed = int(torch.initial_seed()) & 0x7fffffffffffffff
        self._step += 1
        return (self.base_seed * 0x9E3779B97F4A7C15 + self._step * 0xD1B54A32D192ED03) & 0x7fffffffffffffff

    def _gemm_names(self):
        return [n for n, t in self.w.items()
                if t.dim() >= 2 and "embed" not in n and "speaker" not in n and "language" not in n and "stop_net" not in n]

    def refresh_bf16(self):
        """bf16 copies of every GEMM weight, refreshed in ONE launch when any master weight changed (optimizer step,
        load_state_dict).  Conv weights are additionally packed [Cout][5][Cin] (forward / wgrad layout) and flipped +
        transposed [Cin][5][Cout] (input-gradient layout)."""
        names = self._gemm_names()
        key = tuple((self.w[n].data_ptr(), self.w[n]._version) for n in names)
        if key == self._wb_key:
            return
        if self._cast_table is None or any(self._wb.get(n) is None or self._wb[n].shape != self.w[n].shape for n in names):
            self._wb = {n: torch.empty(self.w[n].shape, device=self.device, dtype=BF16) for n in names}
            self._cast_table = None
        ptrs = tuple(self.w[n].data_ptr() for n in names)
        if self._cast_table is None or self._cast_ptrs != ptrs:
            tab = TO.MultiTable(self.device)
            tab.build_cast([(self.w[n].detach(), self._wb[n]) for n in names])
            self._cast_table, self._cast_ptrs = tab, ptrs
        TO.multi_cast(self._cast_table)
        self._conv = {}
        for n in names:
            if n.startswith("postnet.conv_layers"):
                wb = self._wb[n]                                   # [Cout][Cin][5]
                self._conv[n] = (wb.permute(0, 2, 1).reshape(wb.shape[0], -1).contiguous(),                   # [Cout][5*Cin]
                                 wb.flip(2).permute(1, 2, 0).reshape(wb.shape[1], -1).contiguous())           # [Cin][5*Cout]
        self._wb_key = key

    def wb(self, name):
        return self._wb[name]

    # ---- one pre-LN sub-block pieces ----------------------------------------------------------------------------------
    def _lin(self, x_bf, wname, **kw):
        return ops.gemm_bf16(x_bf, self._wb[wname], **kw)

    def _wgrad(self, dy_bf, x_bf):
        """dW[N,K] = dY^T X (contraction over rows)."""
        R = dy_bf.shape[0]
        Nn, K = dy_bf.shape[1], x_bf.shape[1]
        tiles = ((Nn + 127) // 128) * ((K + 255) // 256)
        return ops.gemm_bf16(dy_bf, x_bf, a_mn=True, b_mn=True, out_dtype=torch.float32, split_k=_split_for(R, tiles))

    def _dgrad(self, dy_bf, wname, **kw):
        """dX[R,K] = dY[R,N] W[N,K]: the weight is the MN-major B operand."""
        return ops.gemm_bf16(dy_bf, self._wb[wname], b_mn=True, **kw)

    # ==================================================================================================================
    # encoder (tacotron.py:33-44, modules.py:49-69)
    # ==================================================================================================================
    def encoder_fwd(self, ids, input_lengths, spk_ids, lang_vecs, train):
        cfg, w = self.cfg, self.w
        self.refresh_bf16()
        B, S = ids.shape
        E, H = cfg.encoder_hidden, cfg.n_attention_head
        dh = E // H
        p = float(cfg.transformer_dropout_rate) if train else 0.0
        seed = self.next_seed()
        ids = ids.to(self.device).long().contiguous()
        lens = ops._i32(input_lengths.to(self.device))
        pre = "encoder.encoder."
        pe = self.pe(S, E)
        x = TO.embed_fwd(ids, lens, w["encoder.embed.weight"], pe, w[pre + "pe_scale"], B, S, p, seed, 1)
        layers = []
        for l in range(cfg.n_encoder_layer):
            sid = 10 + 4 * l
            h, m1, r1 = TO.ln_fwd(x, w[f"{pre}attn_layer_norms.{l}.weight"], w[f"{pre}attn_layer_norms.{l}.bias"])
            qkv = self._lin(h, f"{pre}self_attentions.{l}.qkv_transform.weight")
            ctx, lse = TO.attn_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, H, S, S, dh, False, lens, p, seed, sid)
            x2 = self._lin(ctx, f"{pre}self_attentions.{l}.output_transform.weight", out_dtype=torch.float32, residual=x,
                           drop_p=p, seed=seed, rng_stream=sid + 1)
            h2, m2, r2 = TO.ln_fwd(x2, w[f"{pre}ffn_layer_norms.{l}.weight"], w[f"{pre}ffn_layer_norms.{l}.bias"])
            a = self._lin(h2, f"{pre}ffn_layers.{l}.input_layer.weight", act=ops.ACT_RELU, drop_p=p, seed=seed, rng_stream=sid + 2)
            x3 = self._lin(a, f"{pre}ffn_layers.{l}.output_layer.weight", out_dtype=torch.float32, residual=x2, drop_p=p,
                           seed=seed, rng_stream=sid + 3)
            layers.append(dict(x=x, h=h, m1=m1, r1=r1, qkv=qkv, ctx=ctx, lse=lse, x2=x2, h2=h2, m2=m2, r2=r2, a=a))
            x = x3
        y, mf, rf = TO.ln_fwd(x, w[pre + "output_layer_norm.weight"], w[pre + "output_layer_norm.bias"])
        width = cfg.decoder_hidden
        mem = torch.empty((B, S, width), device=self.device, dtype=torch.float32)
        mem[:, :, :E].copy_(y.view(B, S, E))   # strided, dtype-converting copy (the torch.cat of tacotron.py:39,43)
        saved = dict(ids=ids, lens=lens, B=B, S=S, p=p, seed=seed, layers=layers, x_out=x, mf=mf, rf=rf, cond={})
        off = E
        if cfg.multi_speaker:
            spk = spk_ids.to(self.device).long().contiguous()
            ops.cond_embed(mem, off, w["encoder.speaker_layer.weight"], w["encoder.speaker_layer.bias"],
                           w1=w["encoder.speaker_embed.weight"], ids=spk)
            saved["cond"]["spk"] = (spk, off)
            off += cfg.speaker_embedding_size
        if cfg.multi_lingual:
            lv = N.f32c(lang_vecs.to(self.device))
            ops.cond_embed(mem, off, w["encoder.language_layer.weight"], w["encoder.language_layer.bias"], vec=lv,
                           w1=w["encoder.language_embed.weight"])
            saved["cond"]["lang"] = (lv, off)
        return mem, saved

    def encoder_bwd(self, saved, d_mem):
        """d_mem: fp32 [B,S,width] -> {name: grad}."""
        cfg, w = self.cfg, self.w
        B, S, p, seed = saved["B"], saved["S"], saved["p"], saved["seed"]
        E, H = cfg.encoder_hidden, cfg.n_attention_head
        dh = E // H
        R = B * S
        lens = saved["lens"]
        pre = "encoder.encoder."
        grads = {}
        d_mem = N.f32c(d_mem)
        d2 = d_mem.view(R, -1)
        # conditioning vectors: O(B x 128) glue on the host side of the API (tacotron.py:21-31,36-43)
        for kind, (src, off) in saved["cond"].items():
            size = cfg.speaker_embedding_size if kind == "spk" else cfg.language_embedding_size
            g = d_mem[:, :, off:off + size].sum(1)                                           # [B, size]
            lw, lb = ("encoder.speaker_layer.weight", "encoder.speaker_layer.bias") if kind == "spk" else \
                     ("encoder.language_layer.weight", "encoder.language_layer.bias")
            if kind == "spk":
                hvec = w["encoder.speaker_embed.weight"][src]
            else:
                hvec = src @ w["encoder.language_embed.weight"].t()
            zz = hvec @ w[lw].t() + w[lb]
            dz = g / (1 + zz.abs()) ** 2                                                     # softsign'
            grads[lw], grads[lb] = dz.t() @ hvec, dz.sum(0)
            dh_ = dz @ w[lw]
            if kind == "spk":
                ge = torch.zeros_like(w["encoder.speaker_embed.weight"])
                ge.index_add_(0, src, dh_)
                grads["encoder.speaker_embed.weight"] = ge
            else:
                grads["encoder.language_embed.weight"] = dh_.t() @ src
        dy = TO.dropout_cast(d2[:, :E])                                                      # bf16 [R,E] from the strided view
        # every LayerNorm backward also emits dyb = dropout_backward(dx) in bf16 for the Linear that precedes it (fused cast)
        dx, grads[pre + "output_layer_norm.weight"], grads[pre + "output_layer_norm.bias"], dyb = TO.ln_bwd(
            dy, saved["x_out"], saved["mf"], saved["rf"], w[pre + "output_layer_norm.weight"],
            cast_drop=(p, seed, 10 + 4 * (cfg.n_encoder_layer - 1) + 3))
        for l in reversed(range(cfg.n_encoder_layer)):
            sv = saved["layers"][l]
            sid = 10 + 4 * l
            # FFN
            grads[f"{pre}ffn_layers.{l}.output_layer.weight"] = self._wgrad(dyb, sv["a"])
            da = self._dgrad(dyb, f"{pre}ffn_layers.{l}.output_layer.weight", gate=sv["a"], gate_scale=1.0 / (1.0 - p))
            grads[f"{pre}ffn_layers.{l}.input_layer.weight"] = self._wgrad(da, sv["h2"])
            dh2 = self._dgrad(da, f"{pre}ffn_layers.{l}.input_layer.weight")
            dx, grads[f"{pre}ffn_layer_norms.{l}.weight"], grads[f"{pre}ffn_layer_norms.{l}.bias"], dyb = TO.ln_bwd(
                dh2, sv["x2"], sv["m2"], sv["r2"], w[f"{pre}ffn_layer_norms.{l}.weight"], dres=dx, cast_drop=(p, seed, sid + 1))
            # self-attention
            grads[f"{pre}self_attentions.{l}.output_transform.weight"] = self._wgrad(dyb, sv["ctx"])
            dctx = self._dgrad(dyb, f"{pre}self_attentions.{l}.output_transform.weight")
            qkv = sv["qkv"]
            dqkv = torch.empty_like(qkv)
            TO.attn_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], sv["ctx"], sv["lse"], dctx, dqkv[:, :E], dqkv[:, E:2 * E],
                        dqkv[:, 2 * E:], B, H, S, S, dh, False, lens, p, seed, sid)
            grads[f"{pre}self_attentions.{l}.qkv_transform.weight"] = self._wgrad(dqkv, sv["h"])
            dh1 = self._dgrad(dqkv, f"{pre}self_attentions.{l}.qkv_transform.weight")
            res = TO.ln_bwd(dh1, sv["x"], sv["m1"], sv["r1"], w[f"{pre}attn_layer_norms.{l}.weight"], dres=dx,
                            cast_drop=(p, seed, 10 + 4 * (l - 1) + 3) if l > 0 else None)
            dx, grads[f"{pre}attn_layer_norms.{l}.weight"], grads[f"{pre}attn_layer_norms.{l}.bias"] = res[:3]
            dyb = res[3] if l > 0 else None
        ge, gs = TO.embed_bwd(dx, saved["ids"], lens, self.pe(S, E), w["encoder.embed.weight"].shape[0], B, S, p, seed, 1)
        grads["encoder.embed.weight"], grads[pre + "pe_scale"] = ge, gs
        return grads

    # ==================================================================================================================
    # decoder (tacotron.py:107-116, modules.py:108-145)
    # ==================================================================================================================
    def decoder_fwd(self, memory, input_lengths, targets, target_lengths, train, leave_one=False):
        cfg, w = self.cfg, self.w
        self.refresh_bf16()
        B, T, M = targets.shape
        S, D, H = memory.shape[1], cfg.decoder_hidden, cfg.n_attention_head
        dh = D // H
        R = B * T
        p = float(cfg.transformer_dropout_rate) if train else 0.0
        pp = float(cfg.decoder_dropout_rate) if train else 0.0
        seed = self.next_seed()
        in_len, tg_len = ops._i32(input_lengths.to(self.device)), ops._i32(target_lengths.to(self.device))
        mem_bf = TO.dropout_cast(N.f32c(memory).view(B * S, D))
        tgt_bf = TO.dropout_cast(N.f32c(targets).view(R, M))
        q = "decoder.prenet."
        h0 = self._lin(tgt_bf, q + "dense0.weight", bias=w[q + "dense0.bias"], act=ops.ACT_RELU, drop_p=pp, seed=seed, rng_stream=3)
        h1 = self._lin(h0, q + "dense1.weight", bias=w[q + "dense1.bias"], act=ops.ACT_RELU, drop_p=pp, seed=seed, rng_stream=4)
        pre_out = self._lin(h1, q + "dense_final.weight", out_dtype=torch.float32)
        # leave_one zeroes the last prenet row, which the shift-right drops: a no-op (tacotron.py:109-110)
        pre = "decoder.decoder."
        pe = self.pe(T, D)
        x = TO.shift_pe_fwd(pre_out, tg_len, pe, w[pre + "pe_scale"], B, T, p, seed, 2)
        del pre_out
        layers = []
        for l in range(cfg.n_decoder_layer):
            sid = 100 + 8 * l
            h, m1, r1 = TO.ln_fwd(x, w[f"{pre}attn_layer_norms.{l}.weight"], w[f"{pre}attn_layer_norms.{l}.bias"])
            qkv = self._lin(h, f"{pre}self_attentions.{l}.qkv_transform.weight")
            ctx, lse = TO.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, T, T, dh, True, None, p, seed, sid)
            x2 = self._lin(ctx, f"{pre}self_attentions.{l}.output_transform.weight", out_dtype=torch.float32, residual=x,
                           drop_p=p, seed=seed, rng_stream=sid + 1)
            h2, m2, r2 = TO.ln_fwd(x2, w[f"{pre}encdec_layer_norms.{l}.weight"], w[f"{pre}encdec_layer_norms.{l}.bias"])
            q2 = self._lin(h2, f"{pre}encdec_attentions.{l}.q_transform.weight")
            kv = self._lin(mem_bf, f"{pre}encdec_attentions.{l}.kv_transform.weight")
            ctx2, lse2 = TO.attn_fwd(q2, kv[:, :D], kv[:, D:], B, H, T, S, dh, False, in_len, p, seed, sid + 2)
            x3 = self._lin(ctx2, f"{pre}encdec_attentions.{l}.output_transform.weight", out_dtype=torch.float32, residual=x2,
                           drop_p=p, seed=seed, rng_stream=sid + 3)
            h3, m3, r3 = TO.ln_fwd(x3, w[f"{pre}ffn_layer_norms.{l}.weight"], w[f"{pre}ffn_layer_norms.{l}.bias"])
            a = self._lin(h3, f"{pre}ffn_layers.{l}.input_layer.weight", act=ops.ACT_RELU, drop_p=p, seed=seed, rng_stream=sid + 4)
            x4 = self._lin(a, f"{pre}ffn_layers.{l}.output_layer.weight", out_dtype=torch.float32, residual=x3, drop_p=p,
                           seed=seed, rng_stream=sid + 5)
            layers.append(dict(x=x, h=h, m1=m1, r1=r1, qkv=qkv, ctx=ctx, lse=lse, x2=x2, h2=h2, m2=m2, r2=r2, q2=q2, kv=kv,
                               ctx2=ctx2, lse2=lse2, x3=x3, h3=h3, m3=m3, r3=r3, a=a))
            x = x4
        o, mo, ro = TO.ln_fwd(x, w[pre + "output_layer_norm.weight"], w[pre + "output_layer_norm.bias"], row_len=tg_len,
                              rows_per_batch=T)
        mels = self._lin(o, "decoder.mel_net.weight", out_dtype=torch.float32, row_len=tg_len, rows_per_batch=T)
        stop = TO.rowdot(o, w["decoder.stop_net.weight"].view(-1), w["decoder.stop_net.bias"], tg_len, T)
        saved = dict(B=B, T=T, S=S, p=p, pp=pp, seed=seed, in_len=in_len, tg_len=tg_len, mem_bf=mem_bf, tgt_bf=tgt_bf, h0=h0, h1=h1,
                     layers=layers, x_out=x, o=o, mo=mo, ro=ro)
        return mels.view(B, T, M), stop.view(B, T), saved

    def decoder_bwd(self, saved, d_mels, d_stop, need_dmem=True):
        cfg, w = self.cfg, self.w
        B, T, S, p, pp, seed = saved["B"], saved["T"], saved["S"], saved["p"], saved["pp"], saved["seed"]
        D, H, M = cfg.decoder_hidden, cfg.n_attention_head, cfg.num_mels
        dh = D // H
        R = B * T
        in_len, tg_len = saved["in_len"], saved["tg_len"]
        pre = "decoder.decoder."
        grads = {}
        o = saved["o"]
        do = None
        if d_mels is not None:
            dmel = TO.dropout_cast(N.f32c(d_mels).view(R, M), row_len=tg_len, rows_per_batch=T)   # impute' (tacotron.py:113)
            grads["decoder.mel_net.weight"] = self._wgrad(dmel, o)
            do = self._dgrad(dmel, "decoder.mel_net.weight")
        if d_stop is not None:   # the stop head reads DETACHED features (tacotron.py:114): parameters only
            live = (torch.arange(T, device=self.device)[None, :] < tg_len[:, None]).reshape(-1)
            ds = (N.f32c(d_stop).view(R) * live).contiguous()
            grads["decoder.stop_net.weight"] = TO.colsum(o, ds).view(1, D)
            grads["decoder.stop_net.bias"] = TO.sum_f32(ds).view(1)
        if do is None:
            do = torch.zeros((R, D), device=self.device, dtype=BF16)
        dx, grads[pre + "output_layer_norm.weight"], grads[pre + "output_layer_norm.bias"], dyb = TO.ln_bwd(
            do, saved["x_out"], saved["mo"], saved["ro"], w[pre + "output_layer_norm.weight"], row_len=tg_len, rows_per_batch=T,
            cast_drop=(p, seed, 100 + 8 * (cfg.n_decoder_layer - 1) + 5))
        mem_bf = saved["mem_bf"]
        dmem = None
        for l in reversed(range(cfg.n_decoder_layer)):
            sv = saved["layers"][l]
            sid = 100 + 8 * l
            # FFN (dyb = dropout_backward(dx), bf16, came out of the LayerNorm backward that produced dx)
            grads[f"{pre}ffn_layers.{l}.output_layer.weight"] = self._wgrad(dyb, sv["a"])
            da = self._dgrad(dyb, f"{pre}ffn_layers.{l}.output_layer.weight", gate=sv["a"], gate_scale=1.0 / (1.0 - p))
            grads[f"{pre}ffn_layers.{l}.input_layer.weight"] = self._wgrad(da, sv["h3"])
            dh3 = self._dgrad(da, f"{pre}ffn_layers.{l}.input_layer.weight")
            del da
            dx, grads[f"{pre}ffn_layer_norms.{l}.weight"], grads[f"{pre}ffn_layer_norms.{l}.bias"], dyb = TO.ln_bwd(
                dh3, sv["x3"], sv["m3"], sv["r3"], w[f"{pre}ffn_layer_norms.{l}.weight"], dres=dx, cast_drop=(p, seed, sid + 3))
            # cross-attention
            grads[f"{pre}encdec_attentions.{l}.output_transform.weight"] = self._wgrad(dyb, sv["ctx2"])
            dctx2 = self._dgrad(dyb, f"{pre}encdec_attentions.{l}.output_transform.weight")
            kv = sv["kv"]
            dq2, dkv = torch.empty_like(sv["q2"]), torch.empty_like(kv)
            TO.attn_bwd(sv["q2"], kv[:, :D], kv[:, D:], sv["ctx2"], sv["lse2"], dctx2, dq2, dkv[:, :D], dkv[:, D:], B, H, T, S, dh,
                        False, in_len, p, seed, sid + 2)
            grads[f"{pre}encdec_attentions.{l}.q_transform.weight"] = self._wgrad(dq2, sv["h2"])
            dh2 = self._dgrad(dq2, f"{pre}encdec_attentions.{l}.q_transform.weight")
            grads[f"{pre}encdec_attentions.{l}.kv_transform.weight"] = self._wgrad(dkv, mem_bf)
            if need_dmem:
                dmem = self._dgrad(dkv, f"{pre}encdec_attentions.{l}.kv_transform.weight", out_dtype=torch.float32, residual=dmem)
            dx, grads[f"{pre}encdec_layer_norms.{l}.weight"], grads[f"{pre}encdec_layer_norms.{l}.bias"], dyb = TO.ln_bwd(
                dh2, sv["x2"], sv["m2"], sv["r2"], w[f"{pre}encdec_layer_norms.{l}.weight"], dres=dx, cast_drop=(p, seed, sid + 1))
            # self-attention
            grads[f"{pre}self_attentions.{l}.output_transform.weight"] = self._wgrad(dyb, sv["ctx"])
            dctx = self._dgrad(dyb, f"{pre}self_attentions.{l}.output_transform.weight")
            qkv = sv["qkv"]
            dqkv = torch.empty_like(qkv)
            TO.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], sv["ctx"], sv["lse"], dctx, dqkv[:, :D], dqkv[:, D:2 * D],
                        dqkv[:, 2 * D:], B, H, T, T, dh, True, None, p, seed, sid)
            grads[f"{pre}self_attentions.{l}.qkv_transform.weight"] = self._wgrad(dqkv, sv["h"])
            dh1 = self._dgrad(dqkv, f"{pre}self_attentions.{l}.qkv_transform.weight")
            del dqkv
            res = TO.ln_bwd(dh1, sv["x"], sv["m1"], sv["r1"], w[f"{pre}attn_layer_norms.{l}.weight"], dres=dx,
                            cast_drop=(p, seed, 100 + 8 * (l - 1) + 5) if l > 0 else None)
            dx, grads[f"{pre}attn_layer_norms.{l}.weight"], grads[f"{pre}attn_layer_norms.{l}.bias"] = res[:3]
            dyb = res[3] if l > 0 else None
            saved["layers"][l] = None   # free the layer's activations as soon as its backward is done
        dpre, grads[pre + "pe_scale"] = TO.shift_pe_bwd(dx, tg_len, self.pe(T, D), B, T, p, seed, 2)
        q = "decoder.prenet."
        gs = 1.0 / (1.0 - pp)
        grads[q + "dense_final.weight"] = self._wgrad(dpre, saved["h1"])
        dh1p = self._dgrad(dpre, q + "dense_final.weight", gate=saved["h1"], gate_scale=gs)
        grads[q + "dense1.bias"] = TO.colsum(dh1p)
        grads[q + "dense1.weight"] = self._wgrad(dh1p, saved["h0"])
        dh0p = self._dgrad(dh1p, q + "dense1.weight", gate=saved["h0"], gate_scale=gs)
        grads[q + "dense0.bias"] = TO.colsum(dh0p)
        grads[q + "dense0.weight"] = self._wgrad(dh0p, saved["tgt_bf"])
        d_memory = dmem.view(B, S, D) if dmem is not None else None
        return grads, d_memory

    # ==================================================================================================================
    # postnet in train() mode (tacotron.py:81-90): batch-statistics BatchNorm over all B x T positions, dropout 0.5
    # ==================================================================================================================
    def postnet_fwd(self, mels, lengths, train, add_input=False):
        cfg, w = self.cfg, self.w
        self.refresh_bf16()
        B, T, M = mels.shape
        pp = float(cfg.decoder_dropout_rate) if train else 0.0
        seed = self.next_seed()
        lens = ops._i32(lengths.to(self.device))
        mels2 = N.f32c(mels).view(B * T, M)
        xpad = TO.pad_cast(mels2, lens, B, T)
        n = cfg.n_postnet_layer
        rows = B * (T + 4) - 4
        layers = []
        out = None
        for i in range(n):
            last = i == n - 1
            wp, _ = self._conv[f"postnet.conv_layers.{i}.weight"]
            cin, cout = xpad.shape[-1], wp.shape[0]
            z = ops.gemm_bf16(xpad.view(-1, cin), wp, m=rows, k=cin, taps=5, out_dtype=torch.float32, rows_per_batch=T + 4,
                              valid_rows=T, out_rows_per_batch=T, out_rows=B * T, a_rows=B * (T + 4))
            b = f"postnet.batchnorm_layers.{i}."
            if last:
                out = torch.empty((B * T, cout), device=self.device, dtype=torch.float32)
                mean, invstd = TO.bn_fwd(z, w[b + "weight"], w[b + "bias"], w[b + "running_mean"], w[b + "running_var"],
                                         w[b + "num_batches_tracked"], False, pp, seed, 200 + i, None, B, T, out_f32=out,
                                         residual=mels2 if add_input else None)
                nxt = None
            else:
                nxt = torch.empty((B, T + 4, cout), device=self.device, dtype=BF16)
                TO.pad_cast(None, None, B, T, out=nxt, only_pads=True)
                mean, invstd = TO.bn_fwd(z, w[b + "weight"], w[b + "bias"], w[b + "running_mean"], w[b + "running_var"],
                                         w[b + "num_batches_tracked"], True, pp, seed, 200 + i, lens, B, T, out_pad=nxt)
            layers.append(dict(xpad=xpad, z=z, mean=mean, invstd=invstd))
            xpad = nxt
        saved = dict(B=B, T=T, pp=pp, seed=seed, lens=lens, layers=layers, add_input=add_input)
        return out.view(B, T, M), saved

    def postnet_bwd(self, saved, d_out):
        """d_out: gradient of the returned tensor [B,T,M] -> (grads, d_mels [B,T,M])."""
        cfg, w = self.cfg, self.w
        B, T, pp, seed, lens = saved["B"], saved["T"], saved["pp"], saved["seed"], saved["lens"]
        n = cfg.n_postnet_layer
        rows = B * (T + 4) - 4
        grads = {}
        dout = N.f32c(d_out).view(B * T, -1)
        d_in = dout if saved["add_input"] else None
        for i in reversed(range(n)):
            last = i == n - 1
            sv = saved["layers"][i]
            z, xpad = sv["z"], sv["xpad"]
            cout, cin = z.shape[1], xpad.shape[-1]
            b = f"postnet.batchnorm_layers.{i}."
            dz = torch.empty((B, T + 4, cout), device=self.device, dtype=BF16)
            TO.pad_cast(None, None, B, T, out=dz, only_pads=True)
            grads[b + "weight"], grads[b + "bias"] = TO.bn_bwd(z, dout, w[b + "weight"], w[b + "bias"], sv["mean"], sv["invstd"],
                                                               not last, pp, seed, 200 + i, lens, not last, B, T, dz)
            # weight gradient: dW[co][tap][ci] = sum_r dz_pad[r + 2][co] * x_pad[r + tap][ci] (pad rows are zero)
            dz2, x2 = dz.view(-1, cout), xpad.view(-1, cin)
            dwp = torch.zeros((cout, 5 * cin), device=self.device, dtype=torch.float32)
            tiles = ((cout + 127) // 128) * ((cin + 255) // 256)
            split = max(2, _split_for(rows, tiles))
            for tap in range(5):
                ops.gemm_bf16(dz2[2:2 + rows], x2[tap:tap + rows], a_mn=True, b_mn=True, out=dwp[:, tap * cin:(tap + 1) * cin],
                              split_k=split)
            grads[f"postnet.conv_layers.{i}.weight"] = dwp.view(cout, 5, cin).permute(0, 2, 1).contiguous()
            # input gradient: the same 5-tap GEMM over dz_pad with the flipped, transposed weight
            _, wd = self._conv[f"postnet.conv_layers.{i}.weight"]
            # (layer 0 reads impute(mels): rows at or beyond the length get no gradient from the convolution)
            dout = ops.gemm_bf16(dz2, wd, m=rows, k=cout, taps=5, out_dtype=torch.float32, rows_per_batch=T + 4, valid_rows=T,
                                 out_rows_per_batch=T, out_rows=B * T, a_rows=B * (T + 4), row_len=lens if i == 0 else None)
            saved["layers"][i] = None
        if d_in is not None:   # mel_aft = mel_bef + postnet(mel_bef): the identity branch (an [R, 80] add)
            dout = dout + d_in
        return grads, dout.view(B, T, -1)
