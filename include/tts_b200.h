/*
 * tts_b200.h — C ABI of the B200-native Transformer-TTS mel path.
 *
 * The reference (mutiann/few-shot-transformer-tts) has no FFI: its boundary for this path is
 * the Python class API of the `transformer/` package.  This header is the
 * C-ABI shared-library boundary underneath our drop-in `transformer/` package: plain
 * pointers and sizes, no torch types.  Every entry point names the reference code it
 * replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - all matrices are row-major fp32, activations are [rows][channels] (batch-first,
 *     channels-last, transformer/tacotron.py conventions), weights are torch Linear layout
 *     [out][in];
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - return value 0 = success, non-zero = error; tts_last_error() gives the message.
 *     Nothing in this library aborts or allocates device memory: scratch is caller-owned.
 */
#ifndef TTS_B200_H
#define TTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTS_MAX_LAYERS 16
#define TTS_ABI_VERSION 10

/* ---- library / diagnostics -------------------------------------------------------------- */
int tts_abi_version(void);
const char* tts_last_error(void);
/* number of kernels this library launched since the last reset (bench.py's gpu_launches) */
int64_t tts_launch_count(void);
void tts_launch_count_reset(void);

/* ---- dense building blocks (prefill / teacher-forced path) ------------------------------- */

/* Epilogue of tts_gemm_nt.  v = alpha*acc; v = v*scale[n]+shift[n]; v += bias[n]; act;
 * v += residual[m][n]; rows at or beyond row_len[batch] are zeroed; then stored.
 * Rows are grouped in batches of `rows_per_batch` (0 = one batch): row m -> (b, r);
 * rows with r >= valid_rows are not stored; the output row is b*out_rows_per_batch + r +
 * out_row_offset.  This is how the k=5 Conv1d of the Postnet runs as a GEMM over a
 * zero-padded [B][T+4][C] activation buffer with overlapping rows (lda = C, K = 5C). */
typedef struct TtsGemmEpilogue {
  float alpha;                 /* 1.0f for none */
  const float* scale;          /* [N] or NULL */
  const float* shift;          /* [N] or NULL */
  const float* bias;           /* [N] or NULL */
  int32_t act;                 /* 0 none, 1 relu, 2 tanh */
  const float* residual;       /* [M_out][ldr] or NULL (indexed by OUTPUT row) */
  int32_t ldr;
  const int32_t* row_len;      /* [n_batches] or NULL */
  int32_t rows_per_batch;      /* 0 = M */
  int32_t valid_rows;          /* 0 = rows_per_batch */
  int32_t out_rows_per_batch;  /* 0 = rows_per_batch */
  int32_t out_row_offset;
  /* head-split store (cross K/V precompute, attention.py:66-68 + split_heads :6-15):
   * if head_dim > 0 the N axis is [2][H][head_dim] and element (m=(b,s), n=(w,h,d)) goes to
   * (w ? out_v : c)[((b*H + h)*head_rows + s)*head_dim + d]. */
  int32_t head_dim;
  int32_t n_heads;
  int32_t head_rows;
  float* out_v;
} TtsGemmEpilogue;

/* C[M,N] = epilogue(A[M,K] * W[N,K]^T).  Replaces every nn.Linear / nn.Conv1d call of the
 * teacher-forced path: transformer/attention.py:43-47,63-68,119; modules.py:11-19;
 * tacotron.py:50-52,56-64,78,85,112,114.  K % 4 == 0, lda % 4 == 0, ldw % 4 == 0. */
int tts_gemm_nt(const float* A, int32_t lda, const float* W, int32_t ldw, float* C, int32_t ldc,
                int32_t M, int32_t N, int32_t K, const TtsGemmEpilogue* epi, void* stream);

/* Route large tts_gemm_nt problems with K % 32 == 0 through the tcgen05 (TMEM-accumulating) 3xTF32 kernel (1) or keep
 * the packed-FFMA2 kernels (0); on < 0 only queries.  Returns the previous setting.  Initial value: TTS_GEMM_TC. */
int tts_gemm_use_tensor_cores(int on);

/* y[r,:] = LayerNorm(x[r,:]) * gamma + beta, eps = 1e-6 (modules.py:36,43,47,88,95,102,106).
 * row_len/rows_per_batch (optional) zero rows at or beyond the length (common.py:51 impute). */
int tts_layernorm(const float* x, float* y, const float* gamma, const float* beta, int32_t rows,
                  int32_t channels, float eps, const int32_t* row_len, int32_t rows_per_batch,
                  void* stream);

/* Encoder prologue: out[b,s,:] = embed[ids[b,s],:] * (s < len[b]) + pe[s,:] * pe_scale
 * (tacotron.py:34, modules.py:49-55, common.py:4-29).  pe is a [>=S][C] device table.
 * ids == NULL: `embed` is the already embedded input [B*S][C] (vocab = B*S). */
int tts_embed_pe(const int64_t* ids, const int32_t* lengths, const float* embed, const float* pe,
                 const float* pe_scale, float* out, int32_t batch, int32_t seq, int32_t channels,
                 int32_t vocab, void* stream);

/* Decoder prologue (teacher forced): out[b,t,:] = (t==0 ? 0 : pre[b,t-1,:] * (t-1 < len[b])
 * [* (t-1 != T-1 if leave_one)]) + pe[t,:] * pe_scale   (modules.py:114-118, tacotron.py:109-110) */
int tts_shift_pe(const float* pre, const int32_t* lengths, const float* pe, const float* pe_scale,
                 float* out, int32_t batch, int32_t frames, int32_t channels, void* stream);

/* out[b, pad + t, :] = t < len[b] ? x[b, t, :] : 0 for t < T; the `pad` rows before and after
 * each sequence are zeroed.  out is [B][T + 2*pad][C].  This is the `impute` + zero padding in
 * front of the first Postnet convolution (tacotron.py:83-84, Conv1d padding=2). */
int tts_pad_rows(const float* x, const int32_t* lengths, float* out, int32_t batch, int32_t frames,
                 int32_t channels, int32_t pad, void* stream);

/* Speaker / language conditioning written into the tail of the encoder memory:
 * mem[b,s,off:off+E] = softsign(W2 * h + b2) broadcast over s (tacotron.py:21-31,36-43), with
 * h = w1[ids[b], :] (speaker: nn.Embedding table [n][E], ids != NULL) or
 * h = w1 * vec[b]   (language: bias-free Linear [E][vec_dim] on the one-hot vector, ids == NULL). */
int tts_cond_embed(const float* vec, int32_t vec_dim, const int64_t* ids, const float* w1, const float* w2,
                   const float* b2, int32_t emb, float* mem, int32_t batch, int32_t seq,
                   int32_t mem_width, int32_t col_offset, void* stream);

/* Multi-head scaled dot-product attention over full sequences (attention.py:72-122 minus the
 * projections).  q/k/v are addressed as ptr[(b*rows + r)*ld + h*head_dim + d].
 * mask: causal != 0 -> key j allowed iff j <= i (modules.py:112); key_len != NULL -> key j
 * allowed iff j < key_len[b] (modules.py:50-52,109-111).  Masked logits are -1e20 like
 * common.py:32.  ctx is [B][Tq][H*head_dim]; align (optional) is [B][H][Tq][Tk] (the
 * reference returns the transposed view, attention.py:88).  head_dim in {32, 64, 96}. */
int tts_attention(const float* q, int32_t ldq, const float* k, int32_t ldk, const float* v,
                  int32_t ldv, float* ctx, float* align, int32_t batch, int32_t n_heads,
                  int32_t tq, int32_t tk, int32_t head_dim, float q_scale, int32_t causal,
                  const int32_t* key_len, void* stream);

/* ---- teacher-forced TRAINING path: bf16 tensor-core GEMM ---------------------------------------
 * C[M,N] = epilogue( sum over taps t, k < K of A[m + t][k] * B[n][t*K + k] ), bf16 operands, fp32 accumulation in
 * tensor memory (tcgen05.mma kind::f16, TMA-fed).  Replaces, for the training step of train.py:171-174, the forward,
 * input-gradient and weight-gradient products of every nn.Linear / nn.Conv1d (transformer/attention.py:43-47,63-68,
 * 119; modules.py:11-19; tacotron.py:50-52,56-64,78,85,112,114) that autograd runs as fp32 cuBLAS / cuDNN calls.
 *   operands: bf16 bit patterns (uint16_t).  K-major = row-major [rows][K] (ld = row stride in elements);
 *             MN-major = row-major [K][rows].  dgrad passes the weight [N][K] as the MN-major B of dX = dY W; wgrad
 *             passes dY [R][N] and X [R][K] as MN-major A and B of dW = dY^T X (contraction over the R rows).
 *   epilogue (in this order): alpha; + bias[n]; ReLU (act 1); * gate_scale where gate[m][n] > 0 else 0 (backward of
 *             ReLU+dropout through the saved bf16 forward output); Philox dropout(drop_p) keyed by (seed, rng_stream,
 *             output element); + residual[m][n] (fp32); rows at or beyond row_len[b] zeroed; row remapping exactly
 *             as TtsGemmEpilogue (the k5 Conv1d over a zero-padded buffer: taps = 5, K = C_in).
 *   split_k > 1: the contraction is cut into split_k slices reduced with fp32 atomics into a ZERO-INITIALISED fp32 C
 *             (weight gradients: tiny outputs, contraction over all B x T rows).
 * Base pointers and row strides of A and B must be 16-byte aligned (ld % 8 == 0). */
typedef struct TtsGemmBf16 {
  const uint16_t* A; int64_t lda; int32_t a_mn_major; int64_t a_rows;   /* a_rows: rows of the K-major A buffer (0 = M + taps - 1) */
  const uint16_t* B; int64_t ldb; int32_t b_mn_major;
  void* C; int64_t ldc; int32_t out_bf16;                              /* C: bf16 (out_bf16 != 0) or fp32 */
  int32_t M, N, K, taps, split_k;
  const float* bias; int32_t act; float alpha;                          /* alpha 0 = 1 */
  const float* residual; int64_t ldr;
  float drop_p; uint64_t seed; uint32_t rng_stream;
  const uint16_t* gate; int64_t ldg; float gate_scale;
  const int32_t* row_len; int32_t rows_per_batch, valid_rows, out_rows_per_batch, out_row_offset;
} TtsGemmBf16;
int tts_gemm_bf16(const TtsGemmBf16* g, void* stream);
/* 0 = no barrier timeout has been recorded by any tts_gemm_bf16 kernel so far (synchronous device read) */
int tts_gemm_bf16_status(void);

/* Flash-style attention of the training path on bf16 tensor cores (mma.sync m16n8k16), masks from indices / lengths,
 * Philox dropout on the weights regenerated in backward; nothing of size [B,H,Tq,Tk] is written.  Replaces
 * transformer/attention.py:72-122 (minus the projections) and its autograd backward.  q/k/v/out/dq/dk/dv are bf16,
 * addressed as ptr[(b*rows + r)*ld + h*head_dim + d]; lse / delta are fp32 [B][H][Tq] (lse in the log2 domain).
 * causal: key j allowed iff j <= i (modules.py:112); key_len: key j allowed iff j < key_len[b] (modules.py:50-52). */
typedef struct TtsAttnTrain {
  const uint16_t *q, *k, *v; int64_t ldq, ldk, ldv;
  uint16_t* out; int64_t ldo;
  float* lse;
  int32_t batch, n_heads, tq, tk, head_dim, causal;
  const int32_t* key_len;
  float drop_p; uint64_t seed; uint32_t rng_stream;
  /* backward only */
  const uint16_t* d_out; int64_t lddo;
  float* delta;
  uint16_t *dq, *dk, *dv; int64_t lddq, lddk, lddv;
  float* dq_acc;   /* optional scratch, fp32 [B*tq][n_heads*head_dim]: when given, backward runs the single-pass kernel (dK, dV and
                      dQ from one recomputation of S / dP; dQ summed with fp32 atomics); NULL = the deterministic two-kernel path */
  uint32_t* keep_mask;   /* optional cache of the dropout keep bits, [batch*n_heads][tts_attn_keep_words(tk)][tq] words (1 bit per
                      attention weight): written by the forward call (drop_p > 0), read by the single-pass backward instead of
                      regenerating the Philox bits; NULL = regenerate (same bits either way) */
} TtsAttnTrain;
int32_t tts_attn_keep_words(int32_t tk);
int tts_attn_tc_trace(long long* out);   /* diagnostics: 3 x 32 x 8 SM-clock stamps of one CTA (TTS_ATTN_TC_TRACE=1) */
int tts_attn_tc_status(void);   /* 1 after a barrier wait of the tcgen05 backward kernel timed out; reading resets it */
int tts_attn_train_fwd(const TtsAttnTrain* t, void* stream);
int tts_attn_train_bwd(const TtsAttnTrain* t, void* stream);

/* LayerNorm of the training path (modules.py:36,43,47,88,95,102,106): y (bf16, feeds the next GEMM) = LN(x) gamma + beta,
 * rows at or beyond row_len zeroed (modules.py:144); mean / rstd saved for backward. */
int tts_ln_fwd_train(const float* x, uint16_t* y, int64_t ldy, const float* gamma, const float* beta, float* mean,
                     float* rstd, int32_t rows, int32_t channels, float eps, const int32_t* row_len,
                     int32_t rows_per_batch, void* stream);
/* dx = LN'(dy) (+ dres: the gradient arriving over the residual connection), dgamma, dbeta.  scratch: tts_ln_bwd_scratch_floats.
 * dyb (optional, bf16): also dyb = keep(seed, rng_stream, element) ? dx / (1 - drop_p) : 0, the tts_dropout_cast of dx. */
int tts_ln_bwd_train(const uint16_t* dy, int64_t lddy, const float* x, const float* mean, const float* rstd,
                     const float* gamma, const float* dres, float* dx, float* dgamma, float* dbeta, float* scratch,
                     int32_t rows, int32_t channels, const int32_t* row_len, int32_t rows_per_batch, uint16_t* dyb,
                     int64_t lddyb, float drop_p, uint64_t seed, uint32_t rng_stream, void* stream);
size_t tts_ln_bwd_scratch_floats(int32_t channels);
/* dst (bf16) = keep(seed, stream, element) ? src / (1 - p) : 0: the backward of an epilogue dropout, or (p = 0) a cast. */
int tts_dropout_cast(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int32_t channels,
                     float drop_p, uint64_t seed, uint32_t rng_stream, const int32_t* row_len, int32_t rows_per_batch,
                     void* stream);
/* fp32 -> bf16 copies of many tensors in one launch; table_dev: device array of {const float* src; uint16_t* dst;
 * int64 n; int64 first_chunk} with chunks of tts_multi_chunk_elems() elements. */
int tts_multi_cast_bf16(const void* table_dev, int32_t n_entries, int64_t n_chunks, void* stream);
int32_t tts_multi_chunk_elems(void);
/* Encoder / decoder prologues with dropout and their backward (tacotron.py:34; modules.py:49-55,114-120). */
int tts_embed_train_fwd(const int64_t* ids, const int32_t* lengths, const float* embed, const float* pe,
                        const float* pe_scale, float* out, int32_t batch, int32_t seq, int32_t channels, int32_t vocab,
                        float drop_p, uint64_t seed, uint32_t rng_stream, void* stream);
int tts_embed_train_bwd(const float* dx, const int64_t* ids, const int32_t* lengths, const float* pe, float* d_embed,
                        float* d_pe_scale, int32_t batch, int32_t seq, int32_t channels, int32_t vocab, float drop_p,
                        uint64_t seed, uint32_t rng_stream, void* stream);
int tts_shift_pe_train_fwd(const float* pre, const int32_t* lengths, const float* pe, const float* pe_scale, float* out,
                           int32_t batch, int32_t frames, int32_t channels, float drop_p, uint64_t seed,
                           uint32_t rng_stream, void* stream);
int tts_shift_pe_train_bwd(const float* dx, const int32_t* lengths, const float* pe, uint16_t* dpre, float* d_pe_scale,
                           int32_t batch, int32_t frames, int32_t channels, float drop_p, uint64_t seed,
                           uint32_t rng_stream, void* stream);
/* out[c] += sum_r row_weight[r] * x[r][c] (bias gradients; the stop head's weight gradient); out must be initialised. */
int tts_colsum_bf16(const uint16_t* x, int64_t ldx, const float* row_weight, float* out, int64_t rows, int32_t channels,
                    void* stream);
int tts_sum_f32(const float* x, int64_t n, float* out, void* stream);
/* stop head (tacotron.py:114-115): out[r] = (x[r] . w + bias) masked by row_len. */
int tts_rowdot_bf16(const uint16_t* x, int64_t ldx, const float* w, const float* bias, const int32_t* row_len,
                    int32_t rows_per_batch, float* out, int64_t rows, int32_t k, void* stream);
/* Postnet BatchNorm1d in train() mode: batch statistics over all B x T positions, running-stat update, tanh, dropout,
 * length mask into the next layer's zero-padded bf16 input (tacotron.py:83-89). */
size_t tts_bn_scratch_floats(int32_t channels);
int tts_bn_train_fwd(const float* z, const float* gamma, const float* beta, float* mean, float* invstd,
                     float* running_mean, float* running_var, int64_t* num_batches, float momentum, float eps,
                     int32_t act_tanh, float drop_p, uint64_t seed, uint32_t rng_stream, const int32_t* lengths,
                     int32_t batch, int32_t frames, int32_t channels, uint16_t* out_pad, float* out_f32,
                     const float* residual, float* scratch, void* stream);
int tts_bn_train_bwd(const float* z, const float* dout, const float* gamma, const float* beta, const float* mean,
                     const float* invstd, int32_t act_tanh, float drop_p, uint64_t seed, uint32_t rng_stream,
                     const int32_t* lengths, int32_t mask_rows, int32_t batch, int32_t frames, int32_t channels,
                     uint16_t* dz_pad, float* dgamma, float* dbeta, float* scratch, void* stream);
int tts_pad_cast_bf16(const float* x, const int32_t* lengths, uint16_t* out, int32_t batch, int32_t frames,
                      int32_t channels, int32_t only_pads, void* stream);
/* compute_loss (tacotron.py:136-158) without the L2 term: sums3 = {sum mse_bef, sum mse_aft, sum bce} over valid frames,
 * per-sample after-loss sums, and the gradients of bef_loss + aft_loss + stop_loss w.r.t. the three model outputs. */
int tts_loss_train(const float* mel_bef, const float* mel_aft, const float* stop_logits, const float* targets,
                   const int32_t* lengths, const int32_t* total_len, int32_t batch, int32_t frames, int32_t n_mels,
                   float pos_weight, float* sums3, float* aft_per_sample, float* d_bef, float* d_aft, float* d_stop,
                   void* stream);
/* Multi-tensor L2 term and fused Adam over a device table of {float* p; const float* g; float* m; float* v; int64 n;
 * int64 first_chunk; int32 decay; int32 pad} (train.py:130-131,188; tacotron.py:144-146). */
int tts_sumsq_multi(const void* table_dev, int32_t n_entries, int64_t n_chunks, float* out, void* stream);
int tts_adam_multi(const void* table_dev, int32_t n_entries, int64_t n_chunks, float lr, float beta1, float beta2,
                   float eps, int64_t step, float reg_weight, float grad_scale, void* stream);

/* ---- mel -> waveform ------------------------------------------------------ */

/* utils/audio.py:53-99 of the reference: mel2wav = de-normalise, dB -> amplitude, mel_to_linear (pseudo-inverse of
 * librosa.filters.mel, clamp 1e-10), ^power, griffin_lim (n_iter x {librosa.istft, librosa.stft, phase of the estimate on
 * the target magnitude} + a last istft; librosa 0.6.0: Hann window of win_length centred in n_fft, centre = True with
 * reflect padding, window-sum-of-squares normalisation), scipy.signal.lfilter([1], [1, -preemphasis]).  A whole batch at
 * once; utterance b has lengths[b] frames and hop_length * (lengths[b] - 1) output samples.  Host-computed constants:
 * inv_basis_t [n_mels][n_fft/2+1] = pinv(mel basis) transposed, window [win_length] (periodic Hann), twiddle [n_fft]
 * complex: exp(-2 pi i k / n_fft) for k < n_fft/2, then the per-pass radix-4 twiddles (tts_b200/vocoder.py).  Scratch: mag [batch][frames_max][n_fft/2+1], frames [batch][frames_max][win_length],
 * y [batch][ldy]; output wav [batch][ldw].  Built for n_fft 2048 / hop 200 / win 800 (hyperparams.py:7-15). */
typedef struct TtsGriffinLim {
  const float* mel;           /* [batch][frames_max][n_mels], the model's normalised mel (synthesize.py:82) */
  const int32_t* lengths;     /* [batch] frames per utterance, each >= 7 */
  const float* inv_basis_t;
  const float* window;
  const float* twiddle;       /* [n_fft][2] (re, im) */
  int32_t batch, frames_max, min_frames, n_mels;
  int32_t n_fft, hop_length, win_length, n_iter;
  float max_abs, max_db, ref_db, power, preemphasis;
  float* mag;
  float* frames;
  float* y; int64_t ldy;
  float* wav; int64_t ldw;
} TtsGriffinLim;
int tts_griffin_lim(const TtsGriffinLim* g, void* stream);

/* ---- autoregressive decode (the hot path) ------------------------------------------------ */

typedef struct TtsDecLayerWeights {
  const float* ln_self_g;   /* decoder.decoder.attn_layer_norms.{l}.weight   [D] */
  const float* ln_self_b;
  const float* w_qkv;       /* ...self_attentions.{l}.qkv_transform.weight   [3D][D] */
  const float* w_self_out;  /* ...self_attentions.{l}.output_transform.weight [D][D] */
  const float* ln_cross_g;  /* ...encdec_layer_norms.{l} */
  const float* ln_cross_b;
  const float* w_cross_q;   /* ...encdec_attentions.{l}.q_transform.weight   [D][D] */
  const float* w_cross_kv;  /* ...encdec_attentions.{l}.kv_transform.weight  [2D][D] */
  const float* w_cross_out; /* ...encdec_attentions.{l}.output_transform.weight */
  const float* ln_ffn_g;    /* ...ffn_layer_norms.{l} */
  const float* ln_ffn_b;
  const float* w_ffn_in;    /* ...ffn_layers.{l}.input_layer.weight  [4D][D] */
  const float* w_ffn_out;   /* ...ffn_layers.{l}.output_layer.weight [D][4D] */
  /* Derived ("packed") operands of the fused kernel: the LayerNorm in front of a projection folded into it,
   * W_ln = W * diag(gamma) and c_ln = W * beta, so that LN(x) W^T = ((x - mean) * rstd) W_ln^T + c_ln.
   * Rebuilt by the host whenever W / gamma / beta change; NULL = fused kernel unavailable. */
  const float* w_qkv_ln;     /* [3D][D] */
  const float* c_qkv_ln;     /* [3D] */
  const float* w_cross_q_ln; /* [D][D] */
  const float* c_cross_q_ln; /* [D] */
  const float* w_ffn_in_ln;  /* [4D][D] */
  const float* c_ffn_in_ln;  /* [4D] */
  /* Packed operands of the pipelined kernel: one row per output n, K + 16 floats:
   *   [0, K)  the weight row (LayerNorm-folded where a LayerNorm precedes the projection),
   *   [K]     the additive constant (bias, or c_ln[n]),
   *   [K+1]   s_ln[n] = sum_k W_ln[n][k]: LayerNorm is applied AFTER the product,
   *           LN(x) W^T = rstd * (x W_ln^T) - rstd * mean * s_ln + c_ln,
   *   rest    zero.  A CTA's slice of rows is one contiguous run: one TMA bulk copy per phase brings weights
   *           and epilogue constants, already laid out with the bank-spreading row padding the MMA loads want.
   * pk_ffn_out is K-split: [pk_ksplit][D][4D / pk_ksplit + 16]. */
  const float* pk_qkv;       /* [3D][D+16] */
  const float* pk_self_out;  /* [D][D+16] */
  const float* pk_cross_q;   /* [D][D+16] */
  const float* pk_cross_out; /* [D][D+16] */
  const float* pk_ffn_in;    /* [4D][D+16] */
  const float* pk_ffn_out;   /* [pk_ksplit][D][4D/pk_ksplit+16] */
} TtsDecLayerWeights;

typedef struct TtsDecoderWeights {
  int32_t n_layers, d_model, n_heads, d_ffn, n_mels, prenet_hidden;
  const float* prenet_w0;   /* decoder.prenet.dense0.weight [P][M] */
  const float* prenet_b0;
  const float* prenet_w1;   /* [P][P] */
  const float* prenet_b1;
  const float* prenet_w2;   /* dense_final.weight [D][P] */
  const float* pe_scale;    /* decoder.decoder.pe_scale, device scalar */
  const float* pe_table;    /* [t_max][D] sinusoid table (common.py:4-29), built once */
  const float* ln_out_g;    /* decoder.decoder.output_layer_norm */
  const float* ln_out_b;
  const float* w_mel;       /* decoder.mel_net.weight [M][D] */
  const float* w_stop;      /* decoder.stop_net.weight [1][D] */
  const float* b_stop;      /* decoder.stop_net.bias [1] */
  const float* w_mel_ln;    /* [M][D]  mel_net with the output LayerNorm folded in (see TtsDecLayerWeights) */
  const float* w_stop_ln;   /* [1][D] */
  const float* c_out_ln;    /* [M+1]   W_mel * beta_out, w_stop * beta_out */
  const float* pk_pre0;     /* [P][M+16]   packed like TtsDecLayerWeights.pk_* */
  const float* pk_pre1;     /* [P][P+16] */
  const float* pk_pre2;     /* [D][P+16] */
  const float* pk_final;    /* [M+1+P][D+16]  w_mel_ln rows, w_stop_ln (constant = c_out_ln), then prenet_w0 * w_mel_ln
                             * (constant = prenet_w0 * c_mel + prenet_b0): the next step's first prenet layer */
  int32_t pk_ksplit;        /* K split of pk_ffn_out: ceil(4D / 768) */
  int32_t pk_reserved;
  TtsDecLayerWeights layer[TTS_MAX_LAYERS];
} TtsDecoderWeights;

/* Per-utterance-batch decode state.  All buffers are caller-owned device memory.
 * Cache layout: [L][B][H][rows][head_dim] with K and V separate, so that the K (or V) stream
 * of one (layer, sample, head) is one contiguous run of rows*head_dim floats. */
typedef struct TtsDecodeState {
  int32_t batch, mem_len, t_max;
  const float* memory;          /* [B][S][D]  encoder outputs (tacotron.py:44) */
  const int32_t* input_lengths; /* [B] */
  float* self_k;                /* [L][B][H][t_max][dh] */
  float* self_v;
  float* cross_k;               /* [L][B][H][S][dh] */
  float* cross_v;
  int32_t* lengths;             /* [B] target_lengths, starts at 1 (synthesize.py:23) */
  uint8_t* finished;            /* [B] (synthesize.py:24) */
  float* frames;                /* [B][t_max][M]  mel_pre (synthesize.py:43) */
  float* stop_logits;           /* [B][t_max] */
  float* align_self;            /* NULL or [L][B][H][t_max][t_max]  rows = query step */
  float* align_cross;           /* NULL or [L][B][H][t_max][S] */
  int32_t* step_counter;        /* device scalar: next step index t */
  int32_t* n_unfinished;        /* device scalar, refreshed every step */
  float* scratch;               /* tts_decode_scratch_bytes() bytes */
  /* decoder.train() at synthesis time (eval.py:116-117, train.py:229-234): Philox dropout with rate drop_p_prenet after
   * the two prenet ReLUs (tacotron.py:58,62) and drop_p_transformer after the position-encoding add, on the attention
   * weights (after the softmax normalisation; the recorded alignment rows are pre-dropout like attention.py:88), on
   * every residual-branch output and on the FFN hidden (modules.py:18,120,132,138,141).  Masks are a pure function of
   * (drop_seed, site, step, row, element): csrc/philox.cuh.  0 / 0 = deterministic eval() semantics. */
  float drop_p_prenet, drop_p_transformer;
  uint64_t drop_seed;
} TtsDecodeState;

size_t tts_decode_scratch_bytes(const TtsDecoderWeights* w, int32_t batch, int32_t mem_len,
                                int32_t t_max);

/* Once per utterance batch: cross_k/v[l] = split_heads(kv_transform_l(memory))
 * (attention.py:66-68, recomputed EVERY step by the reference, synthesize.py:39-41), and
 * reset lengths=1, finished=0, step_counter=0. */
int tts_decode_begin(const TtsDecoderWeights* w, const TtsDecodeState* st, void* stream);

/* Run `n_steps` cached decode steps starting at *step_counter (the body
 * of the loop synthesize.py:35-45 with transformer/tacotron.py:107-116 inside).
 *   prev_mel / prev_mel_stride: where step t reads frame t-1 from.  NULL = st->frames.
 *   update_state: 1 = also perform synthesize.py:42-45 on device (finished |= stop>0,
 *                 lengths += !finished); 0 = the caller does it (unchanged eval_batch).
 *   impl: 0 = default (4 when the shape is supported, else 2),
 *         1 = per-phase kernels, 2 = per-phase kernels replayed from a CUDA graph,
 *         4 = pipelined persistent kernel (row-group pipelining, producer warp, 3xTF32 mma.sync),
 *         5 = experiment: two CTAs per SM, one row group each on its own phase clock (batch > 16 rows; measured slower).
 * *st->n_unfinished after the call: rows still decoding (>= 0); < 0 = a barrier wait timed out; -2 or <= -2^29 = the
 * call asked for steps beyond t_max (nothing past t_max was touched). */
int tts_decode_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int32_t n_steps,
                     const float* prev_mel, int64_t prev_mel_stride, int32_t update_state,
                     int32_t impl, void* stream);

/* Diagnostics: SM-clock stamps {phase start, compute done, barrier passed} that CTA 0 of the fused
 * kernel recorded for every phase of the last step it ran (synchronous device-to-host copy). */
int tts_decode_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, int64_t* out_host,
                       int32_t max_entries);

#ifdef __cplusplus
}
#endif
#endif /* TTS_B200_H */
