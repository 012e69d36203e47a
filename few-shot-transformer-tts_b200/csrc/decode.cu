// The autoregressive decode step over a K/V cache — per-phase kernels (impl 1/2).
//
// One step = SURVEY.md Appendix A = the body of synthesize.py:35-45 with
// transformer/tacotron.py:107-116 and transformer/modules.py:108-145 inside, restated so that
// only ONE decoder row per sample is computed and the self/cross K/V are read from a cache
// instead of being recomputed (the reference re-runs the whole decoder over all frames so far
// and re-projects the encoder memory every step).
//
// Phases of a step (all read the step index t from device memory so that one CUDA graph can be
// replayed for every step):
//   prenet x3            skinny GEMMs 80->P->P->D (+bias, ReLU), then mask + PE      (tacotron.py:55-65)
//   per layer:  LN+QKV   skinny GEMM, q scaled, k/v appended to the cache            (attention.py:63-64)
//               self-attention over the cache, optional split-KV + combine           (attention.py:83-91)
//               out-proj + residual                                                  (attention.py:119, modules.py:132)
//               LN+Q (cross), cross-attention over the per-utterance cache, out-proj + residual
//               LN+FFN-in+ReLU, FFN-out + residual                                   (modules.py:8-20,140-141)
//   final LN + mel/stop projections + masks                                          (modules.py:142-144, tacotron.py:112-115)
//   advance: finished |= stop>0 ; lengths += !finished ; t += 1                      (synthesize.py:42-45)
//
// The skinny GEMM streams each weight row exactly once from HBM straight into registers
// (128-bit no-allocate loads) and multiplies it against all <=32 batch rows held in shared
// memory; products run on packed FFMA2.
#include <math_constants.h>
#include <string.h>

#include "common.cuh"

namespace tts {

constexpr int kRowsPerBlock = 32;  // batch rows per CTA of the skinny GEMM
constexpr int kMaxWRows = 8;       // weight rows per CTA pass
constexpr int kKChunk = 768;       // K slice of X resident in shared memory

enum SkinnyMode { kPlain = 0, kQkv = 1, kPrenetOut = 2, kFinal = 3 };

struct SkinnyArgs {
  const float* X; long long ldx; int B, K, N;
  int x_from_frames;                 // X = frames[:, t-1, :] (zeros at t == 0)
  const float* W; const float* W2; int n_w1;
  const float* ln_g; const float* ln_b;
  const float* bias; int relu; int mode; int w_rows;
  float* Y; long long ldy; const float* R; long long ldr; float out_scale;
  float* kcache; float* vcache; int H, dh, t_max;
  const int32_t* lengths; const float* pe; const float* pe_scale;
  float* frames; float* stop_logits; const float* b_stop; int n_mels;
  const int32_t* step;
};

template <bool LN>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(SkinnyArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int kc_max = a.K < kKChunk ? a.K : kKChunk;
  const int ldxs = kc_max + 4;  // row stride = 4 (mod 32) words: conflict-free 128-bit reads, 4 rows apart
  float* Xs = smem;                                   // [32][ldxs]
  float* red = Xs + kRowsPerBlock * ldxs;             // [8 warps][kMaxWRows][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, s4 = (lane & 3) * 4;       // rows 4g..4g+3, floats s4..s4+3 of a 16-chunk
  const int t = *a.step;
  if (t >= a.t_max) return;   // past the session's capacity: advance_kernel reports it, nothing is written
  const int b0 = blockIdx.y * kRowsPerBlock;
  const int n0 = blockIdx.x * a.w_rows;
  const int nrows = min(a.w_rows, a.N - n0);
  const float* X = a.x_from_frames ? a.frames + (size_t)(t > 0 ? t - 1 : 0) * a.n_mels : a.X;

  f32x2 acc[kMaxWRows][4];
#pragma unroll
  for (int r = 0; r < kMaxWRows; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0ull;

  for (int k0 = 0; k0 < a.K; k0 += kKChunk) {
    const int kc = min(kKChunk, a.K - k0);
    if (k0 > 0) __syncthreads();
    // ---- stage X[b0:b0+32, k0:k0+kc] (zero rows beyond B) -----------------------------------
    for (int i = tid; i < kRowsPerBlock * (kc / 4); i += 256) {
      const int r = i / (kc / 4), c = (i % (kc / 4)) * 4;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < a.B && !(a.x_from_frames && t == 0))
        val = *reinterpret_cast<const float4*>(X + (size_t)(b0 + r) * a.ldx + k0 + c);
      *reinterpret_cast<float4*>(Xs + r * ldxs + c) = val;
    }
    __syncthreads();
    if (LN) {  // whole row resident (K <= kKChunk): normalise in place, 4 rows per warp
      for (int r = warp * 4; r < warp * 4 + 4; ++r) {
        float* xr = Xs + r * ldxs;
        float s = 0.f;
        for (int c = lane; c < kc; c += 32) s += xr[c];
        const float mean = warp_sum(s) / kc;
        float q = 0.f;
        for (int c = lane; c < kc; c += 32) {
          const float d = xr[c] - mean;
          q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / kc + 1e-6f);
        for (int c = lane; c < kc; c += 32) xr[c] = (xr[c] - mean) * rstd * a.ln_g[c] + a.ln_b[c];
      }
      __syncthreads();
    }
    // ---- stream the weight rows: warp w owns 16-float chunks w, w+8, ... ---------------------
    const int n_chunks = kc / 16;
#pragma unroll 2
    for (int c = warp; c < n_chunks; c += 8) {
      const int kk = c * 16 + s4;
      f32x4 wv[kMaxWRows];
#pragma unroll
      for (int r = 0; r < kMaxWRows; ++r) {
        if (r < nrows) {
          const int n = n0 + r;
          const float* wrow = n < a.n_w1 ? a.W + (size_t)n * a.K : a.W2 + (size_t)(n - a.n_w1) * a.K;
          wv[r] = ldg_stream(wrow + k0 + kk);
        } else {
          wv[r].lo = wv[r].hi = 0ull;
        }
      }
      f32x4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = lds128(Xs + (4 * g + i) * ldxs + kk);
#pragma unroll
      for (int r = 0; r < kMaxWRows; ++r)
        if (r < nrows) {  // CTA-uniform
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[r][i] = fma2(xv[i].lo, wv[r].lo, acc[r][i]);
            acc[r][i] = fma2(xv[i].hi, wv[r].hi, acc[r][i]);
          }
        }
    }
  }

  // ---- reduce: pairs -> 4 k-split lanes -> 8 warps ---------------------------------------------
#pragma unroll
  for (int r = 0; r < kMaxWRows; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = hsum2(acc[r][i]);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if ((lane & 3) == 0) red[(warp * kMaxWRows + r) * 32 + 4 * g + i] = v;
    }
  __syncthreads();

  // ---- epilogue: one thread per (weight row, batch row) ---------------------------------------
  const int r = tid >> 5, br = tid & 31;
  if (r >= nrows || b0 + br >= a.B) return;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) v += red[(w * kMaxWRows + r) * 32 + br];
  const int n = n0 + r, b = b0 + br;
  if (a.bias) v += a.bias[n];
  if (a.relu) v = fmaxf(v, 0.f);
  switch (a.mode) {
    case kPlain: {
      v *= a.out_scale;
      if (a.R) v += a.R[(size_t)b * a.ldr + n];
      a.Y[(size_t)b * a.ldy + n] = v;
    } break;
    case kQkv: {
      const int D = a.H * a.dh;
      const int which = n / D, c = n - which * D;
      if (which == 0) {
        a.Y[(size_t)b * a.ldy + c] = v * a.out_scale;
      } else {
        const int h = c / a.dh, d = c - h * a.dh;
        float* dst = which == 1 ? a.kcache : a.vcache;
        dst[(((size_t)b * a.H + h) * a.t_max + t) * a.dh + d] = v;
      }
    } break;
    case kPrenetOut: {  // modules.py:114-118: impute, shift right, + pe * pe_scale
      const bool have = t > 0 && (t - 1) < a.lengths[b];
      a.Y[(size_t)b * a.ldy + n] = (have ? v : 0.f) + a.pe[(size_t)t * a.N + n] * (*a.pe_scale);
    } break;
    case kFinal: {      // modules.py:144, tacotron.py:112-115
      const bool live = t < a.lengths[b];
      if (n < a.n_mels) a.frames[((size_t)b * a.t_max + t) * a.n_mels + n] = live ? v : 0.f;
      else a.stop_logits[(size_t)b * a.t_max + t] = live ? v + a.b_stop[0] : 0.f;
    } break;
  }
}

// final LayerNorm is masked by `live` BEFORE the projections in the reference; since the
// projections are linear and bias-free (mel) or masked again (stop), masking the outputs is
// identical — except for the stop bias, which the reference also masks (tacotron.py:115).

// ---------------------------------------------------------------------------------------------
// decode attention: one query row per (sample, head) against a cached K/V stream
// ---------------------------------------------------------------------------------------------
struct DecAttnArgs {
  const float* q; int B, H; int n_split;
  const float* kc; const float* vc; int rows_alloc;      // [B][H][rows_alloc][dh]
  int n_keys_fixed;                                      // >0: cross (S); 0: self (t+1)
  const int32_t* key_len;                                // cross: input_lengths
  float* out;                                            // [B][H*dh]
  float* part_o; float* part_m; float* part_l;           // [B][H][n_split][dh|1|1]
  float* align; long long align_bh_stride; int align_row_len;  // row for this step: align + bh*stride + t*row_len
  const int32_t* step;
  int t_max;
};

template <int DH>
__global__ void __launch_bounds__(256) decode_attn_kernel(DecAttnArgs a) {
  constexpr int F4 = DH / 32;  // float4 per lane per key (8 lanes span one key row)
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_red[8][DH + 1];
  __shared__ float s_max[8], s_sum[8];
  float* sc = smem;  // scores of this CTA's key range

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kslot = lane >> 3, l8 = lane & 7;
  const int split = blockIdx.x % a.n_split, bh = blockIdx.x / a.n_split;
  const int b = bh / a.H;
  const int t = *a.step;
  if (t >= a.t_max) return;
  const int n_keys = a.n_keys_fixed > 0 ? a.n_keys_fixed : t + 1;
  const int klen = a.key_len ? a.key_len[b] : n_keys;
  const int per = ceil_div(n_keys, a.n_split);
  const int j0 = split * per, j1 = min(n_keys, j0 + per);
  const float* kbase = a.kc + (size_t)bh * a.rows_alloc * DH;
  const float* vbase = a.vc + (size_t)bh * a.rows_alloc * DH;

  f32x4 qv[F4];
#pragma unroll
  for (int i = 0; i < F4; ++i) qv[i] = ldg_cg(a.q + (size_t)bh * DH + 4 * (l8 + 8 * i));

  // ---- pass 1: scores -------------------------------------------------------------------------
  float lmax = -CUDART_INF_F;
  for (int jb = j0 + warp * 4; jb < j1; jb += 32 * 4) {
    f32x4 kv[4][F4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * 32 + kslot;
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        if (j < j1) kv[u][i] = ldg_cg(kbase + (size_t)j * DH + 4 * (l8 + 8 * i));
        else kv[u][i].lo = kv[u][i].hi = 0ull;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * 32 + kslot;
      f32x2 acc = 0ull;
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        acc = fma2(qv[i].lo, kv[u][i].lo, acc);
        acc = fma2(qv[i].hi, kv[u][i].hi, acc);
      }
      float s = hsum2(acc);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (j < j1) {
        if (j >= klen) s = kNegBias;
        if (l8 == 0) sc[j - j0] = s;
        lmax = fmaxf(lmax, s);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_max[warp] = lmax;
  __syncthreads();
  float m = s_max[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w]);

  // ---- pass 2: weights and weighted sum of V -----------------------------------------------------
  f32x2 o[F4][2];
#pragma unroll
  for (int i = 0; i < F4; ++i) o[i][0] = o[i][1] = 0ull;
  float lsum = 0.f;
  for (int jb = j0 + warp * 4; jb < j1; jb += 32 * 4) {
    f32x4 vv[4][F4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * 32 + kslot;
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        if (j < j1) vv[u][i] = ldg_cg(vbase + (size_t)j * DH + 4 * (l8 + 8 * i));
        else vv[u][i].lo = vv[u][i].hi = 0ull;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * 32 + kslot;
      const float p = j < j1 ? expf(sc[j - j0] - m) : 0.f;
      if (l8 == 0) lsum += p;
      const f32x2 pp = pack2(p, p);
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        o[i][0] = fma2(pp, vv[u][i].lo, o[i][0]);
        o[i][1] = fma2(pp, vv[u][i].hi, o[i][1]);
      }
    }
  }
  // reduce over the 4 key slots of the warp, then over warps
#pragma unroll
  for (int i = 0; i < F4; ++i)
#pragma unroll
    for (int hsel = 0; hsel < 2; ++hsel) {
      float x, y;
      unpack2(o[i][hsel], x, y);
      x += __shfl_xor_sync(0xffffffffu, x, 8);  y += __shfl_xor_sync(0xffffffffu, y, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16); y += __shfl_xor_sync(0xffffffffu, y, 16);
      if (kslot == 0) {
        const int d = 4 * (l8 + 8 * i) + 2 * hsel;
        s_red[warp][d] = x;
        s_red[warp][d + 1] = y;
      }
    }
  lsum = warp_sum(lsum);
  if (lane == 0) s_sum[warp] = lsum;
  __syncthreads();
  float l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += s_sum[w];

  float* arow = a.align ? a.align + (size_t)bh * a.align_bh_stride + (size_t)t * a.align_row_len : nullptr;
  if (a.n_split == 1) {
    const float inv = 1.f / l;
    if (tid < DH) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s_red[w][tid];
      a.out[(size_t)bh * DH + tid] = v * inv;
    }
    if (arow)
      for (int j = j0 + tid; j < j1; j += 256) arow[j] = expf(sc[j - j0] - m) * inv;
  } else {
    const size_t pidx = (size_t)bh * a.n_split + split;
    if (tid < DH) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s_red[w][tid];
      a.part_o[pidx * DH + tid] = v;
    }
    if (tid == 0) {
      a.part_m[pidx] = m;
      a.part_l[pidx] = l;
    }
    if (arow)
      for (int j = j0 + tid; j < j1; j += 256) arow[j] = sc[j - j0];  // raw logits, normalised by the combine
  }
}

template <int DH>
__global__ void __launch_bounds__(128) decode_attn_combine_kernel(DecAttnArgs a) {
  const int bh = blockIdx.x, tid = threadIdx.x;
  const int t = *a.step;
  if (t >= a.t_max) return;
  const int n_keys = a.n_keys_fixed > 0 ? a.n_keys_fixed : t + 1;
  float m = -CUDART_INF_F;
  for (int s = 0; s < a.n_split; ++s) m = fmaxf(m, a.part_m[(size_t)bh * a.n_split + s]);
  float l = 0.f;
  for (int s = 0; s < a.n_split; ++s) {
    const float pm = a.part_m[(size_t)bh * a.n_split + s];
    if (pm > -CUDART_INF_F) l += a.part_l[(size_t)bh * a.n_split + s] * expf(pm - m);
  }
  const float inv = 1.f / l;
  if (tid < DH) {
    float v = 0.f;
    for (int s = 0; s < a.n_split; ++s) {
      const float pm = a.part_m[(size_t)bh * a.n_split + s];
      if (pm > -CUDART_INF_F) v += a.part_o[((size_t)bh * a.n_split + s) * DH + tid] * expf(pm - m);
    }
    a.out[(size_t)bh * DH + tid] = v * inv;
  }
  if (a.align) {
    float* arow = a.align + (size_t)bh * a.align_bh_stride + (size_t)t * a.align_row_len;
    for (int j = tid; j < n_keys; j += 128) arow[j] = expf(arow[j] - m) * inv;
  }
}

// synthesize.py:42-45 on device + step counter
__global__ void advance_kernel(const float* __restrict__ stop_logits, int32_t* lengths, uint8_t* finished,
                               int32_t* step, int32_t* n_unfinished, int B, int t_max, int update_state) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int t = *step;
  if (t >= t_max) {   // a caller stepped past t_max: no buffer was touched; report instead of corrupting memory
    if (threadIdx.x == 0) *n_unfinished = -2;
    return;
  }
  int mine = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    bool fin = finished[b] != 0;
    if (update_state) {
      fin = fin || stop_logits[(size_t)b * t_max + t] > 0.f;
      finished[b] = fin ? 1 : 0;
      if (!fin) lengths[b] += 1;
    }
    mine += fin ? 0 : 1;
  }
  atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    *n_unfinished = s_cnt;
    *step = t + 1;
  }
}

__global__ void decode_reset_kernel(int32_t* lengths, uint8_t* finished, int32_t* step, int32_t* n_unfinished,
                                    int B) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    lengths[b] = 1;
    finished[b] = 0;
  }
  if (threadIdx.x == 0) {
    *step = 0;
    *n_unfinished = B;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Scratch {
  float *x, *q, *ctx, *hid, *p0, *p1, *part_o, *part_m, *part_l;
  size_t floats;
};

static int split_count(int B, int H) {
  int ns = ceil_div(2 * 148, B * H);
  return ns < 1 ? 1 : (ns > 32 ? 32 : ns);
}

static Scratch carve(const TtsDecoderWeights* w, int B, float* base) {
  Scratch s;
  const size_t D = w->d_model, F = w->d_ffn, P = w->prenet_hidden, H = w->n_heads;
  const size_t ns = split_count(B, (int)H);
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 3) / 4 * 4;
    return p;
  };
  s.x = take(B * D); s.q = take(B * D); s.ctx = take(B * D); s.hid = take(B * F);
  s.p0 = take(B * P); s.p1 = take(B * P);
  s.part_o = take(B * D * ns); s.part_m = take(B * H * ns); s.part_l = take(B * H * ns);
  s.floats = off;
  return s;
}

static int launch_skinny(SkinnyArgs a, bool ln, cudaStream_t s) {
  TTS_REQUIRE(a.K % 16 == 0, "decode: K=%d must be a multiple of 16", a.K);
  TTS_REQUIRE(!ln || a.K <= kKChunk, "decode: LayerNorm prologue needs K<=%d (got %d)", kKChunk, a.K);
  int wr = a.N / 148;
  wr = wr < 1 ? 1 : (wr > kMaxWRows ? kMaxWRows : wr);
  a.w_rows = wr;
  const int kc = a.K < kKChunk ? a.K : kKChunk;
  const size_t smem = (size_t)(kRowsPerBlock * (kc + 4) + 8 * kMaxWRows * 32) * sizeof(float);
  dim3 grid(ceil_div(a.N, wr), ceil_div(a.B, kRowsPerBlock));
  if (ln) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skinny_gemm_kernel<true><<<grid, 256, smem, s>>>(a);
  } else {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skinny_gemm_kernel<false><<<grid, 256, smem, s>>>(a);
  }
  TTS_CHECK_LAUNCH();
  return 0;
}

template <int DH>
static int launch_dec_attn_t(const DecAttnArgs& a, int max_keys, cudaStream_t s) {
  const size_t smem = (size_t)ceil_div(max_keys, a.n_split) * sizeof(float) + 16;
  TTS_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  decode_attn_kernel<DH><<<a.B * a.H * a.n_split, 256, smem, s>>>(a);
  TTS_CHECK_LAUNCH();
  if (a.n_split > 1) {
    decode_attn_combine_kernel<DH><<<a.B * a.H, 128, 0, s>>>(a);
    TTS_CHECK_LAUNCH();
  }
  return 0;
}

static int launch_dec_attn(const DecAttnArgs& a, int dh, int max_keys, cudaStream_t s) {
  switch (dh) {
    case 32: return launch_dec_attn_t<32>(a, max_keys, s);
    case 64: return launch_dec_attn_t<64>(a, max_keys, s);
    case 96: return launch_dec_attn_t<96>(a, max_keys, s);
    default: set_error("decode: head_dim %d not in {32,64,96}", dh); return 2;
  }
}

// enqueue the kernels of ONE step (reads t from st->step_counter)
int enqueue_step_phases(const TtsDecoderWeights* w, const TtsDecodeState* st, int update_state, cudaStream_t s) {
  const int B = st->batch, D = w->d_model, H = w->n_heads, dh = D / H, F = w->d_ffn, P = w->prenet_hidden;
  const int M = w->n_mels, S = st->mem_len, T = st->t_max, L = w->n_layers;
  const Scratch sc = carve(w, B, st->scratch);
  const int ns = split_count(B, H);
  const float qscale = 1.0f / sqrtf((float)dh);

  SkinnyArgs base;
  memset(&base, 0, sizeof(base));
  base.B = B; base.step = st->step_counter; base.out_scale = 1.f; base.lengths = st->lengths;
  base.frames = st->frames; base.n_mels = M; base.t_max = T; base.H = H; base.dh = dh;

  int rc;
  {  // prenet (tacotron.py:55-65) + shift/mask/PE (modules.py:114-118)
    SkinnyArgs a = base;
    a.x_from_frames = 1; a.ldx = (long long)T * M; a.K = M; a.N = P; a.W = w->prenet_w0; a.n_w1 = P;
    a.bias = w->prenet_b0; a.relu = 1; a.mode = kPlain; a.Y = sc.p0; a.ldy = P;
    if ((rc = launch_skinny(a, false, s))) return rc;
    a = base;
    a.X = sc.p0; a.ldx = P; a.K = P; a.N = P; a.W = w->prenet_w1; a.n_w1 = P; a.bias = w->prenet_b1; a.relu = 1;
    a.mode = kPlain; a.Y = sc.p1; a.ldy = P;
    if ((rc = launch_skinny(a, false, s))) return rc;
    a = base;
    a.X = sc.p1; a.ldx = P; a.K = P; a.N = D; a.W = w->prenet_w2; a.n_w1 = D; a.mode = kPrenetOut;
    a.Y = sc.x; a.ldy = D; a.pe = w->pe_table; a.pe_scale = w->pe_scale;
    if ((rc = launch_skinny(a, false, s))) return rc;
  }
  for (int l = 0; l < L; ++l) {
    const TtsDecLayerWeights& lw = w->layer[l];
    const size_t self_off = (size_t)l * B * H * T * dh, cross_off = (size_t)l * B * H * S * dh;
    SkinnyArgs a = base;  // LN + QKV, append k/v at position t
    a.X = sc.x; a.ldx = D; a.K = D; a.N = 3 * D; a.W = lw.w_qkv; a.n_w1 = 3 * D; a.ln_g = lw.ln_self_g;
    a.ln_b = lw.ln_self_b; a.mode = kQkv; a.Y = sc.q; a.ldy = D; a.out_scale = qscale;
    a.kcache = st->self_k + self_off; a.vcache = st->self_v + self_off;
    if ((rc = launch_skinny(a, true, s))) return rc;

    DecAttnArgs at;
    memset(&at, 0, sizeof(at));
    at.q = sc.q; at.B = B; at.H = H; at.n_split = ns; at.kc = st->self_k + self_off; at.vc = st->self_v + self_off;
    at.rows_alloc = T; at.n_keys_fixed = 0; at.key_len = nullptr; at.out = sc.ctx;
    at.part_o = sc.part_o; at.part_m = sc.part_m; at.part_l = sc.part_l; at.step = st->step_counter; at.t_max = T;
    if (st->align_self) {
      at.align = st->align_self + (size_t)l * B * H * T * T; at.align_bh_stride = (long long)T * T; at.align_row_len = T;
    }
    if ((rc = launch_dec_attn(at, dh, T, s))) return rc;

    a = base;  // out-proj + residual
    a.X = sc.ctx; a.ldx = D; a.K = D; a.N = D; a.W = lw.w_self_out; a.n_w1 = D; a.mode = kPlain;
    a.Y = sc.x; a.ldy = D; a.R = sc.x; a.ldr = D;
    if ((rc = launch_skinny(a, false, s))) return rc;

    a = base;  // LN + cross q
    a.X = sc.x; a.ldx = D; a.K = D; a.N = D; a.W = lw.w_cross_q; a.n_w1 = D; a.ln_g = lw.ln_cross_g;
    a.ln_b = lw.ln_cross_b; a.mode = kPlain; a.Y = sc.q; a.ldy = D; a.out_scale = qscale;
    if ((rc = launch_skinny(a, true, s))) return rc;

    memset(&at, 0, sizeof(at));
    at.q = sc.q; at.B = B; at.H = H; at.n_split = ns; at.kc = st->cross_k + cross_off; at.vc = st->cross_v + cross_off;
    at.rows_alloc = S; at.n_keys_fixed = S; at.key_len = st->input_lengths; at.out = sc.ctx;
    at.part_o = sc.part_o; at.part_m = sc.part_m; at.part_l = sc.part_l; at.step = st->step_counter; at.t_max = T;
    if (st->align_cross) {
      at.align = st->align_cross + (size_t)l * B * H * T * S; at.align_bh_stride = (long long)T * S; at.align_row_len = S;
    }
    if ((rc = launch_dec_attn(at, dh, S, s))) return rc;

    a = base;  // cross out-proj + residual
    a.X = sc.ctx; a.ldx = D; a.K = D; a.N = D; a.W = lw.w_cross_out; a.n_w1 = D; a.mode = kPlain;
    a.Y = sc.x; a.ldy = D; a.R = sc.x; a.ldr = D;
    if ((rc = launch_skinny(a, false, s))) return rc;

    a = base;  // LN + FFN-in + ReLU
    a.X = sc.x; a.ldx = D; a.K = D; a.N = F; a.W = lw.w_ffn_in; a.n_w1 = F; a.ln_g = lw.ln_ffn_g; a.ln_b = lw.ln_ffn_b;
    a.relu = 1; a.mode = kPlain; a.Y = sc.hid; a.ldy = F;
    if ((rc = launch_skinny(a, true, s))) return rc;

    a = base;  // FFN-out + residual
    a.X = sc.hid; a.ldx = F; a.K = F; a.N = D; a.W = lw.w_ffn_out; a.n_w1 = D; a.mode = kPlain;
    a.Y = sc.x; a.ldy = D; a.R = sc.x; a.ldr = D;
    if ((rc = launch_skinny(a, false, s))) return rc;
  }
  {  // final LN + mel / stop projections
    SkinnyArgs a = base;
    a.X = sc.x; a.ldx = D; a.K = D; a.N = M + 1; a.W = w->w_mel; a.W2 = w->w_stop; a.n_w1 = M;
    a.ln_g = w->ln_out_g; a.ln_b = w->ln_out_b; a.mode = kFinal; a.stop_logits = st->stop_logits; a.b_stop = w->b_stop;
    if ((rc = launch_skinny(a, true, s))) return rc;
  }
  advance_kernel<<<1, 128, 0, s>>>(st->stop_logits, st->lengths, st->finished, st->step_counter, st->n_unfinished, B, T,
                                   update_state);
  TTS_CHECK_LAUNCH();
  return 0;
}

int decode_reset(const TtsDecodeState* st, cudaStream_t s) {
  decode_reset_kernel<<<1, 128, 0, s>>>(st->lengths, st->finished, st->step_counter, st->n_unfinished, st->batch);
  TTS_CHECK_LAUNCH();
  return 0;
}

size_t decode_scratch_floats(const TtsDecoderWeights* w, int B) { return carve(w, B, nullptr).floats; }

}  // namespace tts
