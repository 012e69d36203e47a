// Dense building blocks of the teacher-forced / prefill side of the mel path (sm_100a):
// fp32 GEMM (packed-FFMA2 register tiles) with the fused epilogues the model needs, LayerNorm,
// the encoder/decoder prologues and the conditioning embeddings.
//
// Reference call sites replaced: every nn.Linear / nn.Conv1d / nn.LayerNorm / nn.Embedding on
// the path — transformer/attention.py:43-47, modules.py:11-19,36-47,49-56,88-106,114-118,
// tacotron.py:21-44,50-64,78-90,112-115.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace tts {

// ---------------------------------------------------------------------------------------------
// GEMM: C[M,N] = epi(A[M,K] * W[N,K]^T).  Both operands are K-contiguous ("NT").
// CTA tile BM x BN, BK = 16, 256 threads, each thread owns a TM x TN register tile whose
// accumulators are packed pairs along N so the inner product runs on FFMA2.
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ A, int lda,
                                                      const float* __restrict__ W, int ldw,
                                                      float* __restrict__ C, int ldc, int M, int N,
                                                      int K, TtsGemmEpilogue epi) {
  constexpr int BK = 16;
  constexpr int PAD = 4;
  static_assert((BM / TM) * (BN / TN) == 256, "thread tile must cover the CTA tile");
  static_assert(TN % 4 == 0 && TM % 4 == 0, "");
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  constexpr int A_LD = BM * (BK / 4) / 256;  // float4 loads per thread per tile
  constexpr int B_LD = BN * (BK / 4) / 256;
  static_assert(A_LD >= 1 && B_LD >= 1, "");
  float4 ra[A_LD], rb[B_LD];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256, r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
      const int gm = m0 + r, gk = k0 + c;
      ra[i] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256, r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
      const int gn = n0 + r, gk = k0 + c;
      rb[i] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256, r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
      As[buf][c + 0][r] = ra[i].x; As[buf][c + 1][r] = ra[i].y;
      As[buf][c + 2][r] = ra[i].z; As[buf][c + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256, r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
      Bs[buf][c + 0][r] = rb[i].x; Bs[buf][c + 1][r] = rb[i].y;
      Bs[buf][c + 2][r] = rb[i].z; Bs[buf][c + 3][r] = rb[i].w;
    }
  };

  f32x2 acc[TM][TN / 2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = 0ull;

  const int n_tiles = ceil_div(K, BK);
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) load_tiles((t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM];
      f32x2 b[TN / 2];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const f32x4 v = lds128(&Bs[buf][k][tx * TN + j]);
        b[j / 2] = v.lo; b[j / 2 + 1] = v.hi;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const f32x2 aa = pack2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc[i][j] = fma2(aa, b[j], acc[i][j]);
      }
    }
    if (t + 1 < n_tiles) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue -----------------------------------------------------------------------------
  const int rpb = epi.rows_per_batch > 0 ? epi.rows_per_batch : M;
  const int valid = epi.valid_rows > 0 ? epi.valid_rows : rpb;
  const int orpb = epi.out_rows_per_batch > 0 ? epi.out_rows_per_batch : rpb;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    const int b = m / rpb, r = m - b * rpb;
    if (r >= valid) continue;
    const bool dead = epi.row_len != nullptr && r >= epi.row_len[b];
    const size_t orow = (size_t)b * orpb + r + epi.out_row_offset;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float lo, hi;
      unpack2(acc[i][j / 2], lo, hi);
      float v = ((j & 1) ? hi : lo) * epi.alpha;
      if (epi.scale) v = v * epi.scale[n] + epi.shift[n];
      if (epi.bias) v += epi.bias[n];
      if (epi.act == 1) v = fmaxf(v, 0.f);
      else if (epi.act == 2) v = tanhf(v);
      if (epi.residual) v += epi.residual[orow * epi.ldr + n];
      if (dead) v = 0.f;
      if (epi.head_dim > 0) {
        const int dm = epi.n_heads * epi.head_dim;
        const int w = n / dm, hn = n - w * dm, h = hn / epi.head_dim, d = hn - h * epi.head_dim;
        float* dst = w ? epi.out_v : C;
        dst[(((size_t)b * epi.n_heads + h) * epi.head_rows + r) * epi.head_dim + d] = v;
      } else {
        C[orow * ldc + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel axis, one warp per row, two-pass statistics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ g,
                                                        const float* __restrict__ bta, int rows, int C,
                                                        float eps, const int32_t* __restrict__ row_len,
                                                        int rows_per_batch) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  bool dead = false;
  if (row_len != nullptr) {
    const int b = row / rows_per_batch;
    dead = (row - b * rows_per_batch) >= row_len[b];
  }
  float* yr = y + (size_t)row * C;
  for (int c = lane; c < C; c += 32) yr[c] = dead ? 0.f : (xr[c] - mean) * rstd * g[c] + bta[c];
}

// ---------------------------------------------------------------------------------------------
// prologues / conditioning
// ---------------------------------------------------------------------------------------------
__global__ void embed_pe_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ len,
                                const float* __restrict__ embed, const float* __restrict__ pe,
                                const float* __restrict__ pe_scale, float* __restrict__ out, int B, int S,
                                int C, int vocab) {
  const int row = blockIdx.x;  // b*S + s
  const int b = row / S, s = row - b * S;
  int64_t id = ids ? ids[row] : row;  // ids == NULL: `embed` already holds one row per position
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float live = s < len[b] ? 1.f : 0.f;
  const float sc = *pe_scale;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[(size_t)row * C + c] = embed[(size_t)id * C + c] * live + pe[(size_t)s * C + c] * sc;
}

__global__ void shift_pe_kernel(const float* __restrict__ pre, const int32_t* __restrict__ len,
                                const float* __restrict__ pe, const float* __restrict__ pe_scale,
                                float* __restrict__ out, int B, int T, int C) {
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row - b * T;
  const float sc = *pe_scale;
  const bool have = t > 0 && (t - 1) < len[b];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float p = have ? pre[(size_t)(row - 1) * C + c] : 0.f;
    out[(size_t)row * C + c] = p + pe[(size_t)t * C + c] * sc;
  }
}

__global__ void pad_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ len,
                                float* __restrict__ out, int T, int C, int pad) {
  const int TP = T + 2 * pad;
  const int row = blockIdx.x;  // b*TP + r
  const int b = row / TP, r = row - b * TP, t = r - pad;
  const bool live = t >= 0 && t < T && (len == nullptr || t < len[b]);
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[(size_t)row * C + c] = live ? x[((size_t)b * T + t) * C + c] : 0.f;
}

// mem[b, s, off + e] = softsign(b2[e] + sum_j w2[e][j] * h[j])
// ids == NULL: h[j] = sum_i w1[j][i] * vec[b][i]  (language one-hot -> Linear, tacotron.py:21-25)
// ids != NULL: h[j] = w1[ids[b]][j]               (speaker embedding row, tacotron.py:27-31)
__global__ void cond_embed_kernel(const float* __restrict__ vec, int vec_dim, const int64_t* __restrict__ ids,
                                  const float* __restrict__ w1,
                                  const float* __restrict__ w2, const float* __restrict__ b2, int E,
                                  float* __restrict__ mem, int S, int width, int off) {
  extern __shared__ float sh[];  // h[E], o[E]
  float* h = sh;
  float* o = sh + E;
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    float a;
    if (ids == nullptr) {
      a = 0.f;
      for (int i = 0; i < vec_dim; ++i) a += w1[(size_t)j * vec_dim + i] * vec[(size_t)b * vec_dim + i];
    } else {
      a = w1[(size_t)ids[b] * E + j];
    }
    h[j] = a;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float a = b2[e];
    for (int j = 0; j < E; ++j) a += w2[(size_t)e * E + j] * h[j];
    o[e] = a / (1.f + fabsf(a));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * E; i += blockDim.x) {
    const int s = i / E, e = i - s * E;
    mem[((size_t)b * S + s) * width + off + e] = o[e];
  }
}

}  // namespace tts

using namespace tts;

namespace tts {
// 1 = large dense GEMMs with K % 32 == 0 run on the tcgen05 kernel (default); TTS_GEMM_TC=0 keeps the FFMA2 kernels
static std::atomic<int> g_gemm_tc{[]() { const char* e = getenv("TTS_GEMM_TC"); return e != nullptr ? (e[0] == '1' ? 1 : 0) : 1; }()};
bool gemm_tc_supported(int M, int N, int K, int lda, int ldw);
int launch_gemm_tc(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                   const TtsGemmEpilogue& epi, cudaStream_t s);
}  // namespace tts

extern "C" int tts_gemm_use_tensor_cores(int on) {
  const int prev = g_gemm_tc.load();
  if (on >= 0) g_gemm_tc.store(on ? 1 : 0);
  return prev;
}

extern "C" int tts_gemm_nt(const float* A, int32_t lda, const float* W, int32_t ldw, float* C, int32_t ldc,
                           int32_t M, int32_t N, int32_t K, const TtsGemmEpilogue* epi_in, void* stream) {
  TTS_REQUIRE(M > 0 && N > 0 && K > 0, "tts_gemm_nt: empty problem M=%d N=%d K=%d", M, N, K);
  TTS_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "tts_gemm_nt: K/lda/ldw must be multiples of 4 (%d,%d,%d)", K, lda, ldw);
  TTS_REQUIRE((((uintptr_t)A | (uintptr_t)W) & 15) == 0, "tts_gemm_nt: operands must be 16-byte aligned");
  TtsGemmEpilogue epi;
  if (epi_in) {
    epi = *epi_in;
  } else {
    epi = TtsGemmEpilogue{};
    epi.alpha = 1.f;
  }
  TTS_REQUIRE(epi.scale == nullptr || epi.shift != nullptr, "tts_gemm_nt: scale without shift");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long big_ctas = (long)ceil_div(M, 128) * ceil_div(N, 128);
  // large, K % 32 == 0 problems run on the tcgen05 3xTF32 kernel (gemm_tc.cu); TTS_GEMM_TC=0 keeps the FFMA2 kernels
  if (g_gemm_tc.load(std::memory_order_relaxed) != 0 && M >= 256 && big_ctas >= 32 && gemm_tc_supported(M, N, K, lda, ldw) && (reinterpret_cast<uintptr_t>(C) & 3) == 0)
    return launch_gemm_tc(A, lda, W, ldw, C, ldc, M, N, K, epi, s);
  if (big_ctas >= 148) {
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128));
    gemm_nt_kernel<128, 128, 8, 8><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, epi);
  } else {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    gemm_nt_kernel<64, 64, 4, 4><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, epi);
  }
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_layernorm(const float* x, float* y, const float* gamma, const float* beta, int32_t rows,
                             int32_t channels, float eps, const int32_t* row_len, int32_t rows_per_batch,
                             void* stream) {
  TTS_REQUIRE(rows > 0 && channels > 0, "tts_layernorm: empty input");
  TTS_REQUIRE(row_len == nullptr || rows_per_batch > 0, "tts_layernorm: row_len needs rows_per_batch");
  layernorm_kernel<<<ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, gamma, beta, rows, channels, eps, row_len, rows_per_batch);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_embed_pe(const int64_t* ids, const int32_t* lengths, const float* embed, const float* pe,
                            const float* pe_scale, float* out, int32_t batch, int32_t seq, int32_t channels,
                            int32_t vocab, void* stream) {
  TTS_REQUIRE(batch > 0 && seq > 0, "tts_embed_pe: empty input");
  embed_pe_kernel<<<batch * seq, 128, 0, static_cast<cudaStream_t>(stream)>>>(ids, lengths, embed, pe, pe_scale,
                                                                             out, batch, seq, channels, vocab);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_shift_pe(const float* pre, const int32_t* lengths, const float* pe, const float* pe_scale,
                            float* out, int32_t batch, int32_t frames, int32_t channels, void* stream) {
  TTS_REQUIRE(batch > 0 && frames > 0, "tts_shift_pe: empty input");
  shift_pe_kernel<<<batch * frames, 128, 0, static_cast<cudaStream_t>(stream)>>>(pre, lengths, pe, pe_scale, out,
                                                                                batch, frames, channels);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_pad_rows(const float* x, const int32_t* lengths, float* out, int32_t batch, int32_t frames,
                            int32_t channels, int32_t pad, void* stream) {
  TTS_REQUIRE(batch > 0 && frames > 0 && channels > 0 && pad >= 0, "tts_pad_rows: bad shape");
  pad_rows_kernel<<<batch * (frames + 2 * pad), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, lengths, out, frames,
                                                                                          channels, pad);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_cond_embed(const float* vec, int32_t vec_dim, const int64_t* ids, const float* w1, const float* w2,
                              const float* b2, int32_t emb, float* mem, int32_t batch, int32_t seq,
                              int32_t mem_width, int32_t col_offset, void* stream) {
  TTS_REQUIRE(batch > 0 && seq > 0 && emb > 0, "tts_cond_embed: empty input");
  cond_embed_kernel<<<batch, 256, 2 * emb * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      vec, vec_dim, ids, w1, w2, b2, emb, mem, seq, mem_width, col_offset);
  TTS_CHECK_LAUNCH();
  return 0;
}
