// Fused persistent decode kernel (impl 3): ONE cooperative launch runs `n_steps` whole decode
// steps.  One CTA per SM stays resident; the phases of a step are separated
// by a software grid barrier instead of kernel launches, so a step costs ~52 barriers instead of
// ~65 launches and every phase starts with its operands' addresses already known.
//
//   GEMM phases   weight rows are split evenly over the CTAs; each CTA stages the <=32 batch rows
//                 of the activation in shared memory (cp.async, LayerNorm applied in place from
//                 registers), streams its weight rows once from HBM/L2 with 128-bit no-allocate
//                 loads and multiplies with packed FFMA2.
//   attention     the (sample, head, key) space of a phase is flattened and cut into equal
//                 contiguous spans, one per CTA (perfect balance for any batch size / step); a
//                 span yields at most I/G+2 partial (max, sum, weighted V) records, which the
//                 NEXT phase (the output projection) merges while staging its input, so there is
//                 no separate combine phase.
//   state         every CTA keeps an identical replica of lengths/finished in shared memory and
//                 applies synthesize.py:42-45 itself after the final projection; CTA 0 publishes it.
//
// Reference semantics: transformer/tacotron.py:107-116, transformer/modules.py:108-145,
// transformer/attention.py:53-122, synthesize.py:35-45.
#include <math_constants.h>
#include <string.h>

#include "common.cuh"

namespace tts {
namespace fused {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kKC = 768;       // K slice of the activation resident in shared memory
constexpr int kXLd = kKC + 4;  // row stride = 4 (mod 32) words -> conflict-free 128-bit reads 4 rows apart
constexpr int kRowBlk = 32;    // batch rows per activation tile
constexpr int kPass = 8;       // weight rows per register pass
constexpr int kMaxBatch = 1024;
constexpr long long kSpinLimit = 1LL << 22;

enum Mode { kPlain = 0, kQkv = 1, kPrenetOut = 2, kFinal = 3 };
enum XSrc { kXPlain = 0, kXFrames = 1, kXCombine = 2 };

struct Args {
  TtsDecoderWeights w;
  TtsDecodeState st;
  float *x, *q, *hid, *p0, *p1, *part;
  int max_seg;
  unsigned* bar;
  int* err;
  int n_steps, update_state;
};

struct Gemm {
  int xsrc; const float* X; long long ldx; int K, N;
  const float* W; const float* W2; int n_w1;
  const float* ln_g; const float* ln_b;
  const float* bias; int relu; int mode;
  float* Y; long long ldy; const float* R; long long ldr; float out_scale;
  float* kcache; float* vcache;
  int comb_keys;                       // kXCombine: keys of the attention phase being merged
  float* align; long long align_bh_stride; int align_row_len;  // kXCombine: rows to normalise (or null)
};

struct Smem {
  float* xs[2];   // [32][kXLd] each
  float* red;     // [8][8][32]
  float* ml;      // [32*H][2]  merged (max, 1/sum) of the row block's attention items
  int* len;       // [B]
  int* fin;       // [B]
};

// ---- grid barrier -----------------------------------------------------------------------------
struct GridBar {
  unsigned* ctr;
  int* err;
  unsigned epoch, n;
};

__device__ __forceinline__ void bar_arrive(GridBar& gb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(gb.ctr, 1u);
  }
  gb.epoch++;
}

__device__ __forceinline__ void bar_wait(GridBar& gb) {
  if (threadIdx.x == 0) {
    const unsigned target = gb.epoch * gb.n;
    if (*reinterpret_cast<volatile int*>(gb.err) == 0) {
      long long spins = 0;
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
        if (static_cast<int>(v - target) >= 0) break;
        if (++spins > kSpinLimit) {  // never hang the GPU: flag the error and fall through
          atomicExch(gb.err, 1);
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void grid_sync(GridBar& gb) {
  bar_arrive(gb);
  bar_wait(gb);
}

// ---- async copy helpers -------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- activation staging -------------------------------------------------------------------------
// rows b0..b0+31 (zero beyond B), columns k0..k0+kc of a [B][ldx] activation -> Xs (async)
__device__ __forceinline__ void stage_plain(float* Xs, const float* X, long long ldx, int B, int b0, int k0, int kc,
                                            bool zero_all) {
  const int f4 = kc >> 2;
  for (int i = threadIdx.x; i < kRowBlk * f4; i += kThreads) {
    const int r = i / f4, c = (i - r * f4) << 2;
    float* dst = Xs + r * kXLd + c;
    if (!zero_all && b0 + r < B) cp_async16(dst, X + (size_t)(b0 + r) * ldx + k0 + c);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Merge the attention partials of the previous phase into the [32][D] context tile
// (flash-decoding combine, fused into the staging of the output projection).
template <int DH>
__device__ __noinline__ void stage_combined(const Args& a, const Gemm& g, const Smem& sm, float* Xs, int b0, int t) {
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, n = g.comb_keys;
  const long long total = (long long)B * H * n;
  const long long sl = (total + G - 1) / G;
  constexpr int PS = DH + 4;  // record = o[DH], max, sum, pad (16-byte aligned)
  // 1) merged max and 1/sum per (row, head)
  for (int it = threadIdx.x; it < kRowBlk * H; it += kThreads) {
    const int b = b0 + it / H;
    float m = 0.f, inv = 0.f;
    if (b < B) {
      const int item = b * H + (it % H);
      const long long pa = (long long)item * n, pb = pa + n - 1;
      const int c0 = (int)(pa / sl), c1 = (int)(pb / sl);
      m = -CUDART_INF_F;
      for (int c = c0; c <= c1; ++c) {
        const int slot = item - (int)(((long long)c * sl) / n);
        m = fmaxf(m, __ldcg(a.part + ((size_t)c * a.max_seg + slot) * PS + DH));
      }
      float l = 0.f;
      for (int c = c0; c <= c1; ++c) {
        const int slot = item - (int)(((long long)c * sl) / n);
        const float* pr = a.part + ((size_t)c * a.max_seg + slot) * PS;
        l += __ldcg(pr + DH + 1) * expf(__ldcg(pr + DH) - m);
      }
      inv = 1.f / l;
    }
    sm.ml[2 * it] = m;
    sm.ml[2 * it + 1] = inv;
  }
  __syncthreads();
  // 2) context tile
  const int D = H * DH, f4 = D >> 2;
  for (int i = threadIdx.x; i < kRowBlk * f4; i += kThreads) {
    const int r = i / f4, col = (i - r * f4) << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int b = b0 + r;
    if (b < B) {
      const int h = col / DH, d = col - h * DH, item = b * H + h;
      const float m = sm.ml[2 * (r * H + h)], inv = sm.ml[2 * (r * H + h) + 1];
      const long long pa = (long long)item * n, pb = pa + n - 1;
      const int c0 = (int)(pa / sl), c1 = (int)(pb / sl);
      for (int c = c0; c <= c1; ++c) {
        const int slot = item - (int)(((long long)c * sl) / n);
        const float* pr = a.part + ((size_t)c * a.max_seg + slot) * PS;
        const float wgt = expf(__ldcg(pr + DH) - m);
        const float4 o = __ldcg(reinterpret_cast<const float4*>(pr + d));
        acc.x = fmaf(o.x, wgt, acc.x); acc.y = fmaf(o.y, wgt, acc.y);
        acc.z = fmaf(o.z, wgt, acc.z); acc.w = fmaf(o.w, wgt, acc.w);
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    *reinterpret_cast<float4*>(Xs + r * kXLd + col) = acc;
  }
  // 3) normalise the recorded attention rows of this step (raw logits -> softmax weights)
  if (g.align != nullptr) {
    for (int it = 0; it < kRowBlk * H; ++it) {
      const int b = b0 + it / H;
      if (b >= B) break;
      const int item = b * H + (it % H);
      if (item % G != (int)blockIdx.x) continue;
      const float m = sm.ml[2 * it], inv = sm.ml[2 * it + 1];
      float* row = g.align + (size_t)item * g.align_bh_stride + (size_t)t * g.align_row_len;
      for (int j = threadIdx.x; j < n; j += kThreads) row[j] = expf(__ldcg(row + j) - m) * inv;
    }
  }
}

// LayerNorm of the 32 staged rows, in place, 4 rows per warp, statistics from registers
__device__ __forceinline__ void ln_inplace(float* Xs, int K, const float* __restrict__ gam,
                                           const float* __restrict__ bet) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nf4 = K >> 2;
  float4 g4[6], b4[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int f = lane + 32 * j;
    if (f < nf4) {
      g4[j] = __ldg(reinterpret_cast<const float4*>(gam) + f);
      b4[j] = __ldg(reinterpret_cast<const float4*>(bet) + f);
    }
  }
  const float invK = 1.f / (float)K;
  for (int r = warp * 4; r < warp * 4 + 4; ++r) {
    float* xr = Xs + r * kXLd;
    float4 v[6];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int f = lane + 32 * j;
      if (f < nf4) {
        v[j] = *reinterpret_cast<const float4*>(xr + 4 * f);
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(s) * invK;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int f = lane + 32 * j;
      if (f < nf4) {
        const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw = v[j].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invK + 1e-6f);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int f = lane + 32 * j;
      if (f < nf4) {
        float4 o;
        o.x = (v[j].x - mean) * rstd * g4[j].x + b4[j].x;
        o.y = (v[j].y - mean) * rstd * g4[j].y + b4[j].y;
        o.z = (v[j].z - mean) * rstd * g4[j].z + b4[j].z;
        o.w = (v[j].w - mean) * rstd * g4[j].w + b4[j].w;
        *reinterpret_cast<float4*>(xr + 4 * f) = o;
      }
    }
  }
}

// acc[r][i] += sum_k W[n0+r][k0+k] * Xs[4g+i][k] over this warp's 16-float chunks of the slice
__device__ __forceinline__ void fma_slice(f32x2 (&acc)[kPass][4], const float* Xs, int kc,
                                          const float* wbase, int K, int nrows, int k0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, s4 = (lane & 3) * 4;
  const int nch = kc >> 4;
#pragma unroll 2
  for (int c = warp; c < nch; c += kWarps) {
    const int kk = c * 16 + s4;
    f32x4 wv[kPass];
#pragma unroll
    for (int r = 0; r < kPass; ++r) {
      if (r < nrows) wv[r] = ldg_stream(wbase + (size_t)r * K + k0 + kk);
      else wv[r].lo = wv[r].hi = 0ull;
    }
    f32x4 xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = lds128(Xs + (4 * g + i) * kXLd + kk);
#pragma unroll
    for (int r = 0; r < kPass; ++r)
      if (r < nrows) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[r][i] = fma2(xv[i].lo, wv[r].lo, acc[r][i]);
          acc[r][i] = fma2(xv[i].hi, wv[r].hi, acc[r][i]);
        }
      }
  }
}

template <bool LN, int DH>
__device__ __noinline__ void gemm_phase(const Args& a, const Gemm& g, const Smem& sm, int t) {
  const int G = gridDim.x, c = blockIdx.x, B = a.st.batch;
  const int n_lo = (int)(((long long)c * g.N) / G), n_hi = (int)(((long long)(c + 1) * g.N) / G);
  const bool has_rows = n_hi > n_lo;
  const bool owns_align = g.xsrc == kXCombine && g.align != nullptr;
  if (!has_rows && !owns_align) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2;
  const bool single = g.K <= kKC;
  const int n_kc = (g.K + kKC - 1) / kKC;

  for (int b0 = 0; b0 < B; b0 += kRowBlk) {
    if (single) {
      if (g.xsrc == kXCombine) {
        stage_combined<DH>(a, g, sm, sm.xs[0], b0, t);
      } else {
        const bool from_frames = g.xsrc == kXFrames;
        const float* X = from_frames ? a.st.frames + (size_t)(t > 0 ? t - 1 : 0) * a.w.n_mels : g.X;
        stage_plain(sm.xs[0], X, g.ldx, B, b0, 0, g.K, from_frames && t == 0);
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
      if (LN) {
        ln_inplace(sm.xs[0], g.K, g.ln_g, g.ln_b);
        __syncthreads();
      }
    }
    for (int n0 = n_lo, nstep = 0; n0 < n_hi; n0 += nstep) {
      int nrows = min(kPass, n_hi - n0);
      if (n0 < g.n_w1) nrows = min(nrows, g.n_w1 - n0);  // a pass never straddles the W / W2 boundary
      const float* wbase = n0 < g.n_w1 ? g.W + (size_t)n0 * g.K : g.W2 + (size_t)(n0 - g.n_w1) * g.K;
      f32x2 acc[kPass][4];
#pragma unroll
      for (int r = 0; r < kPass; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = 0ull;

      if (single) {
        fma_slice(acc, sm.xs[0], g.K, wbase, g.K, nrows, 0);
      } else {  // K > 768 (FFN-out): double-buffered K slices of the activation
        stage_plain(sm.xs[0], g.X, g.ldx, B, b0, 0, kKC, false);
        cp_async_commit();
        for (int ki = 0; ki < n_kc; ++ki) {
          const int k0 = ki * kKC;
          if (ki + 1 < n_kc) {
            stage_plain(sm.xs[(ki + 1) & 1], g.X, g.ldx, B, b0, k0 + kKC, min(kKC, g.K - k0 - kKC), false);
            cp_async_commit();
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncthreads();
          fma_slice(acc, sm.xs[ki & 1], min(kKC, g.K - k0), wbase, g.K, nrows, k0);
          __syncthreads();
        }
      }
      // reduce: packed pairs -> 4 k-split lanes -> 8 warps
#pragma unroll
      for (int r = 0; r < kPass; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = hsum2(acc[r][i]);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if ((lane & 3) == 0) sm.red[(warp * kPass + r) * 32 + 4 * gq + i] = v;
        }
      __syncthreads();
      const int r = tid >> 5, br = tid & 31, b = b0 + br;
      if (r < nrows && b < B) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += sm.red[(w * kPass + r) * 32 + br];
        const int n = n0 + r;
        if (g.bias) v += __ldg(g.bias + n);
        if (g.relu) v = fmaxf(v, 0.f);
        switch (g.mode) {
          case kPlain: {
            v *= g.out_scale;
            if (g.R) v += __ldcg(g.R + (size_t)b * g.ldr + n);
            g.Y[(size_t)b * g.ldy + n] = v;
          } break;
          case kQkv: {
            const int H = a.w.n_heads, D = H * DH;
            const int which = n / D, cc = n - which * D;
            if (which == 0) {
              g.Y[(size_t)b * g.ldy + cc] = v * g.out_scale;
            } else {
              const int h = cc / DH, d = cc - h * DH;
              float* dst = which == 1 ? g.kcache : g.vcache;
              dst[(((size_t)b * H + h) * a.st.t_max + t) * DH + d] = v;
            }
          } break;
          case kPrenetOut: {  // modules.py:114-118
            const bool have = t > 0 && (t - 1) < sm.len[b];
            g.Y[(size_t)b * g.ldy + n] = (have ? v : 0.f) + __ldg(a.w.pe_table + (size_t)t * g.N + n) * __ldg(a.w.pe_scale);
          } break;
          case kFinal: {  // modules.py:144, tacotron.py:112-115
            const bool live = t < sm.len[b];
            if (n < a.w.n_mels) a.st.frames[((size_t)b * a.st.t_max + t) * a.w.n_mels + n] = live ? v : 0.f;
            else a.st.stop_logits[(size_t)b * a.st.t_max + t] = live ? v + __ldg(a.w.b_stop) : 0.f;
          } break;
        }
      }
      __syncthreads();  // red is reused by the next pass
      nstep = nrows;
    }
  }
}

// ---- attention phase ------------------------------------------------------------------------------
template <int DH>
__device__ void attn_segment(const Args& a, const float* q, const float* kbase, const float* vbase, int j0, int j1,
                             int klen, float* part_rec, float* arow, float* sc, float* s_red, float* s_max,
                             float* s_sum) {
  constexpr int F4 = DH / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kslot = lane >> 3, l8 = lane & 7;
  f32x4 qv[F4];
#pragma unroll
  for (int i = 0; i < F4; ++i) qv[i] = ldg_cg(q + 4 * (l8 + 8 * i));

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
  for (int w = 1; w < kWarps; ++w) m = fmaxf(m, s_max[w]);

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
#pragma unroll
  for (int i = 0; i < F4; ++i)
#pragma unroll
    for (int hs = 0; hs < 2; ++hs) {
      float x, y;
      unpack2(o[i][hs], x, y);
      x += __shfl_xor_sync(0xffffffffu, x, 8);  y += __shfl_xor_sync(0xffffffffu, y, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16); y += __shfl_xor_sync(0xffffffffu, y, 16);
      if (kslot == 0) {
        const int d = 4 * (l8 + 8 * i) + 2 * hs;
        s_red[warp * (DH + 1) + d] = x;
        s_red[warp * (DH + 1) + d + 1] = y;
      }
    }
  lsum = warp_sum(lsum);
  if (lane == 0) s_sum[warp] = lsum;
  __syncthreads();
  if (tid < DH) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += s_red[w * (DH + 1) + tid];
    part_rec[tid] = v;
  }
  if (tid == 0) {
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) l += s_sum[w];
    part_rec[DH] = m;
    part_rec[DH + 1] = l;
  }
  if (arow != nullptr)
    for (int j = j0 + tid; j < j1; j += kThreads) arow[j] = sc[j - j0];  // raw logits; normalised by the consumer
  __syncthreads();  // sc / s_red are reused by the next segment
}

template <int DH>
__device__ __noinline__ void attn_phase(const Args& a, const Smem& sm, const float* kc, const float* vc, int rows_alloc,
                           int n_keys, const int32_t* key_len, float* align, long long align_bh_stride,
                           int align_row_len, int t) {
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, c = blockIdx.x;
  const long long total = (long long)B * H * n_keys;
  const long long sl = (total + G - 1) / G;
  long long p = (long long)c * sl;
  const long long p1 = min(total, p + sl);
  float* sc = sm.xs[0];           // up to sl floats (host checks it fits)
  float* s_red = sm.red;          // 8 * (DH+1)
  float* s_max = sm.red + kWarps * (DH + 1);
  float* s_sum = s_max + kWarps;
  int seg = 0;
  while (p < p1) {
    const int item = (int)(p / n_keys);
    const int j0 = (int)(p - (long long)item * n_keys);
    const int j1 = (int)min((long long)n_keys, j0 + (p1 - p));
    const int b = item / H;
    const int klen = key_len ? key_len[b] : n_keys;
    float* rec = a.part + ((size_t)c * a.max_seg + seg) * (DH + 4);
    float* arow = align ? align + (size_t)item * align_bh_stride + (size_t)t * align_row_len : nullptr;
    attn_segment<DH>(a, a.q + (size_t)item * DH, kc + (size_t)item * rows_alloc * DH,
                     vc + (size_t)item * rows_alloc * DH, j0, j1, klen, rec, arow, sc, s_red, s_max, s_sum);
    p += j1 - j0;
    ++seg;
  }
}

// ---- the kernel -------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kThreads, 1) fused_decode_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm;
  sm.xs[0] = smem_raw;
  sm.xs[1] = sm.xs[0] + kRowBlk * kXLd;
  sm.red = sm.xs[1] + kRowBlk * kXLd;
  sm.ml = sm.red + kWarps * kPass * 32;
  sm.len = reinterpret_cast<int*>(sm.ml + 2 * kRowBlk * a.w.n_heads);
  sm.fin = sm.len + a.st.batch;

  const int B = a.st.batch, D = a.w.d_model, H = a.w.n_heads, F = a.w.d_ffn, P = a.w.prenet_hidden;
  const int M = a.w.n_mels, S = a.st.mem_len, T = a.st.t_max, L = a.w.n_layers;
  GridBar gb{a.bar, a.err, 0u, gridDim.x};
  const float qscale = (float)(1.0 / sqrt((double)DH));

  for (int b = threadIdx.x; b < B; b += kThreads) {
    sm.len[b] = a.st.lengths[b];
    sm.fin[b] = a.st.finished[b];
  }
  __syncthreads();
  const int t0 = *a.st.step_counter;

  Gemm base;
  memset(&base, 0, sizeof(base));
  base.out_scale = 1.f;

  for (int s = 0; s < a.n_steps; ++s) {
    const int t = t0 + s;
    {  // prenet (tacotron.py:55-65) and shift/mask/PE (modules.py:114-118)
      Gemm g = base;
      g.xsrc = kXFrames; g.ldx = (long long)T * M; g.K = M; g.N = P; g.W = a.w.prenet_w0; g.n_w1 = P;
      g.bias = a.w.prenet_b0; g.relu = 1; g.mode = kPlain; g.Y = a.p0; g.ldy = P;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);
      g = base;
      g.X = a.p0; g.ldx = P; g.K = P; g.N = P; g.W = a.w.prenet_w1; g.n_w1 = P; g.bias = a.w.prenet_b1; g.relu = 1;
      g.mode = kPlain; g.Y = a.p1; g.ldy = P;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);
      g = base;
      g.X = a.p1; g.ldx = P; g.K = P; g.N = D; g.W = a.w.prenet_w2; g.n_w1 = D; g.mode = kPrenetOut; g.Y = a.x; g.ldy = D;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);
    }
    for (int l = 0; l < L; ++l) {
      const TtsDecLayerWeights& lw = a.w.layer[l];
      const size_t self_off = (size_t)l * B * H * T * DH, cross_off = (size_t)l * B * H * S * DH;
      Gemm g = base;  // LN + QKV (attention.py:63-64), k/v appended at row t
      g.X = a.x; g.ldx = D; g.K = D; g.N = 3 * D; g.W = lw.w_qkv; g.n_w1 = 3 * D; g.ln_g = lw.ln_self_g;
      g.ln_b = lw.ln_self_b; g.mode = kQkv; g.Y = a.q; g.ldy = D; g.out_scale = qscale;
      g.kcache = a.st.self_k + self_off; g.vcache = a.st.self_v + self_off;
      gemm_phase<true, DH>(a, g, sm, t);
      grid_sync(gb);

      float* al = a.st.align_self ? a.st.align_self + (size_t)l * B * H * T * T : nullptr;
      attn_phase<DH>(a, sm, a.st.self_k + self_off, a.st.self_v + self_off, T, t + 1, nullptr, al, (long long)T * T, T, t);
      grid_sync(gb);

      g = base;  // merge partials + output projection + residual (attention.py:118-119, modules.py:132)
      g.xsrc = kXCombine; g.comb_keys = t + 1; g.align = al; g.align_bh_stride = (long long)T * T; g.align_row_len = T;
      g.K = D; g.N = D; g.W = lw.w_self_out; g.n_w1 = D; g.mode = kPlain; g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);

      g = base;  // LN + cross query
      g.X = a.x; g.ldx = D; g.K = D; g.N = D; g.W = lw.w_cross_q; g.n_w1 = D; g.ln_g = lw.ln_cross_g;
      g.ln_b = lw.ln_cross_b; g.mode = kPlain; g.Y = a.q; g.ldy = D; g.out_scale = qscale;
      gemm_phase<true, DH>(a, g, sm, t);
      grid_sync(gb);

      al = a.st.align_cross ? a.st.align_cross + (size_t)l * B * H * T * S : nullptr;
      attn_phase<DH>(a, sm, a.st.cross_k + cross_off, a.st.cross_v + cross_off, S, S, a.st.input_lengths, al,
                     (long long)T * S, S, t);
      grid_sync(gb);

      g = base;
      g.xsrc = kXCombine; g.comb_keys = S; g.align = al; g.align_bh_stride = (long long)T * S; g.align_row_len = S;
      g.K = D; g.N = D; g.W = lw.w_cross_out; g.n_w1 = D; g.mode = kPlain; g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);

      g = base;  // LN + FFN-in + ReLU (modules.py:14-17)
      g.X = a.x; g.ldx = D; g.K = D; g.N = F; g.W = lw.w_ffn_in; g.n_w1 = F; g.ln_g = lw.ln_ffn_g; g.ln_b = lw.ln_ffn_b;
      g.relu = 1; g.mode = kPlain; g.Y = a.hid; g.ldy = F;
      gemm_phase<true, DH>(a, g, sm, t);
      grid_sync(gb);

      g = base;  // FFN-out + residual
      g.X = a.hid; g.ldx = F; g.K = F; g.N = D; g.W = lw.w_ffn_out; g.n_w1 = D; g.mode = kPlain;
      g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D;
      gemm_phase<false, DH>(a, g, sm, t);
      grid_sync(gb);
    }
    {  // final LN + mel / stop projections
      Gemm g = base;
      g.X = a.x; g.ldx = D; g.K = D; g.N = M + 1; g.W = a.w.w_mel; g.W2 = a.w.w_stop; g.n_w1 = M;
      g.ln_g = a.w.ln_out_g; g.ln_b = a.w.ln_out_b; g.mode = kFinal;
      gemm_phase<true, DH>(a, g, sm, t);
      grid_sync(gb);
    }
    // synthesize.py:42-45, replicated identically in every CTA
    int unfinished = 0;
    if (a.update_state) {
      for (int b = threadIdx.x; b < B; b += kThreads) {
        const bool fin = sm.fin[b] != 0 || __ldcg(a.st.stop_logits + (size_t)b * T + t) > 0.f;
        sm.fin[b] = fin ? 1 : 0;
        if (!fin) sm.len[b] += 1;
      }
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) unfinished += sm.fin[b] ? 0 : 1;
    if (blockIdx.x == 0) {
      if (a.update_state)
        for (int b = threadIdx.x; b < B; b += kThreads) {
          a.st.lengths[b] = sm.len[b];
          a.st.finished[b] = (uint8_t)sm.fin[b];
        }
      if (threadIdx.x == 0) {
        *a.st.step_counter = t + 1;
        *a.st.n_unfinished = (*reinterpret_cast<volatile int*>(a.err) != 0) ? -1 : unfinished;
      }
    }
    if (a.update_state && unfinished == 0) break;  // uniform: every CTA holds the same replica
  }
}

static size_t smem_bytes(const TtsDecoderWeights* w, int B) {
  return (size_t)(2 * kRowBlk * kXLd + kWarps * kPass * 32 + 2 * kRowBlk * w->n_heads) * sizeof(float) +
         (size_t)2 * B * sizeof(int) + 16;
}

struct Carve {
  float *x, *q, *hid, *p0, *p1, *part;
  unsigned* bar;
  int* err;
  int max_seg;
  size_t floats;
};

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static Carve carve(const TtsDecoderWeights* w, int B, float* base) {
  Carve c;
  const size_t D = w->d_model, F = w->d_ffn, P = w->prenet_hidden, H = w->n_heads, dh = D / H;
  const int G = base ? num_sms() : 160;  // size for the worst case when only sizing
  c.max_seg = (int)((size_t)B * H / (base ? G : 132)) + 2;
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 31) / 32 * 32;
    return p;
  };
  c.bar = reinterpret_cast<unsigned*>(take(32));
  c.err = reinterpret_cast<int*>(take(32));
  c.x = take(B * D); c.q = take(B * D); c.hid = take(B * F); c.p0 = take(B * P); c.p1 = take(B * P);
  c.part = take((size_t)G * c.max_seg * (dh + 4));
  c.floats = off;
  return c;
}

template <int DH>
static int launch(const Args& a, size_t smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(fused_decode_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 0;
    TTS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_decode_kernel<DH>, kThreads, 227 * 1024));
    TTS_REQUIRE(per_sm >= 1, "fused decode kernel does not fit on an SM");
    configured = true;
  }
  void* params[] = {const_cast<Args*>(&a)};
  TTS_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(fused_decode_kernel<DH>), dim3(num_sms()),
                                             dim3(kThreads), params, smem, s));
  count_launch();
  return 0;
}

}  // namespace fused

size_t fused_scratch_floats(const TtsDecoderWeights* w, int B) { return fused::carve(w, B, nullptr).floats; }

bool fused_supported(const TtsDecoderWeights* w, const TtsDecodeState* st) {
  using namespace fused;
  if (w->d_model > kKC || w->d_model % 128 != 0) return false;       // LayerNorm prologue tiling
  if (st->batch > kMaxBatch) return false;
  const int dh = w->d_model / w->n_heads;
  if (dh != 32 && dh != 64 && dh != 96) return false;
  const int G = num_sms();
  const long long max_keys = st->t_max > st->mem_len ? st->t_max : st->mem_len;
  const long long span = ((long long)st->batch * w->n_heads * max_keys + G - 1) / G;
  if (span > 2LL * kRowBlk * kXLd) return false;                      // scores of one span live in shared memory
  if (smem_bytes(w, st->batch) > 227 * 1024) return false;
  return true;
}

int launch_fused_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int n_steps, int update_state,
                       cudaStream_t s) {
  using namespace fused;
  TTS_REQUIRE(fused_supported(w, st), "fused decode kernel does not support this shape");
  if (n_steps == 0) return 0;
  const Carve c = carve(w, st->batch, st->scratch);
  Args a;
  memcpy(&a.w, w, sizeof(*w));
  memcpy(&a.st, st, sizeof(*st));
  a.x = c.x; a.q = c.q; a.hid = c.hid; a.p0 = c.p0; a.p1 = c.p1; a.part = c.part;
  a.max_seg = c.max_seg; a.bar = c.bar; a.err = c.err; a.n_steps = n_steps; a.update_state = update_state;
  TTS_CHECK_CUDA(cudaMemsetAsync(c.bar, 0, 256, s));  // barrier counter and error flag
  const size_t smem = smem_bytes(w, st->batch);
  switch (w->d_model / w->n_heads) {
    case 32: return launch<32>(a, smem, s);
    case 64: return launch<64>(a, smem, s);
    case 96: return launch<96>(a, smem, s);
  }
  set_error("fused decode: unsupported head_dim");
  return 2;
}

}  // namespace tts
