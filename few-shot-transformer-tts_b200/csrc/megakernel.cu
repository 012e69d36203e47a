// Fused persistent decode kernel (impl 3): ONE cooperative launch runs `n_steps` whole decode
// steps.  One CTA per SM stays resident; the phases of a step are separated
// by a software grid barrier instead of kernel launches.  All bulk data movement is done by the
// TMA engine (1-D cp.async.bulk, completion on mbarriers), so no phase has a chain of dependent
// global loads:
//
//   GEMM phases   weight rows are split evenly over the CTAs.  A CTA's slice of the weight matrix
//                 is ONE contiguous range: a single bulk copy brings it into shared memory, issued
//                 while the CTA is still waiting on the grid barrier (weights never change).  The
//                 <=32 activation rows arrive as 32 bulk row copies after the barrier; LayerNorm is
//                 applied in place from registers; products run on packed FFMA2 from shared memory.
//   attention     every warp streams its share of the K/V rows of a (sample, head) through a private
//                 ring of bulk-copied tiles (K and V tile per slot, 3 slots in flight per warp) and
//                 keeps an online-softmax state; warps are merged once per (sample, head).
//   state         every CTA keeps an identical replica of lengths/finished in shared memory and
//                 applies synthesize.py:42-45 itself after the final projection; CTA 0 publishes it.
//
// Reference semantics: transformer/tacotron.py:107-116, transformer/modules.py:108-145,
// transformer/attention.py:53-122, synthesize.py:35-45.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tts {
namespace fused {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kKC = 768;        // K slice of the activation staged in shared memory at a time
constexpr int kChunksPerWarp = kKC / 16 / kWarps;  // 16-float chunks of a slice owned by one warp (6)
constexpr int kRowBlk = 32;     // batch rows per activation tile
constexpr int kPass = 8;        // weight rows per register pass
constexpr int kXFloats = kRowBlk * (kKC + 4);           // 32 activation rows
constexpr int kWFloats = 19456;                         // 76 KB: <= 6 rows of K=3072, <= 25 rows of K=768
constexpr int kTK = 8;          // keys per K/V ring tile
constexpr int kSlots = 3;       // ring slots per warp
constexpr int kMaxBatch = 1024;
constexpr int kMaxSplit = 32;
constexpr long long kSpinLimit = 1LL << 22;
constexpr int kProfPhases = 160;

enum Mode { kPlain = 0, kQkv = 1, kPrenetOut = 2, kFinal = 3 };
enum XSrc { kXPlain = 0, kXFrames = 1, kXCombine = 2 };

struct Args {
  TtsDecoderWeights w;
  TtsDecodeState st;
  float *x, *q, *ctx, *hid, *p0, *p1, *part;
  float* stats;         // [B][G][2] per-CTA (mean, M2) of the rows of x over the CTA's columns
  int n_split;          // K/V splits per (sample, head): 1 when B*H >= #CTAs
  unsigned* bar;
  int* err;
  long long* prof;      // [kProfPhases][8] SM clock stamps of CTA 0 for the last step run (diagnostics)
  int n_steps, update_state;
  int cluster2;         // launched as 2-CTA clusters: activation tiles are TMA-multicast to both CTAs of a pair
};

struct Gemm {
  int xsrc; const float* X; long long ldx; int K, N;
  const float* W; const float* W2; int n_w1;
  const float* ln_g; const float* ln_b;
  const float* bias; int relu; int mode;
  float* Y; long long ldy; const float* R; long long ldr; float out_scale;
  float* kcache; float* vcache;
  int comb_keys;
  float* align; long long align_bh_stride; int align_row_len;
  int ln;           // input rows are normalised with the statistics published by the producing phase
  int emit_stats;   // the output is the residual stream x: also publish per-row (mean, M2) over this CTA's columns
};

struct Attn {
  const float* kc; const float* vc; int rows_alloc; int n_keys; const int32_t* key_len;
  float* align; long long align_bh_stride; int align_row_len;
};

struct Prefetch {   // slice [part/parts) of the first `rows` K/V rows of each item this CTA will attend over
  const float* kc; const float* vc; int rows_alloc, rows, part, parts;
};

struct Phase {
  int kind;  // 0 gemm, 1 gemm with LayerNorm prologue, 2 attention
  Gemm g;
  Attn at;
  Prefetch pre;
};

struct Smem {
  float* xs;      // activation tile(s)
  float* wb;      // weight slice
  float* ring;    // K/V rings (aliases xs + wb during attention phases)
  float* red;     // [8][8][32] GEMM cross-warp reduction / attention warp records
  float* ml;      // [32*H][2]
  float* stat;    // [32][2]  (rstd, -mean*rstd) of the staged rows
  float* sstat;   // [8][32]  epilogue values for the per-row statistics
  int* len;       // [B]
  int* fin;       // [B]
  uint64_t* wfull;     // 1
  uint64_t* xfull;     // 2
  uint64_t* rfull;     // [8][kSlots]
};

struct Track {        // per-call working state of a phase (registers; never passed by reference)
  unsigned w_par;              // parity of the weight barrier for this phase
  unsigned issued, consumed;   // K/V ring tiles requested / used by this warp since the kernel started
  long long* prof;             // stamp row of the current phase (CTA 0, thread 0) or null
};

__device__ __forceinline__ void stamp(const Track& tk, int k) {
  if (tk.prof != nullptr) tk.prof[k] = clock64();
}

__device__ __forceinline__ Smem make_smem(const Args& a, float* base) {
  Smem sm;
  sm.xs = base;
  sm.wb = sm.xs + kXFloats;
  sm.ring = base;
  sm.red = sm.wb + kWFloats;
  sm.ml = sm.red + kWarps * kPass * 32;
  sm.stat = sm.ml + 2 * kRowBlk * a.w.n_heads;
  sm.sstat = sm.stat + 2 * kRowBlk;
  sm.wfull = reinterpret_cast<uint64_t*>(sm.sstat + kPass * 32);
  sm.xfull = sm.wfull + 1;
  sm.rfull = sm.wfull + 3;
  sm.len = reinterpret_cast<int*>(sm.rfull + kWarps * kSlots);
  sm.fin = sm.len + a.st.batch;
  return sm;
}

// ---- mbarrier / bulk copy primitives ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, int* err) {
  long long spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    ++spins;
    if ((spins & 4095) == 0 && (spins > kSpinLimit || *reinterpret_cast<volatile int*>(err) != 0)) {
      atomicExch(err, 2);  // never hang the GPU
      break;
    }
  }
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(float* dst, const float* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// same, delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in `mask`
__device__ __forceinline__ void bulk_g2s_mc(float* dst, const float* src, unsigned bytes, uint64_t* bar, unsigned short mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 128-bit loads the compiler is free to hoist and batch (the asm-volatile helpers of common.cuh pin the
// program order, which exposed the full shared-memory latency per weight row)
__device__ __forceinline__ f32x4 ld4s(const float* p) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f32x4 r;
  r.lo = pack2(v.x, v.y);
  r.hi = pack2(v.z, v.w);
  return r;
}
__device__ __forceinline__ f32x4 ld4cg(const float* p) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  f32x4 r;
  r.lo = pack2(v.x, v.y);
  r.hi = pack2(v.z, v.w);
  return r;
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- grid barrier -----------------------------------------------------------------------------
struct GridBar {
  unsigned* ctr;
  int* err;
  unsigned epoch, n;
};

__device__ __forceinline__ void bar_arrive(GridBar& gb, bool async_readers) {
  // global writes of this phase that another CTA will read through the async proxy (TMA bulk copies of the
  // freshly appended K/V row) need a cross-proxy fence; cp.async / ld readers do not
  if (async_readers) fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
  gb.epoch++;
}

__device__ __forceinline__ void bar_wait(GridBar& gb) {
  if (threadIdx.x == 0) {
    const unsigned target = gb.epoch * gb.n;
    long long spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
      ++spins;
      if ((spins & 1023) == 0 && (spins > kSpinLimit || *reinterpret_cast<volatile int*>(gb.err) != 0)) {
        atomicExch(gb.err, 1);
        break;
      }
    }
  }
  __syncthreads();
}

// Ask the memory system to pull K/V rows this CTA will stream in an upcoming attention phase into L2 while the
// current (FMA-bound) GEMM phases leave HBM idle.  Pure hint: correctness never depends on it.
__device__ __forceinline__ void prefetch_l2(const float* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
template <int DH>
__device__ __forceinline__ void prefetch_kv(const Args& a, const Prefetch& pf) {
  if (pf.kc == nullptr || a.n_split != 1 || pf.rows <= 0) return;
  const int n_items = a.st.batch * a.w.n_heads, G = gridDim.x;
  // keep the total under ~80 MB so that it survives in the 126 MB L2 until it is used
  const long long cap = (80ll << 20) / ((long long)n_items * DH * 8);
  const int r = (int)min((long long)pf.rows, cap > 1 ? cap : 1);
  const int r0 = (int)(((long long)r * pf.part) / pf.parts), r1 = (int)(((long long)r * (pf.part + 1)) / pf.parts);
  const int i = (int)threadIdx.x - 32;  // warp 1, lanes 0..7: (unit, K|V)
  if (i < 0 || i >= 8 || r1 <= r0) return;
  const int item = blockIdx.x + G * (i >> 1);
  if (item >= n_items) return;
  const float* base = ((i & 1) ? pf.vc : pf.kc) + ((size_t)item * pf.rows_alloc + r0) * DH;
  prefetch_l2(base, (unsigned)(r1 - r0) * DH * 4u);
}

// ---- GEMM phase -------------------------------------------------------------------------------------
__device__ __forceinline__ void slice_rows(const Gemm& g, int& n_lo, int& n_hi) {
  const int G = gridDim.x, c = blockIdx.x;
  n_lo = (int)(((long long)c * g.N) / G);
  n_hi = (int)(((long long)(c + 1) * g.N) / G);
}

// one thread: bring this CTA's weight rows [n_lo, n_hi) x K into sm.wb (row r at wb + r*K)
__device__ __forceinline__ void issue_weights(const Gemm& g, const Smem& sm) {
  int n_lo, n_hi;
  slice_rows(g, n_lo, n_hi);
  if (n_hi <= n_lo) return;
  const unsigned row_bytes = (unsigned)g.K * 4u;
  mbar_expect_tx(sm.wfull, (unsigned)(n_hi - n_lo) * row_bytes);
  const int a_hi = min(n_hi, g.n_w1);
  if (n_lo < a_hi) bulk_g2s(sm.wb, g.W + (size_t)n_lo * g.K, (unsigned)(a_hi - n_lo) * row_bytes, sm.wfull);
  if (n_hi > g.n_w1) {
    const int b_lo = max(n_lo, g.n_w1);
    bulk_g2s(sm.wb + (size_t)(b_lo - n_lo) * g.K, g.W2 + (size_t)(b_lo - g.n_w1) * g.K,
             (unsigned)(n_hi - b_lo) * row_bytes, sm.wfull);
  }
}

// rows b0..b0+31 (zero beyond B), columns k0..k0+kc of a [B][ldx] activation -> Xs (row stride ld floats),
// as 16-byte cp.async requests that are all in flight at once (one L2 latency for the whole tile).
// 8 consecutive threads cover 128 contiguous bytes of one row; no index division in the loop.
__device__ __forceinline__ void stage_rows(float* Xs, int ld, const float* X, long long ldx, int B, int b0, int k0,
                                           int kc, bool zero_all) {
  constexpr int TPR = kThreads / kRowBlk;  // threads per row (12)
  const int r = threadIdx.x / TPR, cg = threadIdx.x - r * TPR;
  const int nvalid = zero_all ? 0 : min(kRowBlk, B - b0);
  float* dst = Xs + r * ld + 4 * cg;
  if (r < nvalid) {
    const float* src = X + (size_t)(b0 + r) * ldx + k0 + 4 * cg;
    for (int c = 4 * cg; c < kc; c += 4 * TPR, dst += 4 * TPR, src += 4 * TPR) cp_async16(dst, src);
  } else {
    for (int c = 4 * cg; c < kc; c += 4 * TPR, dst += 4 * TPR)
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Cluster variant: the two CTAs of a pair need the same tile, so each fetches half of the rows ONCE from L2 and
// the TMA engine multicasts them into both CTAs' shared memory (halves the L2 broadcast traffic of a GEMM phase).
// Completion is counted on each CTA's own `bar`, which expects the whole tile.
__device__ __forceinline__ void stage_rows_mc(float* Xs, int ld, const float* X, long long ldx, int B, int b0, int k0,
                                              int kc, uint64_t* bar, bool zero_all) {
  const int nvalid = zero_all ? 0 : min(kRowBlk, B - b0);
  if (nvalid < kRowBlk) {
    const int f4 = kc >> 2;
    for (int i = threadIdx.x; i < (kRowBlk - nvalid) * f4; i += kThreads) {
      const int r = nvalid + i / f4, c = (i % f4) << 2;
      *reinterpret_cast<float4*>(Xs + r * ld + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, (unsigned)nvalid * (unsigned)kc * 4u);
    __syncwarp();
    const int r = (int)cluster_ctarank() * (kRowBlk / 2) + (int)threadIdx.x;  // lanes 0..15: this CTA's half
    if (threadIdx.x < kRowBlk / 2 && r < nvalid)
      bulk_g2s_mc(Xs + r * ld, X + (size_t)(b0 + r) * ldx + k0, (unsigned)kc * 4u, bar, (unsigned short)0x3);
  }
}

// Merge split-K/V attention partials into the [32][D] context tile (only when n_split > 1, i.e. small batches)
template <int DH>
__device__ __noinline__ void stage_combined(const Args& a, const Gemm& g, const Smem& sm, float* Xs, int ld, int b0,
                                            int t) {
  const int B = a.st.batch, H = a.w.n_heads, ns = a.n_split, n = g.comb_keys;
  constexpr int PS = DH + 4;
  for (int it = threadIdx.x; it < kRowBlk * H; it += kThreads) {
    const int b = b0 + it / H;
    float m = 0.f, inv = 0.f;
    if (b < B) {
      const float* pr = a.part + (size_t)(b * H + (it % H)) * ns * PS;
      m = -CUDART_INF_F;
      for (int s = 0; s < ns; ++s) m = fmaxf(m, __ldcg(pr + s * PS + DH));
      float l = 0.f;
      for (int s = 0; s < ns; ++s) {
        const float pm = __ldcg(pr + s * PS + DH);
        if (pm > -CUDART_INF_F) l += __ldcg(pr + s * PS + DH + 1) * expf(pm - m);
      }
      inv = 1.f / l;
    }
    sm.ml[2 * it] = m;
    sm.ml[2 * it + 1] = inv;
  }
  __syncthreads();
  const int D = H * DH, f4 = D >> 2;
  for (int i = threadIdx.x; i < kRowBlk * f4; i += kThreads) {
    const int r = i / f4, col = (i - r * f4) << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int b = b0 + r;
    if (b < B) {
      const int h = col / DH, d = col - h * DH;
      const float m = sm.ml[2 * (r * H + h)], inv = sm.ml[2 * (r * H + h) + 1];
      const float* pr = a.part + (size_t)(b * H + h) * ns * PS;
      for (int s = 0; s < ns; ++s) {
        const float pm = __ldcg(pr + s * PS + DH);
        if (pm > -CUDART_INF_F) {
          const float wgt = expf(pm - m);
          const float4 o = __ldcg(reinterpret_cast<const float4*>(pr + s * PS + d));
          acc.x = fmaf(o.x, wgt, acc.x); acc.y = fmaf(o.y, wgt, acc.y);
          acc.z = fmaf(o.z, wgt, acc.z); acc.w = fmaf(o.w, wgt, acc.w);
        }
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    *reinterpret_cast<float4*>(Xs + r * ld + col) = acc;
  }
  if (g.align != nullptr) {  // raw logits of this step -> softmax weights
    const int G = gridDim.x;
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

// LayerNorm without a pass over the tile.  gamma / beta are folded into the packed weights (W_ln, c_ln); the
// phase that produced x published, per CTA, the (mean, M2) of every row over that CTA's columns.  Here warp w
// merges the G records of rows 4w..4w+3 (Chan's parallel variance) into (rstd, -mean * rstd); the
// normalisation itself is applied when the fragments are loaded into registers.
__device__ __forceinline__ void row_stats(const Args& a, const Smem& sm, int b0, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, G = gridDim.x, B = a.st.batch;
  constexpr int E = 5;                      // records per lane (G <= 160)
  constexpr int RW = kRowBlk / kWarps;      // rows per warp (4)
  float cnt[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {             // columns of x owned by CTA c = lane + 32 e (32-bit math on purpose)
    const unsigned c = lane + 32 * e;
    cnt[e] = c < (unsigned)G ? (float)(((c + 1u) * (unsigned)D) / (unsigned)G - (c * (unsigned)D) / (unsigned)G) : 0.f;
  }
  float2 rec[RW][E];
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {         // all loads first: one L2 round trip for the whole warp
    const int b = b0 + warp * RW + rr;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      rec[rr][e] = make_float2(0.f, 0.f);
      if (b < B && cnt[e] > 0.f)
        rec[rr][e] = __ldcg(reinterpret_cast<const float2*>(a.stats) + (size_t)b * G + lane + 32 * e);
    }
  }
  float mean[RW], m2[RW];
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s += cnt[e] * rec[rr][e].x;
    mean[rr] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) mean[rr] += __shfl_xor_sync(0xffffffffu, mean[rr], o);
  const float invD = 1.f / (float)D;
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
    mean[rr] *= invD;
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float d = rec[rr][e].x - mean[rr];
      q += rec[rr][e].y + cnt[e] * d * d;
    }
    m2[rr] = q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) m2[rr] += __shfl_xor_sync(0xffffffffu, m2[rr], o);
  if (lane == 0) {
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      const int r = warp * RW + rr;
      const float rstd = b0 + r < B ? rsqrtf(m2[rr] * invD + 1e-6f) : 0.f;
      sm.stat[2 * r] = rstd;
      sm.stat[2 * r + 1] = -mean[rr] * rstd;
    }
  }
}

// X-stationary product.  A warp owns the 16-float chunks {warp, warp+8, ...} of a <=768-wide slice; a lane
// (g = lane/4, s = lane%4) keeps batch rows 4g..4g+3, floats 4s..4s+3 of each of its chunks in registers for
// the whole phase, so the inner loop only reads weights (one 64-byte wavefront per weight row and chunk).
struct XFrag {
  f32x4 v[kChunksPerWarp][4];
};
// `stat` holds (scale, shift) per row: (rstd, -mean * rstd) for a LayerNorm-folded phase, (1, 0) otherwise
// (a runtime choice keeps a single copy of this code: the kernel is instruction-fetch sensitive).
__device__ __forceinline__ void load_xfrag(XFrag& xf, const float* Xs, int ld, int kc, const float* stat) {
  constexpr bool NORM = true;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, s4 = (lane & 3) * 4, nch = kc >> 4;
  f32x2 sc[4], sh[4];
  {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 st = *reinterpret_cast<const float2*>(stat + 2 * (4 * g + i));
      sc[i] = pack2(st.x, st.x);
      sh[i] = pack2(st.y, st.y);
    }
  }
#pragma unroll
  for (int j = 0; j < kChunksPerWarp; ++j) {
    const int c = warp + kWarps * j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c < nch) {
        xf.v[j][i] = ld4s(Xs + (4 * g + i) * ld + c * 16 + s4);
        if (NORM) {  // (x - mean) * rstd
          xf.v[j][i].lo = fma2(xf.v[j][i].lo, sc[i], sh[i]);
          xf.v[j][i].hi = fma2(xf.v[j][i].hi, sc[i], sh[i]);
        }
      } else {
        xf.v[j][i].lo = xf.v[j][i].hi = 0ull;
      }
    }
  }
}
// acc[r][i] += sum over this warp's chunks of Wb[r][k0 + ...] * X; rows r >= nrows alias the last valid row
// (their accumulators are never stored) so the loop is branch-free; chunks beyond the slice multiply zeros.
__device__ __forceinline__ void fma_rows(f32x2 (&acc)[kPass][4], const XFrag& xf, int kc, const float* wb, int K,
                                         int nrows, int k0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s4 = (lane & 3) * 4, nch = kc >> 4;
  int coff[kChunksPerWarp];
#pragma unroll
  for (int j = 0; j < kChunksPerWarp; ++j) coff[j] = min(warp + kWarps * j, nch - 1) * 16 + k0 + s4;
#pragma unroll
  for (int r = 0; r < kPass; ++r) {
    const float* wr = wb + (size_t)min(r, nrows - 1) * K;
#pragma unroll
    for (int j = 0; j < kChunksPerWarp; ++j) {
      const f32x4 wv = ld4s(wr + coff[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[r][i] = fma2(xf.v[j][i].lo, wv.lo, acc[r][i]);
        acc[r][i] = fma2(xf.v[j][i].hi, wv.hi, acc[r][i]);
      }
    }
  }
}

template <int DH>
__device__ __forceinline__ void gemm_phase(const Args& a, const Gemm& g, const Prefetch& pf, float* smem_base,
                                           unsigned w_par, unsigned x_par, long long* prof, int t) {
  const Smem sm = make_smem(a, smem_base);
  Track tk{w_par, 0u, 0u, prof};
  const int B = a.st.batch;
  int n_lo, n_hi;
  slice_rows(g, n_lo, n_hi);
  const bool has_rows = n_hi > n_lo;
  const bool owns_align = g.xsrc == kXCombine && g.align != nullptr;
  if (!has_rows && !owns_align) {
    prefetch_kv<DH>(a, pf);
    return;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool LN = g.ln != 0;
  const int ld = min(g.K, kKC) + 4;
  const int n_kc = (g.K + kKC - 1) / kKC;
  bool w_ready = !has_rows;
  bool mc = false;

  for (int b0 = 0; b0 < B; b0 += kRowBlk) {
    if (n_kc == 1) {  // the whole K fits one tile: stage (+ LayerNorm) once, every pass re-reads it
      if (g.xsrc == kXCombine) {
        stage_combined<DH>(a, g, sm, sm.xs, ld, b0, t);
      } else {
        const bool from_frames = g.xsrc == kXFrames;
        const float* X = from_frames ? a.st.frames + (size_t)(t > 0 ? t - 1 : 0) * a.w.n_mels : g.X;
        if (a.cluster2 && g.N >= (int)gridDim.x && B <= kRowBlk) {
          stage_rows_mc(sm.xs, ld, X, g.ldx, B, b0, 0, g.K, &sm.xfull[0], from_frames && t == 0);
          mc = true;
        } else {
          stage_rows(sm.xs, ld, X, g.ldx, B, b0, 0, g.K, from_frames && t == 0);
        }
      }
      if (LN) {
        row_stats(a, sm, b0, g.K);  // overlaps the flight of the tile
      } else if (tid < kRowBlk) {
        sm.stat[2 * tid] = 1.f;
        sm.stat[2 * tid + 1] = 0.f;
      }
      if (mc) mbar_wait(&sm.xfull[0], x_par, a.err);
      else cp_async_wait_all();
      __syncthreads();
      stamp(tk, 3);
      if (b0 == 0) prefetch_kv<DH>(a, pf);  // HBM is idle while the products run
      stamp(tk, 4);
    }
    if (!has_rows) continue;
    if (!w_ready) {
      mbar_wait(sm.wfull, tk.w_par, a.err);
      w_ready = true;
      stamp(tk, 5);
    }
    for (int n0 = n_lo; n0 < n_hi; n0 += kPass) {
      const int nrows = min(kPass, n_hi - n0);
      const float* wb = sm.wb + (size_t)(n0 - n_lo) * g.K;
      // operands of the epilogue are requested now so that their latency hides behind the products
      const int er = tid >> 5, eb = b0 + (tid & 31);
      const bool e_on = er < nrows && eb < B;
      float e_res = 0.f, e_bias = 0.f;
      if (e_on && g.R) e_res = __ldcg(g.R + (size_t)eb * g.ldr + n0 + er);
      if (e_on && g.bias) e_bias = __ldg(g.bias + n0 + er);
      f32x2 acc[kPass][4];
#pragma unroll
      for (int r = 0; r < kPass; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = 0ull;
      {
        XFrag xf;  // activation fragments live only while the products run
        const bool mcs = a.cluster2 && g.N >= (int)gridDim.x && B <= kRowBlk;  // multicast the slices to the CTA pair
        unsigned xp = x_par;
        if (n_kc > 1) {  // K > 768 (FFN-out): slices stream through the tile, starting here
          if (b0 == 0 && n0 == n_lo) prefetch_kv<DH>(a, pf);
          if (tid < kRowBlk) {
            sm.stat[2 * tid] = 1.f;
            sm.stat[2 * tid + 1] = 0.f;
          }
          if (mcs) stage_rows_mc(sm.xs, ld, g.X, g.ldx, B, b0, 0, kKC, &sm.xfull[0], false);
          else stage_rows(sm.xs, ld, g.X, g.ldx, B, b0, 0, kKC, false);
        }
        for (int ki = 0; ki < n_kc; ++ki) {
          const int k0 = ki * kKC, kc = min(kKC, g.K - k0);
          if (n_kc > 1) {
            if (mcs) {
              mbar_wait(&sm.xfull[0], xp, a.err);
              xp ^= 1u;
            } else {
              cp_async_wait_all();
            }
            __syncthreads();
          }
          load_xfrag(xf, sm.xs, ld, kc, sm.stat);
          if (n_kc > 1) {
            // every warp (of both CTAs when multicasting) holds its fragments: the tile may be refilled while
            // they are multiplied
            if (mcs) cluster_sync_all();
            else __syncthreads();
            if (ki + 1 < n_kc) {
              const int kn = min(kKC, g.K - k0 - kKC);
              if (mcs) stage_rows_mc(sm.xs, ld, g.X, g.ldx, B, b0, k0 + kKC, kn, &sm.xfull[0], false);
              else stage_rows(sm.xs, ld, g.X, g.ldx, B, b0, k0 + kKC, kn, false);
            }
          }
          fma_rows(acc, xf, kc, wb, g.K, nrows, k0);
        }
      }
      {  // packed pairs -> reduce-scatter over the 4 k-split lanes (each ends up owning 2 weight rows x 4
         // batch rows) -> one 128-bit store per weight row into the cross-warp buffer
        const int ks = lane & 3, gq = lane >> 2;
        float v32[kPass * 4];
#pragma unroll
        for (int r = 0; r < kPass; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) v32[r * 4 + i] = hsum2(acc[r][i]);
        float v16[16], v8[8];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float keep = (ks & 1) ? v32[j + 16] : v32[j], send = (ks & 1) ? v32[j] : v32[j + 16];
          v16[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float keep = (ks & 2) ? v16[j + 8] : v16[j], send = (ks & 2) ? v16[j] : v16[j + 8];
          v8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        const int r0 = ((ks & 1) ? 4 : 0) + ((ks & 2) ? 2 : 0);  // first of the two weight rows this lane owns
        float* dst = sm.red + (warp * kPass + r0) * 32 + 4 * gq;
        *reinterpret_cast<float4*>(dst) = make_float4(v8[0], v8[1], v8[2], v8[3]);
        *reinterpret_cast<float4*>(dst + 32) = make_float4(v8[4], v8[5], v8[6], v8[7]);
      }
      __syncthreads();
      stamp(tk, 6);
      const int r = er, br = tid & 31, b = eb;
      if (e_on) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += sm.red[(w * kPass + r) * 32 + br];
        const int n = n0 + r;
        v += e_bias;
        if (g.relu) v = fmaxf(v, 0.f);
        switch (g.mode) {
          case kPlain: {
            v = v * g.out_scale + e_res;
            g.Y[(size_t)b * g.ldy + n] = v;
            if (g.emit_stats) sm.sstat[r * 32 + br] = v;
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
            v = (have ? v : 0.f) + __ldg(a.w.pe_table + (size_t)t * g.N + n) * __ldg(a.w.pe_scale);
            g.Y[(size_t)b * g.ldy + n] = v;
            if (g.emit_stats) sm.sstat[r * 32 + br] = v;
          } break;
          case kFinal: {  // modules.py:144, tacotron.py:112-115
            const bool live = t < sm.len[b];
            if (n < a.w.n_mels) a.st.frames[((size_t)b * a.st.t_max + t) * a.w.n_mels + n] = live ? v : 0.f;
            else a.st.stop_logits[(size_t)b * a.st.t_max + t] = live ? v + __ldg(a.w.b_stop) : 0.f;
          } break;
        }
      }
      __syncthreads();  // red is reused by the next pass
      if (g.emit_stats && tid < 32 && b0 + tid < B) {  // (mean, M2) of row b over this CTA's columns (one pass: N <= 8 G)
        float m = 0.f;
        for (int rr = 0; rr < nrows; ++rr) m += sm.sstat[rr * 32 + tid];
        m /= (float)nrows;
        float m2 = 0.f;
        for (int rr = 0; rr < nrows; ++rr) {
          const float d = sm.sstat[rr * 32 + tid] - m;
          m2 += d * d;
        }
        reinterpret_cast<float2*>(a.stats)[(size_t)(b0 + tid) * gridDim.x + blockIdx.x] = make_float2(m, m2);
      }
      stamp(tk, 7);
    }
  }
}

// ---- attention phase ------------------------------------------------------------------------------
template <int DH>
__device__ __forceinline__ unsigned attn_phase(const Args& a, const Attn& at, float* smem_base, unsigned ring_count,
                                            long long* prof, int t) {
  const Smem sm = make_smem(a, smem_base);
  Track tk{0u, ring_count, ring_count, prof};
  constexpr int F4 = DH / 32;            // float4 per lane per key row (8 lanes span a row)
  constexpr int kTile = kTK * DH;        // floats per K (or V) tile
  constexpr int kSlotF = 2 * kTile;      // K tile then V tile
  constexpr int kRounds = kTK / 4;       // 4 key slots per warp pass
  constexpr int PS = DH + 4;
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, ns = a.n_split;
  const int n_units = B * H * ns, n_keys = at.n_keys;
  const int per = (n_keys + ns - 1) / ns;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kslot = lane >> 3, l8 = lane & 7;
  float* ring = sm.ring + (size_t)warp * kSlots * kSlotF;
  uint64_t* full = sm.rfull + warp * kSlots;
  float* wrec = sm.red;  // [8][PS]
  float* sc = sm.ring + (size_t)kWarps * kSlots * kSlotF;            // raw logits of the current unit
  constexpr int kScCap = kXFloats + kWFloats - kWarps * kSlots * kSlotF;
  const bool sc_ok = per <= kScCap;                                   // else fall back to read-modify-write in HBM

  // ---- producer cursor (next tile this warp will request) ----
  int pu = blockIdx.x, pi = warp;
  auto issue_next = [&]() {
    int item = 0, j0 = 0, j1 = 0;
    while (pu < n_units) {
      item = ns == 1 ? pu : pu / ns;
      j0 = min(n_keys, (pu - item * ns) * per);
      j1 = min(n_keys, j0 + per);
      if (pi * kTK < j1 - j0) break;
      pu += G;
      pi = warp;
    }
    if (pu >= n_units) return;
    const int key0 = j0 + pi * kTK, nk = min(kTK, j1 - key0);
    const int slot = tk.issued % kSlots;
    if (lane == 0) {
      const unsigned bytes = (unsigned)nk * DH * 4u;
      mbar_expect_tx(&full[slot], 2u * bytes);
      const size_t off = ((size_t)item * at.rows_alloc + key0) * DH;
      bulk_g2s(ring + slot * kSlotF, at.kc + off, bytes, &full[slot]);
      bulk_g2s(ring + slot * kSlotF + kTile, at.vc + off, bytes, &full[slot]);
    }
    tk.issued++;
    pi += kWarps;
  };
#pragma unroll
  for (int s = 0; s < kSlots; ++s) issue_next();

  f32x4 qnext[F4];
#pragma unroll
  for (int i = 0; i < F4; ++i) qnext[i].lo = qnext[i].hi = 0ull;
  for (int u = blockIdx.x; u < n_units; u += G) {
    const int item = ns == 1 ? u : u / ns, split = u - item * ns;
    const int j0 = min(n_keys, split * per), j1 = min(n_keys, j0 + per);
    const int b = item / H;
    const int klen = at.key_len ? at.key_len[b] : n_keys;
    float* arow = at.align ? at.align + (size_t)item * at.align_bh_stride + (size_t)t * at.align_row_len : nullptr;

    f32x4 qv[F4];
    if (u == (int)blockIdx.x) {
#pragma unroll
      for (int i = 0; i < F4; ++i) qv[i] = ld4cg(a.q + (size_t)item * DH + 4 * (l8 + 8 * i));
    } else {
#pragma unroll
      for (int i = 0; i < F4; ++i) qv[i] = qnext[i];
    }
    if (u + G < n_units) {  // latency of the next unit's query hides behind this unit's stream
      const int nitem = ns == 1 ? u + G : (u + G) / ns;
#pragma unroll
      for (int i = 0; i < F4; ++i) qnext[i] = ld4cg(a.q + (size_t)nitem * DH + 4 * (l8 + 8 * i));
    }
    float m_run = -CUDART_INF_F, l_run = 0.f;
    f32x2 o[F4][2];
#pragma unroll
    for (int i = 0; i < F4; ++i) o[i][0] = o[i][1] = 0ull;

    const int n_tiles = (j1 - j0 + kTK - 1) / kTK;
    for (int ti = warp; ti < n_tiles; ti += kWarps) {
      const int key0 = j0 + ti * kTK, nk = min(kTK, j1 - key0);
      const int slot = tk.consumed % kSlots;
      mbar_wait(&full[slot], (tk.consumed / kSlots) & 1u, a.err);
      const float* kt = ring + slot * kSlotF;
      const float* vt = kt + kTile;
      float s[kRounds];
      float mt = -CUDART_INF_F;
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int kl = r * 4 + kslot;
        f32x2 acc = 0ull;
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const f32x4 kv = ld4s(kt + kl * DH + 4 * (l8 + 8 * i));
          acc = fma2(qv[i].lo, kv.lo, acc);
          acc = fma2(qv[i].hi, kv.hi, acc);
        }
        float v = hsum2(acc);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (kl >= nk) v = -CUDART_INF_F;                 // stale smem beyond the tile's keys
        else if (key0 + kl >= klen) v = kNegBias;        // logits + (-1e20), attention.py:84-85
        if (kl < nk && l8 == 0 && arow != nullptr) {  // raw logit, normalised once the unit's (max, sum) is known
          if (ns == 1 && sc_ok) sc[key0 + kl - j0] = v;
          else arow[key0 + kl] = v;
        }
        s[r] = v;
        mt = fmaxf(mt, v);
      }
      mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
      mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
      const float m_new = fmaxf(m_run, mt);
      const float corr = expf(m_run - m_new);
      l_run *= corr;
      const f32x2 c2 = pack2(corr, corr);
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        o[i][0] = fma2(o[i][0], c2, 0ull);
        o[i][1] = fma2(o[i][1], c2, 0ull);
      }
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int kl = r * 4 + kslot;
        const float p = kl < nk ? expf(s[r] - m_new) : 0.f;
        if (l8 == 0) l_run += p;
        const f32x2 pp = pack2(p, p);
#pragma unroll
        for (int i = 0; i < F4; ++i) {
          const f32x4 vv = ld4s(vt + kl * DH + 4 * (l8 + 8 * i));
          o[i][0] = fma2(pp, kl < nk ? vv.lo : 0ull, o[i][0]);
          o[i][1] = fma2(pp, kl < nk ? vv.hi : 0ull, o[i][1]);
        }
      }
      m_run = m_new;
      tk.consumed++;
      __syncwarp();   // every lane is done with this slot before it is refilled
      issue_next();
    }
    // ---- warp record (max, sum, weighted V) -> shared ----
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
          wrec[warp * PS + d] = x;
          wrec[warp * PS + d + 1] = y;
        }
      }
    const float lw = warp_sum(l_run);
    if (lane == 0) {
      wrec[warp * PS + DH] = m_run;
      wrec[warp * PS + DH + 1] = lw;
    }
    __syncthreads();
    float m = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) m = fmaxf(m, wrec[w * PS + DH]);
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const float mw = wrec[w * PS + DH];
      if (mw > -CUDART_INF_F) l += wrec[w * PS + DH + 1] * expf(mw - m);
    }
    if (tid < DH) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const float mw = wrec[w * PS + DH];
        if (mw > -CUDART_INF_F) v += wrec[w * PS + tid] * expf(mw - m);
      }
      if (ns == 1) a.ctx[(size_t)item * DH + tid] = v / l;
      else a.part[(size_t)u * PS + tid] = v;
    }
    if (ns == 1) {
      if (arow != nullptr) {
        const float inv = 1.f / l;
        if (sc_ok) for (int j = j0 + tid; j < j1; j += kThreads) arow[j] = expf(sc[j - j0] - m) * inv;
        else for (int j = j0 + tid; j < j1; j += kThreads) arow[j] = expf(arow[j] - m) * inv;
      }
    } else if (tid == 0) {
      a.part[(size_t)u * PS + DH] = m;       // -inf when the split is empty
      a.part[(size_t)u * PS + DH + 1] = l;
    }
    __syncthreads();  // wrec is reused by the next unit
  }
  return tk.consumed;
}

// ---- phase table --------------------------------------------------------------------------------------
template <int DH>
__device__ __forceinline__ void get_phase(const Args& a, int ph, int t, float qscale, Phase& p) {
  const int B = a.st.batch, D = a.w.d_model, H = a.w.n_heads, F = a.w.d_ffn, P = a.w.prenet_hidden;
  const int M = a.w.n_mels, S = a.st.mem_len, T = a.st.t_max, L = a.w.n_layers;
  Gemm& g = p.g;
  memset(&g, 0, sizeof(g));
  g.out_scale = 1.f;
  p.kind = 0;
  p.pre.kc = nullptr;
  {  // which K/V stream to pull into L2 behind this phase's products (self: the t rows written so far)
    const int k = ph >= 3 ? (ph - 3) % 8 : -1, l = ph >= 3 ? (ph - 3) / 8 : -1;
    int self_layer = -1, part = 0;
    if (ph < 3) { self_layer = 0; part = ph; }
    else if (ph < 3 + 8 * L && k == 0) { self_layer = l; part = 3; }
    else if (ph < 3 + 8 * L && k >= 5 && l + 1 < L) { self_layer = l + 1; part = k - 5; }
    if (self_layer >= 0) {
      const size_t off = (size_t)self_layer * B * H * T * DH;
      p.pre.kc = a.st.self_k + off; p.pre.vc = a.st.self_v + off; p.pre.rows_alloc = T; p.pre.rows = t;
      p.pre.part = part; p.pre.parts = 4;
    } else if (ph >= 3 && ph < 3 + 8 * L && (k == 2 || k == 3)) {
      const size_t off = (size_t)l * B * H * S * DH;
      p.pre.kc = a.st.cross_k + off; p.pre.vc = a.st.cross_v + off; p.pre.rows_alloc = S; p.pre.rows = S;
      p.pre.part = k - 2; p.pre.parts = 2;
    }
  }
  if (ph == 0) {         // prenet (tacotron.py:55-65)
    g.xsrc = kXFrames; g.ldx = (long long)T * M; g.K = M; g.N = P; g.W = a.w.prenet_w0; g.n_w1 = P;
    g.bias = a.w.prenet_b0; g.relu = 1; g.mode = kPlain; g.Y = a.p0; g.ldy = P;
  } else if (ph == 1) {
    g.X = a.p0; g.ldx = P; g.K = P; g.N = P; g.W = a.w.prenet_w1; g.n_w1 = P; g.bias = a.w.prenet_b1; g.relu = 1;
    g.mode = kPlain; g.Y = a.p1; g.ldy = P;
  } else if (ph == 2) {  // + shift / mask / PE (modules.py:114-118)
    g.X = a.p1; g.ldx = P; g.K = P; g.N = D; g.W = a.w.prenet_w2; g.n_w1 = D; g.mode = kPrenetOut; g.Y = a.x; g.ldy = D;
    g.emit_stats = 1;
  } else if (ph == 3 + 8 * L) {  // final LN + mel / stop projections
    p.kind = 1; g.ln = 1;
    g.X = a.x; g.ldx = D; g.K = D; g.N = M + 1; g.W = a.w.w_mel_ln; g.W2 = a.w.w_stop_ln; g.n_w1 = M;
    g.bias = a.w.c_out_ln; g.mode = kFinal;
  } else {
    const int l = (ph - 3) / 8, k = (ph - 3) % 8;
    const TtsDecLayerWeights& lw = a.w.layer[l];
    const size_t self_off = (size_t)l * B * H * T * DH, cross_off = (size_t)l * B * H * S * DH;
    float* al_self = a.st.align_self ? a.st.align_self + (size_t)l * B * H * T * T : nullptr;
    float* al_cross = a.st.align_cross ? a.st.align_cross + (size_t)l * B * H * T * S : nullptr;
    switch (k) {
      case 0:  // LN + QKV (attention.py:63-64), k/v appended at row t
        p.kind = 1; g.ln = 1;
        g.X = a.x; g.ldx = D; g.K = D; g.N = 3 * D; g.W = lw.w_qkv_ln; g.n_w1 = 3 * D; g.bias = lw.c_qkv_ln;
        g.mode = kQkv; g.Y = a.q; g.ldy = D; g.out_scale = qscale;
        g.kcache = a.st.self_k + self_off; g.vcache = a.st.self_v + self_off;
        break;
      case 1:
        p.kind = 2;
        p.at.kc = a.st.self_k + self_off; p.at.vc = a.st.self_v + self_off; p.at.rows_alloc = T; p.at.n_keys = t + 1;
        p.at.key_len = nullptr; p.at.align = al_self; p.at.align_bh_stride = (long long)T * T; p.at.align_row_len = T;
        break;
      case 2:  // output projection + residual (attention.py:118-119, modules.py:132)
        if (a.n_split > 1) {
          g.xsrc = kXCombine; g.comb_keys = t + 1; g.align = al_self; g.align_bh_stride = (long long)T * T; g.align_row_len = T;
        } else {
          g.X = a.ctx; g.ldx = D;
        }
        g.K = D; g.N = D; g.W = lw.w_self_out; g.n_w1 = D; g.mode = kPlain; g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D;
        g.emit_stats = 1;
        break;
      case 3:  // LN + cross query
        p.kind = 1; g.ln = 1;
        g.X = a.x; g.ldx = D; g.K = D; g.N = D; g.W = lw.w_cross_q_ln; g.n_w1 = D; g.bias = lw.c_cross_q_ln;
        g.mode = kPlain; g.Y = a.q; g.ldy = D; g.out_scale = qscale;
        break;
      case 4:
        p.kind = 2;
        p.at.kc = a.st.cross_k + cross_off; p.at.vc = a.st.cross_v + cross_off; p.at.rows_alloc = S; p.at.n_keys = S;
        p.at.key_len = a.st.input_lengths; p.at.align = al_cross; p.at.align_bh_stride = (long long)T * S; p.at.align_row_len = S;
        break;
      case 5:
        if (a.n_split > 1) {
          g.xsrc = kXCombine; g.comb_keys = S; g.align = al_cross; g.align_bh_stride = (long long)T * S; g.align_row_len = S;
        } else {
          g.X = a.ctx; g.ldx = D;
        }
        g.K = D; g.N = D; g.W = lw.w_cross_out; g.n_w1 = D; g.mode = kPlain; g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D;
        g.emit_stats = 1;
        break;
      case 6:  // LN + FFN-in + ReLU (modules.py:14-17)
        p.kind = 1; g.ln = 1;
        g.X = a.x; g.ldx = D; g.K = D; g.N = F; g.W = lw.w_ffn_in_ln; g.n_w1 = F; g.bias = lw.c_ffn_in_ln;
        g.relu = 1; g.mode = kPlain; g.Y = a.hid; g.ldy = F;
        break;
      default:  // FFN-out + residual
        g.X = a.hid; g.ldx = F; g.K = F; g.N = D; g.W = lw.w_ffn_out; g.n_w1 = D; g.mode = kPlain;
        g.Y = a.x; g.ldy = D; g.R = a.x; g.ldr = D; g.emit_stats = 1;
        break;
    }
  }
}

// ---- the kernel -------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kThreads, 1) fused_decode_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) float smem_raw[];
  const Smem sm = make_smem(a, smem_raw);

  const int B = a.st.batch, T = a.st.t_max;
  const int n_phases = 3 + 8 * a.w.n_layers + 1;
  GridBar gb{a.bar, a.err, 0u, gridDim.x};
  unsigned w_par = 0u, ring_count = 0u, x_par = 0u;
  const float qscale = (float)(1.0 / sqrt((double)DH));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3 + kWarps * kSlots; ++i) mbar_init(&sm.wfull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int b = threadIdx.x; b < B; b += kThreads) {
    sm.len[b] = a.st.lengths[b];
    sm.fin[b] = a.st.finished[b];
  }
  __syncthreads();
  if (a.cluster2) cluster_sync_all();  // the partner's mbarriers exist before anything is multicast into them
  const int t0 = *a.st.step_counter;

  // phase descriptors live in shared memory (one copy per CTA, written by thread 0): per-thread copies
  // would sit in local memory, and with ~200 KB of shared memory in use the L1 left for it is tiny
  __shared__ Phase ph_buf[2];
  Phase* cur = &ph_buf[0];
  Phase* nxt = &ph_buf[1];
  if (threadIdx.x == 0) {
    get_phase<DH>(a, 0, t0, qscale, *cur);
    issue_weights(cur->g, sm);
  }
  __syncthreads();

  for (int s = 0; s < a.n_steps; ++s) {
    const int t = t0 + s;
    for (int ph = 0; ph < n_phases; ++ph) {
      long long* prof = (blockIdx.x == 0 && threadIdx.x == 0 && ph < kProfPhases) ? a.prof + 8 * ph : nullptr;
      if (prof) prof[0] = clock64();
      if (cur->kind == 2) {
        ring_count = attn_phase<DH>(a, cur->at, smem_raw, ring_count, prof, t);
      } else {
        gemm_phase<DH>(a, cur->g, cur->pre, smem_raw, w_par, x_par, prof, t);
        int n_lo, n_hi;
        slice_rows(cur->g, n_lo, n_hi);
        if (n_hi > n_lo) w_par ^= 1u;  // this CTA consumed one completion of the weight barrier
        if (a.cluster2 && cur->g.N >= (int)gridDim.x && B <= kRowBlk && cur->g.xsrc != kXCombine) {
          // ... and completions of the multicast tile barrier: one per K slice and pass
          const int slices = (cur->g.K + kKC - 1) / kKC;
          const int passes = slices > 1 ? (n_hi - n_lo + kPass - 1) / kPass : 1;
          if ((slices * passes) & 1) x_par ^= 1u;
        }
      }
      const bool last = ph == n_phases - 1;
      if (prof) prof[1] = clock64();
      bar_arrive(gb, cur->kind == 1 && cur->g.mode == kQkv);
      if (!last && threadIdx.x == 0) {  // the next phase's weights stream in while we wait for the other CTAs
        get_phase<DH>(a, ph + 1, t, qscale, *nxt);
        if (nxt->kind != 2) issue_weights(nxt->g, sm);
      }
      bar_wait(gb);
      if (prof) prof[2] = clock64();
      if (!last) {
        Phase* tmp = cur;
        cur = nxt;
        nxt = tmp;
      }
    }
    // synthesize.py:42-45, replicated identically in every CTA
    if (a.update_state) {
      for (int b = threadIdx.x; b < B; b += kThreads) {
        const bool fin = sm.fin[b] != 0 || __ldcg(a.st.stop_logits + (size_t)b * T + t) > 0.f;
        sm.fin[b] = fin ? 1 : 0;
        if (!fin) sm.len[b] += 1;
      }
    }
    __syncthreads();
    int unfinished = 0;
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
    if ((a.update_state && unfinished == 0) || s + 1 == a.n_steps) break;  // uniform across CTAs
    if (threadIdx.x == 0) {
      get_phase<DH>(a, 0, t + 1, qscale, *cur);
      issue_weights(cur->g, sm);
    }
    __syncthreads();
  }
  if (a.cluster2) cluster_sync_all();  // nobody leaves while the partner could still write into its shared memory
}

static size_t smem_bytes(const TtsDecoderWeights* w, int B) {
  return (size_t)(kXFloats + kWFloats + kWarps * kPass * 32 + 2 * kRowBlk * w->n_heads + 2 * kRowBlk + kPass * 32) *
             sizeof(float) +
         (size_t)(3 + kWarps * kSlots) * sizeof(uint64_t) + (size_t)2 * B * sizeof(int) + 16;
}

struct Carve {
  float *x, *q, *ctx, *hid, *p0, *p1, *part, *stats;
  unsigned* bar;
  int* err;
  long long* prof;
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

static int split_for(const TtsDecoderWeights* w, int B, int G) {
  int ns = G / (B * w->n_heads);
  return ns < 1 ? 1 : (ns > kMaxSplit ? kMaxSplit : ns);
}

static Carve carve(const TtsDecoderWeights* w, int B, float* base) {
  Carve c;
  const size_t D = w->d_model, F = w->d_ffn, P = w->prenet_hidden, H = w->n_heads, dh = D / H;
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 31) / 32 * 32;
    return p;
  };
  c.bar = reinterpret_cast<unsigned*>(take(32));
  c.err = reinterpret_cast<int*>(take(32));
  c.prof = reinterpret_cast<long long*>(take(2 * 8 * kProfPhases));
  c.x = take(B * D); c.q = take(B * D); c.ctx = take(B * D); c.hid = take(B * F); c.p0 = take(B * P); c.p1 = take(B * P);
  const size_t splits = (size_t)B * H >= 132 ? 1 : kMaxSplit;
  c.part = take((size_t)B * H * splits * (dh + 4));
  c.stats = take((size_t)B * 160 * 2);
  c.floats = off;
  return c;
}

template <int DH>
static int launch(const Args& a, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(fused_decode_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    int per_sm = 0;
    TTS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_decode_kernel<DH>, kThreads, 226 * 1024));
    TTS_REQUIRE(per_sm >= 1, "fused decode kernel does not fit on an SM");
    configured = true;
  }
  static int cluster_ok = -1;  // -1 unknown, 0 no, 1 yes
  Args args = a;
  const size_t smem = smem_bytes(&a.w, a.st.batch);
  if (cluster_ok != 0 && num_sms() % 2 == 0 && getenv("TTS_NO_CLUSTER") == nullptr) {
    args.cluster2 = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(num_sms());
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeCooperative;
    attrs[0].val.cooperative = 1;
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = 2;
    attrs[1].val.clusterDim.y = 1;
    attrs[1].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 2;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, fused_decode_kernel<DH>, args);
    if (e == cudaSuccess) {
      cluster_ok = 1;
      count_launch();
      return 0;
    }
    if (cluster_ok == 1) {
      set_error("fused decode: cluster launch failed: %s", cudaGetErrorString(e));
      return 1;
    }
    (void)cudaGetLastError();  // first attempt failed: this device/driver cannot co-schedule 2-CTA clusters cooperatively
    cluster_ok = 0;
  }
  args.cluster2 = 0;
  void* params[] = {&args};
  TTS_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(fused_decode_kernel<DH>), dim3(num_sms()),
                                             dim3(kThreads), params, smem, s));
  count_launch();
  return 0;
}

}  // namespace fused

int fused_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, long long* out_host, int max_entries) {
  using namespace fused;
  const Carve c = carve(w, st->batch, st->scratch);
  const int n = max_entries < 8 * kProfPhases ? max_entries : 8 * kProfPhases;
  TTS_CHECK_CUDA(cudaMemcpy(out_host, c.prof, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

size_t fused_scratch_floats(const TtsDecoderWeights* w, int B) { return fused::carve(w, B, nullptr).floats; }

bool fused_supported(const TtsDecoderWeights* w, const TtsDecodeState* st) {
  using namespace fused;
  if (w->d_model > kKC || w->d_model % 16 != 0) return false;
  if (!w->w_mel_ln || !w->w_stop_ln || !w->c_out_ln) return false;   // packed LayerNorm-folded operands required
  for (int l = 0; l < w->n_layers; ++l)
    if (!w->layer[l].w_qkv_ln || !w->layer[l].c_qkv_ln || !w->layer[l].w_cross_q_ln || !w->layer[l].c_cross_q_ln ||
        !w->layer[l].w_ffn_in_ln || !w->layer[l].c_ffn_in_ln)
      return false;
  if (num_sms() > 160 || w->d_model > kPass * num_sms()) return false;
  if (w->d_ffn % 16 != 0) return false;
  if (w->prenet_hidden > kKC || w->n_mels > kKC) return false;
  if (st->batch > kMaxBatch) return false;
  const int dh = w->d_model / w->n_heads;
  if (dh != 32 && dh != 64 && dh != 96) return false;
  // the widest weight slice of any phase must fit the shared-memory weight buffer
  const int G = num_sms();
  auto slice = [&](long long N, long long K) { return ((N + G - 1) / G) * K; };
  const long long D = w->d_model, F = w->d_ffn, P = w->prenet_hidden, M = w->n_mels;
  long long worst = slice(3 * D, D);
  worst = worst > slice(F, D) ? worst : slice(F, D);
  worst = worst > slice(D, F) ? worst : slice(D, F);
  worst = worst > slice(P, M) ? worst : slice(P, M);
  worst = worst > slice(P, P) ? worst : slice(P, P);
  worst = worst > slice(D, P) ? worst : slice(D, P);
  if (worst > kWFloats) return false;
  if ((long long)kWarps * kSlots * 2 * kTK * dh > kXFloats + kWFloats) return false;
  if (smem_bytes(w, st->batch) > 226 * 1024) return false;
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
  a.x = c.x; a.q = c.q; a.ctx = c.ctx; a.hid = c.hid; a.p0 = c.p0; a.p1 = c.p1; a.part = c.part; a.stats = c.stats;
  a.n_split = split_for(w, st->batch, num_sms());
  a.bar = c.bar; a.err = c.err; a.prof = c.prof; a.n_steps = n_steps; a.update_state = update_state;
  TTS_CHECK_CUDA(cudaMemsetAsync(c.bar, 0, 256, s));  // barrier counter and error flag
  switch (w->d_model / w->n_heads) {
    case 32: return launch<32>(a, s);
    case 64: return launch<64>(a, s);
    case 96: return launch<96>(a, s);
  }
  set_error("fused decode: unsupported head_dim");
  return 2;
}

}  // namespace tts
