// Flash-style multi-head attention for the teacher-forced TRAINING path: forward, dQ and dK/dV kernels on the tensor
// cores (bf16 mma.sync m16n8k16, fp32 accumulation), masks computed from indices / lengths, Philox dropout on the
// attention weights regenerated in the backward pass.  No [B,H,Tq,Tk] tensor is ever written: the reference
// materialises logits, weights and the dropout mask (transformer/attention.py:83-91; 2.05 GB per layer at B=64,
// T=1000) and autograd keeps them for backward.
//
//   forward : O = dropout(softmax(Q K^T * scale + mask)) V, saves L = log2-sum-exp per (b, h, query)
//   dQ      : one CTA per 64 queries, loops over key blocks: S, dP = dO V^T, dS = P (keep dP / (1-p) - delta), dQ += dS K;
//             also produces delta = rowsum(dO o O) for the dK/dV kernel
//   dK / dV : one CTA per 64 keys, loops over query blocks with the TRANSPOSED products (S^T = K Q^T, dP^T = V dO^T), so
//             P^T and dS^T are already in A-fragment layout for dV += P^T dO and dK += dS^T Q: no shared-memory
//             round trip, no atomics, deterministic
//   fused   : (default when the caller provides the fp32 dQ buffer) the dK / dV kernel also produces dQ: ONE recomputation
//             of S / dP / dropout bits per tile pair instead of two; dQ is accumulated with vector atomics
// Reference semantics: transformer/attention.py:72-122 (scale on q :113-114, additive -1e20 bias :84-85 == exclusion
// from the softmax, dropout on the weights :89), masks of transformer/modules.py:50-52,109-112.
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include <atomic>

#include "philox.cuh"

namespace tts {
namespace attn {

constexpr int kThreads = 128;   // 4 warps x 16 rows
constexpr int BQ = 64, BKV = 64;

struct Args {
  const __nv_bfloat16 *q, *k, *v;
  long long ldq, ldk, ldv;
  __nv_bfloat16* o; long long ldo;
  float* lse;                       // [B][H][Tq], log2 domain
  int B, H, Tq, Tk;
  float scale_log2;                 // head_dim^-0.5 * log2(e)
  float scale;                      // head_dim^-0.5
  int causal;
  const int32_t* key_len;           // [B] or NULL
  float drop_scale; uint32_t drop_thresh; unsigned long long seed; uint32_t stream;
  // backward
  const __nv_bfloat16* d_o; long long lddo;
  float* delta;                     // [B][H][Tq]
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  float* dq_acc;                    // [B*Tq][H*DH] fp32: dQ accumulated with atomics by the fused backward kernel (or NULL)
  // optional cache of the dropout keep bits: word [bh][kw][i] holds keys 32 kw .. 32 kw + 31 of query i (n_kw = 2 ceil(Tk / 64)
  // words per query).  The forward kernel writes it (the same Philox bits), the single-pass backward reads it instead of
  // running Philox again: 1 bit per attention weight (64 MB per decoder layer at B=64, T=1000; the reference keeps 2 GB)
  uint32_t* keep_mask; int n_kw;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// [ROWS][DH] bf16 tile -> shared memory rows of DH + 8 elements (16 bytes of padding: ldmatrix conflict-free); rows at or
// beyond `rows_valid` are zero-filled
template <int DH, int ROWS>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int row0, int rows_total) {
  constexpr int LDS = DH + 8, CPR = DH / 8;   // 16-byte chunks per row
  for (int c = threadIdx.x; c < ROWS * CPR; c += kThreads) {
    const int r = c / CPR, cc = c - r * CPR;
    const bool ok = row0 + r < rows_total;
    const __nv_bfloat16* g = src + (long long)(ok ? row0 + r : 0) * ld + cc * 8;
    const uint32_t d = smem_u32(dst + r * LDS + cc * 8);
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// A fragments (16 rows x DH) of one warp from a shared-memory tile: a[ks] covers k = 16 ks .. 16 ks + 15
template <int DH>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[DH / 16][4], const __nv_bfloat16* tile, int row0) {
  constexpr int LDS = DH + 8;
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
  const uint32_t base = smem_u32(tile + (row0 + r + (mi & 1) * 8) * LDS + (mi >> 1) * 8);
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) ldsm_x4(a[ks], base + ks * 32);
}

// C[16 x 64] += A[16 x DH] . T[64 x DH]^T: the tile rows are the n index, contraction over DH (non-transposed ldmatrix)
template <int DH, int NT>
__device__ __forceinline__ void mma_a_tt(float (&c)[NT][4], const uint32_t (&a)[DH / 16][4], const __nv_bfloat16* tile, int n_row0) {
  constexpr int LDS = DH + 8;
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
  const uint32_t base = smem_u32(tile + (n_row0 + r + (mi >> 1) * 8) * LDS + (mi & 1) * 8);
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4(b, base + (np * 16 * LDS + ks * 16) * 2);
      mma_bf16(c[2 * np], a[ks], b[0], b[1]);
      mma_bf16(c[2 * np + 1], a[ks], b[2], b[3]);
    }
}

// C[16 x DH] += P[16 x KN] . T[KN x DH]: the tile rows are the contraction index (transposed ldmatrix); P comes from
// fp32 accumulator fragments p[KN/8][4] converted on the fly
template <int DH, int KNT>
__device__ __forceinline__ void mma_p_t(float (&c)[DH / 8][4], const float (&p)[KNT][4], const __nv_bfloat16* tile, int k_row0) {
  constexpr int LDS = DH + 8;
  const int lane = threadIdx.x & 31, mi = lane >> 3, r = lane & 7;
  const uint32_t base = smem_u32(tile + (k_row0 + r + (mi & 1) * 8) * LDS + (mi >> 1) * 8);
#pragma unroll
  for (int ks = 0; ks < KNT / 2; ++ks) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
    a[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    a[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < DH / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, base + (ks * 16 * LDS + np * 16) * 2);
      mma_bf16(c[2 * np], a, b[0], b[1]);
      mma_bf16(c[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// keep flags of the 8 weights a thread holds in one 16 x 16 block (two neighbouring n-tiles); TR: the accumulator rows are
// keys and the columns queries (dK/dV kernel).  bit e of the result <-> (n-tile parity e >> 2, accumulator register e & 3).
// philox.cuh: one call covers {i0, i0+8} x {j0, j0+1, j0+8, j0+9}, 16 bits per element, half-word 4*(i bit 3) + 2*(j bit 3) + (j bit 0)
template <bool TR>
__device__ __forceinline__ uint32_t keep_bits(const Args& a, unsigned long long bh, int n_iblk, int n_jblk, int row0, int col0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t2 = (lane & 3) * 2;
  const uint32_t thr = a.drop_thresh;   // 16-bit threshold
  uint32_t bits = 0;
  if (!TR) {   // rows = queries {g, g+8}, columns = keys {t2, t2+1, t2+8, t2+9}: exactly one call
    const uint4 w = philox4x32(a.seed, attn_dropout_index(bh, n_iblk, n_jblk, row0 + g, col0 + t2), a.stream);
    // half-words: x = (i, j) | (i, j+1); y = (i, j+8) | (i, j+9); z = (i+8, j) | (i+8, j+1); w = (i+8, j+8) | (i+8, j+9)
    bits |= ((w.x & 0xffffu) >= thr ? 1u : 0u) << 0;   // tile 0, reg 0
    bits |= ((w.x >> 16) >= thr ? 1u : 0u) << 1;       // tile 0, reg 1
    bits |= ((w.z & 0xffffu) >= thr ? 1u : 0u) << 2;   // tile 0, reg 2 (row g+8)
    bits |= ((w.z >> 16) >= thr ? 1u : 0u) << 3;
    bits |= ((w.y & 0xffffu) >= thr ? 1u : 0u) << 4;   // tile 1, reg 0
    bits |= ((w.y >> 16) >= thr ? 1u : 0u) << 5;
    bits |= ((w.w & 0xffffu) >= thr ? 1u : 0u) << 6;
    bits |= ((w.w >> 16) >= thr ? 1u : 0u) << 7;
  } else {
    // rows = keys {g, g+8}, columns = queries {t2, t2+1, t2+8, t2+9}.  The two query parities are two calls, but lanes g and
    // g ^ 1 (same t) need the SAME two calls (their keys share a column pair of the call): each lane computes one and they
    // swap the halves the other needs with two shuffles - one call per 8 weights here too.
    const int ec = g & 1;                                  // the query parity whose call this lane computes
    const uint4 w = philox4x32(a.seed, attn_dropout_index(bh, n_iblk, n_jblk, col0 + t2 + ec, row0 + (g & ~1)), a.stream);
    const uint32_t sh_mine = 16u * (uint32_t)ec;           // my keys are column parity g & 1 of the call's column pairs
    const uint32_t sh_other = 16u - sh_mine;
    // halves in the order (i, j), (i, j+8), (i+8, j), (i+8, j+8)
    uint32_t mine_lo = ((w.x >> sh_mine) & 0xffffu) | (((w.y >> sh_mine) & 0xffffu) << 16);
    uint32_t mine_hi = ((w.z >> sh_mine) & 0xffffu) | (((w.w >> sh_mine) & 0xffffu) << 16);
    uint32_t send_lo = ((w.x >> sh_other) & 0xffffu) | (((w.y >> sh_other) & 0xffffu) << 16);
    uint32_t send_hi = ((w.z >> sh_other) & 0xffffu) | (((w.w >> sh_other) & 0xffffu) << 16);
    const uint32_t got_lo = __shfl_xor_sync(0xffffffffu, send_lo, 4), got_hi = __shfl_xor_sync(0xffffffffu, send_hi, 4);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const uint32_t lo = e == ec ? mine_lo : got_lo, hi = e == ec ? mine_hi : got_hi;
      bits |= ((lo & 0xffffu) >= thr ? 1u : 0u) << (0 + e);   // (i, j)     : row j,   tile 0, reg e
      bits |= ((lo >> 16) >= thr ? 1u : 0u) << (2 + e);       // (i, j+8)   : row j+8, tile 0, reg 2+e
      bits |= ((hi & 0xffffu) >= thr ? 1u : 0u) << (4 + e);   // (i+8, j)   : tile 1, reg e
      bits |= ((hi >> 16) >= thr ? 1u : 0u) << (6 + e);       // (i+8, j+8) : tile 1, reg 2+e
    }
  }
  return bits;
}

__device__ __forceinline__ bool key_ok(const Args& a, int i, int j, int klen) {
  return j < klen && (!a.causal || j <= i);
}

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const Args a) {
  constexpr int LDS = DH + 8;
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16 *sQ = sm, *sK = sQ + BQ * LDS, *sV = sK + 2 * BKV * LDS;
  // heavy (late) query blocks of a causal problem first: the tail of the grid is made of short CTAs
  const int q0 = (a.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * BQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t2 = (lane & 3) * 2;
  const unsigned long long bh = (unsigned long long)b * a.H + h;
  const int klen = a.key_len ? min(a.key_len[b], a.Tk) : a.Tk;
  const int n_iblk = (a.Tq + 15) >> 4, n_jblk = (a.Tk + 15) >> 4;
  const __nv_bfloat16* qg = a.q + (long long)b * a.Tq * a.ldq + h * DH;
  const __nv_bfloat16* kg = a.k + (long long)b * a.Tk * a.ldk + h * DH;
  const __nv_bfloat16* vg = a.v + (long long)b * a.Tk * a.ldv + h * DH;

  const int i0 = q0 + warp * 16 + g;
  const int k_end = a.causal ? min(klen, q0 + BQ) : klen;
  // K/V tiles are double buffered: tile it+1 streams in (cp.async) while tile it is multiplied
  load_tile<DH, BQ>(sQ, qg, a.ldq, q0, a.Tq);
  if (k_end > 0) {
    load_tile<DH, BKV>(sK, kg, a.ldk, 0, a.Tk);
    load_tile<DH, BKV>(sV, vg, a.ldv, 0, a.Tk);
  }
  cp_commit();
  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m[2] = {-CUDART_INF_F, -CUDART_INF_F}, l[2] = {0.f, 0.f};
  uint32_t qf[DH / 16][4];

  for (int kb = 0, it = 0; kb < k_end; kb += BKV, ++it) {
    cp_wait_all();
    __syncthreads();   // tile `it` has landed for everyone, and everyone is done with the buffer tile it+1 goes into
    if (it == 0) load_a_frags<DH>(qf, sQ, warp * 16);
    const __nv_bfloat16 *cK = sK + (it & 1) * BKV * LDS, *cV = sV + (it & 1) * BKV * LDS;
    if (kb + BKV < k_end) {
      load_tile<DH, BKV>(sK + ((it + 1) & 1) * BKV * LDS, kg, a.ldk, kb + BKV, a.Tk);
      load_tile<DH, BKV>(sV + ((it + 1) & 1) * BKV * LDS, vg, a.ldv, kb + BKV, a.Tk);
      cp_commit();
    }
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    mma_a_tt<DH, 8>(s, qf, cK, 0);
    float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
    // interior tiles (every key of the tile visible to every query row of this warp) skip the mask arithmetic
    const bool open_tile = kb + BKV <= klen && (!a.causal || kb + BKV - 1 <= q0 + warp * 16);
    if (open_tile) {
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[n][e] *= a.scale_log2;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
        }
    } else {
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + (e >> 1) * 8, j = kb + n * 8 + t2 + (e & 1);
          s[n][e] = key_ok(a, i, j, klen) ? s[n][e] * a.scale_log2 : -CUDART_INF_F;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m[r], mx[r]);
      const float mu = mn == -CUDART_INF_F ? 0.f : mn;
      const float corr = ex2(m[r] - mu);   // exp2(-inf) = 0 on the first block
      m[r] = mn;
      l[r] *= corr;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        o[n][2 * r] *= corr;
        o[n][2 * r + 1] *= corr;
      }
      mx[r] = mu;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[n][e] = ex2(s[n][e] - mx[e >> 1]);
        l[e >> 1] += s[n][e];
      }
    if (a.drop_thresh != 0u) {
      uint32_t mw[4] = {0u, 0u, 0u, 0u};   // [key word of the tile (2)][row g / g + 8]
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const uint32_t bits = keep_bits<false>(a, bh, n_iblk, n_jblk, q0 + warp * 16, kb + np * 16);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!((bits >> e) & 1u)) s[2 * np + (e >> 2)][e & 3] = 0.f;
        // keys {t2, t2+1, t2+8, t2+9} of this 16-key block: bits 0,1,4,5 (row g) and 2,3,6,7 (row g + 8)
        const uint32_t sh = (uint32_t)((np & 1) * 16 + t2);
        mw[(np >> 1) * 2 + 0] |= ((bits & 3u) | (((bits >> 4) & 3u) << 8)) << sh;
        mw[(np >> 1) * 2 + 1] |= (((bits >> 2) & 3u) | (((bits >> 6) & 3u) << 8)) << sh;
      }
      if (a.keep_mask != nullptr) {   // the four lanes of a quad hold disjoint key bits of the same two rows
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          mw[w] |= __shfl_xor_sync(0xffffffffu, mw[w], 1);
          mw[w] |= __shfl_xor_sync(0xffffffffu, mw[w], 2);
        }
        const int t = lane & 3;   // lane t of the quad stores word t
        const uint32_t word = t == 0 ? mw[0] : (t == 1 ? mw[1] : (t == 2 ? mw[2] : mw[3]));
        const int row = i0 + (t & 1) * 8;
        if (row < a.Tq) a.keep_mask[((bh * a.n_kw) + (kb >> 5) + (t >> 1)) * a.Tq + row] = word;
      }
    }
    mma_p_t<DH, 8>(o, s, cV, 0);
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i >= a.Tq) continue;
    const float inv = l[r] > 0.f ? a.drop_scale / l[r] : 0.f;
    __nv_bfloat16* op = a.o + ((long long)b * a.Tq + i) * a.ldo + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
      *reinterpret_cast<uint32_t*>(op + n * 8 + t2) = pack_bf16(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
    if ((lane & 3) == 0 && a.lse != nullptr) a.lse[(bh * a.Tq) + i] = m[r] + log2f(l[r]);
  }
}

// =====================================================================================================================
// backward: dQ (+ delta)
// =====================================================================================================================
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_kernel(const Args a) {
  constexpr int LDS = DH + 8;
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16 *sQ = sm, *sDO = sQ + BQ * LDS, *sK = sDO + BQ * LDS, *sV = sK + 2 * BKV * LDS;
  const int q0 = (a.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * BQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t2 = (lane & 3) * 2;
  const unsigned long long bh = (unsigned long long)b * a.H + h;
  const int klen = a.key_len ? min(a.key_len[b], a.Tk) : a.Tk;
  const int n_iblk = (a.Tq + 15) >> 4, n_jblk = (a.Tk + 15) >> 4;
  const __nv_bfloat16* qg = a.q + (long long)b * a.Tq * a.ldq + h * DH;
  const __nv_bfloat16* dog = a.d_o + (long long)b * a.Tq * a.lddo + h * DH;
  const __nv_bfloat16* og = a.o + (long long)b * a.Tq * a.ldo + h * DH;
  const __nv_bfloat16* kg = a.k + (long long)b * a.Tk * a.ldk + h * DH;
  const __nv_bfloat16* vg = a.v + (long long)b * a.Tk * a.ldv + h * DH;
  const int i0 = q0 + warp * 16 + g;

  // delta = rowsum(dO o O) (fp32), O staged through the K tile buffer
  load_tile<DH, BQ>(sQ, qg, a.ldq, q0, a.Tq);
  load_tile<DH, BQ>(sDO, dog, a.lddo, q0, a.Tq);
  load_tile<DH, BQ>(sK, og, a.ldo, q0, a.Tq);
  cp_wait_all();
  __syncthreads();
  float dl[2];
  {
    // thread (row = tid / 2, half = tid % 2) sums half a row; pairs combine by shuffle
    const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
    float acc = 0.f;
    for (int d = half * (DH / 2); d < (half + 1) * (DH / 2); ++d)
      acc += __bfloat162float(sDO[row * LDS + d]) * __bfloat162float(sK[row * LDS + d]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    float* sdelta = reinterpret_cast<float*>(sV);   // 64 floats
    if (half == 0) {
      sdelta[row] = acc;
      if (q0 + row < a.Tq) a.delta[bh * a.Tq + q0 + row] = acc;
    }
    __syncthreads();
    dl[0] = sdelta[warp * 16 + g];
    dl[1] = sdelta[warp * 16 + g + 8];
  }
  float lse[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) lse[r] = i0 + r * 8 < a.Tq ? a.lse[bh * a.Tq + i0 + r * 8] : 0.f;

  uint32_t qf[DH / 16][4], dof[DH / 16][4];
  load_a_frags<DH>(qf, sQ, warp * 16);
  load_a_frags<DH>(dof, sDO, warp * 16);
  float dq[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
  const int k_end = a.causal ? min(klen, q0 + BQ) : klen;
  __syncthreads();   // delta staging (O in the first K buffer, sums in the first V buffer) is finished
  if (k_end > 0) {
    load_tile<DH, BKV>(sK, kg, a.ldk, 0, a.Tk);
    load_tile<DH, BKV>(sV, vg, a.ldv, 0, a.Tk);
    cp_commit();
  }

  for (int kb = 0, it = 0; kb < k_end; kb += BKV, ++it) {
    cp_wait_all();
    __syncthreads();
    const __nv_bfloat16 *cK = sK + (it & 1) * BKV * LDS, *cV = sV + (it & 1) * BKV * LDS;
    if (kb + BKV < k_end) {   // the next K/V tile streams in while this one is multiplied
      load_tile<DH, BKV>(sK + ((it + 1) & 1) * BKV * LDS, kg, a.ldk, kb + BKV, a.Tk);
      load_tile<DH, BKV>(sV + ((it + 1) & 1) * BKV * LDS, vg, a.ldv, kb + BKV, a.Tk);
      cp_commit();
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
    mma_a_tt<DH, 8>(s, qf, cK, 0);
    mma_a_tt<DH, 8>(dp, dof, cV, 0);
    // interior tiles (every key visible to every query row of this warp) skip the mask arithmetic
    const bool open_tile = kb + BKV <= klen && (!a.causal || kb + BKV - 1 <= q0 + warp * 16);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bits = 0xffu;
      if (a.drop_thresh != 0u) bits = keep_bits<false>(a, bh, n_iblk, n_jblk, q0 + warp * 16, kb + np * 16);
#pragma unroll
      for (int e8 = 0; e8 < 8; ++e8) {
        const int n = 2 * np + (e8 >> 2), e = e8 & 3;
        float p = ex2(s[n][e] * a.scale_log2 - lse[e >> 1]);
        if (!open_tile) {
          const int i = i0 + (e >> 1) * 8, j = kb + n * 8 + t2 + (e & 1);
          p = key_ok(a, i, j, klen) ? p : 0.f;
        }
        const float dpe = ((bits >> e8) & 1u) ? dp[n][e] * a.drop_scale : 0.f;
        s[n][e] = p * (dpe - dl[e >> 1]) * a.scale;   // dS (scaled: dQ = scale * dS K)
      }
    }
    mma_p_t<DH, 8>(dq, s, cK, 0);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i >= a.Tq) continue;
    __nv_bfloat16* dp_ = a.dq + ((long long)b * a.Tq + i) * a.lddq + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) *reinterpret_cast<uint32_t*>(dp_ + n * 8 + t2) = pack_bf16(dq[n][2 * r], dq[n][2 * r + 1]);
  }
}

// =====================================================================================================================
// backward: dK, dV (transposed products)
// =====================================================================================================================
template <int DH>
__global__ void __launch_bounds__(kThreads) attn_bwd_dkv_kernel(const Args a) {
  constexpr int LDS = DH + 8, BQ2 = 32;
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16 *sK = sm, *sV = sK + BKV * LDS, *sQ = sV + BKV * LDS, *sDO = sQ + 2 * BQ2 * LDS;
  float* sL = reinterpret_cast<float*>(sDO + 2 * BQ2 * LDS);   // 2 x ([32] lse, [32] delta)
  const int j0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t2 = (lane & 3) * 2;
  const unsigned long long bh = (unsigned long long)b * a.H + h;
  const int klen = a.key_len ? min(a.key_len[b], a.Tk) : a.Tk;
  const int n_iblk = (a.Tq + 15) >> 4, n_jblk = (a.Tk + 15) >> 4;
  const __nv_bfloat16* qg = a.q + (long long)b * a.Tq * a.ldq + h * DH;
  const __nv_bfloat16* dog = a.d_o + (long long)b * a.Tq * a.lddo + h * DH;
  const __nv_bfloat16* kg = a.k + (long long)b * a.Tk * a.ldk + h * DH;
  const __nv_bfloat16* vg = a.v + (long long)b * a.Tk * a.ldv + h * DH;
  const int jr = j0 + warp * 16 + g;   // this thread's key rows: jr, jr + 8

  load_tile<DH, BKV>(sK, kg, a.ldk, j0, a.Tk);
  load_tile<DH, BKV>(sV, vg, a.ldv, j0, a.Tk);
  cp_wait_all();
  __syncthreads();
  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  load_a_frags<DH>(kf, sK, warp * 16);
  load_a_frags<DH>(vf, sV, warp * 16);
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  const bool any_key = j0 < klen;
  const int q_begin = a.causal ? (j0 / BQ2) * BQ2 : 0;   // queries before the first key of the block never see it

  auto load_q = [&](int qb, int buf) {   // Q / dO tiles and the per-query lse / delta of block qb -> buffer buf
    load_tile<DH, BQ2>(sQ + buf * BQ2 * LDS, qg, a.ldq, qb, a.Tq);
    load_tile<DH, BQ2>(sDO + buf * BQ2 * LDS, dog, a.lddo, qb, a.Tq);
    if (threadIdx.x < 2 * BQ2) {
      const int which = threadIdx.x >> 5, i = qb + (threadIdx.x & 31);
      const float* src = which ? a.delta : a.lse;
      const bool ok = i < a.Tq;
      const uint32_t d = smem_u32(sL + buf * 2 * BQ2 + threadIdx.x);
      const int bytes = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src + bh * a.Tq + (ok ? i : 0)), "r"(bytes) : "memory");
    }
    cp_commit();
  };
  if (any_key && q_begin < a.Tq) load_q(q_begin, 0);

  for (int qb = q_begin, it = 0; qb < a.Tq && any_key; qb += BQ2, ++it) {
    cp_wait_all();
    __syncthreads();
    const __nv_bfloat16 *cQ = sQ + (it & 1) * BQ2 * LDS, *cDO = sDO + (it & 1) * BQ2 * LDS;
    const float* cL = sL + (it & 1) * 2 * BQ2;
    if (qb + BQ2 < a.Tq) load_q(qb + BQ2, (it + 1) & 1);
    float st[4][4], dpt[4][4];   // S^T and dP^T: rows = keys, columns = the 32 queries of the block
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
    }
    mma_a_tt<DH, 4>(st, kf, cQ, 0);
    mma_a_tt<DH, 4>(dpt, vf, cDO, 0);
    float pt[4][4];
    // the per-query statistics of this thread's 8 columns, once per tile (not once per element)
    float lq[4][2], dq_[4][2];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float2 l2 = *reinterpret_cast<const float2*>(cL + n * 8 + t2), d2 = *reinterpret_cast<const float2*>(cL + BQ2 + n * 8 + t2);
      lq[n][0] = l2.x; lq[n][1] = l2.y; dq_[n][0] = d2.x; dq_[n][1] = d2.y;
    }
    // interior tiles: all 32 queries exist and see all 16 keys of this warp
    const bool open_tile = qb + BQ2 <= a.Tq && j0 + warp * 16 + 16 <= klen && (!a.causal || j0 + warp * 16 + 15 <= qb);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bits = 0xffu;
      if (a.drop_thresh != 0u) bits = keep_bits<true>(a, bh, n_iblk, n_jblk, j0 + warp * 16, qb + np * 16);
#pragma unroll
      for (int e8 = 0; e8 < 8; ++e8) {
        const int n = 2 * np + (e8 >> 2), e = e8 & 3;
        float p = ex2(st[n][e] * a.scale_log2 - lq[n][e & 1]);
        if (!open_tile) {
          const int j = jr + (e >> 1) * 8, i = qb + n * 8 + t2 + (e & 1);
          p = (i < a.Tq && j < a.Tk && key_ok(a, i, j, klen)) ? p : 0.f;
        }
        const bool keep = (bits >> e8) & 1u;
        pt[n][e] = keep ? p * a.drop_scale : 0.f;                         // dropped-and-scaled weights: dV = P_drop^T dO
        const float dpe = keep ? dpt[n][e] * a.drop_scale : 0.f;
        st[n][e] = p * (dpe - dq_[n][e & 1]) * a.scale;                  // dS^T (scaled: dK = scale * dS^T Q)
      }
    }
    mma_p_t<DH, 4>(dv, pt, cDO, 0);
    mma_p_t<DH, 4>(dk, st, cQ, 0);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = jr + r * 8;
    if (j >= a.Tk) continue;
    __nv_bfloat16* kp = a.dk + ((long long)b * a.Tk + j) * a.lddk + h * DH;
    __nv_bfloat16* vp = a.dv + ((long long)b * a.Tk + j) * a.lddv + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(kp + n * 8 + t2) = pack_bf16(dk[n][2 * r], dk[n][2 * r + 1]);
      *reinterpret_cast<uint32_t*>(vp + n * 8 + t2) = pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
    }
  }
}


// =====================================================================================================================
// backward, single pass: dK, dV and dQ from ONE recomputation of S and dP per (query block, key block) pair.
// The two-kernel path above recomputes S / dP / the dropout bits twice (7 tile products per pair); this kernel does 5:
// the dK/dV kernel's transposed products, then dS^T goes through shared memory (bf16) and every warp multiplies a
// [16 queries x 64 keys] slice of dS with the resident K tile; the partial dQ of this key block is added to an fp32
// buffer with vector atomics (red.global.add.v2.f32), which a small kernel converts to bf16 afterwards.  The sum order
// of dQ over key blocks is not fixed (like the reference's own atomics in backward); dK / dV stay deterministic.
// delta = rowsum(dO o O) comes from attn_bwd_prep_kernel.
// =====================================================================================================================
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <int DH>
__global__ void __launch_bounds__(kThreads) attn_bwd_fused_kernel(const Args a) {
  constexpr int LDS = DH + 8, BQ2 = 32, LDP = BQ2 + 8;   // dS^T rows: 32 queries + 8 (80 bytes: ldmatrix / 4-byte stores conflict-free)
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16 *sK = sm, *sV = sK + BKV * LDS, *sQ = sV + BKV * LDS, *sDO = sQ + 2 * BQ2 * LDS;
  float* sL = reinterpret_cast<float*>(sDO + 2 * BQ2 * LDS);   // 2 x ([32] lse, [32] delta)
  uint32_t* sM = reinterpret_cast<uint32_t*>(sL + 4 * BQ2);    // 2 x [2 key words][32 queries] cached keep bits
  __nv_bfloat16* sDS = reinterpret_cast<__nv_bfloat16*>(sM + 4 * BQ2);   // [64 keys][LDP]
  // heavy (early) key blocks of a causal problem first
  const int j0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t2 = (lane & 3) * 2;
  const unsigned long long bh = (unsigned long long)b * a.H + h;
  const int klen = a.key_len ? min(a.key_len[b], a.Tk) : a.Tk;
  const int n_iblk = (a.Tq + 15) >> 4, n_jblk = (a.Tk + 15) >> 4;
  const __nv_bfloat16* qg = a.q + (long long)b * a.Tq * a.ldq + h * DH;
  const __nv_bfloat16* dog = a.d_o + (long long)b * a.Tq * a.lddo + h * DH;
  const __nv_bfloat16* kg = a.k + (long long)b * a.Tk * a.ldk + h * DH;
  const __nv_bfloat16* vg = a.v + (long long)b * a.Tk * a.ldv + h * DH;
  const int jr = j0 + warp * 16 + g;   // this thread's key rows: jr, jr + 8
  const long long ldacc = (long long)a.H * DH;
  float* accg = a.dq_acc + (long long)b * a.Tq * ldacc + h * DH;

  load_tile<DH, BKV>(sK, kg, a.ldk, j0, a.Tk);
  load_tile<DH, BKV>(sV, vg, a.ldv, j0, a.Tk);
  cp_wait_all();
  __syncthreads();
  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  load_a_frags<DH>(kf, sK, warp * 16);
  load_a_frags<DH>(vf, sV, warp * 16);
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  const bool use_mask = a.keep_mask != nullptr && a.drop_thresh != 0u;
  const bool any_key = j0 < klen;
  const int q_begin = a.causal ? (j0 / BQ2) * BQ2 : 0;   // queries before the first key of the block never see it

  auto load_q = [&](int qb, int buf) {   // Q / dO tiles and the per-query lse / delta of block qb -> buffer buf
    load_tile<DH, BQ2>(sQ + buf * BQ2 * LDS, qg, a.ldq, qb, a.Tq);
    load_tile<DH, BQ2>(sDO + buf * BQ2 * LDS, dog, a.lddo, qb, a.Tq);
    if (threadIdx.x < 2 * BQ2) {
      const int which = threadIdx.x >> 5, i = qb + (threadIdx.x & 31);
      const float* src = which ? a.delta : a.lse;
      const bool ok = i < a.Tq;
      const uint32_t d = smem_u32(sL + buf * 2 * BQ2 + threadIdx.x);
      const int bytes = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src + bh * a.Tq + (ok ? i : 0)), "r"(bytes) : "memory");
    } else if (use_mask) {
      const int idx = threadIdx.x - 2 * BQ2, kw = idx >> 5, i = qb + (idx & 31);
      const bool ok = i < a.Tq;
      const uint32_t d = smem_u32(sM + buf * 2 * BQ2 + idx);
      const int bytes = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d),
                   "l"(a.keep_mask + (bh * a.n_kw + (j0 >> 5) + kw) * a.Tq + (ok ? i : 0)), "r"(bytes) : "memory");
    }
    cp_commit();
  };
  if (any_key && q_begin < a.Tq) load_q(q_begin, 0);

  // dQ stage: warp w owns query rows (w & 1) * 16 .. + 15 of the block and head columns (w >> 1) * DH/2 .. + DH/2 - 1
  const int mt = warp & 1, nh = warp >> 1;
  constexpr int NP = DH / 32;   // pairs of n-tiles per warp in the dQ stage (DH / 2 columns)
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t ds_base = smem_u32(sDS + ((mi >> 1) * 8 + r8) * LDP + mt * 16 + (mi & 1) * 8);
  const uint32_t kb_base = smem_u32(sK + (r8 + (mi & 1) * 8) * LDS + (mi >> 1) * 8 + nh * (DH / 2));

  for (int qb = q_begin, it = 0; qb < a.Tq && any_key; qb += BQ2, ++it) {
    cp_wait_all();
    __syncthreads();   // block qb has landed; everyone is past the dQ stage of the previous block (sDS may be rewritten)
    const __nv_bfloat16 *cQ = sQ + (it & 1) * BQ2 * LDS, *cDO = sDO + (it & 1) * BQ2 * LDS;
    const float* cL = sL + (it & 1) * 2 * BQ2;
    if (qb + BQ2 < a.Tq) load_q(qb + BQ2, (it + 1) & 1);
    float st[4][4], dpt[4][4];   // S^T and dP^T: rows = keys, columns = the 32 queries of the block
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
    }
    mma_a_tt<DH, 4>(st, kf, cQ, 0);
    mma_a_tt<DH, 4>(dpt, vf, cDO, 0);
    float pt[4][4];
    float lq[4][2], dq_[4][2];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float2 l2 = *reinterpret_cast<const float2*>(cL + n * 8 + t2), d2 = *reinterpret_cast<const float2*>(cL + BQ2 + n * 8 + t2);
      lq[n][0] = l2.x; lq[n][1] = l2.y; dq_[n][0] = d2.x; dq_[n][1] = d2.y;
    }
    const bool open_tile = qb + BQ2 <= a.Tq && j0 + warp * 16 + 16 <= klen && (!a.causal || j0 + warp * 16 + 15 <= qb);
    const uint32_t* cM = sM + (it & 1) * 2 * BQ2 + (warp >> 1) * BQ2;   // the key word of this warp's 16 keys, per query
    const uint32_t kbit = (uint32_t)((warp & 1) * 16 + g);              // bit of key row jr (jr + 8: kbit + 8)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bits = 0xffu;
      if (use_mask) {   // bit e8 <-> n-tile 2 np + (e8 >> 2), register e8 & 3: query n * 8 + t2 + (e & 1), key row jr + 8 (e >> 1)
        const uint2 w0 = *reinterpret_cast<const uint2*>(cM + (2 * np) * 8 + t2), w1 = *reinterpret_cast<const uint2*>(cM + (2 * np + 1) * 8 + t2);
        bits = ((w0.x >> kbit) & 1u) | (((w0.y >> kbit) & 1u) << 1) | (((w0.x >> (kbit + 8)) & 1u) << 2) | (((w0.y >> (kbit + 8)) & 1u) << 3) |
               (((w1.x >> kbit) & 1u) << 4) | (((w1.y >> kbit) & 1u) << 5) | (((w1.x >> (kbit + 8)) & 1u) << 6) | (((w1.y >> (kbit + 8)) & 1u) << 7);
      } else if (a.drop_thresh != 0u) bits = keep_bits<true>(a, bh, n_iblk, n_jblk, j0 + warp * 16, qb + np * 16);
#pragma unroll
      for (int e8 = 0; e8 < 8; ++e8) {
        const int n = 2 * np + (e8 >> 2), e = e8 & 3;
        float p = ex2(st[n][e] * a.scale_log2 - lq[n][e & 1]);
        if (!open_tile) {
          const int j = jr + (e >> 1) * 8, i = qb + n * 8 + t2 + (e & 1);
          p = (i < a.Tq && j < a.Tk && key_ok(a, i, j, klen)) ? p : 0.f;
        }
        const bool keep = (bits >> e8) & 1u;
        pt[n][e] = keep ? p * a.drop_scale : 0.f;
        const float dpe = keep ? dpt[n][e] * a.drop_scale : 0.f;
        st[n][e] = p * (dpe - dq_[n][e & 1]) * a.scale;   // dS^T, scaled (dK = scale dS^T Q, dQ = scale dS K)
      }
    }
    // dS^T of this warp's 16 keys -> shared memory (bf16), read back transposed by the dQ stage
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<uint32_t*>(sDS + (warp * 16 + g) * LDP + n * 8 + t2) = pack_bf16(st[n][0], st[n][1]);
      *reinterpret_cast<uint32_t*>(sDS + (warp * 16 + g + 8) * LDP + n * 8 + t2) = pack_bf16(st[n][2], st[n][3]);
    }
    mma_p_t<DH, 4>(dv, pt, cDO, 0);
    mma_p_t<DH, 4>(dk, st, cQ, 0);
    __syncthreads();   // all four warps' dS^T rows are in shared memory
    {
      float dqa[2 * NP][4];
#pragma unroll
      for (int n = 0; n < 2 * NP; ++n) dqa[n][0] = dqa[n][1] = dqa[n][2] = dqa[n][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < BKV / 16; ++ks) {   // contraction over the 64 keys of the block
        uint32_t af[4];
        ldsm_x4_t(af, ds_base + (uint32_t)(ks * 16 * LDP) * 2u);
#pragma unroll
        for (int np = 0; np < NP; ++np) {
          uint32_t bfr[4];
          ldsm_x4_t(bfr, kb_base + (uint32_t)(ks * 16 * LDS + np * 16) * 2u);
          mma_bf16(dqa[2 * np], af, bfr[0], bfr[1]);
          mma_bf16(dqa[2 * np + 1], af, bfr[2], bfr[3]);
        }
      }
      const int i0 = qb + mt * 16 + g;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int i = i0 + rr * 8;
        if (i < a.Tq) {
          float* dst = accg + (long long)i * ldacc + nh * (DH / 2) + t2;
#pragma unroll
          for (int n = 0; n < 2 * NP; ++n) red_add_v2(dst + n * 8, dqa[n][2 * rr], dqa[n][2 * rr + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int j = jr + rr * 8;
    if (j >= a.Tk) continue;
    __nv_bfloat16* kp = a.dk + ((long long)b * a.Tk + j) * a.lddk + h * DH;
    __nv_bfloat16* vp = a.dv + ((long long)b * a.Tk + j) * a.lddv + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(kp + n * 8 + t2) = pack_bf16(dk[n][2 * rr], dk[n][2 * rr + 1]);
      *reinterpret_cast<uint32_t*>(vp + n * 8 + t2) = pack_bf16(dv[n][2 * rr], dv[n][2 * rr + 1]);
    }
  }
}

// delta[b][h][i] = sum_d dO[i][h][d] * O[i][h][d] (fp32): four threads per (query row, head), each takes every fourth
// 16-byte chunk of the head's row, so a quad reads 64 contiguous bytes per instruction (full sectors; one thread per
// head read 16 of every 32-byte sector and ran at half the bandwidth).  dq_acc is zeroed by a memset node.
template <int DH>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const Args a) {
  const long long n = (long long)a.B * a.Tq * a.H;
  long long idx = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
  const bool valid = idx < n;
  if (!valid) idx = n - 1;
  const int t = threadIdx.x & 3;
  const int h = (int)(idx % a.H);
  const long long row = idx / a.H;           // b * Tq + i
  const int b = (int)(row / a.Tq), i = (int)(row - (long long)b * a.Tq);
  const uint4* dp = reinterpret_cast<const uint4*>(a.d_o + row * a.lddo + h * DH);
  const uint4* op = reinterpret_cast<const uint4*>(a.o + row * a.ldo + h * DH);
  float acc = 0.f;
#pragma unroll
  for (int c = t; c < DH / 8; c += 4) {
    const uint4 x = __ldg(dp + c), y = __ldg(op + c);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc = fmaf(__uint_as_float(xs[e] << 16), __uint_as_float(ys[e] << 16), acc);
      acc = fmaf(__uint_as_float(xs[e] & 0xffff0000u), __uint_as_float(ys[e] & 0xffff0000u), acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (valid && t == 0) a.delta[((long long)b * a.H + h) * a.Tq + i] = acc;
}

// dq (bf16, row stride lddq) = dq_acc (fp32 [B*Tq][H*DH])
__global__ void __launch_bounds__(256) attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq,
                                                                  long long lddq, long long rows, int width, float f) {
  const int per_row = width / 8;
  const long long n = rows * per_row;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / per_row;
    const int c = (int)(idx - row * per_row) * 8;
    const float4 x = *reinterpret_cast<const float4*>(acc + row * width + c), y = *reinterpret_cast<const float4*>(acc + row * width + c + 4);
    uint4 o;
    o.x = pack_bf16(x.x * f, x.y * f); o.y = pack_bf16(x.z * f, x.w * f); o.z = pack_bf16(y.x * f, y.y * f); o.w = pack_bf16(y.z * f, y.w * f);
    *reinterpret_cast<uint4*>(dq + row * lddq + c) = o;
  }
}

template <int DH>
static int launch_fwd(const Args& a, cudaStream_t s) {
  const size_t smem = (size_t)(BQ + 4 * BKV) * (DH + 8) * 2;
  TTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_kernel<DH><<<dim3(ceil_div(a.Tq, BQ), a.H, a.B), kThreads, smem, s>>>(a);
  TTS_CHECK_LAUNCH();
  return 0;
}
template <int DH>
static int launch_bwd_fused(const Args& a, cudaStream_t s) {
  const size_t smem = (size_t)(2 * BKV + 4 * 32) * (DH + 8) * 2 + 256 * sizeof(float) + (size_t)BKV * (32 + 8) * 2;
  TTS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n = (long long)a.B * a.Tq * a.H;
  TTS_CHECK_CUDA(cudaMemsetAsync(a.dq_acc, 0, (size_t)n * DH * sizeof(float), s));
  attn_bwd_prep_kernel<DH><<<(unsigned)((n + 63) / 64), 256, 0, s>>>(a);
  TTS_CHECK_LAUNCH();
  attn_bwd_fused_kernel<DH><<<dim3(ceil_div(a.Tk, BKV), a.H, a.B), kThreads, smem, s>>>(a);
  TTS_CHECK_LAUNCH();
  const long long rows = (long long)a.B * a.Tq;
  const long long work = rows * (a.H * DH / 8);
  attn_bwd_dq_convert_kernel<<<(unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16), 256, 0, s>>>(
      a.dq_acc, a.dq, a.lddq, rows, a.H * DH, 1.f);
  TTS_CHECK_LAUNCH();
  return 0;
}
template <int DH>
static int launch_bwd(const Args& a, cudaStream_t s) {
  if (a.dq_acc != nullptr) return launch_bwd_fused<DH>(a, s);
  const size_t smem_q = (size_t)(2 * BQ + 4 * BKV) * (DH + 8) * 2;
  const size_t smem_kv = (size_t)(2 * BKV + 4 * 32) * (DH + 8) * 2 + 128 * sizeof(float);
  TTS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
  TTS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
  attn_bwd_dq_kernel<DH><<<dim3(ceil_div(a.Tq, BQ), a.H, a.B), kThreads, smem_q, s>>>(a);
  TTS_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<DH><<<dim3(ceil_div(a.Tk, BKV), a.H, a.B), kThreads, smem_kv, s>>>(a);
  TTS_CHECK_LAUNCH();
  return 0;
}

static int fill(Args& a, const TtsAttnTrain* t) {
  TTS_REQUIRE(t && t->q && t->k && t->v && t->out && t->lse, "attn_train: null argument");
  TTS_REQUIRE(t->batch > 0 && t->n_heads > 0 && t->tq > 0 && t->tk > 0, "attn_train: empty problem");
  TTS_REQUIRE(t->head_dim == 32 || t->head_dim == 64 || t->head_dim == 96, "attn_train: head_dim %d not in {32,64,96}", t->head_dim);
  TTS_REQUIRE(t->ldq % 8 == 0 && t->ldk % 8 == 0 && t->ldv % 8 == 0 && t->ldo % 8 == 0, "attn_train: row strides must be multiples of 8");
  TTS_REQUIRE(!t->causal || t->tq == t->tk, "attn_train: causal needs tq == tk");
  memset(&a, 0, sizeof(a));
  a.q = reinterpret_cast<const __nv_bfloat16*>(t->q); a.k = reinterpret_cast<const __nv_bfloat16*>(t->k);
  a.v = reinterpret_cast<const __nv_bfloat16*>(t->v);
  a.ldq = t->ldq; a.ldk = t->ldk; a.ldv = t->ldv;
  a.o = reinterpret_cast<__nv_bfloat16*>(t->out); a.ldo = t->ldo;
  a.lse = t->lse;
  a.B = t->batch; a.H = t->n_heads; a.Tq = t->tq; a.Tk = t->tk;
  a.scale = 1.f / sqrtf((float)t->head_dim);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.causal = t->causal; a.key_len = t->key_len;
  a.drop_scale = 1.f;
  if (t->drop_p > 0.f) {
    TTS_REQUIRE(t->drop_p < 1.f, "attn_train: drop_p must be < 1");
    a.drop_thresh = drop_threshold16(t->drop_p);
    a.drop_scale = 1.f / (1.f - t->drop_p);
    a.seed = t->seed; a.stream = t->rng_stream;
    a.keep_mask = t->keep_mask; a.n_kw = 2 * ((t->tk + 63) / 64);
  }
  return 0;
}

}  // namespace attn
}  // namespace tts

#include "attn_bwd_tc.cuh"
#include "attn_fwd_tc.cuh"

using namespace tts;

extern "C" int32_t tts_attn_keep_words(int32_t tk) { return 2 * ((tk + 63) / 64); }

extern "C" int tts_attn_train_fwd(const TtsAttnTrain* t, void* stream) {
  attn::Args a;
  int rc = attn::fill(a, t);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // head_dim 96 (the decoder's self- and cross-attention): the tcgen05 kernel (TTS_ATTN_FWD_TC=0: the mma.sync kernel)
  static const bool tc_on = !(getenv("TTS_ATTN_FWD_TC") != nullptr && atoi(getenv("TTS_ATTN_FWD_TC")) == 0);
  if (tc_on && t->head_dim == 96) return attn::launch_fwd_tc(a, s);
  switch (t->head_dim) {
    case 32: return attn::launch_fwd<32>(a, s);
    case 64: return attn::launch_fwd<64>(a, s);
    default: return attn::launch_fwd<96>(a, s);
  }
}

extern "C" int tts_attn_train_bwd(const TtsAttnTrain* t, void* stream) {
  attn::Args a;
  int rc = attn::fill(a, t);
  if (rc) return rc;
  TTS_REQUIRE(t->d_out && t->delta && t->dq && t->dk && t->dv, "attn_train_bwd: null gradient buffers");
  TTS_REQUIRE(t->lddo % 8 == 0 && t->lddq % 8 == 0 && t->lddk % 8 == 0 && t->lddv % 8 == 0, "attn_train_bwd: row strides must be multiples of 8");
  a.d_o = reinterpret_cast<const __nv_bfloat16*>(t->d_out); a.lddo = t->lddo;
  a.delta = t->delta;
  a.dq = reinterpret_cast<__nv_bfloat16*>(t->dq); a.dk = reinterpret_cast<__nv_bfloat16*>(t->dk);
  a.dv = reinterpret_cast<__nv_bfloat16*>(t->dv);
  a.lddq = t->lddq; a.lddk = t->lddk; a.lddv = t->lddv;
  a.dq_acc = t->dq_acc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // head_dim 96 (the decoder): the tcgen05 kernel; it takes the dropout bits from the forward kernel's cache only
  static const bool tc_on = !(getenv("TTS_ATTN_TC") != nullptr && atoi(getenv("TTS_ATTN_TC")) == 0);
  if (tc_on && t->head_dim == 96 && a.dq_acc != nullptr && (a.drop_thresh == 0u || a.keep_mask != nullptr)) return attn::launch_bwd_tc(a, s);
  switch (t->head_dim) {
    case 32: return attn::launch_bwd<32>(a, s);
    case 64: return attn::launch_bwd<64>(a, s);
    default: return attn::launch_bwd<96>(a, s);
  }
}

/* 1 after a barrier wait of the tcgen05 attention kernel timed out (protocol error); reading resets it */
extern "C" int tts_attn_tc_status(void) {
  int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, attn::tc::g_err, sizeof(int)) != cudaSuccess) return -1;
  if (v != 0) cudaMemcpyToSymbol(attn::tc::g_err, &zero, sizeof(int));
  return v;
}

/* diagnostics: copies the 3 x 32 x 8 SM-clock stamps recorded with TTS_ATTN_TC_TRACE=1 (tests/tools_attn_trace.py) */
extern "C" int tts_attn_tc_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, attn::tc::g_trace, sizeof(long long) * 3 * 32 * 8) == cudaSuccess ? 0 : 1;
}
