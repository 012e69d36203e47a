// Attention backward of the teacher-forced TRAINING path on the 5th-generation tensor cores (tcgen05 + tensor memory):
// dK, dV and dQ of transformer/attention.py:72-122 for head_dim 96, one CTA per (sample, head, 128-key block), looping over
// 64-query blocks.  Included by attn_train.cu (shares Args, the delta / dq_acc prologue and the dQ conversion kernel).
//
// Per query block the tensor cores run five products with fp32 accumulators in TENSOR MEMORY (512 columns, all used):
//   S^T  [128 keys x 64 q] = K Q^T          (K-major A = K tile, K-major B = Q tile; 6 k-steps of 16 over head_dim 96)
//   dP^T [128 x 64]        = V dO^T
//   dV   [128 x 96]       += P^T dO         (A = P^T from shared memory, MN-major B = dO tile; accumulates over the loop)
//   dK   [128 x 96]       += dS^T Q
//   dQ^T [128(d) x 64 q]   = K^T dS^T       (MN-major A = K tile, MN-major B = dS^T; rows 96..127 are padding)
// and eight "softmax" warps turn S^T / dP^T into P^T / dS^T: one thread per key row (tcgen05.ld 32 lanes x 32 columns, two
// warps per lane quadrant), p = exp2(s scale - lse[q]), masks from indices / lengths, the dropout keep bit of the element
// from the forward kernel's bit cache (one 32-bit word per query and 32 keys = one word per warp and query: no Philox in
// backward), bf16 results written to shared memory in the 128-byte-swizzled layout both MMAs read (the SAME dS^T buffer
// is the K-major A operand of dK and the MN-major B operand of dQ^T).  Four more warps add dQ^T to the fp32 dQ buffer with
// coalesced red.global.add (lane = head column), a TMA warp streams Q / dO tiles (cp.async.bulk.tensor.2d, 128-byte
// swizzle) and the per-query statistics two stages ahead, one elected lane issues every tcgen05.mma.  S^T / dP^T are
// double-buffered in tensor memory so the products of block i+1 run while the softmax warps work on block i.
// head_dim 96 = 1.5 swizzle atoms: tiles are loaded 128 columns wide (the extra 32 columns are the next head's, or zero
// past the tensor) and the k-loops / N extents simply stop at 96.
#pragma once
#include <cuda.h>

namespace tts {
int make_tma_map_bf16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_outer);   // gemm_bf16.cu

namespace attn {
namespace tc {

constexpr int DH = 96;
constexpr int BKT = 128, BQT = 64;
constexpr int kSoftWarps = 16;                      // 4 per tensor-memory lane quadrant: 16 query columns per thread
constexpr int CPT = BQT * 4 / kSoftWarps;
constexpr int kWarps = 2 + kSoftWarps + 4 + 1;        // TMA, MMA issuer A, softmax, 4 dQ warps, MMA issuer B
constexpr int kThreads = kWarps * 32;
constexpr int kQStages = 3;   // Q / dO / statistics ring: the tile of block i+2 is requested while block i is worked on
constexpr uint32_t oK = 0, oV = 32768, oQ = 65536, kStageBytes = 32768, oDS = oQ + kQStages * kStageBytes,   // dS^T, two buffers of 16 KB
                   oStat = oDS + 2 * 16384, kStatBytes = 1536, oBar = oStat + kQStages * kStatBytes;
constexpr size_t kSmem = oBar + 256 + 1024;
constexpr uint32_t cS = 0, cDP = 128, cDV = 256, cDK = 352, cDQ = 448;   // tensor-memory columns
constexpr long long kTimeout = 1LL << 28;

__device__ int g_err;   // sticky: a barrier wait timed out
__device__ long long g_trace[3][32][8];   // diagnostics (TTS_ATTN_TC_TRACE=1): SM-clock stamps of CTA (0,0,0): [softmax warp, MMA, dQ warp][block][event]
// compiled in only with -DTTS_ATTN_TC_TRACE_BUILD (the stamps lengthen the single-thread MMA issue path by ~7 %)
#ifdef TTS_ATTN_TC_TRACE_BUILD
#define TC_STAMP(role, k) do { if (tracing && it < 32) g_trace[role][it][k] = clock64(); } while (0)
#else
#define TC_STAMP(role, k) do { } while (0)
#endif

struct Bars {
  uint64_t kv_full, q_full[kQStages], q_empty[kQStages], s_full[2], p_full[2], p_empty[2], dq_full, dq_empty, acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error raises g_err (every later wait of the launch then falls through) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  long long t0 = 0, spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if ((++spins & 255) == 0) {
      if (*reinterpret_cast<volatile int*>(&g_err) != 0) return;
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeout) {
        atomicExch(&g_err, 1);
        return;
      }
    }
  }
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle (see gemm_bf16.cu make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (128 lanes x 8 columns of bf16 pairs), B from shared memory
__device__ __forceinline__ void umma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

struct Params {
  int B, H, Tq, Tk, causal, use_mask, n_kw;
  float scale_log2, scale, drop_scale;
  const int32_t* key_len;
  const float *lse, *delta;
  const uint32_t* keep_mask;
  __nv_bfloat16 *dk, *dv;
  long long lddk, lddv;
  float* dq_acc;
  int trace;
};

// One thread's CPT (key row, query) elements of a block: P^T (kept weights, NOT yet scaled by 1 / (1 - p)) and dS^T (not yet
// scaled by head_dim^-0.5) as packed bf16 pairs.  Both scale factors are applied once, to the accumulated dV / dK / dQ.
//   p = exp2(s scale - lse);  dS = p (keep ? dP / (1 - p_drop) : 0  -  delta)
template <bool OPEN>
__device__ __forceinline__ void soft_block(const Params& p, const uint32_t (&sv)[CPT], const uint32_t (&dv)[CPT], uint32_t stl,
                                           int quad, int lane, int i0, int j, int klen, uint32_t (&pw)[CPT / 2], uint32_t (&dw)[CPT / 2]) {
  // stl: shared-memory address of this warp's CPT lse values; + 256: delta; + 512 + 256 quad: the keep words of its 32 keys
  const uint32_t stm = stl + 512 + quad * 256;
  const uint32_t lane_bit = 1u << lane;
#pragma unroll
  for (int c4 = 0; c4 < CPT / 4; ++c4) {
    uint32_t ls_[4], dl_[4], mk[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    ld_shared_v4(stl + 16 * c4, ls_);
    ld_shared_v4(stl + 256 + 16 * c4, dl_);
    if (p.use_mask) ld_shared_v4(stm + 16 * c4, mk);
    const float ls[4] = {__uint_as_float(ls_[0]), __uint_as_float(ls_[1]), __uint_as_float(ls_[2]), __uint_as_float(ls_[3])};
    const float dl[4] = {__uint_as_float(dl_[0]), __uint_as_float(dl_[1]), __uint_as_float(dl_[2]), __uint_as_float(dl_[3])};
    float pt[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * c4 + e;
      float pe = ex2(fmaf(__uint_as_float(sv[c]), p.scale_log2, -ls[e]));
      if (!OPEN) {
        const int i = i0 + c;
        pe = (i < p.Tq && j < klen && (!p.causal || j <= i)) ? pe : 0.f;
      }
      const bool keep = (mk[e] & lane_bit) != 0u;
      pt[e] = keep ? pe : 0.f;
      const float dpk = keep ? __uint_as_float(dv[c]) : 0.f;
      ds[e] = pe * fmaf(dpk, p.drop_scale, -dl[e]);
    }
    pw[2 * c4] = pack_bf16(pt[0], pt[1]); pw[2 * c4 + 1] = pack_bf16(pt[2], pt[3]);
    dw[2 * c4] = pack_bf16(ds[0], ds[1]); dw[2 * c4 + 1] = pack_bf16(ds[2], ds[3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                  const __grid_constant__ CUtensorMap tmK,
                                                                  const __grid_constant__ CUtensorMap tmV,
                                                                  const __grid_constant__ CUtensorMap tmDO,
                                                                  const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  Bars* bars = reinterpret_cast<Bars*>(smem + oBar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.x * BKT, h = blockIdx.y, b = blockIdx.z;
  const long long bh = (long long)b * p.H + h;
  const int klen = p.key_len ? min(p.key_len[b], p.Tk) : p.Tk;
  const int q_begin = p.causal ? (j0 / BQT) * BQT : 0;   // queries before the first key of the block never see it
  const int n_it = (j0 < klen && q_begin < p.Tq) ? (p.Tq - q_begin + BQT - 1) / BQT : 0;

  if (tid == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&bars->q_full[s], 33);   // the TMA lane's expect_tx arrival + one cp.async arrival per producer lane
      mbar_init(&bars->q_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], kSoftWarps);
      mbar_init(&bars->p_empty[s], 1);
    }
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_empty, 4);
    mbar_init(&bars->acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
  }
  auto load_q = [&](int it) {   // Q / dO tiles of query block `it` -> ring stage it % kQStages (one lane)
    const int s = it % kQStages, qb = q_begin + it * BQT;
    uint64_t* bar = &bars->q_full[s];
    mbar_expect_tx(bar, 32768u);
    const uint32_t dst = sbase + oQ + s * kStageBytes;
    tma_2d(dst, &tmQ, h * DH, b * p.Tq + qb, bar);
    tma_2d(dst + 8192, &tmQ, h * DH + 64, b * p.Tq + qb, bar);
    tma_2d(dst + 16384, &tmDO, h * DH, b * p.Tq + qb, bar);
    tma_2d(dst + 24576, &tmDO, h * DH + 64, b * p.Tq + qb, bar);
  };
  if (tid == 0 && n_it > 0) {
    // the first tiles are requested before the CTA-wide sync: tensor-memory allocation and the loads' latency overlap
    mbar_expect_tx(&bars->kv_full, 65536u);
    tma_2d(sbase + oK, &tmK, h * DH, b * p.Tk + j0, &bars->kv_full);
    tma_2d(sbase + oV, &tmV, h * DH, b * p.Tk + j0, &bars->kv_full);
    load_q(0);
    tma_2d(sbase + oK + 16384, &tmK, h * DH + 64, b * p.Tk + j0, &bars->kv_full);
    tma_2d(sbase + oV + 16384, &tmV, h * DH + 64, b * p.Tk + j0, &bars->kv_full);
    for (int it = 1; it < kQStages && it < n_it; ++it) load_q(it);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_base;
  [[maybe_unused]] const bool tracing = p.trace == 1 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 2 + kSoftWarps || warp == kWarps - 1);

  if (warp == 0) {
    // ================= producer: K / V once, then Q / dO tiles and the per-query statistics, two stages =================
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kQStages, qb = q_begin + it * BQT;
      if (it >= kQStages) {   // the first kQStages tiles were requested in the prologue
        mbar_wait(&bars->q_empty[s], ((it / kQStages) & 1u) ^ 1u);
        TC_STAMP(2, 3);
        if (lane == 0) load_q(it);
      }
      // [64] lse | [64] delta | [4 key words][64] keep bits
      const uint32_t st = sbase + oStat + s * kStatBytes;
#pragma unroll
      for (int w = 0; w < 12; ++w) {
        const int idx = lane + 32 * w, which = idx >> 6, c = idx & 63, i = qb + c;
        bool ok = i < p.Tq;
        const void* src;
        if (which == 0) src = p.lse + bh * p.Tq + (ok ? i : 0);
        else if (which == 1) src = p.delta + bh * p.Tq + (ok ? i : 0);
        else {
          const int kw = (j0 >> 5) + which - 2;
          ok = ok && p.use_mask && kw < p.n_kw;
          src = ok ? (const void*)(p.keep_mask + (bh * p.n_kw + kw) * p.Tq + i) : (const void*)p.lse;
        }
        const int bytes = ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(st + idx * 4), "l"(src), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars->q_full[s])) : "memory");
      TC_STAMP(2, 4);
    }
  } else if (warp == 1) {
    // ================= MMA issuer A: S^T = K Q^T and dP^T = V dO^T of every block, as soon as its tile and a tensor-memory buffer exist =================
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idS = idesc(128, BQT, 0, 0);
      // base descriptors, built once: an operand slice is base + (byte offset >> 4) (the 14-bit address field cannot overflow:
      // shared-memory addresses are below 256 KB), so the code between two tcgen05.mma is one 64-bit add per operand
      const uint64_t dKk = make_desc(sbase + oK, 16, 1024), dVk = make_desc(sbase + oV, 16, 1024);      // K-major A: K / V tile
      const uint64_t dQk0 = make_desc(sbase + oQ, 16, 1024);                                           // K-major B: Q tile of stage 0 (dO: + 16384)
      mbar_wait(&bars->kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const uint32_t s = it & 1;
        TC_STAMP(1, 0);
        mbar_wait(&bars->q_full[it % kQStages], (it / kQStages) & 1u);
        TC_STAMP(1, 1);
        // buffer s of tensor memory held S^T / dP^T / P^T of block it - 2: free once that block's gradient products are done
        mbar_wait(&bars->p_empty[s], ((it >> 1) & 1u) ^ 1u);
        fence_after();
        const uint64_t qk = dQk0 + (uint64_t)(((it % kQStages) * kStageBytes) >> 4), dok = qk + (16384 >> 4);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16(tmem + cS + s * BQT, dKk + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4), qk + (((ks >> 2) * 8192 + (ks & 3) * 32) >> 4), idS, ks > 0);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16(tmem + cDP + s * BQT, dVk + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4), dok + (((ks >> 2) * 8192 + (ks & 3) * 32) >> 4), idS, ks > 0);
        umma_commit(&bars->s_full[s]);
        TC_STAMP(1, 2);
      }
    }
    __syncwarp();
  } else if (warp == kWarps - 1) {
    // ================= MMA issuer B: the three gradient products of a block, as soon as the softmax warps have written P^T / dS^T.
    // Its own thread, so that a late Q tile (issuer A waiting) never delays these products and the stage / buffer releases
    // that hang on their completion =================
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idKV = idesc(128, DH, 0, 1), idQ = idesc(128, BQT, 1, 1);
      const uint64_t dKm = make_desc(sbase + oK, 16384, 1024);                                         // MN-major A: K tile (dQ^T)
      const uint64_t dDSk0 = make_desc(sbase + oDS, 16, 1024);                                         // K-major A: dS^T, buffer 0
      const uint64_t dDSm0 = make_desc(sbase + oDS, 8192, 1024);                                       // MN-major B: dS^T (dQ^T), buffer 0
      const uint64_t dQm0 = make_desc(sbase + oQ, 8192, 1024);                                         // MN-major B: Q tile of stage 0 (dO: + 16384)
      mbar_wait(&bars->kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const uint32_t s = it & 1;
        TC_STAMP(1, 3);
        mbar_wait(&bars->p_full[s], (it >> 1) & 1u);
        TC_STAMP(1, 4);
        mbar_wait(&bars->dq_empty, (it & 1u) ^ 1u);
        fence_after();
        const uint64_t dDSk = dDSk0 + (uint64_t)((s * 16384) >> 4), dDSm = dDSm0 + (uint64_t)((s * 16384) >> 4);
        const uint64_t qm = dQm0 + (uint64_t)(((it % kQStages) * kStageBytes) >> 4), dom = qm + (16384 >> 4);
        const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
        for (int ks = 0; ks < BQT / 16; ++ks)   // dV += P^T dO
          umma_bf16_ta(tmem + cDV, tmem + cS + s * BQT + ks * 16, dom + ((ks * 2048) >> 4), idKV, ks > 0 ? 1u : acc);
#pragma unroll
        for (int ks = 0; ks < BQT / 16; ++ks)   // dK += dS^T Q
          umma_bf16(tmem + cDK, dDSk + ((ks * 32) >> 4), qm + ((ks * 2048) >> 4), idKV, ks > 0 ? 1u : acc);
#pragma unroll
        for (int ks = 0; ks < BKT / 16; ++ks)   // dQ^T = K^T dS^T
          umma_bf16(tmem + cDQ, dKm + ((ks * 2048) >> 4), dDSm + ((ks * 2048) >> 4), idQ, ks > 0 ? 1u : 0u);
        umma_commit(&bars->p_empty[s]);
        umma_commit(&bars->q_empty[it % kQStages]);
        umma_commit(&bars->dq_full);
        TC_STAMP(1, 5);
      }
      umma_commit(&bars->acc_full);
    }
    __syncwarp();
  } else if (warp < 2 + kSoftWarps) {
    // ================= softmax warps: thread = key row (TMEM lane quadrant = warp % 4), CPT query columns per thread =================
    const int quad = warp & 3, cg = (warp - 2) >> 2;
    const int r = quad * 32 + lane, j = j0 + r;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1, qb = q_begin + it * BQT;
      const int qst = it % kQStages;
      TC_STAMP(0, 0);
      mbar_wait(&bars->q_full[qst], (it / kQStages) & 1u);   // the statistics of this block are in shared memory
      TC_STAMP(0, 1);
      mbar_wait(&bars->s_full[s], (it >> 1) & 1u);
      fence_after();
      TC_STAMP(0, 2);
      uint32_t pw[CPT / 2], dw[CPT / 2];
      // a warp whose 32 key rows no query of this block can see (beyond the sample's key length - two valid keys in the
      // last 128-key block of the 258-token cross-attention - or above the causal diagonal) hands over zeros without
      // reading the logits: the kernel is bound by these warps' instruction issue
      const bool dead_warp = j0 + quad * 32 >= klen || (p.causal && j0 + quad * 32 > qb + BQT - 1);
      if (dead_warp) {
#pragma unroll
        for (int e = 0; e < CPT / 2; ++e) pw[e] = dw[e] = 0u;
      } else {
        uint32_t sv[CPT], dv[CPT];
        tmem_ld16(lane_addr + cS + s * BQT + cg * CPT, sv);
        tmem_ld16(lane_addr + cDP + s * BQT + cg * CPT, dv);
        tmem_wait_ld();
        TC_STAMP(0, 3);
        const uint32_t stl = sbase + oStat + qst * kStatBytes + cg * CPT * 4;
        // interior tiles: every query exists and sees every key of the tile
        const bool open = qb + BQT <= p.Tq && j0 + BKT <= klen && (!p.causal || j0 + BKT - 1 <= qb);
        if (open) soft_block<true>(p, sv, dv, stl, quad, lane, 0, 0, 0, pw, dw);
        else soft_block<false>(p, sv, dv, stl, quad, lane, qb + cg * CPT, j, klen, pw, dw);
      }
      TC_STAMP(0, 4);
      // P^T goes back into tensor memory, over this thread's own (already loaded) S^T columns: columns 16 cg .. 16 cg + 7 of the
      // buffer hold queries 16 cg .. 16 cg + 15 as bf16 pairs = the A operand of k-step cg of dV += P^T dO.  dS^T goes to
      // shared-memory buffer s.  Neither needs a wait: S^T of this block was only issued once block it - 2 had released both.
      tmem_st8(lane_addr + cS + s * BQT + cg * CPT, pw);
      const uint32_t rowD = sbase + oDS + s * 16384 + r * 128;
#pragma unroll
      for (int c8 = 0; c8 < CPT / 8; ++c8) {   // 16-byte chunks of the 128-byte row, XOR-swizzled with the row (SWIZZLE_128B)
        const uint32_t off = (uint32_t)(((cg * (CPT / 8) + c8) ^ (r & 7)) << 4);
        st_shared_v4(rowD + off, dw[4 * c8], dw[4 * c8 + 1], dw[4 * c8 + 2], dw[4 * c8 + 3]);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[s]);
      TC_STAMP(0, 6);
    }
    // dV (first half of the warps) / dK (second half) of this thread's key row -> bf16; each warp takes kEpiCols columns
    if (n_it > 0) {
      mbar_wait(&bars->acc_full, 0);
      fence_after();
    }
    constexpr int kEpiCols = 2 * DH / (kSoftWarps / 4);   // 48
    const int col0 = cg * kEpiCols;                        // column of the 192-wide [dV | dK] accumulator pair
    const bool is_dv = col0 < DH;
    const int c_in = is_dv ? col0 : col0 - DH;
    __nv_bfloat16* dst = is_dv ? p.dv + ((long long)b * p.Tk + j) * p.lddv + h * DH + c_in : p.dk + ((long long)b * p.Tk + j) * p.lddk + h * DH + c_in;
    const float f = is_dv ? p.drop_scale : p.scale;   // dV = P_drop^T dO / (1 - p);  dK = head_dim^-0.5 dS^T Q
#pragma unroll
    for (int c0 = 0; c0 < kEpiCols; c0 += 16) {
      uint32_t v[16];
      if (n_it > 0) {
        tmem_ld16(lane_addr + cDV + col0 + c0, v);   // dK follows dV in tensor memory
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = 0u;
      }
      if (j < p.Tk) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[8 * q + 0]) * f, __uint_as_float(v[8 * q + 1]) * f);
          o.y = pack_bf16(__uint_as_float(v[8 * q + 2]) * f, __uint_as_float(v[8 * q + 3]) * f);
          o.z = pack_bf16(__uint_as_float(v[8 * q + 4]) * f, __uint_as_float(v[8 * q + 5]) * f);
          o.w = pack_bf16(__uint_as_float(v[8 * q + 6]) * f, __uint_as_float(v[8 * q + 7]) * f);
          *reinterpret_cast<uint4*>(dst + c0 + 8 * q) = o;
        }
      }
    }
  } else {
    // ================= dQ warps: lane = head column d (TMEM lane), 64 query columns; coalesced fp32 reductions =================
    const int quad = warp & 3, d = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + cDQ;
    const long long ldacc = (long long)p.H * DH;
    float* accg = p.dq_acc + (long long)b * p.Tq * ldacc + h * DH + d;
    for (int it = 0; it < n_it; ++it) {
      const int qb = q_begin + it * BQT;
      TC_STAMP(2, 0);
      mbar_wait(&bars->dq_full, it & 1u);
      fence_after();
      TC_STAMP(2, 1);
      uint32_t v0[32], v1[32];
      if (quad < 3) {
        tmem_ld32(lane_addr, v0);
        tmem_ld32(lane_addr + 32, v1);
        tmem_wait_ld();
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dq_empty);
      TC_STAMP(2, 2);
      if (quad < 3) {   // unscaled: head_dim^-0.5 is applied by the fp32 -> bf16 conversion kernel
        float* dst = accg + (long long)qb * ldacc;
        const int nq = min(BQT, p.Tq - qb);
        if (nq == BQT && ldacc == 8 * DH) {   // the decoder's 8 heads: constant row stride, immediate offsets
#pragma unroll
          for (int c = 0; c < 32; ++c) red_add_f32(dst + c * 8 * DH, __uint_as_float(v0[c]));
#pragma unroll
          for (int c = 0; c < 32; ++c) red_add_f32(dst + (32 + c) * 8 * DH, __uint_as_float(v1[c]));
        } else if (nq == BQT) {
#pragma unroll
          for (int c = 0; c < 32; ++c, dst += ldacc) red_add_f32(dst, __uint_as_float(v0[c]));
#pragma unroll
          for (int c = 0; c < 32; ++c, dst += ldacc) red_add_f32(dst, __uint_as_float(v1[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nq) red_add_f32(dst + (long long)c * ldacc, __uint_as_float(v0[c]));
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (32 + c < nq) red_add_f32(dst + (long long)(32 + c) * ldacc, __uint_as_float(v1[c]));
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace tc

// single-pass backward on tcgen05 (head_dim 96): delta / dq_acc prologue, the kernel above, fp32 -> bf16 dQ conversion
static int launch_bwd_tc(const Args& a, cudaStream_t s) {
  constexpr int DH = tc::DH;
  CUtensorMap mq, mk, mv, mdo;
  const long long width = (long long)a.H * DH;
  int rc;
  if ((rc = make_tma_map_bf16(&mq, a.q, width, (long long)a.B * a.Tq, a.ldq, tc::BQT))) return rc;
  if ((rc = make_tma_map_bf16(&mdo, a.d_o, width, (long long)a.B * a.Tq, a.lddo, tc::BQT))) return rc;
  if ((rc = make_tma_map_bf16(&mk, a.k, width, (long long)a.B * a.Tk, a.ldk, tc::BKT))) return rc;
  if ((rc = make_tma_map_bf16(&mv, a.v, width, (long long)a.B * a.Tk, a.ldv, tc::BKT))) return rc;
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk; p.causal = a.causal;
  p.use_mask = a.drop_thresh != 0u ? 1 : 0; p.n_kw = a.n_kw;
  p.scale_log2 = a.scale_log2; p.scale = a.scale; p.drop_scale = a.drop_scale;
  p.key_len = a.key_len; p.lse = a.lse; p.delta = a.delta; p.keep_mask = a.keep_mask;
  p.dk = a.dk; p.dv = a.dv; p.lddk = a.lddk; p.lddv = a.lddv; p.dq_acc = a.dq_acc;
  static const bool trace_on = getenv("TTS_ATTN_TC_TRACE") != nullptr && atoi(getenv("TTS_ATTN_TC_TRACE")) != 0;
  p.trace = trace_on ? atoi(getenv("TTS_ATTN_TC_TRACE")) : 0;
  static std::atomic<unsigned long long> attr{0ull};   // per device: cudaFuncSetAttribute applies to the current device only
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long dev_bit = 1ull << (dev & 63);
  if (!(attr.load(std::memory_order_acquire) & dev_bit)) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(tc::attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmem));
    attr.fetch_or(dev_bit, std::memory_order_release);
  }
  const long long n = (long long)a.B * a.Tq * a.H;
  TTS_CHECK_CUDA(cudaMemsetAsync(a.dq_acc, 0, (size_t)n * DH * sizeof(float), s));
  attn_bwd_prep_kernel<DH><<<(unsigned)((n + 63) / 64), 256, 0, s>>>(a);
  TTS_CHECK_LAUNCH();
  tc::attn_bwd_tc_kernel<<<dim3(ceil_div(a.Tk, tc::BKT), a.H, a.B), tc::kThreads, tc::kSmem, s>>>(mq, mk, mv, mdo, p);
  TTS_CHECK_LAUNCH();
  const long long rows = (long long)a.B * a.Tq;
  const long long work = rows * (a.H * DH / 8);
  attn_bwd_dq_convert_kernel<<<(unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16), 256, 0, s>>>(
      a.dq_acc, a.dq, a.lddq, rows, a.H * DH, a.scale);
  TTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace attn
}  // namespace tts
