// bf16 tcgen05 GEMM of the teacher-forced TRAINING path (forward, dgrad and wgrad of every nn.Linear / nn.Conv1d of
// transformer/attention.py:43-47, modules.py:11-19, tacotron.py:50-52,78,103-105 - the reference runs them as fp32
// cuBLAS / cuDNN calls through autograd).
//
//   C[M,N] = epilogue( sum_k A[m,k] * B[n,k] ),  A and B bf16, fp32 accumulation in tensor memory.
//
// Blackwell-native structure (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d tiles (128-byte swizzle) into a 4/6-stage shared-memory ring,
//              completion on mbarriers (SASS: UTMALDG)
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16; SASS UTCHMMA)
//              from shared-memory descriptors, tcgen05.commit releases ring stages and publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld (LDTM) 32 lanes x 32 columns at a time, fused bias / ReLU / Philox dropout /
//              residual / length mask / ReLU-gate (backward) / bf16 or fp32 store / split-K fp32 reduction
//   tensor memory holds TWO accumulators (2 x BN columns), so the epilogue of tile i overlaps the MMAs of tile i+1.
//   CTA PAIRS (thread-block clusters of 2 along M): the two CTAs work on vertically adjacent output tiles, which share the
//   B operand; each CTA fetches HALF of the B tile and TMA-multicasts it into both CTAs' shared memory
//   (.multicast::cluster), so the L2 -> SM traffic per k-block drops from 48 KB to 32 KB per CTA (a 128 x 256 tile at the
//   tensor peak needs ~28 TB/s of operand traffic chip-wide without sharing - more than the L2 delivers).  A ring stage is
//   released to BOTH producers by a multicast tcgen05.commit.
// Operands may be K-major (row-major [rows][K], the forward layout) or MN-major (row-major [K][rows]): dgrad reads the
// weight [N][K] as the MN-major B operand of dX = dY W, wgrad reads dY and X as MN-major operands of dW = dY^T X, so
// no transposed copy of a weight or an activation is ever written.  A k5 Conv1d runs as 5 accumulated taps over a
// zero-padded channels-last buffer (the producer shifts the A row coordinate per tap).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "philox.cuh"

namespace tts {
namespace bf {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;   // two warps per tensor-memory lane quadrant; each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kABytes = BM * BK * 2;   // 16 KB
constexpr long long kTimeout = 2LL << 30;

// MODE 0: one CTA per tile (cta_group::1).  MODE 1: CTA pair, cta_group::1 MMAs, B tile TMA-multicast into both CTAs.
// MODE 2: CTA pair, ONE cta_group::2 MMA (M = 256) per k-step issued by the even CTA: each CTA keeps its own 128 rows of A
// and HALF of the B tile in shared memory (a stage is 32 KB instead of 48, six stages instead of four).
template <int BN, int MODE = 0>
struct Cfg {
  static constexpr int kBBytes = BN * BK * 2;                      // whole B tile
  static constexpr int kBStage = MODE == 2 ? kBBytes / 2 : kBBytes;   // B bytes resident per CTA and stage
  static constexpr int kStage = kABytes + kBStage;
  static constexpr int kStages = MODE == 2 ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : 6);
  static constexpr int kTmemCols = 2 * BN;
};

struct Params {
  int M, N;
  int taps, kb_per_tap, tap_b_stride;   // contraction = taps x kb_per_tap k-blocks; tap t: A rows + t, B k + t * tap_b_stride
  int a_mn, b_mn;                       // operand stored MN-major ([K][rows])
  int split_k, n_mblk, n_nblk;
  void* C; long long ldc; int out_bf16;
  const float* bias; int act; float alpha;
  const float* residual; long long ldr;
  float drop_scale; uint32_t drop_thresh; unsigned long long seed; uint32_t stream;
  const __nv_bfloat16* gate; long long ldg; float gate_scale;
  const int32_t* row_len; int rows_per_batch, valid_rows, out_rows_per_batch, out_row_offset;
};

__device__ int g_err;   // sticky: a barrier wait timed out (read by tts_gemm_bf16_status)

struct Shared {
  uint64_t full[8], empty[8], tfull[2], tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity) {   // false on timeout: never hangs the GPU
  long long t0 = 0, spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if ((++spins & 1023) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeout) {
        atomicExch(&g_err, 1);
        return false;
      }
    }
  }
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// cta_group::2 forms: the TMA of either CTA of the pair completes its bytes on the EVEN CTA's barrier (`bar` is a
// shared::cluster address obtained with mapa), the MMA spans both CTAs' shared and tensor memory, the commit arrives on the
// same barrier offset in both CTAs.
__device__ __forceinline__ void tma_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope): the explicit .release.cluster form compiled to an ERRBAR that held the warp
  // until its global stores of the tile had drained (16 % of the kernel's stall samples, profiles/r2_gemm_relu_drop_sass_top.txt)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor; canonical layouts
// in cute/atom/mma_traits_sm100.hpp).  K-major: rows of 128 bytes (64 bf16 of K), 8-row atoms SBO = 1024 bytes apart.
// MN-major: rows of 128 bytes hold 64 consecutive MN elements of one k, 8-k atoms SBO = 1024 bytes apart, the next 64
// MN elements LBO = 64 k-rows x 128 bytes = 8192 bytes away.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version (sm_100)
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 256-bit global store (sm_100: STG.256), address 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}

__device__ __forceinline__ void ld_global_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

struct Unit {
  int m0, n0, kb0, kb1;
};
// u: unit index of the CTA (CL = 1) or of the CTA pair (CL = 2: rows of m-block pairs; `rank` picks the m-block of the pair)
template <int CL>
__device__ __forceinline__ Unit decode_unit(const Params& p, int u, int rank) {
  const int n_mrows = (p.n_mblk + CL - 1) / CL;
  const int n_tiles = n_mrows * p.n_nblk;
  const int split = u / n_tiles, tile = u - split * n_tiles;   // split slowest: concurrent CTAs share operand slices in L2
  const int mr = tile / p.n_nblk, nb = tile - mr * p.n_nblk;
  const int mb = mr * CL + rank;
  const int total = p.taps * p.kb_per_tap, per = (total + p.split_k - 1) / p.split_k;
  Unit r;
  r.m0 = mb * BM;
  r.kb0 = split * per;
  r.kb1 = min(total, r.kb0 + per);
  r.n0 = nb;   // scaled by BN by the caller
  return r;
}

template <int BN, int MODE>
__device__ __forceinline__ void gemm_bf16_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p) {
  using C = Cfg<BN, MODE>;
  constexpr int CL = MODE == 0 ? 1 : 2;
  constexpr bool TWO = MODE == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Shared* sh = reinterpret_cast<Shared*>(smem + C::kStages * C::kStage);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_units = ((p.n_mblk + CL - 1) / CL) * p.n_nblk * p.split_k;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int u0 = (int)blockIdx.x / CL, ustep = (int)gridDim.x / CL;   // the CTAs of a pair walk the same units

  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&sh->full[s], 1);
      // MODE 1: a stage is refilled by multicast, BOTH CTAs' MMAs must have read it.  MODE 2: one cta_group::2 commit
      mbar_init(&sh->empty[s], TWO ? 1 : CL);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->tfull[s], 1);
      mbar_init(&sh->tempty[s], TWO ? 2 * kEpiWarps : kEpiWarps);   // MODE 2: both CTAs' epilogues free the pair's accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {   // the MMA warp owns the tensor-memory allocation (MODE 2: the same warp of both CTAs, collectively)
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)C::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)C::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into this CTA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = sh->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int u = u0; u < n_units && ok; u += ustep) {
        const Unit un = decode_unit<CL>(p, u, rank);
        const int n0 = un.n0 * BN;
        for (int kb = un.kb0; kb < un.kb1 && ok; ++kb, ++it) {
          const int s = it % C::kStages;
          ok = mbar_wait(&sh->empty[s], ((it / C::kStages) & 1u) ^ 1u);
          uint64_t* bar = &sh->full[s];
          const uint32_t a_dst = smem_u32(smem + s * C::kStage), b_dst = a_dst + kABytes;
          const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
          const int ka = kc * BK, kbc = tap * p.tap_b_stride + kc * BK;
          if (TWO) {   // both CTAs' tiles complete on the even CTA's barrier, which expects the bytes of the pair
            if (rank == 0) mbar_expect_tx(bar, 2u * (unsigned)C::kStage);
            const uint32_t lbar = mapa_rank(smem_u32(bar), 0u);
            if (!p.a_mn) {
              tma_2d_2sm(a_dst, &tmA, ka, un.m0 + tap, lbar);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_2d_2sm(a_dst + j * 8192, &tmA, un.m0 + 64 * j, ka, lbar);
            }
            if (!p.b_mn) {   // my half of the tile's B rows, at the start of my B region
              tma_2d_2sm(b_dst, &tmB, kbc, n0 + rank * (BN / 2), lbar);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) tma_2d_2sm(b_dst + j * 8192, &tmB, n0 + rank * (BN / 2) + 64 * j, kbc, lbar);
            }
            continue;
          }
          mbar_expect_tx(bar, (unsigned)C::kStage);
          if (!p.a_mn) {
            tma_2d(a_dst, &tmA, ka, un.m0 + tap, bar);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_2d(a_dst + j * 8192, &tmA, un.m0 + 64 * j, ka, bar);
          }
          if (CL == 1) {
            if (!p.b_mn) {
              tma_2d(b_dst, &tmB, kbc, n0, bar);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_2d(b_dst + j * 8192, &tmB, n0 + 64 * j, kbc, bar);
            }
          } else {   // my half of the B tile, multicast into both CTAs of the pair (same offsets, each CTA's own barrier)
            if (!p.b_mn) {
              tma_2d_mc(b_dst + rank * (C::kBBytes / 2), &tmB, kbc, n0 + rank * (BN / 2), bar, (uint16_t)0x3);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = rank * (BN / 128) + j;
                tma_2d_mc(b_dst + jj * 8192, &tmB, n0 + 64 * jj, kbc, bar, (uint16_t)0x3);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (MODE 2: the even CTA issues for the pair) =================
    if (lane == 0 && (!TWO || rank == 0)) {
      // kind::f16 instruction descriptor: fp32 accumulate, bf16 x bf16, per-operand major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn ? 1 : 0) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((TWO ? 2 * BM : BM) >> 4) << 24);
      const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;   // bytes per K = 16
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      uint32_t it = 0, lu = 0;
      bool ok = true;
      for (int u = u0; u < n_units && ok; u += ustep, ++lu) {
        const Unit un = decode_unit<CL>(p, u, rank);
        const uint32_t as = lu & 1u;
        ok = mbar_wait(&sh->tempty[as], ((lu >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_d + as * BN;
        for (int kb = un.kb0; kb < un.kb1 && ok; ++kb, ++it) {
          const int s = it % C::kStages;
          ok = mbar_wait(&sh->full[s], (it / C::kStages) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_base = smem_u32(smem + s * C::kStage), b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_desc(a_base + k * a_step, a_lbo, 1024u), db = make_desc(b_base + k * b_step, b_lbo, 1024u);
            if (TWO) umma_bf16_2sm(acc, da, db, idesc, (kb > un.kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(acc, da, db, idesc, (kb > un.kb0 || k > 0) ? 1u : 0u);
          }
          if (TWO) umma_commit_2sm(&sh->empty[s]);   // the stage may be refilled once these MMAs have read it, in both CTAs
          else if (CL == 1) umma_commit(&sh->empty[s]);
          else umma_commit_mc(&sh->empty[s], (uint16_t)0x3);
        }
        if (TWO) umma_commit_2sm(&sh->tfull[as]);   // accumulator complete: both CTAs' epilogues read their 128 rows
        else umma_commit(&sh->tfull[as]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps: TMEM lane quadrant = warp % 4, thread = output row; warps 2-5 take the first half
    // of the tile's columns, warps 6-9 the second.  Global accesses are 32 bytes per lane (LDG/STG.256: one full sector
    // per lane and instruction; with 16-byte accesses the fp32-output GEMMs of K = 768 were bound by the store requests of
    // the epilogue, profiles/r2_gemm_bf16_sweep.txt) and the residual / gate operands of chunk c+1 are requested before
    // chunk c is processed (their latency was exposed once per chunk) =================
    const int quad = warp & 3;
    const int c_lo = ((warp - 2) >> 2) * (BN / 2), c_hi = c_lo + BN / 2;
    const int rpb = p.rows_per_batch > 0 ? p.rows_per_batch : p.M;
    const int valid = p.valid_rows > 0 ? p.valid_rows : rpb;
    const int orpb = p.out_rows_per_batch > 0 ? p.out_rows_per_batch : rpb;
    const bool res_vec = p.residual != nullptr && (p.ldr & 7) == 0;
    const bool gate_vec = p.gate != nullptr && (p.ldg & 15) == 0;
    uint32_t lu = 0;
    bool ok = true;
    for (int u = u0; u < n_units; u += ustep, ++lu) {
      const Unit un = decode_unit<CL>(p, u, rank);
      const int n0 = un.n0 * BN;
      const uint32_t as = lu & 1u;
      const int m = un.m0 + quad * 32 + lane;
      const int b = m / rpb, r = m - b * rpb;
      const bool row_in = m < p.M && r < valid && un.kb1 > un.kb0;
      const size_t orow = (size_t)b * orpb + r + p.out_row_offset;
      float rn[32];       // residual of the next chunk
      uint32_t gn[16];    // gate (32 bf16) of the next chunk
      bool pre_n = false;
      auto prefetch = [&](int c0) {   // 32-column chunk c0 of this row, if it is complete: vector loads, else the scalar path
        pre_n = row_in && n0 + c0 + 32 <= p.N && (res_vec || gate_vec);
        if (!pre_n) return;
        if (res_vec) {
          const float* rp = p.residual + orow * p.ldr + n0 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) ld_global_v8(rp + 8 * q, *reinterpret_cast<float(*)[8]>(&rn[8 * q]));
        }
        if (gate_vec) {
          const __nv_bfloat16* gp = p.gate + orow * p.ldg + n0 + c0;
#pragma unroll
          for (int q = 0; q < 2; ++q) ld_global_v8(reinterpret_cast<const float*>(gp + 16 * q), *reinterpret_cast<float(*)[8]>(&gn[8 * q]));
        }
      };
      if (n0 + c_lo < p.N) prefetch(c_lo);   // in flight while the MMAs of this tile finish
      if (ok) ok = mbar_wait(&sh->tfull[as], (lu >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool row_ok = ok && row_in;
      const bool dead = row_ok && p.row_len != nullptr && r >= p.row_len[b];
      for (int c0 = c_lo; c0 < c_hi && n0 + c0 < p.N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + as * BN + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        float rc[32];
        uint32_t gc[16];
        const bool pre_c = pre_n;
#pragma unroll
        for (int j = 0; j < 32; ++j) rc[j] = rn[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) gc[j] = gn[j];
        if (c0 + 32 < c_hi && n0 + c0 + 32 < p.N) prefetch(c0 + 32);
        else pre_n = false;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
        const int n = n0 + c0;
        const bool full = n + 32 <= p.N;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n + j < p.N) x[j] += __ldg(p.bias + n + j);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (p.gate != nullptr) {   // backward of ReLU (+ dropout): the saved forward output is > 0 exactly where both kept it
          const __nv_bfloat16* gp = p.gate + orow * p.ldg + n;
          if (pre_c && gate_vec) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              // bf16 > 0: sign clear and magnitude non-zero
              const uint32_t lo = gc[e] & 0xffffu, hi = gc[e] >> 16;
              x[2 * e] = (lo != 0u && lo < 0x8000u) ? x[2 * e] * p.gate_scale : 0.f;
              x[2 * e + 1] = (hi != 0u && hi < 0x8000u) ? x[2 * e + 1] * p.gate_scale : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) x[j] = __bfloat162float(gp[j]) > 0.f ? x[j] * p.gate_scale : 0.f;
          }
        }
        if (p.drop_thresh != 0u) {   // nn.Dropout on the product (before the residual add: modules.py:132,138,141)
          const unsigned long long e0 = (unsigned long long)orow * (unsigned long long)p.N + (unsigned long long)n;
#pragma unroll
          for (int q = 0; q < 4; ++q) {   // one Philox call per 8 outputs: 16-bit lanes (philox.cuh)
            const uint4 rw = dropout_words_linear(p.seed, p.stream, (e0 >> 3) + q);
            const uint32_t ws[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[8 * q + 2 * e] = (ws[e] & 0xffffu) >= p.drop_thresh ? x[8 * q + 2 * e] * p.drop_scale : 0.f;
              x[8 * q + 2 * e + 1] = (ws[e] >> 16) >= p.drop_thresh ? x[8 * q + 2 * e + 1] * p.drop_scale : 0.f;
            }
          }
        }
        if (p.residual != nullptr) {
          const float* rp = p.residual + orow * p.ldr + n;
          if (pre_c && res_vec) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] += rc[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) x[j] += rp[j];
          }
        }
        if (dead) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.f;
        }
        if (p.split_k > 1) {   // fp32 reduction of the K splits into a zero-initialised C
          float* cp = reinterpret_cast<float*>(p.C) + orow * p.ldc + n;
          if (full && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0) {   // 16-byte vector reductions: a quarter of the L2 atomic operations
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * q), "f"(x[4 * q]), "f"(x[4 * q + 1]),
                           "f"(x[4 * q + 2]), "f"(x[4 * q + 3]) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) atomicAdd(cp + j, x[j]);
          }
        } else if (p.out_bf16) {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + n;
          if (full && (p.ldc & 15) == 0) {   // 32-byte stores: one full sector per lane and instruction
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = pack_bf16x2(x[16 * q + 2 * e], x[16 * q + 2 * e + 1]);
              st_global_v8(cp + 16 * q, o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) cp[j] = __float2bfloat16_rn(x[j]);
          }
        } else {
          float* cp = reinterpret_cast<float*>(p.C) + orow * p.ldc + n;
          if (full && (p.ldc & 7) == 0) {   // 32-byte stores: one full sector per lane and instruction
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(x[8 * q + e]);
              st_global_v8(cp + 8 * q, o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < p.N) cp[j] = x[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {   // this warp's quadrant of the accumulator is free again (MODE 2: tell the issuing CTA)
        if (TWO) mbar_arrive_cluster(mapa_rank(smem_u32(&sh->tempty[as]), 0u));
        else mbar_arrive(&sh->tempty[as]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // nobody leaves while the peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)C::kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ Params p) {
  gemm_bf16_body<BN, 0>(tmA, tmB, p);
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ Params p) {
  gemm_bf16_body<BN, 1>(tmA, tmB, p);
}
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ Params p) {
  gemm_bf16_body<BN, 2>(tmA, tmB, p);
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static std::atomic<EncodeTiledFn> fn{nullptr};
  EncodeTiledFn f = fn.load(std::memory_order_acquire);
  if (f == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && ptr != nullptr) {
      f = reinterpret_cast<EncodeTiledFn>(ptr);
      fn.store(f, std::memory_order_release);
    }
  }
  return f;
}

// operand stored row-major [outer][inner] with `ld` elements between rows; box = [box_outer][64 inner elements]
static int make_map(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_outer) {
  EncodeTiledFn f = encode_fn();
  TTS_REQUIRE(f != nullptr, "gemm_bf16: cuTensorMapEncodeTiled is not available from this driver");
  TTS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0,
              "gemm_bf16: operands need 16-byte aligned base pointers and row strides (ld=%lld)", ld);
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult rc = f(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TTS_REQUIRE(rc == CUDA_SUCCESS, "gemm_bf16: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)rc,
              inner, outer, ld);
  return 0;
}

static int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// TTS_GEMM_PAIRS: 0 = one CTA per tile, 1 = CTA pairs with a multicast B tile (cta_group::1), 2 (default) = cta_group::2 pairs
static int pair_mode(int n_mblk) {
  static const int mode = getenv("TTS_GEMM_PAIRS") != nullptr ? atoi(getenv("TTS_GEMM_PAIRS")) : 2;
  return n_mblk >= 2 ? (mode < 0 || mode > 2 ? 2 : mode) : 0;
}
static bool use_pairs(int n_mblk) { return pair_mode(n_mblk) != 0; }

template <int BN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t s) {
  using C = Cfg<BN>;
  const size_t smem = (size_t)C::kStages * C::kStage + sizeof(Shared) + 1024 + 64;
  const size_t smem2 = (size_t)Cfg<BN, 2>::kStages * Cfg<BN, 2>::kStage + sizeof(Shared) + 1024 + 64;
  static std::atomic<unsigned long long> configured{0ull};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TTS_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TTS_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_2sm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    configured.fetch_or(bit, std::memory_order_release);
  }
  const int mode = pair_mode(p.n_mblk);
  if (mode != 0) {   // CTA pairs: one cta_group::2 MMA per pair, or cta_group::1 MMAs with a TMA-multicast B tile
    const int units = ((p.n_mblk + 1) / 2) * p.n_nblk * p.split_k;
    const int pairs = sm_count() / 2;
    const int grid = 2 * (units < pairs ? units : pairs);
    if (mode == 2) gemm_bf16_2sm_kernel<BN><<<grid, kThreads, smem2, s>>>(ma, mb, p);
    else gemm_bf16_pair_kernel<BN><<<grid, kThreads, smem, s>>>(ma, mb, p);
    TTS_CHECK_LAUNCH();
    return 0;
  }
  const int units = p.n_mblk * p.n_nblk * p.split_k;
  const int grid = units < sm_count() ? units : sm_count();
  gemm_bf16_kernel<BN><<<grid, kThreads, smem, s>>>(ma, mb, p);
  TTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace bf

// 2-D bf16 tensor map, 128-byte swizzle, box = [box_outer rows][64 elements] (also used by the tcgen05 attention kernels)
int make_tma_map_bf16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_outer) {
  return bf::make_map(map, ptr, inner, outer, ld, box_outer);
}
}  // namespace tts

using namespace tts;

extern "C" int tts_gemm_bf16(const TtsGemmBf16* g, void* stream) {
  TTS_REQUIRE(g != nullptr && g->A && g->B && g->C, "gemm_bf16: null operand");
  TTS_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm_bf16: empty problem M=%d N=%d K=%d", g->M, g->N, g->K);
  const int taps = g->taps > 0 ? g->taps : 1;
  TTS_REQUIRE(taps == 1 || (!g->a_mn_major && !g->b_mn_major), "gemm_bf16: taps need K-major operands");
  bf::Params p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N;
  p.taps = taps;
  p.kb_per_tap = (g->K + bf::BK - 1) / bf::BK;
  p.tap_b_stride = g->K;
  p.a_mn = g->a_mn_major ? 1 : 0;
  p.b_mn = g->b_mn_major ? 1 : 0;
  const int bn = g->N > 128 ? 256 : 128;
  p.n_mblk = (g->M + bf::BM - 1) / bf::BM;
  p.n_nblk = (g->N + bn - 1) / bn;
  const int total_kb = taps * p.kb_per_tap;
  int split = g->split_k > 0 ? g->split_k : 1;
  if (split > total_kb) split = total_kb;
  while (split > 1 && ((total_kb + split - 1) / split) * (split - 1) >= total_kb) --split;   // no empty split
  p.split_k = split;
  TTS_REQUIRE(split == 1 || (!g->out_bf16 && !g->bias && !g->residual && g->act == 0 && g->drop_p == 0.f && !g->gate),
              "gemm_bf16: split-K supports only a plain fp32 reduction into a zeroed C");
  p.C = g->C; p.ldc = g->ldc; p.out_bf16 = g->out_bf16;
  p.bias = g->bias; p.act = g->act; p.alpha = g->alpha == 0.f ? 1.f : g->alpha;
  p.residual = g->residual; p.ldr = g->ldr;
  if (g->drop_p > 0.f) {
    TTS_REQUIRE(g->drop_p < 1.f && g->N % 8 == 0, "gemm_bf16: dropout needs p < 1 and N %% 8 == 0");
    p.drop_thresh = drop_threshold16(g->drop_p);
    p.drop_scale = 1.f / (1.f - g->drop_p);
    p.seed = g->seed; p.stream = g->rng_stream;
  }
  p.gate = reinterpret_cast<const __nv_bfloat16*>(g->gate); p.ldg = g->ldg; p.gate_scale = g->gate_scale == 0.f ? 1.f : g->gate_scale;
  p.row_len = g->row_len; p.rows_per_batch = g->rows_per_batch; p.valid_rows = g->valid_rows;
  p.out_rows_per_batch = g->out_rows_per_batch; p.out_row_offset = g->out_row_offset;

  // A: M rows (+ taps - 1 for the shifted conv reads), contraction K;  B: N rows, contraction taps * K
  CUtensorMap ma, mb;
  int rc;
  const long long a_rows = g->a_rows > 0 ? g->a_rows : (long long)g->M + taps - 1;
  if (!p.a_mn) rc = bf::make_map(&ma, g->A, g->K, a_rows, g->lda, bf::BM);
  else rc = bf::make_map(&ma, g->A, g->M, g->K, g->lda, 64);
  if (rc) return rc;
  // K-major B: one box per CTA = the whole tile, or (CTA pairs) the half of the tile that this CTA multicasts
  if (!p.b_mn) rc = bf::make_map(&mb, g->B, (long long)taps * g->K, g->N, g->ldb, bf::use_pairs(p.n_mblk) ? bn / 2 : bn);
  else rc = bf::make_map(&mb, g->B, g->N, g->K, g->ldb, 64);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bn == 256 ? bf::launch<256>(ma, mb, p, s) : bf::launch<128>(ma, mb, p, s);
}

extern "C" int tts_gemm_bf16_status(void) {
  int v = 0;
  if (cudaMemcpyFromSymbol(&v, bf::g_err, sizeof(int)) != cudaSuccess) return -1;
  return v;
}
