// Attention forward of the teacher-forced TRAINING path on the 5th-generation tensor cores (tcgen05 + tensor memory):
// softmax(Q K^T / sqrt(d)) V of transformer/attention.py:72-98 with the masks from indices / lengths and nn.Dropout on the
// weights (attention.py:89), head_dim 96.  Included by attn_train.cu after attn_bwd_tc.cuh (shares its PTX helpers).
//
// One CTA per (sample, head, 256 queries) = TWO 128-query tiles that share every K / V tile and take turns on the tensor
// core: while the four softmax warps of one tile work on S_w(j), the tensor core runs the products of the other tile.
//   S_w [128 q x 128 keys] = Q_w K_j^T     (K-major A = Q tile, K-major B = K tile; 6 k-steps of 16 over head_dim 96)
//   O_w [128 q x 96]      += P_w V_j       (A = P_w from TENSOR MEMORY, written over the consumed S_w columns as bf16 pairs;
//                                           MN-major B = V tile; 8 k-steps of 16 keys)
// fp32 accumulators in tensor memory: S_0 | S_1 | O_0 | O_1 = 128 + 128 + 96 + 96 columns.  Softmax: one THREAD per query
// row (tcgen05.ld 32 lanes x 32 columns), two passes over the row's 128 logits (maximum, then exp2 / sum / dropout / bf16
// pack), online rescaling of O_w in tensor memory only when a row's maximum moved.  The dropout keep bits are the documented
// Philox function of philox.cuh (one call covers rows {i, i+8}: lanes l and l ^ 8 of a warp hold exactly those rows, each
// computes half of the calls and they swap the other half with shuffles - one call per 8 weights) and are written to the
// keep-bit cache the tcgen05 backward reads (1 bit per weight).  A TMA warp streams K / V tiles through two-stage rings
// (cp.async.bulk.tensor.2d, 128-byte swizzle), one elected lane issues every tcgen05.mma.
#pragma once

namespace tts {
namespace attn {
namespace tcf {

using namespace tc;   // PTX helpers of attn_bwd_tc.cuh

constexpr int BQ = 128, BK = 128, kTiles = 2;
constexpr int kWarpsF = 8 * kTiles + 4;   // 8 softmax warps per query tile (warpgroups 0-3), then TMA, MMA issuer, two idle warps
constexpr int kTmaWarp = 8 * kTiles, kMmaWarp = kTmaWarp + 1;
constexpr int kThreadsF = kWarpsF * 32;
constexpr uint32_t oQf = 0, oKf = 65536, oVf = oKf + 2 * 32768, oXchF = oVf + 2 * 32768, oBarF = oXchF + kTiles * 2048;
constexpr size_t kSmemF = oBarF + 256 + 1024;
constexpr uint32_t cSf = 0, cOf = 256;   // tensor-memory columns: S_w at 128 w, O_w at 256 + 96 w

struct BarsF {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], o_done[2];
  uint32_t tmem_base;
};

struct ParamsF {
  int B, H, Tq, Tk, causal, n_kw;
  float scale_log2, drop_scale;
  uint32_t drop_thresh, stream;
  unsigned long long seed;
  const int32_t* key_len;
  float* lse;
  uint32_t* keep_mask;
  __nv_bfloat16* o;
  long long ldo;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
      "r"(v[30]), "r"(v[31]) : "memory");
}
// bounded wait with a suspend-time hint: the hardware parks the thread until the phase completes (or ~10 us pass) instead
// of the thread spinning through try_wait - the helper warps share their schedulers with softmax warps
__device__ __forceinline__ void wait_bar(uint64_t* bar, unsigned parity) {
  long long t0 = 0, spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(10000u) : "memory");
    if (ok) return;
    if ((++spins & 15) == 0) {
      if (*reinterpret_cast<volatile int*>(&g_err) != 0) return;
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeout) {
        atomicExch(&g_err, 1);
        return;
      }
    }
  }
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// keep bits of the 32 weights (row i, keys 16 jb0 .. 16 jb0 + 31): bit c <-> key 16 jb0 + c.  philox.cuh: call index
// ((bh n_iblk + i / 16) n_jblk + j / 16) * 32 + (i & 7) * 4 + t covers rows {i & ~8, i | 8} x keys {2t, 2t+1, 2t+8, 2t+9} of the
// 16 x 16 block; words (x, y) belong to the row with bit 3 clear, (z, w) to the other.  The lane with bit 3 clear computes
// t = 0, 1, its partner (lane ^ 8) t = 2, 3.
// two keep bits of one Philox word (two 16-bit lanes): bit 0 <-> low half >= thresh, bit 1 <-> high half >= thresh
// (half + add carries into bit 16  <=>  half >= thresh, add = 2^16 - thresh)
__device__ __forceinline__ uint32_t keep2(uint32_t x, uint32_t add) {
  const uint32_t lo = ((x & 0xffffu) + add) >> 16, hi = ((x >> 16) + add) >> 16;
  return lo + 2u * hi;
}
__device__ __forceinline__ uint32_t keep_bits32(const ParamsF& p, unsigned long long blk_idx, int i, int lane) {
  const bool hi = (lane & 8) != 0;
  const uint32_t add = 0x10000u - p.drop_thresh;
  const uint32_t sh_mine = hi ? 4u : 0u, sh_got = 4u - sh_mine;   // my calls are t = 2 hi, 2 hi + 1: keys 4 hi .. 4 hi + 3 (+ 8)
  uint32_t bits = 0;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    const unsigned long long idx = ((blk_idx + (unsigned)blk) << 5) + (unsigned)((i & 7) * 4 + (hi ? 2 : 0));
    const uint4 w0 = philox4x32(p.seed, idx, p.stream), w1 = philox4x32(p.seed, idx + 1, p.stream);
    // my rows' words of my two calls, and the partner's rows' words of the same calls
    const uint32_t mx0 = hi ? w0.z : w0.x, my0 = hi ? w0.w : w0.y, mx1 = hi ? w1.z : w1.x, my1 = hi ? w1.w : w1.y;
    const uint32_t sx0 = hi ? w0.x : w0.z, sy0 = hi ? w0.y : w0.w, sx1 = hi ? w1.x : w1.z, sy1 = hi ? w1.y : w1.w;
    const uint32_t gx0 = __shfl_xor_sync(0xffffffffu, sx0, 8), gy0 = __shfl_xor_sync(0xffffffffu, sy0, 8);
    const uint32_t gx1 = __shfl_xor_sync(0xffffffffu, sx1, 8), gy1 = __shfl_xor_sync(0xffffffffu, sy1, 8);
    // X words cover keys (2t, 2t+1), Y words keys (2t+8, 2t+9): two calls = 4 + 4 keys, placed by one shift per source
    const uint32_t mine = keep2(mx0, add) | (keep2(mx1, add) << 2) | (keep2(my0, add) << 8) | (keep2(my1, add) << 10);
    const uint32_t got = keep2(gx0, add) | (keep2(gx1, add) << 2) | (keep2(gy0, add) << 8) | (keep2(gy1, add) << 10);
    bits |= ((mine << sh_mine) | (got << sh_got)) << (16 * blk);
  }
  return bits;
}

__global__ void __launch_bounds__(kThreadsF, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const __grid_constant__ ParamsF p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  BarsF* bars = reinterpret_cast<BarsF*>(smem + oBarF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // heavy (late) query blocks of a causal problem first
  const int q0 = (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * (kTiles * BQ), h = blockIdx.y, b = blockIdx.z;
  const unsigned long long bh = (unsigned long long)b * p.H + h;
  const int klen = p.key_len ? min(p.key_len[b], p.Tk) : p.Tk;
  // key tiles of query tile w: the keys its LAST row sees (causal), none for a tile that starts past the last query
  int n_w[kTiles];
#pragma unroll
  for (int w = 0; w < kTiles; ++w) {
    const int k_end = p.causal ? min(klen, q0 + BQ * (w + 1)) : klen;
    n_w[w] = (q0 + BQ * w < p.Tq && k_end > 0) ? (k_end + BK - 1) / BK : 0;
  }
  const int n_max = max(n_w[0], n_w[1]);

  if (tid == 0) {
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], 1);
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 8);
      mbar_init(&bars->o_done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    if (n_max > 0) {
      // Q and the first K / V tiles are requested before the CTA-wide sync: their latency overlaps the tensor-memory allocation
      mbar_expect_tx(&bars->q_full, 65536u);
      mbar_expect_tx(&bars->k_full[0], 32768u);
      tma_2d(sbase + oQf, &tmQ, h * DH, b * p.Tq + q0, &bars->q_full);
      tma_2d(sbase + oQf + 16384, &tmQ, h * DH + 64, b * p.Tq + q0, &bars->q_full);
      tma_2d(sbase + oKf, &tmK, h * DH, b * p.Tk, &bars->k_full[0]);
      tma_2d(sbase + oKf + 16384, &tmK, h * DH + 64, b * p.Tk, &bars->k_full[0]);
      tma_2d(sbase + oQf + 32768, &tmQ, h * DH, b * p.Tq + q0 + BQ, &bars->q_full);
      tma_2d(sbase + oQf + 32768 + 16384, &tmQ, h * DH + 64, b * p.Tq + q0 + BQ, &bars->q_full);
      mbar_expect_tx(&bars->v_full[0], 32768u);
      tma_2d(sbase + oVf, &tmV, h * DH, b * p.Tk, &bars->v_full[0]);
      tma_2d(sbase + oVf + 16384, &tmV, h * DH + 64, b * p.Tk, &bars->v_full[0]);
    }
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = bars->tmem_base;
  // the softmax warpgroups keep a row's 128 logits in registers: take the helper warpgroup's share of the register file

  if (warp >= kTmaWarp) {
    // helper warpgroup (TMA, MMA issuer, two idle warps): ONE setmaxnreg for the whole warpgroup, then the roles
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
  }
  if (warp == kTmaWarp) {
    // ================= producer: both Q tiles once, then K / V tiles through two-stage rings =================
    if (lane == 0 && n_max > 0) {
      // tile 0 was requested in the prologue
      for (int j = 1; j < n_max; ++j) {
        const int s = j & 1;
        const uint32_t par = ((j >> 1) & 1u) ^ 1u;
        wait_bar(&bars->k_empty[s], par);
        mbar_expect_tx(&bars->k_full[s], 32768u);
        tma_2d(sbase + oKf + s * 32768, &tmK, h * DH, b * p.Tk + j * BK, &bars->k_full[s]);
        tma_2d(sbase + oKf + s * 32768 + 16384, &tmK, h * DH + 64, b * p.Tk + j * BK, &bars->k_full[s]);
        wait_bar(&bars->v_empty[s], par);
        mbar_expect_tx(&bars->v_full[s], 32768u);
        tma_2d(sbase + oVf + s * 32768, &tmV, h * DH, b * p.Tk + j * BK, &bars->v_full[s]);
        tma_2d(sbase + oVf + s * 32768 + 16384, &tmV, h * DH + 64, b * p.Tk + j * BK, &bars->v_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: PV_w(j) as soon as P_w(j) is published, then at once S_w(j+1): a tile's next logits are
    // computed while the OTHER tile's softmax warps work =================
    if (lane == 0 && n_max > 0) {
      constexpr uint32_t idS = idesc(128, BK, 0, 0), idO = idesc(128, DH, 0, 1);
      const uint64_t dQ0 = make_desc(sbase + oQf, 16, 1024);     // K-major A: Q tile 0 (tile 1: + 32768)
      const uint64_t dK0 = make_desc(sbase + oKf, 16, 1024);     // K-major B: K stage 0 (stage 1: + 32768)
      const uint64_t dV0 = make_desc(sbase + oVf, 16384, 1024);  // MN-major B: V stage 0
      auto issue_s = [&](int w, int j) {
        const uint64_t qd = dQ0 + (uint64_t)((w * 32768) >> 4), kd = dK0 + (uint64_t)(((j & 1) * 32768) >> 4);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint64_t off = (uint64_t)(((ks >> 2) * 16384 + (ks & 3) * 32) >> 4);
          umma_bf16(tmem + cSf + w * BK, qd + off, kd + off, idS, ks > 0);
        }
        umma_commit(&bars->s_full[w]);
      };
      wait_bar(&bars->q_full, 0);
      wait_bar(&bars->k_full[0], 0);
      fence_after();
#pragma unroll
      for (int w = 0; w < kTiles; ++w)
        if (n_w[w] > 0) issue_s(w, 0);
      umma_commit(&bars->k_empty[0]);
      for (int j = 0; j < n_max; ++j) {
        const int s = j & 1;
        wait_bar(&bars->v_full[s], (j >> 1) & 1u);
        if (j + 1 < n_max) wait_bar(&bars->k_full[s ^ 1], ((j + 1) >> 1) & 1u);
        const uint64_t vd = dV0 + (uint64_t)((s * 32768) >> 4);
#pragma unroll
        for (int w = 0; w < kTiles; ++w) {
          if (j < n_w[w]) {
            wait_bar(&bars->p_full[w], j & 1u);
            fence_after();
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks)   // O_w += P_w V_j
              umma_bf16_ta(tmem + cOf + w * DH, tmem + cSf + w * BK + (ks >> 2) * (BK / 2) + (ks & 3) * 8, vd + (uint64_t)((ks * 2048) >> 4), idO,
                           (j > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&bars->o_done[w]);
            if (j + 1 < n_w[w]) issue_s(w, j + 1);
          }
        }
        umma_commit(&bars->v_empty[s]);
        if (j + 1 < n_max) umma_commit(&bars->k_empty[s ^ 1]);
      }
    }
    __syncwarp();
  } else if (warp < kTmaWarp) {
    // ================= softmax warps: query tile w = warp / 8; TWO threads per query row (TMEM lane quadrant = warp % 4): the
    // thread of half 0 owns keys 0..63 of the tile (and head columns 0..47 of O), its partner in warp + 4 keys 64..127 (columns
    // 48..95).  Sixteen warps = four per scheduler hide each other's dependent chains (Philox, exp2, tensor-memory round
    // trips); the halves of a row exchange their maxima through shared memory at one named barrier per tile =================
    // register pool of the CTA = 96 (launch bound of 640 threads) x 640 = 61440 = 512 x 104 + 128 x 64: an increase beyond the
    // pool would wait forever
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    const int w = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
    const int r = quad * 32 + lane, i = q0 + BQ * w + r;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t sS = lane_addr + cSf + w * BK + half * (BK / 2), sP = sS;   // P goes over this thread's OWN consumed S columns
    const uint32_t sO = lane_addr + cOf + w * DH + half * (DH / 2);
    float* xch = reinterpret_cast<float*>(smem + oXchF) + w * 512;   // [2 buffers][2 halves][128 rows]
    const int n_it = n_w[w];
    const int n_iblk = (p.Tq + 15) >> 4, n_jblk = (p.Tk + 15) >> 4;
    const unsigned long long blk_row = (bh * (unsigned long long)n_iblk + (unsigned)(i >> 4)) * (unsigned long long)n_jblk;
    const bool drop = p.drop_thresh != 0u;
    float m = -CUDART_INF_F, l = 0.f;   // l: this thread's half of the row sum
    for (int j = 0; j < n_it; ++j) {
      const int kb = j * BK + half * (BK / 2);   // first key of this thread's 64
      wait_bar(&bars->s_full[w], j & 1u);
      fence_after();
      // interior tiles: every key of the tile exists and is visible to every row of the query tile
      const bool open = j * BK + BK <= klen && (!p.causal || j * BK + BK - 1 <= q0 + BQ * w);
      // ---- pass 1: maximum of the half row (rolled loops over the two 32-key chunks keep the code small: the unrolled
      //      version was ~80 KB of instructions and the warps starved on instruction fetch), exchanged with the partner
      float mx = -CUDART_INF_F;
      // keys this WARP's 32 rows can see at all (causal diagonal tiles / the last key tile): chunks past it are skipped
      const int warp_lim = min(klen, p.causal ? q0 + BQ * w + quad * 32 + 32 : klen);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (kb + 32 * c >= warp_lim) break;   // warp-uniform
        uint32_t sv[32];
        tmem_ld32(sS + 32 * c, sv);
        tmem_wait_ld();
        float m4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        if (open) {
#pragma unroll
          for (int e = 0; e < 32; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(sv[e]));
        } else {
          const int lim = min(klen, p.causal ? i + 1 : klen) - (kb + 32 * c);   // keys kb + 32 c + e with e < lim are visible
#pragma unroll
          for (int e = 0; e < 32; ++e) m4[e & 3] = fmaxf(m4[e & 3], e < lim ? __uint_as_float(sv[e]) : -CUDART_INF_F);
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      float* xb = xch + (j & 1) * 256;
      xb[half * 128 + r] = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
      mx = fmaxf(mx, xb[(half ^ 1) * 128 + r]) * p.scale_log2;   // scale > 0: max(s) scale = max(s scale); -inf stays -inf
      // The reference point of the exponentials only moves when the row maximum grew by more than 2^8 (or on the first
      // tile): p <= 256 is harmless in bf16 / fp32, l and lse = m + log2(l) stay exact, and O_w is rescaled rarely.
      const bool move = mx > m + 8.f || (m == -CUDART_INF_F && mx > -CUDART_INF_F);
      const float mn = move ? mx : m;
      const float mu = mn == -CUDART_INF_F ? 0.f : mn;
      const float corr = move ? ex2(m - mu) : 1.f;   // exp2(-inf) = 0 on the first tile
      m = mn;
      l *= corr;
      // ---- rescale this thread's 48 columns of O_w when a row's reference moved (warp-uniform decision: the tensor-memory
      //      accesses are collective; both halves of a row see the same maxima and decide alike)
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        wait_bar(&bars->o_done[w], (j - 1) & 1u);   // PV_w(j-1) has been added
        fence_after();
        uint32_t ov[32], ow[16];
        tmem_ld32(sO, ov);
        tmem_ld16(sO + 32, ow);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * corr);
#pragma unroll
        for (int e = 0; e < 16; ++e) ow[e] = __float_as_uint(__uint_as_float(ow[e]) * corr);
        tmem_st32(sO, ov);
        tmem_st16(sO + 32, ow);
      }
      // ---- pass 2: p = exp2(s scale - reference), row sum, dropout, bf16 pairs back into tensor memory over the S columns
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (kb + 32 * c >= warp_lim) {   // nothing visible to any row of the warp: P = 0 (no Philox, no exp2; the cached keep
          uint32_t z[16];               // bits of masked weights are never read)
#pragma unroll
          for (int e = 0; e < 16; ++e) z[e] = 0u;
          tmem_st16(sP + 16 * c, z);
          continue;
        }
        uint32_t sv[32];
        tmem_ld32(sS + 32 * c, sv);
        uint32_t bits = 0xffffffffu;
        if (drop) {   // the Philox calls run while the logits travel
          bits = keep_bits32(p, blk_row + (unsigned)((kb >> 4) + 2 * c), i, lane);
          const int kw = (kb >> 5) + c;
          if (p.keep_mask != nullptr && i < p.Tq && kw < p.n_kw) p.keep_mask[(bh * p.n_kw + kw) * p.Tq + i] = bits;
        }
        tmem_wait_ld();
        if (!open) {
          const int lim = min(klen, p.causal ? i + 1 : klen) - (kb + 32 * c);
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e >= lim) sv[e] = 0xff800000u;   // -inf: exp2 gives 0
        }
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {   // 16 weights at a time: few live registers
          float pe[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pe[e] = ex2(fmaf(__uint_as_float(sv[16 * hc + e]), p.scale_log2, -mu));
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 16; ++e) s4[e & 3] += pe[e];
          l += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          if (drop) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (!((bits >> (16 * hc + e)) & 1u)) pe[e] = 0.f;
          }
          uint32_t pw[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) pw[e] = pack_bf16(pe[2 * e], pe[2 * e + 1]);
          tmem_st8(sP + 16 * c + 8 * hc, pw);   // columns 64 half + 16 c + 8 hc ..: k-step 4 half + 2 c + hc of P_w V_j
        }
      }
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[w]);
    }
    // ---- epilogue: O_w / l (and the 1 / (1 - p) of the dropout) -> bf16 context, log2-domain log-sum-exp
    {
      float* xb = xch + (n_it & 1) * 256;   // the buffer the last tile did not use
      xb[half * 128 + r] = l;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
      l += xb[(half ^ 1) * 128 + r];
    }
    if (n_it > 0) {
      wait_bar(&bars->o_done[w], (n_it - 1) & 1u);
      fence_after();
    }
    const float inv = l > 0.f ? p.drop_scale / l : 0.f;
    __nv_bfloat16* op = p.o + ((long long)b * p.Tq + i) * p.ldo + h * DH + half * (DH / 2);
    uint32_t v[48];
    if (n_it > 0) {
      tmem_ld32(sO, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld16(sO + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int e = 0; e < 48; ++e) v[e] = 0u;
    }
    if (i < p.Tq) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
        o.y = pack_bf16(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
        o.z = pack_bf16(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
        o.w = pack_bf16(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
        *reinterpret_cast<uint4*>(op + 8 * q) = o;
      }
      if (half == 0 && p.lse != nullptr) p.lse[bh * p.Tq + i] = m + log2f(l);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace tcf

// forward on tcgen05 (head_dim 96)
static int launch_fwd_tc(const Args& a, cudaStream_t s) {
  constexpr int DH = tc::DH;
  CUtensorMap mq, mk, mv;
  const long long width = (long long)a.H * DH;
  int rc;
  if ((rc = make_tma_map_bf16(&mq, a.q, width, (long long)a.B * a.Tq, a.ldq, tcf::BQ))) return rc;
  if ((rc = make_tma_map_bf16(&mk, a.k, width, (long long)a.B * a.Tk, a.ldk, tcf::BK))) return rc;
  if ((rc = make_tma_map_bf16(&mv, a.v, width, (long long)a.B * a.Tk, a.ldv, tcf::BK))) return rc;
  tcf::ParamsF p;
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk; p.causal = a.causal; p.n_kw = a.n_kw;
  p.scale_log2 = a.scale_log2; p.drop_scale = a.drop_scale; p.drop_thresh = a.drop_thresh; p.stream = a.stream; p.seed = a.seed;
  p.key_len = a.key_len; p.lse = a.lse; p.keep_mask = a.keep_mask; p.o = a.o; p.ldo = a.ldo;
  static std::atomic<unsigned long long> attr{0ull};   // per device: cudaFuncSetAttribute applies to the current device only
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long dev_bit = 1ull << (dev & 63);
  if (!(attr.load(std::memory_order_acquire) & dev_bit)) {
    // setmaxnreg redistributes the registers the CTA was launched with: 16 softmax warps x 104 + 4 helper warps x 64
    cudaFuncAttributes fa;
    TTS_CHECK_CUDA(cudaFuncGetAttributes(&fa, tcf::attn_fwd_tc_kernel));
    TTS_REQUIRE((long long)fa.numRegs * tcf::kThreadsF >= 512LL * 104 + 128LL * 64,
                "attn_fwd_tc: built with %d registers per thread, the kernel's setmaxnreg budget needs 96", fa.numRegs);
    TTS_CHECK_CUDA(cudaFuncSetAttribute(tcf::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcf::kSmemF));
    attr.fetch_or(dev_bit, std::memory_order_release);
  }
  tcf::attn_fwd_tc_kernel<<<dim3(ceil_div(a.Tq, tcf::kTiles * tcf::BQ), a.H, a.B), tcf::kThreadsF, tcf::kSmemF, s>>>(mq, mk, mv, p);
  TTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace attn
}  // namespace tts
