// tcgen05 (5th-generation tensor core) GEMM for the dense, once-per-utterance side of the mel path:
// C[M,N] = epi(A[M,K] * W[N,K]^T) with fp32-class accuracy through the 3-term TF32 split
//     A W^T ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T,   hi = the upper 19 bits the tensor core reads, lo = x - hi.
// One CTA computes a 128 x 128 tile:
//   warps 0-7  loaders + epilogue: 16-byte chunks of A and W travel global -> registers (requested one stage ahead)
//              -> shared memory in the canonical K-major, no-swizzle core-matrix layout (8 rows x 16 bytes = 128
//              contiguous bytes per core matrix), once as `hi` (rounded to TF32) and once as `lo` = x - hi;
//              cross-proxy fence, arrive on the stage's `full` mbarrier;
//   warp 8     one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8): 4 k-steps x 3 terms per stage,
//              accumulating in tensor memory (128 lanes x 128 fp32 columns); tcgen05.commit frees the stage;
//   epilogue   tcgen05.ld 32x32b.x8 (thread = one output row, 8 columns at a time), the TtsGemmEpilogue of
//              tts_b200.h applied in registers, row-wise stores.
// Both operands are K-contiguous ("NT"), so no transposes are needed.  Replaces the FFMA2 kernel of dense.cu for
// K % 32 == 0 (every Linear / Conv-as-GEMM of the encoder, the cross-K/V precompute and the Postnet except the
// 80-channel input layers).  Reference call sites: transformer/attention.py:43-47,63-68,119; modules.py:11-19;
// tacotron.py:78,85.
#include <atomic>

#include "common.cuh"

namespace tts {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32, kStages = 3;
constexpr int kLoaders = 256;                  // warps 0-7
constexpr int kThreads = kLoaders + 32;        // + MMA warp
constexpr int kTileBytes = BM * BK * 4;        // 16 KB: one operand tile (hi or lo) of one stage
constexpr int kStageBytes = 4 * kTileBytes;    // A_hi, A_lo, W_hi, W_lo
constexpr int kKBlockBytes = BM * 16;          // 2048: all rows of one 16-byte k-block (LBO)
constexpr int kAcc = 4;                        // accumulators in tensor memory, k-blocks dealt round-robin: the tensor
                                               // core adds into fp32 with truncation, so the error grows linearly with
                                               // the number of accumulations (measured 1.1e-4 at K = 2560 with one
                                               // accumulator); four partial sums added in registers (RN) cut it 4x
constexpr int kTmemCols = kAcc * BN;           // 512 = all of tensor memory (one CTA per SM)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity) {   // false on timeout (never hangs the GPU)
  long long t0 = 0, spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if ((++spins & 255) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > (2LL << 30)) return false;
    }
  }
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
// Canonical layout in 16-byte units ((8,n),2):((1,SBO),LBO): 8 rows of a core matrix are 16 bytes apart, 8-row
// groups SBO apart, the two 16-byte k-blocks of one K=8 instruction LBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((kKBlockBytes >> 4) & 0x3FFF) << 16;   // LBO: next 16-byte k-block
  d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;            // SBO: next group of 8 rows
  d |= 1ull << 46;                                       // descriptor version (sm_100)
  return d;
}
// kind::tf32 instruction descriptor (InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Shared {
  uint64_t full[kStages], empty[kStages], accum;
  uint32_t tmem_base;
  int err;
};

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const float* __restrict__ A, int lda,
                                                              const float* __restrict__ W, int ldw,
                                                              float* __restrict__ C, int ldc, int M, int N, int K,
                                                              TtsGemmEpilogue epi) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Shared* sh = reinterpret_cast<Shared*>(smem + kStages * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int n_kb = K / BK;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full[s], kLoaders);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->accum, 1);
    sh->err = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLoaders / 32) {   // the MMA warp owns the tensor-memory allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = sh->tmem_base;

  if (warp < kLoaders / 32) {
    // ================= loaders: 4 chunks of A and 4 of W per thread and stage =================
    // chunk c = i * 256 + tid -> k-block = 2 * (c >> 8) + (c & 1), row = (c >> 1) & 127: two neighbouring threads
    // read one 32-byte sector of a row, and their shared-memory stores conflict at most 2-way
    float4 ra[4], rw[4];
    auto load = [&](int kb) {
      const int k0 = kb * BK;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * kLoaders + tid, kq = 2 * (c >> 8) + (c & 1), row = (c >> 1) & 127;
        const int gm = m0 + row, gn = n0 + row;
        ra[i] = gm < M ? __ldg(reinterpret_cast<const float4*>(A + (size_t)gm * lda + k0 + kq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        rw[i] = gn < N ? __ldg(reinterpret_cast<const float4*>(W + (size_t)gn * ldw + k0 + kq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto split = [](float x, float& hi, float& lo) {   // hi: round to nearest TF32 (low 13 bits zero), lo exact
      hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
      lo = x - hi;
    };
    auto store = [&](int s, const float4 (&va)[4], const float4 (&vw)[4]) {
      uint8_t* st = smem + s * kStageBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * kLoaders + tid, kq = 2 * (c >> 8) + (c & 1), row = (c >> 1) & 127;
        const uint32_t off = (uint32_t)kq * kKBlockBytes + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
        float4 h, l;
        split(va[i].x, h.x, l.x); split(va[i].y, h.y, l.y); split(va[i].z, h.z, l.z); split(va[i].w, h.w, l.w);
        *reinterpret_cast<float4*>(st + off) = h;
        *reinterpret_cast<float4*>(st + kTileBytes + off) = l;
        split(vw[i].x, h.x, l.x); split(vw[i].y, h.y, l.y); split(vw[i].z, h.z, l.z); split(vw[i].w, h.w, l.w);
        *reinterpret_cast<float4*>(st + 2 * kTileBytes + off) = h;
        *reinterpret_cast<float4*>(st + 3 * kTileBytes + off) = l;
      }
    };
    bool ok = true;
    load(0);
    for (int kb = 0; kb < n_kb && ok; ++kb) {
      float4 ca[4], cw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ca[i] = ra[i];
        cw[i] = rw[i];
      }
      if (kb + 1 < n_kb) load(kb + 1);   // in flight while this stage is converted and stored
      const int s = kb % kStages;
      if (kb >= kStages) ok = mbar_wait(&sh->empty[s], ((kb / kStages) - 1) & 1u);
      store(s, ca, cw);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor-core reads
      mbar_arrive(&sh->full[s]);
    }
    if (!ok) atomicExch(&sh->err, 1);

    // ================= epilogue: thread = output row, 8 columns per tensor-memory load =================
    ok = mbar_wait(&sh->accum, 0u) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = (warp & 3) * 32 + lane, m = m0 + row;   // a warp reads the TMEM lanes 32 * (warp % 4) ...
    const int rpb = epi.rows_per_batch > 0 ? epi.rows_per_batch : M;
    const int valid = epi.valid_rows > 0 ? epi.valid_rows : rpb;
    const int orpb = epi.out_rows_per_batch > 0 ? epi.out_rows_per_batch : rpb;
    const int b = m / rpb, r = m - b * rpb;
    const bool store_row = m < M && r < valid;   // on a barrier timeout (!ok) the tile is stored as NaN: loud, never stale
    const bool dead = store_row && epi.row_len != nullptr && r >= epi.row_len[b];
    const size_t orow = (size_t)b * orpb + r + epi.out_row_offset;
    for (int c0 = (warp >> 2) * (BN / 2); c0 < (warp >> 2) * (BN / 2) + BN / 2; c0 += 8) {   // ... and half of the columns
      float v[8];
      const int n_acc = n_kb < kAcc ? n_kb : kAcc;   // accumulators that were written
#pragma unroll
      for (int a = 0; a < kAcc; ++a) {
        if (a < n_acc) {   // uniform
          uint32_t u[8];
          const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * BN + c0);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                       : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = a == 0 ? __uint_as_float(u[j]) : v[j] + __uint_as_float(u[j]);
        }
      }
      if (!store_row) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + c0 + j;
        if (n >= N) continue;
        float x = v[j] * epi.alpha;
        if (epi.scale) x = x * epi.scale[n] + epi.shift[n];
        if (epi.bias) x += epi.bias[n];
        if (epi.act == 1) x = fmaxf(x, 0.f);
        else if (epi.act == 2) x = tanhf(x);
        if (epi.residual) x += epi.residual[orow * epi.ldr + n];
        if (dead) x = 0.f;
        if (!ok) x = __int_as_float(0x7fc00000);
        if (epi.head_dim > 0) {
          const int dm = epi.n_heads * epi.head_dim;
          const int w = n / dm, hn = n - w * dm, h = hn / epi.head_dim, d = hn - h * epi.head_dim;
          float* dst = w ? epi.out_v : C;
          dst[(((size_t)b * epi.n_heads + h) * epi.head_rows + r) * epi.head_dim + d] = x;
        } else {
          C[orow * ldc + n] = x;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ================= MMA issuer (one elected lane) =================
    if (lane == 0) {
      bool ok = true;
      for (int kb = 0; kb < n_kb && ok; ++kb) {
        const int s = kb % kStages;
        ok = mbar_wait(&sh->full[s], (kb / kStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(smem + s * kStageBytes);
        const uint32_t a_hi = base, a_lo = base + kTileBytes, w_hi = base + 2 * kTileBytes, w_lo = base + 3 * kTileBytes;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t ko = (uint32_t)ks * 2u * kKBlockBytes;   // two 16-byte k-blocks per K=8 instruction
          const uint32_t acc = tmem_d + (uint32_t)((kb % kAcc) * BN);
          umma_tf32(acc, make_desc(a_lo + ko), make_desc(w_hi + ko), (kb >= kAcc || ks > 0) ? 1u : 0u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(w_lo + ko), 1u);
          umma_tf32(acc, make_desc(a_hi + ko), make_desc(w_hi + ko), 1u);
        }
        umma_commit(&sh->empty[s]);   // the stage may be refilled once these MMAs have read it
      }
      umma_commit(&sh->accum);        // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kLoaders / 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)kTmemCols) : "memory");
}

static size_t smem_bytes() { return (size_t)kStages * kStageBytes + sizeof(Shared) + 64; }

}  // namespace tc

bool gemm_tc_supported(int M, int N, int K, int lda, int ldw) {
  return K % tc::BK == 0 && K >= (tc::kStages - 1) * tc::BK && lda % 4 == 0 && ldw % 4 == 0 && M >= 1 && N >= 1;
}

int launch_gemm_tc(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                   const TtsGemmEpilogue& epi, cudaStream_t s) {
  static std::atomic<unsigned long long> configured{0ull};   // bit per device
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes()));
    configured.fetch_or(bit, std::memory_order_release);
  }
  dim3 grid(ceil_div(N, tc::BN), ceil_div(M, tc::BM));
  tc::gemm_tc_kernel<<<grid, tc::kThreads, tc::smem_bytes(), s>>>(A, lda, W, ldw, C, ldc, M, N, K, epi);
  TTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace tts
