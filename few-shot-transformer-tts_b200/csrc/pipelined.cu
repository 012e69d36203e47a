// Pipelined persistent decode kernel (impl 4, the default): ONE cooperative launch runs `n_steps` whole decode
// steps with one CTA per SM, like megakernel.cu, but the step is software-pipelined over ROW GROUPS of the batch
// so that the latency of the grid-wide dependency between phases (barrier + activation broadcast, ~2 us measured,
// profiles/r1_microbench.txt) hides behind the work of the other group(s):
//
//   * the batch is cut into groups of <= 16 rows (one m16 MMA tile).  Every (phase, group) pair has its own
//     grid-barrier counter: group g of phase p only waits for group g of phase p-1, so while the last CTAs
//     finish (p, g) everybody else already works on (p, g+1) or (p+1, g-1).
//   * warp specialisation (11 warps): 8 consumer warps compute; a LOADER warp polls the barrier counters, stages the
//     next activation tile with a 1-D TMA bulk copy (one mbarrier, one slot: the consumers pull the tile into
//     registers with ldmatrix and release the slot at once) and streams the NEXT phase's packed weight slice into the
//     half of the weight region the current phase does not use; a SIGNALER warp publishes finished group-phases
//     with the gpu-scope release (a memory barrier) off the consumers' path; a FEEDER warp (one lane per ring slot)
//     keeps the K/V ring of the attention phases full.
//   * products run on the legacy tensor-core path: mma.sync m16n8k8 TF32 with the 3-term split
//     (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate) which keeps fp32-class accuracy (parity 1e-3 on mel
//     frames needs it; a single TF32 pass does not).  tcgen05 needs M >= 64 per CTA; a CTA owns 5-21 weight
//     rows here, so the 128-row tensor-memory path would waste > 80 % of every instruction.
//   * LayerNorm is applied in the epilogue: y = rstd * (W_ln x) - rstd * mean * rowsum(W_ln) + c_ln, the row
//     statistics are computed by the consumers from the fragments they already hold (no pass over the tile).
//   * FFN-out (K = 3072) is split 4-way along K across CTAs (each CTA then stages the same 16 x 768 tile shape as
//     every other phase instead of 16 x 3072) and a short reduce phase adds the four partials to the residual.
//   * attention phases: one (sample, head) stream per CTA and group through a CTA-wide ring of 20 x 6 KB tiles
//     (120 KB in flight per SM, partly filled while the preceding GEMM still runs), online softmax in the log2
//     domain, static tile-to-warp assignment (bit-reproducible).
// Measured (profiles/r1_*): 414 us per decode step at B=32, S=258 averaged over 1000 frames = 0.40 of the HBM
// roofline; DESIGN.md section 4.1 has the breakdown and what was tried.
//
// Reference semantics: transformer/tacotron.py:107-116, transformer/modules.py:108-145,
// transformer/attention.py:53-122, synthesize.py:35-45.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "philox.cuh"

namespace tts {
namespace pipe {

constexpr int kCWarps = 8;                 // consumer warps
constexpr int kConsumers = kCWarps * 32;   // 256
constexpr int kThreads = kConsumers + 96;  // + loader warp + signaler warp + K/V feeder warp
constexpr int kGroupRows = 16;             // batch rows per group (one m16 tile)
constexpr int kKC = 768;                   // widest K slice of an activation tile / weight row in shared memory
constexpr int kPad = 16;                   // floats of padding per weight row in shared memory (bank spread of LDS.128)
constexpr int kXPad = 4;                   // floats of padding per activation row: rows 4 banks apart -> ldmatrix conflict-free
constexpr int kLdMax = kKC + kPad;
constexpr int kMaxRows = 24;               // weight rows per CTA and phase (3 n-tiles of 8)
constexpr int kChunks = kKC / 16 / kCWarps;             // 16-float K chunks per warp (6)
constexpr int kXsFloats = kGroupRows * kLdMax;          // activation slot
constexpr int kWFloats = 2 * kMaxRows * kLdMax;         // weight region: two phases in flight
constexpr int kRedFloats = kCWarps * kGroupRows * kMaxRows;
constexpr int kTK = 8;                     // keys per K/V ring tile (16-key tiles: register spills, no gain; 4-key tiles: slower)
constexpr int kSlots = 24;                 // most slots of the CTA-wide K/V ring (K tile + V tile each; one feeder lane per slot)
constexpr int kMaxBatch = 1024;
constexpr int kMaxGroups = 64;
constexpr int kMaxSplit = 32;
constexpr int kDescRing = 4;
constexpr long long kTimeoutCycles = 3LL << 30;   // ~1.6 s of SM clock: no wait may hang the GPU
constexpr int kProfPhases = 160;
constexpr int kProfStride = 16;

enum Kind { kGemm = 0, kAttn = 1, kReduce = 2, kCombine = 3 };
enum Mode { kPlain = 0, kQkv = 1, kPrenetOut = 2, kFinal = 3, kPartial = 4, kHid = 5 };

struct Args {
  TtsDecoderWeights w;
  TtsDecodeState st;
  float *x, *q, *ctx, *hid, *p0, *p1, *part, *fpart;
  unsigned* bar;        // [kMaxGroups] counters, 32 words apart
  int* err;
  long long* prof;
  int n_steps, update_state;
  int group_rows, n_groups, n_split, ksplit;
  int prefetch;         // L2 prefetch of upcoming K/V streams (TTS_PREFETCH=1 enables it)
  // decoder.train() at synthesis time (eval.py:116-117): Philox dropout, p = decoder_dropout_rate after the two prenet
  // ReLUs (tacotron.py:58,62) and p = transformer_dropout_rate after the PE add, on every attention weight, every
  // residual-branch output and the FFN hidden (modules.py:120,132,138,141,18; attention.py:89).  0 = off.
  uint32_t thr_d, thr_t; float sc_d, sc_t; unsigned long long seed;
  int ring_lo, n_slots, n_hi;   // K/V ring.  Slots [0, n_hi) sit above the weight tiles of the GEMM that precedes an
                                // attention phase (from float ring_lo on, below the next GEMM's tiles): they may be
                                // filled while that GEMM still runs.  Slots [n_hi, n_slots) reuse the space of those
                                // weight tiles and are filled once the consumers have finished the GEMM.
};

struct Desc {
  int kind;
  // ---- GEMM: Y[b][n] = epilogue(sum_k X[b][k] W[n][k]) for the rows of one group
  const float* X; long long ldx; int K, N, ksplit;
  const float* W;       // packed rows [ksplit][N][K/ksplit + 16]: weights | constant | LayerNorm row sum (tts_b200.h pk_*)
  int ln, relu, mode, hi, zero_x;
  float inv_k;          // 1 / K
  float* Y; long long ldy; const float* R; long long ldr; float out_scale;
  float* kcache; float* vcache;
  uint32_t drop_thr; float drop_sc; uint32_t drop_site;   // dropout on the stored value (before the residual add); 0 = none
  // ---- attention over a K/V stream
  const float* kc; const float* vc; int rows_alloc, n_keys; const int32_t* key_len;
  float* align; long long align_bh_stride; int align_row_len;
  // ---- reduce: Y[b][n] += sum_s part[s][b][n]
  const float* part; int n_parts;
  // ---- this CTA's share of the phase (filled in by the loader warp: constant for the phase)
  int s_n_lo, s_n_hi, s_k_lo, s_kc, s_ks;
};

struct Smem {
  float* xs;       // activation slot [16][ld]
  float* wreg;     // weight region; K/V rings + logits during attention phases
  float* red;      // [8][16][24] cross-warp reduction / attention warp records
  float* spart;    // [8][16][2] LayerNorm partial (sum, sum of squares) about the row's first element
  float* sshift;   // [16]
  float* qbuf;     // [2][96] query of the CTA's first attention unit of a group-phase, prefetched by the feeder warp
  int* len;        // [B]
  int* fin;        // [B]
  int* klen;       // [B] input_lengths (key mask of the cross attention), cached once per launch
  uint64_t* x_full;   // 1: producer -> consumers, tile landed / group may start
  uint64_t* x_empty;  // 1: consumers -> producer, slot free (8 arrivals)
  uint64_t* w_full;   // 2: weight slice landed (low / high placement)
  uint64_t* pdone;    // phases the consumers have finished, as a plain counter (an mbarrier's parity would alias:
                      // with a single row group the consumers can complete two phases before the loader looks)
  uint64_t* qfull;    // 2: prefetched query landed
  uint64_t* rfull;    // [kSlots] ring slot filled (feeder warp -> consumers)
  unsigned* drained;  // [kSlots] how many tiles have been consumed out of each slot.  A plain counter, not an mbarrier:
                      // a consumer may reach the slot's use n+2 while use n is still being read, and a parity
                      // wait cannot tell "two phases behind" from "done" (this aliasing was hit on hardware)
  unsigned* kv_go;    // attention group-phases the loader has released to the feeder
  unsigned* sig;      // group-phases the consumers have finished (polled by the signaler warp)
  Desc* desc;         // [kDescRing]
};

__host__ __device__ inline size_t smem_floats_fixed() {
  return (size_t)kXsFloats + kWFloats + kRedFloats + kCWarps * kGroupRows * 2 + kGroupRows + 2 * 96;
}
static size_t smem_bytes(int B) {
  return smem_floats_fixed() * sizeof(float) + (size_t)3 * B * sizeof(int) + (10 + 2 * kSlots) * sizeof(uint64_t) +
         kDescRing * sizeof(Desc) + 64;
}

__device__ __forceinline__ Smem make_smem(const Args& a, float* base) {
  Smem sm;
  sm.xs = base;
  sm.wreg = sm.xs + kXsFloats;
  sm.red = sm.wreg + kWFloats;
  sm.spart = sm.red + kRedFloats;
  sm.sshift = sm.spart + kCWarps * kGroupRows * 2;
  sm.qbuf = sm.sshift + kGroupRows;
  sm.len = reinterpret_cast<int*>(sm.qbuf + 2 * 96);
  sm.fin = sm.len + a.st.batch;
  sm.klen = sm.fin + a.st.batch;
  uintptr_t p = reinterpret_cast<uintptr_t>(sm.klen + a.st.batch);
  p = (p + 15) & ~(uintptr_t)15;
  sm.x_full = reinterpret_cast<uint64_t*>(p);
  sm.x_empty = sm.x_full + 1;
  sm.w_full = sm.x_full + 2;
  sm.pdone = sm.x_full + 4;
  sm.qfull = sm.x_full + 5;
  sm.rfull = sm.x_full + 7;
  sm.drained = reinterpret_cast<unsigned*>(sm.rfull + kSlots);
  uint64_t* after = sm.rfull + 2 * kSlots;
  sm.sig = reinterpret_cast<unsigned*>(after);
  sm.kv_go = sm.sig + 1;
  sm.desc = reinterpret_cast<Desc*>(after + 2);
  return sm;
}

// ---- primitives -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, int* err) {
  long long spins = 0, t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeoutCycles || *reinterpret_cast<volatile int*>(err) != 0) {
        atomicExch(err, 2);  // never hang the GPU
        break;
      }
    }
  }
}
__device__ __forceinline__ void bulk_g2s(float* dst, const float* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy global writes (observed through an acquire) -> later async-proxy (TMA) reads of global memory.
// The state-space-qualified form is a single FENCE.VIEW.ASYNC.G; the unqualified one adds a MEMBAR.ALL.GPU that
// waits for every outstanding store of the warp (8 % of all stall samples in profiles/r2_pipe_a).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
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

// TF32 split: the tensor core reads only the upper 19 bits of an fp32 operand (sign, 8-bit exponent, 10-bit
// mantissa), so hi is x itself (truncated by the hardware) and lo = x - trunc(x) is exact in fp32;
// lo*hi + hi*lo + hi*hi then recovers the fp32 product to ~2^-20 relative.  Two full-rate ALU ops per element
// (the rounding conversion cvt.rna.tf32.f32 of the textbook split runs on a quarter-rate pipe and dominated the
// product time).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// dropout of ONE value: element e of dropout site `site` (philox.cuh: four consecutive elements share a call)
__device__ __forceinline__ float drop1(unsigned long long seed, uint32_t thr, float sc, uint32_t site, unsigned long long e, float v) {
  const uint4 w = philox4x32(seed, e >> 2, site);
  const uint32_t k = (uint32_t)(e & 3ull);
  const uint32_t r = k == 0u ? w.x : (k == 1u ? w.y : (k == 2u ? w.z : w.w));
  return r >= thr ? v * sc : 0.f;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- per-group grid barrier ---------------------------------------------------------------------------
__device__ __forceinline__ void grid_arrive(const Args& a, int g) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar + 32 * g) : "memory");
}
__device__ __forceinline__ void grid_wait(const Args& a, int g, unsigned target) {
  const unsigned* ctr = a.bar + 32 * g;
  long long spins = 0, t0 = 0;
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (static_cast<int>(v - target) >= 0) break;
    if ((++spins & 255) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeoutCycles || *reinterpret_cast<volatile int*>(a.err) != 0) {
        atomicExch(a.err, 1);
        break;
      }
    }
  }
}

// spin until a shared-memory counter written with st.release.cta reaches `target`
__device__ __forceinline__ void wait_count(const Args& a, const void* ctr, unsigned target) {
  long long spins = 0, t0 = 0;
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(ctr)) : "memory");
    if (static_cast<int>(v - target) >= 0) break;
    if ((++spins & 4095) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeoutCycles || *reinterpret_cast<volatile int*>(a.err) != 0) {
        atomicExch(a.err, 3);
        break;
      }
    }
  }
}

// ---- work split of a GEMM phase over the CTAs -----------------------------------------------------------
struct Slice {
  int n_lo, n_hi, k_lo, kc, ks;
};
__device__ __forceinline__ Slice compute_slice(const Desc& d, unsigned c, unsigned G) {   // N * G < 2^32
  Slice s;
  if (d.ksplit <= 1) {
    s.ks = 0;
    s.n_lo = (int)((c * (unsigned)d.N) / G);
    s.n_hi = (int)(((c + 1u) * (unsigned)d.N) / G);
    s.k_lo = 0;
    s.kc = d.K;
  } else {  // CTA c works on K slice c % ksplit; the CTAs of one slice share the N rows
    const unsigned ksp = (unsigned)d.ksplit;
    s.ks = (int)(c % ksp);
    const unsigned i = c / ksp, nc = (G - (unsigned)s.ks + ksp - 1u) / ksp;
    s.n_lo = (int)((i * (unsigned)d.N) / nc);
    s.n_hi = (int)(((i + 1u) * (unsigned)d.N) / nc);
    s.kc = d.K / d.ksplit;
    s.k_lo = s.ks * s.kc;
  }
  return s;
}
__device__ __forceinline__ Slice slice_of(const Desc& d, int, int) {
  Slice s;
  s.n_lo = d.s_n_lo; s.n_hi = d.s_n_hi; s.k_lo = d.s_k_lo; s.kc = d.s_kc; s.ks = d.s_ks;
  return s;
}
__device__ __forceinline__ float* weight_base(const Smem& sm, const Desc& d, const Slice& s) {
  const int ld = s.kc + kPad;
  const int tiles = (s.n_hi - s.n_lo + 7) >> 3;
  return d.hi ? sm.wreg + kWFloats - tiles * 8 * ld : sm.wreg;
}

// ---- producer side ---------------------------------------------------------------------------------------
// packed rows [n][kc + 16] (weights | constant | LayerNorm row sum | zeros): a CTA's slice is one contiguous run
__device__ __forceinline__ void issue_weights(const Smem& sm, const Desc& d) {
  const Slice s = slice_of(d, blockIdx.x, gridDim.x);
  if (s.n_hi <= s.n_lo) return;
  const int ld = s.kc + kPad;
  const unsigned bytes = (unsigned)(s.n_hi - s.n_lo) * (unsigned)ld * 4u;
  uint64_t* bar = &sm.w_full[d.hi];
  mbar_expect_tx(bar, bytes);
  bulk_g2s(weight_base(sm, d, s), d.W + ((size_t)s.ks * d.N + s.n_lo) * ld, bytes, bar);
}

__device__ __forceinline__ void stage_tile(const Args& a, const Smem& sm, const Desc& d, int g) {
  const Slice s = slice_of(d, blockIdx.x, gridDim.x);
  const int b0 = g * a.group_rows, rows = min(a.group_rows, a.st.batch - b0);
  if (d.kind != kGemm || d.zero_x || s.n_hi <= s.n_lo || rows <= 0) {
    mbar_arrive(sm.x_full);
    return;
  }
  const int ld = s.kc + kXPad;
  if (d.ldx == ld) {   // stored at the shared-memory stride (K-split-major if split): the group's rows are one run
    const unsigned bytes = (unsigned)rows * (unsigned)ld * 4u;
    mbar_expect_tx(sm.x_full, bytes);
    bulk_g2s(sm.xs, d.X + ((size_t)s.ks * a.st.batch + b0) * d.ldx, bytes, sm.x_full);
    return;
  }
  const unsigned row_bytes = (unsigned)s.kc * 4u;
  mbar_expect_tx(sm.x_full, (unsigned)rows * row_bytes);
  for (int r = 0; r < rows; ++r)
    bulk_g2s(sm.xs + r * ld, d.X + (size_t)(b0 + r) * d.ldx + s.k_lo, row_bytes, sm.x_full);
}

// C[16 rows][8 NT cols] += A (activation fragments) x B (weight rows from shared memory), 3 x TF32; the LayerNorm
// partial sums (about the row's first element) ride along on the FMA pipe while the tensor pipe works.
// A fragments come from ldmatrix (one instruction = the four registers of an m16n8k8 TF32 A operand, already in
// consecutive registers: assembling them from 128-bit loads cost ~4 MOVs per MMA).  xq[j][s] covers k8-step s of
// chunk j: {(row gq, k tq), (row gq+8, k tq), (row gq, k tq+4), (row gq+8, k tq+4)}.  The packed weight rows are
// permuted within every 16-float chunk so that floats 4tq..4tq+3 of a row are k = tq, tq+4, tq+8, tq+12.
// FULL: every warp owns exactly kChunks chunks and NTMAX n-tiles -> no branches, the compiler interleaves freely.
struct RowStats {
  float s0[2], q0[2], s1[2], q1[2];   // rows gq / gq+8, two interleaved accumulators each
};
template <int NTMAX, bool FULL>
__device__ __forceinline__ void mma_tiles(float (&acc)[3][2][4], const float (&xq)[kChunks][2][4], const float* wb, int ld,
                                          int nch, int warp, int nt_run, float sh0, float sh1, RowStats& rs) {
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const int c = warp + kCWarps * j;
    if (FULL || c < nch) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(xq[j][st][i], ah[st][i], al[st][i]);
#pragma unroll
      for (int nt = 0; nt < NTMAX; ++nt) {
        if (FULL || nt < nt_run) {
          const float4 wv = lds4(wb + nt * 8 * ld + c * 16);
          uint32_t bh[4], bl[4];
          split_tf32(wv.x, bh[0], bl[0]); split_tf32(wv.y, bh[1], bl[1]);
          split_tf32(wv.z, bh[2], bl[2]); split_tf32(wv.w, bh[3], bl[3]);
          mma_tf32(acc[nt][0], al[0][0], al[0][1], al[0][2], al[0][3], bh[0], bh[1]);
          mma_tf32(acc[nt][1], al[1][0], al[1][1], al[1][2], al[1][3], bh[2], bh[3]);
          mma_tf32(acc[nt][0], ah[0][0], ah[0][1], ah[0][2], ah[0][3], bl[0], bl[1]);
          mma_tf32(acc[nt][1], ah[1][0], ah[1][1], ah[1][2], ah[1][3], bl[2], bl[3]);
          mma_tf32(acc[nt][0], ah[0][0], ah[0][1], ah[0][2], ah[0][3], bh[0], bh[1]);
          mma_tf32(acc[nt][1], ah[1][0], ah[1][1], ah[1][2], ah[1][3], bh[2], bh[3]);
        }
      }
      {
        const float d0 = xq[j][0][0] - sh0, d1 = xq[j][0][2] - sh0, d2 = xq[j][1][0] - sh0, d3 = xq[j][1][2] - sh0;
        const float e0 = xq[j][0][1] - sh1, e1 = xq[j][0][3] - sh1, e2 = xq[j][1][1] - sh1, e3 = xq[j][1][3] - sh1;
        rs.s0[0] += d0 + d2; rs.s0[1] += d1 + d3;
        rs.q0[0] = fmaf(d0, d0, fmaf(d2, d2, rs.q0[0])); rs.q0[1] = fmaf(d1, d1, fmaf(d3, d3, rs.q0[1]));
        rs.s1[0] += e0 + e2; rs.s1[1] += e1 + e3;
        rs.q1[0] = fmaf(e0, e0, fmaf(e2, e2, rs.q1[0])); rs.q1[1] = fmaf(e1, e1, fmaf(e3, e3, rs.q1[1]));
      }
    }
  }
}

// ---- GEMM group-phase (consumers) --------------------------------------------------------------------------
struct CState {
  unsigned gp;            // group-phases consumed so far (parity of x_full)
  unsigned wpar0, wpar1;  // parity of the two weight barriers (scalars: a dynamically indexed array would live in local memory)
  unsigned ring_seq;      // K/V ring tiles of all attention units this CTA has finished (slot and parity of the next)
  unsigned q_uses;        // prefetched queries consumed (slot = uses & 1, parity = (uses >> 1) & 1)
};

template <int DH, bool DROP>
__device__ __forceinline__ void store_out(const Args& a, const Desc& d, const Smem& sm, const Slice& s, int b, int n, int t,
                                          float v, float res) {
  if (DROP && d.drop_thr != 0u && d.mode != kFinal && d.mode != kPrenetOut)   // CTA-uniform: decode in decoder.train() mode only
    v = drop1(a.seed, d.drop_thr, d.drop_sc, d.drop_site, ((unsigned long long)t * a.st.batch + b) * (unsigned)d.N + (unsigned)n, v);
  switch (d.mode) {
    case kPlain:
      d.Y[(size_t)b * d.ldy + n] = v * d.out_scale + res;
      break;
    case kPartial:
      d.Y[((size_t)s.ks * a.st.batch + b) * d.ldy + n] = v;
      break;
    case kHid: {   // FFN hidden, K-split-major for the FFN-out phase: [n / kc][b][n % kc], row stride ldy = kc + kXPad
      const int kc = (int)d.ldy - kXPad;
      int sl = 0;
      for (int e = kc; e <= n; e += kc) ++sl;   // n / kc for a handful of slices, without a runtime division
      d.Y[((size_t)sl * a.st.batch + b) * d.ldy + (n - sl * kc)] = v;
    } break;
    case kQkv: {
      const int H = a.w.n_heads, D = H * DH;
      const int which = (n >= D) + (n >= 2 * D), cc = n - which * D;   // q | k | v without a runtime division
      if (which == 0) {
        d.Y[(size_t)b * d.ldy + cc] = v * d.out_scale;
      } else {
        const int h = cc / DH, dd = cc - h * DH;
        float* dst = which == 1 ? d.kcache : d.vcache;
        dst[(((size_t)b * H + h) * a.st.t_max + t) * DH + dd] = v;
      }
    } break;
    case kPrenetOut: {  // modules.py:114-118
      const bool have = t > 0 && (t - 1) < sm.len[b];
      float o = (have ? v : 0.f) + __ldg(a.w.pe_table + (size_t)t * d.N + n) * __ldg(a.w.pe_scale);
      if (DROP && d.drop_thr != 0u)   // modules.py:120: dropout on the sum
        o = drop1(a.seed, d.drop_thr, d.drop_sc, d.drop_site, ((unsigned long long)t * a.st.batch + b) * (unsigned)d.N + (unsigned)n, o);
      d.Y[(size_t)b * d.ldy + n] = o;
    } break;
    case kFinal: {  // modules.py:144, tacotron.py:112-115
      const bool on = t < sm.len[b];
      if (n < a.w.n_mels) a.st.frames[((size_t)b * a.st.t_max + t) * a.w.n_mels + n] = on ? v : 0.f;
      else if (n == a.w.n_mels) a.st.stop_logits[(size_t)b * a.st.t_max + t] = on ? v + __ldg(a.w.b_stop) : 0.f;
      else {   // first prenet layer of the NEXT step (its dropout uses that step's element index, site 1)
        const int pn = n - a.w.n_mels - 1;
        float o = fmaxf(v, 0.f);
        if (DROP && a.thr_d != 0u)
          o = drop1(a.seed, a.thr_d, a.sc_d, 1u, ((unsigned long long)(t + 1) * a.st.batch + b) * (unsigned)a.w.prenet_hidden + (unsigned)pn, o);
        a.p0[(size_t)b * (a.w.prenet_hidden + kXPad) + pn] = o;
      }
    } break;
  }
}

template <int DH, bool DROP>
__device__ __forceinline__ void gemm_group(const Args& a, const Desc& d, const Smem& sm, CState& cs, int g, int t,
                                           bool first_group, long long* prof) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int B = a.st.batch;
  const int b0 = g * a.group_rows, rows = min(a.group_rows, B - b0);
  const Slice s = slice_of(d, blockIdx.x, gridDim.x);
  const bool has_rows = s.n_hi > s.n_lo;
  const int ld = s.kc + kPad, nch = s.kc >> 4;
  const int ncols = s.n_hi - s.n_lo, NT = (ncols + 7) >> 3;

  // ---- activation fragments (ldmatrix.x4: lanes 8i..8i+7 give the row addresses of 8x4-float matrix i =
  //      {rows 0-7 | 8-15} x {k 0-3 | 4-7} of the k8-step), chunks {warp, warp+8, ...} of 16 floats
  float xq[kChunks][2][4];
  float sh0 = 0.f, sh1 = 0.f;
  const bool live = has_rows && !d.zero_x;
  const int ldx = s.kc + kXPad;
  {
    const int mid = lane >> 3, rin = lane & 7;
    const uint32_t lane_addr = smem_u32(sm.xs + (rin + (mid & 1) * 8) * ldx + (mid >> 1) * 4);
#pragma unroll
    for (int j = 0; j < kChunks; ++j) {
      const int c = warp + kCWarps * j;
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        if (live && c < nch) {
          uint32_t r0, r1, r2, r3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(lane_addr + (uint32_t)(c * 16 + st * 8) * 4u));
          xq[j][st][0] = __uint_as_float(r0); xq[j][st][1] = __uint_as_float(r1);
          xq[j][st][2] = __uint_as_float(r2); xq[j][st][3] = __uint_as_float(r3);
        } else {
          xq[j][st][0] = xq[j][st][1] = xq[j][st][2] = xq[j][st][3] = 0.f;
        }
      }
    }
  }
  if (live && d.ln) {
    sh0 = sm.xs[gq * ldx];
    sh1 = sm.xs[(gq + 8) * ldx];
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(sm.x_empty);   // the tile lives in registers now: the producer may refill the slot
  if (!has_rows) return;
  if (prof) prof[6] = clock64();

  // epilogue mapping: thread (row = tid / 16, cg = tid % 16) finishes columns cg and cg + 16 of its row.
  // The residual is requested now so that its latency hides behind the products.
  const int erow = tid >> 4, ecg = tid & 15;
  const bool on0 = erow < rows && ecg < ncols, on1 = erow < rows && ecg + 16 < ncols;
  float res0 = 0.f, res1 = 0.f;
  if (d.R != nullptr) {
    const float* rp = d.R + (size_t)(b0 + erow) * d.ldr + s.n_lo + ecg;
    if (on0) res0 = __ldcg(rp);
    if (on1) res1 = __ldcg(rp + 16);
  }

  if (prof) prof[11] = clock64();
  if (first_group) {
    mbar_wait(&sm.w_full[d.hi], d.hi ? cs.wpar1 : cs.wpar0, a.err);
    if (d.hi) cs.wpar1 ^= 1u;
    else cs.wpar0 ^= 1u;
  }
  if (prof) prof[7] = clock64();

  // ---- products
  float acc[3][2][4];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][h][i] = 0.f;
  RowStats rs;
  rs.s0[0] = rs.s0[1] = rs.q0[0] = rs.q0[1] = rs.s1[0] = rs.s1[1] = rs.q1[0] = rs.q1[1] = 0.f;
  const float* wbase = weight_base(sm, d, s);
  const float* wb = wbase + gq * ld + 4 * tq;
  if (nch == kChunks * kCWarps) {   // K slice of 768: every warp owns exactly kChunks chunks -> straight-line code
    if (NT == 1) mma_tiles<1, true>(acc, xq, wb, ld, nch, warp, NT, sh0, sh1, rs);
    else if (NT == 2) mma_tiles<2, true>(acc, xq, wb, ld, nch, warp, NT, sh0, sh1, rs);
    else mma_tiles<3, true>(acc, xq, wb, ld, nch, warp, NT, sh0, sh1, rs);
  } else {
    mma_tiles<3, false>(acc, xq, wb, ld, nch, warp, NT, sh0, sh1, rs);
  }
  if (prof) prof[8] = clock64();
  if (d.ln) {  // row statistics: merge the 4 lanes that share a row, one record per warp and row
    float s0 = rs.s0[0] + rs.s0[1], q0 = rs.q0[0] + rs.q0[1], s1 = rs.s1[0] + rs.s1[1], q1 = rs.q1[0] + rs.q1[1];
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if (tq == 0) {
      *reinterpret_cast<float2*>(sm.spart + (warp * kGroupRows + gq) * 2) = make_float2(s0, q0);
      *reinterpret_cast<float2*>(sm.spart + (warp * kGroupRows + gq + 8) * 2) = make_float2(s1, q1);
      if (warp == 0) {
        sm.sshift[gq] = sh0;
        sm.sshift[gq + 8] = sh1;
      }
    }
  }
  {  // fragments -> cross-warp buffer [warp][row][24]
    float* rw = sm.red + warp * (kGroupRows * kMaxRows);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      if (nt < NT) {
        *reinterpret_cast<float2*>(rw + gq * kMaxRows + nt * 8 + 2 * tq) =
            make_float2(acc[nt][0][0] + acc[nt][1][0], acc[nt][0][1] + acc[nt][1][1]);
        *reinterpret_cast<float2*>(rw + (gq + 8) * kMaxRows + nt * 8 + 2 * tq) =
            make_float2(acc[nt][0][2] + acc[nt][1][2], acc[nt][0][3] + acc[nt][1][3]);
      }
    }
  }
  consumer_bar();
  if (prof) prof[9] = clock64();

  // ---- epilogue: every load first (branch-free), then the arithmetic of both outputs
  {
    const int nl0 = min(ecg, NT * 8 - 1), nl1 = min(ecg + 16, NT * 8 - 1);   // clamped: loads stay inside the slice's tiles
    const float* rr = sm.red + erow * kMaxRows;
    float v0 = 0.f, v1 = 0.f, S = 0.f, Q = 0.f;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) {
      v0 += rr[w * (kGroupRows * kMaxRows) + nl0];
      v1 += rr[w * (kGroupRows * kMaxRows) + nl1];
      const float2 p = *reinterpret_cast<const float2*>(sm.spart + (w * kGroupRows + erow) * 2);
      S += p.x;
      Q += p.y;
    }
    const float2 c0 = *reinterpret_cast<const float2*>(wbase + nl0 * ld + s.kc);   // (constant, LayerNorm row sum)
    const float2 c1 = *reinterpret_cast<const float2*>(wbase + nl1 * ld + s.kc);
    if (d.ln) {  // LayerNorm (eps 1e-6, modules.py:88) applied to the finished product
      const float invK = d.inv_k;
      const float ms = S * invK;
      const float var = fmaxf(Q * invK - ms * ms, 0.f);
      const float rstd = rsqrtf(var + 1e-6f);
      const float nm = -rstd * (sm.sshift[erow] + ms);
      v0 = fmaf(rstd, v0, nm * c0.y);
      v1 = fmaf(rstd, v1, nm * c1.y);
    }
    v0 += c0.x;
    v1 += c1.x;
    if (d.relu) {
      v0 = fmaxf(v0, 0.f);
      v1 = fmaxf(v1, 0.f);
    }
    if (on0) store_out<DH, DROP>(a, d, sm, s, b0 + erow, s.n_lo + ecg, t, v0, res0);
    if (on1) store_out<DH, DROP>(a, d, sm, s, b0 + erow, s.n_lo + ecg + 16, t, v1, res1);
  }
  if (prof) prof[10] = clock64();
}

// ---- reduce group-phase: x[b][n] += sum_s part[s][b][n] (FFN-out partials + residual) ------------------------
template <bool DROP>
__device__ __forceinline__ void reduce_group(const Args& a, const Desc& d, int g, int t) {
  const int B = a.st.batch, G = gridDim.x, c = blockIdx.x;
  const int b0 = g * a.group_rows, rows = min(a.group_rows, B - b0);
  const int n_lo = d.s_n_lo, n_hi = d.s_n_hi, ncols = n_hi - n_lo;
  (void)G; (void)c;
  for (int idx = threadIdx.x; idx < rows * ncols; idx += kConsumers) {
    const int row = idx / ncols, n = n_lo + idx - row * ncols, b = b0 + row;
    const float x0 = __ldcg(d.Y + (size_t)b * d.ldy + n);
    float v = 0.f;
    for (int s0 = 0; s0 < d.n_parts; s0 += 4) {   // four independent loads in flight
      float pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pv[i] = s0 + i < d.n_parts ? __ldcg(d.part + ((size_t)(s0 + i) * B + b) * d.N + n) : 0.f;
      v += (pv[0] + pv[1]) + (pv[2] + pv[3]);
    }
    if (DROP && d.drop_thr != 0u)   // modules.py:141: x + dropout(ffn(x))
      v = drop1(a.seed, d.drop_thr, d.drop_sc, d.drop_site, ((unsigned long long)t * B + b) * (unsigned)d.N + (unsigned)n, v);
    d.Y[(size_t)b * d.ldy + n] = x0 + v;
  }
}

// ---- attention group-phase -----------------------------------------------------------------------------------
// The K/V stream of a unit = (sample, head[, key split]) flows through ONE CTA-wide ring of kSlots tiles (8 K rows +
// 8 V rows each) that the feeder warp keeps full with TMA bulk copies; a consumer warp takes the next tile nobody
// has taken yet (shared counter), so fast and slow warps balance themselves and no consumer instruction is spent
// on issuing copies.  Scores are kept in the log2 domain (q is pre-multiplied by log2 e): one EX2 per weight.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

struct UnitRange {
  int item, split, j0, j1, n_tiles;
};
__device__ __forceinline__ UnitRange unit_range(int u, int ns, int n_keys, int item0) {
  UnitRange r;
  if (ns == 1) {   // one stream per (sample, head): no divisions on the consumers' critical path
    r.split = 0; r.item = item0 + u; r.j0 = 0; r.j1 = n_keys; r.n_tiles = (n_keys + kTK - 1) / kTK;
    return r;
  }
  const int per = (n_keys + ns - 1) / ns;
  const int litem = ns == 1 ? u : u / ns;
  r.split = u - litem * ns;
  r.item = item0 + litem;
  r.j0 = min(n_keys, r.split * per);
  r.j1 = min(n_keys, r.j0 + per);
  r.n_tiles = (r.j1 - r.j0 + kTK - 1) / kTK;
  return r;
}

template <int DH>
__device__ __forceinline__ float* slot_ptr(const Args& a, const Smem& sm, unsigned slot) {
  constexpr int kSlotF = 2 * kTK * DH;
  return (int)slot < a.n_hi ? sm.wreg + a.ring_lo + (size_t)slot * kSlotF : sm.wreg + (size_t)((int)slot - a.n_hi) * kSlotF;
}

// feeder warp: lane `sl` owns ring slot `sl`.  Every lane polls (non-blocking) whether its slot has been drained
// and, if so, issues the next tile that maps to it; the lanes never block each other, so a freed slot is refilled
// within one polling round (a single issuing thread was 2x too slow: ~700 cycles per tile, measured).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int DH>
__device__ __forceinline__ void feed_group(const Args& a, const Smem& sm, const float* kc, const float* vc, int rows_alloc,
                                           int n_keys, int g, unsigned& f_seq, unsigned go_target, bool gate_last,
                                           unsigned low_target, unsigned& f_q) {
  // go_target: value of sm.kv_go once the loader has seen the previous phase of this group complete everywhere.
  // Cross K/V never changes during decode and self K/V rows < t were written in earlier steps, so only the tile
  // that holds the row appended in this step (gate_last: the last key of the stream) has to wait for it; all
  // other tiles are requested as soon as their ring slot is free, i.e. while the consumers still run the GEMMs
  // in front of the attention phase.
  constexpr int kTile = kTK * DH, kSlotF = 2 * kTile;
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, ns = a.n_split, lane = threadIdx.x & 31;
  const int b0 = g * a.group_rows, rows = min(a.group_rows, B - b0);
  const int n_units = rows * H * ns, item0 = b0 * H;
  const unsigned nsl = (unsigned)a.n_slots;
  int total = 0;
  for (int u = blockIdx.x; u < n_units; u += G) total += unit_range(u, ns, n_keys, item0).n_tiles;
  // local index of the first tile of this group-phase that lands in my slot
  int k = lane < (int)nsl ? (int)((lane + nsl - f_seq % nsl) % nsl) : total;
  int u = blockIdx.x, base = 0;
  UnitRange r = unit_range(u, ns, n_keys, item0);
  const unsigned* my_drained = &sm.drained[lane < (int)nsl ? lane : 0];
  uint64_t* my_full = &sm.rfull[lane < (int)nsl ? lane : 0];
  float* dst = slot_ptr<DH>(a, sm, lane < (int)nsl ? lane : 0);
  bool low_ok = lane < a.n_hi;   // low slots: only after the GEMM in front of the attention phase (pdone >= low_target)
  // lane 31: the query of the first unit (written in the phase before this one) -> shared memory, as soon as the
  // loader has seen that phase complete; the consumers then find it without an L2 round trip
  bool q_todo = lane == 31 && (int)blockIdx.x < n_units;
  long long spins = 0, t0 = 0;
  while (__any_sync(0xffffffffu, k < total || q_todo)) {
    if (q_todo) {
      unsigned gv;
      asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(gv) : "r"(smem_u32(sm.kv_go)) : "memory");
      if (static_cast<int>(gv - go_target) >= 0) {
        fence_proxy_async();
        uint64_t* qb = &sm.qfull[f_q & 1u];
        mbar_expect_tx(qb, (unsigned)DH * 4u);
        bulk_g2s(sm.qbuf + (f_q & 1u) * 96, a.q + (size_t)unit_range(blockIdx.x, ns, n_keys, item0).item * DH, (unsigned)DH * 4u, qb);
        q_todo = false;
      }
    }
    if (k < total) {
      bool continue_spin = false;
      const unsigned seq = f_seq + (unsigned)k;
      unsigned dv;   // use number seq / nsl of my slot needs that many earlier tiles drained
      asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(dv) : "r"(smem_u32(my_drained)) : "memory");
      if (!low_ok) {
        unsigned pv;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(pv) : "r"(smem_u32(sm.pdone)) : "memory");
        low_ok = static_cast<int>(pv - low_target) >= 0;
      }
      if (low_ok && static_cast<int>(dv - seq / nsl) >= 0) {
        while (k >= base + r.n_tiles) {   // unit that contains local tile k
          base += r.n_tiles;
          u += G;
          r = unit_range(u, ns, n_keys, item0);
        }
        const int key0 = r.j0 + (k - base) * kTK, nk = min(kTK, r.j1 - key0);
        if (gate_last && key0 + nk == n_keys) {   // the freshly appended row: visible once kv_go says so
          unsigned gv;
          asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(gv) : "r"(smem_u32(sm.kv_go)) : "memory");
          if (static_cast<int>(gv - go_target) < 0) continue_spin = true;
          else fence_proxy_async();
        }
        if (!continue_spin) {
        const unsigned bytes = (unsigned)nk * DH * 4u;
        const size_t off = ((size_t)r.item * rows_alloc + key0) * DH;
        mbar_expect_tx(my_full, 2u * bytes);
        bulk_g2s(dst, kc + off, bytes, my_full);
        bulk_g2s(dst + kTile, vc + off, bytes, my_full);
        k += (int)nsl;
        }
      }
    }
    if ((++spins & 1023) == 0) {
      if (t0 == 0) t0 = clock64();
      if (clock64() - t0 > kTimeoutCycles || *reinterpret_cast<volatile int*>(a.err) != 0) {
        atomicExch(a.err, 4);
        break;
      }
    }
  }
  f_seq += (unsigned)total;
  if ((int)blockIdx.x < n_units) ++f_q;
}

// one tile of 8 keys: scores, online softmax update, weighted V.  FULL: all 8 keys exist and none is masked.
struct DropA {   // dropout on the attention weights of one stream (attention.py:89): element = ebase + key index
  uint32_t thr; float sc; unsigned long long seed; uint32_t site; unsigned long long ebase;
};
template <int DH, bool FULL, bool DROP>
__device__ __forceinline__ void attn_tile(const float* kt, const float* vt, const f32x4 (&qv)[DH / 32], int nk, int key0,
                                          int klen, float* logit_dst, float& m_run, float& l_run,
                                          f32x2 (&o)[DH / 32][2], int kslot, int l8, const DropA& da) {
  constexpr int F4 = DH / 32, kRounds = kTK / 4;
  float sv[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int kl = r * 4 + kslot;
    f32x2 acc0 = 0ull, acc1 = 0ull;
#pragma unroll
    for (int i = 0; i < F4; ++i) {
      const f32x4 kv = ld4s(kt + kl * DH + 4 * (l8 + 8 * i));
      acc0 = fma2(qv[i].lo, kv.lo, acc0);
      acc1 = fma2(qv[i].hi, kv.hi, acc1);
    }
    float x0, x1, y0, y1;
    unpack2(acc0, x0, y0);
    unpack2(acc1, x1, y1);
    sv[r] = (x0 + y0) + (x1 + y1);
  }
#pragma unroll
  for (int o8 = 1; o8 <= 4; o8 <<= 1)
#pragma unroll
    for (int r = 0; r < kRounds; ++r) sv[r] += __shfl_xor_sync(0xffffffffu, sv[r], o8);
  float mt = -CUDART_INF_F;
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int kl = r * 4 + kslot;
    if (!FULL) {
      if (kl >= nk) sv[r] = -CUDART_INF_F;                  // stale shared memory beyond the tile's keys
      else if (key0 + kl >= klen) sv[r] = kNegBias;         // logits + (-1e20), attention.py:84-85
    }
    if (logit_dst != nullptr && l8 == 0 && (FULL || kl < nk)) logit_dst[kl] = sv[r];
    mt = fmaxf(mt, sv[r]);
  }
  mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
  mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
  if (mt > m_run) {   // warp-uniform; rare after the first tiles
    const float corr = ex2(m_run - mt);
    l_run *= corr;
    const f32x2 c2 = pack2(corr, corr);
#pragma unroll
    for (int i = 0; i < F4; ++i) {
      o[i][0] = fma2(o[i][0], c2, 0ull);
      o[i][1] = fma2(o[i][1], c2, 0ull);
    }
    m_run = mt;
  }
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int kl = r * 4 + kslot;
    const float p = (FULL || kl < nk) ? ex2(sv[r] - m_run) : 0.f;
    if (l8 == 0) l_run += p;   // the softmax normaliser sums the weights BEFORE dropout (attention.py:87-89)
    const float pd = (DROP && da.thr != 0u) ? drop1(da.seed, da.thr, da.sc, da.site, da.ebase + (unsigned)(key0 + kl), p) : p;
    const f32x2 pp = pack2(pd, pd);
#pragma unroll
    for (int i = 0; i < F4; ++i) {
      const f32x4 vv = ld4s(vt + kl * DH + 4 * (l8 + 8 * i));
      if (FULL) {
        o[i][0] = fma2(pp, vv.lo, o[i][0]);
        o[i][1] = fma2(pp, vv.hi, o[i][1]);
      } else {
        o[i][0] = fma2(pp, kl < nk ? vv.lo : 0ull, o[i][0]);
        o[i][1] = fma2(pp, kl < nk ? vv.hi : 0ull, o[i][1]);
      }
    }
  }
}

template <int DH, bool DROP>
__device__ __forceinline__ void attn_group(const Args& a, const Desc& at, const Smem& sm, CState& cs, int g, int t,
                                           long long* prof) {
  constexpr int F4 = DH / 32;            // float4 per lane per key row (8 lanes span a row)
  constexpr int kTile = kTK * DH;        // floats per K (or V) tile
  constexpr int kSlotF = 2 * kTile;      // K tile then V tile
  constexpr int PS = DH + 4;
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, ns = a.n_split;
  const int b0 = g * a.group_rows, rows = min(a.group_rows, B - b0);
  const int n_units = rows * H * ns, n_keys = at.n_keys;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kslot = lane >> 3, l8 = lane & 7;
  float* wrec = sm.red;  // [8][PS]
  float* sc = sm.red + kCWarps * PS;                                  // scaled logits of the current unit
  const int sc_cap = kRedFloats - kCWarps * PS;
  unsigned seq_base = cs.ring_seq;

  for (int u = blockIdx.x; u < n_units; u += G) {
    const UnitRange ur = unit_range(u, ns, n_keys, b0 * H);
    const int item = ur.item, j0 = ur.j0, j1 = ur.j1;
    const bool sc_ok = j1 - j0 <= sc_cap;                             // else fall back to read-modify-write in HBM
    const int b = item / H;
    const int klen = at.key_len ? sm.klen[b] : n_keys;
    float* arow = at.align ? at.align + (size_t)item * at.align_bh_stride + (size_t)t * at.align_row_len : nullptr;
    if (prof) prof[12] = clock64();

    f32x4 qv[F4];
    {
      const f32x2 sc2 = pack2(kLog2e, kLog2e);
      const bool pre = u == (int)blockIdx.x;   // the first unit's query was prefetched into shared memory by the feeder
      const float* qsrc = a.q + (size_t)item * DH;
      if (pre) {
        mbar_wait(&sm.qfull[cs.q_uses & 1u], (cs.q_uses >> 1) & 1u, a.err);
        qsrc = sm.qbuf + (cs.q_uses & 1u) * 96;
        ++cs.q_uses;
      }
      if (prof) prof[13] = clock64();
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        qv[i] = pre ? ld4s(qsrc + 4 * (l8 + 8 * i)) : ld4cg(qsrc + 4 * (l8 + 8 * i));
        qv[i].lo = fma2(qv[i].lo, sc2, 0ull);
        qv[i].hi = fma2(qv[i].hi, sc2, 0ull);
      }
    }
    float m_run = -CUDART_INF_F, l_run = 0.f;
    f32x2 o[F4][2];
#pragma unroll
    for (int i = 0; i < F4; ++i) o[i][0] = o[i][1] = 0ull;
    if (prof) prof[6] = clock64();
    const DropA da{at.drop_thr, at.drop_sc, a.seed, at.drop_site, ((unsigned long long)t * (unsigned)(B * H) + (unsigned)item) << 16};

    // slot and use number of this warp's first tile; advanced incrementally (no division in the loop)
    unsigned slot = (seq_base + (unsigned)warp) % (unsigned)a.n_slots, use = (seq_base + (unsigned)warp) / (unsigned)a.n_slots;
    for (int ti = warp; ti < ur.n_tiles; ti += kCWarps) {   // static assignment: bit-reproducible sums
      const int key0 = j0 + ti * kTK, nk = min(kTK, j1 - key0);
      wait_count(a, &sm.drained[slot], use);      // the slot's previous tile has been consumed (no parity aliasing)
      mbar_wait(&sm.rfull[slot], use & 1u, a.err);
      const float* kt = slot_ptr<DH>(a, sm, slot);
      float* ldst = arow == nullptr ? nullptr : ((ns == 1 && sc_ok) ? sc + (key0 - j0) : arow + key0);
      if (nk == kTK && key0 + kTK <= klen)
        attn_tile<DH, true, DROP>(kt, kt + kTile, qv, nk, key0, klen, ldst, m_run, l_run, o, kslot, l8, da);
      else
        attn_tile<DH, false, DROP>(kt, kt + kTile, qv, nk, key0, klen, ldst, m_run, l_run, o, kslot, l8, da);
      __syncwarp();   // every lane is done with this slot before it is refilled
      if (lane == 0)
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&sm.drained[slot])), "r"(use + 1u) : "memory");
      slot += kCWarps;
      while (slot >= (unsigned)a.n_slots) {
        slot -= (unsigned)a.n_slots;
        ++use;
      }
    }
    seq_base += (unsigned)ur.n_tiles;
    if (prof) prof[7] = clock64();
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
          const int dd = 4 * (l8 + 8 * i) + 2 * hs;
          wrec[warp * PS + dd] = x;
          wrec[warp * PS + dd + 1] = y;
        }
      }
    const float lw = warp_sum(l_run);
    if (lane == 0) {
      wrec[warp * PS + DH] = m_run;
      wrec[warp * PS + DH + 1] = lw;
    }
    consumer_bar();
    if (prof) prof[8] = clock64();
    float m = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) m = fmaxf(m, wrec[w * PS + DH]);
    float l = 0.f;
    float wgt[kCWarps];
#pragma unroll
    for (int w = 0; w < kCWarps; ++w) {
      const float mw = wrec[w * PS + DH];
      wgt[w] = mw > -CUDART_INF_F ? ex2(mw - m) : 0.f;
      l = fmaf(wrec[w * PS + DH + 1], wgt[w], l);
    }
    if (tid < DH) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kCWarps; ++w) v = fmaf(wrec[w * PS + tid], wgt[w], v);
      if (ns == 1) a.ctx[(size_t)b * (H * DH + kXPad) + (item - b * H) * DH + tid] = v / l;
      else a.fpart[((size_t)item * ns + ur.split) * PS + tid] = v;
    }
    if (ns == 1) {
      if (arow != nullptr) {
        const float inv = 1.f / l;
        if (sc_ok) for (int j = j0 + tid; j < j1; j += kConsumers) arow[j] = ex2(sc[j - j0] - m) * inv;
        else for (int j = j0 + tid; j < j1; j += kConsumers) arow[j] = ex2(arow[j] - m) * inv;
      }
    } else if (tid == 0) {
      a.fpart[((size_t)item * ns + ur.split) * PS + DH] = m;       // -inf when the split is empty
      a.fpart[((size_t)item * ns + ur.split) * PS + DH + 1] = l;
    }
    if (prof) prof[9] = clock64();
    consumer_bar();  // wrec / sc are reused by the next unit
    if (prof) prof[10] = clock64();
  }
  cs.ring_seq = seq_base;
}

// ---- combine group-phase (only when the K/V streams were split, i.e. small batches): partials -> ctx, align rows
template <int DH>
__device__ __forceinline__ void combine_group(const Args& a, const Desc& d, const Smem& sm, int g, int t) {
  constexpr int PS = DH + 4;
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x, ns = a.n_split;
  const int b0 = g * a.group_rows, rows = min(a.group_rows, B - b0);
  const int tid = threadIdx.x;
  float* ml = sm.red;
  for (int li = blockIdx.x; li < rows * H; li += G) {
    const int item = b0 * H + li;
    const float* pr = a.fpart + (size_t)item * ns * PS;
    if (tid == 0) {
      float m = -CUDART_INF_F;
      for (int sidx = 0; sidx < ns; ++sidx) m = fmaxf(m, __ldcg(pr + sidx * PS + DH));
      float l = 0.f;
      for (int sidx = 0; sidx < ns; ++sidx) {
        const float pm = __ldcg(pr + sidx * PS + DH);
        if (pm > -CUDART_INF_F) l += __ldcg(pr + sidx * PS + DH + 1) * ex2(pm - m);
      }
      ml[0] = m;
      ml[1] = 1.f / l;
    }
    consumer_bar();
    const float m = ml[0], inv = ml[1];
    if (tid < DH) {
      float acc = 0.f;
      for (int sidx = 0; sidx < ns; ++sidx) {
        const float pm = __ldcg(pr + sidx * PS + DH);
        if (pm > -CUDART_INF_F) acc = fmaf(__ldcg(pr + sidx * PS + tid), ex2(pm - m), acc);
      }
      a.ctx[(size_t)(item / H) * (H * DH + kXPad) + (item % H) * DH + tid] = acc * inv;
    }
    if (d.align != nullptr) {  // raw logits of this step -> softmax weights
      float* row = d.align + (size_t)item * d.align_bh_stride + (size_t)t * d.align_row_len;
      for (int j = tid; j < d.n_keys; j += kConsumers) row[j] = ex2(__ldcg(row + j) - m) * inv;
    }
    consumer_bar();
  }
}

// ---- phase table ---------------------------------------------------------------------------------------------
// per step: 3 prenet GEMMs, per layer {qkv, self, [combine], oproj, cq, cross, [combine], coproj, ffn1, ffn2, [reduce]},
// final projection.  `comb` = attention streams are split (n_split > 1), `red` = FFN-out is K-split.
__device__ __forceinline__ int phases_per_layer(const Args& a) { return 8 + (a.n_split > 1 ? 2 : 0) + (a.ksplit > 1 ? 1 : 0); }
__device__ __forceinline__ int n_phases(const Args& a) { return 3 + phases_per_layer(a) * a.w.n_layers + 1; }

// canonical id of phase `ph` (see get_phase_body): 0 qkv, 1 self, 2 comb, 3 oproj, 4 cq, 5 cross, 6 comb, 7 coproj,
// 8 ffn1, 9 ffn2, 10 reduce; 100 + i for the prenet GEMMs, 200 for the final projection
__device__ __forceinline__ int phase_id(const Args& a, int ph, int& layer) {
  const int ppl = phases_per_layer(a);
  layer = 0;
  if (ph < 3) return 100 + ph;
  if (ph == 3 + ppl * a.w.n_layers) return 200;
  layer = (ph - 3) / ppl;
  const int k = (ph - 3) % ppl;
  if (a.n_split > 1) return k;
  return k < 2 ? k : (k < 5 ? k + 1 : k + 2);
}

// Ask the memory system to pull the K/V rows this CTA will stream in an upcoming attention phase into L2 while the
// GEMM phases leave HBM idle (the ring then refills at L2 latency instead of HBM latency).  Pure hint.
template <int DH>
__device__ __forceinline__ void prefetch_kv(const Args& a, const float* kc, const float* vc, int rows_alloc, int rows) {
  if (a.n_split != 1 || rows <= 0) return;
  const int B = a.st.batch, H = a.w.n_heads, G = gridDim.x;
  // keep the total under ~88 MB so that it survives in the 126 MB L2 until it is used
  const long long cap = (88ll << 20) / ((long long)B * H * DH * 8);
  const int r = (int)min((long long)rows, cap > 1 ? cap : 1);
  for (int g = 0; g < a.n_groups; ++g) {
    const int b0 = g * a.group_rows, n_items = min(a.group_rows, B - b0) * H;
    for (int u = blockIdx.x; u < n_items; u += G) {
      const size_t off = (size_t)(b0 * H + u) * rows_alloc * DH;
      const unsigned bytes = (unsigned)r * DH * 4u;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kc + off), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vc + off), "r"(bytes) : "memory");
    }
  }
}

template <int DH>
__device__ __forceinline__ void get_phase_body(const Args& a, int ph, int t, float qscale, Desc& d);
template <int DH>
__device__ __forceinline__ void get_phase(const Args& a, int ph, int t, float qscale, Desc& d) {
  get_phase_body<DH>(a, ph, t, qscale, d);
  if (d.kind == kGemm) {
    d.inv_k = 1.f / (float)d.K;
    const Slice s = compute_slice(d, blockIdx.x, gridDim.x);
    d.s_n_lo = s.n_lo; d.s_n_hi = s.n_hi; d.s_k_lo = s.k_lo; d.s_kc = s.kc; d.s_ks = s.ks;
  } else if (d.kind == kReduce) {
    d.s_n_lo = (int)((blockIdx.x * (unsigned)d.N) / gridDim.x);
    d.s_n_hi = (int)(((blockIdx.x + 1u) * (unsigned)d.N) / gridDim.x);
  }
}
template <int DH>
__device__ __forceinline__ void get_phase_body(const Args& a, int ph, int t, float qscale, Desc& d) {
  const int B = a.st.batch, D = a.w.d_model, H = a.w.n_heads, F = a.w.d_ffn, P = a.w.prenet_hidden;
  const int M = a.w.n_mels, S = a.st.mem_len, T = a.st.t_max, L = a.w.n_layers;
  const int ppl = phases_per_layer(a);
  const bool comb = a.n_split > 1, red = a.ksplit > 1;
  memset(&d, 0, sizeof(d));
  d.out_scale = 1.f;
  d.kind = kGemm;
  d.ksplit = 1;
  if (ph == 0) {         // prenet (tacotron.py:55-65)
    d.X = a.st.frames + (size_t)(t > 0 ? t - 1 : 0) * M; d.ldx = (long long)T * M; d.zero_x = t == 0;
    d.K = M; d.N = P; d.W = a.w.pk_pre0; d.relu = 1;
    d.mode = kPlain; d.Y = a.p0; d.ldy = P + kXPad; d.hi = 1;
    d.drop_thr = a.thr_d; d.drop_sc = a.sc_d; d.drop_site = 1u;
    return;
  }
  if (ph == 1) {
    d.X = a.p0; d.ldx = P + kXPad; d.K = P; d.N = P; d.W = a.w.pk_pre1;
    d.relu = 1; d.mode = kPlain; d.Y = a.p1; d.ldy = P + kXPad; d.hi = 0;
    d.drop_thr = a.thr_d; d.drop_sc = a.sc_d; d.drop_site = 2u;
    return;
  }
  if (ph == 2) {         // + shift / mask / PE (modules.py:114-118)
    d.X = a.p1; d.ldx = P + kXPad; d.K = P; d.N = D; d.W = a.w.pk_pre2; d.mode = kPrenetOut;
    d.Y = a.x; d.ldy = D + kXPad; d.hi = 1;
    d.drop_thr = a.thr_t; d.drop_sc = a.sc_t; d.drop_site = 3u;
    return;
  }
  if (ph == 3 + ppl * L) {  // final LN + mel / stop projections
    // ... and, fused into the same rows, the first prenet layer of the NEXT step: relu(W0 mel + b0) with
    // mel = W_mel LN(x) is relu((W0 W_mel_ln) xhat + W0 c_mel + b0), so rows M+1.. of pk_final give p0 directly and
    // steps t > 0 skip the prenet's first phase (a dead row's p0 differs from the reference's relu(b0), but its
    // prenet output is masked in the third prenet phase either way: modules.py:114-116)
    d.ln = 1; d.X = a.x; d.ldx = D + kXPad; d.K = D; d.N = M + 1 + P; d.W = a.w.pk_final;
    d.mode = kFinal; d.hi = 0;
    return;
  }
  const int l = (ph - 3) / ppl;
  int k = (ph - 3) % ppl;
  // canonical phase ids: 0 qkv, 1 self, 2 comb, 3 oproj, 4 cq, 5 cross, 6 comb, 7 coproj, 8 ffn1, 9 ffn2, 10 reduce
  int id;
  if (comb) id = k;
  else id = k < 2 ? k : (k < 5 ? k + 1 : k + 2);
  const TtsDecLayerWeights& lw = a.w.layer[l];
  const size_t self_off = (size_t)l * B * H * T * DH, cross_off = (size_t)l * B * H * S * DH;
  float* al_self = a.st.align_self ? a.st.align_self + (size_t)l * B * H * T * T : nullptr;
  float* al_cross = a.st.align_cross ? a.st.align_cross + (size_t)l * B * H * T * S : nullptr;
  const uint32_t site0 = 16u + 8u * (uint32_t)l;   // dropout sites of layer l: +0 self weights, +1 self out, +2 cross weights,
                                                   // +3 cross out, +4 FFN hidden, +5 FFN out
  const bool drop_here = id == 1 || id == 3 || id == 5 || id == 7 || id == 8 || (id == 9 && !red) || id == 10;
  if (drop_here) {
    d.drop_thr = a.thr_t; d.drop_sc = a.sc_t;
    d.drop_site = site0 + (id == 1 ? 0u : id == 3 ? 1u : id == 5 ? 2u : id == 7 ? 3u : id == 8 ? 4u : 5u);
  }
  switch (id) {
    case 0:  // LN + QKV (attention.py:63-64), k/v appended at row t
      d.ln = 1; d.X = a.x; d.ldx = D + kXPad; d.K = D; d.N = 3 * D; d.W = lw.pk_qkv;
      d.mode = kQkv; d.Y = a.q; d.ldy = D; d.out_scale = qscale;
      d.kcache = a.st.self_k + self_off; d.vcache = a.st.self_v + self_off; d.hi = 0;
      break;
    case 1:
      d.kind = kAttn;
      d.kc = a.st.self_k + self_off; d.vc = a.st.self_v + self_off; d.rows_alloc = T; d.n_keys = t + 1;
      d.key_len = nullptr; d.align = al_self; d.align_bh_stride = (long long)T * T; d.align_row_len = T;
      break;
    case 2:
      d.kind = kCombine; d.n_keys = t + 1; d.align = al_self; d.align_bh_stride = (long long)T * T; d.align_row_len = T;
      break;
    case 3:  // output projection + residual (attention.py:118-119, modules.py:132)
      d.X = a.ctx; d.ldx = D + kXPad; d.K = D; d.N = D; d.W = lw.pk_self_out; d.mode = kPlain;
      d.Y = a.x; d.ldy = D + kXPad; d.R = a.x; d.ldr = D + kXPad; d.hi = 1;
      break;
    case 4:  // LN + cross query
      d.ln = 1; d.X = a.x; d.ldx = D + kXPad; d.K = D; d.N = D; d.W = lw.pk_cross_q;
      d.mode = kPlain; d.Y = a.q; d.ldy = D; d.out_scale = qscale;
      d.hi = 0;
      break;
    case 5:
      d.kind = kAttn;
      d.kc = a.st.cross_k + cross_off; d.vc = a.st.cross_v + cross_off; d.rows_alloc = S; d.n_keys = S;
      d.key_len = a.st.input_lengths; d.align = al_cross; d.align_bh_stride = (long long)T * S; d.align_row_len = S;
      break;
    case 6:
      d.kind = kCombine; d.n_keys = S; d.align = al_cross; d.align_bh_stride = (long long)T * S; d.align_row_len = S;
      break;
    case 7:
      d.X = a.ctx; d.ldx = D + kXPad; d.K = D; d.N = D; d.W = lw.pk_cross_out; d.mode = kPlain;
      d.Y = a.x; d.ldy = D + kXPad; d.R = a.x; d.ldr = D + kXPad; d.hi = 1;
      break;
    case 8:  // LN + FFN-in + ReLU (modules.py:14-17)
      d.ln = 1; d.X = a.x; d.ldx = D + kXPad; d.K = D; d.N = F; d.W = lw.pk_ffn_in;
      d.relu = 1; d.mode = kHid; d.Y = a.hid; d.ldy = F / a.ksplit + kXPad; d.hi = 0;
      break;
    case 9:  // FFN-out: K-split partials, or + residual directly
      d.X = a.hid; d.ldx = F / a.ksplit + kXPad; d.K = F; d.N = D; d.W = lw.pk_ffn_out; d.hi = 1;
      if (red) {
        d.ksplit = a.ksplit; d.mode = kPartial; d.Y = a.part; d.ldy = D;
      } else {
        d.mode = kPlain; d.Y = a.x; d.ldy = D + kXPad; d.R = a.x; d.ldr = D + kXPad;
      }
      break;
    default:
      d.kind = kReduce; d.Y = a.x; d.ldy = D + kXPad; d.N = D; d.part = a.part; d.n_parts = a.ksplit;
      break;
  }
}

// ---- the kernel ------------------------------------------------------------------------------------------------
// DROP = false compiles the Philox paths out (the deterministic build measured ~2 % faster than the one that only
// tests the thresholds at run time); DROP = true is decoder.train() at synthesis time.
template <int DH, bool DROP>
__global__ void __launch_bounds__(kThreads, 1) pipelined_decode_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) float smem_raw[];
  const Smem sm = make_smem(a, smem_raw);
  const int B = a.st.batch, T = a.st.t_max, G = gridDim.x, NG = a.n_groups;
  const int n_ph = n_phases(a);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float qscale = (float)(1.0 / sqrt((double)DH));

  if (tid == 0) {
    mbar_init(sm.x_full, 1);
    mbar_init(sm.x_empty, kCWarps);
    mbar_init(&sm.w_full[0], 1);
    mbar_init(&sm.w_full[1], 1);
    *reinterpret_cast<volatile unsigned*>(sm.pdone) = 0u;
    mbar_init(&sm.qfull[0], 1);
    mbar_init(&sm.qfull[1], 1);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&sm.rfull[i], 1);
      sm.drained[i] = 0u;
    }
    *sm.kv_go = 0u;
    *sm.sig = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int b = tid; b < B; b += kThreads) {
    sm.len[b] = a.st.lengths[b];
    sm.fin[b] = a.st.finished[b];
    sm.klen[b] = a.st.input_lengths ? a.st.input_lengths[b] : a.st.mem_len;
  }
  __syncthreads();
  const int t0 = *a.st.step_counter;
  // never step past the session's capacity (K/V rows, frames and the PE table end at t_max): run what fits and
  // report -2 through n_unfinished instead of writing out of bounds
  const int n_run = min(a.n_steps, T - t0);
  if (n_run <= 0) {
    if (blockIdx.x == 0 && tid == 0) *a.st.n_unfinished = -2;
    return;
  }

  CState cs{0u, 0u, 0u, 0u, 0u};
  unsigned p_gp = 0u, p_go = 0u;        // loader: group-phases staged, attention group-phases released to the feeder
  unsigned f_go = 0u, f_seq = 0u, f_q = 0u;   // feeder: attention group-phases served, ring tiles issued, queries prefetched
  unsigned s_gp = 0u;                   // signaler: group-phases published
  unsigned epoch = 0u;                  // phases completed per group since the kernel started

  for (int s = 0; s < n_run; ++s) {
    const int t = t0 + s;
    // Steps after the first of a launch skip table phase 0 (the previous step's final phase already produced p0 from
    // the frame it emitted).  The FIRST step of every launch runs it from st->frames[:, t-1]: that is the ABI
    // contract (a caller may have written, primed or clamped that frame between launches) and it makes the scratch
    // content irrelevant across launches and implementations.
    const int off = s > 0 ? 1 : 0;
    const int n_ph_s = n_ph - off;          // phases of this step; `ph` below counts them, table index = ph + off
    if (warp == kCWarps) {
      // =========================== producer warp (one elected lane) ===========================
      if (lane == 0) {
        fence_proxy_async();  // frames / activations written by other CTAs in the previous step are read by TMA
        get_phase<DH>(a, off, t, qscale, sm.desc[0]);
        if (sm.desc[0].kind == kGemm) issue_weights(sm, sm.desc[0]);
        for (int ph = 0; ph < n_ph_s; ++ph) {
          const Desc& d = sm.desc[ph % kDescRing];
          for (int g = 0; g < NG; ++g) {
            if (ph > 0) {
              grid_wait(a, g, (epoch + ph) * G);      // (ph-1, g) is complete everywhere
              fence_proxy_async();
            }
            if (d.kind == kAttn)   // q and the appended K/V row of this group are visible: the feeder may start
              asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(sm.kv_go)), "r"(++p_go) : "memory");
            mbar_wait(sm.x_empty, (p_gp & 1u) ^ 1u, a.err);
            stage_tile(a, sm, d, g);
            ++p_gp;
            if (g == 0 && ph + 1 < n_ph_s) {
              Desc& dn = sm.desc[(ph + 1) % kDescRing];
              get_phase<DH>(a, ph + 1 + off, t, qscale, dn);
              if (ph >= 1) wait_count(a, sm.pdone, epoch + ph);   // the consumers are done with phase ph-1: its weight
                                                                  // half (and the K/V rings) may be overwritten
              if (dn.kind == kGemm) issue_weights(sm, dn);
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == kCWarps + 2) {
      // =========================== K/V feeder warp (one lane per ring slot) ===========================
      fence_proxy_async();   // K/V rows appended in earlier steps (observed through the step-end barrier) are read by TMA
      for (int ph = 0; ph < n_ph_s; ++ph) {
        int l;
        const int id = phase_id(a, ph + off, l);
        if (id != 1 && id != 5) continue;
        const size_t BH = (size_t)B * a.w.n_heads;
        const size_t off = id == 1 ? (size_t)l * BH * T * DH : (size_t)l * BH * a.st.mem_len * DH;
        const float* kc = (id == 1 ? a.st.self_k : a.st.cross_k) + off;
        const float* vc = (id == 1 ? a.st.self_v : a.st.cross_v) + off;
        const int rows_alloc = id == 1 ? T : a.st.mem_len, n_keys = id == 1 ? t + 1 : a.st.mem_len;
        // the ring sits above the weight tiles of the GEMM in front of an attention phase; FFN weights are larger
        // and reach into it, so self-attention of layer l > 0 may be pre-filled once FFN-out of layer l-1 is done
        if (id == 1 && l > 0) wait_count(a, sm.pdone, epoch + (unsigned)ph - 1u - (a.ksplit > 1 ? 1u : 0u));
        for (int g = 0; g < NG; ++g) {
          ++f_go;
          feed_group<DH>(a, sm, kc, vc, rows_alloc, n_keys, g, f_seq, f_go, id == 1, epoch + (unsigned)ph, f_q);
        }
      }
      __syncwarp();
    } else if (warp == kCWarps + 1) {
      // =========================== signaler warp (one elected lane) ===========================
      // publishes the consumers' finished group-phases to the other CTAs: the gpu-scope release (a memory barrier
      // that waits for the CTA's outstanding stores) runs here, off the consumers' critical path
      if (lane == 0) {
        for (int ph = 0; ph < n_ph_s; ++ph)
          for (int g = 0; g < NG; ++g) {
            ++s_gp;
            wait_count(a, sm.sig, s_gp);
            grid_arrive(a, g);
            if (g == 0 && a.prefetch) {   // L2 prefetch of the K/V streams of the attention phase two GEMMs ahead
              int l;
              const int id = phase_id(a, ph + off, l);
              const size_t BH = (size_t)B * a.w.n_heads;
              if (id == 3) {
                const size_t off = (size_t)l * BH * a.st.mem_len * DH;
                prefetch_kv<DH>(a, a.st.cross_k + off, a.st.cross_v + off, a.st.mem_len, a.st.mem_len);
              } else if (id == 7 && l + 1 < a.w.n_layers) {
                const size_t off = (size_t)(l + 1) * BH * T * DH;
                prefetch_kv<DH>(a, a.st.self_k + off, a.st.self_v + off, T, t);
              } else if (id == 100) {
                prefetch_kv<DH>(a, a.st.self_k, a.st.self_v, T, t);
              }
            }
          }
      }
      __syncwarp();
    } else {
      // =========================== consumer warps ===========================
      for (int ph = 0; ph < n_ph_s; ++ph) {
        long long* prof = (blockIdx.x == 0 && tid == 0 && ph + off < kProfPhases) ? a.prof + kProfStride * (ph + off) : nullptr;
        for (int g = 0; g < NG; ++g) {
          if (prof && g < 2) prof[3 * g] = clock64();
          mbar_wait(sm.x_full, cs.gp & 1u, a.err);
          ++cs.gp;
          if (prof && g < 2) prof[3 * g + 1] = clock64();
          const Desc& d = sm.desc[ph % kDescRing];
          if (d.kind == kGemm) {
            gemm_group<DH, DROP>(a, d, sm, cs, g, t, g == 0, g == 0 ? prof : nullptr);
          } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.x_empty);
            if (prof && g == 0) prof[11] = clock64();
            if (d.kind == kAttn) attn_group<DH, DROP>(a, d, sm, cs, g, t, g == 0 ? prof : nullptr);
            else if (d.kind == kReduce) reduce_group<DROP>(a, d, g, t);
            else combine_group<DH>(a, d, sm, g, t);
          }
          consumer_bar();
          if (tid == 0)   // every consumer's stores of this group-phase happen-before this (bar.sync): hand over
            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(sm.sig)), "r"(cs.gp) : "memory");
          if (prof && g < 2) prof[3 * g + 2] = clock64();
        }
        if (tid == 0)
          asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(sm.pdone)), "r"(epoch + ph + 1u) : "memory");
      }
      if (tid == 0)
        for (int g = 0; g < NG; ++g) grid_wait(a, g, (epoch + n_ph_s) * G);   // the step is complete everywhere
    }
    epoch += n_ph_s;
    __syncthreads();
    // synthesize.py:42-45, replicated identically in every CTA
    if (a.update_state) {
      for (int b = tid; b < B; b += kThreads) {
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
        for (int b = tid; b < B; b += kThreads) {
          a.st.lengths[b] = sm.len[b];
          a.st.finished[b] = (uint8_t)sm.fin[b];
        }
      if (tid == 0) {
        *a.st.step_counter = t + 1;
        *a.st.n_unfinished = (*reinterpret_cast<volatile int*>(a.err) != 0) ? -1
                             : ((s + 1 == n_run && n_run < a.n_steps) ? -2 : unfinished);
      }
    }
    if ((a.update_state && unfinished == 0) || s + 1 == n_run) break;  // uniform across CTAs
    if (*reinterpret_cast<volatile int*>(a.err) != 0) break;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct Carve {
  float *x, *q, *ctx, *hid, *p0, *p1, *part, *fpart;
  unsigned* bar;
  int* err;
  long long* prof;
  size_t floats;
};

static int num_sms() {   // per device: a process may drive several GPUs
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

static int ksplit_for(const TtsDecoderWeights* w) { return (w->d_ffn + kKC - 1) / kKC; }

// The K/V ring lives in the weight region between the tiles of the GEMM before an attention phase (QKV or the
// cross query: low placement) and the tiles of the GEMM after it (an output projection: high placement), so the
// feeder may start while the consumers still read the former and the loader already streams the latter.
static int ring_slots(const TtsDecoderWeights* w, int G, int* ring_lo, int* n_hi) {
  const int D = w->d_model, dh = D / w->n_heads;
  auto tiles = [&](int N) { return ((N + G - 1) / G + 7) / 8; };
  const int lo = tiles(3 * D) * 8 * (D + kPad), hi = kWFloats - tiles(D) * 8 * (D + kPad);
  const int slot_f = 2 * kTK * dh;
  int nh = (hi - lo) / slot_f, nl = lo / slot_f;
  if (nh > kSlots) nh = kSlots;
  if (nh + nl > kSlots) nl = kSlots - nh;
  if (ring_lo) *ring_lo = lo;
  if (n_hi) *n_hi = nh;
  return nh + nl;
}

static int group_rows_for(int B) {
  const char* e = getenv("TTS_GROUP_ROWS");
  int r = e ? atoi(e) : 0;
  // up to 16 rows: one group (measured: two half-size groups cost more in doubled per-phase work and split-K/V
  // combine phases than the barrier latency they hide: 408 vs 311 us/step at B=16, 389 vs 312 at B=4)
  if (r <= 0) r = B > kGroupRows ? kGroupRows : B;
  if (r > kGroupRows) r = kGroupRows;
  if (r < 1) r = 1;
  while ((B + r - 1) / r > kMaxGroups) ++r;
  return r;
}

static int split_for(const TtsDecoderWeights* w, int group_rows, int G) {
  int ns = G / (group_rows * w->n_heads);
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
  c.bar = reinterpret_cast<unsigned*>(take(32 * kMaxGroups));
  c.err = reinterpret_cast<int*>(take(32));
  c.prof = reinterpret_cast<long long*>(take(2 * kProfStride * kProfPhases));
  // activations that are staged by TMA are stored with the shared-memory row stride (K + kXPad): a row group is one
  // contiguous run = one bulk copy.  hid is K-split-major: [ksplit][B][F / ksplit + kXPad].
  const size_t ks = ksplit_for(w);
  c.x = take(B * (D + kXPad)); c.q = take(B * D); c.ctx = take(B * (D + kXPad)); c.hid = take(ks * B * (F / ks + kXPad));
  c.p0 = take(B * (P + kXPad)); c.p1 = take(B * (P + kXPad));
  c.part = take((size_t)ksplit_for(w) * B * D);
  c.fpart = take((size_t)B * H * kMaxSplit * (dh + 4));
  c.floats = off;
  return c;
}

template <int DH, bool DROP>
static int launch(const Args& a, cudaStream_t s) {
  static std::atomic<unsigned long long> configured{0ull};   // bit per device
  const size_t smem = smem_bytes(a.st.batch);
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    TTS_CHECK_CUDA(cudaFuncSetAttribute(pipelined_decode_kernel<DH, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 0;
    TTS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pipelined_decode_kernel<DH, DROP>, kThreads, smem));
    TTS_REQUIRE(per_sm >= 1, "pipelined decode kernel does not fit on an SM");
    configured.fetch_or(bit, std::memory_order_release);
  }
  Args args = a;
  void* params[] = {&args};
  TTS_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(pipelined_decode_kernel<DH, DROP>), dim3(num_sms()),
                                             dim3(kThreads), params, smem, s));
  count_launch();
  return 0;
}

}  // namespace pipe

int pipelined_profile(const TtsDecoderWeights* w, const TtsDecodeState* st, long long* out_host, int max_entries) {
  using namespace pipe;
  const Carve c = carve(w, st->batch, st->scratch);
  const int n = max_entries < kProfStride * kProfPhases ? max_entries : kProfStride * kProfPhases;
  TTS_CHECK_CUDA(cudaMemcpy(out_host, c.prof, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

size_t pipelined_scratch_floats(const TtsDecoderWeights* w, int B) { return pipe::carve(w, B, nullptr).floats; }

bool pipelined_supported(const TtsDecoderWeights* w, const TtsDecodeState* st) {
  using namespace pipe;
  const int G = num_sms();
  const int D = w->d_model, F = w->d_ffn, P = w->prenet_hidden, M = w->n_mels;
  if (D > kKC || D % 16 != 0 || P > kKC || P % 16 != 0 || M > kKC || M % 16 != 0 || F % 16 != 0) return false;
  if (!w->pk_pre0 || !w->pk_pre1 || !w->pk_pre2 || !w->pk_final) return false;   // packed operands required
  for (int l = 0; l < w->n_layers; ++l) {
    const TtsDecLayerWeights& lw = w->layer[l];
    if (!lw.pk_qkv || !lw.pk_self_out || !lw.pk_cross_q || !lw.pk_cross_out || !lw.pk_ffn_in || !lw.pk_ffn_out) return false;
  }
  const int ks = ksplit_for(w);
  if (w->pk_ksplit != ks) return false;
  if (F % ks != 0 || (F / ks) % 16 != 0 || F / ks > kKC || ks > G) return false;
  if (st->batch > kMaxBatch) return false;
  const int dh = D / w->n_heads;
  if (dh != 32 && dh != 64 && dh != 96) return false;
  auto rows = [&](long long N, int parts) { return (int)((N + parts - 1) / parts); };
  int worst = rows(3 * D, G);
  worst = worst > rows(F, G) ? worst : rows(F, G);
  worst = worst > rows(D, G / ks) ? worst : rows(D, G / ks);
  worst = worst > rows(P, G) ? worst : rows(P, G);
  worst = worst > rows(M + 1 + P, G) ? worst : rows(M + 1 + P, G);
  if (worst > kMaxRows) return false;
  if (ring_slots(w, G, nullptr, nullptr) < 2) return false;
  if (smem_bytes(st->batch) > 227 * 1024) return false;
  return true;
}

int launch_pipelined_steps(const TtsDecoderWeights* w, const TtsDecodeState* st, int n_steps, int update_state,
                           cudaStream_t s) {
  using namespace pipe;
  TTS_REQUIRE(pipelined_supported(w, st), "pipelined decode kernel does not support this shape");
  if (n_steps == 0) return 0;
  const Carve c = carve(w, st->batch, st->scratch);
  Args a;
  memcpy(&a.w, w, sizeof(*w));
  memcpy(&a.st, st, sizeof(*st));
  a.x = c.x; a.q = c.q; a.ctx = c.ctx; a.hid = c.hid; a.p0 = c.p0; a.p1 = c.p1; a.part = c.part; a.fpart = c.fpart;
  a.bar = c.bar; a.err = c.err; a.prof = c.prof; a.n_steps = n_steps; a.update_state = update_state;
  a.group_rows = group_rows_for(st->batch);
  a.n_groups = (st->batch + a.group_rows - 1) / a.group_rows;
  a.n_split = split_for(w, a.group_rows, num_sms());
  a.ksplit = ksplit_for(w);
  a.prefetch = getenv("TTS_PREFETCH") != nullptr;   // off by default: it competes with the weight stream (measured -3 %)
  a.thr_d = st->drop_p_prenet > 0.f ? drop_threshold(st->drop_p_prenet) : 0u;
  a.sc_d = st->drop_p_prenet > 0.f ? 1.f / (1.f - st->drop_p_prenet) : 1.f;
  a.thr_t = st->drop_p_transformer > 0.f ? drop_threshold(st->drop_p_transformer) : 0u;
  a.sc_t = st->drop_p_transformer > 0.f ? 1.f / (1.f - st->drop_p_transformer) : 1.f;
  a.seed = st->drop_seed;
  a.n_slots = ring_slots(w, num_sms(), &a.ring_lo, &a.n_hi);
  TTS_CHECK_CUDA(cudaMemsetAsync(c.bar, 0, (32 * kMaxGroups + 32) * sizeof(unsigned), s));  // counters + error flag
  const bool drop = a.thr_d != 0u || a.thr_t != 0u;
  switch (w->d_model / w->n_heads) {
    case 32: return drop ? launch<32, true>(a, s) : launch<32, false>(a, s);
    case 64: return drop ? launch<64, true>(a, s) : launch<64, false>(a, s);
    case 96: return drop ? launch<96, true>(a, s) : launch<96, false>(a, s);
  }
  set_error("pipelined decode: unsupported head_dim");
  return 2;
}

}  // namespace tts
