// Full-sequence multi-head attention (encoder self-attention, teacher-forced decoder causal
// self-attention and cross-attention): tiled, online-softmax, no [B,H,Tq,Tk] logits tensor in
// HBM.  The mask is computed from indices/lengths in-kernel instead of the additive bias tensors
// the reference builds on the CPU every call (transformer/common.py:32-48, modules.py:50-52,
// 109-112).  Replaces transformer/attention.py:72-91 (+ split/combine_heads :6-26, the q scale
// :113-114).  The attention map is only materialised when the caller asks for it.
#include <math_constants.h>

#include "common.cuh"

namespace tts {

constexpr int kQT = 32;  // queries per CTA (4 per warp)
constexpr int kKT = 64;  // keys per tile (2 per lane)

template <int DH>
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ q, int ldq,
                                                        const float* __restrict__ k, int ldk,
                                                        const float* __restrict__ v, int ldv,
                                                        float* __restrict__ ctx, float* __restrict__ align,
                                                        int H, int Tq, int Tk, float q_scale, int causal,
                                                        const int32_t* __restrict__ key_len) {
  constexpr int LD = DH + 4;   // padded smem row: conflict-free 128-bit reads at one row per lane
  constexpr int DPL = DH / 32; // output dims per lane
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                 // [kKT][LD]
  float* Vs = Ks + kKT * LD;        // [kKT][LD]
  float* Qs = Vs + kKT * LD;        // [kQT][DH]
  float* Ps = Qs + kQT * DH;        // [8][4][kKT]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQT;
  const int klen = key_len ? key_len[b] : Tk;

  // stage the (pre-scaled) queries of this CTA
  for (int i = tid; i < kQT * (DH / 4); i += 256) {
    const int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tq) val = *reinterpret_cast<const float4*>(q + ((size_t)b * Tq + q0 + r) * ldq + h * DH + c);
    val.x *= q_scale; val.y *= q_scale; val.z *= q_scale; val.w *= q_scale;
    *reinterpret_cast<float4*>(Qs + r * DH + c) = val;
  }

  float m_run[4], l_run[4], o[4][DPL];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F;
    l_run[i] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[i][d] = 0.f;
  }

  const int q_last = min(q0 + kQT, Tq) - 1;
  const int n_tiles_all = ceil_div(Tk, kKT);
  const int n_tiles = causal ? min(n_tiles_all, q_last / kKT + 1) : n_tiles_all;

  auto load_kv = [&](int kt, bool with_v) {
    for (int i = tid; i < kKT * (DH / 4); i += 256) {
      const int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      const int j = kt * kKT + r;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < Tk) {
        kv = *reinterpret_cast<const float4*>(k + ((size_t)b * Tk + j) * ldk + h * DH + c);
        if (with_v) vv = *reinterpret_cast<const float4*>(v + ((size_t)b * Tk + j) * ldv + h * DH + c);
      }
      *reinterpret_cast<float4*>(Ks + r * LD + c) = kv;
      if (with_v) *reinterpret_cast<float4*>(Vs + r * LD + c) = vv;
    }
  };

  // scores of the warp's 4 queries against keys (lane, lane+32) of the staged tile
  auto tile_scores = [&](int kt, float (&s)[4][2]) {
    f32x2 acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0ull;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const f32x4 k0 = lds128(Ks + lane * LD + c);
      const f32x4 k1 = lds128(Ks + (lane + 32) * LD + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x4 qv = lds128(Qs + (warp * 4 + i) * DH + c);
        acc[i][0] = fma2(qv.lo, k0.lo, acc[i][0]);
        acc[i][0] = fma2(qv.hi, k0.hi, acc[i][0]);
        acc[i][1] = fma2(qv.lo, k1.lo, acc[i][1]);
        acc[i][1] = fma2(qv.hi, k1.hi, acc[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + warp * 4 + i;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int j = kt * kKT + lane + 32 * kk;
        float val = hsum2(acc[i][kk]);
        if (j >= Tk) val = -CUDART_INF_F;                                  // no such key
        else if (j >= klen || (causal && j > qi)) val = kNegBias;          // logits + (-1e20)
        s[i][kk] = val;
      }
    }
  };

  for (int kt = 0; kt < n_tiles; ++kt) {
    __syncthreads();  // previous tile fully consumed (and Qs visible on the first pass)
    load_kv(kt, true);
    __syncthreads();
    float s[4][2];
    tile_scores(kt, s);
    float* Pw = Ps + warp * 4 * kKT;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float tmax = warp_max(fmaxf(s[i][0], s[i][1]));
      const float m_new = fmaxf(m_run[i], tmax);
      const float corr = expf(m_run[i] - m_new);
      const float p0 = expf(s[i][0] - m_new), p1 = expf(s[i][1] - m_new);
      l_run[i] = l_run[i] * corr + warp_sum(p0 + p1);
      m_run[i] = m_new;
#pragma unroll
      for (int d = 0; d < DPL; ++d) o[i][d] *= corr;
      Pw[i * kKT + lane] = p0;
      Pw[i * kKT + lane + 32] = p1;
    }
    __syncwarp();
#pragma unroll 4
    for (int j = 0; j < kKT; j += 4) {
      float vv[4][DPL];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int d = 0; d < DPL; ++d) vv[jj][d] = Vs[(j + jj) * LD + lane + 32 * d];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 p = *reinterpret_cast<const float4*>(Pw + i * kKT + j);
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
          o[i][d] = fmaf(p.x, vv[0][d], o[i][d]);
          o[i][d] = fmaf(p.y, vv[1][d], o[i][d]);
          o[i][d] = fmaf(p.z, vv[2][d], o[i][d]);
          o[i][d] = fmaf(p.w, vv[3][d], o[i][d]);
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + warp * 4 + i;
    if (qi < Tq) {
      const float inv = 1.f / l_run[i];
#pragma unroll
      for (int d = 0; d < DPL; ++d)
        ctx[((size_t)b * Tq + qi) * (H * DH) + h * DH + lane + 32 * d] = o[i][d] * inv;
    }
  }

  if (align != nullptr) {  // second sweep: recompute the logits, emit normalised weights
    for (int kt = 0; kt < n_tiles_all; ++kt) {
      float s[4][2];
      const bool live_tile = kt < n_tiles;
      __syncthreads();
      if (live_tile) load_kv(kt, false);
      __syncthreads();
      if (live_tile) tile_scores(kt, s);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q0 + warp * 4 + i;
        if (qi >= Tq) continue;
        float* row = align + (((size_t)b * H + h) * Tq + qi) * Tk;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int j = kt * kKT + lane + 32 * kk;
          if (j < Tk) row[j] = live_tile ? expf(s[i][kk] - m_run[i]) / l_run[i] : 0.f;
        }
      }
    }
  }
}

template <int DH>
static int launch_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                            float* ctx, float* align, int B, int H, int Tq, int Tk, float q_scale, int causal,
                            const int32_t* key_len, cudaStream_t s) {
  const size_t smem = (size_t)(2 * kKT * (DH + 4) + kQT * DH + 8 * 4 * kKT) * sizeof(float);
  TTS_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(Tq, kQT), H, B);
  attention_kernel<DH><<<grid, 256, smem, s>>>(q, ldq, k, ldk, v, ldv, ctx, align, H, Tq, Tk, q_scale, causal,
                                                key_len);
  TTS_CHECK_LAUNCH();
  return 0;
}

}  // namespace tts

using namespace tts;

extern "C" int tts_attention(const float* q, int32_t ldq, const float* k, int32_t ldk, const float* v,
                             int32_t ldv, float* ctx, float* align, int32_t batch, int32_t n_heads, int32_t tq,
                             int32_t tk, int32_t head_dim, float q_scale, int32_t causal,
                             const int32_t* key_len, void* stream) {
  TTS_REQUIRE(batch > 0 && n_heads > 0 && tq > 0 && tk > 0, "tts_attention: empty problem");
  TTS_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "tts_attention: row strides must be multiples of 4");
  TTS_REQUIRE(batch <= 65535 && n_heads <= 65535, "tts_attention: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 32: return launch_attention<32>(q, ldq, k, ldk, v, ldv, ctx, align, batch, n_heads, tq, tk, q_scale, causal, key_len, s);
    case 64: return launch_attention<64>(q, ldq, k, ldk, v, ldv, ctx, align, batch, n_heads, tq, tk, q_scale, causal, key_len, s);
    case 96: return launch_attention<96>(q, ldq, k, ldk, v, ldv, ctx, align, batch, n_heads, tq, tk, q_scale, causal, key_len, s);
    default: set_error("tts_attention: head_dim %d not in {32,64,96}", head_dim); return 2;
  }
}
