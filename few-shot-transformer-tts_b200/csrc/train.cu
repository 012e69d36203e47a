// Memory-bound kernels of the teacher-forced TRAINING path (everything around the bf16 GEMMs and the attention):
// LayerNorm forward / backward, dropout-backward + bf16 casts, encoder / decoder prologues and their backward
// (embedding scatter, pe_scale), batch-statistics BatchNorm + tanh + dropout of the Postnet, bias / stop-net
// reductions, the fused masked loss with its gradients, the L2 term and a fused multi-tensor Adam.
// Every kernel is a single pass over its operands with 128-bit accesses; dropout masks are regenerated from
// (seed, stream, element) instead of being stored (philox.cuh).
// Reference semantics: transformer/modules.py:49-69,108-145; tacotron.py:33-44,55-65,81-90,136-158; train.py:130,188.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <math_constants.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"

namespace tts {
namespace tr {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float block_sum(float v, float* sh) {   // blockDim.x <= 1024, result in every thread
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = l < nw ? sh[l] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ bool keep1(unsigned long long seed, uint32_t stream, unsigned long long e, uint32_t thresh) {
  const uint4 w = dropout_words_linear(seed, stream, e >> 3);
  return dropout_lane16(w, (int)(e & 7)) >= thresh;
}
struct Drop {
  uint32_t thresh; float scale; unsigned long long seed; uint32_t stream;
};
static Drop make_drop(float p, unsigned long long seed, uint32_t stream) {
  Drop d;
  d.thresh = p > 0.f ? drop_threshold16(p) : 0u;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.seed = seed;
  d.stream = stream;
  return d;
}
// 4 consecutive elements starting at e (e % 4 == 0)
__device__ __forceinline__ void drop4(const Drop& d, unsigned long long e, float4& v) {
  if (d.thresh == 0u) return;
  const uint4 w = dropout_words_linear(d.seed, d.stream, e >> 3);   // 16-bit lanes: this half of the call's eight
  const uint32_t a = (e & 4) ? w.z : w.x, b = (e & 4) ? w.w : w.y;
  v.x = (a & 0xffffu) >= d.thresh ? v.x * d.scale : 0.f;
  v.y = (a >> 16) >= d.thresh ? v.y * d.scale : 0.f;
  v.z = (b & 0xffffu) >= d.thresh ? v.z * d.scale : 0.f;
  v.w = (b >> 16) >= d.thresh ? v.w * d.scale : 0.f;
}
__device__ __forceinline__ uint2 pack4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}
__device__ __forceinline__ float4 unpack4(uint2 u) {
  const __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, C % 4 == 0, C <= 1024.  y (bf16) feeds the next GEMM; mean / rstd are saved.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLnMaxV = 8;   // float4 per lane: C <= 1024
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long ldy,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int rows, int C,
                                                     float eps, const int32_t* row_len, int rpb) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
  float4 v[kLnMaxV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < nv ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  const bool dead = row_len != nullptr && (row % rpb) >= row_len[row / rpb];
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * ldy);
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o = make_float4((v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y, (v[i].z - mu) * rs * g.z + b.z,
                             (v[i].w - mu) * rs * g.w + b.w);
      if (dead) o = make_float4(0.f, 0.f, 0.f, 0.f);
      yr[c] = pack4(o);
    }
  }
}

// LayerNorm backward.  dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy gamma; dx (+= dres) in fp32;
// per-CTA partial sums of dgamma = sum dy xhat and dbeta = sum dy go to part[blockIdx][2][C] (finalised below).
// One warp per row.  The column partials live in SHARED memory (one private [2][C] slice per warp), not in registers, and
// xhat / g are recomputed for the output pass instead of being kept: 3 CTAs (24 warps) per SM instead of the 1 CTA that
// the 193-register version got (ncu: 12.5 % occupancy, 1.4-2.6 TB/s).  Optionally also writes dyb = dropout_backward(dx) in
// bf16 (the dY operand of the preceding Linear's wgrad / dgrad), which saves the separate cast kernel's re-read of dx.
template <int NV, bool EARLY>
__global__ void __launch_bounds__(256, 3) ln_bwd_kernel(const bf16* __restrict__ dy, long long lddy, const float* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ dres,
                                                        float* __restrict__ dx, bf16* __restrict__ dyb, long long lddyb, Drop drop,
                                                        float* __restrict__ part, int rows, int C, const int32_t* row_len, int rpb) {
  extern __shared__ float red[];   // [8 warps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nv = C >> 2;
  float4* wa = reinterpret_cast<float4*>(red + (size_t)(warp * 2) * C);
  float4* wb = reinterpret_cast<float4*>(red + (size_t)(warp * 2 + 1) * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) wa[c] = wb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  const float inv_c = 1.f / (float)C;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    const bool dead = row_len != nullptr && (row % rpb) >= row_len[row / rpb];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + (size_t)row * lddy);
    const float4* rr = dres ? reinterpret_cast<const float4*>(dres + (size_t)row * C) : nullptr;
    float4 xv[NV], rv[NV];
    uint2 dv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {   // every operand of the row is requested up front
      const int c = lane + 32 * i;
      const bool in = c < nv;
      xv[i] = in ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      dv[i] = (in && !dead) ? dr[c] : make_uint2(0u, 0u);
      if (EARLY) rv[i] = (in && rr != nullptr) ? rr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 d = unpack4(dv[i]);
        const float4 xh = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
        float4 a = wa[c], bsum = wb[c];
        a.x += d.x * xh.x; a.y += d.y * xh.y; a.z += d.z * xh.z; a.w += d.w * xh.w;
        bsum.x += d.x; bsum.y += d.y; bsum.z += d.z; bsum.w += d.w;
        wa[c] = a; wb[c] = bsum;
        const float4 gg = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += (gg.x + gg.y) + (gg.z + gg.w);
        s2 += gg.x * xh.x + gg.y * xh.y + gg.z * xh.z + gg.w * xh.w;
      }
    }
    if (!EARLY) {   // the residual-path gradient is requested once the first pass' temporaries are dead (no spills at 80 registers)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        rv[i] = (c < nv && rr != nullptr) ? rr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
    float4* ox = reinterpret_cast<float4*>(dx + (size_t)row * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 d = unpack4(dv[i]);
        const float4 xh = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
        float4 o = make_float4(rs * (d.x * g.x - s1 - xh.x * s2) + rv[i].x, rs * (d.y * g.y - s1 - xh.y * s2) + rv[i].y,
                               rs * (d.z * g.z - s1 - xh.z * s2) + rv[i].z, rs * (d.w * g.w - s1 - xh.w * s2) + rv[i].w);
        ox[c] = o;
        if (dyb != nullptr) {
          drop4(drop, (unsigned long long)row * C + 4 * c, o);
          *reinterpret_cast<uint2*>(dyb + (size_t)row * lddyb + 4 * c) = pack4(o);
        }
      }
    }
  }
  // cross-warp reduction of the column partials, one [2][C] record per CTA
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[(size_t)(w * 2) * C + idx];   // [w][which][c] flattened: which * C + c == idx
    part[(size_t)blockIdx.x * 2 * C + idx] = t;
  }
}
// out[idx] = sum_b part[b][idx]: CTA = 32 columns x 8 block-groups (coalesced 128-byte reads, independent loads in flight)
__global__ void __launch_bounds__(256) colpart_finalize_kernel(const float* __restrict__ part, int n_blocks, int width,
                                                                float* __restrict__ out0, float* __restrict__ out1, int C) {
  __shared__ float sh[8][33];
  const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5, idx = blockIdx.x * 32 + cg;
  float t0 = 0.f, t1 = 0.f;
  if (idx < width) {
    int b = rg;
    for (; b + 8 < n_blocks; b += 16) {
      t0 += part[(size_t)b * width + idx];
      t1 += part[(size_t)(b + 8) * width + idx];
    }
    if (b < n_blocks) t0 += part[(size_t)b * width + idx];
  }
  sh[rg][cg] = t0 + t1;
  __syncthreads();
  if (rg == 0 && idx < width) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sh[r][cg];
    if (idx < C) out0[idx] = t;
    else if (out1) out1[idx - C] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dst (bf16) = dropout_backward(src fp32) = keep ? src * scale : 0, optional row mask; p = 0: plain cast
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) drop_cast_kernel(const float* __restrict__ src, long long lds, bf16* __restrict__ dst,
                                                        long long ldd, long long rows, int C, Drop d, const int32_t* row_len,
                                                        int rpb) {
  const int nv = C >> 2;
  const long long total = rows * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i - r * nv);
    float4 v = *reinterpret_cast<const float4*>(src + r * lds + 4 * c);
    drop4(d, (unsigned long long)r * C + 4 * c, v);
    if (row_len != nullptr && (r % rpb) >= row_len[r / rpb]) v = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<uint2*>(dst + r * ldd + 4 * c) = pack4(v);
  }
}

// table-driven cast of many fp32 tensors to bf16 (the per-step bf16 copies of the fp32 master weights)
struct CastEntry {
  const float* src; bf16* dst; long long n; long long first_chunk;
};
constexpr int kChunk = 8192;
__global__ void __launch_bounds__(256) multi_cast_kernel(const CastEntry* __restrict__ tab, int n_entries) {
  const long long chunk = blockIdx.x;
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {   // last entry whose first_chunk <= chunk
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].first_chunk <= chunk) lo = mid;
    else hi = mid - 1;
  }
  const CastEntry e = tab[lo];
  const long long base = (chunk - e.first_chunk) * kChunk, end = min(e.n, base + kChunk);
  if ((((uintptr_t)e.src | (uintptr_t)e.dst) & 15) == 0) {
    for (long long i = base + 4 * threadIdx.x; i + 3 < end; i += 4 * blockDim.x)
      *reinterpret_cast<uint2*>(e.dst + i) = pack4(*reinterpret_cast<const float4*>(e.src + i));
    for (long long i = (end & ~3ll) + threadIdx.x; i < end; i += blockDim.x)
      if (i >= base) e.dst[i] = __float2bfloat16_rn(e.src[i]);
  } else {
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) e.dst[i] = __float2bfloat16_rn(e.src[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// prologues
// ---------------------------------------------------------------------------------------------------------------------
// encoder: out[b,s,:] = dropout(embed[ids[b,s]] * (s < len[b]) + pe[s] * pe_scale)    (tacotron.py:34, modules.py:49-55)
__global__ void __launch_bounds__(128) embed_fwd_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ len,
                                                        const float* __restrict__ embed, const float* __restrict__ pe,
                                                        const float* __restrict__ pe_scale, float* __restrict__ out, int B, int S,
                                                        int C, int vocab, Drop d) {
  const int row = blockIdx.x, b = row / S, s = row - b * S;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const bool on = s < len[b];
  const float sc = *pe_scale;
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) {
    float4 v = on ? *reinterpret_cast<const float4*>(embed + (size_t)id * C + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s * C + 4 * c);
    v.x += p.x * sc; v.y += p.y * sc; v.z += p.z * sc; v.w += p.w * sc;
    drop4(d, (unsigned long long)row * C + 4 * c, v);
    *reinterpret_cast<float4*>(out + (size_t)row * C + 4 * c) = v;
  }
}
// backward: g = dropout_bwd(dx); d_embed[ids] += g * mask (fp32 atomics: 6000 x 512 table, rows repeat); d_pe_scale += sum g pe
__global__ void __launch_bounds__(128) embed_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids,
                                                        const int32_t* __restrict__ len, const float* __restrict__ pe,
                                                        float* __restrict__ d_embed, float* __restrict__ d_pe_scale, int B, int S, int C,
                                                        int vocab, Drop d) {
  __shared__ float sh[32];
  const int row = blockIdx.x, b = row / S, s = row - b * S;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const bool on = s < len[b];
  float acc = 0.f;
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) {
    float4 g = *reinterpret_cast<const float4*>(dx + (size_t)row * C + 4 * c);
    drop4(d, (unsigned long long)row * C + 4 * c, g);
    const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s * C + 4 * c);
    acc += g.x * p.x + g.y * p.y + g.z * p.z + g.w * p.w;
    if (on) {
      float* e = d_embed + (size_t)id * C + 4 * c;
      atomicAdd(e, g.x); atomicAdd(e + 1, g.y); atomicAdd(e + 2, g.z); atomicAdd(e + 3, g.w);
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(d_pe_scale, acc);
}
// decoder: out[b,t,:] = dropout((t ? pre[b,t-1,:] * (t-1 < len[b]) : 0) + pe[t] * pe_scale)      (modules.py:114-120)
__global__ void __launch_bounds__(192) shift_fwd_kernel(const float* __restrict__ pre, const int32_t* __restrict__ len,
                                                        const float* __restrict__ pe, const float* __restrict__ pe_scale,
                                                        float* __restrict__ out, int B, int T, int C, Drop d) {
  const int row = blockIdx.x, b = row / T, t = row - b * T;
  const bool have = t > 0 && (t - 1) < len[b];
  const float sc = *pe_scale;
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) {
    float4 v = have ? *reinterpret_cast<const float4*>(pre + (size_t)(row - 1) * C + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)t * C + 4 * c);
    v.x += p.x * sc; v.y += p.y * sc; v.z += p.z * sc; v.w += p.w * sc;
    drop4(d, (unsigned long long)row * C + 4 * c, v);
    *reinterpret_cast<float4*>(out + (size_t)row * C + 4 * c) = v;
  }
}
// backward: dpre[b,t] = dropout_bwd(dx)[b,t+1] * (t < len[b]) for t < T-1, 0 for t = T-1 (bf16: operand of the prenet
// dgrad / wgrad); d_pe_scale += sum dropout_bwd(dx) pe
__global__ void __launch_bounds__(192) shift_bwd_kernel(const float* __restrict__ dx, const int32_t* __restrict__ len,
                                                        const float* __restrict__ pe, bf16* __restrict__ dpre,
                                                        float* __restrict__ d_pe_scale, int B, int T, int C, Drop d) {
  __shared__ float sh[32];
  const int row = blockIdx.x, b = row / T, t = row - b * T;   // row of dx
  float acc = 0.f;
  const bool to_pre = t > 0 && (t - 1) < len[b];
  for (int c = threadIdx.x; c < (C >> 2); c += blockDim.x) {
    float4 g = *reinterpret_cast<const float4*>(dx + (size_t)row * C + 4 * c);
    drop4(d, (unsigned long long)row * C + 4 * c, g);
    const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)t * C + 4 * c);
    acc += g.x * p.x + g.y * p.y + g.z * p.z + g.w * p.w;
    if (t > 0) *reinterpret_cast<uint2*>(dpre + (size_t)(row - 1) * C + 4 * c) = pack4(to_pre ? g : make_float4(0.f, 0.f, 0.f, 0.f));
    if (t == T - 1) *reinterpret_cast<uint2*>(dpre + (size_t)row * C + 4 * c) = make_uint2(0u, 0u);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(d_pe_scale, acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// reductions over rows: out[c] += sum_r w[r] * x[r][c]  (bias gradients, stop-net weight gradient)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                     float* __restrict__ out, long long rows, int C) {
  // thread = column pair; CTA strides over row chunks
  const int nc2 = (C + 1) >> 1;
  for (int c2 = threadIdx.x; c2 < nc2; c2 += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    const bool pair = 2 * c2 + 1 < C && (ldx & 1) == 0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
      const float wr = w ? w[r] : 1.f;
      if (pair) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * ldx + 2 * c2));
        a0 += wr * v.x;
        a1 += wr * v.y;
      } else {
        a0 += wr * __bfloat162float(x[r * ldx + 2 * c2]);
        if (2 * c2 + 1 < C) a1 += wr * __bfloat162float(x[r * ldx + 2 * c2 + 1]);
      }
    }
    atomicAdd(out + 2 * c2, a0);
    if (2 * c2 + 1 < C) atomicAdd(out + 2 * c2 + 1, a1);
  }
}
__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += x[i];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) atomicAdd(out, a);
}
// stop head: out[r] = (sum_k x[r][k] w[k] + bias) * (r % rpb < len)      (tacotron.py:114-115; one warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const int32_t* row_len, int rpb,
                                                     float* __restrict__ out, long long rows, int K) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float a = 0.f;
  for (int k = 4 * lane; k < K; k += 128) {
    const float4 xv = unpack4(*reinterpret_cast<const uint2*>(x + row * ldx + k));
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
    a += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
  }
  a = warp_sum(a);
  if (lane == 0) {
    const bool dead = row_len != nullptr && (row % rpb) >= row_len[row / rpb];
    out[row] = dead ? 0.f : a + (bias ? *bias : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Postnet BatchNorm1d with batch statistics over ALL B x T positions (tacotron.py:85-86; SURVEY hard part 9)
// ---------------------------------------------------------------------------------------------------------------------
// pass 1: per-CTA column sums of z and z^2 -> part[blk][2][C]
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, long long rows, int C, float* __restrict__ part) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
      const float v = z[r * C + c];
      s += v;
      q += v * v;
    }
    part[(size_t)blockIdx.x * 2 * C + c] = s;
    part[(size_t)blockIdx.x * 2 * C + C + c] = q;
  }
}
// finalize in fp64: mean, invstd (biased variance, eps 1e-5) + running statistics (momentum 0.1, unbiased variance).
// One CTA per 32 channels, 8 warps: warp w sums partial blocks w, w + 8, ... of its lane's channel (coalesced 128-byte
// reads), the eight sums meet in shared memory (a single thread per channel walked all 592 partials: 93 us per layer).
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ part, int n_blocks, long long rows, int C, float eps,
                                                          float* __restrict__ mean, float* __restrict__ invstd, float* running_mean,
                                                          float* running_var, long long* num_batches, float momentum) {
  __shared__ double sh[2][8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int b = warp; b < n_blocks; b += 8) {
      s += (double)part[(size_t)b * 2 * C + c];
      q += (double)part[(size_t)b * 2 * C + C + c];
    }
  }
  sh[0][warp][lane] = s;
  sh[1][warp][lane] = q;
  __syncthreads();
  if (warp != 0 || c >= C) return;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {   // fixed order: reproducible
    s += sh[0][w][lane];
    q += sh[1][w][lane];
  }
  const double mu = s / (double)rows;
  double var = q / (double)rows - mu * mu;
  var = var < 0.0 ? 0.0 : var;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    if (c == 0 && num_batches != nullptr) *num_batches += 1;
  }
}
// pass 2: a = dropout(tanh?(gamma (z - mean) invstd + beta)); hidden layers store a * (t < len) as the next layer's
// zero-padded bf16 input [B][T+4][C]; the last layer stores a (+ residual) in fp32 [B][T][C], unmasked.
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int act_tanh, Drop d,
                                                       const int32_t* __restrict__ len, int B, int T, int C, bf16* out_pad,
                                                       float* out_f32, const float* residual) {
  const int nv = C >> 2;
  const long long total = (long long)B * T * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i - r * nv), b = (int)(r / T), t = (int)(r - (long long)b * T);
    const float4 zv = *reinterpret_cast<const float4*>(z + r * C + 4 * c);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c), is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), be = __ldg(reinterpret_cast<const float4*>(beta) + c);
    float4 y = make_float4((zv.x - mu.x) * is.x * g.x + be.x, (zv.y - mu.y) * is.y * g.y + be.y, (zv.z - mu.z) * is.z * g.z + be.z,
                           (zv.w - mu.w) * is.w * g.w + be.w);
    if (act_tanh) y = make_float4(tanhf(y.x), tanhf(y.y), tanhf(y.z), tanhf(y.w));
    drop4(d, (unsigned long long)r * C + 4 * c, y);
    if (out_pad != nullptr) {
      if (t >= len[b]) y = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<uint2*>(out_pad + ((size_t)b * (T + 4) + t + 2) * C + 4 * c) = pack4(y);
    } else {
      if (residual != nullptr) {
        const float4 rv = *reinterpret_cast<const float4*>(residual + r * C + 4 * c);
        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
      }
      *reinterpret_cast<float4*>(out_f32 + r * C + 4 * c) = y;
    }
  }
}
// backward.  dout = gradient w.r.t. the layer output (fp32 [B][T][C]; hidden layers: already the gradient of the NEXT
// layer's masked input, so the length mask is applied here).  dy = dout keep scale mask (1 - tanh^2).
// pass A: column partials of sum dy and sum dy xhat; pass B: dz = gamma invstd (dy - mean(dy) - xhat mean(dy xhat)) as the
// zero-padded bf16 buffer [B][T+4][C] the conv dgrad / wgrad GEMMs read.
__device__ __forceinline__ float4 bn_dy(const float* z, const float* dout, long long r, int c, int C, const float4& mu, const float4& is,
                                        const float4& g, const float4& be, int act_tanh, const Drop& d, bool masked, float4& xh) {
  const float4 zv = *reinterpret_cast<const float4*>(z + r * C + 4 * c);
  xh = make_float4((zv.x - mu.x) * is.x, (zv.y - mu.y) * is.y, (zv.z - mu.z) * is.z, (zv.w - mu.w) * is.w);
  float4 dv = masked ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dout + r * C + 4 * c);
  drop4(d, (unsigned long long)r * C + 4 * c, dv);
  if (act_tanh) {
    const float a = tanhf(xh.x * g.x + be.x), b = tanhf(xh.y * g.y + be.y), cc = tanhf(xh.z * g.z + be.z), e = tanhf(xh.w * g.w + be.w);
    dv.x *= 1.f - a * a; dv.y *= 1.f - b * b; dv.z *= 1.f - cc * cc; dv.w *= 1.f - e * e;
  }
  return dv;
}
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ dout,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int act_tanh, Drop d, const int32_t* __restrict__ len, int mask_rows,
                                                            int B, int T, int C, float* __restrict__ part) {
  // thread = 4 columns, CTA strides over rows; blockDim.x >= C / 4 is not required (loop)
  const int nv = C >> 2;
  for (int c = threadIdx.x; c < nv; c += blockDim.x) {
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c), is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), be = __ldg(reinterpret_cast<const float4*>(beta) + c);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (long long r = blockIdx.x; r < (long long)B * T; r += gridDim.x) {
      const int b = (int)(r / T), t = (int)(r - (long long)b * T);
      float4 xh;
      const float4 dy = bn_dy(z, dout, r, c, C, mu, is, g, be, act_tanh, d, mask_rows && t >= len[b], xh);
      s.x += dy.x; s.y += dy.y; s.z += dy.z; s.w += dy.w;
      q.x += dy.x * xh.x; q.y += dy.y * xh.y; q.z += dy.z * xh.z; q.w += dy.w * xh.w;
    }
    reinterpret_cast<float4*>(part + (size_t)blockIdx.x * 2 * C)[c] = s;
    reinterpret_cast<float4*>(part + (size_t)blockIdx.x * 2 * C + C)[c] = q;
  }
}
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ dout,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int act_tanh, Drop d, const int32_t* __restrict__ len, int mask_rows,
                                                           int B, int T, int C, const float* __restrict__ sums /* [2][C] */,
                                                           bf16* __restrict__ dz_pad) {
  const int nv = C >> 2;
  const long long total = (long long)B * T * nv;
  const float inv_n = 1.f / (float)((long long)B * T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i - r * nv), b = (int)(r / T), t = (int)(r - (long long)b * T);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c), is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), be = __ldg(reinterpret_cast<const float4*>(beta) + c);
    const float4 s = __ldg(reinterpret_cast<const float4*>(sums) + c), q = __ldg(reinterpret_cast<const float4*>(sums + C) + c);
    float4 xh;
    const float4 dy = bn_dy(z, dout, r, c, C, mu, is, g, be, act_tanh, d, mask_rows && t >= len[b], xh);
    const float4 o = make_float4(g.x * is.x * (dy.x - s.x * inv_n - xh.x * q.x * inv_n), g.y * is.y * (dy.y - s.y * inv_n - xh.y * q.y * inv_n),
                                 g.z * is.z * (dy.z - s.z * inv_n - xh.z * q.z * inv_n), g.w * is.w * (dy.w - s.w * inv_n - xh.w * q.w * inv_n));
    *reinterpret_cast<uint2*>(dz_pad + ((size_t)b * (T + 4) + t + 2) * C + 4 * c) = pack4(o);
  }
}
// [B][T][C] fp32 -> masked, zero-padded bf16 [B][T+4][C] (the input of the first Postnet convolution; also zeroes the pads)
__global__ void __launch_bounds__(256) pad_cast_kernel(const float* __restrict__ x, const int32_t* __restrict__ len, bf16* __restrict__ out,
                                                       int B, int T, int C, int only_pads) {
  const int nv = C >> 2;
  const long long total = (long long)B * (T + 4) * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i - r * nv), b = (int)(r / (T + 4)), t = (int)(r - (long long)b * (T + 4)) - 2;
    const bool pad = t < 0 || t >= T;
    if (only_pads && !pad) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!pad && (len == nullptr || t < len[b])) v = *reinterpret_cast<const float4*>(x + ((size_t)b * T + t) * C + 4 * c);
    *reinterpret_cast<uint2*>(out + r * C + 4 * c) = pack4(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// loss (tacotron.py:136-158 without the L2 term) and its gradients in one pass
// out[0] = sum_valid mean_m (bef - tgt)^2, out[1] = same for aft, out[2] = sum_valid bce, per-sample aft sums in aft_b[B]
// d_bef / d_aft / d_stop are the gradients of (bef_loss + aft_loss + stop_loss) (each a masked mean over sum(len) frames)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) loss_kernel(const float* __restrict__ bef, const float* __restrict__ aft,
                                                   const float* __restrict__ stop, const float* __restrict__ tgt,
                                                   const int32_t* __restrict__ len, const int32_t* __restrict__ total_len, int B, int T,
                                                   int M, float pos_weight, float* __restrict__ out, float* __restrict__ aft_b,
                                                   float* __restrict__ d_bef, float* __restrict__ d_aft, float* __restrict__ d_stop) {
  __shared__ float sh[32];
  const int row = blockIdx.x, b = row / T, t = row - b * T;
  const bool on = t < len[b];
  const float n_valid = (float)(*total_len);
  const float gs = on ? 2.f / ((float)M * n_valid) : 0.f;
  float sb = 0.f, sa = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const size_t i = (size_t)row * M + m;
    const float tg = tgt[i], eb = bef[i] - tg, ea = aft[i] - tg;
    sb += eb * eb;
    sa += ea * ea;
    if (d_bef) d_bef[i] = gs * eb;
    if (d_aft) d_aft[i] = gs * ea;
  }
  sb = block_sum(sb, sh);
  sa = block_sum(sa, sh);
  if (threadIdx.x == 0) {
    const float x = stop[row], y = (t == len[b] - 1) ? 1.f : 0.f;
    // BCE with logits, pos_weight on the positive term: l = pw y softplus(-x) + (1 - y) softplus(x)
    const float sp_neg = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))), sp_pos = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    const float ce = pos_weight * y * sp_neg + (1.f - y) * sp_pos;
    const float sig = 1.f / (1.f + expf(-x));
    if (d_stop) d_stop[row] = on ? (-(pos_weight * y) * (1.f - sig) + (1.f - y) * sig) / n_valid : 0.f;
    if (on) {
      atomicAdd(out + 0, sb / (float)M);
      atomicAdd(out + 1, sa / (float)M);
      atomicAdd(out + 2, ce);
      atomicAdd(aft_b + b, sa / (float)M);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-tensor kernels over a device table: L2 term (sum of squares of the decayed tensors) and fused Adam
// ---------------------------------------------------------------------------------------------------------------------
struct OptEntry {
  float* p; const float* g; float* m; float* v; long long n; long long first_chunk; int decay; int pad;
};
__device__ __forceinline__ int find_entry(const OptEntry* tab, int n_entries, long long chunk) {
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].first_chunk <= chunk) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256) sumsq_kernel(const OptEntry* __restrict__ tab, int n_entries, float* __restrict__ out) {
  __shared__ float sh[32];
  const OptEntry e = tab[find_entry(tab, n_entries, blockIdx.x)];
  float a = 0.f;
  if (e.decay) {
    const long long base = ((long long)blockIdx.x - e.first_chunk) * kChunk, end = min(e.n, base + kChunk);
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) a += e.p[i] * e.p[i];
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0 && e.decay) atomicAdd(out, a);
}
// torch.optim.Adam semantics (train.py:130: lr, eps, betas (0.9, 0.999), no amsgrad), with the reference's L2 loss term
// folded in as its exact gradient reg_weight * p on the tensors compute_loss decays (tacotron.py:144-146)
__global__ void __launch_bounds__(256) adam_kernel(const OptEntry* __restrict__ tab, int n_entries, float lr, float beta1, float beta2,
                                                   float eps, float bc1, float bc2_sqrt, float reg_weight, float grad_scale) {
  const OptEntry e = tab[find_entry(tab, n_entries, blockIdx.x)];
  const long long base = ((long long)blockIdx.x - e.first_chunk) * kChunk, end = min(e.n, base + kChunk);
  const float step = lr / bc1;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    const float p = e.p[i];
    float g = e.g[i] * grad_scale;
    if (e.decay) g += reg_weight * p;
    const float m = beta1 * e.m[i] + (1.f - beta1) * g;
    const float v = beta2 * e.v[i] + (1.f - beta2) * g * g;
    e.m[i] = m;
    e.v[i] = v;
    e.p[i] = p - step * m / (sqrtf(v) / bc2_sqrt + eps);
  }
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace tr
}  // namespace tts

using namespace tts;
using tts::tr::bf16;

extern "C" int tts_ln_fwd_train(const float* x, uint16_t* y, int64_t ldy, const float* gamma, const float* beta, float* mean,
                                float* rstd, int32_t rows, int32_t channels, float eps, const int32_t* row_len,
                                int32_t rows_per_batch, void* stream) {
  TTS_REQUIRE(x && y && gamma && beta && mean && rstd && rows > 0, "ln_fwd_train: bad arguments");
  TTS_REQUIRE(channels % 4 == 0 && channels <= 128 * tr::kLnMaxV && ldy % 4 == 0, "ln_fwd_train: channels %d unsupported", channels);
  tr::ln_fwd_kernel<<<ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<bf16*>(y), ldy, gamma, beta, mean, rstd, rows, channels, eps, row_len, rows_per_batch > 0 ? rows_per_batch : rows);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_ln_bwd_train(const uint16_t* dy, int64_t lddy, const float* x, const float* mean, const float* rstd,
                                const float* gamma, const float* dres, float* dx, float* dgamma, float* dbeta, float* scratch,
                                int32_t rows, int32_t channels, const int32_t* row_len, int32_t rows_per_batch, uint16_t* dyb,
                                int64_t lddyb, float drop_p, uint64_t seed, uint32_t rng_stream, void* stream) {
  TTS_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && scratch && rows > 0, "ln_bwd_train: bad arguments");
  TTS_REQUIRE(channels % 4 == 0 && channels <= 128 * tr::kLnMaxV && lddy % 4 == 0 && lddyb % 4 == 0, "ln_bwd_train: channels %d unsupported", channels);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int blocks = ceil_div(rows, 8);
  if (blocks > 444) blocks = 444;   // 3 CTAs of 8 warps per SM
  const size_t smem = (size_t)16 * channels * sizeof(float);
  const int nvl = ceil_div(channels / 4, 32);   // float4 per lane
  const tr::Drop d = tr::make_drop(dyb ? drop_p : 0.f, seed, rng_stream);
  const int rpb = rows_per_batch > 0 ? rows_per_batch : rows;
  static const bool early = getenv("TTS_LN_BWD_EARLY") != nullptr && atoi(getenv("TTS_LN_BWD_EARLY")) != 0;
#define TTS_LN_BWD(NV)                                                                                                              \
  do {                                                                                                                              \
    static bool attr = false;                                                                                                       \
    if (!attr) {                                                                                                                    \
      TTS_CHECK_CUDA(cudaFuncSetAttribute(tr::ln_bwd_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));    \
      TTS_CHECK_CUDA(cudaFuncSetAttribute(tr::ln_bwd_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));   \
      attr = true;                                                                                                                  \
    }                                                                                                                               \
    if (early)                                                                                                                      \
      tr::ln_bwd_kernel<NV, true><<<blocks, 256, smem, s>>>(reinterpret_cast<const bf16*>(dy), lddy, x, mean, rstd, gamma, dres, dx, \
                                                    reinterpret_cast<bf16*>(dyb), lddyb, d, scratch, rows, channels, row_len, rpb); \
    else                                                                                                                            \
      tr::ln_bwd_kernel<NV, false><<<blocks, 256, smem, s>>>(reinterpret_cast<const bf16*>(dy), lddy, x, mean, rstd, gamma, dres, dx, \
                                                    reinterpret_cast<bf16*>(dyb), lddyb, d, scratch, rows, channels, row_len, rpb); \
  } while (0)
  if (nvl <= 2) TTS_LN_BWD(2);
  else if (nvl <= 4) TTS_LN_BWD(4);
  else if (nvl <= 6) TTS_LN_BWD(6);
  else TTS_LN_BWD(8);
#undef TTS_LN_BWD
  TTS_CHECK_LAUNCH();
  tr::colpart_finalize_kernel<<<ceil_div(2 * channels, 32), 256, 0, s>>>(scratch, blocks, 2 * channels, dgamma, dbeta, channels);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" size_t tts_ln_bwd_scratch_floats(int32_t channels) { return (size_t)592 * 2 * channels; }

extern "C" int tts_dropout_cast(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int32_t channels,
                                float drop_p, uint64_t seed, uint32_t rng_stream, const int32_t* row_len, int32_t rows_per_batch,
                                void* stream) {
  TTS_REQUIRE(src && dst && rows > 0 && channels % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "dropout_cast: bad arguments (C=%d)", channels);
  tr::drop_cast_kernel<<<tr::grid_for(rows * (channels / 4), 1024), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, lds, reinterpret_cast<bf16*>(dst), ldd, rows, channels, tr::make_drop(drop_p, seed, rng_stream), row_len,
      rows_per_batch > 0 ? rows_per_batch : (int)rows);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_multi_cast_bf16(const void* table_dev, int32_t n_entries, int64_t n_chunks, void* stream) {
  TTS_REQUIRE(table_dev && n_entries > 0 && n_chunks > 0, "multi_cast: bad arguments");
  tr::multi_cast_kernel<<<(unsigned)n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const tr::CastEntry*>(table_dev), n_entries);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_embed_train_fwd(const int64_t* ids, const int32_t* lengths, const float* embed, const float* pe,
                                   const float* pe_scale, float* out, int32_t batch, int32_t seq, int32_t channels, int32_t vocab,
                                   float drop_p, uint64_t seed, uint32_t rng_stream, void* stream) {
  TTS_REQUIRE(ids && lengths && embed && pe && pe_scale && out && channels % 4 == 0, "embed_train_fwd: bad arguments");
  tr::embed_fwd_kernel<<<batch * seq, 128, 0, static_cast<cudaStream_t>(stream)>>>(ids, lengths, embed, pe, pe_scale, out, batch, seq,
                                                                                    channels, vocab, tr::make_drop(drop_p, seed, rng_stream));
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_embed_train_bwd(const float* dx, const int64_t* ids, const int32_t* lengths, const float* pe, float* d_embed,
                                   float* d_pe_scale, int32_t batch, int32_t seq, int32_t channels, int32_t vocab, float drop_p,
                                   uint64_t seed, uint32_t rng_stream, void* stream) {
  TTS_REQUIRE(dx && ids && lengths && pe && d_embed && d_pe_scale && channels % 4 == 0, "embed_train_bwd: bad arguments");
  tr::embed_bwd_kernel<<<batch * seq, 128, 0, static_cast<cudaStream_t>(stream)>>>(dx, ids, lengths, pe, d_embed, d_pe_scale, batch, seq,
                                                                                    channels, vocab, tr::make_drop(drop_p, seed, rng_stream));
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_shift_pe_train_fwd(const float* pre, const int32_t* lengths, const float* pe, const float* pe_scale, float* out,
                                      int32_t batch, int32_t frames, int32_t channels, float drop_p, uint64_t seed,
                                      uint32_t rng_stream, void* stream) {
  TTS_REQUIRE(pre && lengths && pe && pe_scale && out && channels % 4 == 0, "shift_pe_train_fwd: bad arguments");
  tr::shift_fwd_kernel<<<batch * frames, 192, 0, static_cast<cudaStream_t>(stream)>>>(pre, lengths, pe, pe_scale, out, batch, frames,
                                                                                       channels, tr::make_drop(drop_p, seed, rng_stream));
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_shift_pe_train_bwd(const float* dx, const int32_t* lengths, const float* pe, uint16_t* dpre, float* d_pe_scale,
                                      int32_t batch, int32_t frames, int32_t channels, float drop_p, uint64_t seed,
                                      uint32_t rng_stream, void* stream) {
  TTS_REQUIRE(dx && lengths && pe && dpre && d_pe_scale && channels % 4 == 0, "shift_pe_train_bwd: bad arguments");
  tr::shift_bwd_kernel<<<batch * frames, 192, 0, static_cast<cudaStream_t>(stream)>>>(
      dx, lengths, pe, reinterpret_cast<bf16*>(dpre), d_pe_scale, batch, frames, channels, tr::make_drop(drop_p, seed, rng_stream));
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_colsum_bf16(const uint16_t* x, int64_t ldx, const float* row_weight, float* out, int64_t rows, int32_t channels,
                               void* stream) {
  TTS_REQUIRE(x && out && rows > 0 && channels > 0, "colsum: bad arguments");
  int blocks = (int)(rows < 592 ? rows : 592);
  tr::colsum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(x), ldx, row_weight, out, rows, channels);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_sum_f32(const float* x, int64_t n, float* out, void* stream) {
  TTS_REQUIRE(x && out && n > 0, "sum_f32: bad arguments");
  tr::sum_f32_kernel<<<tr::grid_for(n, 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_rowdot_bf16(const uint16_t* x, int64_t ldx, const float* w, const float* bias, const int32_t* row_len,
                               int32_t rows_per_batch, float* out, int64_t rows, int32_t k, void* stream) {
  TTS_REQUIRE(x && w && out && rows > 0 && k % 4 == 0 && ldx % 4 == 0, "rowdot: bad arguments");
  tr::rowdot_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), ldx, w, bias, row_len, rows_per_batch > 0 ? rows_per_batch : (int)rows, out, rows, k);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t tts_bn_scratch_floats(int32_t channels) { return (size_t)(592 + 1) * 2 * channels; }
extern "C" int tts_bn_train_fwd(const float* z, const float* gamma, const float* beta, float* mean, float* invstd,
                                float* running_mean, float* running_var, int64_t* num_batches, float momentum, float eps,
                                int32_t act_tanh, float drop_p, uint64_t seed, uint32_t rng_stream, const int32_t* lengths,
                                int32_t batch, int32_t frames, int32_t channels, uint16_t* out_pad, float* out_f32,
                                const float* residual, float* scratch, void* stream) {
  TTS_REQUIRE(z && gamma && beta && mean && invstd && scratch && (out_pad || out_f32) && channels % 4 == 0, "bn_train_fwd: bad arguments");
  TTS_REQUIRE(out_pad == nullptr || lengths != nullptr, "bn_train_fwd: the padded output needs lengths");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long rows = (long long)batch * frames;
  const int blocks = (int)(rows < 592 ? rows : 592);
  tr::bn_stats_kernel<<<blocks, 256, 0, s>>>(z, rows, channels, scratch);
  TTS_CHECK_LAUNCH();
  tr::bn_finalize_kernel<<<ceil_div(channels, 32), 256, 0, s>>>(scratch, blocks, rows, channels, eps, mean, invstd, running_mean,
                                                                 running_var, reinterpret_cast<long long*>(num_batches), momentum);
  TTS_CHECK_LAUNCH();
  tr::bn_apply_kernel<<<tr::grid_for(rows * (channels / 4), 1024), 256, 0, s>>>(
      z, mean, invstd, gamma, beta, act_tanh, tr::make_drop(drop_p, seed, rng_stream), lengths, batch, frames, channels,
      reinterpret_cast<bf16*>(out_pad), out_f32, residual);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_bn_train_bwd(const float* z, const float* dout, const float* gamma, const float* beta, const float* mean,
                                const float* invstd, int32_t act_tanh, float drop_p, uint64_t seed, uint32_t rng_stream,
                                const int32_t* lengths, int32_t mask_rows, int32_t batch, int32_t frames, int32_t channels,
                                uint16_t* dz_pad, float* dgamma, float* dbeta, float* scratch, void* stream) {
  TTS_REQUIRE(z && dout && gamma && beta && mean && invstd && dz_pad && dgamma && dbeta && scratch && channels % 4 == 0,
              "bn_train_bwd: bad arguments");
  TTS_REQUIRE(!mask_rows || lengths != nullptr, "bn_train_bwd: mask_rows needs lengths");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long rows = (long long)batch * frames;
  const int blocks = (int)(rows < 592 ? rows : 592);
  const tr::Drop d = tr::make_drop(drop_p, seed, rng_stream);
  tr::bn_bwd_reduce_kernel<<<blocks, 128, 0, s>>>(z, dout, mean, invstd, gamma, beta, act_tanh, d, lengths, mask_rows, batch, frames,
                                                  channels, scratch);
  TTS_CHECK_LAUNCH();
  float* sums = scratch + (size_t)592 * 2 * channels;   // [2][C]: sum dy (= dbeta), sum dy xhat (= dgamma)
  tr::colpart_finalize_kernel<<<ceil_div(2 * channels, 32), 256, 0, s>>>(scratch, blocks, 2 * channels, sums, sums + channels, channels);
  TTS_CHECK_LAUNCH();
  TTS_CHECK_CUDA(cudaMemcpyAsync(dbeta, sums, channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  TTS_CHECK_CUDA(cudaMemcpyAsync(dgamma, sums + channels, channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  tr::bn_bwd_apply_kernel<<<tr::grid_for(rows * (channels / 4), 1024), 256, 0, s>>>(z, dout, mean, invstd, gamma, beta, act_tanh, d, lengths,
                                                                                    mask_rows, batch, frames, channels, sums,
                                                                                    reinterpret_cast<bf16*>(dz_pad));
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_pad_cast_bf16(const float* x, const int32_t* lengths, uint16_t* out, int32_t batch, int32_t frames,
                                 int32_t channels, int32_t only_pads, void* stream) {
  TTS_REQUIRE(out && (x || only_pads) && channels % 4 == 0, "pad_cast: bad arguments");
  tr::pad_cast_kernel<<<tr::grid_for((long long)batch * (frames + 4) * (channels / 4), 1024), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, lengths, reinterpret_cast<bf16*>(out), batch, frames, channels, only_pads);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_loss_train(const float* mel_bef, const float* mel_aft, const float* stop_logits, const float* targets,
                              const int32_t* lengths, const int32_t* total_len, int32_t batch, int32_t frames, int32_t n_mels,
                              float pos_weight, float* sums3, float* aft_per_sample, float* d_bef, float* d_aft, float* d_stop,
                              void* stream) {
  TTS_REQUIRE(mel_bef && mel_aft && stop_logits && targets && lengths && total_len && sums3 && aft_per_sample, "loss_train: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TTS_CHECK_CUDA(cudaMemsetAsync(sums3, 0, 4 * sizeof(float), s));
  TTS_CHECK_CUDA(cudaMemsetAsync(aft_per_sample, 0, batch * sizeof(float), s));
  tr::loss_kernel<<<batch * frames, 128, 0, s>>>(mel_bef, mel_aft, stop_logits, targets, lengths, total_len, batch, frames, n_mels,
                                                pos_weight, sums3, aft_per_sample, d_bef, d_aft, d_stop);
  TTS_CHECK_LAUNCH();
  return 0;
}

extern "C" int tts_sumsq_multi(const void* table_dev, int32_t n_entries, int64_t n_chunks, float* out, void* stream) {
  TTS_REQUIRE(table_dev && out && n_entries > 0 && n_chunks > 0, "sumsq_multi: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TTS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
  tr::sumsq_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(reinterpret_cast<const tr::OptEntry*>(table_dev), n_entries, out);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int tts_adam_multi(const void* table_dev, int32_t n_entries, int64_t n_chunks, float lr, float beta1, float beta2,
                              float eps, int64_t step, float reg_weight, float grad_scale, void* stream) {
  TTS_REQUIRE(table_dev && n_entries > 0 && n_chunks > 0 && step >= 1, "adam_multi: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  tr::adam_kernel<<<(unsigned)n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const tr::OptEntry*>(table_dev), n_entries, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), reg_weight, grad_scale);
  TTS_CHECK_LAUNCH();
  return 0;
}
extern "C" int32_t tts_multi_chunk_elems(void) { return tr::kChunk; }
