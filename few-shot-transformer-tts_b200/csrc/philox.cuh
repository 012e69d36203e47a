// Counter-based dropout masks (Philox4x32-7: the 7-round variant of Random123, the fewest rounds that pass BigCrush)
// shared by every kernel of the training path and by the decode
// kernel's train()-mode dropout: a mask is a pure function of (seed, stream, element), so the backward kernels
// regenerate it instead of storing it (the reference stores nothing either: torch keeps a byte mask per
// nn.Dropout call, transformer/modules.py:18,120,132,138,141, attention.py:89, tacotron.py:58,62,88).
#pragma once
#include <stdint.h>

namespace tts {

constexpr int kPhiloxRounds = 7;

__device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned long long idx, uint32_t stream) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0x5eedu;
#pragma unroll
  for (int i = 0; i < kPhiloxRounds; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// drop probability p -> 32-bit threshold: an element is KEPT iff its random word >= threshold
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}

// Row-major tensors of the TRAINING path (GEMM epilogues, element-wise kernels): the eight elements 8*g .. 8*g+7 share one
// call, 16 random bits each: element e uses half (e & 1) of word (e & 7) >> 1 and is KEPT iff its 16 bits >= p * 65536
// (drop_threshold16).  (The decode kernel's train()-mode dropout, pipelined.cu drop1, keeps 32-bit lanes: four elements
// per call.)
__device__ __forceinline__ uint4 dropout_words_linear(unsigned long long seed, uint32_t stream, unsigned long long group) {
  return philox4x32(seed, group, stream);
}
__device__ __forceinline__ uint32_t dropout_lane16(const uint4& w, int lane) {   // lane = e & 7
  const uint32_t word = (lane >> 1) == 0 ? w.x : ((lane >> 1) == 1 ? w.y : ((lane >> 1) == 2 ? w.z : w.w));
  return (lane & 1) ? (word >> 16) : (word & 0xffffu);
}

// Attention weights [bh][i][j]: one call covers the 2 x 4 elements {i0, i0+8} x {j0, j0+1, j0+8, j0+9} of a 16 x 16 block
// (i0 in [0, 8), j0 even in [0, 8)) with 16 random bits each: half-word h = 4 * (i bit 3) + 2 * (j bit 3) + (j bit 0) of
// the 128-bit result.  A thread of an m16n8 accumulator of S = Q K^T holds rows {g, g+8} and, over two neighbouring
// n-tiles, columns {2t, 2t+1, 2t+8, 2t+9}: ONE call per 8 weights in the forward and dQ kernels; the dK/dV kernel
// (S^T = K Q^T: rows are keys) needs two calls per 8 weights.  An element is kept iff its 16 bits >= p * 65536.
__device__ __forceinline__ unsigned long long attn_dropout_index(unsigned long long bh, int n_iblk, int n_jblk, int i, int j) {
  return (((bh * (unsigned long long)n_iblk + (unsigned)(i >> 4)) * (unsigned long long)n_jblk + (unsigned)(j >> 4)) << 5) +
         (unsigned)((i & 7) * 4 + ((j & 7) >> 1));
}
__device__ __forceinline__ uint32_t attn_dropout_half(const uint4& w, int i, int j) {   // 16 random bits of element (i, j)
  const int h = 4 * ((i >> 3) & 1) + 2 * ((j >> 3) & 1) + (j & 1);
  const uint32_t word = (h >> 1) == 0 ? w.x : ((h >> 1) == 1 ? w.y : ((h >> 1) == 2 ? w.z : w.w));
  return (h & 1) ? (word >> 16) : (word & 0xffffu);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold16(float p) {
  const float t = p * 65536.f;
  return t >= 65535.f ? 65535u : (uint32_t)t;
}

}  // namespace tts
