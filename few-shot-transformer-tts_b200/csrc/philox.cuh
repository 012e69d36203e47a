// Counter-based dropout masks (Philox4x32-10) shared by every kernel of the training path and by the decode
// kernel's train()-mode dropout: a mask is a pure function of (seed, stream, element), so the backward kernels
// regenerate it instead of storing it (the reference stores nothing either: torch keeps a byte mask per
// nn.Dropout call, transformer/modules.py:18,120,132,138,141, attention.py:89, tacotron.py:58,62,88).
#pragma once
#include <stdint.h>

namespace tts {

__device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned long long idx, uint32_t stream) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0x5eedu;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
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

// Row-major tensors (GEMM epilogues, element-wise kernels): the four elements 4*g .. 4*g+3 share one call.
__device__ __forceinline__ uint4 dropout_words_linear(unsigned long long seed, uint32_t stream, unsigned long long group) {
  return philox4x32(seed, group, stream);
}

// Attention weights [bh][i][j]: one call covers the 2 x 2 elements {i0, i0+8} x {j0, j0+8} of a 16 x 16 block
// (i0, j0 in [0, 8)), word = 2 * (i bit 3) + (j bit 3).  A thread of an m16n8 MMA accumulator holds rows {g, g+8} and,
// over two neighbouring n-tiles, columns {c, c+8}: exactly one call per 4 elements, both for S = Q K^T (rows are
// queries) and for S^T = K Q^T (rows are keys) - the forward, dQ and dK/dV kernels all pay one call per 4 weights.
__device__ __forceinline__ unsigned long long attn_dropout_index(unsigned long long bh, int n_iblk, int n_jblk, int i, int j) {
  return (((bh * (unsigned long long)n_iblk + (unsigned)(i >> 4)) * (unsigned long long)n_jblk + (unsigned)(j >> 4)) << 6) +
         (unsigned)((i & 7) * 8 + (j & 7));
}

}  // namespace tts
