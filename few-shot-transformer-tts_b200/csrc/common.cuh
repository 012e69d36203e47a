// Shared device/host helpers for the sm_100a kernels of the Transformer-TTS mel path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tts_b200.h"

namespace tts {

constexpr float kNegBias = -1e20f;  // transformer/common.py:32

// ---- error plumbing (never abort: transformer callers catch Python exceptions) -------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TTS_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::tts::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define TTS_REQUIRE(cond, ...)           \
  do {                                   \
    if (!(cond)) {                       \
      ::tts::set_error(__VA_ARGS__);     \
      return 2;                          \
    }                                    \
  } while (0)

#define TTS_CHECK_LAUNCH()                     \
  do {                                         \
    ::tts::count_launch();                     \
    TTS_CHECK_CUDA(cudaPeekAtLastError());     \
  } while (0)

// ---- warp helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- packed fp32 (Blackwell FFMA2): the only way to reach the full fp32 FMA rate on sm_100 ----
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float hsum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}
// d = a * b + c, element-wise on the two packed floats
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

struct __align__(16) f32x4 {
  f32x2 lo, hi;
};

// streaming (read-once) 128-bit global load that does not pollute L1
__device__ __forceinline__ f32x4 ldg_stream(const float* p) {
  f32x4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
  return r;
}
// coherent 128-bit global load (data written earlier in the same kernel by other CTAs)
__device__ __forceinline__ f32x4 ldg_cg(const float* p) {
  f32x4 r;
  asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
  return r;
}
__device__ __forceinline__ f32x4 lds128(const float* p) {
  f32x4 r;
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "r"(a));
  return r;
}
__device__ __forceinline__ float4 as_float4(const f32x4& v) {
  float4 r;
  unpack2(v.lo, r.x, r.y);
  unpack2(v.hi, r.z, r.w);
  return r;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace tts
