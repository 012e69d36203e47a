// Standalone B200 micro-benchmarks behind the design choices of the pipelined decode kernel:
//   mma     legacy mma.sync throughput (TF32 m16n8k8, BF16 m16n8k16) and packed FFMA2 throughput per SM
//   bar     latency of a software grid barrier (red.release.gpu + ld.acquire.gpu poll), 148 CTAs
//   bcast   every CTA bulk-copies (TMA, 1-D) the same 16 x 3 KB activation tile from L2
//   chain   barrier followed by the tile copy (the serial latency a non-pipelined GEMM phase pays)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tests/tools_microbench.cu
// Diagnostics only: nothing here is on the product path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) pipe_kernel(float* out, int iters) {
  float acc[8][4];
  unsigned long long acc2[16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc2[i] = 0ull;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 1, b1 = 5;
  unsigned long long x2 = 0x3f8000003f800000ull, y2 = 0x3f0000003f000000ull;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(x2), "l"(y2));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += (float)(acc2[i] & 0xff);
  if (s == 12345.f) out[0] = s;
}

__device__ __forceinline__ void grid_bar(unsigned* ctr, unsigned& epoch, unsigned n) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned target = epoch * n;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

// mode 0: barriers only; 1: tile copies only; 2: barrier then tile copy; rows_per_copy: 1 (16 row copies) or 16 (one copy)
__global__ void __launch_bounds__(288, 1) sync_kernel(unsigned* ctr, const float* tiles, int n_tiles, int iters, int mode,
                                                      int one_copy, float* sink) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t full;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned epoch = 0, par = 0;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (mode != 1) grid_bar(ctr, epoch, gridDim.x);
    if (mode != 0) {
      const float* src = tiles + (size_t)(it % n_tiles) * 16 * 768;
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(16u * 3072u) : "memory");
        if (one_copy) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sm)), "l"(src), "r"(16u * 3072u), "r"(smem_u32(&full)) : "memory");
        } else {
          for (int r = 0; r < 16; ++r)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm + r * 784)), "l"(src + r * 768), "r"(3072u), "r"(smem_u32(&full)) : "memory");
        }
      }
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&full)), "r"(par) : "memory");
      par ^= 1u;
      acc += sm[threadIdx.x];
      __syncthreads();
    }
  }
  if (acc == 12345.f) sink[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  printf("SMs %d, max clock %d MHz\n", sms, khz / 1000);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float* out; CK(cudaMalloc(&out, 1024));
  const int iters = 20000;
  const char* names[3] = {"mma.sync m16n8k8 tf32", "mma.sync m16n8k16 bf16", "fma.rn.f32x2"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      if (mode == 0) pipe_kernel<0><<<sms, 256>>>(out, iters);
      if (mode == 1) pipe_kernel<1><<<sms, 256>>>(out, iters);
      if (mode == 2) pipe_kernel<2><<<sms, 256>>>(out, iters);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    const double ms = time_ms(e0, e1);
    const double per_warp = (mode == 2 ? 16.0 : 8.0) * iters;   // instructions per warp
    const double inst_per_sm = per_warp * 8;
    const double ns_per_inst_sm = ms * 1e6 / inst_per_sm;
    const double fma_per_inst = mode == 0 ? 1024.0 : (mode == 1 ? 2048.0 : 64.0);
    printf("%-24s %.3f ms: %.2f ns per instruction per SM (8 warps) -> %.1f FMA/ns/SM, %.1f TFLOP/s chip\n", names[mode], ms,
           ns_per_inst_sm, fma_per_inst / ns_per_inst_sm, 2.0 * fma_per_inst / ns_per_inst_sm * sms / 1000.0);
  }

  unsigned* ctr; CK(cudaMalloc(&ctr, 256));
  const int n_tiles = 64;
  float* tiles; CK(cudaMalloc(&tiles, (size_t)n_tiles * 16 * 768 * 4)); CK(cudaMemset(tiles, 0, (size_t)n_tiles * 16 * 768 * 4));
  CK(cudaFuncSetAttribute(sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int n_it = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int one = 0; one < (mode == 0 ? 1 : 2); ++one) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(ctr, 0, 256));
        int m = mode, o = one, nt = n_tiles, ni = n_it;
        void* args[] = {&ctr, &tiles, &nt, &ni, &m, &o, &out};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)sync_kernel, dim3(sms), dim3(288), args, 64 * 1024, 0));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      }
      const double us = time_ms(e0, e1) * 1e3 / n_it;
      const char* what = mode == 0 ? "grid barrier" : (mode == 1 ? "tile copy 48 KB / CTA" : "barrier + tile copy");
      printf("%-24s %s: %.2f us per iteration", what, mode == 0 ? "" : (one ? "(1 copy)" : "(16 row copies)"), us);
      if (mode == 1) printf("  (%.2f TB/s L2->SM aggregate)", (double)sms * 16 * 3072 / us / 1e6);
      printf("\n");
    }
  return 0;
}
