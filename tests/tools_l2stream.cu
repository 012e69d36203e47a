// Standalone B200 micro-benchmark: how fast can 148 CTAs stream K/V-like data through a TMA-fed shared-memory ring
// (a) from HBM (cold), (b) from L2 (warm), (c) after cp.async.bulk.prefetch.L2 of the region, as a function of the
// ring depth (bytes in flight per SM).  Behind the decode kernel's attention phases.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2stream tests/tools_l2stream.cu
// Diagnostics only: nothing here is on the product path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kTile = 3072;   // bytes of a K (or V) tile: 8 keys x 96 floats

// lane `l` of warp 0 owns ring slot l: issue (K tile + V tile), wait, reissue.  No consumers: pure memory system.
__global__ void __launch_bounds__(64, 1) stream_kernel(const char* src, size_t bytes_per_cta, int slots, float* sink) {
  extern __shared__ __align__(128) char ring[];
  __shared__ uint64_t full[32];
  const int lane = threadIdx.x;
  if (lane < 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[lane])));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x >= 32 || lane >= slots) return;
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  const int n_tiles = (int)(bytes_per_cta / (2 * kTile));
  unsigned par = 0;
  float acc = 0.f;
  for (int k = lane; k < n_tiles; k += slots) {
    char* dst = ring + (size_t)lane * 2 * kTile;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[lane])), "r"(2u * kTile) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(base + (size_t)k * 2 * kTile), "r"((unsigned)kTile), "r"(smem_u32(&full[lane])) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst + kTile)), "l"(base + (size_t)k * 2 * kTile + kTile), "r"((unsigned)kTile), "r"(smem_u32(&full[lane])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&full[lane])), "r"(par) : "memory");
    par ^= 1u;
    acc += *reinterpret_cast<float*>(dst);
  }
  if (acc == 12345.f) sink[0] = acc;
}

// every CTA asks for its region to be pulled into L2 in `chunk`-byte prefetches issued by `lanes` lanes
__global__ void prefetch_kernel(const char* src, size_t bytes_per_cta, unsigned chunk) {
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  for (size_t off = (size_t)threadIdx.x * chunk; off < bytes_per_cta; off += (size_t)blockDim.x * chunk) {
    const unsigned n = (unsigned)((bytes_per_cta - off) < chunk ? (bytes_per_cta - off) : chunk);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(n) : "memory");
  }
}

__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

__global__ void fill_kernel(float4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float* sink; CK(cudaMalloc(&sink, 1024));
  const size_t big = (size_t)sms * (4u << 20);   // 592 MB: > L2
  char *a, *flush;
  CK(cudaMalloc(&a, big)); CK(cudaMalloc(&flush, big));
  fill_kernel<<<1024, 256>>>((float4*)a, big / 16);
  fill_kernel<<<1024, 256>>>((float4*)flush, big / 16);
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 2 * kTile));
  CK(cudaDeviceSynchronize());
  auto run_stream = [&](const char* src, size_t per_cta, int slots) {
    CK(cudaEventRecord(e0));
    stream_kernel<<<sms, 64, 32 * 2 * kTile>>>(src, per_cta, slots, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
  };
  auto do_flush = [&]() { fill_kernel<<<1024, 256>>>((float4*)flush, big / 16); CK(cudaDeviceSynchronize()); };

  printf("SMs %d.  ring slot = 6 KB (3 KB K tile + 3 KB V tile); TB/s = bytes streamed / kernel time (incl. ~3 us launch)\n", sms);
  printf("== (a) HBM cold: 592 MB region, 4 MB per CTA\n");
  for (int slots : {4, 6, 8, 10, 12, 16, 20, 24, 32}) {
    do_flush();
    const float ms = run_stream(a, 4u << 20, slots);
    printf("  slots %2d (%3d KB in flight / SM): %.3f ms  %.2f TB/s\n", slots, slots * 6, ms, big / ms / 1e9);
  }
  for (size_t mb_total : {24, 48, 72, 96}) {
    const size_t per_cta = (mb_total << 20) / sms / (2 * kTile) * (2 * kTile);
    const double tot = (double)per_cta * sms;
    printf("== region %zu MB total (%zu KB per CTA)\n", mb_total, per_cta >> 10);
    for (int slots : {6, 10, 12, 20}) {
      do_flush();
      const float cold = run_stream(a, per_cta, slots);
      const float warm = run_stream(a, per_cta, slots);
      printf("  slots %2d: cold %.1f us %.2f TB/s | warm (L2) %.1f us %.2f TB/s", slots, cold * 1e3, tot / cold / 1e9, warm * 1e3, tot / warm / 1e9);
      for (unsigned chunk : {6144u, 49152u}) {
        do_flush();
        CK(cudaEventRecord(e0));
        prefetch_kernel<<<sms, 32>>>(a, per_cta, chunk);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float pms; CK(cudaEventElapsedTime(&pms, e0, e1));
        spin_kernel<<<1, 1>>>(60000);   // ~30 us: let the prefetches land
        const float pf = run_stream(a, per_cta, slots);
        printf(" | prefetch(chunk %u: issue %.1f us) then %.1f us %.2f TB/s", chunk, pms * 1e3, pf * 1e3, tot / pf / 1e9);
      }
      printf("\n");
    }
  }
  // (d) prefetch racing the demand stream: prefetch region B while streaming region A cold, then stream B
  {
    const size_t per_cta = (48u << 20) / sms / (2 * kTile) * (2 * kTile);
    const double tot = (double)per_cta * sms;
    const char* regB = a + ((size_t)300 << 20);
    do_flush();
    const float a_alone = run_stream(a, per_cta, 20);
    do_flush();
    cudaStream_t s2; CK(cudaStreamCreate(&s2));
    prefetch_kernel<<<sms, 32, 0, s2>>>(regB, per_cta, 49152u);
    const float a_with = run_stream(a, per_cta, 20);
    CK(cudaDeviceSynchronize());
    const float b_after = run_stream(regB, per_cta, 20);
    printf("== (d) 48 MB regions: A cold alone %.1f us (%.2f TB/s); A cold while B is prefetched %.1f us; B afterwards %.1f us (%.2f TB/s)\n",
           a_alone * 1e3, tot / a_alone / 1e9, a_with * 1e3, b_after * 1e3, tot / b_after / 1e9);
  }
  return 0;
}
