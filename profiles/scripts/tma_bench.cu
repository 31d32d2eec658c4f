#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one CTA per SM streams its contiguous share of `src` through a ring of `slots` buffers of `chunk` bytes (thread 0 only)
__global__ void k(const unsigned char* src, size_t per_cta, int chunk, int slots, size_t misalign, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < slots; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned char* base = src + (size_t)blockIdx.x * per_cta + misalign;
    const int n = (int)(per_cta / chunk) - 1;
    unsigned acc = 0;
    for (int i = 0; i < slots && i < n; ++i) { mbar_expect_tx(&full[i], chunk); bulk_g2s(smem + (size_t)i * chunk, base + (size_t)i * chunk, chunk, &full[i]); }
    for (int i = 0; i < n; ++i) {
      const int s = i % slots;
      mbar_wait(&full[s], (i / slots) & 1);
      acc += *reinterpret_cast<unsigned*>(smem + (size_t)s * chunk);
      if (i + slots < n) { mbar_expect_tx(&full[s], chunk); bulk_g2s(smem + (size_t)s * chunk, base + (size_t)(i + slots) * chunk, chunk, &full[s]); }
    }
    sink[blockIdx.x] = acc;
  }
}
int main() {
  const size_t total = (size_t)148 * 8 * 1024 * 1024;   // 1.2 GB > L2
  unsigned char* src; unsigned* sink;
  cudaMalloc(&src, total + 4096); cudaMalloc(&sink, 148 * 4);
  cudaMemset(src, 1, total + 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (size_t mis : {(size_t)0, (size_t)16}) {
    for (int chunk : {4096, 8192, 16384, 32768}) {
      for (int slots : {2, 4, 6}) {
        if ((size_t)chunk * slots > 196 * 1024) continue;
        const size_t per = total / 148;
        k<<<148, 32, (size_t)chunk * slots>>>(src, per, chunk, slots, mis, sink);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148, 32, (size_t)chunk * slots>>>(src, per, chunk, slots, mis, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("misalign %2zu chunk %6d slots %d: %.3f ms  %.0f GB/s total, %.1f GB/s per SM, in flight %d KB -> implied latency %.2f us\n", mis, chunk, slots, ms,
               total / ms / 1e6, total / ms / 1e6 / 148, chunk * slots / 1024, (double)chunk * slots / (total / ms / 1e6 / 148 * 1e9) * 1e6);
      }
    }
  }
  return 0;
}
