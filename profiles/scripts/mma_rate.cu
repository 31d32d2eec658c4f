// mma_rate.cu - how long does one tcgen05.mma take with shared-memory operands in the layouts this repository uses?
// One CTA, thread 0 issues `iters` MMAs (M = 128) back to back into TMEM and waits for the commit; cycles per MMA from clock64.
// Operand contents are irrelevant (zeros).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sparsernns_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include "tc.cuh"
using namespace s5;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return tc_smem_desc(addr, lbo, sbo) | ((uint64_t)layout << 61);
}
template <int KIND>   // 0 tf32 (K = 8), 1 f16/bf16 (K = 16), 2 i8 (K = 32): 32 operand bytes per row and MMA in every case
__global__ void rate_kernel(int N, int iters, uint32_t lbo, uint32_t sbo, uint32_t layout, int alternate, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  if (threadIdx.x < 32) tc_alloc(&slot, 512);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    const uint32_t idesc = KIND == 0 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                         : KIND == 1 ? tc_idesc_bf16(128, N) : tc_idesc_i8(128, N, true, true);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // walk over 4 K steps inside the tile like the real kernels do (start address advanced by 32 bytes or 2 LBO)
      const uint32_t ko = layout ? 32u * (i & 3) : 2u * lbo * (i & 3);
      const uint32_t d = slot + (alternate ? 256u * (i & 1) : 0u);
      const uint64_t da = desc(a0 + ko, lbo, sbo, layout), db = desc(b0 + ko, lbo, sbo, layout);
      if (KIND == 0) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      else if (KIND == 1) tc_mma_f16(d, da, db, idesc, true);
      else tc_mma_i8(d, da, db, idesc, true);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc_dealloc(slot, 512);
}

template <int KIND>
static void run(const char* name, int N, uint32_t lbo, uint32_t sbo, uint32_t layout, int alternate) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 512;
  cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<KIND><<<1, 128, 160 * 1024>>>(N, iters, lbo, sbo, layout, alternate, d);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s N=%3d alt=%d : %7.1f cycles / MMA  (%s)\n", name, N, alternate, (double)h / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int alt = 0; alt < 2; ++alt) {
    run<0>("tf32 K=8  no swizzle, LBO 2064 (padded)", 128, 2064, 128, 0, alt);
    run<0>("tf32 K=8  no swizzle, LBO 2080 (padded 32)", 128, 2080, 128, 0, alt);
    run<0>("tf32 K=8  no swizzle, LBO 2048", 128, 2048, 128, 0, alt);
    run<0>("tf32 K=8  SWIZZLE_128B, SBO 1024", 128, 16, 1024, 2, alt);
    run<0>("tf32 K=8  SWIZZLE_64B, SBO 512", 128, 16, 512, 4, alt);
    run<1>("bf16 K=16 no swizzle, LBO 2064", 128, 2064, 128, 0, alt);
    run<1>("bf16 K=16 SWIZZLE_128B", 128, 16, 1024, 2, alt);
    run<2>("i8   K=32 no swizzle, LBO 2064", 128, 2064, 128, 0, alt);
    run<2>("i8   K=32 no swizzle, LBO 2064", 96, 2064, 128, 0, alt);
    run<2>("i8   K=32 SWIZZLE_128B", 128, 16, 1024, 2, alt);
    run<2>("i8   K=32 no swizzle, LBO 4112 (256 rows)", 256, 4112, 128, 0, alt);
    run<2>("i8   K=32 SWIZZLE_128B", 256, 16, 1024, 2, alt);
  }
  return 0;
}
