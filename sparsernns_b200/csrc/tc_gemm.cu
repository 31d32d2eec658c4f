// tc_gemm.cu - plain dense-tile GEMM on the tcgen05 tensor cores: Y[rows][N] (f32) = A[rows][K] (f16) * W[N][K]^T (f16).
// The op-level form of the dense-tile projection path; the fused model kernels reuse exactly this operand layout,
// descriptor construction and TMEM read-back (tc.cuh).
#include "tc.cuh"

namespace s5 {

__global__ void __launch_bounds__(128) tc_gemm_f16_kernel(const __half* __restrict__ A, const __half* __restrict__ W,
                                                          float* __restrict__ Y, long long rows, int K, int N, uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Wt = smem;                          // [N][K] canonical
  unsigned char* At = smem + (size_t)N * K * 2;      // [128][K] canonical
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * 128;
  if (warp == 0) tc_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  const int kchunks = K >> 3;
  for (int i = tid; i < N * kchunks; i += blockDim.x) {
    const int n = i / kchunks, kc = i % kchunks;
    *reinterpret_cast<int4*>(Wt + (size_t)kc * (N * 16) + (size_t)n * 16) =
        *reinterpret_cast<const int4*>(W + (size_t)n * K + kc * 8);
  }
  for (int i = tid; i < 128 * kchunks; i += blockDim.x) {
    const int r = i / kchunks, kc = i % kchunks;
    int4 v = make_int4(0, 0, 0, 0);
    if (row0 + r < rows) v = *reinterpret_cast<const int4*>(A + (size_t)(row0 + r) * K + kc * 8);
    *reinterpret_cast<int4*>(At + (size_t)kc * (128 * 16) + (size_t)r * 16) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(At), b0 = smem_u32(Wt);
    for (int j = 0; j < (K >> 4); ++j) {
      const uint64_t ad = tc_smem_desc(a0 + j * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t bd = tc_smem_desc(b0 + j * 2 * (N * 16), N * 16, 128);
      tc_mma_f16(taddr, ad, bd, idesc, j > 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const long long row = row0 + warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tc_ld16(taddr + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    if (row < rows) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(Y + (size_t)row * N + c0 + 4 * q) =
            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(taddr, tmem_cols);
}

// Y[rows][N] (int32) = A[rows][K] (8-bit) * W[N][K]^T (int8): the exact-integer form used by the model kernels
// (hi plane: signed A, lo plane: unsigned A).
__global__ void __launch_bounds__(128) tc_gemm_i8_kernel(const unsigned char* __restrict__ A, const signed char* __restrict__ W,
                                                         int* __restrict__ Y, long long rows, int K, int N, int a_signed,
                                                         uint32_t tmem_cols) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Wt = smem;                      // [N][K] canonical
  unsigned char* At = smem + (size_t)N * K;      // [128][K] canonical
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * 128;
  if (warp == 0) tc_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  const int kchunks = K >> 4;
  for (int i = tid; i < N * kchunks; i += blockDim.x) {
    const int n = i / kchunks, kc = i % kchunks;
    *reinterpret_cast<int4*>(Wt + (size_t)kc * (N * 16) + (size_t)n * 16) =
        *reinterpret_cast<const int4*>(W + (size_t)n * K + kc * 16);
  }
  for (int i = tid; i < 128 * kchunks; i += blockDim.x) {
    const int r = i / kchunks, kc = i % kchunks;
    int4 v = make_int4(0, 0, 0, 0);
    if (row0 + r < rows) v = *reinterpret_cast<const int4*>(A + (size_t)(row0 + r) * K + kc * 16);
    *reinterpret_cast<int4*>(At + (size_t)kc * (128 * 16) + (size_t)r * 16) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc_idesc_i8(128, N, a_signed != 0, true);
    const uint32_t a0 = smem_u32(At), b0 = smem_u32(Wt);
    for (int j = 0; j < (K >> 5); ++j) {
      const uint64_t ad = tc_smem_desc(a0 + j * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t bd = tc_smem_desc(b0 + j * 2 * (N * 16), N * 16, 128);
      tc_mma_i8(taddr, ad, bd, idesc, j > 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const long long row = row0 + warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tc_ld16(taddr + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    if (row < rows) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<int4*>(Y + (size_t)row * N + c0 + 4 * q) = make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(taddr, tmem_cols);
}

}  // namespace s5

extern "C" int s5b200_tc_gemm_i8(const void* A, const void* W, int32_t* Y, long long rows, int K, int N, int a_signed, void* stream) {
  using namespace s5;
  if (!A || !W || !Y) S5_FAIL(S5_E_BADARG, "tc_gemm_i8: null pointer");
  if (rows <= 0 || K < 32 || (K & 31) || N < 16 || (N & 15) || N > 256) S5_FAIL(S5_E_BADSHAPE, "tc_gemm_i8: rows=%lld K=%d N=%d (K multiple of 32, N multiple of 16, N <= 256)", rows, K, N);
  if (!aligned16(A) || !aligned16(W) || !aligned16(Y)) S5_FAIL(S5_E_BADARG, "tc_gemm_i8: pointers must be 16-byte aligned");
  const size_t smem = (size_t)(N + 128) * K;
  if (smem > 220 * 1024) S5_FAIL(S5_E_BADSHAPE, "tc_gemm_i8: tile does not fit in shared memory");
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  S5_CUDA(cudaFuncSetAttribute(tc_gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (rows + 127) / 128;
  ProfScope prof(S5_K_OTHER, (cudaStream_t)stream);
  tc_gemm_i8_kernel<<<(unsigned)tiles, 128, smem, (cudaStream_t)stream>>>((const unsigned char*)A, (const signed char*)W, Y, rows, K, N, a_signed, cols);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

extern "C" int s5b200_tc_gemm_f16(const void* A, const void* W, float* Y, long long rows, int K, int N, void* stream) {
  using namespace s5;
  if (!A || !W || !Y) S5_FAIL(S5_E_BADARG, "tc_gemm_f16: null pointer");
  if (rows <= 0 || K < 16 || (K & 15) || N < 16 || (N & 15) || N > 256) S5_FAIL(S5_E_BADSHAPE, "tc_gemm_f16: rows=%lld K=%d N=%d (K, N multiples of 16, N <= 256)", rows, K, N);
  if (!aligned16(A) || !aligned16(W) || !aligned16(Y)) S5_FAIL(S5_E_BADARG, "tc_gemm_f16: pointers must be 16-byte aligned");
  const size_t smem = (size_t)(N + 128) * K * 2;
  if (smem > 220 * 1024) S5_FAIL(S5_E_BADSHAPE, "tc_gemm_f16: tile does not fit in shared memory");
  uint32_t cols = 32;
  while (cols < (uint32_t)N) cols <<= 1;
  S5_CUDA(cudaFuncSetAttribute(tc_gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (rows + 127) / 128;
  ProfScope prof(S5_K_OTHER, (cudaStream_t)stream);
  tc_gemm_f16_kernel<<<(unsigned)tiles, 128, smem, (cudaStream_t)stream>>>((const __half*)A, (const __half*)W, Y, rows, K, N, cols);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
