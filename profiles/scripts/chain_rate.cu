// chain_rate.cu - cycles per step of the integer recurrence of fxpmodel.py:147-208 on ONE warp (32 states), in isolation:
// how far is scan_i32's chain role from what the pipes allow?  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int mulw(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

template <int MODE>
__global__ void chain(const int* bu, int* out, long long* cyc, int eA, int steps) {
  __shared__ int2 tile[32 * 32];
  __shared__ int2 tile2[32 * 32];
  int* plan = reinterpret_cast<int*>(tile2);
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(tile2) + (threadIdx.x & 31) * 8;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) tile[i] = make_int2(bu[i], bu[i + 1024]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32 * (MODE == 3 ? 2 : 1)) return;
  const int ar = 31000 + lane, ai = 4000 - lane;
  int xr = 0, xi = 0;
  const int2* c2 = tile + lane;
  const long long t0 = clock64();
  for (int it = 0; it < steps / 32; ++it) {
    if (MODE == 0) {            // registers only: the pure dependent chain
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int nr = (int)((unsigned)(mulw(ar, xr) >> eA) - (unsigned)(mulw(ai, xi) >> eA) + (unsigned)(t + it));
        const int ni = (int)((unsigned)(mulw(ar, xi) >> eA) + (unsigned)(mulw(ai, xr) >> eA) + (unsigned)(t - it));
        xr = nr; xi = ni;
      }
    } else {                    // 1: + LDS of the input and STS of the state per step (the kernel's loop); 2: LDS only
      int vr[32], vi[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) { const int2 v = c2[t * 32]; vr[t] = v.x; vi[t] = v.y; }
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int nr = (int)((unsigned)(mulw(ar, xr) >> eA) - (unsigned)(mulw(ai, xi) >> eA) + (unsigned)vr[t]);
        const int ni = (int)((unsigned)(mulw(ar, xi) >> eA) + (unsigned)(mulw(ai, xr) >> eA) + (unsigned)vi[t]);
        xr = nr; xi = ni; vr[t] = nr; vi[t] = ni;
      }
      if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int t = 0; t < 32; ++t) const_cast<int2*>(c2)[t * 32] = make_int2(vr[t], vi[t]);
      }
      if (MODE == 4) {          // a different buffer than the one the next iteration loads from
#pragma unroll
        for (int t = 0; t < 32; ++t) tile2[t * 32 + lane] = make_int2(vr[t], vi[t]);
      }
      if (MODE == 5) {          // stores fenced behind the whole chain of the tile
        asm volatile("" ::: "memory");
#pragma unroll
        for (int t = 0; t < 32; ++t) asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(saddr + t * 256), "r"(vr[t]), "r"(vi[t]) : "memory");
      }
      if (MODE == 6) {          // two 32-bit stores (planar)
#pragma unroll
        for (int t = 0; t < 32; ++t) { plan[t * 64 + lane] = vr[t]; plan[t * 64 + 32 + lane] = vi[t]; }
      }
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = xr + xi;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int *bu, *out; long long* cyc;
  cudaMalloc(&bu, 2048 * 4); cudaMalloc(&out, 64 * 4); cudaMalloc(&cyc, 8);
  cudaMemset(bu, 1, 2048 * 4);
  const int steps = 3744;
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
    chain<0><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("registers only        : %.1f cycles per step\n", (double)h / steps);
    chain<2><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("+ LDS per step        : %.1f cycles per step\n", (double)h / steps);
    chain<1><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("+ LDS + STS per step  : %.1f cycles per step\n", (double)h / steps);
    chain<3><<<1, 64>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("two chain warps / SM  : %.1f cycles per step\n", (double)h / steps);
    chain<4><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("STS to another buffer : %.1f cycles per step\n", (double)h / steps);
    chain<5><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("STS after the chain   : %.1f cycles per step\n", (double)h / steps);
    chain<6><<<1, 32>>>(bu, out, cyc, 14, steps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("2 x STS.32 planar     : %.1f cycles per step\n", (double)h / steps);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
