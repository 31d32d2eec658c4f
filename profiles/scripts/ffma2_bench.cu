#include <cstdio>
#include <cuda_runtime.h>
// throughput of scalar FFMA / FADD / FMNMX vs packed FFMA2 / FADD2 on sm_100a, and a fake-quant like mix
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else if (MODE == 2) {   // scalar fake-quant: mul, max, min, add magic, sub magic, mul
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float t = fminf(fmaxf(x[i] * a, -32768.f), 32767.f);
        t = (t + 12582912.f) - 12582912.f;
        x[i] = t * b;
      }
    } else if (MODE == 3) {   // packed fake-quant: mul2, 2 x (max, min), add2 magic, fma2 (t - magic) * b
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float2 t = __fmul2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a));
        t.x = fminf(fmaxf(t.x, -32768.f), 32767.f);
        t.y = fminf(fmaxf(t.y, -32768.f), 32767.f);
        t = __fadd2_rn(t, make_float2(12582912.f, 12582912.f));
        float2 r = __ffma2_rn(t, make_float2(b, b), make_float2(-12582912.f * b, -12582912.f * b));
        x[i] = r.x; x[i + 1] = r.y;
      }
    } else if (MODE == 4) {   // FMNMX only
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fminf(fmaxf(x[i], a), b);
    } else if (MODE == 5) {   // FADD scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] + a;
    } else if (MODE == 6) {   // FADD2
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float2 r = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a));
        x[i] = r.x; x[i + 1] = r.y;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_elem, int warps_per_sm) {
  float* out;
  const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
  cudaMalloc(&out, blocks * threads * 4);
  const int iters = 20000;
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double elems = (double)blocks * threads * 16 * iters;
  printf("%-28s warps/SM %2d: %.3f ms, %.1f elements/clk/SM (at 1.965 GHz), %.1f scalar-op-equivalents/clk/SM\n", name, warps_per_sm, ms,
         elems / (ms * 1e-3) / 148 / 1.965e9, elems * ops_per_elem / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  for (int w : {16, 32}) {
    if (w == 16) {
      run<0>("FFMA scalar", 1, 16); run<1>("FFMA2 packed", 1, 16); run<5>("FADD scalar", 1, 16); run<6>("FADD2 packed", 1, 16);
      run<4>("FMNMX x2 (clamp)", 2, 16); run<2>("fake-quant scalar (6 ops)", 6, 16); run<3>("fake-quant packed", 6, 16);
    } else {
      run<0>("FFMA scalar", 1, 32); run<1>("FFMA2 packed", 1, 32); run<2>("fake-quant scalar (6 ops)", 6, 32); run<3>("fake-quant packed", 6, 32);
    }
  }
  return 0;
}
