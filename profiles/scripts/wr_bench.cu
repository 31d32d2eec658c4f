#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(float4* out, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = make_float4(v, v, v, v);
}
__global__ void wr4(float* out, size_t n, float v) {   // 4-byte stores, warp-coalesced 128 B, rows of 257 floats (unaligned rows)
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i + 1] = v;
}
__global__ void rd(const float4* in, size_t n, float* sink) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  float a = 0;
  for (; i < n; i += st) { float4 v = in[i]; a += v.x + v.y + v.z + v.w; }
  if (a == 123.456f) *sink = a;
}
__global__ void cp(const float4* in, float4* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}
int main() {
  const size_t bytes = (size_t)2 << 30;
  float4 *a, *b; float* sink;
  cudaMalloc(&a, bytes + 64); cudaMalloc(&b, bytes + 64); cudaMalloc(&sink, 4);
  cudaMemset(a, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); wr<<<148 * 16, 512>>>(a, bytes / 16, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("write-only  STG.128: %.0f GB/s\n", bytes / ms / 1e6);
    cudaEventRecord(e0); wr4<<<148 * 16, 512>>>((float*)a, bytes / 4 - 8, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("write-only  STG.32 (misaligned 128 B per warp): %.0f GB/s\n", bytes / ms / 1e6);
    cudaEventRecord(e0); rd<<<148 * 16, 512>>>(a, bytes / 16, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("read-only   LDG.128: %.0f GB/s\n", bytes / ms / 1e6);
    cudaEventRecord(e0); cp<<<148 * 16, 512>>>(a, b, bytes / 16); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("copy (read + write bytes): %.0f GB/s\n", 2.0 * bytes / ms / 1e6);
  }
  return 0;
}
