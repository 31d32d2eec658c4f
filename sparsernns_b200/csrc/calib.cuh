// calib.cuh - small element-wise / reduction kernels of the calibration pass (s5b200_model_calibrate, float_model.cu).
#pragma once
#include "common.cuh"

namespace s5 {
#ifdef __CUDACC__

// monotone map float -> uint32 (so that atomicMin / atomicMax on the images order the floats)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// Running min / max of x[offset + stride * i] (stride 2 = one part of an interleaved complex tensor), optionally of the
// BatchNorm image (x - sub[k]) * mul[k] + add[k] with k = column index (H columns).  out: [min, max] as ordered uint32.
__global__ void __launch_bounds__(256) calib_minmax_kernel(const float* __restrict__ x, long long n, int stride, int offset, int H,
                                                           const float* __restrict__ sub, const float* __restrict__ mul,
                                                           const float* __restrict__ add, uint32_t* out) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[offset + stride * i];
    if (sub) {
      const int k = (int)(i % H);
      v = (v - sub[k]) * mul[k] + add[k];
    }
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out, f2ord(lo));
    atomicMax(out + 1, f2ord(hi));
  }
}
__global__ void calib_init_kernel(uint32_t* mm, int sites) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < sites) { mm[2 * i] = 0xFFFFFFFFu; mm[2 * i + 1] = 0u; }
}
__global__ void calib_decode_kernel(uint32_t* mm, int n) {   // ordered images -> floats, in place
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = __float_as_uint(ord2f(mm[i]));
}

enum { CAL_RELU = 0, CAL_CRELU = 1, CAL_GELU = 2, CAL_SIGMOID = 3, CAL_GATE = 4, CAL_GATE_RELU = 5 };
// in-place element-wise steps between the observed tensors; CAL_GATE: a[i] = a[i] * b[i] + c[i]
__global__ void __launch_bounds__(256) calib_ew_kernel(int mode, float* a, const float* __restrict__ b, const float* __restrict__ c,
                                                       long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mode == CAL_CRELU) {                                   // n complex values: lexicographic maximum with 0 (ssm.py:152-159)
      float2 v = reinterpret_cast<float2*>(a)[i];
      const bool keep = v.x > 0.f || (v.x == 0.f && v.y > 0.f);
      reinterpret_cast<float2*>(a)[i] = keep ? v : make_float2(0.f, 0.f);
      continue;
    }
    float v = a[i];
    if (mode == CAL_RELU) v = fmaxf(v, 0.f);
    else if (mode == CAL_GELU) { const float k = 0.7978845608028654f; v = v * (0.5f * (1.0f + tanhf(k * (v + 0.044715f * (v * v * v))))); }
    else if (mode == CAL_SIGMOID) v = 1.0f / (1.0f + expf(-v));
    else { v = v * b[i] + c[i]; if (mode == CAL_GATE_RELU) v = fmaxf(v, 0.f); }
    a[i] = v;
  }
}

#endif
}  // namespace s5
