// fxp_prims.cuh - device restatement of the reference's fixed-point primitives
// (sparseRNNs/fxparray.py, cited per function).  int32 payloads, wrap-around arithmetic, arithmetic >>.
#pragma once

#include "common.cuh"

namespace s5 {

__host__ __device__ __forceinline__ int fx_sat(int v, int bits) {
  // fxp_clip, fxparray.py:346-357 (signed).  bits >= 32 cannot clip an int32.
  if (bits >= 32) return v;
  const int hi = (1 << (bits - 1)) - 1;
  const int lo = -hi - 1;
  return v < lo ? lo : (v > hi ? hi : v);
}
__host__ __device__ __forceinline__ int fx_addw(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__host__ __device__ __forceinline__ int fx_mulw(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__host__ __device__ __forceinline__ int fx_shl(int v, int n) { return (int)((unsigned)v << n); }

__host__ __device__ __forceinline__ int fx_rshift_round(int x, int r, int mode) {
  // fxp_rshift_round, fxparray.py:274-284
  if (mode == S5_ROUND_CEIL) return fx_addw(x, (1 << r) - 1) >> r;
  if (mode == S5_ROUND_HALF_UP) return fx_addw(x, 1 << (r - 1)) >> r;
  return x >> r;
}

__host__ __device__ __forceinline__ int fx_change_exp(int v, int bits, int e, int ne, int mode) {
  // fxp_change_exp, fxparray.py:310-326: shift, then saturate at the CURRENT width; no-op when equal
  if (ne == e) return v;
  v = (ne > e) ? fx_shl(v, ne - e) : fx_rshift_round(v, e - ne, mode);
  return fx_sat(v, bits);
}

__host__ __device__ __forceinline__ int fx_change_cfg(int v, int bits, int e, int nbits, int ne, int mode) {
  // fxp_change_cfg, fxparray.py:232-271 (signed -> signed)
  if (bits == nbits && e == ne) return v;
  v = fx_change_exp(v, bits, e, ne, mode);
  if (bits > nbits) v = fx_sat(v, nbits);
  return v;
}

#ifdef __CUDACC__

__device__ __forceinline__ float fx_to_float(int v, int e) {
  // FxpArray.to_float, fxparray.py:72-73: int32 -> float32 (RN) / 2**e (exact)
  return __int2float_rn(v) * exp2f((float)-e);
}

__device__ __forceinline__ int fx_ceil_log2(float v) {
  // int(ceil(log2(v))) on a float32 scalar with a correctly rounded float32 log2 (see oracle/fxp.py)
  return (int)ceilf((float)log2((double)v));
}

struct AddPlan {
  int res_exp, agg_bits, agg_exp;
};

// fxp_add(..., result_exp="compute_best"), fxparray.py:420-432: exponent bookkeeping from the three maxima
__device__ __forceinline__ AddPlan fx_plan_add(float m_sum, float m1, float m2, int e1, int e2, int result_bits) {
  AddPlan p;
  const int intbits = max(0, fx_ceil_log2(m_sum + 1e-6f));
  p.res_exp = result_bits - intbits - 1;
  const int agg_int = max(max(0, fx_ceil_log2(m1 + 1e-8f)), max(0, fx_ceil_log2(m2 + 1e-8f)));
  p.agg_exp = max(e1, e2);
  p.agg_bits = agg_int + p.agg_exp + 1;
  return p;
}

// fxparray.py:433-448 + clip :458-464
__device__ __forceinline__ int fx_add_cb_apply(int a, int a_bits, int a_exp, int b, int b_bits, int b_exp,
                                               const AddPlan& p, int result_bits) {
  const int ac = fx_change_cfg(a, a_bits, a_exp, max(p.agg_bits, a_bits), p.agg_exp, S5_ROUND_FLOOR);
  const int bc = fx_change_cfg(b, b_bits, b_exp, max(p.agg_bits, b_bits), p.agg_exp, S5_ROUND_FLOOR);
  int s = fx_addw(ac, bc);
  const int d = p.res_exp - p.agg_exp;
  if (d > 0) s = fx_shl(s, d);
  else if (d < 0) s = s >> (-d);
  return fx_sat(s, result_bits);
}

// fxp_mul(..., result_exp="compute_best"), fxparray.py:601-608
__device__ __forceinline__ int fx_plan_mul(float m_prod, int result_bits) {
  return result_bits - max(0, fx_ceil_log2(1e-6f + m_prod)) - 1;
}

// fxp_mul body, fxparray.py:619-635
__device__ __forceinline__ int fx_mul_apply(int a, int b, int rshift, int mode, int result_bits) {
  return fx_sat(fx_rshift_round(fx_mulw(a, b), rshift, mode), result_bits);
}

// fxp_relu on a complex value, fxpmodel.py:30-45 (float32 round trip of both parts, lexicographic order)
__device__ __forceinline__ void fx_relu_complex(int& re, int& im) {
  const float fr = (float)re, fi = (float)im;
  const bool keep = (fr > 0.f) || (fr == 0.f && fi > 0.f);
  re = keep ? __float2int_rz(fr) : 0;
  im = keep ? __float2int_rz(fi) : 0;
}

// FxpSigmoid.apply, fxpmodel.py:97-144.  xx is already at exponent x_exp.
__device__ __forceinline__ int fx_sigmoid(int xx, int x_exp, int y_exp, const int* lut8) {
  const int sign = 2 * (xx > 0 ? 1 : 0) - 1;
  const int ax = xx < 0 ? (int)(0u - (unsigned)xx) : xx;
  const int delta = 1 << x_exp;
  int ind = ax >> x_exp;
  ind = ind < 6 ? ind : 6;
  ind = ind < 0 ? 0 : ind;  // |INT_MIN| stays negative; unreachable for saturated 16-bit inputs
  const int mu = ax & (delta - 1);
  const int yy = fx_addw(fx_mulw(delta - mu, lut8[ind]) >> x_exp, fx_mulw(mu, lut8[ind + 1]) >> x_exp);
  return fx_addw(1 << (y_exp - 1), fx_mulw(sign, yy));
}

// block-wide max of a non-negative float, accumulated into a global slot with an integer atomic
__device__ __forceinline__ void block_max_to(unsigned* slot, float v, float* red /* [32] shared */) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) atomicMax(slot, __float_as_uint(v));
  }
}

#endif  // __CUDACC__
}  // namespace s5
