// fxp_model.cu - integer-only S5 / NDNS forward, bit-exact with the reference's fixed-point model.
//
// Reference path (paths relative to the reference repo):
//   run_validation                    sparseRNNs/fxprun.py:63-77   (input fxp_from_fp FLOOR, model(fxp_x), to_float)
//   FxpRegressionModel / StackedEnc.  sparseRNNs/fxpmodel.py:1261-1271, 1431-1439
//   FxpSequenceLayer.forward          sparseRNNs/fxpmodel.py:1110-1161 (+ mult_gate :1075-1093)
//   FxpBatchNorm.forward              sparseRNNs/fxpmodel.py:890-944   (four "compute_best" ops)
//   FxpSSM.forward                    sparseRNNs/fxpmodel.py:610-794
//   FxpDense.forward                  sparseRNNs/fxpmodel.py:331-366
//   fxparray ops                      sparseRNNs/fxparray.py
//
// "compute_best" exponents depend on a float32 max over the WHOLE tensor, so the forward is a chain of
// kernels separated by grid-wide reductions.  No host round trip: maxima and exponents live in a small
// device state block; each kernel re-derives the exponent plans it needs in its prologue and re-applies the
// cheap element-wise BatchNorm chain from the block input instead of materialising the intermediates.
#include <cstdlib>
#include <algorithm>
#include <mutex>
#include <vector>

#include "fxp_chain.cuh"
#include "fxp_tc.cuh"
#include "tc_model.cuh"
#include "proj.cuh"
#include "scan.cuh"

namespace s5 {


constexpr size_t kSmemMaxF = 227 * 1024;
// ------------------------------------------------------------------------------------------- BN reduce
// Applies stages 1..upto of the BatchNorm chain to x and accumulates the float maxima stage upto+1 needs.
struct BnReduceParams {
  const int* X; long long n; int H;
  BnChain c; unsigned* fm; int* ex; int* status;
  int upto, next;  // next in {2, 3, 4}
};

__global__ void __launch_bounds__(256) fxp_bn_reduce_kernel(BnReduceParams p) {
  __shared__ BnPlan pl;
  __shared__ float red[32];
  if (threadIdx.x == 0) {
    bn_make_plan(pl, p.c, p.fm, p.ex, p.upto);
    if (blockIdx.x == 0) {
      p.ex[E_1] = pl.e1; p.ex[E_2] = pl.e2; p.ex[E_3] = pl.e3;
      if (pl.bad) atomicOr(p.status, 1);
    }
  }
  __syncthreads();
  float m0 = 0.f, m1 = 0.f;
  const int e_prev = p.upto == 1 ? pl.e1 : (p.upto == 2 ? pl.e2 : pl.e3);
  // to_float scale factors hoisted out of the element loop (powers of two: exact)
  const float s_prev = exp2f((float)-e_prev);
  const int* pv = p.next == 2 ? p.c.inv : (p.next == 3 ? p.c.scale : p.c.bias);
  const float s_par = exp2f((float)-(p.next == 2 ? p.c.e_inv : (p.next == 3 ? p.c.e_scale : p.c.e_bias)));
  const bool is_mul = p.next != 4;
  const int cgn = p.H >> 2;                                   // int4 column groups per row
  if ((p.H & 3) == 0 && cgn <= 256) {
    // Column-resident form: a thread owns four fixed columns and strides over rows, so the per-column parameters sit in
    // registers (no modulo, no parameter loads in the loop); consecutive threads still read consecutive int4.
    const int rpb = 256 / cgn;                                // rows per block and iteration
    const int cg = threadIdx.x % cgn, rl = threadIdx.x / cgn;
    if (rl < rpb) {
      int mm[4], iv[4], sc[4], bi[4];
      float g4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * cg + e;
        mm[e] = p.c.mmean[k]; iv[e] = p.c.inv[k];
        sc[e] = p.c.has_scale ? p.c.scale[k] : 0; bi[e] = p.c.has_bias ? p.c.bias[k] : 0;
        g4[e] = __int2float_rn(pv[k]) * s_par;
      }
      const long long rows = p.n / p.H;
      const int4* X4 = reinterpret_cast<const int4*>(p.X);
      for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
        const int4 xv = X4[r * cgn + cg];
        const int xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int v = bn_apply_v(xs4[e], mm[e], iv[e], sc[e], bi[e], p.c, pl, p.upto);
          const float f = __int2float_rn(v) * s_prev;
          m0 = fmaxf(m0, fabsf(is_mul ? f * g4[e] : f + g4[e]));
          m1 = fmaxf(m1, fabsf(f));
        }
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
      const int k = (int)(i % p.H);
      const int v = bn_apply(p.X[i], k, p.c, pl, p.upto);
      const float f = __int2float_rn(v) * s_prev;
      const float g = __int2float_rn(pv[k]) * s_par;
      m0 = fmaxf(m0, fabsf(is_mul ? f * g : f + g));
      m1 = fmaxf(m1, fabsf(f));
    }
  }
  if (p.next == 2) block_max_to(p.fm + M_2, m0, red);
  else if (p.next == 3) block_max_to(p.fm + M_3, m0, red);
  else { block_max_to(p.fm + M_S4, m0, red); block_max_to(p.fm + M_A4, m1, red); }
}

// ------------------------------------------------------------------------------------------- BN maxima from column extremes
// Every stage of the chain is a monotone function of the block input for a fixed column (shifts, clamps, a product with a
// per-column constant, a sum with a per-column constant), and so are the float images the "compute_best" maxima are taken of
// (int -> float RN, times a power of two, times / plus a per-column constant, abs).  The maximum over all rows of a column is
// therefore attained at the column's smallest or largest input: ONE CTA derives all six maxima of a block from the 2 H column
// extremes the producing kernel left behind, instead of three full passes over the tensor (0.29 of 1.43 ms per forward).
// Monotonicity needs wrap-free integer steps; every step is re-done in 64 bits and, should one wrap (never seen with
// w8a16 payloads), this CTA falls back to the plain passes over the whole tensor - exact either way, only slow.
struct BnDeriveParams {
  const int* X; long long n; int H;
  BnChain c; unsigned* fm; int* ex; int* status;
  const unsigned* cmax; const unsigned* cmin;
  int force_slow;
};

__device__ __forceinline__ float block_allmax(float v, float* red) {   // every thread gets the block-wide maximum
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = 0.f;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) m = fmaxf(m, red[i]);
  return m;
}
__device__ __forceinline__ int sc_apply_chk(int v, const ShiftClamp& s, int& wrap) {
  const long long t = (long long)v << s.shl;
  if (t != (long long)(int)t) wrap = 1;
  return sc_apply(v, s);
}
__device__ __forceinline__ int add_fast_chk(int a, int b, const AddFast& f, int& wrap) {
  const int a1 = sc_apply_chk(a, f.a, wrap), b1 = sc_apply_chk(b, f.b, wrap);
  const long long s = (long long)a1 + (long long)b1;
  if (s != (long long)(int)s) wrap = 1;
  return sc_apply_chk((int)s, f.post, wrap);
}
__device__ __forceinline__ int mul_chk(int a, int b, int rs, int lo, int hi, int& wrap) {
  const long long t = (long long)a * (long long)b;
  if (t != (long long)(int)t) wrap = 1;
  return min(max(fx_mulw(a, b) >> rs, lo), hi);
}

__global__ void __launch_bounds__(1024) fxp_bn_derive_kernel(BnDeriveParams p) {
  __shared__ BnPlan pl;
  __shared__ float red[32];
  // the maxima and the input exponent are staged in shared memory: thread 0 plans from there (four plans per launch) and the
  // global state block is written once at the end - no global round trip between the stages
  __shared__ unsigned fml[kStateStride];
  __shared__ int exl[kStateStride];
  if (threadIdx.x < kStateStride) { fml[threadIdx.x] = 0u; exl[threadIdx.x] = threadIdx.x == E_IN ? p.ex[E_IN] : 0; }
  __syncthreads();
  const int k = threadIdx.x;
  const bool act = k < p.H;
  const BnChain& c = p.c;
  int wrap = p.force_slow;
  int lo = 0, hi = 0, mm = 0, iv = 0, scl = 0, bia = 0;
  if (act) {
    hi = ord_dec(p.cmax[k]);
    lo = ~ord_dec(p.cmin[k]);
    mm = c.mmean[k]; iv = c.inv[k];
    scl = c.has_scale ? c.scale[k] : 0; bia = c.has_bias ? c.bias[k] : 0;
  }
  const int e_in = exl[E_IN];
  {  // maxima of the first add: |x + (-mean)| and |x|
    const float s = exp2f((float)-e_in), g = __int2float_rn(mm) * exp2f((float)-c.e_mean);
    const float fl = __int2float_rn(lo) * s, fh = __int2float_rn(hi) * s;
    const float mS = block_allmax(act ? fmaxf(fabsf(fl + g), fabsf(fh + g)) : 0.f, red);
    const float mA = block_allmax(act ? fmaxf(fabsf(fl), fabsf(fh)) : 0.f, red);
    if (k == 0) {
      fml[M_S1] = __float_as_uint(mS); fml[M_A1] = __float_as_uint(mA);
      bn_make_plan(pl, c, fml, exl, 1);
    }
    __syncthreads();
  }
  int bad = pl.bad;
  {  // stage 1 -> maxima of the product with 1 / sqrt(var + eps)
    const int a = add_fast_chk(lo, mm, pl.f.s1, wrap), b = add_fast_chk(hi, mm, pl.f.s1, wrap);
    lo = min(a, b); hi = max(a, b);
    const float s = exp2f((float)-pl.e1), g = __int2float_rn(iv) * exp2f((float)-c.e_inv);
    const float m = block_allmax(act ? fmaxf(fabsf(__int2float_rn(lo) * s * g), fabsf(__int2float_rn(hi) * s * g)) : 0.f, red);
    if (k == 0) {
      fml[M_2] = __float_as_uint(m);
      bn_make_plan(pl, c, fml, exl, 2);
    }
    __syncthreads();
  }
  bad |= pl.bad;
  {  // stage 2
    const int a = mul_chk(lo, iv, pl.f.rs2, pl.f.lo2, pl.f.hi2, wrap), b = mul_chk(hi, iv, pl.f.rs2, pl.f.lo2, pl.f.hi2, wrap);
    lo = min(a, b); hi = max(a, b);
  }
  if (c.has_scale) {
    const float s = exp2f((float)-pl.e2), g = __int2float_rn(scl) * exp2f((float)-c.e_scale);
    const float m = block_allmax(act ? fmaxf(fabsf(__int2float_rn(lo) * s * g), fabsf(__int2float_rn(hi) * s * g)) : 0.f, red);
    if (k == 0) {
      fml[M_3] = __float_as_uint(m);
      bn_make_plan(pl, c, fml, exl, 3);
    }
    __syncthreads();
    bad |= pl.bad;
    const int a = mul_chk(lo, scl, pl.f.rs3, pl.f.lo3, pl.f.hi3, wrap), b = mul_chk(hi, scl, pl.f.rs3, pl.f.lo3, pl.f.hi3, wrap);
    lo = min(a, b); hi = max(a, b);
  }
  if (c.has_bias) {
    const float s = exp2f((float)-pl.e3), g = __int2float_rn(bia) * exp2f((float)-c.e_bias);
    const float fl = __int2float_rn(lo) * s, fh = __int2float_rn(hi) * s;
    const float mS = block_allmax(act ? fmaxf(fabsf(fl + g), fabsf(fh + g)) : 0.f, red);
    const float mA = block_allmax(act ? fmaxf(fabsf(fl), fabsf(fh)) : 0.f, red);
    if (k == 0) {
      fml[M_S4] = __float_as_uint(mS); fml[M_A4] = __float_as_uint(mA);
      bn_make_plan(pl, c, fml, exl, 4);
    }
    __syncthreads();
    // the second add is only range-checked here; its result is not needed for a maximum
    if (act) { (void)add_fast_chk(lo, bia, pl.f.s4, wrap); (void)add_fast_chk(hi, bia, pl.f.s4, wrap); }
  }
  wrap = __syncthreads_or(act ? wrap : p.force_slow);
  if (k < kStateStride && k != M_SR && k != M_G) p.fm[k] = fml[k];   // M_SR / M_G belong to the GLU kernel of this block
  __syncthreads();
  if (wrap) {
    // plain passes over the whole tensor by this one CTA (the arithmetic of fxp_bn_reduce_kernel)
    for (int upto = 1; upto <= 3; ++upto) {
      const int next = upto + 1;
      if ((next == 3 && !c.has_scale) || (next == 4 && !c.has_bias)) continue;
      if (k == 0) {
        if (next == 2) p.fm[M_2] = 0u; else if (next == 3) p.fm[M_3] = 0u; else { p.fm[M_S4] = 0u; p.fm[M_A4] = 0u; }
        bn_make_plan(pl, c, p.fm, p.ex, upto);
      }
      __syncthreads();
      const int e_prev = upto == 1 ? pl.e1 : (upto == 2 ? pl.e2 : pl.e3);
      const float s_prev = exp2f((float)-e_prev);
      const int* pv = next == 2 ? c.inv : (next == 3 ? c.scale : c.bias);
      const float s_par = exp2f((float)-(next == 2 ? c.e_inv : (next == 3 ? c.e_scale : c.e_bias)));
      float m0 = 0.f, m1 = 0.f;
      for (long long i = threadIdx.x; i < p.n; i += blockDim.x) {
        const int kk = (int)(i % p.H);
        const float f = __int2float_rn(bn_apply(p.X[i], kk, c, pl, upto)) * s_prev, g = __int2float_rn(pv[kk]) * s_par;
        m0 = fmaxf(m0, fabsf(next != 4 ? f * g : f + g));
        m1 = fmaxf(m1, fabsf(f));
      }
      m0 = block_allmax(m0, red); m1 = block_allmax(m1, red);
      if (k == 0) {
        if (next == 2) p.fm[M_2] = __float_as_uint(m0); else if (next == 3) p.fm[M_3] = __float_as_uint(m0);
        else { p.fm[M_S4] = __float_as_uint(m0); p.fm[M_A4] = __float_as_uint(m1); }
      }
      __syncthreads();
    }
    if (k == 0) bn_make_plan(pl, c, p.fm, p.ex, 4);
    __syncthreads();
    bad = pl.bad;
  }
  if (k == 0) {
    p.ex[E_1] = pl.e1; p.ex[E_2] = pl.e2; p.ex[E_3] = pl.e3;
    if (bad | pl.bad) atomicOr(p.status, 1);
  }
}

// Materialises u = change_cfg(norm_output) (the "pre_s5" tap).
struct BnApplyParams { const int* X; int* U; long long n; int H; BnChain c; unsigned* fm; int* ex; };
__global__ void __launch_bounds__(256) fxp_bn_apply_kernel(BnApplyParams p) {
  __shared__ BnPlan pl;
  if (threadIdx.x == 0) bn_make_plan(pl, p.c, p.fm, p.ex, 4);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % p.H);
    const int v = bn_apply(p.X[i], k, p.c, pl, 4);
    p.U[i] = bn_to_u(v, pl);
  }
}

// the integer model is always ReLU-ified (fxpmodel.py:1316): C projection, out2 and decoder inputs are post-ReLU
static bool fxp_skip_zero() {
  static const bool on = [] { const char* e = getenv("S5B200_SKIP_ZERO"); return e && atoi(e) != 0; }();   // off by default: exact but measured slower on B200
  return on;
}

// ------------------------------------------------------------------------------------------- linear kernels

struct FxpLinearParams {
  const void* X; int* Y; float* Yf; long long rows; WView w;
  int mode;
  // ENC: float input quantised with (in_bits, in_exp) FLOOR when x_is_float; int input otherwise
  int x_is_float, in_bits, in_exp;
  // matmul epilogue: rshift (even/odd output column), round mode, saturation width
  int rshift0, rshift1, round_mode, out_bits;
  const int* bias_aligned;  // [N] bias already at out_exp (FxpDense, fxpmodel.py:358-364) or null
  int relu;
  float out_scale;          // 2^-out_exp for Yf
  // BPROJ: BatchNorm chain of this block
  BnChain c; unsigned* fm; int* ex; int* status;
  // ENC / RESID epilogue: maxima for the next block's first BatchNorm stage
  const int* next_mmean; int next_e_mean; unsigned* next_fm; int out_exp;
  unsigned* next_cmax; unsigned* next_cmin;   // dense-tile encoder only: column extremes instead of the float maxima
  unsigned short* U16;                        // dense-tile B projection only: 16-bit copy of u for the GLU kernel
  // DEC: dynamic input exponent (ex[E_IN]) versus the static FxpDense input config
  int dec_inp_bits, dec_inp_exp, dec_w_exp, dec_out_exp;
  int skip_zero;   // proj.cuh:zero_group_mask (exact: a zero activation contributes 0 to the wrapping int32 sum)
};

template <int TR>
__global__ void __launch_bounds__(kProjThreads) fxp_linear_kernel(FxpLinearParams p) {
  constexpr int R = TR / 32;
  extern __shared__ __align__(16) int smem_i[];
  __shared__ BnPlan pl;
  __shared__ float red[32];
  __shared__ int dyn[4];
  const int K = p.w.K, N = p.w.N;
  size_t ww = (wsmem_words(p.w) + 1) & ~(size_t)1;
  const WSmem w = wsmem_load(smem_i, p.w);
  int* xt = smem_i + ww;
  int* yt = xt + (size_t)K * (TR + 1);
  uint32_t* zmask = p.skip_zero ? reinterpret_cast<uint32_t*>(yt + (size_t)N * (TR + 1)) : nullptr;   // [K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (p.mode == LIN_BPROJ) {
      bn_make_plan(pl, p.c, p.fm, p.ex, 4);
      if (blockIdx.x == 0) {
        p.ex[E_1] = pl.e1; p.ex[E_2] = pl.e2; p.ex[E_3] = pl.e3; p.ex[E_4] = pl.e4;
        if (pl.bad) atomicOr(p.status, 1);
      }
    } else if (p.mode == LIN_DEC) {
      // FxpDense.forward input check (fxpmodel.py:335-347) with a data-dependent input exponent
      const int e = p.ex[E_IN];
      const int change = (16 > p.dec_inp_bits) || (e > p.dec_inp_exp);
      const int e_eff = change ? p.dec_inp_exp : e;
      const int rs = e_eff + p.dec_w_exp - p.dec_out_exp;
      dyn[0] = change; dyn[1] = e; dyn[2] = rs & 31;
      if ((rs < 0 || rs > 31 || (change && e - p.dec_inp_exp > 31)) && blockIdx.x == 0) atomicOr(p.status, 1);
    }
  }
  __syncthreads();
  float m0 = 0.f, m1 = 0.f;
  const long long ntiles = (p.rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int rv = (int)min((long long)TR, p.rows - row0);
    if (p.mode == LIN_ENC && p.x_is_float) {
      const float sc = exp2f((float)p.in_exp);
      stage_rows<float, int, TR>(xt, reinterpret_cast<const float*>(p.X) + row0 * K, rv, K, [&](float v, int) {
        return fx_sat(__float2int_rd(v * sc), p.in_bits);  // fxp_from_fp FLOOR, fxprun.py:69-75
      });
    } else if (p.mode == LIN_BPROJ) {
      stage_rows<int, int, TR>(xt, reinterpret_cast<const int*>(p.X) + row0 * K, rv, K, [&](int v, int k) {
        v = bn_apply(v, k, p.c, pl, 4);
        return bn_to_u(v, pl);  // fxpmodel.py:620-624
      });
    } else if (p.mode == LIN_DEC) {
      const int change = dyn[0], e = dyn[1];
      stage_rows<int, int, TR>(xt, reinterpret_cast<const int*>(p.X) + row0 * K, rv, K, [&](int v, int) {
        return change ? fx_change_cfg(v, 16, e, p.dec_inp_bits, p.dec_inp_exp, S5_ROUND_FLOOR) : v;
      });
    } else {
      stage_rows<int, int, TR>(xt, reinterpret_cast<const int*>(p.X) + row0 * K, rv, K, [](int v, int) { return v; });
    }
    __syncthreads();
    if (zmask) {
      zero_group_mask<int, R, TR>(xt, K, zmask);
      __syncthreads();
    }
    const int rs_dyn = p.mode == LIN_DEC ? dyn[2] : 0;
    for (int n = warp; n < N; n += kProjWarps) {
      int acc[R];
      column_dot<int, R, TR>(w, K, n, xt, lane, acc, zmask);
      const int rs = p.mode == LIN_DEC ? rs_dyn : ((n & 1) ? p.rshift1 : p.rshift0);
      const int bn = p.bias_aligned ? p.bias_aligned[n] : 0;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        int v = fx_sat(fx_rshift_round(acc[j], rs, p.round_mode), p.out_bits);  // fxp_matmul, fxparray.py:662-676
        if (p.bias_aligned) v = fx_sat(fx_addw(v, bn), p.out_bits);             // fxp_add at out_exp
        if (p.relu) v = v > 0 ? v : 0;
        yt[(size_t)n * (TR + 1) + lane + 32 * j] = v;
      }
    }
    __syncthreads();
    if (p.next_fm) {
      const float so = exp2f((float)-p.out_exp), sm = exp2f((float)-p.next_e_mean);
      unstage_rows<int, int, TR>(p.Y + row0 * N, yt, rv, N, [&](int v, int, int n) {
        const float f = __int2float_rn(v) * so;
        m0 = fmaxf(m0, fabsf(f + __int2float_rn(p.next_mmean[n]) * sm));
        m1 = fmaxf(m1, fabsf(f));
        return v;
      });
    } else {
      if (p.Y) unstage_rows<int, int, TR>(p.Y + row0 * N, yt, rv, N, [](int v, int, int) { return v; });
      if (p.Yf) {
        const float so = p.out_scale;
        unstage_rows<int, float, TR>(p.Yf + row0 * N, yt, rv, N, [=](int v, int, int) { return __int2float_rn(v) * so; });
      }
    }
  }
  if (p.next_fm) {
    block_max_to(p.next_fm + M_S1, m0, red);
    block_max_to(p.next_fm + M_A1, m1, red);
  }
}

static inline size_t wwords_h(const WView& s) {
  size_t w = s.dense ? (size_t)s.K * s.N : (size_t)s.nnz * 2 + (size_t)(s.N + 1);
  return (w + 1) & ~(size_t)1;
}
static int pick_tr_f(size_t fixed_bytes, size_t per_row_bytes) {
  const int cands[3] = {128, 64, 32};
  for (int tr : cands)
    if (fixed_bytes + per_row_bytes * (tr + 1) <= kSmemMaxF / 2 - 2048) return tr;
  for (int tr : cands)
    if (fixed_bytes + per_row_bytes * (tr + 1) <= kSmemMaxF - 1024) return tr;
  return 0;
}
template <typename KernelT>
static int pgrid(KernelT kernel, size_t smem, long long ntiles, int* grid_out) {
  int per_sm = 0;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kProjThreads, smem));
  if (per_sm < 1) S5_FAIL(S5_E_BADSHAPE, "kernel does not fit on an SM with %zu bytes of shared memory", smem);
  *grid_out = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)per_sm * sm_count()));
  return S5_OK;
}

static int launch_fxp_linear_tc(const FxpLinearParams& p, const FxpTcLinearDev& tc, cudaStream_t st, int kclass) {
  FxpTcLinArgs a{};
  a.X = p.X; a.Y = p.Y; a.Yf = p.Yf; a.rows = p.rows; a.mode = p.mode;
  a.x_is_float = p.x_is_float; a.in_bits = p.in_bits; a.in_exp = p.in_exp;
  a.rshift0 = p.rshift0; a.rshift1 = p.rshift1; a.out_bits = p.out_bits; a.bias_aligned = p.bias_aligned; a.relu = p.relu;
  a.out_scale = p.out_scale; a.c = p.c; a.fm = p.fm; a.ex = p.ex; a.status = p.status;
  a.next_mmean = p.next_mmean; a.next_e_mean = p.next_e_mean; a.next_fm = p.next_fm; a.out_exp = p.out_exp;
  a.next_cmax = p.next_cmax; a.next_cmin = p.next_cmin; a.U16 = p.U16;
  a.dec_inp_bits = p.dec_inp_bits; a.dec_inp_exp = p.dec_inp_exp; a.dec_w_exp = p.dec_w_exp; a.dec_out_exp = p.dec_out_exp;
  return fxp_tc_linear_launch(tc, a, kclass, st);
}

static int launch_fxp_linear(const FxpLinearParams& p, cudaStream_t st, int kclass = S5_K_OTHER) {
  if (p.rows <= 0) S5_FAIL(S5_E_BADSHAPE, "fxp linear: rows=%lld must be positive", p.rows);
  const size_t wbytes = wwords_h(p.w) * 4;
  const size_t zbytes = p.skip_zero ? (size_t)p.w.K * 4 : 0;
  const int tr = pick_tr_f(wbytes + zbytes, (size_t)(p.w.K + p.w.N) * 4);
  if (!tr) S5_FAIL(S5_E_BADSHAPE, "fxp linear: K=%d N=%d weights do not fit in shared memory", p.w.K, p.w.N);
  const size_t smem = wbytes + (size_t)(p.w.K + p.w.N) * (tr + 1) * 4 + zbytes;
  const long long ntiles = (p.rows + tr - 1) / tr;
  int grid = 1, rc;
  ProfScope prof(kclass, st);
  switch (tr) {
    case 128:
      if ((rc = pgrid(fxp_linear_kernel<128>, smem, ntiles, &grid))) return rc;
      fxp_linear_kernel<128><<<grid, kProjThreads, smem, st>>>(p);
      break;
    case 64:
      if ((rc = pgrid(fxp_linear_kernel<64>, smem, ntiles, &grid))) return rc;
      fxp_linear_kernel<64><<<grid, kProjThreads, smem, st>>>(p);
      break;
    default:
      if ((rc = pgrid(fxp_linear_kernel<32>, smem, ntiles, &grid))) return rc;
      fxp_linear_kernel<32><<<grid, kProjThreads, smem, st>>>(p);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// ------------------------------------------------------------------------------------------- C-proj + GLU
struct FxpGluParams {
  const int* XS;  // [rows][2P] interleaved (re, im), ReLU applied
  const int* X;   // [rows][H] block input
  int* G;         // [rows][H] post-GLU
  long long rows;
  WView wcr, wci, wo;  // C_re / C_im (K = 2P rows, only even / odd k populated), out2
  BnChain c; unsigned* fm; int* ex; int* status;
  const int* D; int rs_cre, rs_cim, rs_D, y_bits, conj;
  // out2 FxpDense
  int o_in_change, y_exp, o_inp_bits, o_inp_exp, o_rs, o_out_bits, o_out_exp; const int* o_bias_aligned;
  // sigmoid + gate
  int sig_x_exp, sig_y_exp; int lut[8];
  int gl_bits, gl_exp, gr_bits, gr_exp, g_rs, g_bits, g_exp;
  int skip_zero;
  // fix-up pass behind the dense-tile kernel (fxp_tc.cu): only the listed row ranges (row0, count <= 128), i.e. the tiles
  // whose states do not fit 16 bits; the count lives in device memory, a launch with an empty list returns at once
  const int* bad_list; const int* bad_count;
};

template <int TR>
__global__ void __launch_bounds__(kProjThreads) fxp_glu_kernel(FxpGluParams p) {
  constexpr int R = TR / 32;
  constexpr int LD = TR + 1;
  constexpr int SUB = (128 + TR - 1) / TR;                      // sub-tiles per listed range
  const int nbad = p.bad_list ? *p.bad_count : 0;
  if (p.bad_list && nbad == 0) return;
  extern __shared__ __align__(16) int smem_i[];
  __shared__ BnPlan pl;
  __shared__ float red[32];
  __shared__ int lut_s[8];
  const int K1 = p.wcr.K, H = p.wcr.N;
  const size_t w1 = (wsmem_words(p.wcr) + 1) & ~(size_t)1;
  const size_t w2 = (wsmem_words(p.wci) + 1) & ~(size_t)1;
  const size_t w3 = (wsmem_words(p.wo) + 1) & ~(size_t)1;
  const WSmem wcr = wsmem_load(smem_i, p.wcr);
  const WSmem wci = wsmem_load(smem_i + w1, p.wci);
  const WSmem wo = wsmem_load(smem_i + w1 + w2, p.wo);
  const int rowsA = K1 > H ? K1 : H;
  int* bufA = smem_i + w1 + w2 + w3;       // xs tile, later g
  int* bufB = bufA + (size_t)rowsA * LD;   // block input tile, later out2 input
  int* bufC = bufB + (size_t)H * LD;       // x1
  uint32_t* zmask = p.skip_zero ? reinterpret_cast<uint32_t*>(bufC + (size_t)H * LD) : nullptr;   // [max(K1, H)]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) bn_make_plan(pl, p.c, p.fm, p.ex, 4);
  if (threadIdx.x < 8) lut_s[threadIdx.x] = p.lut[threadIdx.x];
  __syncthreads();
  const float sg = exp2f((float)-p.g_exp), sk = exp2f((float)-pl.e_in);
  float m0 = 0.f, m1 = 0.f;
  const long long ntiles = p.bad_list ? (long long)nbad * SUB : (p.rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long row0 = tile * TR;
    int rv = (int)min((long long)TR, p.rows - row0);
    if (p.bad_list) {
      const int ent = (int)(tile / SUB), sub = (int)(tile % SUB);
      row0 = (long long)p.bad_list[2 * ent] + (long long)sub * TR;
      rv = min(TR, p.bad_list[2 * ent + 1] - sub * TR);
      if (rv <= 0) continue;                                    // CTA-uniform
    }
    __syncthreads();
    stage_rows<int, int, TR>(bufA, p.XS + row0 * K1, rv, K1, [](int v, int) { return v; });
    stage_rows<int, int, TR>(bufB, p.X + row0 * H, rv, H, [&](int v, int k) {
      v = bn_apply(v, k, p.c, pl, 4);
      return bn_to_u(v, pl);  // u, fxpmodel.py:620-624
    });
    __syncthreads();
    if (zmask) {
      zero_group_mask<int, R, TR>(bufA, K1, zmask);
      __syncthreads();
    }
    for (int n = warp; n < H; n += kProjWarps) {
      int a1[R], a2[R];
      column_dot<int, R, TR>(wcr, K1, n, bufA, lane, a1, zmask);
      column_dot<int, R, TR>(wci, K1, n, bufA, lane, a2, zmask);
      const int dn = p.D[n];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const size_t o = (size_t)n * LD + lane + 32 * j;
        const int y1 = fx_sat(a1[j] >> p.rs_cre, p.y_bits);                  // fxpmodel.py:747-753
        const int y2 = fx_sat(a2[j] >> p.rs_cim, p.y_bits);                  // :754-760
        int ys = fx_sat(fx_addw(y1, fx_mulw(y2, -1)), p.y_bits);              // fxp_sub :746-763
        if (p.conj) ys = fx_mulw(ys, 2);                                      // :765-767, after the clip
        const int du = fx_mul_apply(dn, bufB[o], p.rs_D, S5_ROUND_FLOOR, p.y_bits);  // :771-776
        const int y = fx_sat(fx_addw(ys, du), p.y_bits);                      // :778-783
        const int x1 = y > 0 ? y : 0;                                         // fxp_relu :1125
        bufC[o] = x1;
        bufB[o] = p.o_in_change ? fx_change_cfg(x1, p.y_bits, p.y_exp, p.o_inp_bits, p.o_inp_exp, S5_ROUND_FLOOR) : x1;
      }
    }
    __syncthreads();
    if (zmask) {
      zero_group_mask<int, R, TR>(bufB, H, zmask);
      __syncthreads();
    }
    for (int n = warp; n < H; n += kProjWarps) {
      int acc[R];
      column_dot<int, R, TR>(wo, H, n, bufB, lane, acc, zmask);
      const int bn = p.o_bias_aligned[n];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const size_t o = (size_t)n * LD + lane + 32 * j;
        int ov = fx_sat(acc[j] >> p.o_rs, p.o_out_bits);
        ov = fx_sat(fx_addw(ov, bn), p.o_out_bits);                           // FxpDense :352-364
        const int xx = fx_change_exp(ov, p.o_out_bits, p.o_out_exp, p.sig_x_exp, S5_ROUND_FLOOR);
        const int r = fx_sigmoid(xx, p.sig_x_exp, p.sig_y_exp, lut_s);        // bits = o_out_bits, exp = sig_y_exp
        const int xg = fx_change_cfg(bufC[o], p.y_bits, p.y_exp, p.gl_bits, p.gl_exp, S5_ROUND_FLOOR);
        const int rg = fx_change_cfg(r, p.o_out_bits, p.sig_y_exp, p.gr_bits, p.gr_exp, S5_ROUND_FLOOR);
        bufA[o] = fx_mul_apply(xg, rg, p.g_rs, S5_ROUND_FLOOR, p.g_bits);     // mult_gate :1075-1093
      }
    }
    __syncthreads();
    const int* skip = p.X + row0 * H;
    unstage_rows<int, int, TR>(p.G + row0 * H, bufA, rv, H, [&](int g, int r, int n) {
      const float fg = __int2float_rn(g) * sg;
      m0 = fmaxf(m0, fabsf(fg + __int2float_rn(skip[(size_t)r * H + n]) * sk));
      m1 = fmaxf(m1, fabsf(fg));
      return g;
    });
  }
  block_max_to(p.fm + M_SR, m0, red);
  block_max_to(p.fm + M_G, m1, red);
}

static int launch_fxp_glu(const FxpGluParams& p, cudaStream_t st) {
  const int K1 = p.wcr.K, H = p.wcr.N;
  const size_t wbytes = (wwords_h(p.wcr) + wwords_h(p.wci) + wwords_h(p.wo)) * 4;
  const size_t per_row = (size_t)(std::max(K1, H) + 2 * H) * 4;
  const size_t zbytes = p.skip_zero ? (size_t)std::max(K1, H) * 4 : 0;
  const int tr = pick_tr_f(wbytes + zbytes, per_row);
  if (!tr) S5_FAIL(S5_E_BADSHAPE, "fxp glu: H=%d 2P=%d weights do not fit in shared memory", H, K1);
  const size_t smem = wbytes + per_row * (tr + 1) + zbytes;
  // fix-up mode: the number of listed ranges is only known on the device; one wave of CTAs walks the list
  const long long ntiles = p.bad_list ? (long long)sm_count() * 4 : (p.rows + tr - 1) / tr;
  int grid = 1, rc;
  ProfScope prof(S5_K_FXP_GLU, st);
  switch (tr) {
    case 128:
      if ((rc = pgrid(fxp_glu_kernel<128>, smem, ntiles, &grid))) return rc;
      fxp_glu_kernel<128><<<grid, kProjThreads, smem, st>>>(p);
      break;
    case 64:
      if ((rc = pgrid(fxp_glu_kernel<64>, smem, ntiles, &grid))) return rc;
      fxp_glu_kernel<64><<<grid, kProjThreads, smem, st>>>(p);
      break;
    default:
      if ((rc = pgrid(fxp_glu_kernel<32>, smem, ntiles, &grid))) return rc;
      fxp_glu_kernel<32><<<grid, kProjThreads, smem, st>>>(p);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// ------------------------------------------------------------------------------------------- residual
// out = relu(fxp_add(g, skip, "compute_best", res_bits))  (fxpmodel.py:1147-1159) + maxima for the next block
struct FxpResidParams {
  const int* G; const int* X; int* OUT; long long n; int H;
  unsigned* fm; int* ex; int* status;     // this block's state
  int g_bits, g_exp, x_bits, res_bits;
  const int* next_mmean; int next_e_mean; unsigned* next_fm; int* next_ex;  // next block (or decoder: only next_ex)
  unsigned* next_cmax; unsigned* next_cmin;   // column extremes of the output instead of float maxima (fxp_bn_derive_kernel)
};

__global__ void __launch_bounds__(256) fxp_resid_kernel(FxpResidParams p) {
  __shared__ AddPlan pl;
  __shared__ float red[32];
  __shared__ int e_in_s;
  __shared__ unsigned scmax[1024], scmin[1024];                 // H <= 1024 (fxp_model_create)
  const bool cstat = p.next_cmax != nullptr;
  if (cstat)
    for (int i = threadIdx.x; i < p.H; i += blockDim.x) { scmax[i] = 0u; scmin[i] = 0u; }
  if (threadIdx.x == 0) {
    const int e_in = p.ex[E_IN];
    pl = fx_plan_add(__uint_as_float(p.fm[M_SR]), __uint_as_float(p.fm[M_G]), __uint_as_float(p.fm[M_A1]), p.g_exp, e_in,
                     p.res_bits);
    e_in_s = e_in;
    if (blockIdx.x == 0) {
      p.ex[E_OUT] = pl.res_exp;
      p.next_ex[E_IN] = pl.res_exp;
      if (pl.agg_bits > 32 || pl.agg_exp - p.g_exp > 31 || pl.agg_exp - e_in > 31 || pl.res_exp - pl.agg_exp > 31 ||
          pl.agg_exp - pl.res_exp > 31)
        atomicOr(p.status, 1);
    }
  }
  __syncthreads();
  const int e_in = e_in_s;
  const float so = exp2f((float)-pl.res_exp), sm = exp2f((float)-p.next_e_mean);
  float m0 = 0.f, m1 = 0.f;
  const int cgn = p.H >> 2;
  if ((p.H & 3) == 0 && cgn <= 256) {
    // column-resident, 16-byte accesses (see fxp_bn_reduce_kernel)
    const int rpb = 256 / cgn, cg = threadIdx.x % cgn, rl = threadIdx.x / cgn;
    if (rl < rpb) {
      float mmf[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.next_fm) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mmf[e] = __int2float_rn(p.next_mmean[4 * cg + e]) * sm;
      }
      const long long rows = p.n / p.H;
      const int4* G4 = reinterpret_cast<const int4*>(p.G);
      const int4* X4 = reinterpret_cast<const int4*>(p.X);
      int4* O4 = reinterpret_cast<int4*>(p.OUT);
      int vmax[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN}, vmin[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
      for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
        const int4 gv = G4[r * cgn + cg], xv = X4[r * cgn + cg];
        const int g4[4] = {gv.x, gv.y, gv.z, gv.w}, x4[4] = {xv.x, xv.y, xv.z, xv.w};
        int o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int v = fx_add_cb_apply(g4[e], p.g_bits, p.g_exp, x4[e], p.x_bits, e_in, pl, p.res_bits);
          v = v > 0 ? v : 0;
          o4[e] = v;
          vmax[e] = max(vmax[e], v); vmin[e] = min(vmin[e], v);
          if (p.next_fm) {
            const float f = __int2float_rn(v) * so;
            m0 = fmaxf(m0, fabsf(f + mmf[e]));
            m1 = fmaxf(m1, fabsf(f));
          }
        }
        O4[r * cgn + cg] = make_int4(o4[0], o4[1], o4[2], o4[3]);
      }
      if (cstat && vmax[0] >= vmin[0]) {                         // this thread saw at least one row
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicMax(&scmax[4 * cg + e], ord_enc(vmax[e]));
          atomicMax(&scmin[4 * cg + e], ord_enc(~vmin[e]));
        }
      }
    }
  } else
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    int v = fx_add_cb_apply(p.G[i], p.g_bits, p.g_exp, p.X[i], p.x_bits, e_in, pl, p.res_bits);
    v = v > 0 ? v : 0;
    p.OUT[i] = v;
    if (cstat) {
      atomicMax(&scmax[(int)(i % p.H)], ord_enc(v));
      atomicMax(&scmin[(int)(i % p.H)], ord_enc(~v));
    }
    if (p.next_fm) {
      const float f = __int2float_rn(v) * so;
      m0 = fmaxf(m0, fabsf(f + __int2float_rn(p.next_mmean[(int)(i % p.H)]) * sm));
      m1 = fmaxf(m1, fabsf(f));
    }
  }
  if (p.next_fm) {
    block_max_to(p.next_fm + M_S1, m0, red);
    block_max_to(p.next_fm + M_A1, m1, red);
  }
  if (cstat) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.H; i += blockDim.x) {
      if (scmax[i]) atomicMax(p.next_cmax + i, scmax[i]);
      if (scmin[i]) atomicMax(p.next_cmin + i, scmin[i]);
    }
  }
}

static int ew_grid(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sm_count() * 8)); }

}  // namespace s5

// ------------------------------------------------------------------------------------------- model handle
namespace s5 {
struct FxpDenseDev {
  SparseDev w;
  FxpTcLinearDev tc;   // dense-tile form of the same weights (fxp_tc.cu); !ok: CUDA-core kernels only
  int* bias_aligned = nullptr;
  s5_fxp_dense_desc q;
};
struct FxpLayerDev {
  int *mmean = nullptr, *inv = nullptr, *scale = nullptr, *bias = nullptr;
  int *A_re = nullptr, *A_im = nullptr, *D = nullptr;
  SparseDev wb, wcr, wci;
  FxpTcLinearDev tcb;  // B projection on the tensor cores
  FxpTcGluDev tcg;     // C projection + out2 on the tensor cores
  FxpDenseDev out2;
  s5_fxp_layer_desc q;
  float pm_mean = 0.f, pm_bias = 0.f;
};

static int upload_i(const int32_t* h, size_t n, int** d) {
  if (!h) S5_FAIL(S5_E_BADARG, "fxp_model_create: null parameter array");
  S5_CUDA(cudaMalloc(d, std::max<size_t>(1, n) * sizeof(int)));
  S5_CUDA(cudaMemcpy(*d, h, n * sizeof(int), cudaMemcpyHostToDevice));
  return S5_OK;
}
static float absmax_float(const int32_t* v, int n, int e) {
  // np.abs(op.to_float()).max(): int32 -> float32 (RN), divided by 2**e
  float m = 0.f;
  for (int i = 0; i < n; ++i) m = std::max(m, std::fabs((float)v[i] / (float)(1ll << e)));
  return m;
}
static int build_dense(const s5_fxp_dense_desc& q, int K, int N, int dense, FxpDenseDev* out) {
  int rc;
  out->q = q;
  if (!q.weight_host || !q.bias_host) S5_FAIL(S5_E_BADARG, "fxp_model_create: null dense parameters");
  if ((rc = sparse_build(q.weight_host, K, N, S5_I32, dense, &out->w))) return rc;
  if ((rc = fxp_tc_linear_build(q.weight_host, K, N, &out->tc))) return rc;
  // fxp_add(wx, bias, result_exp=out_exp): bias.change_exp(out_exp) saturates at the bias width (fxparray.py:449-455)
  if (std::abs(q.out_exp - q.b_exp) > 31) S5_FAIL(S5_E_UNSUPPORTED, "fxp dense: bias/out exponent gap above 31");
  std::vector<int> b(N);
  for (int n = 0; n < N; ++n) b[n] = fx_change_exp(q.bias_host[n], q.b_bits, q.b_exp, q.out_exp, S5_ROUND_FLOOR);
  return upload_i(b.data(), N, &out->bias_aligned);
}
}  // namespace s5

struct s5_fxp_model {
  s5_fxp_model_desc desc;
  s5::FxpDenseDev enc, dec;
  std::vector<s5::FxpLayerDev> layers;
  // Dense-tile (tensor core) form of the integer projections: from S5B200_FXP_TC_MIN_ROWS (default 512) frame rows up, when
  // the weights are 8-bit payloads and the activation operand is at most 16 bits wide (w8a16).  S5B200_FXP_TC=0: CUDA cores
  // only.  Both are read when the handle is created.
  bool tc_on = true; long long tc_min_rows = 512;
  // S5B200_FXP_BN_DERIVE=0: the three BatchNorm-reduce passes per block instead of fxp_bn_derive_kernel;
  // S5B200_FXP_BN_FORCE_SLOW=1 (testing): the derive kernel takes its whole-tensor fallback
  bool bn_derive = true, bn_force_slow = false;
  bool dec_ws = true;        // S5B200_FXP_DEC_WS=0: the generic dense-tile kernel for the decoder instead of the warp-specialised one
  bool enc_ws = true;        // S5B200_FXP_ENC_WS=0: the generic dense-tile kernel for the encoder instead of the warp-specialised one
  bool u16 = true;           // S5B200_FXP_U16=0: the dense-tile GLU kernel re-runs the BatchNorm chain instead of reading u from the B projection
  bool relu_in_glu = true;   // S5B200_FXP_RELU_IN_GLU=0: the scan applies fxp_relu itself even when the dense-tile GLU follows
  bool use_tc(long long rows) const { return tc_on && rows >= tc_min_rows; }
  s5_fxp_sync_fn sync = nullptr;   // global-batch parity under sharding: max-all-reduce of the running maxima (s5b200.h)
  void* sync_user = nullptr;
};

namespace s5 {

struct FxpWs {
  size_t off_a, off_b, off_g, off_bu, off_state, off_bad, off_u16, total;
  size_t state_ints, max_words;
};
static FxpWs fxp_ws(const s5_fxp_model* m, int B, int L) {
  const size_t rows = (size_t)B * L;
  FxpWs w;
  const size_t act = round_up(rows * m->desc.H * 4, 256);
  w.off_a = 0; w.off_b = act; w.off_g = 2 * act; w.off_bu = 3 * act;
  w.off_state = w.off_bu + round_up(rows * m->desc.P * 8, 256);
  // maxima | column extremes (max, ~min) of every block input | exponents | 8 status words | one bad-tile counter per block
  // (dense-tile GLU, fxp_tc.cu).  Maxima and extremes are contiguous: one max-all-reduce covers both (global-batch parity).
  w.max_words = (size_t)(m->desc.n_layers + 1) * kStateStride + (size_t)m->desc.n_layers * 2 * m->desc.H;
  w.state_ints = w.max_words + (size_t)(m->desc.n_layers + 1) * kStateStride + 8 + (size_t)m->desc.n_layers;
  w.off_bad = w.off_state + round_up(w.state_ints * 4, 256);
  // bad-tile list: (row0, count) per tile; tiles are at least 64 rows except one short tile per CTA
  w.off_u16 = w.off_bad + round_up(((rows + 63) / 64 + 2048) * 8, 256);
  w.total = w.off_u16 + round_up(rows * m->desc.H * 2, 256);   // mixer input u as 16-bit payloads (B projection -> GLU, dense tiles)
  return w;
}

static BnChain make_chain(const FxpLayerDev& l) {
  BnChain c{};
  c.mmean = l.mmean; c.inv = l.inv; c.scale = l.scale; c.bias = l.bias;
  c.e_mean = l.q.bn_mean_exp; c.b_mean = l.q.bn_mean_bits; c.e_inv = l.q.bn_invsq_exp; c.b_inv = l.q.bn_invsq_bits;
  c.e_scale = l.q.bn_scale_exp; c.b_scale = l.q.bn_scale_bits; c.e_bias = l.q.bn_bias_exp; c.b_bias = l.q.bn_bias_bits;
  c.pm_mean = l.pm_mean; c.pm_bias = l.pm_bias;
  c.has_scale = l.scale != nullptr; c.has_bias = l.bias != nullptr;
  c.x_bits = 16; c.u_bits = l.q.u_bits; c.u_exp = l.q.u_exp;
  return c;
}

struct FxpRun {
  const s5_fxp_model* m; int B, L; unsigned char* base; FxpWs w; cudaStream_t st;
  unsigned* fm(int l) const { return reinterpret_cast<unsigned*>(base + w.off_state) + (size_t)l * kStateStride; }
  unsigned* cmax(int l) const { return reinterpret_cast<unsigned*>(base + w.off_state) + (size_t)(m->desc.n_layers + 1) * kStateStride + (size_t)l * 2 * m->desc.H; }
  unsigned* cmin(int l) const { return cmax(l) + m->desc.H; }
  int* ex(int l) const { return reinterpret_cast<int*>(base + w.off_state) + w.max_words + (size_t)l * kStateStride; }
  int* status() const { return reinterpret_cast<int*>(base + w.off_state) + w.max_words + (size_t)(m->desc.n_layers + 1) * kStateStride; }
  int* bad_count(int l) const { return status() + 8 + l; }
  int* bad_list() const { return reinterpret_cast<int*>(base + w.off_bad); }
  unsigned short* u16() const { return reinterpret_cast<unsigned short*>(base + w.off_u16); }
};

// Runs encoder + blocks [0, upto_layer]; stop_stage: 0 full block, 1 stop after BN (pre_s5 materialised into g),
// 2 stop after the scan, 3 stop after the GLU.  Returns with `cur` = buffer holding the last block output / input.
static int fxp_run(const FxpRun& r, const void* x, bool x_is_float, int upto_layer, int stop_stage, int** cur_out,
                   int** other_out) {
  const s5_fxp_model* m = r.m;
  const s5_fxp_model_desc& d = m->desc;
  const long long rows = (long long)r.B * r.L;
  int* a = reinterpret_cast<int*>(r.base + r.w.off_a);
  int* b = reinterpret_cast<int*>(r.base + r.w.off_b);
  int* g = reinterpret_cast<int*>(r.base + r.w.off_g);
  int* bu = reinterpret_cast<int*>(r.base + r.w.off_bu);
  int rc;
  // Sharded runs with global-batch parity: every kernel that produced running maxima is followed by the caller's
  // max-all-reduce over all blocks' maxima (max is idempotent, so reducing the whole region every time is exact).
  auto sync_maxima = [&]() {
    if (m->sync) m->sync(m->sync_user, r.fm(0), (int)r.w.max_words, r.st);
  };
  S5_CUDA(cudaMemsetAsync(r.base + r.w.off_state, 0, r.w.state_ints * 4, r.st));
  // BatchNorm maxima from column extremes (fxp_bn_derive_kernel) instead of three passes per block: needs the dense-tile
  // encoder (its write-out has lanes along columns) and a block width the derive CTA covers with one thread per column
  bool cs = m->bn_derive && m->use_tc(rows) && m->enc.tc.ok && m->enc.q.inp_bits <= 16 && d.H <= 1024 && m->enc.q.out_bits <= 30;
  for (const FxpLayerDev& l : m->layers) cs = cs && l.q.gate_res_bits <= 30;
  {  // encoder FxpDense + fxp_relu (fxpmodel.py:1262-1266) + maxima for block 0
    const s5_fxp_dense_desc& q = m->enc.q;
    FxpLinearParams p{};
    p.X = x; p.Y = a; p.rows = rows; p.w = wview(m->enc.w); p.mode = LIN_ENC;
    p.x_is_float = x_is_float; p.in_bits = q.inp_bits; p.in_exp = q.inp_exp;
    p.rshift0 = p.rshift1 = q.inp_exp + q.w_exp - q.out_exp; p.round_mode = S5_ROUND_FLOOR; p.out_bits = q.out_bits;
    p.bias_aligned = m->enc.bias_aligned; p.relu = 1;
    p.next_mmean = m->layers[0].mmean; p.next_e_mean = m->layers[0].q.bn_mean_exp; p.next_fm = r.fm(0); p.out_exp = q.out_exp;
    p.status = r.status();
    S5_CUDA(cudaMemcpyAsync(r.ex(0) + E_IN, &m->enc.q.out_exp, sizeof(int), cudaMemcpyHostToDevice, r.st));
    const bool tc = m->use_tc(rows) && m->enc.tc.ok && q.inp_bits <= 16 && p.round_mode == S5_ROUND_FLOOR;
    if (tc && cs) { p.next_fm = nullptr; p.next_cmax = r.cmax(0); p.next_cmin = r.cmin(0); }
    // the warp-specialised encoder (TMA input ring; the column extremes fall out of its lane = column epilogue) where the
    // shape is covered and the maxima come from the extremes; else the generic dense-tile kernel / the CUDA cores
    rc = S5_E_UNSUPPORTED;
    if (tc && cs && m->enc_ws)
      rc = tc_enc_fx_launch(m->enc.tc.W8[1], m->enc.tc.Kfull, m->enc.tc.Kp[1], m->enc.tc.Np, m->enc.tc.Nfull, m->enc.tc.KV, x, x_is_float ? 1 : 0,
                            q.inp_bits, q.inp_exp, a, rows, p.rshift0, q.out_bits, m->enc.bias_aligned, 1, r.cmax(0), r.cmin(0), r.st);
    if (rc == S5_E_UNSUPPORTED) rc = tc ? launch_fxp_linear_tc(p, m->enc.tc, r.st, S5_K_FXP_ENC) : launch_fxp_linear(p, r.st, S5_K_FXP_ENC);
    if (rc) return rc;
    sync_maxima();
  }
  for (int li = 0; li <= upto_layer; ++li) {
    const FxpLayerDev& l = m->layers[li];
    const BnChain c = make_chain(l);
    const long long n = rows * d.H;
    auto reduce = [&](int upto, int next) -> int {
      BnReduceParams p{a, n, d.H, c, r.fm(li), r.ex(li), r.status(), upto, next};
      ProfScope prof(S5_K_FXP_BN, r.st);
      fxp_bn_reduce_kernel<<<ew_grid(n), 256, 0, r.st>>>(p);
      S5_LAUNCH_CHECK();
      sync_maxima();
      return S5_OK;
    };
    if (cs) {
      BnDeriveParams p{a, n, d.H, c, r.fm(li), r.ex(li), r.status(), r.cmax(li), r.cmin(li), m->bn_force_slow ? 1 : 0};
      ProfScope prof(S5_K_FXP_BN, r.st);
      fxp_bn_derive_kernel<<<1, (d.H <= 256 && !m->bn_force_slow) ? 256 : 1024, 0, r.st>>>(p);
      S5_LAUNCH_CHECK();
    } else {
      if ((rc = reduce(1, 2))) return rc;
      if (c.has_scale && (rc = reduce(2, 3))) return rc;
      if (c.has_bias && (rc = reduce(3, 4))) return rc;
    }
    if (stop_stage == 1 && li == upto_layer) {
      BnApplyParams p{a, g, n, d.H, c, r.fm(li), r.ex(li)};
      fxp_bn_apply_kernel<<<ew_grid(n), 256, 0, r.st>>>(p);
      S5_LAUNCH_CHECK();
      *cur_out = g; *other_out = a;
      return S5_OK;
    }
    // dense tiles: 16-bit operands everywhere (w8a16); tiles whose states exceed 16 bits come back in the list and
    // are redone by the CUDA-core kernel, which multiplies full int32 states
    const s5_fxp_dense_desc& o2 = l.out2.q;
    const bool tc_glu = m->use_tc(rows) && l.tcg.ok && l.q.y_bits <= 16 && o2.inp_bits <= 16 && o2.out_bits <= 16 &&
                        l.q.gate_l_bits <= 16 && l.q.gate_r_bits <= 16 && l.q.u_bits <= 16;
    const bool tc_bproj = m->use_tc(rows) && l.tcb.ok && l.q.u_bits <= 16 && (d.H % 128) != 1;
    // the dense-tile B projection leaves u behind as 16-bit payloads for the dense-tile GLU kernel (S5B200_FXP_U16=0: recompute)
    unsigned short* u16 = (tc_bproj && tc_glu && m->u16 && (d.H & 3) == 0) ? r.u16() : nullptr;
    {  // Bu = fxp_matmul(u, B^T) for re / im (fxpmodel.py:631-644), interleaved
      FxpLinearParams p{};
      p.X = a; p.Y = bu; p.rows = rows; p.w = wview(l.wb); p.mode = LIN_BPROJ;
      p.rshift0 = l.q.u_exp + l.q.eB_re - l.q.Bu_re_exp; p.rshift1 = l.q.u_exp + l.q.eB_im - l.q.Bu_im_exp;
      p.round_mode = S5_ROUND_FLOOR; p.out_bits = l.q.Bu_re_bits;
      p.c = c; p.fm = r.fm(li); p.ex = r.ex(li); p.status = r.status();
      const bool tc = tc_bproj;
      p.U16 = u16;
      if ((rc = tc ? launch_fxp_linear_tc(p, l.tcb, r.st, S5_K_FXP_BPROJ) : launch_fxp_linear(p, r.st, S5_K_FXP_BPROJ))) return rc;
    }
    // fxp_relu of the states: in the scan's post role, or - cheaper, the scan is latency bound with few clips - while the
    // dense-tile GLU kernel stages them (not when the states themselves are tapped)
    const bool relu_in_glu = tc_glu && m->relu_in_glu && !(stop_stage == 2 && li == upto_layer);
    {  // recurrent_loop + fxp_relu (fxpmodel.py:682-742)
      ScanI32Params q{};
      q.Bu_re = bu; q.Bu_im = nullptr; q.A_re = l.A_re; q.A_im = l.A_im; q.xs_re = bu; q.xs_im = nullptr;
      q.L = r.L; q.P = d.P; q.lanes = (long long)r.B * d.P;
      q.eA_re = l.q.eA_re; q.eA_im = l.q.eA_im; q.eBu_re = l.q.Bu_re_exp; q.eBu_im = l.q.Bu_im_exp;
      q.ex_re = l.q.x_re_exp; q.ex_im = l.q.x_im_exp; q.relu = relu_in_glu ? 0 : 1;
      if ((rc = launch_scan_i32(q, true, r.st))) return rc;
    }
    if (stop_stage == 2 && li == upto_layer) { *cur_out = bu; *other_out = a; return S5_OK; }
    {  // C projection, D feed-through, out2, sigmoid, mult_gate
      FxpGluParams p{};
      p.XS = bu; p.X = a; p.G = g; p.rows = rows; p.skip_zero = fxp_skip_zero();
      p.wcr = wview(l.wcr); p.wci = wview(l.wci); p.wo = wview(l.out2.w);
      p.c = c; p.fm = r.fm(li); p.ex = r.ex(li); p.status = r.status();
      p.D = l.D; p.rs_cre = l.q.x_re_exp + l.q.eC_re - l.q.y_exp; p.rs_cim = l.q.x_im_exp + l.q.eC_im - l.q.y_exp;
      p.rs_D = l.q.eD + l.q.u_exp - l.q.y_exp; p.y_bits = l.q.y_bits; p.conj = 1;
      const s5_fxp_dense_desc& o = l.out2.q;
      p.o_in_change = (l.q.y_bits > o.inp_bits) || (l.q.y_exp > o.inp_exp);
      p.y_exp = l.q.y_exp; p.o_inp_bits = o.inp_bits; p.o_inp_exp = o.inp_exp;
      const int eff_in = p.o_in_change ? o.inp_exp : l.q.y_exp;
      p.o_rs = eff_in + o.w_exp - o.out_exp; p.o_out_bits = o.out_bits; p.o_out_exp = o.out_exp;
      p.o_bias_aligned = l.out2.bias_aligned;
      p.sig_x_exp = l.q.sig_x_exp; p.sig_y_exp = l.q.sig_y_exp;
      for (int i = 0; i < 8; ++i) p.lut[i] = l.q.sig_lut[i];
      p.gl_bits = l.q.gate_l_bits; p.gl_exp = l.q.gate_l_exp; p.gr_bits = l.q.gate_r_bits; p.gr_exp = l.q.gate_r_exp;
      p.g_rs = l.q.gate_l_exp + l.q.gate_r_exp - l.q.gate_res_exp; p.g_bits = l.q.gate_res_bits; p.g_exp = l.q.gate_res_exp;
      const bool tc = tc_glu;
      if (tc) {
        FxpTcGluArgs t{};
        t.XS = bu; t.relu_in = relu_in_glu ? 1 : 0; t.U16 = u16; t.X = p.X; t.G = p.G; t.rows = rows; t.c = c; t.fm = p.fm; t.ex = p.ex; t.status = p.status;
        t.D = p.D; t.rs_cre = p.rs_cre; t.rs_cim = p.rs_cim; t.rs_D = p.rs_D; t.y_bits = p.y_bits; t.conj = p.conj;
        t.o_in_change = p.o_in_change; t.y_exp = p.y_exp; t.o_inp_bits = p.o_inp_bits; t.o_inp_exp = p.o_inp_exp; t.o_rs = p.o_rs;
        t.o_out_bits = p.o_out_bits; t.o_out_exp = p.o_out_exp; t.o_bias_aligned = p.o_bias_aligned;
        t.sig_x_exp = p.sig_x_exp; t.sig_y_exp = p.sig_y_exp;
        for (int i = 0; i < 8; ++i) t.lut[i] = p.lut[i];
        t.gl_bits = p.gl_bits; t.gl_exp = p.gl_exp; t.gr_bits = p.gr_bits; t.gr_exp = p.gr_exp; t.g_rs = p.g_rs;
        t.g_bits = p.g_bits; t.g_exp = p.g_exp;
        t.bad_list = r.bad_list(); t.bad_count = r.bad_count(li);
        if ((rc = fxp_tc_glu_launch(l.tcg, t, r.st))) return rc;
        p.bad_list = r.bad_list(); p.bad_count = r.bad_count(li);
      }
      if ((rc = launch_fxp_glu(p, r.st))) return rc;
      sync_maxima();
    }
    if (stop_stage == 3 && li == upto_layer) { *cur_out = g; *other_out = a; return S5_OK; }
    {  // residual add (compute_best) + fxp_relu, maxima for the next block
      FxpResidParams p{};
      p.G = g; p.X = a; p.OUT = b; p.n = n; p.H = d.H;
      p.fm = r.fm(li); p.ex = r.ex(li); p.status = r.status();
      p.g_bits = l.q.gate_res_bits; p.g_exp = l.q.gate_res_exp; p.x_bits = 16; p.res_bits = l.q.gate_res_bits;
      p.next_ex = r.ex(li + 1);
      if (li + 1 < d.n_layers) {
        p.next_mmean = m->layers[li + 1].mmean; p.next_e_mean = m->layers[li + 1].q.bn_mean_exp; p.next_fm = r.fm(li + 1);
        if (cs) { p.next_fm = nullptr; p.next_cmax = r.cmax(li + 1); p.next_cmin = r.cmin(li + 1); }
      }
      ProfScope prof(S5_K_FXP_RESID, r.st);
      fxp_resid_kernel<<<ew_grid(n), 256, 0, r.st>>>(p);
      S5_LAUNCH_CHECK();
      sync_maxima();
    }
    std::swap(a, b);
  }
  *cur_out = a; *other_out = b;
  return S5_OK;
}

static int fxp_forward_impl(const s5_fxp_model* m, const void* x, bool x_is_float, int32_t* y_int, float* y_float, int B,
                            int L, void* workspace, size_t workspace_bytes, int32_t* status_dev, void* stream) {
  if (!m || !x || !workspace || (!y_int && !y_float)) S5_FAIL(S5_E_BADARG, "fxp_model_forward: null pointer");
  if (B <= 0 || L <= 0) S5_FAIL(S5_E_BADSHAPE, "fxp_model_forward: B=%d L=%d must be positive", B, L);
  FxpRun r{m, B, L, reinterpret_cast<unsigned char*>(workspace), fxp_ws(m, B, L), (cudaStream_t)stream};
  if (workspace_bytes < r.w.total) S5_FAIL(S5_E_BADARG, "fxp_model_forward: workspace %zu < %zu bytes", workspace_bytes, r.w.total);
  if (!aligned16(workspace)) S5_FAIL(S5_E_BADARG, "fxp_model_forward: workspace must be 16-byte aligned");
  int *cur = nullptr, *other = nullptr, rc;
  if ((rc = fxp_run(r, x, x_is_float, m->desc.n_layers - 1, 0, &cur, &other))) return rc;
  const s5_fxp_dense_desc& q = m->dec.q;
  FxpLinearParams p{};
  p.X = cur; p.Y = y_int; p.Yf = y_float; p.rows = (long long)B * L; p.w = wview(m->dec.w); p.mode = LIN_DEC;
  p.skip_zero = fxp_skip_zero();
  p.round_mode = S5_ROUND_FLOOR; p.out_bits = q.out_bits; p.bias_aligned = m->dec.bias_aligned; p.relu = 0;
  p.out_scale = 1.0f / (float)(1ll << q.out_exp);
  p.ex = r.ex(m->desc.n_layers); p.status = r.status();
  p.dec_inp_bits = q.inp_bits; p.dec_inp_exp = q.inp_exp; p.dec_w_exp = q.w_exp; p.dec_out_exp = q.out_exp;
  const bool tc = m->use_tc(p.rows) && m->dec.tc.ok && q.inp_bits <= 16 && (m->desc.H % 128) != 1 && (m->desc.H & 3) == 0;
  // the warp-specialised decoder (TMA input ring, bulk stores of whole half-tiles of both outputs) where the shape is covered
  rc = S5_E_UNSUPPORTED;
  if (tc && m->dec_ws && m->dec.tc.xn)
    rc = tc_dec_fx_launch(m->dec.tc.W8[0], m->dec.tc.WN[0], m->dec.tc.Kfull, m->dec.tc.Kp[0], m->dec.tc.Np, m->dec.tc.Nfull, m->dec.tc.KV, cur,
                          y_int, y_float, p.rows, r.ex(m->desc.n_layers) + E_IN, q.inp_bits, q.inp_exp, q.w_exp, q.out_exp, q.out_bits,
                          m->dec.bias_aligned, p.out_scale, r.status(), r.st);
  if (rc == S5_E_UNSUPPORTED) rc = tc ? launch_fxp_linear_tc(p, m->dec.tc, r.st, S5_K_FXP_DEC) : launch_fxp_linear(p, r.st, S5_K_FXP_DEC);
  if (rc) return rc;
  if (status_dev) S5_CUDA(cudaMemcpyAsync(status_dev, r.status(), 4 * sizeof(int), cudaMemcpyDeviceToDevice, r.st));
  return S5_OK;
}

}  // namespace s5

extern "C" int s5b200_fxp_model_set_maxima_sync(s5_fxp_model* m, s5_fxp_sync_fn fn, void* user) {
  if (!m) S5_FAIL(S5_E_BADARG, "fxp_model_set_maxima_sync: null model");
  m->sync = fn;
  m->sync_user = user;
  return S5_OK;
}

extern "C" int s5b200_fxp_model_create(const s5_fxp_model_desc* d, s5_fxp_model** out) {
  using namespace s5;
  if (!d || !out || !d->layers) S5_FAIL(S5_E_BADARG, "fxp_model_create: null pointer");
  if (d->n_layers < 1 || d->n_layers > 64 || d->H < 1 || d->H > 1024 || d->P < 1 || d->P > 1024 || d->d_input < 1 ||
      d->d_input > 4096 || d->d_output < 1 || d->d_output > 4096)
    S5_FAIL(S5_E_BADSHAPE, "fxp_model_create: unsupported dimensions");
  s5_fxp_model* m = new s5_fxp_model();
  m->desc = *d;
  m->desc.layers = nullptr;
  if (const char* e = getenv("S5B200_FXP_TC")) m->tc_on = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_TC_MIN_ROWS")) m->tc_min_rows = atoll(e);
  if (const char* e = getenv("S5B200_FXP_BN_DERIVE")) m->bn_derive = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_RELU_IN_GLU")) m->relu_in_glu = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_U16")) m->u16 = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_ENC_WS")) m->enc_ws = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_DEC_WS")) m->dec_ws = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FXP_BN_FORCE_SLOW")) m->bn_force_slow = atoi(e) != 0;
  const int H = d->H, P = d->P, dense = d->use_sparse ? 0 : 1;
  int rc = S5_OK;
  auto fail = [&](int code) { s5b200_fxp_model_destroy(m); return code; };
  auto shift_ok = [](int s) { return s >= 0 && s <= 31; };
  if ((rc = build_dense(d->encoder, d->d_input, H, dense, &m->enc))) return fail(rc);
  if ((rc = build_dense(d->decoder, H, d->d_output, dense, &m->dec))) return fail(rc);
  if (!shift_ok(d->encoder.inp_exp + d->encoder.w_exp - d->encoder.out_exp))
    return fail((set_error("fxp_model_create: encoder matmul shift is negative (the reference shifts by a negative count)"), S5_E_UNSUPPORTED));
  m->layers.resize(d->n_layers);
  for (int i = 0; i < d->n_layers; ++i) {
    const s5_fxp_layer_desc& q = d->layers[i];
    FxpLayerDev& l = m->layers[i];
    l.q = q;
    if (!q.bn_minus_mean_host || !q.bn_invsq_var_host || !q.A_re_host || !q.A_im_host || !q.B_re_host || !q.B_im_host ||
        !q.C_re_host || !q.C_im_host || !q.D_host)
      return fail((set_error("fxp_model_create: null layer parameter"), S5_E_BADARG));
    if ((rc = upload_i(q.bn_minus_mean_host, H, &l.mmean))) return fail(rc);
    if ((rc = upload_i(q.bn_invsq_var_host, H, &l.inv))) return fail(rc);
    if (q.bn_scale_host && (rc = upload_i(q.bn_scale_host, H, &l.scale))) return fail(rc);
    if (q.bn_bias_host && (rc = upload_i(q.bn_bias_host, H, &l.bias))) return fail(rc);
    l.pm_mean = absmax_float(q.bn_minus_mean_host, H, q.bn_mean_exp);
    if (q.bn_bias_host) l.pm_bias = absmax_float(q.bn_bias_host, H, q.bn_bias_exp);
    if ((rc = upload_i(q.A_re_host, P, &l.A_re))) return fail(rc);
    if ((rc = upload_i(q.A_im_host, P, &l.A_im))) return fail(rc);
    if ((rc = upload_i(q.D_host, H, &l.D))) return fail(rc);
    // B: [P][H] -> W[h][2p + c] ; C_re/C_im: [H][P] -> W[2p + c][h] with only the matching parity populated
    std::vector<int> wb((size_t)H * 2 * P, 0), wcr((size_t)2 * P * H, 0), wci((size_t)2 * P * H, 0);
    for (int p = 0; p < P; ++p)
      for (int h = 0; h < H; ++h) {
        wb[(size_t)h * 2 * P + 2 * p] = q.B_re_host[(size_t)p * H + h];
        wb[(size_t)h * 2 * P + 2 * p + 1] = q.B_im_host[(size_t)p * H + h];
        wcr[(size_t)(2 * p) * H + h] = q.C_re_host[(size_t)h * P + p];
        wci[(size_t)(2 * p + 1) * H + h] = q.C_im_host[(size_t)h * P + p];
      }
    if ((rc = sparse_build(wb.data(), H, 2 * P, S5_I32, dense, &l.wb))) return fail(rc);
    if ((rc = fxp_tc_linear_build(wb.data(), H, 2 * P, &l.tcb))) return fail(rc);
    // the C halves are always kept compressed: a dense tile would store the structural zeros of the other parity
    if ((rc = sparse_build(wcr.data(), 2 * P, H, S5_I32, 0, &l.wcr))) return fail(rc);
    if ((rc = sparse_build(wci.data(), 2 * P, H, S5_I32, 0, &l.wci))) return fail(rc);
    if ((rc = build_dense(q.out2, H, H, dense, &l.out2))) return fail(rc);
    if ((rc = fxp_tc_glu_build(q.C_re_host, q.C_im_host, q.out2.weight_host, H, P, &l.tcg))) return fail(rc);
    const int eff_in = ((q.y_bits > q.out2.inp_bits) || (q.y_exp > q.out2.inp_exp)) ? q.out2.inp_exp : q.y_exp;
    const int shifts[] = {q.u_exp + q.eB_re - q.Bu_re_exp, q.u_exp + q.eB_im - q.Bu_im_exp, q.x_re_exp + q.eC_re - q.y_exp,
                          q.x_im_exp + q.eC_im - q.y_exp, q.eD + q.u_exp - q.y_exp, eff_in + q.out2.w_exp - q.out2.out_exp,
                          q.gate_l_exp + q.gate_r_exp - q.gate_res_exp, q.eA_re, q.eA_im};
    for (int s : shifts)
      if (!shift_ok(s))
        return fail((set_error("fxp_model_create: layer %d has a shift of %d (negative shifts raise in the reference, fxparray.py:619-621)", i, s), S5_E_UNSUPPORTED));
    if (q.sig_x_exp < 0 || q.sig_x_exp > 15 || q.sig_y_exp < 1 || q.sig_y_exp > 30)
      return fail((set_error("fxp_model_create: sigmoid exponents out of range"), S5_E_UNSUPPORTED));
    if (q.Bu_re_bits != q.Bu_im_bits) return fail((set_error("fxp_model_create: Bu_re/Bu_im widths differ"), S5_E_UNSUPPORTED));
  }
  *out = m;
  return S5_OK;
}

extern "C" int s5b200_fxp_model_destroy(s5_fxp_model* m) {
  if (!m) return S5_OK;
  auto fd = [](s5::FxpDenseDev& d) { s5::sparse_free(&d.w); s5::fxp_tc_linear_free(&d.tc); cudaFree(d.bias_aligned); };
  fd(m->enc); fd(m->dec);
  for (auto& l : m->layers) {
    cudaFree(l.mmean); cudaFree(l.inv); cudaFree(l.scale); cudaFree(l.bias);
    cudaFree(l.A_re); cudaFree(l.A_im); cudaFree(l.D);
    s5::sparse_free(&l.wb); s5::sparse_free(&l.wcr); s5::sparse_free(&l.wci);
    s5::fxp_tc_linear_free(&l.tcb);
    s5::fxp_tc_glu_free(&l.tcg);
    fd(l.out2);
  }
  delete m;
  return S5_OK;
}

extern "C" size_t s5b200_fxp_model_workspace_bytes(const s5_fxp_model* m, int B, int L) {
  if (!m || B <= 0 || L <= 0) return 0;
  return s5::fxp_ws(m, B, L).total;
}

extern "C" int s5b200_fxp_model_forward(const s5_fxp_model* m, const float* x, int32_t* y_int, float* y_float, int B, int L,
                                        void* workspace, size_t workspace_bytes, int32_t* status_dev, void* stream) {
  return s5::fxp_forward_impl(m, x, true, y_int, y_float, B, L, workspace, workspace_bytes, status_dev, stream);
}
extern "C" int s5b200_fxp_model_forward_int(const s5_fxp_model* m, const int32_t* x_int, int32_t* y_int, float* y_float,
                                            int B, int L, void* workspace, size_t workspace_bytes, int32_t* status_dev,
                                            void* stream) {
  return s5::fxp_forward_impl(m, x_int, false, y_int, y_float, B, L, workspace, workspace_bytes, status_dev, stream);
}

extern "C" int s5b200_fxp_model_tap(const s5_fxp_model* m, const int32_t* x_int, int layer, int tap, int32_t* dst,
                                    int* exp_out_host, int B, int L, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace s5;
  if (!m || !x_int || !dst || !workspace) S5_FAIL(S5_E_BADARG, "fxp_model_tap: null pointer");
  if (layer < 0 || layer >= m->desc.n_layers) S5_FAIL(S5_E_BADARG, "fxp_model_tap: layer %d out of range", layer);
  FxpRun r{m, B, L, reinterpret_cast<unsigned char*>(workspace), fxp_ws(m, B, L), (cudaStream_t)stream};
  if (workspace_bytes < r.w.total) S5_FAIL(S5_E_BADARG, "fxp_model_tap: workspace too small");
  int *cur = nullptr, *other = nullptr, rc;
  const size_t rows = (size_t)B * L;
  size_t bytes = rows * m->desc.H * 4;
  const int* src = nullptr;
  const int* exp_src = nullptr;
  int exp_static = 0;
  const s5_fxp_layer_desc& q = m->layers[layer].q;
  switch (tap) {
    case S5_TAP_LAYER_INPUT:
      if (layer == 0) { if ((rc = fxp_run(r, x_int, false, 0, 1, &cur, &other))) return rc; src = other; }
      else { if ((rc = fxp_run(r, x_int, false, layer - 1, 0, &cur, &other))) return rc; src = cur; }
      exp_src = r.ex(layer) + E_IN;
      break;
    case S5_TAP_PRE_S5:
      if ((rc = fxp_run(r, x_int, false, layer, 1, &cur, &other))) return rc;
      src = cur; exp_static = q.u_exp;
      break;
    case S5_TAP_XS:
      if ((rc = fxp_run(r, x_int, false, layer, 2, &cur, &other))) return rc;
      src = cur; bytes = rows * m->desc.P * 8; exp_static = q.x_re_exp;
      break;
    case S5_TAP_POST_GLU:
      if ((rc = fxp_run(r, x_int, false, layer, 3, &cur, &other))) return rc;
      src = cur; exp_static = q.gate_res_exp;
      break;
    case S5_TAP_OUTPUT:
      if ((rc = fxp_run(r, x_int, false, layer, 0, &cur, &other))) return rc;
      src = cur; exp_src = r.ex(layer) + E_OUT;
      break;
    default:
      S5_FAIL(S5_E_BADARG, "fxp_model_tap: unknown tap %d", tap);
  }
  S5_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, r.st));
  int e = exp_static;
  if (exp_src) S5_CUDA(cudaMemcpyAsync(&e, exp_src, sizeof(int), cudaMemcpyDeviceToHost, r.st));
  S5_CUDA(cudaStreamSynchronize(r.st));
  if (exp_out_host) *exp_out_host = e;
  return S5_OK;
}

extern "C" int s5b200_spmm_i32(const s5_sparse* w, const int32_t* X, int32_t* Y, long long rows, int rshift, int round_mode,
                               int bits, int skip_zero_cols, void* stream) {
  if (!w || !X || !Y) S5_FAIL(S5_E_BADARG, "spmm_i32: null pointer");
  if (w->d.dtype != S5_I32) S5_FAIL(S5_E_BADARG, "spmm_i32: handle holds float32 weights");
  if (rshift < 0 || rshift > 31 || (round_mode == S5_ROUND_HALF_UP && rshift < 1))
    S5_FAIL(S5_E_BADARG, "spmm_i32: rshift %d invalid for round mode %d", rshift, round_mode);
  if (bits < 1 || bits > 32) S5_FAIL(S5_E_BADARG, "spmm_i32: bits %d outside [1, 32]", bits);
  s5::FxpLinearParams p{};
  p.X = X; p.Y = Y; p.rows = rows; p.w = s5::wview(w->d); p.mode = s5::LIN_PLAIN;
  p.rshift0 = p.rshift1 = rshift; p.round_mode = round_mode; p.out_bits = bits; p.skip_zero = skip_zero_cols != 0;
  return s5::launch_fxp_linear(p, (cudaStream_t)stream);
}
