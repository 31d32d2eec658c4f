// fxp_chain.cuh - device state block and BatchNorm exponent plans shared by the integer model kernels
// (fxp_model.cu: CUDA-core kernels, fxp_tc.cu: dense-tile kernels).  Reference: FxpBatchNorm.forward,
// sparseRNNs/fxpmodel.py:890-944 (four "compute_best" ops).
#pragma once

#include "fxp_prims.cuh"

namespace s5 {

enum { LIN_ENC = 0, LIN_BPROJ = 1, LIN_DEC = 2, LIN_PLAIN = 3 };
constexpr int kStateStride = 8;  // per-layer slots in the device state block
// fmax slots
enum { M_S1 = 0, M_A1 = 1, M_2 = 2, M_3 = 3, M_S4 = 4, M_A4 = 5, M_SR = 6, M_G = 7 };
// exponent slots
enum { E_IN = 0, E_1 = 1, E_2 = 2, E_3 = 3, E_4 = 4, E_OUT = 5 };

struct BnChain {
  const int* mmean; const int* inv; const int* scale; const int* bias;  // device [H]
  int e_mean, b_mean, e_inv, b_inv, e_scale, b_scale, e_bias, b_bias;
  float pm_mean, pm_bias;  // max|to_float(param)|
  int has_scale, has_bias;
  int x_bits;              // width of the block input (16)
  int u_bits, u_exp;
};

// Every exponent, width and shift of the chain is the same for all elements of a launch, so the reference's op sequence
// (fxp_change_cfg -> fxp_change_exp -> shift -> saturate at the old width -> saturate at the new width, fxparray.py:232-326,
// 420-464, 573-637) collapses into "shift left, shift right, clamp" triples that are derived ONCE per launch; applying the
// chain to an element is then ~25 integer instructions instead of ~160 (the chain was the larger part of the B-projection
// and BatchNorm-reduce kernels).  Rounding is FLOOR everywhere in the model (arithmetic >>).
struct ShiftClamp { int shl, sar, lo, hi; };   // v -> clamp((v << shl) >> sar, lo, hi), wrapping shift
struct AddFast { ShiftClamp a, b, post; };     // fxp_add "compute_best": post(a(x) + b(y))

__device__ __forceinline__ int sc_apply(int v, const ShiftClamp& s) {
  return min(max((int)((unsigned)v << s.shl) >> s.sar, s.lo), s.hi);
}
__device__ inline void sc_bounds(int bits, int& lo, int& hi) {
  if (bits >= 32) { lo = INT_MIN; hi = INT_MAX; }
  else { hi = (1 << (bits - 1)) - 1; lo = -hi - 1; }
}
// fx_change_cfg(v, bits, e, nbits, ne, FLOOR) as one ShiftClamp (nested symmetric ranges: two clamps == the tighter one)
__device__ inline ShiftClamp sc_change_cfg(int bits, int e, int nbits, int ne) {
  ShiftClamp s{0, 0, INT_MIN, INT_MAX};
  if (bits == nbits && e == ne) return s;
  if (ne != e) {
    if (ne > e) s.shl = (ne - e) & 31; else s.sar = (e - ne) & 31;
    sc_bounds(bits, s.lo, s.hi);
  }
  if (bits > nbits) {
    int lo, hi;
    sc_bounds(nbits, lo, hi);
    s.lo = max(s.lo, lo); s.hi = min(s.hi, hi);
  }
  return s;
}
// fx_add_cb_apply(a, a_bits, a_exp, b, b_bits, b_exp, plan, result_bits)
__device__ inline AddFast add_fast_make(int a_bits, int a_exp, int b_bits, int b_exp, const AddPlan& p, int result_bits) {
  AddFast f;
  f.a = sc_change_cfg(a_bits, a_exp, max(p.agg_bits, a_bits), p.agg_exp);
  f.b = sc_change_cfg(b_bits, b_exp, max(p.agg_bits, b_bits), p.agg_exp);
  const int d = p.res_exp - p.agg_exp;
  f.post = ShiftClamp{d > 0 ? (d & 31) : 0, d < 0 ? ((-d) & 31) : 0, 0, 0};
  sc_bounds(result_bits, f.post.lo, f.post.hi);
  return f;
}
__device__ __forceinline__ int add_fast_apply(int a, int b, const AddFast& f) {
  return sc_apply(fx_addw(sc_apply(a, f.a), sc_apply(b, f.b)), f.post);
}

// Order-preserving non-negative image of a payload (|v| < 2^30: block inputs are saturated to <= 30 bits, checked on the host),
// so that per-column extremes can be accumulated with atomicMax on zero-initialised words (0 lies below every image) and
// travel through the same max-all-reduce as the float maxima, whether the reducer reads the words as signed or unsigned.
// A column's minimum is kept as the maximum of ~v (order reversing, no overflow).
__host__ __device__ __forceinline__ unsigned ord_enc(int v) { return (unsigned)(v + 0x40000000); }
__host__ __device__ __forceinline__ int ord_dec(unsigned u) { return (int)u - 0x40000000; }

struct BnFast {
  AddFast s1, s4;
  int rs2, lo2, hi2, rs3, lo3, hi3;
  ShiftClamp u;   // change_cfg of the chain output to the mixer-input format (fxpmodel.py:620-624)
};

struct BnPlan {
  AddPlan p1;
  int e_in, e1, e2, rs2, e3, rs3, e4;
  AddPlan p4;
  int rb1, rb2, rb3, rb4;
  int bad;
  BnFast f;
};

// exponent plans of the BatchNorm chain up to stage `upto` (1 add mean, 2 mul invsq, 3 mul scale, 4 add bias)
__device__ inline void bn_make_plan(BnPlan& pl, const BnChain& c, const unsigned* fm, const int* ex, int upto) {
  pl.bad = 0;
  pl.e_in = ex[E_IN];
  pl.rb1 = max(c.x_bits, c.b_mean);
  pl.rb2 = max(pl.rb1, c.b_inv);
  pl.rb3 = c.has_scale ? max(pl.rb2, c.b_scale) : pl.rb2;
  pl.rb4 = c.has_bias ? max(pl.rb3, c.b_bias) : pl.rb3;
  pl.p1 = fx_plan_add(__uint_as_float(fm[M_S1]), __uint_as_float(fm[M_A1]), c.pm_mean, pl.e_in, c.e_mean, pl.rb1);
  pl.e1 = pl.p1.res_exp;
  pl.e2 = pl.e1; pl.rs2 = 0; pl.e3 = pl.e1; pl.rs3 = 0; pl.e4 = pl.e1; pl.p4 = pl.p1;
  if (upto >= 2) {
    pl.e2 = fx_plan_mul(__uint_as_float(fm[M_2]), pl.rb2);
    pl.rs2 = pl.e1 + c.e_inv - pl.e2;
    if (pl.rs2 < 0 || pl.rs2 > 31) pl.bad = 1;
    pl.e3 = pl.e2; pl.e4 = pl.e2;
  }
  if (upto >= 3 && c.has_scale) {
    pl.e3 = fx_plan_mul(__uint_as_float(fm[M_3]), pl.rb3);
    pl.rs3 = pl.e2 + c.e_scale - pl.e3;
    if (pl.rs3 < 0 || pl.rs3 > 31) pl.bad = 1;
    pl.e4 = pl.e3;
  }
  if (upto >= 4 && c.has_bias) {
    pl.p4 = fx_plan_add(__uint_as_float(fm[M_S4]), __uint_as_float(fm[M_A4]), c.pm_bias, pl.e3, c.e_bias, pl.rb4);
    pl.e4 = pl.p4.res_exp;
  }
  if (pl.p1.agg_bits > 32 || (upto >= 4 && c.has_bias && pl.p4.agg_bits > 32)) pl.bad = 1;
  const int sh[4] = {pl.p1.agg_exp - pl.e_in, pl.p1.agg_exp - c.e_mean, pl.p1.res_exp - pl.p1.agg_exp, pl.p1.agg_exp - pl.p1.res_exp};
  for (int i = 0; i < 4; ++i)
    if (sh[i] > 31) pl.bad = 1;
  pl.f.s1 = add_fast_make(c.x_bits, pl.e_in, c.b_mean, c.e_mean, pl.p1, pl.rb1);
  pl.f.rs2 = pl.rs2 & 31; sc_bounds(pl.rb2, pl.f.lo2, pl.f.hi2);
  pl.f.rs3 = pl.rs3 & 31; sc_bounds(pl.rb3, pl.f.lo3, pl.f.hi3);
  pl.f.s4 = add_fast_make(pl.rb3, pl.e3, c.b_bias, c.e_bias, pl.p4, pl.rb4);
  pl.f.u = sc_change_cfg(pl.rb4, pl.e4, c.u_bits, c.u_exp);
}

// the chain with the per-column parameters in registers
__device__ __forceinline__ int bn_apply_v(int v, int mm, int inv, int sc, int bi, const BnChain& c, const BnPlan& pl, int upto) {
  const BnFast& f = pl.f;
  v = add_fast_apply(v, mm, f.s1);
  if (upto >= 2) v = min(max(fx_mulw(v, inv) >> f.rs2, f.lo2), f.hi2);
  if (upto >= 3 && c.has_scale) v = min(max(fx_mulw(v, sc) >> f.rs3, f.lo3), f.hi3);
  if (upto >= 4 && c.has_bias) v = add_fast_apply(v, bi, f.s4);
  return v;
}
__device__ __forceinline__ int bn_apply(int v, int k, const BnChain& c, const BnPlan& pl, int upto) {
  return bn_apply_v(v, c.mmean[k], upto >= 2 ? c.inv[k] : 0, (upto >= 3 && c.has_scale) ? c.scale[k] : 0,
                    (upto >= 4 && c.has_bias) ? c.bias[k] : 0, c, pl, upto);
}
// the chain with the per-column operands of the two adds already aligned (sc_apply(param, f.b), once per column and launch)
__device__ __forceinline__ int bn_apply_pre(int v, int mm_al, int inv, int sc, int bi_al, const BnChain& c, const BnPlan& pl) {
  const BnFast& f = pl.f;
  v = sc_apply(fx_addw(sc_apply(v, f.s1.a), mm_al), f.s1.post);
  v = min(max(fx_mulw(v, inv) >> f.rs2, f.lo2), f.hi2);
  if (c.has_scale) v = min(max(fx_mulw(v, sc) >> f.rs3, f.lo3), f.hi3);
  if (c.has_bias) v = sc_apply(fx_addw(sc_apply(v, f.s4.a), bi_al), f.s4.post);
  return v;
}
// u = change_cfg(BatchNorm(x)) to the mixer-input format: the whole chain + the final format change
__device__ __forceinline__ int bn_to_u(int v, const BnPlan& pl) { return sc_apply(v, pl.f.u); }

}  // namespace s5
