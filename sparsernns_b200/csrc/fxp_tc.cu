// fxp_tc.cu - the integer (fixed-point) model on the tensor cores: exact int8 x int16 dense-tile products through
// tcgen05.mma.kind::i8 with integer epilogues, bit-exact with the reference's fixed-point model.
//
// Reference path (paths relative to the reference repo):
//   fxp_matmul                        sparseRNNs/fxparray.py:640-678   (int32 `@`, rshift FLOOR, saturate)
//   FxpDense.forward                  sparseRNNs/fxpmodel.py:331-366   (matmul + bias add at out_exp)
//   FxpSSM.forward                    sparseRNNs/fxpmodel.py:610-794   (B projection :631-644, C projection :746-783)
//   FxpSequenceLayer.forward          sparseRNNs/fxpmodel.py:1110-1161 (+ mult_gate :1075-1093, FxpSigmoid :70-144)
//   run_validation input quantise     sparseRNNs/fxprun.py:69-75
//
// The kernels follow the float dense-tile kernels of tc_model.cu (same operand layout, same tile schedule); what differs is
// that every conversion and epilogue is the reference's integer arithmetic (fxp_prims.cuh) instead of fake-quant floats.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "fxp_tc.cuh"
#include "tc.cuh"

namespace s5 {

constexpr int kFtThreads = 256;

struct FxpTcLinParams {
  FxpTcLinArgs a;
  int Kfull, Nfull, Kp, Np, NP, alias_out;
  const signed char* W8; const signed char* WN;
};

__device__ __forceinline__ int sat_lh(int v, int lo, int hi) { return min(max(v, lo), hi); }
// int -> float without the conversion pipe, exact for |v| < 2^22 (every value here is a saturated <= 16-bit payload)
__device__ __forceinline__ float i2f_small(int v) { return __uint_as_float(0x4B400000u + (uint32_t)v) - 12582912.0f; }

// Per 128-row tile (the skeleton of tc_linear_kernel, tc_model.cu):
//   stage   : rows -> mode-specific integer conversion -> 16-bit payloads -> signed high-byte / unsigned low-byte planes
//   MMA     : acc_hi = Ah * W^T, acc_lo = Al * W^T  (kind::i8, int32 in TMEM), NP output columns per pass
//   epilogue: acc = 256 acc_hi + acc_lo == the reference's int32 matmul; >> rshift, saturate, bias, saturate, ReLU
//             -> shared tile -> coalesced stores (+ float maxima for the next block's BatchNorm, ENC)
template <int MODE, bool VEC4, int KV, bool XN>
__global__ void __launch_bounds__(kFtThreads, 2) fxp_tc_linear_kernel(const __grid_constant__ FxpTcLinParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FxpTcLinArgs& a = p.a;
  const int Kp = p.Kp, Np = p.Np, NP = p.NP;
  const size_t a_bytes = tc_tile_bytes8(kTcRows, Kp);
  unsigned char* Wt = smem;
  unsigned char* Ah = Wt + tc_tile_bytes8(Np, Kp);
  unsigned char* Al = Ah + a_bytes;
  const int ld_out = NP + 4;
  int* outT = reinterpret_cast<int*>(p.alias_out ? Ah : Al + a_bytes);
  int2* cst = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(outT) +
                                      (p.alias_out ? max(2 * a_bytes, (size_t)kTcRows * ld_out * 4) : (size_t)kTcRows * ld_out * 4));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int extra_n[XN ? kTcRows : 1];
  __shared__ BnPlan pl;
  __shared__ float red[32];
  __shared__ int dyn[4];
  __shared__ unsigned scmax[MODE == LIN_ENC ? 512 : 1], scmin[MODE == LIN_ENC ? 512 : 1];   // Nfull <= 512 (fxp_tc_linear_build)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kFtThreads / 32;
  const bool cstat = MODE == LIN_ENC && a.next_cmax != nullptr;
  if (cstat)
    for (int i = tid; i < p.Nfull; i += kFtThreads) { scmax[i] = 0u; scmin[i] = 0u; }
  if (warp == 0) tc_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    if (MODE == LIN_BPROJ) {
      bn_make_plan(pl, a.c, a.fm, a.ex, 4);
      if (blockIdx.x == 0) {
        a.ex[E_1] = pl.e1; a.ex[E_2] = pl.e2; a.ex[E_3] = pl.e3; a.ex[E_4] = pl.e4;
        if (pl.bad) atomicOr(a.status, 1);
      }
    } else if (MODE == LIN_DEC) {
      // FxpDense.forward input check (fxpmodel.py:335-347) with a data-dependent input exponent
      const int e = a.ex[E_IN];
      const int change = (16 > a.dec_inp_bits) || (e > a.dec_inp_exp);
      const int e_eff = change ? a.dec_inp_exp : e;
      const int rs = e_eff + a.dec_w_exp - a.dec_out_exp;
      dyn[0] = change; dyn[1] = e; dyn[2] = rs & 31;
      if ((rs < 0 || rs > 31 || (change && e - a.dec_inp_exp > 31)) && blockIdx.x == 0) atomicOr(a.status, 1);
    }
  }
  tc_load_weights8(Wt, p.W8, Np, Kp);
  for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kFtThreads * 16) *reinterpret_cast<int4*>(Ah + i) = make_int4(0, 0, 0, 0);
  const float sm_next = a.next_fm ? exp2f((float)-a.next_e_mean) : 0.f;
  for (int n = tid; n < Np; n += kFtThreads) {
    const bool in = n < p.Nfull;
    cst[n] = make_int2((in && a.bias_aligned) ? a.bias_aligned[n] : 0,
                       (in && a.next_fm) ? __float_as_int(__int2float_rn(a.next_mmean[n]) * sm_next) : 0);
  }
  const uint32_t lbo = tc_lbo(kTcRows);
  int koff[KV][4];
  int wn[KV][4];
  bool wval[KV];
  uint32_t loff[KV];
  int mm[KV][4], iv[KV][4], sc[KV][4], bi[KV][4];
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    loff[i] = (uint32_t)(8 * i + (lane >> 2)) * lbo + 4u * (lane & 3);
    wval[i] = 128 * i + 4 * lane < Kp;
#pragma unroll
    for (int e = 0; e < 4; ++e) wn[i][e] = (XN && wval[i]) ? (int)p.WN[128 * i + 4 * lane + e] : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = VEC4 ? 128 * i + 4 * lane + e : 128 * i + lane + 32 * e;
      const bool in = k < p.Kfull;
      koff[i][e] = in ? k : (VEC4 ? e : p.Kfull - 1);
      const int kk = in ? k : 0;
      mm[i][e] = MODE == LIN_BPROJ ? a.c.mmean[kk] : 0;
      iv[i][e] = MODE == LIN_BPROJ ? a.c.inv[kk] : 0;
      sc[i][e] = (MODE == LIN_BPROJ && a.c.has_scale) ? a.c.scale[kk] : 0;
      bi[i][e] = (MODE == LIN_BPROJ && a.c.has_bias) ? a.c.bias[kk] : 0;
    }
  }
  const bool tail = p.Kfull > 128 * KV;                       // one more input column (encoder: 257 = 2 * 128 + 1)
  const float in_sc = exp2f((float)a.in_exp);
  const int olo = a.out_bits >= 32 ? INT_MIN : -(1 << (a.out_bits - 1)), ohi = a.out_bits >= 32 ? INT_MAX : (1 << (a.out_bits - 1)) - 1;
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  const int egrp = warp >> 2;
  const bool wvec = (p.Nfull & 3) == 0 && (NP & 3) == 0 && (!a.Y || (reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) &&
                    (!a.Yf || (reinterpret_cast<uintptr_t>(a.Yf) & 15) == 0);
  const float so = exp2f((float)-a.out_exp);
  float m0 = 0.f, m1 = 0.f;
  uint32_t phase = 0;
  __syncthreads();
  if (MODE == LIN_BPROJ) {                                    // the plan is in shared memory now: align the add operands once
#pragma unroll
    for (int i = 0; i < KV; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mm[i][e] = sc_apply(mm[i][e], pl.f.s1.b);
        bi[i][e] = sc_apply(bi[i][e], pl.f.s4.b);
      }
  }
  const int change = MODE == LIN_DEC ? dyn[0] : 0, e_dyn = MODE == LIN_DEC ? dyn[1] : 0;
  const int rs0 = MODE == LIN_DEC ? dyn[2] : a.rshift0, rs1 = MODE == LIN_DEC ? dyn[2] : a.rshift1;
  // the conversion of one raw 32-bit word (float or int32 payload) of input column (i, e) to the 16-bit matmul operand
  auto convert = [&](uint32_t raw, int i, int e) -> int {
    if (MODE == LIN_ENC) {
      if (a.x_is_float) return fx_sat(__float2int_rd(__uint_as_float(raw) * in_sc), a.in_bits);   // fxp_from_fp FLOOR
      return (int)raw;
    } else if (MODE == LIN_BPROJ) {
      int v = bn_apply_pre((int)raw, mm[i][e], iv[i][e], sc[i][e], bi[i][e], a.c, pl);
      return bn_to_u(v, pl);             // fxpmodel.py:620-624
    } else {
      return change ? fx_change_cfg((int)raw, 16, e_dyn, a.dec_inp_bits, a.dec_inp_exp, S5_ROUND_FLOOR) : (int)raw;
    }
  };
  const long long row_begin = a.rows * blockIdx.x / gridDim.x, row_end = a.rows * (blockIdx.x + 1) / gridDim.x;
  constexpr int RB = 8 / KV;
  const long long last_row = a.rows - 1;
  const uint32_t* X = reinterpret_cast<const uint32_t*>(a.X);
  uint32_t raw[RB][KV][4];
  uint32_t rawt[RB];
  bool have_first = false;
  auto load_batch = [&](long long r0, int rb) {
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const uint32_t* xr = X + (size_t)min(r0 + rb + q, last_row) * p.Kfull;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        if constexpr (VEC4) {
          const uint4 v = *reinterpret_cast<const uint4*>(xr + koff[i][0]);
          raw[q][i][0] = v.x; raw[q][i][1] = v.y; raw[q][i][2] = v.z; raw[q][i][3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) raw[q][i][e] = xr[koff[i][e]];
        }
      }
      rawt[q] = tail ? xr[128 * KV] : 0u;
    }
  };
  for (long long row0 = row_begin; row0 < row_end; row0 += kTcRows) {
    const int rv = (int)min((long long)kTcRows, row_end - row0);
    unsigned short* utile = (MODE == LIN_BPROJ && VEC4 && a.U16) ? a.U16 + (size_t)row0 * p.Kfull + 4 * lane : nullptr;
    for (int rb = warp * RB; rb < rv; rb += NW * RB) {
      if (!(have_first && rb == warp * RB)) load_batch(row0, rb);
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const uint32_t roff = (uint32_t)(rb + q) * 16u;
        int part = 0;
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          uint32_t b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int v = convert(raw[q][i][e], i, e);
            b[e] = (uint32_t)v;
            if constexpr (XN) part += (int)(short)v * wn[i][e];
          }
          uint32_t lo, hi;
          split_planes(b[0], b[1], b[2], b[3], lo, hi);
          if (wval[i]) {
            *reinterpret_cast<uint32_t*>(Ah + loff[i] + roff) = hi;
            *reinterpret_cast<uint32_t*>(Al + loff[i] + roff) = lo;
          }
          if constexpr (MODE == LIN_BPROJ && VEC4) {
            // the GLU kernel of the same block needs u again (D u, fxpmodel.py:771-776): 2 bytes per element from here are
            // cheaper than re-running the BatchNorm chain there
            if (utile && 128 * i + 4 * lane < p.Kfull && rb + q < rv)
              *reinterpret_cast<uint2*>(utile + (uint32_t)(rb + q) * (uint32_t)p.Kfull + 128 * i) =
                  make_uint2(__byte_perm(b[0], b[1], 0x5410), __byte_perm(b[2], b[3], 0x5410));
          }
        }
        if (tail && lane == 0) {
          const uint32_t bt = (uint32_t)convert(rawt[q], 0, 0);   // ENC only: no per-column parameters
          *reinterpret_cast<uint32_t*>(Ah + (size_t)(8 * KV) * lbo + roff) = (bt >> 8) & 0xFFu;
          *reinterpret_cast<uint32_t*>(Al + (size_t)(8 * KV) * lbo + roff) = bt & 0xFFu;
        }
        if constexpr (XN) {
          part = __reduce_add_sync(0xffffffffu, part);
          if (lane == 0) extra_n[rb + q] = part;
        }
      }
    }
    have_first = false;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    if (tid == 0) {
      const int np = min(NP, Np);
      tc_issue8(taddr, Ah, Wt, Kp, np, Np, true);
      tc_issue8(taddr + 128, Al, Wt, Kp, np, Np, false);
      tc_commit(&bar);
    }
    if (row0 + kTcRows < row_end) {
      load_batch(row0 + kTcRows, warp * RB);
      have_first = true;
    }
    for (int pass0 = 0; pass0 < Np; pass0 += NP) {
      const int np = min(NP, Np - pass0);
      const int npv = min(np, p.Nfull - pass0);
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      const int cw = np >> 1;
      const bool row_ok = erow < rv;
      for (int c0 = egrp * cw; c0 < (erow0 < rv ? (egrp + 1) * cw : 0); c0 += 8) {
        uint32_t vh[8], vl[8];
        tc_ld8(taddr + lane_base + c0, vh);
        tc_ld8(taddr + lane_base + 128 + c0, vl);
        tc_wait_ld();
        int out[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int4 c2 = *reinterpret_cast<const int4*>(cst + pass0 + c0 + j);   // (bias, next mean) of two columns
          const int acc0 = (int)(vh[j] * 256u + vl[j]), acc1 = (int)(vh[j + 1] * 256u + vl[j + 1]);
          int v0 = sat_lh(acc0 >> rs0, olo, ohi), v1 = sat_lh(acc1 >> rs1, olo, ohi);   // fxp_matmul, fxparray.py:662-676
          if (a.bias_aligned) {                                                         // fxp_add at out_exp
            v0 = sat_lh(fx_addw(v0, c2.x), olo, ohi);
            v1 = sat_lh(fx_addw(v1, c2.z), olo, ohi);
          }
          if (a.relu) { v0 = max(v0, 0); v1 = max(v1, 0); }
          out[j] = v0; out[j + 1] = v1;
          if (MODE == LIN_ENC && a.next_fm && row_ok) {
            const float f0 = i2f_small(v0) * so, f1 = i2f_small(v1) * so;
            m0 = fmaxf(m0, fmaxf(fabsf(f0 + __int_as_float(c2.y)), fabsf(f1 + __int_as_float(c2.w))));
            m1 = fmaxf(m1, fmaxf(fabsf(f0), fabsf(f1)));
          }
        }
        int* dst = outT + (size_t)erow * ld_out + c0;
        *reinterpret_cast<int4*>(dst) = make_int4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<int4*>(dst + 4) = make_int4(out[4], out[5], out[6], out[7]);
      }
      tc_fence_before();
      __syncthreads();
      if (tid == 0 && pass0 + NP < Np) {
        tc_fence_after();
        const int np2 = min(NP, Np - pass0 - NP);
        tc_issue8(taddr, Ah, Wt + (size_t)(pass0 + NP) * 16, Kp, np2, Np, true);
        tc_issue8(taddr + 128, Al, Wt + (size_t)(pass0 + NP) * 16, Kp, np2, Np, false);
        tc_commit(&bar);
      }
      // ---- write-out, coalesced
      if (wvec) {
        const int q4 = npv >> 2;
        const int rpi = q4 >= 32 ? 1 : 32 / q4;
        const int lr = lane / q4, lc = lane - lr * q4;
        if (lr < rpi) {
          int cmx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN}, cmn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
          for (int rb = warp * rpi + lr; rb < rv; rb += 4 * NW * rpi) {
            int4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const int4*>(outT + (size_t)min(rb + q * NW * rpi, kTcRows - 1) * ld_out + 4 * lc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (rb + q * NW * rpi < rv) {
                if (cstat) {
                  cmx[0] = max(cmx[0], v[q].x); cmx[1] = max(cmx[1], v[q].y); cmx[2] = max(cmx[2], v[q].z); cmx[3] = max(cmx[3], v[q].w);
                  cmn[0] = min(cmn[0], v[q].x); cmn[1] = min(cmn[1], v[q].y); cmn[2] = min(cmn[2], v[q].z); cmn[3] = min(cmn[3], v[q].w);
                }
                const size_t o = (size_t)(row0 + rb + q * NW * rpi) * p.Nfull + pass0 + 4 * lc;
                if (a.Y) *reinterpret_cast<int4*>(a.Y + o) = v[q];
                if (a.Yf)
                  *reinterpret_cast<float4*>(a.Yf + o) = make_float4(__int2float_rn(v[q].x) * a.out_scale, __int2float_rn(v[q].y) * a.out_scale,
                                                                     __int2float_rn(v[q].z) * a.out_scale, __int2float_rn(v[q].w) * a.out_scale);
              }
          }
          if (cstat && cmx[0] >= cmn[0]) {                       // this lane saw at least one row
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              atomicMax(&scmax[pass0 + 4 * lc + e], ord_enc(cmx[e]));
              atomicMax(&scmin[pass0 + 4 * lc + e], ord_enc(~cmn[e]));
            }
          }
        }
      } else {
        for (int rb = warp * 4; rb < rv; rb += NW * 4) {
          int v[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[q][j] = outT[(size_t)(rb + q) * ld_out + lane + 32 * j];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const size_t o = (size_t)(row0 + rb + q) * p.Nfull + pass0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + q < rv && lane + 32 * j < npv) {
                if (cstat) {
                  atomicMax(&scmax[pass0 + lane + 32 * j], ord_enc(v[q][j]));
                  atomicMax(&scmin[pass0 + lane + 32 * j], ord_enc(~v[q][j]));
                }
                if (a.Y) a.Y[o + lane + 32 * j] = v[q][j];
                if (a.Yf) a.Yf[o + lane + 32 * j] = __int2float_rn(v[q][j]) * a.out_scale;
              }
          }
        }
      }
      if constexpr (XN) {
        if (pass0 + NP >= Np && tid < rv) {
          const int n = p.Nfull - 1;
          int v = sat_lh(extra_n[tid] >> ((n & 1) ? rs1 : rs0), olo, ohi);
          if (a.bias_aligned) v = sat_lh(fx_addw(v, a.bias_aligned[n]), olo, ohi);
          if (a.relu) v = max(v, 0);
          const size_t o = (size_t)(row0 + tid) * p.Nfull + n;
          if (a.Y) a.Y[o] = v;
          if (a.Yf) a.Yf[o] = __int2float_rn(v) * a.out_scale;
        }
      }
      __syncthreads();
    }
  }
  if (MODE == LIN_ENC && a.next_fm) {
    block_max_to(a.next_fm + M_S1, m0, red);
    block_max_to(a.next_fm + M_A1, m1, red);
  }
  if (cstat) {
    __syncthreads();
    for (int i = tid; i < p.Nfull; i += kFtThreads) {
      if (scmax[i]) atomicMax(a.next_cmax + i, scmax[i]);
      if (scmin[i]) atomicMax(a.next_cmin + i, scmin[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 256);
}

// ------------------------------------------------------------------------------------------- host side
int fxp_tc_linear_build(const int32_t* W, int Kfull, int Nfull, FxpTcLinearDev* out) {
  out->ok = false;
  const bool tail = Kfull > 128 && (Kfull % 128) == 1;
  const int KV = tail ? Kfull / 128 : (Kfull + 127) / 128;
  if (KV < 1 || KV > 2 || Kfull < 16 || Nfull < 16 || Nfull > 512) return S5_OK;
  const int Kp[2] = {(Kfull + 31) & ~31, 128 * KV + (tail ? 32 : 0)};
  const bool xn = (Nfull % 16) == 1 && Nfull > 16 && !tail;
  const int Np = xn ? Nfull - 1 : (Nfull + 15) & ~15;
  auto perm = [&](int k) { return k >= 128 * KV ? k : (k & ~127) + 4 * (k & 31) + ((k & 127) >> 5); };
  std::vector<signed char> w8[2], wnx[2];
  for (int o = 0; o < 2; ++o) {
    w8[o].assign((size_t)Np * Kp[o], 0);
    wnx[o].assign((size_t)Kp[o] + 256, 0);
  }
  for (int n = 0; n < Nfull; ++n)
    for (int k = 0; k < Kfull; ++k) {
      const int c = W[(size_t)k * Nfull + n];
      if (c < -128 || c > 127) return S5_OK;                  // not an 8-bit payload: CUDA-core kernels
      if (n < Np) {
        w8[0][(size_t)n * Kp[0] + k] = (signed char)c;
        w8[1][(size_t)n * Kp[1] + perm(k)] = (signed char)c;
      } else {
        wnx[0][k] = (signed char)c;
        wnx[1][perm(k)] = (signed char)c;
      }
    }
  out->Kfull = Kfull; out->Nfull = Nfull; out->Kp[0] = Kp[0]; out->Kp[1] = Kp[1]; out->Np = Np; out->KV = KV; out->xn = xn;
  for (int o = 0; o < 2; ++o) {
    S5_CUDA(cudaMalloc(&out->W8[o], w8[o].size()));
    S5_CUDA(cudaMemcpy(out->W8[o], w8[o].data(), w8[o].size(), cudaMemcpyHostToDevice));
    S5_CUDA(cudaMalloc(&out->WN[o], wnx[o].size()));
    S5_CUDA(cudaMemcpy(out->WN[o], wnx[o].data(), wnx[o].size(), cudaMemcpyHostToDevice));
  }
  out->ok = true;
  return S5_OK;
}

void fxp_tc_linear_free(FxpTcLinearDev* d) {
  cudaFree(d->W8[0]); cudaFree(d->W8[1]); cudaFree(d->WN[0]); cudaFree(d->WN[1]);
  *d = FxpTcLinearDev();
}

template <int MODE, bool VEC4, int KV, bool XN>
static int fxp_tc_linear_go(const FxpTcLinParams& p0, int kclass, cudaStream_t st) {
  FxpTcLinParams p = p0;
  auto kernel = fxp_tc_linear_kernel<MODE, VEC4, KV, XN>;
  const size_t w_bytes = tc_tile_bytes8(p.Np, p.Kp), a2 = 2 * tc_tile_bytes8(kTcRows, p.Kp);
  size_t smem = 0;
  auto fits2 = [&](size_t b) { return 2 * (b + 2048 + 64) <= (size_t)227 * 1024; };
  auto smem_for = [&](int NP) {
    p.NP = std::min(NP, p.Np);
    p.alias_out = p.NP >= p.Np;
    const size_t out_bytes = (size_t)kTcRows * (p.NP + 4) * 4;
    return w_bytes + (p.alias_out ? std::max(a2, out_bytes) : a2 + out_bytes) + (size_t)p.Np * sizeof(int2);
  };
  bool two = false;
  for (int NP : {128, 64}) {
    smem = smem_for(NP);
    if (fits2(smem)) { two = true; break; }
  }
  if (!two)
    for (int NP : {128, 64, 32, 16}) {
      smem = smem_for(NP);
      if (smem + 2048 + 64 <= (size_t)227 * 1024) break;
    }
  if (smem + 2048 + 64 > 227 * 1024) S5_FAIL(S5_E_BADSHAPE, "fxp dense-tile linear: tile does not fit in shared memory");
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int per_sm = two ? 2 : 1;
  const int grid = (int)std::max<long long>(1, std::min<long long>((p.a.rows + 31) / 32, (long long)per_sm * sm_count()));
  ProfScope prof(kclass, st);
  kernel<<<grid, kFtThreads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

int fxp_tc_linear_launch(const FxpTcLinearDev& d, const FxpTcLinArgs& a, int kclass, cudaStream_t st) {
  if (!d.ok) S5_FAIL(S5_E_BADARG, "fxp dense-tile linear: not eligible");
  if (a.rows <= 0) S5_FAIL(S5_E_BADSHAPE, "fxp dense-tile linear: rows=%lld must be positive", a.rows);
  FxpTcLinParams p{};
  p.a = a; p.Kfull = d.Kfull; p.Nfull = d.Nfull; p.Np = d.Np;
  const bool vec4 = (d.Kfull & 3) == 0 && aligned16(a.X);
  p.W8 = d.W8[vec4 ? 0 : 1]; p.WN = d.WN[vec4 ? 0 : 1]; p.Kp = d.Kp[vec4 ? 0 : 1];
  if (d.xn && !vec4) S5_FAIL(S5_E_BADARG, "fxp dense-tile linear: extra output column needs aligned rows");
#define FT_GO(MODE_)                                                                                                \
  do {                                                                                                               \
    if (d.xn) return d.KV == 1 ? fxp_tc_linear_go<MODE_, true, 1, true>(p, kclass, st) : fxp_tc_linear_go<MODE_, true, 2, true>(p, kclass, st); \
    if (vec4) return d.KV == 1 ? fxp_tc_linear_go<MODE_, true, 1, false>(p, kclass, st) : fxp_tc_linear_go<MODE_, true, 2, false>(p, kclass, st); \
    return d.KV == 1 ? fxp_tc_linear_go<MODE_, false, 1, false>(p, kclass, st) : fxp_tc_linear_go<MODE_, false, 2, false>(p, kclass, st); \
  } while (0)
  switch (a.mode) {
    case LIN_ENC: FT_GO(LIN_ENC);
    case LIN_BPROJ: FT_GO(LIN_BPROJ);
    case LIN_DEC: FT_GO(LIN_DEC);
    default: S5_FAIL(S5_E_BADARG, "fxp dense-tile linear: unknown mode %d", a.mode);
  }
#undef FT_GO
}


// ------------------------------------------------------------------------------------------- C projection + GLU
// The tail of a fixed-point S5 block on the tensor cores (the integer twin of tc_glu_kernel, tc_model.cu).  One CTA of 16 warps per
// SM, all 512 TMEM columns.  Per tile of <= 128 frame rows:
//   S   : states (int32 pairs re, im) -> signed high-byte / unsigned low-byte planes.  The recurrence does not saturate
//         (fxpmodel.py:147-208), so a state may need more than 16 bits: such a tile is NOT processed here - it is appended to
//         a list and left to the CUDA-core kernel (fxp_glu_kernel), which multiplies full int32 values.  Exact either way.
//   MMA1: [hi | lo] x [C_re on the even k ; C_im on the odd k]^T -> four int32 accumulators (N = 2H per plane)
//   E1  : y1 = sat(acc_re >> rs), y2 = sat(acc_im >> rs), ys = 2 sat(y1 - y2)  -> int32 tile T          (TMEM lane = row)
//   P1  : lanes along columns, x read coalesced: u = change_cfg(BatchNorm(x)); y = sat(ys + D u); x1 = relu(y);
//         left-gate payload -> T (16 bit); out2 input payload -> int8 planes
//   MMA2: out2 (kind::i8)
//   E2  : o = sat(sat(acc >> rs) + bias); sigmoid table; right-gate payload -> T (16 bit)
//   P2  : g = sat((l * r) >> rs) -> G, coalesced; float maxima for the residual's compute_best add
constexpr int kFgThreads = 512;
constexpr int kFgK2 = 128;

struct GluFast {
  int ylo, yhi, olo, ohi, glo, ghi;
  ShiftClamp oin, gl, sx, gr;
};

struct FxpTcGluParams {
  FxpTcGluArgs a;
  int H, K1;
  const signed char* Wc; const signed char* Wo;
};

template <int HJ>
__global__ void __launch_bounds__(kFgThreads, 1) fxp_tc_glu_kernel(const __grid_constant__ FxpTcGluParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FxpTcGluArgs& a = p.a;
  constexpr int H = 32 * HJ;
  const int K1 = p.K1;
  constexpr int ldT = H + 4;
  const size_t wc_bytes = tc_tile_bytes8(2 * H, K1), wo_bytes = tc_tile_bytes8(H, kFgK2);
  const size_t pl_bytes = tc_tile_bytes8(kTcRows, kFgK2);       // one int8 plane (K1 <= 128)
  unsigned char* Wc = smem;
  unsigned char* Wo = Wc + wc_bytes;
  unsigned char* Ph = Wo + wo_bytes;                            // state planes, later the out2 input planes
  unsigned char* Pl = Ph + pl_bytes;
  int* T = reinterpret_cast<int*>(Pl + pl_bytes);               // [128][H + 4]
  int* ob = T + (size_t)kTcRows * ldT;                          // [H] out2 bias at out_exp
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ BnPlan pl;
  __shared__ GluFast gf;
  __shared__ float red[32];
  __shared__ int lut_s[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kFgThreads / 32;
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    bn_make_plan(pl, a.c, a.fm, a.ex, 4);
    sc_bounds(a.y_bits, gf.ylo, gf.yhi);
    sc_bounds(a.o_out_bits, gf.olo, gf.ohi);
    sc_bounds(a.g_bits, gf.glo, gf.ghi);
    gf.oin = a.o_in_change ? sc_change_cfg(a.y_bits, a.y_exp, a.o_inp_bits, a.o_inp_exp) : ShiftClamp{0, 0, INT_MIN, INT_MAX};
    gf.gl = sc_change_cfg(a.y_bits, a.y_exp, a.gl_bits, a.gl_exp);
    gf.sx = sc_change_cfg(a.o_out_bits, a.o_out_exp, a.o_out_bits, a.sig_x_exp);    // fx_change_exp
    gf.gr = sc_change_cfg(a.o_out_bits, a.sig_y_exp, a.gr_bits, a.gr_exp);
  }
  if (tid < 8) lut_s[tid] = a.lut[tid];
  tc_load_weights8(Wc, p.Wc, 2 * H, K1);
  tc_load_weights8(Wo, p.Wo, H, kFgK2);
  for (size_t i = (size_t)tid * 16; i < 2 * pl_bytes; i += kFgThreads * 16) *reinterpret_cast<int4*>(Ph + i) = make_int4(0, 0, 0, 0);
  for (int n = tid; n < H; n += kFgThreads) ob[n] = a.o_bias_aligned[n];
  const uint32_t lbo = tc_lbo(kTcRows);
  const bool kin = 4 * lane < K1;
  const int xs_col = kin ? 4 * lane : 0;
  const uint32_t aoff = (uint32_t)(lane >> 2) * lbo + 4u * (lane & 3);
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  constexpr int cw = H / 4;                                     // columns per warp in the TMEM passes (4 column groups)
  const int cbeg = (warp >> 2) * cw;
  // per-lane column constants of the lanes-along-columns passes
  int mm[HJ], iv[HJ], sc[HJ], bi[HJ], dd[HJ];
#pragma unroll
  for (int j = 0; j < HJ; ++j) {
    const int k = lane + 32 * j;
    mm[j] = a.c.mmean[k]; iv[j] = a.c.inv[k];
    sc[j] = a.c.has_scale ? a.c.scale[k] : 0; bi[j] = a.c.has_bias ? a.c.bias[k] : 0;
    dd[j] = a.D[k];
  }
  uint32_t phase = 0;
  __syncthreads();
  const uint32_t tcol = tmem_slot;
  const float sg = exp2f((float)-a.g_exp), sk = exp2f((float)-pl.e_in);
  const int rs_cre = a.rs_cre, rs_cim = a.rs_cim, rs_D = a.rs_D, o_rs = a.o_rs, g_rs = a.g_rs, conj = a.conj;
  const int ylo = gf.ylo, yhi = gf.yhi;
  float m0 = 0.f, m1 = 0.f;
  // balanced partition: every CTA owns one contiguous row range, cut into equal tiles of at most 128 rows
  const long long row_begin = a.rows * blockIdx.x / gridDim.x, row_end = a.rows * (blockIdx.x + 1) / gridDim.x;
  const int nt = (int)((row_end - row_begin + kTcRows - 1) / kTcRows);
  const int th = nt ? (int)((((row_end - row_begin) + nt - 1) / nt + 7) & ~7LL) : kTcRows;   // tile height, multiple of 8
  const long long last_row = a.rows - 1;
  constexpr int PB = 8;
  for (long long row0 = row_begin; row0 < row_end; row0 += th) {
    const int rv = (int)min((long long)th, row_end - row0);
    // ---- S: 8 rows per warp, one 16-byte load per lane and row
    int bad = 0;
    const int rb_s = warp * PB;
    int4 raw[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q)
      raw[q] = *reinterpret_cast<const int4*>(a.XS + (size_t)min(row0 + rb_s + q, last_row) * K1 + xs_col);
    if (kin && rb_s < rv) {
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        int4 v = raw[q];
        if (rb_s + q < rv)
          bad |= (int)((((uint32_t)v.x + 0x8000u) | ((uint32_t)v.y + 0x8000u) | ((uint32_t)v.z + 0x8000u) | ((uint32_t)v.w + 0x8000u)) >> 16);
        if (a.relu_in) {
          // fxp_relu on the complex state (fxpmodel.py:30-45, :740-742) when the scan left it to this kernel: lexicographic
          // test; the float32 round trip of the kept values is the identity below 2^24 (and a tile with such a value is bad)
          const bool k0 = v.x > 0 || (v.x == 0 && v.y > 0), k1 = v.z > 0 || (v.z == 0 && v.w > 0);
          v.x = k0 ? v.x : 0; v.y = k0 ? v.y : 0; v.z = k1 ? v.z : 0; v.w = k1 ? v.w : 0;
        }
        uint32_t lo, hi;
        split_planes((uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w, lo, hi);
        *reinterpret_cast<uint32_t*>(Ph + aoff + (uint32_t)(rb_s + q) * 16u) = hi;
        *reinterpret_cast<uint32_t*>(Pl + aoff + (uint32_t)(rb_s + q) * 16u) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    bad = __syncthreads_or(bad);
    tc_fence_after();
    if (bad) {                                                  // CTA-uniform: leave the tile to the CUDA-core kernel
      if (a.relu_in && kin) {
        // that kernel expects ReLU-ed states in memory: write this tile back with the exact fxp_relu (float round trip)
#pragma unroll
        for (int q = 0; q < PB; ++q)
          if (rb_s + q < rv) {
            int4 v = raw[q];
            fx_relu_complex(v.x, v.y);
            fx_relu_complex(v.z, v.w);
            *reinterpret_cast<int4*>(a.XS + (size_t)(row0 + rb_s + q) * K1 + xs_col) = v;
          }
      }
      if (tid == 0) {
        const int slot = atomicAdd(a.bad_count, 1);
        a.bad_list[2 * slot] = (int)row0;
        a.bad_list[2 * slot + 1] = rv;
      }
      continue;
    }
    if (tid == 0) {
      tc_issue8(tcol, Ph, Wc, K1, 2 * H, 2 * H, true);
      tc_issue8(tcol + 256, Pl, Wc, K1, 2 * H, 2 * H, false);
      tc_commit(&bar);
    }
    // block-input rows of this warp's P1 batch travel while the tensor core works
    const int* xtile = a.X + (size_t)row0 * H;
    const int rlim = (int)min((long long)kTcRows, last_row - row0);
    int xin[PB][HJ];
    unsigned short uin[PB][HJ];
    if (a.U16) {                                                // u from the B projection (2 bytes per element)
      const unsigned short* utile = a.U16 + (size_t)row0 * H;
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const unsigned short* ur = utile + (size_t)min(warp * PB + q, rlim) * H;
#pragma unroll
        for (int j = 0; j < HJ; ++j) uin[q][j] = ur[lane + 32 * j];
      }
    } else {
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int* xr = xtile + (size_t)min(warp * PB + q, rlim) * H;
#pragma unroll
        for (int j = 0; j < HJ; ++j) { xin[q][j] = xr[lane + 32 * j]; uin[q][j] = 0; }
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- E1: the four accumulators -> ys (fxpmodel.py:746-767)
    if (erow0 < rv) {
#pragma unroll
      for (int c0 = 0; c0 < cw; c0 += 8) {
        const int c = cbeg + c0;
        uint32_t rh[8], rl[8], ih[8], il[8];
        tc_ld8(tcol + lane_base + c, rh);
        tc_ld8(tcol + lane_base + 256 + c, rl);
        tc_ld8(tcol + lane_base + H + c, ih);
        tc_ld8(tcol + lane_base + 256 + H + c, il);
        tc_wait_ld();
        int o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int a1 = (int)(rh[j] * 256u + rl[j]), a2 = (int)(ih[j] * 256u + il[j]);
          const int y1 = sat_lh(a1 >> rs_cre, ylo, yhi), y2 = sat_lh(a2 >> rs_cim, ylo, yhi);
          int ys = sat_lh(fx_addw(y1, fx_mulw(y2, -1)), ylo, yhi);
          if (conj) ys = fx_mulw(ys, 2);
          o[j] = ys;
        }
        int* dst = T + (size_t)erow * ldT + c;
        *reinterpret_cast<int4*>(dst) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4*>(dst + 4) = make_int4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- P1: one warp per row, PB rows per warp
    {
      const int rb = warp * PB;
      if (rb < rv) {
        int tin[PB][HJ];
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
          for (int j = 0; j < HJ; ++j) tin[q][j] = T[(size_t)(rb + q) * ldT + lane + 32 * j];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const int r = rb + q;
          uint32_t ob4[4] = {0u, 0u, 0u, 0u};
          unsigned short* trow = reinterpret_cast<unsigned short*>(T + (size_t)r * ldT);
#pragma unroll
          for (int j = 0; j < HJ; ++j) {
            const int u = a.U16 ? (int)(short)uin[q][j] : bn_to_u(bn_apply_v(xin[q][j], mm[j], iv[j], sc[j], bi[j], a.c, pl, 4), pl);
            const int du = sat_lh(fx_mulw(dd[j], u) >> rs_D, ylo, yhi);                        // fxpmodel.py:771-776
            const int y = sat_lh(fx_addw(tin[q][j], du), ylo, yhi);                            // :778-783
            const int x1 = max(y, 0);                                                          // fxp_relu :1125
            trow[lane + 32 * j] = (unsigned short)sc_apply(x1, gf.gl);                         // left gate (>= 0, <= 16 bits)
            ob4[j] = (uint32_t)sc_apply(x1, gf.oin);                                           // out2 input (>= 0, <= 16 bits)
          }
          uint32_t lo, hi;
          split_planes(ob4[0], ob4[1], ob4[2], ob4[3], lo, hi);
          *reinterpret_cast<uint32_t*>(Ph + aoff + (uint32_t)r * 16u) = hi;
          *reinterpret_cast<uint32_t*>(Pl + aoff + (uint32_t)r * 16u) = lo;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      tc_issue8(tcol, Ph, Wo, kFgK2, H, H, true);
      tc_issue8(tcol + 256, Pl, Wo, kFgK2, H, H, false);
      tc_commit(&bar);
    }
    if (a.U16) {                                                // the skip rows (maxima of the residual add) travel during MMA 2 / E2
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int* xr = xtile + (size_t)min(warp * PB + q, rlim) * H;
#pragma unroll
        for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- E2: out2 epilogue, sigmoid table, right-gate payload -> second half of the T row
    if (erow0 < rv) {
#pragma unroll
      for (int c0 = 0; c0 < cw; c0 += 8) {
        const int c = cbeg + c0;
        uint32_t vh[8], vl[8];
        tc_ld8(tcol + lane_base + c, vh);
        tc_ld8(tcol + lane_base + 256 + c, vl);
        tc_wait_ld();
        uint32_t sb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int acc = (int)(vh[j] * 256u + vl[j]);
          int ov = sat_lh(acc >> o_rs, gf.olo, gf.ohi);
          ov = sat_lh(fx_addw(ov, ob[c + j]), gf.olo, gf.ohi);                                 // FxpDense :352-364
          const int xx = sc_apply(ov, gf.sx);
          const int r = fx_sigmoid(xx, a.sig_x_exp, a.sig_y_exp, lut_s);
          sb[j] = (uint32_t)sc_apply(r, gf.gr);
        }
        uint4 pk;
        pk.x = __byte_perm(sb[0], sb[1], 0x5410); pk.y = __byte_perm(sb[2], sb[3], 0x5410);
        pk.z = __byte_perm(sb[4], sb[5], 0x5410); pk.w = __byte_perm(sb[6], sb[7], 0x5410);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(T + (size_t)erow * ldT) + 2 * H + 2 * c) = pk;
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- P2: g = mult_gate (fxpmodel.py:1075-1093), coalesced stores, maxima for the residual add
    {
      const int rb = warp * PB;
      const short* trow = reinterpret_cast<const short*>(T + (size_t)rb * ldT) + lane;
      int* gr = a.G + (size_t)(row0 + rb) * H + lane;
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        if (rb + q < rv) {
#pragma unroll
          for (int j = 0; j < HJ; ++j) {
            const int l = trow[(size_t)q * 2 * ldT + 32 * j], r = trow[(size_t)q * 2 * ldT + H + 32 * j];
            const int g = sat_lh(fx_mulw(l, r) >> g_rs, gf.glo, gf.ghi);
            gr[(size_t)q * H + 32 * j] = g;
            const float fg = __int2float_rn(g) * sg;
            m0 = fmaxf(m0, fabsf(fg + __int2float_rn(xin[q][j]) * sk));
            m1 = fmaxf(m1, fabsf(fg));
          }
        }
      }
    }
    __syncthreads();
  }
  block_max_to(a.fm + M_SR, m0, red);
  block_max_to(a.fm + M_G, m1, red);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

int fxp_tc_glu_build(const int32_t* C_re, const int32_t* C_im, const int32_t* out2_w, int H, int P, FxpTcGluDev* out) {
  out->ok = false;
  const int K1 = 2 * P;
  if (H % 32 != 0 || H < 32 || H > 128 || K1 % 32 != 0 || K1 < 32 || K1 > 128) return S5_OK;   // shape not covered
  std::vector<signed char> wc((size_t)2 * H * K1, 0), wo((size_t)H * kFgK2, 0);
  for (int h = 0; h < H; ++h)
    for (int q = 0; q < P; ++q) {
      const int cr = C_re[(size_t)h * P + q], ci = C_im[(size_t)h * P + q];
      if (cr < -128 || cr > 127 || ci < -128 || ci > 127) return S5_OK;
      wc[(size_t)h * K1 + 2 * q] = (signed char)cr;
      wc[(size_t)(H + h) * K1 + 2 * q + 1] = (signed char)ci;
    }
  for (int k = 0; k < H; ++k)
    for (int n = 0; n < H; ++n) {
      const int c = out2_w[(size_t)k * H + n];
      if (c < -128 || c > 127) return S5_OK;
      wo[(size_t)n * kFgK2 + 4 * (k & 31) + (k >> 5)] = (signed char)c;
    }
  S5_CUDA(cudaMalloc(&out->Wc, wc.size()));
  S5_CUDA(cudaMemcpy(out->Wc, wc.data(), wc.size(), cudaMemcpyHostToDevice));
  S5_CUDA(cudaMalloc(&out->Wo, wo.size()));
  S5_CUDA(cudaMemcpy(out->Wo, wo.data(), wo.size(), cudaMemcpyHostToDevice));
  out->H = H; out->K1 = K1;
  out->smem = tc_tile_bytes8(2 * H, K1) + tc_tile_bytes8(H, kFgK2) + 2 * tc_tile_bytes8(kTcRows, kFgK2) +
              (size_t)kTcRows * (H + 4) * 4 + (size_t)H * 4;
  out->ok = out->smem + 4096 <= 227 * 1024;
  return S5_OK;
}

void fxp_tc_glu_free(FxpTcGluDev* d) {
  cudaFree(d->Wc); cudaFree(d->Wo);
  *d = FxpTcGluDev();
}

template <int HJ>
static int fxp_tc_glu_go(const FxpTcGluParams& p, size_t smem, cudaStream_t st) {
  auto kernel = fxp_tc_glu_kernel<HJ>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<long long>(1, std::min<long long>((p.a.rows + 63) / 64, (long long)sm_count()));
  ProfScope prof(S5_K_FXP_GLU, st);
  kernel<<<grid, kFgThreads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

int fxp_tc_glu_launch(const FxpTcGluDev& d, const FxpTcGluArgs& a, cudaStream_t st) {
  if (!d.ok) S5_FAIL(S5_E_BADARG, "fxp dense-tile GLU: not eligible");
  if (a.rows <= 0 || a.rows >= (1ll << 31)) S5_FAIL(S5_E_BADSHAPE, "fxp dense-tile GLU: rows=%lld out of range", a.rows);
  if (!aligned16(a.XS)) S5_FAIL(S5_E_BADARG, "fxp dense-tile GLU: states must be 16-byte aligned");
  FxpTcGluParams p{};
  p.a = a; p.H = d.H; p.K1 = d.K1; p.Wc = d.Wc; p.Wo = d.Wo;
  switch (d.H / 32) {
    case 1: return fxp_tc_glu_go<1>(p, d.smem, st);
    case 2: return fxp_tc_glu_go<2>(p, d.smem, st);
    case 3: return fxp_tc_glu_go<3>(p, d.smem, st);
    default: return fxp_tc_glu_go<4>(p, d.smem, st);
  }
}

}  // namespace s5
