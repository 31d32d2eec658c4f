// tc_model.cu - dense-tile (tcgen05) projection kernels of the static w8a16 flavour, for throughput batches.
//
// Why dense tiles: at rows = B*L >= 1e5 the projections are genuine [rows x K] x [K x N] contractions with tiny
// weight matrices; the CUDA-core compressed kernels are instruction bound there (profiles/r01_*), the tensor cores
// are not.  Exactness: static fake-quant values are integer codes times power-of-two scales.  int8 weight codes and
// the hi/lo bytes of int16 activation codes are exact in f16, their products and the running sums (< 2^24 per
// byte plane) are exact in the fp32 TMEM accumulators, and  acc = 256 * acc_hi + acc_lo  is recombined in int32 -
// the tensor cores reproduce the exact integer sum; a single float rounding follows, as in the reference's fp32 dot.
//
//   tc_linear_kernel : encoder Dense / BatchNorm + B projection / decoder Dense of the reference
//                      (quantization.py:684-816; ssm.py:230-246; layers.py:210-211)
//   tc_glu_kernel    : C projection + D feed-through + ReLU + out2 Dense + sigmoid + mult_gate + skip + ReLU
//                      (ssm.py:332-337, 594-601; layers.py:217-238)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "fxp_chain.cuh"
#include "proj.cuh"
#include "tc.cuh"
#include <cuda_bf16.h>

#include "tc_model.cuh"

namespace s5 {

// -DS5_TC_TRACE: thread 0 of CTA 0 prints the phase timeline (ns, %globaltimer) of its second tile; debugging only
#ifdef S5_TC_TRACE
#include <cstdio>
__device__ int g_trace_budget = 12;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE_DECL unsigned long long tr_[20]; int tr_n = 0; int tr_tile = 0; const unsigned long long tr_entry = tr_t0_; unsigned long long tr_loop = gtime();
#define TC_TRACE_ENTRY const unsigned long long tr_t0_ = gtime();
#define TC_TRACE_MARK(cond) do { if ((cond) && blockIdx.x == 0 && tr_tile == 1 && tr_n < 20) tr_[tr_n++] = gtime(); } while (0)
#define TC_TRACE_NEXT() do { ++tr_tile; } while (0)
#define TC_TRACE_PRINT(cond, name) do { if ((cond) && blockIdx.x == 0 && tr_n > 1 && atomicSub(&g_trace_budget, 1) > 0) { \
    printf("[trace %s] prologue %llu total %llu tiles %d |", name, tr_loop - tr_entry, gtime() - tr_entry, tr_tile); for (int i_ = 1; i_ < tr_n; ++i_) printf(" %llu", tr_[i_] - tr_[i_ - 1]); printf(" ns\n"); } } while (0)
#else
#define TC_TRACE_DECL
#define TC_TRACE_ENTRY
#define TC_TRACE_MARK(cond)
#define TC_TRACE_NEXT()
#define TC_TRACE_PRINT(cond, name)
#endif

#ifndef S5_LOOKBACK_STRIDE
#define S5_LOOKBACK_STRIDE 4   // every S5_LOOKBACK_STRIDE-th tile of a clip publishes its inclusive prefix (fused recurrence)
#endif
constexpr int kTcThreads = 256;
constexpr int kTcKI = 8;  // K <= 256: each lane stages at most 8 input columns

struct TcLinParams {
  const float* X; float* Y; long long rows;
  int Kfull, Nfull;         // matrix dims
  int Kp, Np;               // padded dims of the operand tiles: Kp multiple of 32, Np multiple of 16
  int NP;                   // output columns per pass (MMA N), multiple of 16, <= 128
  int alias_out;            // single pass: the output tile may reuse the (dead) activation planes
  int x_stride, y_stride;   // floats per input / output row
  const signed char* W8;    // [Np][Kp] int8 weight codes, K in the physical (staging) order, zero padded
  const signed char* WN;    // XN: [Kp] codes of the one output column beyond Np (decoder: 257 = 256 + 1), same K order
  const float* in_sub; const float* in_mul; const float* in_add;  // BatchNorm on the input (BN instantiation only)
  float in_inv, in_qmin, in_qmax;   // 1 / s_in (s_in is a power of two, so x * in_inv == x / s_in exactly)
  const float* col_scale;   // [Nfull] s_in * s_w[n]
  const float* bias;        // [Nfull] fake-quantised bias, nullable
  float out_scale0, out_scale1, out_inv0, out_inv1, out_qmin, out_qmax; int relu;
  // FX instantiation of tc_enc2_kernel (integer encoder of the fixed-point model, fxpmodel.py:331-366, fxprun.py:69-75): X holds
  // floats (quantised with FLOOR to fx_in_bits at 2^fx_in_exp) or int32 payloads, Y int32 payloads; the epilogue is
  // shift, saturate, bias, saturate, ReLU, and it leaves the per-column extremes of the output behind (fxp_bn_derive_kernel)
  int fx_x_is_float, fx_in_bits, fx_in_exp, fx_rs, fx_out_bits; const int* fx_bias; unsigned* fx_cmax; unsigned* fx_cmin;
  // FX instantiation of tc_dec2_kernel (integer decoder, fxpmodel.py:331-366, :1431-1439): the input exponent is data dependent
  // (read from *fx_ex_in on the device), Y receives the int32 payloads and fx_Yf their float image (either may be null)
  float* fx_Yf; float fx_out_scale; const int* fx_ex_in; int fx_dec_inp_bits, fx_dec_inp_exp, fx_dec_w_exp, fx_dec_out_exp; int* fx_status;
};

// clip(round_half_even(x / s)) with 1/s passed in: exact for the power-of-two scales of the static recipe
// (quantization.py:368-369, :395-396); the dense-tile path is only enabled when every scale is a power of two.
__device__ __forceinline__ int quant_code(float v, float inv, float qmin, float qmax) {
  return __float2int_rn(fminf(fmaxf(rintf(v * inv), qmin), qmax));
}
__device__ __forceinline__ float fq_from(float v, float s, float inv, float qmin, float qmax) {
  return fminf(fmaxf(rintf(v * inv), qmin), qmax) * s;
}
__device__ __forceinline__ uint32_t pack_h2(int a, int b) {
  const __half2 h = __halves2half2(__int2half_rn(a), __int2half_rn(b));
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Rounding without the conversion pipe.  For |t| <= 2^22, t + 1.5 * 2^23 is rounded to an integer by the add itself
// (round-half-even, the mode of rintf / jnp.round), and the low mantissa bits of the sum hold that integer in two's
// complement: bits = 0x4B400000 + code.  Clipping to the integer bounds BEFORE rounding gives the same result as after.
constexpr float kRoundMagic = 12582912.0f;
__device__ __forceinline__ uint32_t quant_bits(float v, float inv, float qmin, float qmax) {   // low 16 bits = int16 code
  return __float_as_uint(fminf(fmaxf(v * inv, qmin), qmax) + kRoundMagic);
}
__device__ __forceinline__ float round_clip(float t, float qmin, float qmax) {                  // clip(rint(t))
  return (fminf(fmaxf(t, qmin), qmax) + kRoundMagic) - kRoundMagic;
}
// Dense / B-projection tile kernel.  Per 128-row tile:
//   stage   : rows -> (BatchNorm) -> int16 codes -> signed high-byte plane + unsigned low-byte plane (8-bit operands)
//   MMA     : acc_hi = Ah * W^T, acc_lo = Al * W^T  (tcgen05 kind::i8, int32 in TMEM), NP output columns per pass
//   epilogue: acc = 256 acc_hi + acc_lo (exact) -> scale, bias, fake-quant, ReLU -> shared tile -> coalesced stores
// VEC4: rows are 16-byte aligned, lane owns input columns 4 lane .. 4 lane + 3 of each 128-column block.  Otherwise
// lane owns columns lane + 32 e (coalesced 4-byte loads) and the K axis is PERMUTED inside each block (the same
// permutation is baked into W8 by the host), so that the four codes of a lane still form one 32-bit store.
// XN: Nfull = Np + 1; the last output column is an integer dot product on the CUDA cores, accumulated while staging.
template <bool VEC4, int KV, bool BN, bool XN>
__global__ void __launch_bounds__(kTcThreads, 2) tc_linear_kernel(TcLinParams p) {
  TC_TRACE_ENTRY
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = p.Kp, Np = p.Np, NP = p.NP;
  const size_t a_bytes = tc_tile_bytes8(kTcRows, Kp);
  unsigned char* Wt = smem;                                   // [Np][Kp] weight codes
  unsigned char* Ah = Wt + tc_tile_bytes8(Np, Kp);            // [128][Kp] high bytes of the activation codes
  unsigned char* Al = Ah + a_bytes;                           // [128][Kp] low bytes
  const int ld_out = NP + 4;
  float* outT = reinterpret_cast<float*>(p.alias_out ? Ah : Al + a_bytes);      // output tile [128][NP + 4]
  float2* cst = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(outT) +
                                          (p.alias_out ? max(2 * a_bytes, (size_t)kTcRows * ld_out * 4) : (size_t)kTcRows * ld_out * 4));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int extra_n[XN ? kTcRows : 1];                   // integer dot of the extra output column
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kTcThreads / 32;
  if (warp == 0) tc_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_load_weights8(Wt, p.W8, Np, Kp);
  for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kTcThreads * 16) *reinterpret_cast<int4*>(Ah + i) = make_int4(0, 0, 0, 0);
  // per-column epilogue constants: t = acc * a + b with a = s_in s_w / s_out, b = bias / s_out (all scales powers of two)
  for (int n = tid; n < Np; n += kTcThreads) {
    const float inv = (n & 1) ? p.out_inv1 : p.out_inv0;
    cst[n] = n < p.Nfull ? make_float2(p.col_scale[n] * inv, p.bias ? p.bias[n] * inv : 0.f) : make_float2(0.f, 0.f);
  }
  const uint32_t lbo = tc_lbo(kTcRows);
  // per-lane staging constants.  Loads are never predicated (a predicated LDG costs three extra instructions):
  // rows beyond the end are clamped to the last row, columns beyond Kfull to the last column; what they produce
  // is multiplied by zero weights (pad columns of W8) or never written out (pad rows).
  float sub[KV][4], mul[KV][4], add[KV][4];
  int koff[KV][4];
  int wn[KV][4];
  bool wval[KV];
  uint32_t loff[KV];
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    loff[i] = (uint32_t)(8 * i + (lane >> 2)) * lbo + 4u * (lane & 3);
    wval[i] = 128 * i + 4 * lane < Kp;                        // the lane's 32-bit word lies inside the tile
#pragma unroll
    for (int e = 0; e < 4; ++e) wn[i][e] = (XN && wval[i]) ? (int)p.WN[128 * i + 4 * lane + e] : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = VEC4 ? 128 * i + 4 * lane + e : 128 * i + lane + 32 * e;
      const bool in = k < p.Kfull;
      koff[i][e] = in ? k : (VEC4 ? e : p.Kfull - 1);
      sub[i][e] = (BN && in) ? p.in_sub[k] : 0.f;
      mul[i][e] = (BN && in) ? p.in_mul[k] : 1.f;
      add[i][e] = (BN && in) ? p.in_add[k] : 0.f;
    }
  }
  const bool tail = p.Kfull > 128 * KV;                       // one more input column (encoder: 257 = 2 * 128 + 1)
  const float inv = p.in_inv, qmin = p.in_qmin, qmax = p.in_qmax;
  const float oqmin = p.relu ? fmaxf(p.out_qmin, 0.f) : p.out_qmin, oqmax = p.out_qmax, os0 = p.out_scale0, os1 = p.out_scale1;
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;     // epilogue: TMEM lane = row
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  const int egrp = warp >> 2;                                 // column half of a pass
  const bool wvec = (p.y_stride & 3) == 0 && (NP & 3) == 0 && (p.Nfull & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.Y) & 15) == 0;
  uint32_t phase = 0;
  TC_TRACE_DECL
  __syncthreads();
  // Balanced partition: every CTA owns one contiguous row range (full tiles plus one short tile whose row loops and
  // TMEM passes stop at its last row) instead of a whole number of 128-row tiles.
  const long long row_begin = p.rows * blockIdx.x / gridDim.x, row_end = p.rows * (blockIdx.x + 1) / gridDim.x;
  constexpr int RB = 8 / KV;
  const long long last_row = p.rows - 1;
  float raw[RB][KV][4];
  float rawt[RB];
  bool have_first = false;
  auto load_batch = [&](long long r0, int rb) {
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const float* xr = p.X + (size_t)min(r0 + rb + q, last_row) * p.x_stride;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        if constexpr (VEC4) {
          const float4 v = *reinterpret_cast<const float4*>(xr + koff[i][0]);
          raw[q][i][0] = v.x; raw[q][i][1] = v.y; raw[q][i][2] = v.z; raw[q][i][3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) raw[q][i][e] = xr[koff[i][e]];
        }
      }
      rawt[q] = tail ? xr[128 * KV] : 0.f;                    // same address in every lane: one broadcast load
    }
  };
  for (long long row0 = row_begin; row0 < row_end; row0 += kTcRows) {
    const int rv = (int)min((long long)kTcRows, row_end - row0);
    TC_TRACE_MARK(tid == 0);
    // ---- stage: RB rows are loaded before any of them is converted so that every lane keeps many loads in flight;
    //      the first batch of a tile was already requested before the previous tile's MMA (see below)
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
            float v = raw[q][i][e];
            if constexpr (BN) v = (v - sub[i][e]) * mul[i][e] + add[i][e];
            b[e] = quant_bits(v, inv, qmin, qmax);
            if constexpr (XN) part += (int)(short)b[e] * wn[i][e];
          }
          uint32_t lo, hi;
          split_planes(b[0], b[1], b[2], b[3], lo, hi);
          if (wval[i]) {
            *reinterpret_cast<uint32_t*>(Ah + loff[i] + roff) = hi;
            *reinterpret_cast<uint32_t*>(Al + loff[i] + roff) = lo;
          }
        }
        if (tail && lane == 0) {
          const uint32_t bt = quant_bits(rawt[q], inv, qmin, qmax);
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
    TC_TRACE_MARK(tid == 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    const uint32_t taddr = tmem_slot;
    if (tid == 0) {
      const int np = min(NP, Np);
      tc_issue8(taddr, Ah, Wt, Kp, np, Np, true);
      tc_issue8(taddr + 128, Al, Wt, Kp, np, Np, false);
      tc_commit(&bar);
    }
    // the first batch of the next tile travels while this tile is multiplied, post-processed and written out
    if (row0 + kTcRows < row_end) {
      load_batch(row0 + kTcRows, warp * RB);
      have_first = true;
    }
    for (int pass0 = 0; pass0 < Np; pass0 += NP) {
      const int np = min(NP, Np - pass0);                     // MMA width of this pass (multiple of 16)
      const int npv = min(np, p.Nfull - pass0);               // columns that exist in Y
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      TC_TRACE_MARK(tid == 0);
      // ---- epilogue: exact integer recombination, scale, bias, fake-quant, ReLU -> shared tile
      const int cw = np >> 1;
      for (int c0 = egrp * cw; c0 < (erow0 < rv ? (egrp + 1) * cw : 0); c0 += 8) {
        uint32_t vh[8], vl[8];
        tc_ld8(taddr + lane_base + c0, vh);
        tc_ld8(taddr + lane_base + 128 + c0, vl);
        tc_wait_ld();
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 c2 = *reinterpret_cast<const float4*>(cst + pass0 + c0 + j);   // (a, b) of two columns
          const float f0 = __int2float_rn((int)vh[j] * 256 + (int)vl[j]);
          const float f1 = __int2float_rn((int)vh[j + 1] * 256 + (int)vl[j + 1]);
          // ReLU after the fake-quant = lower clip bound raised to zero
          out[j] = round_clip(fmaf(f0, c2.x, c2.y), oqmin, oqmax) * os0;
          out[j + 1] = round_clip(fmaf(f1, c2.z, c2.w), oqmin, oqmax) * os1;
        }
        float* dst = outT + (size_t)erow * ld_out + c0;
        *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(out[4], out[5], out[6], out[7]);
      }
      TC_TRACE_MARK(tid == 0);
      tc_fence_before();
      __syncthreads();                                        // tile complete; TMEM and (last pass) the planes are free
      TC_TRACE_MARK(tid == 0);
      if (tid == 0 && pass0 + NP < Np) {                      // next pass overlaps the write-out below
        tc_fence_after();
        const int np2 = min(NP, Np - pass0 - NP);
        tc_issue8(taddr, Ah, Wt + (size_t)(pass0 + NP) * 16, Kp, np2, Np, true);
        tc_issue8(taddr + 128, Al, Wt + (size_t)(pass0 + NP) * 16, Kp, np2, Np, false);
        tc_commit(&bar);
      }
      // ---- write-out, coalesced.  (One bulk copy per row through the TMA engine was measured 3-4x slower: ~50 ns per
      // 512-byte copy, 128 copies per tile.)
      if (wvec) {
        const int q4 = npv >> 2;                              // float4 per row
        const int rpi = q4 >= 32 ? 1 : 32 / q4;               // rows per warp iteration
        const int lr = lane / q4, lc = lane - lr * q4;
        if (lr < rpi) {
          for (int rb = warp * rpi + lr; rb < rv; rb += 4 * NW * rpi) {   // four rows in flight per lane
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const float4*>(outT + (size_t)min(rb + q * NW * rpi, kTcRows - 1) * ld_out + 4 * lc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (rb + q * NW * rpi < rv)
                *reinterpret_cast<float4*>(p.Y + (size_t)(row0 + rb + q * NW * rpi) * p.y_stride + pass0 + 4 * lc) = v[q];
          }
        }
      } else {
        // unaligned rows (decoder, 257 floats): 4-byte stores, lanes along columns, four rows in flight per warp
        for (int rb = warp * 4; rb < rv; rb += NW * 4) {
          float v[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[q][j] = outT[(size_t)(rb + q) * ld_out + lane + 32 * j];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* yr = p.Y + (size_t)(row0 + rb + q) * p.y_stride + pass0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + q < rv && lane + 32 * j < npv) yr[lane + 32 * j] = v[q][j];
          }
        }
      }
      if constexpr (XN) {
        if (pass0 + NP >= Np && tid < rv) {                   // the extra column: one value per row
          const int n = p.Nfull - 1;
          const float oinv = (n & 1) ? p.out_inv1 : p.out_inv0;
          const float t = fmaf(__int2float_rn(extra_n[tid]), p.col_scale[n] * oinv, p.bias ? p.bias[n] * oinv : 0.f);
          p.Y[(size_t)(row0 + tid) * p.y_stride + n] = round_clip(t, oqmin, oqmax) * ((n & 1) ? os1 : os0);
        }
      }
      __syncthreads();                                        // output tile free again
      TC_TRACE_MARK(tid == 0);
    }
    TC_TRACE_NEXT();
  }
  TC_TRACE_PRINT(tid == 0, BN ? "linear bn: stage sync mma epi sync writeout" : "linear: stage sync mma epi sync writeout [per pass: mma epi sync writeout]");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 256);
}

// ------------------------------------------------------------------------------------------- encoder (odd K)
// Dense layer whose input rows are contiguous but NOT 16-byte aligned (encoder: 257 floats = 1028 bytes per frame).
// A block of 4n rows is contiguous and 16-byte aligned, so the raw rows are streamed by the TMA engine into a ring of
// three 32-row buffers (one bulk copy each, mbarrier completion) while the SM converts / multiplies / post-processes
// earlier rows: loads never stop, and no load instruction is issued by the warps at all.  One CTA of 16 warps per SM.
// K order: lane-interleaved inside each 128-column block (conflict-free 4-byte shared loads with the odd row pitch).
constexpr int kEncThreads = 512;
constexpr int kEncSub = 32;        // rows per ring buffer
constexpr int kEncRing = 3;

template <int KV>
__global__ void __launch_bounds__(kEncThreads, 1) tc_enc_kernel(TcLinParams p) {
  TC_TRACE_ENTRY
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = p.Kp, Np = p.Np, Kf = p.Kfull;
  const size_t a_bytes = tc_tile_bytes8(kTcRows, Kp);
  const int ld_out = Np + 4;
  unsigned char* Wt = smem;                                   // [Np][Kp] weight codes
  unsigned char* Ah = Wt + tc_tile_bytes8(Np, Kp);
  unsigned char* Al = Ah + a_bytes;
  float* outT = reinterpret_cast<float*>(Ah);                 // output tile [128][Np + 4], aliases the dead planes
  const size_t raw_bytes = (((size_t)kEncSub * Kf * 4) + 127) & ~(size_t)127;
  unsigned char* rawb = Al + a_bytes;                         // ring of raw row buffers [32][Kf] floats
  float2* cst = reinterpret_cast<float2*>(rawb + kEncRing * raw_bytes);
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t full[kEncRing];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kEncThreads / 32;
  if (warp == 0) tc_alloc(&tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < kEncRing; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  tc_load_weights8(Wt, p.W8, Np, Kp);
  for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kEncThreads * 16) *reinterpret_cast<int4*>(Ah + i) = make_int4(0, 0, 0, 0);
  for (int n = tid; n < Np; n += kEncThreads) {
    const float oinv = (n & 1) ? p.out_inv1 : p.out_inv0;
    cst[n] = n < p.Nfull ? make_float2(p.col_scale[n] * oinv, p.bias ? p.bias[n] * oinv : 0.f) : make_float2(0.f, 0.f);
  }
  const uint32_t lbo = tc_lbo(kTcRows);
  uint32_t loff[KV];
  int kq[KV][4];
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    loff[i] = (uint32_t)(8 * i + (lane >> 2)) * lbo + 4u * (lane & 3);
#pragma unroll
    for (int e = 0; e < 4; ++e) kq[i][e] = min(128 * i + lane + 32 * e, Kf - 1);   // pad columns meet zero weights
  }
  const bool tail = Kf > 128 * KV;
  const float inv = p.in_inv, qmin = p.in_qmin, qmax = p.in_qmax;
  const float oqmin = p.relu ? fmaxf(p.out_qmin, 0.f) : p.out_qmin, oqmax = p.out_qmax, os0 = p.out_scale0, os1 = p.out_scale1;
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  const int cw = Np >> 2, cbeg = (warp >> 2) * cw;            // 4 column groups (Np multiple of 32)
  __syncthreads();
  const uint32_t taddr = tmem_slot;
  // contiguous row range per CTA, boundaries on multiples of 4 rows (16-byte alignment of the bulk copies)
  const long long rb4 = (p.rows * blockIdx.x / gridDim.x) & ~3LL;
  const long long re4 = blockIdx.x + 1 == gridDim.x ? p.rows : (p.rows * (blockIdx.x + 1) / gridDim.x) & ~3LL;
  const int nsub = (int)((re4 - rb4 + kEncSub - 1) / kEncSub);
  auto issue = [&](int sidx) {                                // thread 0: request sub-tile sidx into its ring slot
    const long long r0 = rb4 + (long long)sidx * kEncSub;
    const int nr = (int)min((long long)kEncSub, re4 - r0) & ~3;   // whole groups of 4 rows; the rest is copied by hand
    const uint32_t bytes = (uint32_t)nr * Kf * 4u;
    uint64_t* fb = &full[sidx % kEncRing];
    if (bytes) {
      mbar_expect_tx(fb, bytes);
      bulk_g2s(rawb + (size_t)(sidx % kEncRing) * raw_bytes, p.X + (size_t)r0 * Kf, bytes, fb);
    } else {
      mbar_arrive(fb);
    }
  };
  if (tid == 0)
    for (int i = 0; i < kEncRing && i < nsub; ++i) issue(i);
  uint32_t phase = 0;
  TC_TRACE_DECL
  for (int sidx = 0; sidx < nsub; ++sidx) {
    TC_TRACE_MARK(tid == 0);
    const long long r0 = rb4 + (long long)sidx * kEncSub;
    const int nr = (int)min((long long)kEncSub, re4 - r0);
    const int slot = sidx % kEncRing;
    float* raw = reinterpret_cast<float*>(rawb + (size_t)slot * raw_bytes);
    const int trow0 = (sidx & 3) * kEncSub;                   // position of the sub-tile inside the 128-row tile
    mbar_wait(&full[slot], (uint32_t)(sidx / kEncRing) & 1u);
    TC_TRACE_MARK(tid == 0);
    if ((nr & 3) && tid < (nr & 3) * 32) {                    // <= 3 rows at the very end of the input: plain loads
      const int rr = (nr & ~3) + (tid >> 5);
      for (int k = tid & 31; k < Kf; k += 32) raw[(size_t)rr * Kf + k] = p.X[(size_t)(r0 + rr) * Kf + k];
    }
    if (nr & 3) __syncthreads();
    // ---- convert: warp w owns rows 2w, 2w + 1 of the sub-tile
#pragma unroll
    for (int q = 0; q < kEncSub / NW; ++q) {
      const int r = warp * (kEncSub / NW) + q;
      const float* xr = raw + (size_t)min(r, nr - 1) * Kf;
      const uint32_t roff = (uint32_t)(trow0 + r) * 16u;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        uint32_t b[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = quant_bits(xr[kq[i][e]], inv, qmin, qmax);
        uint32_t lo, hi;
        split_planes(b[0], b[1], b[2], b[3], lo, hi);
        *reinterpret_cast<uint32_t*>(Ah + loff[i] + roff) = hi;
        *reinterpret_cast<uint32_t*>(Al + loff[i] + roff) = lo;
      }
      if (tail && lane == 0) {
        const uint32_t bt = quant_bits(xr[128 * KV], inv, qmin, qmax);
        *reinterpret_cast<uint32_t*>(Ah + (size_t)(8 * KV) * lbo + roff) = (bt >> 8) & 0xFFu;
        *reinterpret_cast<uint32_t*>(Al + (size_t)(8 * KV) * lbo + roff) = bt & 0xFFu;
      }
    }
    const bool tile_done = (sidx & 3) == 3 || sidx + 1 == nsub;
    if (tile_done) {
      fence_proxy_async_smem();
      tc_fence_before();
    }
    __syncthreads();                                          // the ring slot is free; (tile_done) the planes are complete
    if (tid == 0 && sidx + kEncRing < nsub) issue(sidx + kEncRing);
    if (!tile_done) continue;
    TC_TRACE_MARK(tid == 0);
    tc_fence_after();
    if (tid == 0) {
      tc_issue8(taddr, Ah, Wt, Kp, Np, Np, true);
      tc_issue8(taddr + 128, Al, Wt, Kp, Np, Np, false);
      tc_commit(&bar);
    }
    const long long row0 = rb4 + (long long)(sidx & ~3) * kEncSub;
    const int rv = (int)min((long long)kTcRows, re4 - row0);
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    // ---- epilogue -> padded shared tile (over the dead planes)
    if (erow0 < rv) {
      for (int c0 = cbeg; c0 < cbeg + cw; c0 += 8) {
        uint32_t vh[8], vl[8];
        tc_ld8(taddr + lane_base + c0, vh);
        tc_ld8(taddr + lane_base + 128 + c0, vl);
        tc_wait_ld();
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 c2 = *reinterpret_cast<const float4*>(cst + c0 + j);
          const float f0 = __int2float_rn((int)vh[j] * 256 + (int)vl[j]);
          const float f1 = __int2float_rn((int)vh[j + 1] * 256 + (int)vl[j + 1]);
          out[j] = round_clip(fmaf(f0, c2.x, c2.y), oqmin, oqmax) * os0;
          out[j + 1] = round_clip(fmaf(f1, c2.z, c2.w), oqmin, oqmax) * os1;
        }
        float* dst = outT + (size_t)erow * ld_out + c0;
        *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(out[4], out[5], out[6], out[7]);
      }
    }
    tc_fence_before();
    __syncthreads();
    TC_TRACE_MARK(tid == 0);
    // ---- write-out, coalesced 16-byte stores (output rows are Nfull floats, 16-byte aligned)
    {
      const int q4 = p.Nfull >> 2;
      const int rpi = q4 >= 32 ? 1 : 32 / q4;
      const int lr = lane / q4, lc = lane - lr * q4;
      if (lr < rpi) {
        for (int rb = warp * rpi + lr; rb < rv; rb += 2 * NW * rpi) {
          float4 v[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) v[q] = *reinterpret_cast<const float4*>(outT + (size_t)min(rb + q * NW * rpi, kTcRows - 1) * ld_out + 4 * lc);
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (rb + q * NW * rpi < rv)
              *reinterpret_cast<float4*>(p.Y + (size_t)(row0 + rb + q * NW * rpi) * p.y_stride + 4 * lc) = v[q];
        }
      }
    }
    __syncthreads();                                          // the planes may be overwritten again
    TC_TRACE_MARK(tid == 0);
    TC_TRACE_NEXT();
  }
  TC_TRACE_PRINT(tid == 0, "enc: [wait convert+sync] x4, mma, epi+sync, writeout+sync");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 256);
}

// ------------------------------------------------------------------------------------------- encoder, warp-specialised
// Same layer as tc_enc_kernel (contiguous input rows that are not 16-byte aligned: 257 floats per frame), but no phase of
// the tile ever waits for another one.  The product is formed TRANSPOSED - the weights are the A operand (M = 128 output
// columns, zero padded), a 96-row activation tile the B operand - so the accumulator has TMEM lane = output column and
// TMEM column = row: an epilogue thread owns ONE output column, keeps its scale / bias in registers and stores
// Y[row][column] coalesced across the warp straight from registers (no shared output tile, no write-out pass).  Roles:
//   warp 8        TMA producer: raw rows -> ring of three 32-row slots (1-D bulk copies, mbarrier complete_tx)
//   warps 10..25  converters: ring slot -> FQ16 codes -> hi / lo int8 planes of tile buffer t & 1 (double buffered)
//   warp 9        MMA issuer: tcgen05.mma kind::i8 of both planes into accumulator buffer t & 1 (2 x 64 TMEM columns each)
//   warps 0..7    epilogue (lane quadrant w % 4, row half w / 4): tcgen05.ld -> 256 hi + lo -> scale, bias, fake-quant, ReLU -> global
// linked by mbarriers (raw_full / raw_free per slot, planes_full / mma_done / tmem_free per buffer).  HBM reads never stop
// while a tile is multiplied or written, which the first version (all 16 warps convert -> multiply -> post-process ->
// write in turn) paid with 49 % of the HBM roofline.
constexpr int kEnc2Rows = 64;      // rows per tile (MMA N)
constexpr int kEnc2Sub = 32;       // rows per ring slot: 32 rows of 257 floats are a multiple of 128 bytes
constexpr int kEnc2Ring = 3;
constexpr int kEnc2Conv = 16;      // converter warps
constexpr int kEnc2Epi = 8;        // epilogue warps: TMEM lane quadrant w % 4, row half w / 4
constexpr int kEnc2Threads = 32 * (kEnc2Epi + 2 + kEnc2Conv);
constexpr int kEnc2SubPerTile = kEnc2Rows / kEnc2Sub;

template <int KV, int NRING, bool FX = false>
__global__ void __launch_bounds__(kEnc2Threads, 1) tc_enc2_kernel(TcLinParams p) {
  constexpr int nring = NRING;                                // ring slots: compile-time, the slot / phase arithmetic sits in every loop
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = p.Kp, Np = p.Np, Kf = p.Kfull;
  const int nblk = (Np + 127) >> 7;                           // blocks of 128 output columns (<= 2)
  const size_t w_bytes = tc_tile_bytes8(kTcRows, Kp);         // one A operand block: 128 rows (output columns), zero padded
  // one tile buffer: hi plane (rows 0 .. 63) and lo plane (rows 64 .. 127) STACKED in one 128-row B operand, so that one
  // MMA per K step multiplies both.  A single operand has one signedness: the low byte travels as lo - 128 (lo ^ 0x80 read
  // as int8) and the epilogue adds 128 * sum_k W[n][k] back - an exact integer per output column.
  const size_t a_bytes = tc_tile_bytes8(2 * kEnc2Rows, Kp);
  const size_t raw_bytes = (((size_t)kEnc2Sub * Kf * 4) + 127) & ~(size_t)127;
  unsigned char* Wt = smem;
  unsigned char* planes = Wt + (size_t)nblk * w_bytes;        // [buf][hi | lo]
  unsigned char* rawb = planes + 2 * a_bytes;
  __shared__ __align__(8) uint64_t raw_full[kEnc2Ring], raw_free[kEnc2Ring], planes_full[2], mma_done[2], tmem_free[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef S5_ENC2_TRACE
  long long tw[3] = {0, 0, 0};
  const long long tstart = clock64();
#define E2_WAIT(i, stmt) do { const long long t_ = clock64(); stmt; tw[i] += clock64() - t_; } while (0)
#else
#define E2_WAIT(i, stmt) stmt
#endif
  const int n_epi = kEnc2Epi;                                 // every epilogue warp arrives (warps without real columns store nothing)
  if (warp == 0) tc_alloc(&tmem_slot, nblk == 1 ? 256 : 512);
  if (tid == 0) {
    for (int i = 0; i < nring; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], kEnc2Conv); }
    for (int i = 0; i < 2; ++i) { mbar_init(&planes_full[i], kEnc2Conv); mbar_init(&mma_done[i], 1); mbar_init(&tmem_free[i], n_epi); }
    fence_mbar_init();
  }
  {   // weights: rows < Np from global (K already in the staging order), the rest zero; planes: zero (pad bytes stay zero)
    const int kchunks = Kp >> 4;
    const uint32_t lbw = tc_lbo(kTcRows);
    for (int i = tid; i < nblk * kTcRows * kchunks; i += kEnc2Threads) {
      const int r = i / kchunks, kc = i - r * kchunks;       // r = output column
      const int4 v = r < Np ? *reinterpret_cast<const int4*>(p.W8 + (size_t)r * Kp + kc * 16) : make_int4(0, 0, 0, 0);
      *reinterpret_cast<int4*>(Wt + (size_t)(r >> 7) * w_bytes + (size_t)kc * lbw + (size_t)(r & 127) * 16) = v;
    }
    for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kEnc2Threads * 16) *reinterpret_cast<int4*>(planes + i) = make_int4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_slot;
  // contiguous row range per CTA, boundaries on multiples of 4 rows (16-byte alignment of the bulk copies)
  // (32-row boundaries: every bulk copy then starts on a 128-byte boundary when X does)
  const long long rb4 = (p.rows * blockIdx.x / gridDim.x) & ~31LL;
  const long long re4 = blockIdx.x + 1 == gridDim.x ? p.rows : (p.rows * (blockIdx.x + 1) / gridDim.x) & ~31LL;
  const int nsub = (int)((re4 - rb4 + kEnc2Sub - 1) / kEnc2Sub);
  const int ntiles = (nsub + kEnc2SubPerTile - 1) / kEnc2SubPerTile;

  if (warp == kEnc2Epi) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int s = 0; s < nsub; ++s) {
        const int slot = s % nring;
        if (s >= nring) E2_WAIT(0, mbar_wait(&raw_free[slot], (uint32_t)((s / nring) - 1) & 1u));
        const long long r0 = rb4 + (long long)s * kEnc2Sub;
        const int nr = (int)min((long long)kEnc2Sub, re4 - r0) & ~3;   // whole groups of 4 rows; the rest is copied by hand
        const uint32_t bytes = (uint32_t)nr * Kf * 4u;
        if (bytes) {
          mbar_expect_tx(&raw_full[slot], bytes);
          bulk_g2s(rawb + (size_t)slot * raw_bytes, p.X + (size_t)r0 * Kf, bytes, &raw_full[slot]);
        } else {
          mbar_arrive(&raw_full[slot]);
        }
      }
    }
  } else if (warp == kEnc2Epi + 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t la = tc_lbo(kTcRows);                     // A (weights) and B (stacked planes) are both 128-row tiles
      const uint64_t a_desc = tc_smem_desc(smem_u32(Wt), la, 128);
      const uint32_t step = (2 * la) >> 4;                    // next 32 columns of K, in units of 16 bytes
      const uint32_t idesc = tc_idesc_i8(kTcRows, 2 * kEnc2Rows, true, true);
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t k = (uint32_t)(t >> 1);
        E2_WAIT(0, mbar_wait(&planes_full[buf], k & 1u));
        if (t >= 2) E2_WAIT(1, mbar_wait(&tmem_free[buf], (k - 1u) & 1u));
        tc_fence_after();
        const uint64_t b_desc = tc_smem_desc(smem_u32(planes + (size_t)buf * a_bytes), la, 128);
        for (int blk = 0; blk < nblk; ++blk) {
          const uint32_t d0 = taddr + (uint32_t)(buf * nblk + blk) * (2 * kEnc2Rows);
          const uint64_t ab = tc_desc_advance(a_desc, (uint32_t)(blk * (w_bytes >> 4)));
          for (int j = 0; j < (Kp >> 5); ++j) tc_mma_i8(d0, tc_desc_advance(ab, j * step), tc_desc_advance(b_desc, j * step), idesc, j > 0);
        }
        tc_commit(&mma_done[buf]);
      }
    }
  } else if (warp >= kEnc2Epi + 2) {
    // ---------------------------------------------------------------- converters
    const int cw = warp - (kEnc2Epi + 2);
    const uint32_t lbo = tc_lbo(2 * kEnc2Rows);
    uint32_t loff[KV];
    int kq[KV][4];
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      loff[i] = (uint32_t)(8 * i + (lane >> 2)) * lbo + 4u * (lane & 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) kq[i][e] = min(128 * i + lane + 32 * e, Kf - 1);   // pad columns meet zero weights
    }
    const bool tail = Kf > 128 * KV;
    const float inv = p.in_inv, qmin = p.in_qmin, qmax = p.in_qmax;
    // FX: fxp_from_fp FLOOR (fxprun.py:69-75) = floor(x 2^e), saturated - or the int32 payload as it is
    const float fx_sc = FX ? exp2f((float)p.fx_in_exp) : 0.f;
    auto code = [&](float x) -> uint32_t {
      if constexpr (FX) return (uint32_t)(p.fx_x_is_float ? fx_sat(__float2int_rd(x * fx_sc), p.fx_in_bits) : __float_as_int(x));
      else return quant_bits(x, inv, qmin, qmax);
    };
    for (int s = 0; s < nsub; ++s) {
      const int t = s / kEnc2SubPerTile, ts = s - t * kEnc2SubPerTile, buf = t & 1, slot = s % nring;
      if (ts == 0 && t >= 2) E2_WAIT(1, mbar_wait(&mma_done[buf], (uint32_t)((t >> 1) - 1) & 1u));   // the MMAs of tile t - 2 have read these planes
      unsigned char* Ah = planes + (size_t)buf * a_bytes;
      unsigned char* Al = Ah + kEnc2Rows * 16;                // rows 64 .. 127 of the same operand tile
      const long long r0 = rb4 + (long long)s * kEnc2Sub;
      const int nr = (int)min((long long)kEnc2Sub, re4 - r0);
      float* raw = reinterpret_cast<float*>(rawb + (size_t)slot * raw_bytes);
      E2_WAIT(0, mbar_wait(&raw_full[slot], (uint32_t)(s / nring) & 1u));
      if (nr & 3) {                                           // <= 3 rows at the very end of the input: plain loads
        if (cw < (nr & 3)) {
          const int rr = (nr & ~3) + cw;
          for (int k = lane; k < Kf; k += 32) raw[(size_t)rr * Kf + k] = p.X[(size_t)(r0 + rr) * Kf + k];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kEnc2Conv) : "memory");
      }
      // all shared loads of this warp's rows first, then the arithmetic and the plane stores: a load behind a store it
      // may alias with (raw ring and planes are both shared memory) is never hoisted by the compiler
      constexpr int RPW = kEnc2Sub / kEnc2Conv;
      float xv[RPW][KV][4], xt[RPW];
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        const float* xr = raw + (size_t)min(cw * RPW + q, nr - 1) * Kf;
#pragma unroll
        for (int i = 0; i < KV; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) xv[q][i][e] = xr[kq[i][e]];
        xt[q] = xr[min(128 * KV, Kf - 1)];
      }
#pragma unroll
      for (int q = 0; q < RPW; ++q) {
        const uint32_t roff = (uint32_t)(ts * kEnc2Sub + cw * RPW + q) * 16u;
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          uint32_t b[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) b[e] = code(xv[q][i][e]);
          uint32_t lo, hi;
          split_planes(b[0], b[1], b[2], b[3], lo, hi);
          *reinterpret_cast<uint32_t*>(Ah + loff[i] + roff) = hi;
          *reinterpret_cast<uint32_t*>(Al + loff[i] + roff) = lo ^ 0x80808080u;
        }
        const uint32_t bt = code(xt[q]);
        if (tail && lane == 0) {
          *reinterpret_cast<uint32_t*>(Ah + (size_t)(8 * KV) * lbo + roff) = (bt >> 8) & 0xFFu;
          *reinterpret_cast<uint32_t*>(Al + (size_t)(8 * KV) * lbo + roff) = (bt & 0xFFu) ^ 0x80808080u;
        }
      }
      const bool tile_done = ts == kEnc2SubPerTile - 1 || s + 1 == nsub;
      if (tile_done) fence_proxy_async_smem();                // the planes are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_free[slot]);
        if (tile_done) mbar_arrive(&planes_full[buf]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: lane = output column, warp / 4 = row half
    const int quad = warp & 3, half = warp >> 2;
    const float oqmin = p.relu ? fmaxf(p.out_qmin, 0.f) : p.out_qmin, oqmax = p.out_qmax;
    const uint32_t lane_base = (uint32_t)(32 * quad) << 16;
    constexpr int kHalfRows = kEnc2Rows / 2;
    const size_t ystride = (size_t)p.y_stride;
    float cs[2], cb[2], os[2];
    int corr[2];                                              // 128 * sum_k W[n][k]: undoes the lo - 128 of the stacked low plane
    bool ncol[2];
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      const int n = 128 * blk + 32 * quad + lane;
      ncol[blk] = blk < nblk && n < p.Nfull;
      const float oinv = (n & 1) ? p.out_inv1 : p.out_inv0;
      os[blk] = (n & 1) ? p.out_scale1 : p.out_scale0;
      cs[blk] = (!FX && ncol[blk]) ? p.col_scale[n] * oinv : 0.f;
      cb[blk] = (!FX && ncol[blk] && p.bias) ? p.bias[n] * oinv : 0.f;
      int c = 0;
      if (ncol[blk] && n < Np) {
        const int4* wr = reinterpret_cast<const int4*>(p.W8 + (size_t)n * Kp);
        for (int i = 0; i < (Kp >> 4); ++i) {
          const int4 w = wr[i];
          c = __dp4a(w.x, 0x01010101, c); c = __dp4a(w.y, 0x01010101, c); c = __dp4a(w.z, 0x01010101, c); c = __dp4a(w.w, 0x01010101, c);
        }
      }
      corr[blk] = 128 * c;
    }
    int fbias[2] = {0, 0}, cmx[2] = {INT_MIN, INT_MIN}, cmn[2] = {INT_MAX, INT_MAX};
    const int fhi = (FX && p.fx_out_bits < 32) ? (1 << (p.fx_out_bits - 1)) - 1 : INT_MAX, flo = -fhi - 1;
    if constexpr (FX) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) fbias[blk] = (ncol[blk] && p.fx_bias) ? p.fx_bias[128 * blk + 32 * quad + lane] : 0;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const long long row0 = rb4 + (long long)t * kEnc2Rows;
      const int rv = (int)min((long long)kEnc2Rows, re4 - row0);
      E2_WAIT(0, mbar_wait(&mma_done[buf], (uint32_t)(t >> 1) & 1u));
      tc_fence_after();
      for (int c0 = half * kHalfRows; c0 < min(rv, (half + 1) * kHalfRows); c0 += 16) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          if (blk >= nblk) break;
          const uint32_t d0 = taddr + (uint32_t)(buf * nblk + blk) * (2 * kEnc2Rows) + lane_base;
          uint32_t vh[16], vl[16];
          tc_ld16(d0 + c0, vh);
          tc_ld16(d0 + kEnc2Rows + c0, vl);
          tc_wait_ld();
          // all 16 values first (independent chains), then the stores: a predicate per element would serialise the chains
          float o[16];
          if constexpr (FX) {
            // fxp_matmul + FxpDense bias + fxp_relu (fxparray.py:640-678, fxpmodel.py:352-364, :1262-1266) on the exact int32 sum
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              int v = min(max((int)(vh[j] * 256u + vl[j] + (uint32_t)corr[blk]) >> p.fx_rs, flo), fhi);
              if (p.fx_bias) v = min(max(fx_addw(v, fbias[blk]), flo), fhi);
              if (p.relu) v = max(v, 0);
              if (c0 + j < rv) { cmx[blk] = max(cmx[blk], v); cmn[blk] = min(cmn[blk], v); }
              o[j] = __int_as_float(v);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              o[j] = round_clip(fmaf(__int2float_rn((int)vh[j] * 256 + (int)vl[j] + corr[blk]), cs[blk], cb[blk]), oqmin, oqmax) * os[blk];
          }
          float* q = p.Y + (size_t)(row0 + c0) * ystride + 128 * blk + 32 * quad + lane;
          if (ncol[blk]) {
            if (c0 + 16 <= rv) {
#pragma unroll
              for (int j = 0; j < 16; ++j, q += ystride) *q = o[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j, q += ystride)
                if (c0 + j < rv) *q = o[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[buf]);
    }
    if constexpr (FX) {                                       // per-column extremes: a lane owns one column per block
      if (p.fx_cmax) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
          if (ncol[blk] && cmx[blk] >= cmn[blk]) {
            atomicMax(p.fx_cmax + 128 * blk + 32 * quad + lane, ord_enc(cmx[blk]));
            atomicMax(p.fx_cmin + 128 * blk + 32 * quad + lane, ord_enc(~cmn[blk]));
          }
      }
    }
  }
#ifdef S5_ENC2_TRACE
  if (blockIdx.x == 3 && lane == 0 && (warp == 0 || warp == 4 || warp == 8 || warp == 9 || warp == 10 || warp == 25))
    printf("[enc2 trace] warp %d: total %lld cyc, wait0 %lld, wait1 %lld (tiles %d, subs %d)\n", warp, clock64() - tstart, tw[0], tw[1], ntiles, nsub);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, nblk == 1 ? 256 : 512);
}

// ------------------------------------------------------------------------------------------- decoder (wide, odd N)
// Dense layer with K <= 128 aligned input columns and Nfull = Np + 1 <= 257 output columns whose rows are contiguous
// in memory (decoder: 96 -> 257).  The 1028-byte output rows are not 16-byte aligned, but a 128-row tile of them is one
// contiguous, 16-byte aligned block: the tile is assembled in shared memory with the global row pitch and leaves as ONE
// asynchronous bulk copy (TMA engine) that overlaps the next tile.  One CTA of 16 warps per SM, N = Np in a single
// MMA pass (2 x 256 TMEM columns), the extra column on the CUDA cores.
constexpr int kDecThreads = 512;

__global__ void __launch_bounds__(kDecThreads, 1) tc_dec_kernel(TcLinParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = p.Kp, Np = p.Np, Nf = p.Nfull;
  const size_t a_bytes = tc_tile_bytes8(kTcRows, Kp);
  unsigned char* Wt = smem;                                   // [Np][Kp] weight codes
  unsigned char* Ah = Wt + tc_tile_bytes8(Np, Kp);
  unsigned char* Al = Ah + a_bytes;
  float* outT = reinterpret_cast<float*>(Al + a_bytes);       // [128][Nf] with the pitch of the global rows
  float2* cst = reinterpret_cast<float2*>(outT + (size_t)kTcRows * Nf);   // per column: (s_in s_w / s_out, bias / s_out)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int extra_n[kTcRows];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kDecThreads / 32;
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_load_weights8(Wt, p.W8, Np, Kp);
  for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kDecThreads * 16) *reinterpret_cast<int4*>(Ah + i) = make_int4(0, 0, 0, 0);
  for (int n = tid; n < Nf; n += kDecThreads) {
    const float oinv = (n & 1) ? p.out_inv1 : p.out_inv0;
    cst[n] = make_float2(p.col_scale[n] * oinv, p.bias ? p.bias[n] * oinv : 0.f);
  }
  const uint32_t lbo = tc_lbo(kTcRows);
  const uint32_t loff = (uint32_t)(lane >> 2) * lbo + 4u * (lane & 3);
  const bool wval = 4 * lane < Kp;
  const int kcol = 4 * lane < p.Kfull ? 4 * lane : 0;
  int wn[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) wn[e] = wval ? (int)p.WN[4 * lane + e] : 0;
  const float inv = p.in_inv, qmin = p.in_qmin, qmax = p.in_qmax;
  const float oqmin = p.relu ? fmaxf(p.out_qmin, 0.f) : p.out_qmin, oqmax = p.out_qmax, os0 = p.out_scale0, os1 = p.out_scale1;
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  const int cw = Np >> 2, cbeg = (warp >> 2) * cw;            // 4 column groups
  uint32_t phase = 0;
  __syncthreads();
  const uint32_t taddr = tmem_slot;
  // contiguous row range per CTA, boundaries on multiples of 4 rows (16-byte alignment of the tile in global memory)
  const long long rb4 = (p.rows * blockIdx.x / gridDim.x) & ~3LL;
  const long long re4 = blockIdx.x + 1 == gridDim.x ? p.rows : (p.rows * (blockIdx.x + 1) / gridDim.x) & ~3LL;
  const long long last_row = p.rows - 1;
  constexpr int RB = 8;                                       // 16 warps x 8 rows: the whole tile is one batch
  float4 raw[RB];
  auto load_batch = [&](long long r0) {
#pragma unroll
    for (int q = 0; q < RB; ++q)
      raw[q] = *reinterpret_cast<const float4*>(p.X + (size_t)min(r0 + warp * RB + q, last_row) * p.x_stride + kcol);
  };
  if (rb4 < re4) load_batch(rb4);
  for (long long row0 = rb4; row0 < re4; row0 += kTcRows) {
    const int rv = (int)min((long long)kTcRows, re4 - row0);
    // ---- stage this warp's 8 rows (requested during the previous tile)
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int r = warp * RB + q;
      const float xv[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
      uint32_t b[4];
      int part = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        b[e] = quant_bits(xv[e], inv, qmin, qmax);
        part += (int)(short)b[e] * wn[e];
      }
      uint32_t lo, hi;
      split_planes(b[0], b[1], b[2], b[3], lo, hi);
      if (wval) {
        *reinterpret_cast<uint32_t*>(Ah + loff + (uint32_t)r * 16u) = hi;
        *reinterpret_cast<uint32_t*>(Al + loff + (uint32_t)r * 16u) = lo;
      }
      part = __reduce_add_sync(0xffffffffu, part);
      if (lane == 0) extra_n[r] = part;
    }
    if (tid == 0) bulk_wait_read0();                          // the previous tile has left the output tile
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      tc_issue8(taddr, Ah, Wt, Kp, Np, Np, true);
      tc_issue8(taddr + 256, Al, Wt, Kp, Np, Np, false);
      tc_commit(&bar);
    }
    if (row0 + kTcRows < re4) load_batch(row0 + kTcRows);     // next tile's rows travel during MMA + epilogue
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread = row, 4-byte stores with the odd row pitch (conflict-free)
    if (erow0 < rv) {
      float* orow = outT + (size_t)erow * Nf;
      for (int c0 = cbeg; c0 < cbeg + cw; c0 += 8) {
        uint32_t vh[8], vl[8];
        tc_ld8(taddr + lane_base + c0, vh);
        tc_ld8(taddr + lane_base + 256 + c0, vl);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 c2 = *reinterpret_cast<const float4*>(cst + c0 + j);
          const float f0 = __int2float_rn((int)vh[j] * 256 + (int)vl[j]);
          const float f1 = __int2float_rn((int)vh[j + 1] * 256 + (int)vl[j + 1]);
          orow[c0 + j] = round_clip(fmaf(f0, c2.x, c2.y), oqmin, oqmax) * os0;
          orow[c0 + j + 1] = round_clip(fmaf(f1, c2.z, c2.w), oqmin, oqmax) * os1;
        }
      }
    }
    if (tid < rv) {                                           // the extra column
      const float2 c = cst[Nf - 1];
      outT[(size_t)tid * Nf + Nf - 1] = round_clip(fmaf(__int2float_rn(extra_n[tid]), c.x, c.y), oqmin, oqmax) * (((Nf - 1) & 1) ? os1 : os0);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- one bulk copy for the whole tile; up to 3 rows that would break the 16-byte size rule go by hand
    const int rv4 = rv & ~3;
    if (tid == 0 && rv4 > 0) {
      bulk_s2g(p.Y + (size_t)row0 * Nf, outT, (uint32_t)((size_t)rv4 * Nf * 4));
      bulk_commit();
    }
    for (int i = tid; i < (rv - rv4) * Nf; i += kDecThreads) p.Y[(size_t)(row0 + rv4) * Nf + i] = outT[(size_t)rv4 * Nf + i];
  }
  if (tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- decoder, warp-specialised
// The decoder Dense (K <= 128 aligned input columns, Nfull = Np + 1 <= 257 output columns) in the organisation of
// tc_enc2_kernel: TRANSPOSED product (A = weights, one 128-row operand tile per block of 128 output columns; B = a 64-row
// activation tile whose hi / lo byte planes are stacked into one 128-row operand), double-buffered planes and accumulators
// (2 x 256 TMEM columns), epilogue lane = output column with its scale / bias / low-plane correction in registers and
// coalesced stores straight from registers - the 1028-byte output rows need no shared tile and no alignment.  Roles:
//   warps 0..15   epilogue (TMEM lane quadrant w % 4, rows 16 (w / 4) .. + 15 of the tile), both column blocks; the
//                 quadrant-0 warps also finish the extra output column (an integer dot formed by the converters)
//   warp 16       MMA issuer
//   warps 17..24  converters: 8 rows each per tile, loaded from global one tile ahead (16-byte loads), FQ16 -> planes
constexpr int kDec2Rows = 64;
constexpr int kDec2Epi = 16, kDec2Conv = 8, kDec2Ring = 6, kDec2Sub = 32;   // ring: six slots of 32 input rows (half a tile)
constexpr int kDec2OutBufs = 3;    // output half-tiles (32 rows of Nf floats) in flight
constexpr int kDec2Threads = 32 * (kDec2Epi + 2 + kDec2Conv);

// BN: BatchNorm on the input (B projection).  DIRECT: output rows are line aligned (Nfull = Np, a multiple of 32 floats):
// no extra column, no shared output tile - the epilogue stores from registers as in tc_enc2_kernel (NOUT is unused).
// FX (without BN, DIRECT): the integer decoder of the fixed-point model - integer converters and epilogue, and TWO outputs (payload
// and float image), each half-tile of both leaving as its own bulk store (NOUT = 4: one tile's four buffers).
template <int KV, int NRING, int NOUT, bool BN, bool DIRECT, bool FX = false>
__global__ void __launch_bounds__(kDec2Threads, 1) tc_dec2_kernel(TcLinParams p) {
  static_assert(!FX || (!BN && !DIRECT && NOUT == 4), "the integer instantiation is the decoder");
  constexpr int nring = NRING, nout = NOUT;                   // compile-time: the slot / phase arithmetic sits in every loop
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = p.Kp, Np = p.Np, Nf = p.Nfull;
  const int nblk = (Np + 127) >> 7;                           // blocks of 128 output columns (<= 2)
  const size_t w_bytes = tc_tile_bytes8(kTcRows, Kp);         // one A operand block
  const size_t a_bytes = tc_tile_bytes8(2 * kDec2Rows, Kp);   // one tile buffer: hi rows 0 .. 63, lo - 128 rows 64 .. 127
  unsigned char* Wt = smem;
  unsigned char* planes = Wt + 2 * w_bytes;
  const size_t raw_bytes = (size_t)kDec2Sub * p.Kfull * 4;    // one ring slot: 32 input rows (a multiple of 16 bytes)
  unsigned char* rawb = planes + 2 * a_bytes;
  // output tile with the pitch of the global rows: 64 rows of Nf floats are ONE contiguous block of global memory, which
  // leaves as one bulk store (full lines; per-row warp stores of 128 bytes at a 4-byte alignment reach only half of the
  // HBM write bandwidth: 3.5 of 6.8 TB/s measured, profiles/r02_hbm_write.md)
  float* outT0 = reinterpret_cast<float*>(rawb + (size_t)nring * raw_bytes);   // three half-tiles: one or two leave while the next is assembled
  __shared__ __align__(8) uint64_t planes_full[2], mma_done[2], tmem_free[2], raw_full[kDec2Ring], raw_free[kDec2Ring];
  __shared__ uint32_t tmem_slot;
  __shared__ int extra_n[2][kDec2Rows];
  __shared__ ShiftClamp fx_in;                                // FX: change_cfg of the block output to the decoder's input format
  __shared__ int fx_rs_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (FX && tid == 32) {
    // FxpDense.forward input check (fxpmodel.py:335-347) with a data-dependent input exponent
    const int e = *p.fx_ex_in;
    const bool change = (16 > p.fx_dec_inp_bits) || (e > p.fx_dec_inp_exp);
    const int rs = (change ? p.fx_dec_inp_exp : e) + p.fx_dec_w_exp - p.fx_dec_out_exp;
    fx_in = change ? sc_change_cfg(16, e, p.fx_dec_inp_bits, p.fx_dec_inp_exp) : ShiftClamp{0, 0, INT_MIN, INT_MAX};
    fx_rs_s = rs & 31;
    if ((rs < 0 || rs > 31 || (change && e - p.fx_dec_inp_exp > 31)) && blockIdx.x == 0) atomicOr(p.fx_status, 1);
  }
#ifdef S5_ENC2_TRACE
  long long tw[3] = {0, 0, 0};
  const long long tstart = clock64();
#endif
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&planes_full[i], kDec2Conv); mbar_init(&mma_done[i], 1); mbar_init(&tmem_free[i], kDec2Epi); }
    for (int i = 0; i < nring; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], kDec2Conv / 2); }
    fence_mbar_init();
  }
  {
    const int kchunks = Kp >> 4;
    const uint32_t lbw = tc_lbo(kTcRows);
    for (int i = tid; i < 2 * kTcRows * kchunks; i += kDec2Threads) {
      const int r = i / kchunks, kc = i - r * kchunks;       // r = output column
      const int4 v = r < Np ? *reinterpret_cast<const int4*>(p.W8 + (size_t)r * Kp + kc * 16) : make_int4(0, 0, 0, 0);
      *reinterpret_cast<int4*>(Wt + (size_t)(r >> 7) * w_bytes + (size_t)kc * lbw + (size_t)(r & 127) * 16) = v;
    }
    for (size_t i = (size_t)tid * 16; i < 2 * a_bytes; i += kDec2Threads * 16) *reinterpret_cast<int4*>(planes + i) = make_int4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_slot;
  // row ranges on 32-row boundaries: every tile then starts on a 128-byte boundary of Y (32 x 257 floats = 257 x 128 bytes)
  const long long rb = (p.rows * blockIdx.x / gridDim.x) & ~31LL;
  const long long re = blockIdx.x + 1 == gridDim.x ? p.rows : (p.rows * (blockIdx.x + 1) / gridDim.x) & ~31LL;
  const int ntiles = (int)((re - rb + kDec2Rows - 1) / kDec2Rows);
  const uint32_t la = tc_lbo(kTcRows);

  if (warp == kDec2Epi) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t step = (2 * la) >> 4;
      const uint32_t idesc = tc_idesc_i8(kTcRows, 2 * kDec2Rows, true, true);
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t k = (uint32_t)(t >> 1);
        E2_WAIT(0, mbar_wait(&planes_full[buf], k & 1u));
        if (t >= 2) E2_WAIT(1, mbar_wait(&tmem_free[buf], (k - 1u) & 1u));
        tc_fence_after();
        const uint64_t b_desc = tc_smem_desc(smem_u32(planes + (size_t)buf * a_bytes), la, 128);
        for (int blk = 0; blk < nblk; ++blk) {
          const uint64_t a_desc = tc_smem_desc(smem_u32(Wt + (size_t)blk * w_bytes), la, 128);
          const uint32_t d0 = taddr + (uint32_t)buf * 256u + (uint32_t)blk * 128u;
          for (int j = 0; j < (Kp >> 5); ++j) tc_mma_i8(d0, tc_desc_advance(a_desc, j * step), tc_desc_advance(b_desc, j * step), idesc, j > 0);
        }
        tc_commit(&mma_done[buf]);
      }
    }
  } else if (warp == kDec2Epi + 1) {
    // ---------------------------------------------------------------- TMA producer: half-tile s (32 input rows) -> ring slot s % 3
    if (lane == 0) {
      const int nhalf = (int)((re - rb + kDec2Sub - 1) / kDec2Sub);
      for (int sh = 0; sh < nhalf; ++sh) {
        const int slot = sh % nring;
        if (sh >= nring) mbar_wait(&raw_free[slot], (uint32_t)((sh / nring) - 1) & 1u);
        const long long r0 = rb + (long long)sh * kDec2Sub;
        const uint32_t bytes = (uint32_t)min((long long)kDec2Sub, re - r0) * p.Kfull * 4u;
        mbar_expect_tx(&raw_full[slot], bytes);
        bulk_g2s(rawb + (size_t)slot * raw_bytes, p.X + (size_t)r0 * p.x_stride, bytes, &raw_full[slot]);
      }
    }
  } else if (warp > kDec2Epi + 1) {
    // ---------------------------------------------------------------- converters
    const int cw = warp - kDec2Epi - 2;
    constexpr int RPW = kDec2Rows / kDec2Conv;
    // lane owns input columns 128 i + 4 lane .. + 3 of every 128-column block i (natural K order: one 32-bit plane store)
    uint32_t loff[KV];
    bool wval[KV];
    int kcol[KV], wn[KV][4];
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      loff[i] = (uint32_t)(8 * i + (lane >> 2)) * la + 4u * (lane & 3);
      wval[i] = 128 * i + 4 * lane < Kp;
      kcol[i] = 128 * i + 4 * lane < p.Kfull ? 128 * i + 4 * lane : 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) wn[i][e] = wval[i] && p.WN ? (int)p.WN[128 * i + 4 * lane + e] : 0;
    }
    const float inv = p.in_inv, qmin = p.in_qmin, qmax = p.in_qmax;
    // BatchNorm folded with the (power-of-two) input scale: ((x - sub) * mul + add) / s == (x - sub) * (mul / s) + add / s
    float bsub[BN ? KV : 1][4], bmul[BN ? KV : 1][4], badd[BN ? KV : 1][4];
    if constexpr (BN) {
#pragma unroll
      for (int i = 0; i < KV; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = min(128 * i + 4 * lane + e, p.Kfull - 1);
          bsub[i][e] = p.in_sub[k]; bmul[i][e] = p.in_mul[k] * inv; badd[i][e] = p.in_add ? p.in_add[k] * inv : 0.f;
        }
    }
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int hh = cw / (kDec2Conv / 2), sh = 2 * t + hh, slot = sh % nring;   // this warp's rows lie in half-tile sh
      const int rvh = (int)min((long long)kDec2Sub, re - (rb + (long long)sh * kDec2Sub));   // rows of that half (<= 0: none)
      if (t >= 2) {                                           // tile t - 2: its MMAs have read the planes, its epilogue the extra column
        E2_WAIT(0, mbar_wait(&mma_done[buf], (uint32_t)((t >> 1) - 1) & 1u));
        E2_WAIT(1, mbar_wait(&tmem_free[buf], (uint32_t)((t >> 1) - 1) & 1u));
      }
      unsigned char* Ah = planes + (size_t)buf * a_bytes;
      unsigned char* Al = Ah + kDec2Rows * 16;
      if (rvh > 0) {
        E2_WAIT(2, mbar_wait(&raw_full[slot], (uint32_t)(sh / nring) & 1u));
        const float* rs = reinterpret_cast<const float*>(rawb + (size_t)slot * raw_bytes);
        const int lr0 = (cw - hh * (kDec2Conv / 2)) * RPW;     // first row of this warp inside the half-tile
        // four rows at a time: enough independent work per lane, few enough live registers (no spills at 832 threads)
#pragma unroll
        for (int q0 = 0; q0 < RPW; q0 += 4) {
          float4 raw[4][KV];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < KV; ++i)
              raw[q][i] = *reinterpret_cast<const float4*>(rs + (size_t)min(lr0 + q0 + q, rvh - 1) * p.Kfull + kcol[i]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = cw * RPW + q0 + q;
            int part = 0;
#pragma unroll
            for (int i = 0; i < KV; ++i) {
              const float xv[4] = {raw[q][i].x, raw[q][i].y, raw[q][i].z, raw[q][i].w};
              uint32_t b4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if constexpr (FX) b4[e] = (uint32_t)sc_apply(__float_as_int(xv[e]), fx_in);
                else if constexpr (BN) b4[e] = quant_bits((xv[e] - bsub[i][e]) * bmul[i][e] + badd[i][e], 1.f, qmin, qmax);
                else b4[e] = quant_bits(xv[e], inv, qmin, qmax);
                if constexpr (!DIRECT) part += (int)(short)b4[e] * wn[i][e];
              }
              uint32_t lo, hi;
              split_planes(b4[0], b4[1], b4[2], b4[3], lo, hi);
              if (wval[i]) {
                *reinterpret_cast<uint32_t*>(Ah + loff[i] + (uint32_t)r * 16u) = hi;
                *reinterpret_cast<uint32_t*>(Al + loff[i] + (uint32_t)r * 16u) = lo ^ 0x80808080u;
              }
            }
            if constexpr (!DIRECT) {
              const int tot = __reduce_add_sync(0xffffffffu, part);
              if (lane == 0) extra_n[buf][r] = tot;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_free[slot]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&planes_full[buf]);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: lane = output column, warp / 4 = 16 rows of the tile
    const int quad = warp & 3, rq = warp >> 2;
    const float oqmin = p.relu ? fmaxf(p.out_qmin, 0.f) : p.out_qmin, oqmax = p.out_qmax;
    const uint32_t lane_base = (uint32_t)(32 * quad) << 16;
    const size_t ystride = (size_t)p.y_stride;
    float cs[2], cb[2], os[2];
    int corr[2];
    bool ncol[2];
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      const int n = 128 * blk + 32 * quad + lane;
      ncol[blk] = blk < nblk && n < Np;
      const float oinv = (n & 1) ? p.out_inv1 : p.out_inv0;
      os[blk] = (n & 1) ? p.out_scale1 : p.out_scale0;
      cs[blk] = (!FX && ncol[blk]) ? p.col_scale[n] * oinv : 0.f;
      cb[blk] = (!FX && ncol[blk] && p.bias) ? p.bias[n] * oinv : 0.f;
      int c = 0;                                              // 128 * sum_k W[n][k]: undoes the lo - 128 of the stacked low plane
      if (ncol[blk]) {
        const int4* wr = reinterpret_cast<const int4*>(p.W8 + (size_t)n * Kp);
        for (int i = 0; i < (Kp >> 4); ++i) {
          const int4 w = wr[i];
          c = __dp4a(w.x, 0x01010101, c); c = __dp4a(w.y, 0x01010101, c); c = __dp4a(w.z, 0x01010101, c); c = __dp4a(w.w, 0x01010101, c);
        }
      }
      corr[blk] = 128 * c;
    }
    // the extra column (Nf = Np + 1): (s_in s_w / s_out, bias / s_out) of column Nf - 1
    const bool xn = Nf > Np;
    const float xoinv = ((Nf - 1) & 1) ? p.out_inv1 : p.out_inv0, xos = ((Nf - 1) & 1) ? p.out_scale1 : p.out_scale0;
    const float xcs = (!FX && xn) ? p.col_scale[Nf - 1] * xoinv : 0.f, xcb = (!FX && xn && p.bias) ? p.bias[Nf - 1] * xoinv : 0.f;
    // FX: fxp_matmul + FxpDense bias on the exact int32 sum (fxparray.py:640-678, fxpmodel.py:352-364)
    const int frs = FX ? fx_rs_s : 0;
    const int fhi = (FX && p.fx_out_bits < 32) ? (1 << (p.fx_out_bits - 1)) - 1 : INT_MAX, flo = -fhi - 1;
    int fbias[2] = {0, 0};
    if constexpr (FX) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) fbias[blk] = (ncol[blk] && p.fx_bias) ? p.fx_bias[128 * blk + 32 * quad + lane] : 0;
    }
    const int fxbias = (FX && xn && p.fx_bias) ? p.fx_bias[Nf - 1] : 0;
    auto fx_out = [&](int acc, int bias) -> int {
      int v = min(max(acc >> frs, flo), fhi);
      if (p.fx_bias) v = min(max(fx_addw(v, bias), flo), fhi);
      return v;
    };
    const size_t half_elems = (size_t)kDec2Sub * Nf;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const long long row0 = rb + (long long)t * kDec2Rows;
      const int rv = (int)min((long long)kDec2Rows, re - row0);
      E2_WAIT(0, mbar_wait(&mma_done[buf], (uint32_t)(t >> 1) & 1u));
      tc_fence_after();
      if constexpr (DIRECT) {
        const int c0 = 16 * rq;
        if (c0 < rv) {
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            if (blk >= nblk) break;
            const uint32_t d0 = taddr + (uint32_t)buf * 256u + (uint32_t)blk * 128u + lane_base;
            uint32_t vh[16], vl[16];
            tc_ld16(d0 + c0, vh);
            tc_ld16(d0 + kDec2Rows + c0, vl);
            tc_wait_ld();
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              o[j] = round_clip(fmaf(__int2float_rn((int)vh[j] * 256 + (int)vl[j] + corr[blk]), cs[blk], cb[blk]), oqmin, oqmax) * os[blk];
            float* q = p.Y + (size_t)(row0 + c0) * ystride + 128 * blk + 32 * quad + lane;
            if (ncol[blk]) {
              if (c0 + 16 <= rv) {
#pragma unroll
                for (int j = 0; j < 16; ++j, q += ystride) *q = o[j];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j, q += ystride)
                  if (c0 + j < rv) *q = o[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_free[buf]);
        continue;
      }
      if (tid == 0) {                                         // three buffers: the second half of the previous tile may still be leaving
        if (nout >= 3 && !FX) bulk_wait_read1(); else bulk_wait_read0();   // FX: the four buffers hold one tile (2 halves x 2 outputs)
      }
      asm volatile("bar.sync 2, %0;" ::"r"(32 * kDec2Epi) : "memory");
      const int c0 = 16 * rq, hh = rq >> 1;                   // rows c0 .. c0 + 15 of the tile lie in half-tile 2 t + hh
      float* outT = outT0 + (size_t)(FX ? 2 * hh : (2 * t + hh) % nout) * half_elems;   // FX: payload buffer; its float image follows it
      if (c0 < rv) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          if (blk >= nblk) break;
          const uint32_t d0 = taddr + (uint32_t)buf * 256u + (uint32_t)blk * 128u + lane_base;
          uint32_t vh[16], vl[16];
          tc_ld16(d0 + c0, vh);
          tc_ld16(d0 + kDec2Rows + c0, vl);
          tc_wait_ld();
          float* q = outT + (size_t)(c0 - kDec2Sub * hh) * Nf + 128 * blk + 32 * quad + lane;   // lanes along columns: conflict-free
          if (ncol[blk]) {
            if constexpr (FX) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int v = fx_out((int)(vh[j] * 256u + vl[j] + (uint32_t)corr[blk]), fbias[blk]);
                q[(size_t)j * Nf] = __int_as_float(v);
                q[half_elems + (size_t)j * Nf] = __int2float_rn(v) * p.fx_out_scale;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                q[(size_t)j * Nf] = round_clip(fmaf(__int2float_rn((int)vh[j] * 256 + (int)vl[j] + corr[blk]), cs[blk], cb[blk]), oqmin, oqmax) * os[blk];
            }
          }
        }
        if (xn && quad == 0 && lane < 16) {
          float* qx = outT + (size_t)(c0 - kDec2Sub * hh + lane) * Nf + Nf - 1;
          if constexpr (FX) {
            const int v = fx_out(extra_n[buf][c0 + lane], fxbias);
            qx[0] = __int_as_float(v);
            qx[half_elems] = __int2float_rn(v) * p.fx_out_scale;
          } else {
            qx[0] = round_clip(fmaf(__int2float_rn(extra_n[buf][c0 + lane]), xcs, xcb), oqmin, oqmax) * xos;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[buf]);
      asm volatile("bar.sync 2, %0;" ::"r"(32 * kDec2Epi) : "memory");
      // one bulk copy per half-tile; up to 3 rows that would break the 16-byte size rule go by hand (end of the input)
      const int rv4 = rv & ~3;
      if (tid == 0) {
        for (int h2 = 0; h2 < 2; ++h2) {
          const int nr = min(kDec2Sub, rv4 - kDec2Sub * h2);
          if constexpr (FX) {
            if (nr > 0 && p.Y) bulk_s2g(p.Y + (size_t)(row0 + kDec2Sub * h2) * ystride, outT0 + (size_t)(2 * h2) * half_elems, (uint32_t)((size_t)nr * Nf * 4));
            if (nr > 0 && p.fx_Yf) bulk_s2g(p.fx_Yf + (size_t)(row0 + kDec2Sub * h2) * ystride, outT0 + (size_t)(2 * h2 + 1) * half_elems, (uint32_t)((size_t)nr * Nf * 4));
          } else {
            if (nr > 0) bulk_s2g(p.Y + (size_t)(row0 + kDec2Sub * h2) * ystride, outT0 + (size_t)((2 * t + h2) % nout) * kDec2Sub * Nf, (uint32_t)((size_t)nr * Nf * 4));
          }
          bulk_commit();
        }
      }
      if (rv4 != rv) {
        const int h2 = rv4 / kDec2Sub;
        const float* src = outT0 + (size_t)(FX ? 2 * h2 : (2 * t + h2) % nout) * half_elems + (size_t)(rv4 - kDec2Sub * h2) * Nf;
        for (int i = tid; i < (rv - rv4) * Nf; i += 32 * kDec2Epi) {
          if (!FX || p.Y) p.Y[(size_t)(row0 + rv4) * ystride + i] = src[i];
          if constexpr (FX) { if (p.fx_Yf) p.fx_Yf[(size_t)(row0 + rv4) * ystride + i] = src[half_elems + i]; }
        }
        asm volatile("bar.sync 2, %0;" ::"r"(32 * kDec2Epi) : "memory");
      }
    }
    if (!DIRECT && tid == 0) bulk_wait0();
  }
#ifdef S5_ENC2_TRACE
  if (blockIdx.x == 3 && lane == 0 && (warp == 0 || warp == 5 || warp == 15 || warp == 16 || warp == 18 || warp == 25))
    printf("[dec2 trace] warp %d: total %lld cyc, wait0 %lld, wait1 %lld, wait2 %lld (tiles %d)\n", warp, clock64() - tstart, tw[0], tw[1], tw[2], ntiles);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- C projection + GLU
struct TcGluParams {
  const float* XS;   // [rows][2P] scan output (complex64 interleaved, ReLU applied)
  const float* X;    // [rows][H] block input (BatchNorm source, skip connection)
  float* OUT;        // [rows][H]
  long long rows; int H, K1;   // K1 = 2P <= 128
  const __nv_bfloat16* Wc;     // [H][K1] codes of (C_re, -C_im) interleaved along k
  const signed char* Wo;       // [H][128] int8 codes of the out2 kernel, K in the lane-interleaved order, zero padded
  float xs_odd_mul;            // s_C_im / s_C_re folded into the odd (imaginary) inputs
  float c_scale;               // conj_mul * s_C_re
  const float* bn_sub; const float* bn_mul; const float* bn_add;   // nullable (no BatchNorm)
  const float* D; const float* o_bias;
  float o_col_scale;           // s_o_act * s_w_out2
  float ut_s, ut_inv, yt_s, yt_inv, oa_inv, oo_s, oo_inv, gl_s, gl_inv, gr_s, gr_inv;
  float qs_min, qs_max, qn_min, qn_max;
  int relu_out;
  // fused recurrence (SCAN instantiation): XS holds Bu [B][L][P] complex64 and the kernel runs x_k = lam x_{k-1} + Bu_k itself
  const float2* lam;           // [P]
  int B, L, R, tiles_per_clip; // tile = R <= 128 consecutive frames of one clip
  int relu_scan;               // lexicographic complex ReLU on the states handed to C (relufication)
  uint32_t* flags;             // [B][tiles_per_clip] look-back flags (0 none, 1 aggregate, 2 inclusive prefix), zeroed per launch
  float2* recs;                // [B][tiles_per_clip][2][P] aggregate / inclusive prefix
  // fused B projection (BPJ instantiation): Bu = FQ(B_bar FQ(BN(x))) is computed in the kernel as well, XS is unused
  const signed char* WB;       // [K1][128] int8 codes of B_bar (rows = (re, im) interleaved outputs), lane-interleaved K
  const float* b_col_scale;    // [K1] s_u * s_w[n]
  float b_out_scale0, b_out_scale1, b_out_inv0, b_out_inv1;   // FQ grid of Re(Bu) / Im(Bu)
  // BPJ: the layer's shared-memory image (weight tiles in the operand layout + per-column constants), built once per handle by
  // tc_glu_image_kernel; a layer switch is then ONE bulk copy instead of ~55 KB of strided loads with index arithmetic
  const unsigned char* img; uint32_t img_bytes;
};

// All blocks of the model in ONE launch (block kernel only): the tiles of layer l + 1 follow those of layer l in the same
// round-robin order, so a pipeline that would idle in the last, partial round of a layer starts on the next layer
// instead.  Tile (l + 1, clip, chunk) reads exactly the rows tile (l, clip, chunk) wrote: one done-flag per tile.
constexpr int kMaxFusedLayers = 4;
struct TcGluMulti {
  int nlayers;
  uint32_t* done;              // [nlayers - 1][tiles per layer], zeroed per launch (nullptr when nlayers == 1)
  TcGluParams layer[kMaxFusedLayers];
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cfmaf(float2 a, float2 x, float2 v) {   // a * x + v
  return make_float2(fmaf(a.x, x.x, fmaf(-a.y, x.y, v.x)), fmaf(a.x, x.y, fmaf(a.y, x.x, v.y)));
}

constexpr int kGluThreads = 512;   // two independent 256-thread tile pipelines per CTA that share the weight tiles
constexpr int kGluHalf = 256;
constexpr int kGluK2 = 128;        // padded K of the out2 product (lane-interleaved order: byte 4 lane + j <-> column lane + 32 j)

__device__ __forceinline__ void half_barrier(int half) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(kGluHalf) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {   // {lo half = bf16(a), hi half = bf16(b)}, round to nearest even
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// The per-layer image of the block kernel, laid out exactly as in its shared memory: C tile | out2 tile | B tile (BPJ) |
// out2 bias / s_o_out [H] | B column scales [K1] | per-column constants [4][H].  `dst` is shared memory (direct fill) or the
// handle's global image (tc_glu_image_kernel).  All scales are powers of two, so folding them into the constants changes
// no rounding: (x - sub) * (mul / s_u) + add / s_u == BN(x) / s_u.
__host__ __device__ inline size_t tc_glu_image_bytes(int H, int K1, bool bpj) {
  return tc_tile_bytes(H, K1) + tc_tile_bytes8(H, kGluK2) + (bpj ? tc_tile_bytes8(K1, kGluK2) : 0) + (size_t)(5 * H + K1) * 4;
}
template <bool BPJ>
__device__ __forceinline__ void tc_glu_fill_image(unsigned char* dst, const TcGluParams& p, int H, int K1, int tid, int nthreads) {
  unsigned char* Wc = dst;
  unsigned char* Wo = Wc + tc_tile_bytes(H, K1);
  unsigned char* WB = Wo + tc_tile_bytes8(H, kGluK2);
  float* ob = reinterpret_cast<float*>(WB + (BPJ ? tc_tile_bytes8(K1, kGluK2) : 0));
  float* cstB = ob + H;
  float* lcst = cstB + K1;
  tc_load_weights(Wc, reinterpret_cast<const __half*>(p.Wc), H, K1);
  tc_load_weights8(Wo, p.Wo, H, kGluK2);
  if constexpr (BPJ) {
    tc_load_weights8(WB, p.WB, K1, kGluK2);
    for (int n = tid; n < K1; n += nthreads) cstB[n] = p.b_col_scale[n] * ((n & 1) ? p.b_out_inv1 : p.b_out_inv0);
  }
  for (int n = tid; n < H; n += nthreads) ob[n] = p.o_bias[n] * p.oo_inv;
  for (int c = tid; c < H; c += nthreads) {
    lcst[c] = p.bn_sub ? p.bn_sub[c] : 0.f;
    lcst[H + c] = (p.bn_sub ? p.bn_mul[c] : 1.f) * p.ut_inv;
    lcst[2 * H + c] = (p.bn_sub ? p.bn_add[c] : 0.f) * p.ut_inv;
    lcst[3 * H + c] = p.D[c] * p.ut_s * p.yt_inv;             // D u / s_y = (D s_u / s_y) * code(u)
  }
}
__global__ void tc_glu_image_kernel(TcGluParams p, unsigned char* img) {
  tc_glu_fill_image<true>(img, p, p.H, p.K1, threadIdx.x, blockDim.x);
}

// Per 128-row tile (HJ = H / 32 columns per lane in the lanes-along-columns passes):
//   S  : xs rows -> bf16 planes hi = bf16(x), lo = bf16(x - hi)                       (thread: 4 consecutive k of a row)
//   MMA1 (bf16): acc1 = (hi + lo) * Wc^T
//   E1 : acc1 -> fp32 tile T (TMEM lane = row)
//   P1 : lanes along columns, x read coalesced from global: u = FQ(BN(x)); y = FQ(conj Re(C x) + D u); x1 = ReLU(y);
//        left-gate code FQ_gl(x1) -> T (int16, in place of the row); out2 input codes -> int8 planes (hi / lo byte)
//   MMA2 (int8): acc2 = 256 (hi * Wo^T) + lo * Wo^T, exact
//   E2 : o = FQ(out2) ; right-gate code FQ(sigmoid(o)) -> T (int16, second half of the row)
//   P2 : lanes along columns: out = ReLU(l * sg + x), coalesced stores
//
// SCAN: the S pass runs the diagonal recurrence itself.  A tile is R <= 128 consecutive frames of ONE clip, tiles are
// handed out clip-fastest and round-robin over the pipelines; thread (state p, segment s) keeps its <= 32 Bu values in
// registers, scans them twice (local aggregate, then with the true carry-in) and the carry across tiles of a clip
// travels by decoupled look-back (aggregate / inclusive-prefix records + acquire / release flags in global memory).
// The states never touch HBM: they go from registers straight into the bf16 operand planes of the C projection.
// BPJ (implies SCAN): the B projection runs in the kernel, too: x tile -> u codes -> int8 planes -> MMA -> FQ(Bu) -> a
// shared-memory tile that the recurrence threads read column-wise.  A whole S5 block is then ONE kernel that reads
// the block input and writes the block output; Bu and the states never leave the SM.
// KJ > 0: K1 = 2P = 32 KJ is a compile-time constant as well (the NDNS shapes), so that the column loops of the TMEM passes
// unroll and every tile / plane / TMEM address in them is an immediate offset instead of run-time integer arithmetic.
template <int HJ, bool SCAN, bool BPJ, int KJ = 0>
__global__ void __launch_bounds__(kGluThreads, 1) tc_glu_kernel(const __grid_constant__ TcGluMulti mp) {
  static_assert(SCAN || !BPJ, "the fused B projection needs the fused recurrence");
  TC_TRACE_ENTRY
  extern __shared__ __align__(128) unsigned char smem[];
  const TcGluParams& p0 = mp.layer[0];
  constexpr int H = 32 * HJ;
  const int K1 = KJ > 0 ? 32 * KJ : p0.K1;                    // the same for every layer
  const int ldT = H + 4;                                      // floats per row of T
  const size_t wc_bytes = tc_tile_bytes(H, K1), wo_bytes = tc_tile_bytes8(H, kGluK2);
  const size_t p1_bytes = tc_tile_bytes(kTcRows, K1);         // one bf16 plane of the xs tile
  const size_t t_bytes = (size_t)kTcRows * ldT * 4;
  const size_t a2_bytes = tc_tile_bytes8(kTcRows, kGluK2);    // one int8 plane of the x1 codes
  const int ldB = K1 + 4;                                     // floats per row of the Bu tile (BPJ)
  const size_t bu_bytes = BPJ ? (size_t)kTcRows * ldB * 4 : 0;
  const size_t seg_bytes = SCAN ? (size_t)(2 * kGluHalf + 128) * sizeof(float2) : 0;
  const size_t r_bytes = max(max(2 * p1_bytes, bu_bytes) + seg_bytes, t_bytes + 2 * a2_bytes);
  const size_t wb_bytes = BPJ ? tc_tile_bytes8(K1, kGluK2) : 0;
  unsigned char* Wc = smem;
  unsigned char* Wo = Wc + wc_bytes;
  unsigned char* WB = Wo + wo_bytes;
  float* ob = reinterpret_cast<float*>(WB + wb_bytes);        // [H] out2 bias / s_o_out
  float* cstB = ob + H;                                       // [K1] s_u s_w / s_Bu (BPJ)
  float* lcst = cstB + K1;                                    // [4][H] per-column constants of the lanes-along-columns passes
  const int tid = threadIdx.x, half = tid >> 8, htid = tid & (kGluHalf - 1), warp = htid >> 5, lane = tid & 31;
  unsigned char* Rbase = reinterpret_cast<unsigned char*>(ob) + (((size_t)(5 * H + K1) * 4 + 127) & ~(size_t)127);
  unsigned char* R = Rbase + (size_t)half * r_bytes;
  // SCAN: segment aggregates + multipliers [nseg][P] and the tile carry [P], behind the planes / the Bu tile
  float2* seg_agg = reinterpret_cast<float2*>(R + max(2 * p1_bytes, bu_bytes));
  float2* seg_mul = seg_agg + kGluHalf;
  float2* carry_s = seg_mul + kGluHalf;
  float* BuT = reinterpret_cast<float*>(R);                   // BPJ: FQ(Bu) tile [128][K1 + 4], dead once the recurrence holds it
  unsigned char* P1h = R;                                     // xs planes (dead after MMA 1)
  unsigned char* P1l = R + p1_bytes;
  float* T = reinterpret_cast<float*>(R);                     // aliases the xs planes
  unsigned char* A2h = R + t_bytes;
  unsigned char* A2l = A2h + a2_bytes;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  constexpr int NW = kGluHalf / 32;
  if (tid < 32) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  const long long last_row = p0.rows - 1;
  const uint32_t lbo = tc_lbo(kTcRows);
  const bool kin = 4 * lane < K1;
  const int xs_col = kin ? 4 * lane : 0;                      // loads are never predicated: pad lanes / rows re-read valid data
  const uint32_t soff = (uint32_t)(lane >> 1) * lbo + (uint32_t)(lane & 1) * 8u;   // bf16 planes: k = 4 lane
  const uint32_t aoff = (uint32_t)(lane >> 2) * lbo + 4u * (lane & 3);             // int8 planes: bytes 4 lane .. 4 lane + 3
  const int erow0 = (warp & 3) * 32;                          // first tile row of this warp's TMEM lane quadrant
  const int erow = erow0 + lane;                              // TMEM lane = tile row
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  const int cw = H >> 1, cbeg = (warp >> 2) * cw;             // column half of this warp in the TMEM passes
  const float qs_min = p0.qs_min, qs_max = p0.qs_max, qn_min = p0.qn_min, qn_max = p0.qn_max;
  const float qrelu_min = fmaxf(qs_min, 0.f);
  // SCAN per-thread constants: state, segment
  const int P = K1 >> 1;
  const int sp = SCAN ? htid % P : 0, sseg = SCAN ? htid / P : 0, nseg = SCAN ? kGluHalf / P : 1;
  // rows per segment (<= 32); the KJ instantiations are only launched with full-height tiles (R = 128 = 32 nseg)
  const int sr = !SCAN ? 0 : (KJ > 0 ? 32 : (p0.R + nseg - 1) / nseg);
  const uint32_t poff = (uint32_t)(sp >> 2) * lbo + 4u * (sp & 3);   // plane offset of column k = 2 sp
  // ---- per-layer state: weight tiles in shared memory, per-lane constants in registers
  // All scales are powers of two, so folding them into the BatchNorm / feed-through constants changes no rounding:
  // (x - sub) * (mul / s_u) + add / s_u == BN(x) / s_u.
  float2 lam = make_float2(0.f, 0.f), lamR = make_float2(1.f, 0.f), lamS = make_float2(1.f, 0.f);
  __shared__ __align__(8) uint64_t wbar;                      // BPJ: completion of the layer-image bulk copy
  uint32_t wpar = 0;
  if (tid == 0) {
    mbar_init(&wbar, 1);
    fence_mbar_init();
  }
  auto load_layer = [&](int l) {                              // called by all threads of the CTA, between two __syncthreads
    const TcGluParams& p = mp.layer[l];
    if (BPJ && p.img) {
      __syncthreads();                                        // (first call) the mbarrier is initialised
      if (tid == 0) {
        fence_proxy_async_smem();
        mbar_expect_tx(&wbar, p.img_bytes);
        bulk_g2s(Wc, p.img, p.img_bytes, &wbar);
      }
      mbar_wait(&wbar, wpar);
      wpar ^= 1u;
    } else {
      tc_glu_fill_image<BPJ>(Wc, p, H, K1, tid, kGluThreads);
    }
    if constexpr (SCAN) {
      lam = p.lam[sp];
      lamR = make_float2(1.f, 0.f); lamS = make_float2(1.f, 0.f);
      float2 sq = lam;                                        // lam^R (full tile) and lam^sr (full segment), binary powers
      for (int bit = 0; bit < 8; ++bit) {
        if ((p.R >> bit) & 1) lamR = cmulf(lamR, sq);
        if ((sr >> bit) & 1) lamS = cmulf(lamS, sq);
        sq = cmulf(sq, sq);
      }
    }
  };
  // the int8 planes carry a zero pad byte per lane (column 4 lane + 3 when HJ < 4): clear them once
  for (size_t i = (size_t)htid * 16; i < 2 * a2_bytes; i += kGluHalf * 16) *reinterpret_cast<int4*>(A2h + i) = make_int4(0, 0, 0, 0);
  load_layer(0);
  int cur_layer = 0;
  uint32_t phase = 0;
  TC_TRACE_DECL
  __syncthreads();
  const uint32_t tcol = tmem_slot + 256u * half;
  // Without SCAN: balanced partition, every pipeline owns one contiguous row range (full tiles plus one short tile whose
  // row loops and TMEM passes stop at its last row).  With SCAN: tile t = chunk * B + clip, handed out round-robin.
  const long long npipe = 2LL * gridDim.x, pipe = 2LL * blockIdx.x + half;
  const long long row_begin = SCAN ? 0 : p0.rows * pipe / npipe, row_end = SCAN ? 0 : p0.rows * (pipe + 1) / npipe;
  const long long ntiles = SCAN ? (long long)p0.B * p0.tiles_per_clip : 0;    // per layer
  const long long ntp = (ntiles + 1) & ~1LL;                  // padded to even: the two pipelines of a CTA share a layer
  const long long total = ntp * mp.nlayers;
  // SCAN tile schedule.  Global tile index g = layer * ntp + t, t = chunk * B + clip; round r hands g = r * G + pipe to
  // pipeline `pipe`.  When the last, partial round lies inside the last layer its tiles go to ONE pipeline per SM first
  // (g = full * G + half * grid + blockIdx), so that an SM gets at most one tile more than any other.  Every pipeline
  // walks its tiles in increasing order and a tile only ever waits for tiles with a smaller index that belong to
  // another, co-resident pipeline (look-back: t - B of the same layer; layers: the same t one layer down), so the
  // waits cannot deadlock.  The iteration count and the layer of an iteration are CTA-uniform: the weight tiles are
  // shared by the two pipelines and are reloaded between two __syncthreads when the layer changes.
  const long long rounds_all = SCAN ? (total + npipe - 1) / npipe : 0;
  const long long full_rounds = SCAN ? total / npipe : 0;
  const long long rem = SCAN ? total - full_rounds * npipe : 0;
  const bool spread = rem > 0 && full_rounds * npipe >= ntp * (mp.nlayers - 1);
  const long long last_j = (long long)half * gridDim.x + blockIdx.x;
  long long iter = 0;
  for (long long it0 = row_begin; SCAN ? iter < rounds_all : it0 < row_end; it0 += kTcRows, ++iter) {
    long long it = it0;
    int layer = 0;
    if constexpr (SCAN) {
      const long long g0 = iter * npipe + 2LL * blockIdx.x;   // CTA-uniform
      long long g = g0 + half;
      bool active = g < total;
      layer = (int)((uint32_t)min(g0, total - 1) / (uint32_t)ntp);   // tile counts are far below 2^31: 32-bit division
      if (iter == full_rounds && spread) {
        g = full_rounds * npipe + last_j;
        active = last_j < rem;
        layer = mp.nlayers - 1;
      }
      if (layer != cur_layer) {                               // CTA-uniform
        __syncthreads();
        load_layer(layer);
        cur_layer = layer;
        __syncthreads();
      }
      it = g - (long long)layer * ntp;
      if (!active || it >= ntiles) continue;
    }
    // scalars of the current layer, straight from the parameter bank (cheap to re-derive, no registers held)
    const TcGluParams& p = mp.layer[layer];
    const float k_c = p.c_scale * p.yt_inv;                   // conj Re(C x) / s_y = acc * k_c
    const float a_o = p.o_col_scale * p.oo_inv;
    const float k_e = -p.oo_s * 1.4426950408889634f;          // exp(-o) = 2^(code(o) * k_e)
    const float y_to_gl = p.yt_s * p.gl_inv, y_to_oa = p.yt_s * p.oa_inv;   // powers of two
    const float xs_odd_mul = p.xs_odd_mul;
    const bool relu_scan = p.relu_scan != 0;
    const float gl_s = p.gl_s, gr_s = p.gr_s, gr_inv = p.gr_inv, out_floor = p.relu_out ? 0.f : -INFINITY;
    const float b_os0 = p.b_out_scale0, b_os1 = p.b_out_scale1;
    const float* pXS = p.XS; const float* pX = p.X; float* pOUT = p.OUT;
    uint32_t* pflags = p.flags; float2* precs = p.recs;
    long long row0;
    int rv;
    int clip = 0, chunk = 0;
    if constexpr (SCAN) {
      chunk = (int)((uint32_t)it / (uint32_t)p0.B);
      clip = (int)it - chunk * p0.B;
      row0 = (long long)clip * p0.L + (long long)chunk * p0.R;
      rv = min(p0.R, p0.L - chunk * p0.R);
      if (layer > 0) {
        // The rows of this tile come from the same tile one layer down, possibly written by another SM.  Acquire load of
        // its done flag by one thread (relaxed polling + one acquire fence, which also drops this SM's stale L1 lines), then a barrier: every later load of
        // the pipeline is ordered after the producer's stores.
        if (htid == 0) {
          while (ld_relaxed_u32(mp.done + (size_t)(layer - 1) * ntiles + it) == 0u) {}
          fence_acquire_gpu();
        }
        half_barrier(half);
      }
    } else {
      row0 = it;
      rv = (int)min((long long)kTcRows, row_end - row0);
    }
    // per-tile base pointer and 32-bit row clamp: loads are never predicated, rows past the end re-read the last row
    const float* xtile = pX + (size_t)row0 * H;
    const int rlim = (int)min((long long)kTcRows, last_row - row0);
    TC_TRACE_MARK(tid == 0);
    if constexpr (SCAN) {
      // ---- S (fused recurrence): Bu -> registers -> local aggregate -> look-back -> states -> bf16 planes
      const int len = max(0, min(sr, rv - sseg * sr));
      float2 v[32];
      if constexpr (BPJ) {
        // ---- S0: block input -> u = FQ(BN(x)) codes -> int8 planes (lanes along columns, as in P1)
        float bsub[HJ], bmulq[HJ], baddq[HJ];
#pragma unroll
        for (int j = 0; j < HJ; ++j) {
          bsub[j] = lcst[lane + 32 * j]; bmulq[j] = lcst[H + lane + 32 * j]; baddq[j] = lcst[2 * H + lane + 32 * j];
        }
        for (int rb = warp * 8; rb < rv; rb += NW * 8) {
          float xr8[8][HJ];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float* xr = xtile + min(rb + q, rlim) * H;
#pragma unroll
            for (int j = 0; j < HJ; ++j) xr8[q][j] = xr[lane + 32 * j];
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t ub[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < HJ; ++j)
              ub[j] = __float_as_uint(fminf(fmaxf((xr8[q][j] - bsub[j]) * bmulq[j] + baddq[j], qs_min), qs_max) + kRoundMagic);
            uint32_t lo, hi;
            split_planes(ub[0], ub[1], ub[2], ub[3], lo, hi);
            *reinterpret_cast<uint32_t*>(A2h + aoff + (uint32_t)(rb + q) * 16u) = hi;
            *reinterpret_cast<uint32_t*>(A2l + aoff + (uint32_t)(rb + q) * 16u) = lo;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        half_barrier(half);
        tc_fence_after();
        TC_TRACE_MARK(tid == 0);
        if (htid == 0) {
          tc_issue8(tcol, A2h, WB, kGluK2, K1, K1, true);
          tc_issue8(tcol + 128, A2l, WB, kGluK2, K1, K1, false);
          tc_commit(&bar[half]);
        }
        mbar_wait(&bar[half], phase);
        phase ^= 1;
        tc_fence_after();
        TC_TRACE_MARK(tid == 0);
        // ---- E_B: Bu = FQ(acc * s_u s_w) -> shared tile (the planes are dead)
        if (erow0 < rv) {
          const int cwb = K1 >> 1, cb0 = (warp >> 2) * cwb;
          // two register sets: the TMEM loads of the next 8 columns are in flight while these 8 are post-processed
          auto eb8 = [&](const uint32_t (&vh)[8], const uint32_t (&vl)[8], int c0) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 c4 = *reinterpret_cast<const float4*>(cstB + c0 + j);
              const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float f = __int2float_rn((int)vh[j + e] * 256 + (int)vl[j + e]);
                o[j + e] = round_clip(f * cc[e], qs_min, qs_max) * ((e & 1) ? b_os1 : b_os0);
              }
            }
            float* dst = BuT + (size_t)erow * ldB + c0;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
          };
          uint32_t ah[8], al[8], bh[8], bl[8];
          tc_ld8(tcol + lane_base + cb0, ah);
          tc_ld8(tcol + lane_base + 128 + cb0, al);
          for (int c0 = cb0; c0 < cb0 + cwb; c0 += 16) {     // cwb is a multiple of 16
            tc_wait_ld();
            tc_ld8(tcol + lane_base + c0 + 8, bh);
            tc_ld8(tcol + lane_base + 128 + c0 + 8, bl);
            eb8(ah, al, c0);
            tc_wait_ld();
            if (c0 + 16 < cb0 + cwb) {
              tc_ld8(tcol + lane_base + c0 + 16, ah);
              tc_ld8(tcol + lane_base + 128 + c0 + 16, al);
            }
            eb8(bh, bl, c0 + 8);
          }
        }
        tc_fence_before();
        half_barrier(half);
        TC_TRACE_MARK(tid == 0);
        // the recurrence threads pick their 32 frames of one state out of the tile
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v[i] = *reinterpret_cast<const float2*>(BuT + (size_t)(KJ > 0 ? sseg * sr + i : min(sseg * sr + i, kTcRows - 1)) * ldB + 2 * sp);
      } else {
        const float2* bu = reinterpret_cast<const float2*>(pXS) + (size_t)row0 * P + sp;
        // 32 unpredicated, unclamped loads: rows past the segment are ignored, rows past the end of Bu fall into the
        // slack the workspace layout keeps behind it
        const float2* q = bu + (size_t)(sseg * sr) * P;
#pragma unroll
        for (int i = 0; i < 32; ++i, q += P) v[i] = *q;
      }
      float2 x = make_float2(0.f, 0.f), mm = lamS;
      // full 32-frame segment (every segment of every tile but a clip's last one): straight-line code.  With a predicate
      // per step the compiler wraps every step into its own branch region and re-derives the loop invariants inside it.
      const bool full_seg = len == 32;
      if (full_seg) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x = cfmaf(lam, x, v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < len) x = cfmaf(lam, x, v[i]);
      }
      if (len != sr) {                                        // short segment (end of a clip): lam^len
        mm = make_float2(1.f, 0.f);
        for (int i = 0; i < len; ++i) mm = cmulf(mm, lam);
      }
      seg_agg[sseg * P + sp] = x;
      seg_mul[sseg * P + sp] = mm;
      TC_TRACE_MARK(tid == 0);
      half_barrier(half);
      TC_TRACE_MARK(tid == 0);
      if (htid < P) {                                          // carry threads: segment 0 of every state (whole warps)
        float2 e = seg_agg[sp], M = seg_mul[sp];
        for (int q = 1; q < nseg; ++q) {
          e = cfmaf(seg_mul[q * P + sp], e, seg_agg[q * P + sp]);
          M = cmulf(M, seg_mul[q * P + sp]);
        }
        uint32_t* flag_row = pflags + (size_t)clip * p0.tiles_per_clip;
        float2* rec_row = precs + (size_t)clip * p0.tiles_per_clip * 2 * P;
        // Fixed-stride look-back.  Every tile publishes its aggregate (flag 1); every kStride-th tile (an "anchor") also
        // publishes its inclusive prefix (flag 2).  The carry-in of tile k is ALWAYS
        //     incl[a], then Horner over agg[a+1 .. k-1],   a = the last anchor before k,
        // so the floating-point association is fixed: results do not depend on timing, batching or sharding (a classic
        // decoupled look-back takes whatever prefix happens to be published and is not bit-reproducible).  All flags
        // and records of the walk (<= kStride of each) are requested together: two round trips, not 2 per predecessor.
        constexpr int kStride = S5_LOOKBACK_STRIDE;
        const bool last = chunk + 1 >= p0.tiles_per_clip;      // nobody looks back at the last tile of a clip
        const bool anchor = (chunk % kStride) == 0;
        float2 carry = make_float2(0.f, 0.f);
        if (!last) {
          __stcg(rec_row + ((size_t)chunk * 2 + (chunk == 0 ? 1 : 0)) * P + sp, e);   // tile 0: aggregate == inclusive prefix
          // records -> barrier among the carry threads -> release store of the flag by one of them: the release is
          // cumulative over the writes ordered before it by the barrier, no __threadfence needed
          asm volatile("bar.sync %0, %1;" ::"r"(3 + half), "r"(P) : "memory");
          if (htid == 0) st_release_u32(flag_row + chunk, chunk == 0 ? 2u : 1u);
        }
        if (chunk > 0) {
          const int a0 = ((chunk - 1) / kStride) * kStride;   // anchor
          const int n = chunk - a0;                           // records to read: incl[a0], agg[a0 + 1 .. chunk - 1]
          uint32_t f[kStride];
#pragma unroll
          for (int w = 0; w < kStride; ++w) f[w] = w < n ? ld_relaxed_u32(flag_row + a0 + w) : 2u;
          while (f[0] != 2u) f[0] = ld_relaxed_u32(flag_row + a0);
#pragma unroll
          for (int w = 1; w < kStride; ++w)
            while (f[w] == 0u) f[w] = ld_relaxed_u32(flag_row + a0 + w);
          fence_acquire_gpu();                                // one acquire for the whole walk
          float2 rec[kStride];
#pragma unroll
          for (int w = 0; w < kStride; ++w)
            if (w < n) rec[w] = __ldcg(rec_row + ((size_t)(a0 + w) * 2 + (w == 0 ? 1 : 0)) * P + sp);
          carry = rec[0];
#pragma unroll
          for (int w = 1; w < kStride; ++w)
            if (w < n) carry = cfmaf(lamR, carry, rec[w]);
          if (anchor && !last) {
            __stcg(rec_row + ((size_t)chunk * 2 + 1) * P + sp, cfmaf(M, carry, e));
            asm volatile("bar.sync %0, %1;" ::"r"(3 + half), "r"(P) : "memory");
            if (htid == 0) st_release_u32(flag_row + chunk, 2u);
          }
        }
        carry_s[sp] = carry;
      }
      TC_TRACE_MARK(tid == 0);
      half_barrier(half);
      TC_TRACE_MARK(tid == 0);
      float2 cin = carry_s[sp];
      for (int q = 0; q < sseg; ++q) cin = cfmaf(seg_mul[q * P + sp], cin, seg_agg[q * P + sp]);
      x = cin;
      auto emit_state = [&](float2 z, int i) {                // complex ReLU, bf16 hi / lo split, operand planes
        const bool keep = !relu_scan || (z.x > 0.f) || (z.x == 0.f && z.y > 0.f);
        const float v0 = keep ? z.x : 0.f, v1 = keep ? z.y * xs_odd_mul : 0.f;
        const uint32_t h01 = pack_bf2(v0, v1);
        const uint32_t l01 = pack_bf2(v0 - __uint_as_float(h01 << 16), v1 - __uint_as_float(h01 & 0xFFFF0000u));
        const uint32_t o = poff + (uint32_t)(sseg * sr + i) * 16u;
        *reinterpret_cast<uint32_t*>(P1h + o) = h01;
        *reinterpret_cast<uint32_t*>(P1l + o) = l01;
      };
      if (full_seg) {
        // the dependent chain first, the states back into v[]; the (independent) post-processing of the 32 states afterwards
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x = cfmaf(lam, x, v[i]);
          v[i] = x;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) emit_state(v[i], i);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < len) {
            x = cfmaf(lam, x, v[i]);
            emit_state(x, i);
          }
      }
    } else {
    // ---- S: xs -> two bf16 planes (value = hi + lo to 16 mantissa bits); 8 rows in flight per lane
    for (int rb = warp * 8; rb < rv; rb += NW * 8) {
      float4 raw[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        raw[q] = *reinterpret_cast<const float4*>(pXS + (size_t)min(row0 + rb + q, last_row) * K1 + xs_col);
      if (kin) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float v0 = raw[q].x, v1 = raw[q].y * xs_odd_mul, v2 = raw[q].z, v3 = raw[q].w * xs_odd_mul;
          const uint32_t h01 = pack_bf2(v0, v1), h23 = pack_bf2(v2, v3);
          const uint32_t l01 = pack_bf2(v0 - __uint_as_float(h01 << 16), v1 - __uint_as_float(h01 & 0xFFFF0000u));
          const uint32_t l23 = pack_bf2(v2 - __uint_as_float(h23 << 16), v3 - __uint_as_float(h23 & 0xFFFF0000u));
          const uint32_t o = soff + (uint32_t)(rb + q) * 16u;
          *reinterpret_cast<uint2*>(P1h + o) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(P1l + o) = make_uint2(l01, l23);
        }
      }
    }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    half_barrier(half);
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    if (htid == 0) {
      tc_issue(tcol, P1h, Wc, K1, H, false, true);
      tc_issue(tcol, P1l, Wc, K1, H, true, true);
      tc_commit(&bar[half]);
    }
    // the block-input rows of the first P1 batch travel while the tensor core works
    constexpr int PB = 8;                                     // rows per warp batch in the lanes-along-columns passes
    float xin[PB][HJ];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const float* xr = xtile + min(warp * PB + q, rlim) * H;
#pragma unroll
      for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
    }
    mbar_wait(&bar[half], phase);
    phase ^= 1;
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    // ---- E1: raw accumulators -> T (the xs planes are dead)
    if (erow0 < rv) {
      for (int c0 = cbeg; c0 < cbeg + cw; c0 += 8) {
        uint32_t v[8];
        tc_ld8(tcol + lane_base + c0, v);
        tc_wait_ld();
        float* dst = T + (size_t)erow * ldT + c0;
        *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(dst + 4) = make_uint4(v[4], v[5], v[6], v[7]);
      }
    }
    tc_fence_before();
    half_barrier(half);
    TC_TRACE_MARK(tid == 0);
    // ---- P1: one warp per row, PB rows in flight
    float bsub[HJ], bmulq[HJ], baddq[HJ], ddq[HJ];
#pragma unroll
    for (int j = 0; j < HJ; ++j) {
      bsub[j] = lcst[lane + 32 * j]; bmulq[j] = lcst[H + lane + 32 * j]; baddq[j] = lcst[2 * H + lane + 32 * j];
      ddq[j] = lcst[3 * H + lane + 32 * j];
    }
    for (int rb = warp * PB; rb < rv; rb += NW * PB) {
      float tin[PB][HJ];
      if (rb != warp * PB) {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const float* xr = xtile + min(rb + q, rlim) * H;
#pragma unroll
          for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
        }
      }
#pragma unroll
      for (int q = 0; q < PB; ++q)
#pragma unroll
        for (int j = 0; j < HJ; ++j) tin[q][j] = T[(size_t)(rb + q) * ldT + lane + 32 * j];
      __syncwarp();   // every lane holds its raw values of these rows before the rows are overwritten with codes
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int r = rb + q;
        uint32_t ob4[4] = {0u, 0u, 0u, 0u};
        unsigned short* trow = reinterpret_cast<unsigned short*>(T + (size_t)r * ldT);
#pragma unroll
        for (int j = 0; j < HJ; ++j) {
          const float uc = round_clip((xin[q][j] - bsub[j]) * bmulq[j] + baddq[j], qs_min, qs_max);   // code of u = FQ(BN(x))
          const float cy = round_clip(fmaf(ddq[j], uc, tin[q][j] * k_c), qrelu_min, qs_max);          // code of x1 = ReLU(FQ(y))
          trow[lane + 32 * j] = (unsigned short)__float_as_uint(fminf(cy * y_to_gl, qn_max) + kRoundMagic);   // left gate (>= 0)
          ob4[j] = __float_as_uint(fminf(cy * y_to_oa, qn_max) + kRoundMagic);                        // out2 input code (>= 0)
        }
        uint32_t lo, hi;
        split_planes(ob4[0], ob4[1], ob4[2], ob4[3], lo, hi);
        *reinterpret_cast<uint32_t*>(A2h + aoff + (uint32_t)r * 16u) = hi;
        *reinterpret_cast<uint32_t*>(A2l + aoff + (uint32_t)r * 16u) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    half_barrier(half);
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    if (htid == 0) {
      tc_issue8(tcol, A2h, Wo, kGluK2, H, H, true);
      tc_issue8(tcol + 128, A2l, Wo, kGluK2, H, H, false);
      tc_commit(&bar[half]);
    }
    // skip-connection rows of the first P2 batch (L2 hits) travel during the second product and E2
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const float* xr = xtile + min(warp * PB + q, rlim) * H;
#pragma unroll
      for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
    }
    mbar_wait(&bar[half], phase);
    phase ^= 1;
    tc_fence_after();
    TC_TRACE_MARK(tid == 0);
    // ---- E2: o = FQ(out2(x1)); right-gate code FQ(sigmoid(o)) -> second half of the T row
    if (erow0 < rv) {
      auto e2_8 = [&](const uint32_t (&vh)[8], const uint32_t (&vl)[8], int c0) {
        uint32_t sb[8];
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(ob + c0 + j);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f = __int2float_rn((int)vh[j + e] * 256 + (int)vl[j + e]);
            const float oc = round_clip(fmaf(f, a_o, bb[e]), qn_min, qn_max);                 // code of o = FQ(out2(x1))
            // sigmoid with the SFU exponential and reciprocal (~2 ulp), far below the 2^-15 grid of the FQ16 that follows
            float ex, sg;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(oc * k_e));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex));
            sb[j + e] = __float_as_uint(fminf(sg * gr_inv, qn_max) + kRoundMagic);           // right gate (> 0)
          }
        }
        uint4 pk;
        pk.x = __byte_perm(sb[0], sb[1], 0x5410); pk.y = __byte_perm(sb[2], sb[3], 0x5410);
        pk.z = __byte_perm(sb[4], sb[5], 0x5410); pk.w = __byte_perm(sb[6], sb[7], 0x5410);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(T + (size_t)erow * ldT) + 2 * H + 2 * c0) = pk;
      };
      // two register sets: the TMEM loads of the next 8 columns are in flight while these 8 are post-processed
      uint32_t ah[8], al[8], bh[8], bl[8];
      tc_ld8(tcol + lane_base + cbeg, ah);
      tc_ld8(tcol + lane_base + 128 + cbeg, al);
      for (int c0 = cbeg; c0 < cbeg + cw; c0 += 16) {         // cw = 16 HJ
        tc_wait_ld();
        tc_ld8(tcol + lane_base + c0 + 8, bh);
        tc_ld8(tcol + lane_base + 128 + c0 + 8, bl);
        e2_8(ah, al, c0);
        tc_wait_ld();
        if (c0 + 16 < cbeg + cw) {
          tc_ld8(tcol + lane_base + c0 + 16, ah);
          tc_ld8(tcol + lane_base + 128 + c0 + 16, al);
        }
        e2_8(bh, bl, c0 + 8);
      }
    }
    tc_fence_before();
    half_barrier(half);
    TC_TRACE_MARK(tid == 0);
    // ---- P2: out = ReLU(FQ_gl(x1) * FQ(sigmoid) + x), coalesced
    for (int rb = warp * PB; rb < rv; rb += NW * PB) {
      if (rb != warp * PB) {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const float* xr = xtile + min(rb + q, rlim) * H;
#pragma unroll
          for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
        }
      }
      const unsigned short* trow = reinterpret_cast<const unsigned short*>(T + (size_t)rb * ldT) + lane;
      float* yr = pOUT + (size_t)(row0 + rb) * H + lane;
      if (rb + PB <= rv) {
        // whole batch: four rows at a time, every code first, then the arithmetic, then the stores (a per-row exit test
        // turns every row into its own dependent load -> multiply -> store chain)
#pragma unroll
        for (int q0 = 0; q0 < PB; q0 += 4) {
          uint32_t lc[4][HJ], rc[4][HJ];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
              lc[q][j] = trow[(size_t)(q0 + q) * 2 * ldT + 32 * j];
              rc[q][j] = trow[(size_t)(q0 + q) * 2 * ldT + H + 32 * j];
            }
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
              // codes are non-negative: int -> float through the magic constant (no conversion pipe)
              const float l = (__uint_as_float(0x4B400000u + lc[q][j]) - kRoundMagic) * gl_s;
              const float sg = (__uint_as_float(0x4B400000u + rc[q][j]) - kRoundMagic) * gr_s;
              yr[(size_t)(q0 + q) * H + 32 * j] = fmaxf(fmaf(l, sg, xin[q0 + q][j]), out_floor);
            }
        }
      } else {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          if (rb + q >= rv) break;                              // warp-uniform
#pragma unroll
          for (int j = 0; j < HJ; ++j) {
            const float l = (__uint_as_float(0x4B400000u + trow[(size_t)q * 2 * ldT + 32 * j]) - kRoundMagic) * gl_s;
            const float sg = (__uint_as_float(0x4B400000u + trow[(size_t)q * 2 * ldT + H + 32 * j]) - kRoundMagic) * gr_s;
            const float out = fmaf(l, sg, xin[q][j]);
            yr[(size_t)q * H + 32 * j] = fmaxf(out, out_floor);
          }
        }
      }
    }
    half_barrier(half);   // T is free: the next tile's planes may overwrite it; every row of the tile has been stored
    if constexpr (SCAN) {
      // release (cumulative over the stores the barrier ordered before it): the same tile one layer up may start
      if (layer + 1 < mp.nlayers && htid == 0) st_release_u32(mp.done + (size_t)layer * ntiles + it, 1u);
    }
    TC_TRACE_MARK(tid == 0);
    TC_TRACE_NEXT();
  }
  TC_TRACE_PRINT(tid == 0, BPJ ? "block: S0+sync mmaB EB+sync load+pass1 bar lookback bar pass2+sync mma1 E1+sync P1+sync mma2 E2+sync P2+sync" : SCAN ? "glu+scan: load+pass1 bar lookback bar pass2+sync mma1 E1+sync P1+sync mma2 E2+sync P2+sync" : "glu: S+sync mma1 E1+sync P1+sync mma2 E2+sync P2+sync");
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- wide C projection + GLU
// The same block tail for models whose weight tiles do not fit next to two tile pipelines (dim_scale 1.0: H = 192,
// 2P = 256: C alone is 99 KB in bf16, one tile's state planes 132 KB).  One 512-thread pipeline per SM; the C
// projection is K-CHUNKED: the scan output arrives in chunks of 64 real columns (32 states), each chunk's two bf16
// planes and the matching [H][64] slice of C (streamed from L2 by ONE bulk copy, pre-arranged in the operand layout by
// the host) sit in one of two buffers, and the tensor core accumulates chunk after chunk into the same TMEM columns
// while the warps stage the next chunk.  Afterwards the buffers are dead and hold the fp32 tile T of the passes P1 / E2 /
// P2, which are those of tc_glu_kernel.  out2's K order: column lane + 32 j sits at byte 4 lane + j of the first
// 128-byte block (j < 4) and at byte 128 + w lane + (j - 4) of the second (w = 1, 2 or 4 bytes per lane).
constexpr int kGluwThreads = 512;
constexpr int kGluwKC = 64;        // real columns of the scan output per chunk

__host__ __device__ constexpr int gluw_w1(int HJ) { return HJ <= 4 ? 0 : HJ == 5 ? 1 : HJ == 6 ? 2 : 4; }   // bytes per lane, second block
__host__ __device__ constexpr int gluw_k2(int HJ) { return 128 + 32 * gluw_w1(HJ); }
// physical byte of out2 input column k
static inline int gluw_kpos(int k, int HJ) {
  const int lane = k & 31, j = k >> 5;
  return j < 4 ? 4 * lane + j : 128 + gluw_w1(HJ) * lane + (j - 4);
}

template <int HJ>
__global__ void __launch_bounds__(kGluwThreads, 1) tc_gluw_kernel(const __grid_constant__ TcGluParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int H = 32 * HJ, K2 = gluw_k2(HJ), W1 = gluw_w1(HJ);
  constexpr int PB = HJ > 4 ? 4 : 8;                          // rows per warp batch in the lanes-along-columns passes
  const int K1 = p.K1, NC = K1 / kGluwKC;
  constexpr int ldT = H + 4;
  const size_t wo_bytes = tc_tile_bytes8(H, K2);
  const size_t pl_bytes = tc_tile_bytes(kTcRows, kGluwKC);    // one bf16 plane of one chunk
  const size_t wck_bytes = tc_tile_bytes(H, kGluwKC);         // one chunk of C
  const size_t buf_bytes = 2 * pl_bytes + wck_bytes;
  const size_t t_bytes = (size_t)kTcRows * ldT * 4;
  const size_t a2_bytes = tc_tile_bytes8(kTcRows, K2);
  unsigned char* Wo = smem;
  float* ob = reinterpret_cast<float*>(Wo + wo_bytes);        // [H] out2 bias / s_o_out
  float* lcst = ob + H;                                       // [4][H]
  unsigned char* R = reinterpret_cast<unsigned char*>(ob) + (((size_t)(5 * H) * 4 + 127) & ~(size_t)127);
  float* T = reinterpret_cast<float*>(R);                     // aliases the two chunk buffers
  unsigned char* A2h = R + max(t_bytes, 2 * buf_bytes);
  unsigned char* A2l = A2h + a2_bytes;
  __shared__ __align__(8) uint64_t full[2], mdone[2], bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = kGluwThreads / 32;
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&mdone[0], 1); mbar_init(&mdone[1], 1); mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_load_weights8(Wo, p.Wo, H, K2);
  for (int n = tid; n < H; n += kGluwThreads) {
    ob[n] = p.o_bias[n] * p.oo_inv;
    lcst[n] = p.bn_sub ? p.bn_sub[n] : 0.f;
    lcst[H + n] = (p.bn_sub ? p.bn_mul[n] : 1.f) * p.ut_inv;
    lcst[2 * H + n] = (p.bn_sub ? p.bn_add[n] : 0.f) * p.ut_inv;
    lcst[3 * H + n] = p.D[n] * p.ut_s * p.yt_inv;
  }
  for (size_t i = (size_t)tid * 16; i < 2 * a2_bytes; i += kGluwThreads * 16) *reinterpret_cast<int4*>(A2h + i) = make_int4(0, 0, 0, 0);
  const long long last_row = p.rows - 1;
  const uint32_t lbo = tc_lbo(kTcRows);
  // chunk staging: 16 lanes per row (float4 = 4 real columns each), 32 rows per pass, 4 passes per tile
  const int srow = tid >> 4, l16 = tid & 15;
  const uint32_t soff = (uint32_t)(l16 >> 1) * lbo + (uint32_t)(l16 & 1) * 8u;
  const uint32_t aoff0 = (uint32_t)(lane >> 2) * lbo + 4u * (lane & 3);
  const uint32_t aoff1 = W1 == 4 ? (uint32_t)(8 + (lane >> 2)) * lbo + 4u * (lane & 3)
                       : W1 == 2 ? (uint32_t)(8 + (lane >> 3)) * lbo + 2u * (lane & 7)
                                 : (uint32_t)(8 + (lane >> 4)) * lbo + (uint32_t)(lane & 15);
  const int erow0 = (warp & 3) * 32, erow = erow0 + lane;
  const uint32_t lane_base = (uint32_t)erow0 << 16;
  constexpr int cw = H >> 2;                                  // 4 column groups in the TMEM passes (multiple of 8)
  const int cbeg = (warp >> 2) * cw;
  const float qs_min = p.qs_min, qs_max = p.qs_max, qn_min = p.qn_min, qn_max = p.qn_max;
  const float qrelu_min = fmaxf(qs_min, 0.f);
  const float k_c = p.c_scale * p.yt_inv, a_o = p.o_col_scale * p.oo_inv, k_e = -p.oo_s * 1.4426950408889634f;
  const float y_to_gl = p.yt_s * p.gl_inv, y_to_oa = p.yt_s * p.oa_inv, xs_odd_mul = p.xs_odd_mul;
  const float gl_s = p.gl_s, gr_s = p.gr_s, gr_inv = p.gr_inv, out_floor = p.relu_out ? 0.f : -INFINITY;
  const unsigned char* WcG = reinterpret_cast<const unsigned char*>(p.Wc);
  uint32_t ph_full[2] = {0, 0}, ph_done[2] = {0, 0}, phase = 0;
  __syncthreads();
  const uint32_t tcol = tmem_slot;
  const long long row_begin = p.rows * blockIdx.x / gridDim.x, row_end = p.rows * (blockIdx.x + 1) / gridDim.x;
  float4 raw[4];
  auto load_chunk = [&](long long row0, int c) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      raw[q] = *reinterpret_cast<const float4*>(p.XS + (size_t)min(row0 + 32 * q + srow, last_row) * K1 + c * kGluwKC + 4 * l16);
  };
  if (row_begin < row_end) load_chunk(row_begin, 0);
  for (long long row0 = row_begin; row0 < row_end; row0 += kTcRows) {
    const int rv = (int)min((long long)kTcRows, row_end - row0);
    const float* xtile = p.X + (size_t)row0 * H;
    const int rlim = (int)min((long long)kTcRows, last_row - row0);
    // ---- S + MMA 1, chunk by chunk
    for (int c = 0; c < NC; ++c) {
      const int b = c & 1;
      unsigned char* buf = R + (size_t)b * buf_bytes;
      if (c >= 2) {                                           // the MMAs of chunk c - 2 have read this buffer
        mbar_wait(&mdone[b], ph_done[b]);
        ph_done[b] ^= 1;
      }
      if (tid == 0) {
        fence_proxy_async_smem();                             // earlier generic accesses (T of the previous tile) before the bulk write
        mbar_expect_tx(&full[b], (uint32_t)wck_bytes);
        bulk_g2s(buf + 2 * pl_bytes, WcG + (size_t)c * wck_bytes, (uint32_t)wck_bytes, &full[b]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float v0 = raw[q].x, v1 = raw[q].y * xs_odd_mul, v2 = raw[q].z, v3 = raw[q].w * xs_odd_mul;
        const uint32_t h01 = pack_bf2(v0, v1), h23 = pack_bf2(v2, v3);
        const uint32_t l01 = pack_bf2(v0 - __uint_as_float(h01 << 16), v1 - __uint_as_float(h01 & 0xFFFF0000u));
        const uint32_t l23 = pack_bf2(v2 - __uint_as_float(h23 << 16), v3 - __uint_as_float(h23 & 0xFFFF0000u));
        const uint32_t o = soff + (uint32_t)(32 * q + srow) * 16u;
        *reinterpret_cast<uint2*>(buf + o) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(buf + pl_bytes + o) = make_uint2(l01, l23);
      }
      // the next chunk (or the first chunk of the next tile) travels while this one is multiplied
      if (c + 1 < NC) load_chunk(row0, c + 1);
      else if (row0 + kTcRows < row_end) load_chunk(row0 + kTcRows, 0);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        mbar_wait(&full[b], ph_full[b]);                      // this chunk of C has landed
        tc_issue(tcol, buf, buf + 2 * pl_bytes, kGluwKC, H, c > 0, true);
        tc_issue(tcol, buf + pl_bytes, buf + 2 * pl_bytes, kGluwKC, H, true, true);
        tc_commit(&mdone[b]);
      }
      ph_full[b] ^= 1;
    }
    // the block-input rows of the first P1 batch travel while the tensor core finishes
    float xin[PB][HJ];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const float* xr = xtile + min(warp * PB + q, rlim) * H;
#pragma unroll
      for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
    }
    for (int c = max(0, NC - 2); c < NC; ++c) {               // the last commit on each buffer
      mbar_wait(&mdone[c & 1], ph_done[c & 1]);
      ph_done[c & 1] ^= 1;
    }
    tc_fence_after();
    // ---- E1: raw accumulators -> T (the chunk buffers are dead)
    if (erow0 < rv) {
      for (int c0 = cbeg; c0 < cbeg + cw; c0 += 8) {
        uint32_t v[8];
        tc_ld8(tcol + lane_base + c0, v);
        tc_wait_ld();
        float* dst = T + (size_t)erow * ldT + c0;
        *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(dst + 4) = make_uint4(v[4], v[5], v[6], v[7]);
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- P1: one warp per row, PB rows in flight
    float bsub[HJ], bmulq[HJ], baddq[HJ], ddq[HJ];
#pragma unroll
    for (int j = 0; j < HJ; ++j) {
      bsub[j] = lcst[lane + 32 * j]; bmulq[j] = lcst[H + lane + 32 * j]; baddq[j] = lcst[2 * H + lane + 32 * j];
      ddq[j] = lcst[3 * H + lane + 32 * j];
    }
    for (int rb = warp * PB; rb < rv; rb += NW * PB) {
      float tin[PB][HJ];
      if (rb != warp * PB) {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const float* xr = xtile + min(rb + q, rlim) * H;
#pragma unroll
          for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
        }
      }
#pragma unroll
      for (int q = 0; q < PB; ++q)
#pragma unroll
        for (int j = 0; j < HJ; ++j) tin[q][j] = T[(size_t)(rb + q) * ldT + lane + 32 * j];
      __syncwarp();   // every lane holds its raw values of these rows before the rows are overwritten with codes
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int r = rb + q;
        uint32_t ob8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        unsigned short* trow = reinterpret_cast<unsigned short*>(T + (size_t)r * ldT);
#pragma unroll
        for (int j = 0; j < HJ; ++j) {
          const float uc = round_clip((xin[q][j] - bsub[j]) * bmulq[j] + baddq[j], qs_min, qs_max);   // code of u = FQ(BN(x))
          const float cy = round_clip(fmaf(ddq[j], uc, tin[q][j] * k_c), qrelu_min, qs_max);          // code of x1 = ReLU(FQ(y))
          trow[lane + 32 * j] = (unsigned short)__float_as_uint(fminf(cy * y_to_gl, qn_max) + kRoundMagic);   // left gate (>= 0)
          ob8[j] = __float_as_uint(fminf(cy * y_to_oa, qn_max) + kRoundMagic);                        // out2 input code (>= 0)
        }
        uint32_t lo, hi;
        split_planes(ob8[0], ob8[1], ob8[2], ob8[3], lo, hi);
        *reinterpret_cast<uint32_t*>(A2h + aoff0 + (uint32_t)r * 16u) = hi;
        *reinterpret_cast<uint32_t*>(A2l + aoff0 + (uint32_t)r * 16u) = lo;
        if constexpr (W1 == 4) {
          split_planes(ob8[4], ob8[5], ob8[6], ob8[7], lo, hi);
          *reinterpret_cast<uint32_t*>(A2h + aoff1 + (uint32_t)r * 16u) = hi;
          *reinterpret_cast<uint32_t*>(A2l + aoff1 + (uint32_t)r * 16u) = lo;
        } else if constexpr (W1 == 2) {
          *reinterpret_cast<unsigned short*>(A2h + aoff1 + (uint32_t)r * 16u) = (unsigned short)__byte_perm(ob8[4], ob8[5], 0x0051);
          *reinterpret_cast<unsigned short*>(A2l + aoff1 + (uint32_t)r * 16u) = (unsigned short)__byte_perm(ob8[4], ob8[5], 0x0040);
        } else if constexpr (W1 == 1) {
          A2h[aoff1 + (uint32_t)r * 16u] = (unsigned char)(ob8[4] >> 8);
          A2l[aoff1 + (uint32_t)r * 16u] = (unsigned char)ob8[4];
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      tc_issue8(tcol, A2h, Wo, K2, H, H, true);
      tc_issue8(tcol + 256, A2l, Wo, K2, H, H, false);
      tc_commit(&bar);
    }
    // skip-connection rows of the first P2 batch (L2 hits) travel during the second product and E2
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const float* xr = xtile + min(warp * PB + q, rlim) * H;
#pragma unroll
      for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- E2: o = FQ(out2(x1)); right-gate code FQ(sigmoid(o)) -> second half of the T row
    if (erow0 < rv) {
      auto e2_8 = [&](const uint32_t (&vh)[8], const uint32_t (&vl)[8], int c0) {
        uint32_t sb[8];
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(ob + c0 + j);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f = __int2float_rn((int)vh[j + e] * 256 + (int)vl[j + e]);
            const float oc = round_clip(fmaf(f, a_o, bb[e]), qn_min, qn_max);                 // code of o = FQ(out2(x1))
            float ex, sg;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(oc * k_e));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex));
            sb[j + e] = __float_as_uint(fminf(sg * gr_inv, qn_max) + kRoundMagic);           // right gate (> 0)
          }
        }
        uint4 pk;
        pk.x = __byte_perm(sb[0], sb[1], 0x5410); pk.y = __byte_perm(sb[2], sb[3], 0x5410);
        pk.z = __byte_perm(sb[4], sb[5], 0x5410); pk.w = __byte_perm(sb[6], sb[7], 0x5410);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(T + (size_t)erow * ldT) + 2 * H + 2 * c0) = pk;
      };
      // two register sets: the TMEM loads of the next 8 columns are in flight while these 8 are post-processed
      uint32_t ah[8], al[8], bh[8], bl[8];
      tc_ld8(tcol + lane_base + cbeg, ah);
      tc_ld8(tcol + lane_base + 256 + cbeg, al);
#pragma unroll
      for (int c0 = 0; c0 < cw; c0 += 16) {                   // cw = 8 HJ: a multiple of 16 for even HJ, else one odd group of 8
        tc_wait_ld();
        if (c0 + 8 < cw) {
          tc_ld8(tcol + lane_base + cbeg + c0 + 8, bh);
          tc_ld8(tcol + lane_base + 256 + cbeg + c0 + 8, bl);
        }
        e2_8(ah, al, cbeg + c0);
        if (c0 + 8 < cw) {
          tc_wait_ld();
          if (c0 + 16 < cw) {
            tc_ld8(tcol + lane_base + cbeg + c0 + 16, ah);
            tc_ld8(tcol + lane_base + 256 + cbeg + c0 + 16, al);
          }
          e2_8(bh, bl, cbeg + c0 + 8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- P2: out = ReLU(FQ_gl(x1) * FQ(sigmoid) + x), coalesced
    for (int rb = warp * PB; rb < rv; rb += NW * PB) {
      if (rb != warp * PB) {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          const float* xr = xtile + min(rb + q, rlim) * H;
#pragma unroll
          for (int j = 0; j < HJ; ++j) xin[q][j] = xr[lane + 32 * j];
        }
      }
      const unsigned short* trow = reinterpret_cast<const unsigned short*>(T + (size_t)rb * ldT) + lane;
      float* yr = p.OUT + (size_t)(row0 + rb) * H + lane;
      if (rb + PB <= rv) {
        // whole batch: two rows at a time, every code first, then the arithmetic, then the stores (a per-row exit test turns
        // every row into its own dependent load -> multiply -> store chain)
#pragma unroll
        for (int q0 = 0; q0 < PB; q0 += 2) {
          uint32_t lc[2][HJ], rc[2][HJ];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
              lc[q][j] = trow[(size_t)(q0 + q) * 2 * ldT + 32 * j];
              rc[q][j] = trow[(size_t)(q0 + q) * 2 * ldT + H + 32 * j];
            }
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < HJ; ++j) {
              const float l = (__uint_as_float(0x4B400000u + lc[q][j]) - kRoundMagic) * gl_s;
              const float sg = (__uint_as_float(0x4B400000u + rc[q][j]) - kRoundMagic) * gr_s;
              yr[(size_t)(q0 + q) * H + 32 * j] = fmaxf(fmaf(l, sg, xin[q0 + q][j]), out_floor);
            }
        }
      } else {
#pragma unroll
        for (int q = 0; q < PB; ++q) {
          if (rb + q >= rv) break;                              // warp-uniform
#pragma unroll
          for (int j = 0; j < HJ; ++j) {
            const float l = (__uint_as_float(0x4B400000u + trow[(size_t)q * 2 * ldT + 32 * j]) - kRoundMagic) * gl_s;
            const float sg = (__uint_as_float(0x4B400000u + trow[(size_t)q * 2 * ldT + H + 32 * j]) - kRoundMagic) * gr_s;
            yr[(size_t)q * H + 32 * j] = fmaxf(fmaf(l, sg, xin[q][j]), out_floor);
          }
        }
      }
    }
    __syncthreads();                                          // T is free: the next tile's chunks may overwrite it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- host side

static int upload_f(const std::vector<float>& h, float** d) {
  S5_CUDA(cudaMalloc(d, std::max<size_t>(1, h.size()) * sizeof(float)));
  S5_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return S5_OK;
}

// resident CTAs per SM from the three budgets that matter here: shared memory, registers, TMEM columns
static int tc_ctas_per_sm(const void* kernel, size_t dyn_smem, int tmem_ctas) {
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 1;
  const int by_smem = (int)((227 * 1024) / (dyn_smem + fa.sharedSizeBytes + 1024));
  const int by_regs = 65536 / std::max(1, fa.numRegs * kTcThreads);
  return std::max(1, std::min(std::min(by_smem, by_regs), tmem_ctas));
}

bool is_pow2(float s) {
  int e;
  return s > 0.f && std::frexp(s, &e) == 0.5f;
}

// Largest power of two g such that every value of `w` (stride-strided subset) is an integer multiple of g.
float pow2_grid(const float* w, size_t n, size_t start, size_t stride) {
  int best = 1 << 30;
  for (size_t i = start; i < n; i += stride) {
    const float v = w[i];
    if (v == 0.f || !std::isfinite(v)) continue;
    int e;
    const float m = std::frexp(std::fabs(v), &e);              // v = m * 2^e, m in [0.5, 1)
    const uint32_t mi = (uint32_t)std::ldexp(m, 24);           // 24-bit integer mantissa
    const int low = e - 24 + __builtin_ctz(mi);                // exponent of the lowest set bit
    best = std::min(best, low);
  }
  return best == (1 << 30) ? 1.f : std::ldexp(1.f, best);
}

// W: fake-quantised weights [Kfull][Nfull]; even / odd output columns may sit on different power-of-two grids
// (real / imaginary parts of the B projection).  s_in: activation scale.  out->ok stays false when the matrix is
// not covered (shape, or codes outside int8), in which case the compressed kernels are used.
int tc_linear_build(const float* W, int Kfull, int Nfull, float s_in, TcLinearDev* out) {
  out->ok = false;
  if (!is_pow2(s_in)) return S5_OK;
  const bool tail = Kfull > 128 && (Kfull % 128) == 1;
  const int KV = tail ? Kfull / 128 : (Kfull + 127) / 128;
  if (KV < 1 || KV > 2 || Kfull < 16 || Nfull < 16 || Nfull > 512) return S5_OK;   // shape not covered
  // padded K of the operand tiles: natural order / lane-interleaved order (whole 128-column blocks)
  const int Kp[2] = {(Kfull + 31) & ~31, 128 * KV + (tail ? 32 : 0)};
  // one output column beyond a multiple of 16 (decoder: 257) is computed on the CUDA cores instead of a whole MMA pass
  const bool xn = (Nfull % 16) == 1 && Nfull > 16 && !tail;
  const int Np = xn ? Nfull - 1 : (Nfull + 15) & ~15;
  std::vector<float> col0, col1;
  for (int k = 0; k < Kfull; ++k)
    for (int n = 0; n < Nfull; ++n) ((n & 1) ? col1 : col0).push_back(W[(size_t)k * Nfull + n]);
  const float w_scale0 = pow2_grid(col0.data(), col0.size(), 0, 1);
  const float w_scale1 = pow2_grid(col1.data(), col1.size(), 0, 1);
  // physical position of logical column k: identity for 16-byte aligned rows; lane-interleaved inside each
  // 128-column block for the 4-byte staging path (lane owns k = lane + 32 e  <->  bytes 4 lane + e)
  auto perm = [&](int k) { return k >= 128 * KV ? k : (k & ~127) + 4 * (k & 31) + ((k & 127) >> 5); };
  std::vector<signed char> w8[2], wnx[2];
  for (int o = 0; o < 2; ++o) {
    w8[o].assign((size_t)Np * Kp[o], 0);
    wnx[o].assign((size_t)Kp[o] + 256, 0);                    // + slack: every lane reads its four codes
  }
  std::vector<float> cs(Nfull);
  for (int n = 0; n < Nfull; ++n) {
    const float sw = (n & 1) ? w_scale1 : w_scale0;
    cs[n] = s_in * sw;
    for (int k = 0; k < Kfull; ++k) {
      const float c = std::nearbyint(W[(size_t)k * Nfull + n] / sw);
      if (c < -128.f || c > 127.f) return S5_OK;  // not an int8 code: leave this matrix to the compressed kernels
      if (n < Np) {
        w8[0][(size_t)n * Kp[0] + k] = (signed char)c;
        w8[1][(size_t)n * Kp[1] + perm(k)] = (signed char)c;
      } else {
        wnx[0][k] = (signed char)c;
        wnx[1][perm(k)] = (signed char)c;
      }
    }
  }
  out->Kfull = Kfull; out->Nfull = Nfull; out->Kp[0] = Kp[0]; out->Kp[1] = Kp[1]; out->Np = Np; out->KV = KV;
  out->xn = xn;
  for (int o = 0; o < 2; ++o) {
    S5_CUDA(cudaMalloc(&out->W8[o], w8[o].size()));
    S5_CUDA(cudaMemcpy(out->W8[o], w8[o].data(), w8[o].size(), cudaMemcpyHostToDevice));
    S5_CUDA(cudaMalloc(&out->WN[o], wnx[o].size()));
    S5_CUDA(cudaMemcpy(out->WN[o], wnx[o].data(), wnx[o].size(), cudaMemcpyHostToDevice));
  }
  int rc;
  if ((rc = upload_f(cs, &out->col_scale))) return rc;
  out->ok = true;
  return S5_OK;
}

void tc_linear_free(TcLinearDev* d) {
  cudaFree(d->W8[0]); cudaFree(d->W8[1]); cudaFree(d->WN[0]); cudaFree(d->WN[1]); cudaFree(d->col_scale);
  *d = TcLinearDev();
}

template <bool VEC4, int KV, bool BN, bool XN>
static int tc_linear_go(const TcLinParams& p0, int kclass, cudaStream_t st) {
  TcLinParams p = p0;
  auto kernel = tc_linear_kernel<VEC4, KV, BN, XN>;
  const size_t w_bytes = tc_tile_bytes8(p.Np, p.Kp), a2 = 2 * tc_tile_bytes8(kTcRows, p.Kp);
  // output columns per pass: as wide as possible while two CTAs (2 x 8 warps, 2 x 256 TMEM columns) share an SM
  size_t smem = 0;
  auto fits2 = [&](size_t b) { return 2 * (b + 1024 + 64) <= (size_t)227 * 1024; };
  auto smem_for = [&](int NP) {
    p.NP = std::min(NP, p.Np);
    p.alias_out = p.NP >= p.Np;                               // single pass: the output tile reuses the dead planes
    const size_t out_bytes = (size_t)kTcRows * (p.NP + 4) * 4;
    return w_bytes + (p.alias_out ? std::max(a2, out_bytes) : a2 + out_bytes) + (size_t)p.Np * sizeof(float2);
  };
  // S5B200_LIN_MIN_NP (default 64): narrower passes than this are not worth a second CTA per SM (a pass is an MMA, an
  // epilogue, two barriers and a write-out of NP-float row segments) - one CTA per SM with 128-column passes instead
  static const int min_np = [] { const char* e = getenv("S5B200_LIN_MIN_NP"); return e ? atoi(e) : 64; }();
  bool two = false;
  for (int NP : {128, 64, 32, 16}) {
    smem = smem_for(NP);
    if (fits2(smem) && (NP >= min_np || NP >= p.Np)) { two = true; break; }
    if (NP < min_np) break;
  }
  if (!two)
    for (int NP : {128, 64, 32, 16}) {
      smem = smem_for(NP);
      if (smem + 1024 + 64 <= (size_t)227 * 1024) break;
    }
  if (smem + 64 > 227 * 1024) S5_FAIL(S5_E_BADSHAPE, "dense-tile linear: tile does not fit in shared memory");
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int per_sm = 2 * (smem + 1024 + 64) <= 227 * 1024 ? 2 : 1;
  // rows are split evenly over the CTAs (balanced partition); at least 32 rows per CTA
  const int grid = (int)std::max<long long>(1, std::min<long long>((p.rows + 31) / 32, (long long)per_sm * sm_count()));
  ProfScope prof(kclass, st);
  kernel<<<grid, kTcThreads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

int tc_linear_launch(const TcLinearDev& d, const float* X, float* Y, long long rows, const float* in_sub, const float* in_mul,
                     const float* in_add, float in_scale, float in_qmin, float in_qmax, const float* bias, float out_scale0,
                     float out_scale1, float out_qmin, float out_qmax, int relu, int kclass, cudaStream_t st) {
  if (!d.ok || !is_pow2(in_scale) || !is_pow2(out_scale0) || !is_pow2(out_scale1)) S5_FAIL(S5_E_BADARG, "dense-tile linear: not eligible");
  TcLinParams p{};
  p.X = X; p.Y = Y; p.rows = rows; p.Kfull = d.Kfull; p.Nfull = d.Nfull; p.Np = d.Np;
  p.x_stride = d.Kfull; p.y_stride = d.Nfull;
  p.in_sub = in_sub; p.in_mul = in_mul; p.in_add = in_add; p.in_inv = 1.0f / in_scale; p.in_qmin = in_qmin; p.in_qmax = in_qmax;
  p.col_scale = d.col_scale; p.bias = bias; p.out_scale0 = out_scale0; p.out_scale1 = out_scale1;
  p.out_inv0 = 1.0f / out_scale0; p.out_inv1 = 1.0f / out_scale1;
  p.out_qmin = out_qmin; p.out_qmax = out_qmax; p.relu = relu;
  const bool vec4 = (d.Kfull & 3) == 0 && aligned16(X);
  p.W8 = d.W8[vec4 ? 0 : 1];
  p.WN = d.WN[vec4 ? 0 : 1];
  p.Kp = d.Kp[vec4 ? 0 : 1];
  const bool bn = in_sub != nullptr;
  // instantiated combinations: BatchNorm only occurs on the B projection (no extra column), the extra column only
  // on 16-byte aligned inputs (decoder)
  if (d.xn) {
    if (!vec4 || bn) S5_FAIL(S5_E_BADARG, "dense-tile linear: extra output column needs aligned rows and no BatchNorm");
    // wide decoder: contiguous output rows, the whole N in one MMA pass, one bulk copy per tile
    static const bool dec_tma = [] { const char* e = getenv("S5B200_DEC_TMA"); return !e || atoi(e) != 0; }();
    const size_t dsm = tc_tile_bytes8(d.Np, p.Kp) + 2 * tc_tile_bytes8(kTcRows, p.Kp) + (size_t)kTcRows * d.Nfull * 4 + (size_t)d.Nfull * sizeof(float2);
    // warp-specialised version (S5B200_DEC_V2=0 selects the first one)
    static const bool dec_v2 = [] { const char* e = getenv("S5B200_DEC_V2"); return !e || atoi(e) != 0; }();
    if (dec_v2 && d.Np <= 256 && (d.Np & 31) == 0 && (p.Kp & 31) == 0 && aligned16(Y) && d.Nfull == d.Np + 1) {
      // the deepest input ring / most output half-tiles that fit: (3 out, >= 3 ring), (3, 2), (2, 2)
      const size_t fixed = 2 * tc_tile_bytes8(kTcRows, p.Kp) + 2 * tc_tile_bytes8(2 * kDec2Rows, p.Kp);
      const size_t slot = (size_t)kDec2Sub * d.Kfull * 4, ohalf = (size_t)kDec2Sub * d.Nfull * 4;
      // instantiated shapes: (ring 6, 3 output half-tiles) where it fits (K <= 128), else (2, 2)
      int nring = 0, nout = 0;
      if (fixed + 6 * slot + 3 * ohalf + 2048 <= 227 * 1024) { nring = 6; nout = 3; }
      else if (fixed + 2 * slot + 2 * ohalf + 2048 <= 227 * 1024) { nring = 2; nout = 2; }
      if (nring) {
        const size_t dsm2 = fixed + nring * slot + nout * ohalf;
        auto kernel = nring == 6 ? (d.KV == 1 ? tc_dec2_kernel<1, 6, 3, false, false> : tc_dec2_kernel<2, 6, 3, false, false>)
                                 : (d.KV == 1 ? tc_dec2_kernel<1, 2, 2, false, false> : tc_dec2_kernel<2, 2, 2, false, false>);
        S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm2));
        const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
        ProfScope prof(kclass, st);
        kernel<<<grid, kDec2Threads, dsm2, st>>>(p);
        S5_LAUNCH_CHECK();
        return S5_OK;
      }
    }
    if (dec_tma && d.KV == 1 && d.Np <= 256 && (d.Np & 31) == 0 && aligned16(Y) && dsm + 2048 <= 227 * 1024) {
      S5_CUDA(cudaFuncSetAttribute(tc_dec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
      const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
      ProfScope prof(kclass, st);
      tc_dec_kernel<<<grid, kDecThreads, dsm, st>>>(p);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
    return d.KV == 1 ? tc_linear_go<true, 1, false, true>(p, kclass, st) : tc_linear_go<true, 2, false, true>(p, kclass, st);
  }
  // unaligned contiguous rows (encoder): TMA-fed raw-row ring, one CTA per SM
  static const bool enc_tma = [] { const char* e = getenv("S5B200_ENC_TMA"); return !e || atoi(e) != 0; }();
  if (enc_tma && !vec4 && !bn && aligned16(X) && aligned16(Y) && d.Np <= 256 && (d.Np & 31) == 0 && (d.Nfull & 3) == 0 && d.Nfull <= d.Np) {
    const size_t raw_bytes = (((size_t)kEncSub * d.Kfull * 4) + 127) & ~(size_t)127;
    const size_t esm = tc_tile_bytes8(d.Np, p.Kp) + 2 * tc_tile_bytes8(kTcRows, p.Kp) + kEncRing * raw_bytes + (size_t)d.Np * sizeof(float2);
    // warp-specialised version (S5B200_ENC_V2=0 selects the first one)
    static const bool enc_v2 = [] { const char* e = getenv("S5B200_ENC_V2"); return !e || atoi(e) != 0; }();
    const size_t raw2 = (((size_t)kEnc2Sub * d.Kfull * 4) + 127) & ~(size_t)127;
    const int nblk2 = (d.Np + 127) >> 7;
    int nring = kEnc2Ring;
    size_t esm2 = 0;
    for (; nring >= 2; --nring) {                              // the deepest ring that fits
      esm2 = (size_t)nblk2 * tc_tile_bytes8(kTcRows, p.Kp) + 2 * tc_tile_bytes8(2 * kEnc2Rows, p.Kp) + nring * raw2;
      if (esm2 + 2048 <= 227 * 1024) break;
    }
    if (enc_v2 && nring >= 2) {
      auto kernel = d.KV == 1 ? (nring == 3 ? tc_enc2_kernel<1, 3> : tc_enc2_kernel<1, 2>) : (nring == 3 ? tc_enc2_kernel<2, 3> : tc_enc2_kernel<2, 2>);
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm2));
      const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
      ProfScope prof(kclass, st);
      kernel<<<grid, kEnc2Threads, esm2, st>>>(p);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
    if (d.Np <= 128 && esm + 2048 <= 227 * 1024 && 2 * tc_tile_bytes8(kTcRows, p.Kp) >= (size_t)kTcRows * (d.Np + 4) * 4) {
      auto kernel = d.KV == 1 ? tc_enc_kernel<1> : tc_enc_kernel<2>;
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));
      const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
      ProfScope prof(kclass, st);
      kernel<<<grid, kEncThreads, esm, st>>>(p);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
  }
  // aligned input rows, line-aligned output rows (B projection; any Dense with K % 4 == 0 and N % 32 == 0): the warp-specialised
  // kernel with direct register stores (S5B200_LIN_V2=0 selects the first version)
  static const bool lin_v2 = [] { const char* e = getenv("S5B200_LIN_V2"); return !e || atoi(e) != 0; }();
  if (lin_v2 && vec4 && aligned16(Y) && d.Np <= 256 && (d.Np & 31) == 0 && d.Nfull == d.Np && (p.Kp & 31) == 0 && (!bn || in_mul)) {
    const size_t lsm = 2 * tc_tile_bytes8(kTcRows, p.Kp) + 2 * tc_tile_bytes8(2 * kDec2Rows, p.Kp) + 4 * (size_t)kDec2Sub * d.Kfull * 4;
    if (lsm + 2048 <= 227 * 1024) {
      auto kernel = d.KV == 1 ? (bn ? tc_dec2_kernel<1, 4, 1, true, true> : tc_dec2_kernel<1, 4, 1, false, true>)
                              : (bn ? tc_dec2_kernel<2, 4, 1, true, true> : tc_dec2_kernel<2, 4, 1, false, true>);
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
      const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
      ProfScope prof(kclass, st);
      kernel<<<grid, kDec2Threads, lsm, st>>>(p);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
  }
  if (d.KV == 1) {
    if (vec4) return bn ? tc_linear_go<true, 1, true, false>(p, kclass, st) : tc_linear_go<true, 1, false, false>(p, kclass, st);
    return bn ? tc_linear_go<false, 1, true, false>(p, kclass, st) : tc_linear_go<false, 1, false, false>(p, kclass, st);
  }
  if (vec4) return bn ? tc_linear_go<true, 2, true, false>(p, kclass, st) : tc_linear_go<true, 2, false, false>(p, kclass, st);
  return bn ? tc_linear_go<false, 2, true, false>(p, kclass, st) : tc_linear_go<false, 2, false, false>(p, kclass, st);
}


// Integer encoder of the fixed-point model on the warp-specialised kernel (FX instantiation of tc_enc2_kernel).  W8: [Np][Kp] int8
// payloads in the lane-interleaved K order (FxpTcLinearDev::W8[1]).  S5_E_UNSUPPORTED: shape not covered (the caller falls back).
int tc_enc_fx_launch(const signed char* W8, int Kfull, int Kp, int Np, int Nfull, int KV, const void* X, int x_is_float, int in_bits, int in_exp,
                     int* Y, long long rows, int rshift, int out_bits, const int* bias_aligned, int relu, unsigned* cmax, unsigned* cmin,
                     cudaStream_t st) {
  if (!(KV == 1 || KV == 2) || !aligned16(X) || !aligned16(Y) || Np > 256 || (Np & 31) || (Nfull & 3) || Nfull > Np || (Kp & 31))
    return S5_E_UNSUPPORTED;
  const size_t raw2 = (((size_t)kEnc2Sub * Kfull * 4) + 127) & ~(size_t)127;
  const int nblk2 = (Np + 127) >> 7;
  int nring = kEnc2Ring;
  size_t esm2 = 0;
  for (; nring >= 2; --nring) {
    esm2 = (size_t)nblk2 * tc_tile_bytes8(kTcRows, Kp) + 2 * tc_tile_bytes8(2 * kEnc2Rows, Kp) + nring * raw2;
    if (esm2 + 2048 <= 227 * 1024) break;
  }
  if (nring < 2) return S5_E_UNSUPPORTED;
  TcLinParams p{};
  p.X = reinterpret_cast<const float*>(X); p.Y = reinterpret_cast<float*>(Y); p.rows = rows;
  p.Kfull = Kfull; p.Nfull = Nfull; p.Kp = Kp; p.Np = Np; p.x_stride = Kfull; p.y_stride = Nfull; p.W8 = W8; p.relu = relu;
  p.fx_x_is_float = x_is_float; p.fx_in_bits = in_bits; p.fx_in_exp = in_exp; p.fx_rs = rshift; p.fx_out_bits = out_bits;
  p.fx_bias = bias_aligned; p.fx_cmax = cmax; p.fx_cmin = cmin;
  auto kernel = KV == 1 ? (nring == 3 ? tc_enc2_kernel<1, 3, true> : tc_enc2_kernel<1, 2, true>) : (nring == 3 ? tc_enc2_kernel<2, 3, true> : tc_enc2_kernel<2, 2, true>);
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm2));
  const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
  ProfScope prof(S5_K_FXP_ENC, st);
  kernel<<<grid, kEnc2Threads, esm2, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// Integer decoder of the fixed-point model on the warp-specialised kernel (FX instantiation of tc_dec2_kernel).  W8 / WN: int8
// payloads in the natural K order (FxpTcLinearDev::W8[0], WN[0]).  S5_E_UNSUPPORTED: shape not covered (the caller falls back).
int tc_dec_fx_launch(const signed char* W8, const signed char* WN, int Kfull, int Kp, int Np, int Nfull, int KV, const int* X, int* Y, float* Yf,
                     long long rows, const int* ex_in, int inp_bits, int inp_exp, int w_exp, int out_exp, int out_bits, const int* bias_aligned,
                     float out_scale, int* status, cudaStream_t st) {
  if (!(KV == 1 || KV == 2) || (Kfull & 3) || !aligned16(X) || (Y && !aligned16(Y)) || (Yf && !aligned16(Yf)) || Np > 256 || (Np & 31) ||
      (Kp & 31) || Nfull != Np + 1)
    return S5_E_UNSUPPORTED;
  const size_t fixed = 2 * tc_tile_bytes8(kTcRows, Kp) + 2 * tc_tile_bytes8(2 * kDec2Rows, Kp);
  const size_t slot = (size_t)kDec2Sub * Kfull * 4, ohalf = (size_t)kDec2Sub * Nfull * 4;
  const size_t dsm = fixed + 3 * slot + 4 * ohalf;
  if (dsm + 2048 > 227 * 1024) return S5_E_UNSUPPORTED;
  TcLinParams p{};
  p.X = reinterpret_cast<const float*>(X); p.Y = reinterpret_cast<float*>(Y); p.rows = rows;
  p.Kfull = Kfull; p.Nfull = Nfull; p.Kp = Kp; p.Np = Np; p.x_stride = Kfull; p.y_stride = Nfull; p.W8 = W8; p.WN = WN;
  p.fx_out_bits = out_bits; p.fx_bias = bias_aligned; p.fx_Yf = Yf; p.fx_out_scale = out_scale; p.fx_ex_in = ex_in;
  p.fx_dec_inp_bits = inp_bits; p.fx_dec_inp_exp = inp_exp; p.fx_dec_w_exp = w_exp; p.fx_dec_out_exp = out_exp; p.fx_status = status;
  auto kernel = KV == 1 ? tc_dec2_kernel<1, 3, 4, false, false, true> : tc_dec2_kernel<2, 3, 4, false, false, true>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
  const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, (long long)sm_count()));
  ProfScope prof(S5_K_FXP_DEC, st);
  kernel<<<grid, kDec2Threads, dsm, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// ---- C projection + GLU handle
static size_t tc_glu_smem(int H, int K1, bool scan, bool bpj) {
  const size_t bu_bytes = bpj ? (size_t)kTcRows * (K1 + 4) * 4 : 0;
  const size_t seg_bytes = scan ? (size_t)(2 * kGluHalf + 128) * sizeof(float2) : 0;
  const size_t r_bytes = std::max(std::max(2 * tc_tile_bytes(kTcRows, K1), bu_bytes) + seg_bytes,
                                  (size_t)kTcRows * (H + 4) * 4 + 2 * tc_tile_bytes8(kTcRows, kGluK2));
  return tc_tile_bytes(H, K1) + tc_tile_bytes8(H, kGluK2) + (bpj ? tc_tile_bytes8(K1, kGluK2) : 0) +
         (((size_t)(5 * H + K1) * 4 + 127) & ~(size_t)127) + 2 * r_bytes;
}

static size_t tc_gluw_smem(int H, int K1) {
  const int HJ = H >> 5, K2 = gluw_k2(HJ);
  const size_t buf = 2 * tc_tile_bytes(kTcRows, kGluwKC) + tc_tile_bytes(H, kGluwKC);
  (void)K1;
  return tc_tile_bytes8(H, K2) + (((size_t)(5 * H) * 4 + 127) & ~(size_t)127) + std::max((size_t)kTcRows * (H + 4) * 4, 2 * buf) +
         2 * tc_tile_bytes8(kTcRows, K2);
}

// wide form: chunked C, wide out2 K order; GLU only (B projection and recurrence run as their own kernels)
static int tc_gluw_build(const float* C_tilde, const float* out2_kernel, int H, int P, float conj_mul, float s_o_act, TcGluDev* out) {
  const int K1 = 2 * P, HJ = H >> 5;
  if ((H & 31) || H > 256 || (K1 % kGluwKC) || K1 > 512 || !is_pow2(s_o_act)) return S5_OK;
  const size_t smem = tc_gluw_smem(H, K1);
  if (smem + 2048 > 227 * 1024) return S5_OK;
  const size_t nC = (size_t)H * P * 2;
  const float s_re = pow2_grid(C_tilde, nC, 0, 2), s_im = pow2_grid(C_tilde, nC, 1, 2);
  const float s_w = pow2_grid(out2_kernel, (size_t)H * H, 0, 1);
  const int K2 = gluw_k2(HJ), NC = K1 / kGluwKC;
  const size_t lboH = tc_lbo(H), wck = tc_tile_bytes(H, kGluwKC);
  std::vector<__nv_bfloat16> wc(NC * wck / 2, __float2bfloat16_rn(0.f));
  std::vector<signed char> wo((size_t)H * K2, 0);
  for (int h = 0; h < H; ++h)
    for (int k = 0; k < K1; ++k) {
      const int p = k >> 1;
      const float s = (k & 1) ? s_im : s_re;
      const float c = std::nearbyint(C_tilde[((size_t)h * P + p) * 2 + (k & 1)] / s);
      if (std::fabs(c) > 256.f) return S5_OK;                  // not exact in bf16
      const int ch = k / kGluwKC, kk = k % kGluwKC;
      const size_t byte = (size_t)ch * wck + (size_t)(kk >> 3) * lboH + (size_t)h * 16 + (size_t)(kk & 7) * 2;
      wc[byte / 2] = __float2bfloat16_rn((k & 1) ? -c : c);    // Re(C x) = C_re x_re - C_im x_im
    }
  for (int n = 0; n < H; ++n)
    for (int k = 0; k < H; ++k) {
      const float c = std::nearbyint(out2_kernel[(size_t)k * H + n] / s_w);
      if (c < -128.f || c > 127.f) return S5_OK;
      wo[(size_t)n * K2 + gluw_kpos(k, HJ)] = (signed char)c;
    }
  out->H = H; out->K1 = K1; out->smem_wide = smem;
  out->xs_odd_mul = s_im / s_re; out->c_scale = conj_mul * s_re; out->o_col_scale = s_o_act * s_w;
  S5_CUDA(cudaMalloc(&out->Wc, wc.size() * sizeof(__nv_bfloat16)));
  S5_CUDA(cudaMemcpy(out->Wc, wc.data(), wc.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  S5_CUDA(cudaMalloc(&out->Wo, wo.size()));
  S5_CUDA(cudaMemcpy(out->Wo, wo.data(), wo.size(), cudaMemcpyHostToDevice));
  out->wide_ok = true;
  return S5_OK;
}

int tc_glu_build(const float* C_tilde /*[H][P][2] quantised*/, const float* out2_kernel /*[H][H] quantised*/, int H, int P,
                 float conj_mul, float s_o_act, const float* B_bar /*[H][2P] quantised, nullable*/, float s_ut, TcGluDev* out) {
  out->ok = false;
  const int K1 = 2 * P;
  if ((H & 31) || (K1 & 15) || !is_pow2(s_o_act)) return S5_OK;
  if (H > 128 || K1 > 128) return tc_gluw_build(C_tilde, out2_kernel, H, P, conj_mul, s_o_act, out);
  const size_t smem = tc_glu_smem(H, K1, false, false);
  if (smem > 226 * 1024) return S5_OK;
  const size_t nC = (size_t)H * P * 2;
  const float s_re = pow2_grid(C_tilde, nC, 0, 2), s_im = pow2_grid(C_tilde, nC, 1, 2);
  const float s_w = pow2_grid(out2_kernel, (size_t)H * H, 0, 1);
  std::vector<__nv_bfloat16> wc((size_t)H * K1);
  std::vector<signed char> wo((size_t)H * kGluK2, 0);
  for (int h = 0; h < H; ++h)
    for (int p = 0; p < P; ++p) {
      const float cre = std::nearbyint(C_tilde[((size_t)h * P + p) * 2] / s_re);
      const float cim = std::nearbyint(C_tilde[((size_t)h * P + p) * 2 + 1] / s_im);
      if (std::fabs(cre) > 256.f || std::fabs(cim) > 256.f) return S5_OK;  // not exact in bf16
      wc[(size_t)h * K1 + 2 * p] = __float2bfloat16_rn(cre);
      wc[(size_t)h * K1 + 2 * p + 1] = __float2bfloat16_rn(-cim);          // Re(C x) = C_re x_re - C_im x_im
    }
  for (int n = 0; n < H; ++n)
    for (int k = 0; k < H; ++k) {
      const float c = std::nearbyint(out2_kernel[(size_t)k * H + n] / s_w);
      if (c < -128.f || c > 127.f) return S5_OK;
      wo[(size_t)n * kGluK2 + 4 * (k & 31) + (k >> 5)] = (signed char)c;   // lane-interleaved K order of pass P1
    }
  out->H = H; out->K1 = K1; out->smem = smem;
  out->xs_odd_mul = s_im / s_re; out->c_scale = conj_mul * s_re; out->o_col_scale = s_o_act * s_w;
  S5_CUDA(cudaMalloc(&out->Wc, wc.size() * sizeof(__nv_bfloat16)));
  S5_CUDA(cudaMemcpy(out->Wc, wc.data(), wc.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  S5_CUDA(cudaMalloc(&out->Wo, wo.size()));
  S5_CUDA(cudaMemcpy(out->Wo, wo.data(), wo.size(), cudaMemcpyHostToDevice));
  out->ok = true;
  out->smem_scan = tc_glu_smem(H, K1, true, false);
  out->smem_bpj = tc_glu_smem(H, K1, true, true);
  // fused B projection: int8 codes of B_bar on per-parity power-of-two grids (real / imaginary outputs)
  if (B_bar && is_pow2(s_ut) && out->smem_bpj <= 227 * 1024 - 1280) {
    std::vector<float> c0, c1;
    for (int k = 0; k < H; ++k)
      for (int n = 0; n < K1; ++n) ((n & 1) ? c1 : c0).push_back(B_bar[(size_t)k * K1 + n]);
    const float g0 = pow2_grid(c0.data(), c0.size(), 0, 1), g1 = pow2_grid(c1.data(), c1.size(), 0, 1);
    std::vector<signed char> wb((size_t)K1 * kGluK2, 0);
    std::vector<float> cs(K1);
    bool fits = true;
    for (int n = 0; n < K1 && fits; ++n) {
      const float g = (n & 1) ? g1 : g0;
      cs[n] = s_ut * g;
      for (int k = 0; k < H; ++k) {
        const float c = std::nearbyint(B_bar[(size_t)k * K1 + n] / g);
        if (c < -128.f || c > 127.f) { fits = false; break; }
        wb[(size_t)n * kGluK2 + 4 * (k & 31) + (k >> 5)] = (signed char)c;
      }
    }
    if (fits) {
      S5_CUDA(cudaMalloc(&out->WB, wb.size()));
      S5_CUDA(cudaMemcpy(out->WB, wb.data(), wb.size(), cudaMemcpyHostToDevice));
      int rc;
      if ((rc = upload_f(cs, &out->b_col_scale))) return rc;
      out->bpj_ok = true;
    }
  }
  return S5_OK;
}

void tc_glu_free(TcGluDev* d) {
  cudaFree(d->Wc); cudaFree(d->Wo); cudaFree(d->WB); cudaFree(d->b_col_scale); cudaFree(d->img);
  *d = TcGluDev();
}

template <int HJ, bool SCAN, bool BPJ, int KJ = 0>
static int tc_glu_go_multi(const TcGluMulti& mp, size_t smem, int grid, cudaStream_t st) {
  auto kernel = tc_glu_kernel<HJ, SCAN, BPJ, KJ>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(BPJ ? S5_K_BLOCK : SCAN ? S5_K_GLU_SCAN : S5_K_GLU, st);
  if constexpr (SCAN) {
    // The fused recurrence spins on tiles owned by other CTAs (look-back, done flags) and its persistent round-robin
    // schedule lets a CTA wait for a HIGHER-numbered one (previous round), so every CTA of the grid must be resident.
    // A cooperative launch guarantees exactly that (grid <= SM count, 1 CTA per SM); with a plain launch two such
    // kernels on different streams (the host pipeline runs two forward streams) could each become partially resident
    // and wait for their own unscheduled CTAs forever.
    // The grid is clamped to what the device can keep resident for THIS shared-memory size (an MPS SM limit or another
    // carve-out can make that less than one CTA per SM; the round-robin tile schedule works for any grid).  If the
    // launch is refused all the same, kCoopRefused tells the caller to run the unfused kernels instead.
    int per_sm = 0;
    S5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGluThreads, smem));
    if (const char* e = getenv("S5B200_TEST_REFUSE_COOP")) per_sm = atoi(e) ? 0 : per_sm;   // test hook for the fallback
    if (per_sm < 1) {
      set_error("dense-tile block kernel: no CTA fits an SM with %zu bytes of shared memory", smem);
      return kCoopRefused;
    }
    grid = std::min(grid, sm_count());                       // one CTA per SM: the two tile pipelines share its weights
    void* args[] = {const_cast<TcGluMulti*>(&mp)};
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(grid), dim3(kGluThreads), args, smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
      (void)cudaGetLastError();                               // not sticky: clear it, the fallback launches next
      set_error("dense-tile block kernel: cooperative launch of %d CTAs refused (%s)", grid, cudaGetErrorString(e));
      return kCoopRefused;
    }
    S5_CUDA(e);
  } else {
    kernel<<<grid, kGluThreads, smem, st>>>(mp);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}
template <int HJ, bool SCAN, bool BPJ>
static int tc_glu_go(const TcGluParams& p, size_t smem, int grid, cudaStream_t st) {
  TcGluMulti mp{};
  mp.nlayers = 1; mp.done = nullptr; mp.layer[0] = p;
  return tc_glu_go_multi<HJ, SCAN, BPJ>(mp, smem, grid, st);
}

size_t tc_glu_scan_workspace_bytes(int B, int L, int P) {
  const int cmax = (L + 31) / 32 + 1;                          // upper bound on tiles per clip
  return round_up((size_t)B * cmax * sizeof(uint32_t), 256) + (size_t)B * cmax * 2 * P * sizeof(float2);
}

// Builds (once per handle and set of scales) the block kernel's per-layer image in global memory.  The first call
// synchronises the stream so that the image is complete before any other stream uses the handle.
static int tc_glu_ensure_image(const TcGluDev& d, TcGluParams* p, cudaStream_t st) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  struct Key {
    const void *Wc, *Wo, *WB, *bcs, *bn_sub, *bn_mul, *bn_add, *D, *o_bias;
    float ut_s, ut_inv, yt_inv, oo_inv, b_inv0, b_inv1;
  } key{p->Wc, p->Wo, p->WB, p->b_col_scale, p->bn_sub, p->bn_mul, p->bn_add, p->D, p->o_bias,
        p->ut_s, p->ut_inv, p->yt_inv, p->oo_inv, p->b_out_inv0, p->b_out_inv1};
  static_assert(sizeof(Key) <= sizeof(d.img_key), "image key too large");
  const size_t bytes = tc_glu_image_bytes(d.H, d.K1, true);
  if (!d.img_ok || std::memcmp(d.img_key, &key, sizeof(Key)) != 0) {
    if (!d.img) S5_CUDA(cudaMalloc(&d.img, bytes));
    S5_CUDA(cudaMemsetAsync(d.img, 0, bytes, st));
    TcGluParams q = *p;
    q.img = nullptr;
    tc_glu_image_kernel<<<1, 512, 0, st>>>(q, reinterpret_cast<unsigned char*>(d.img));
    S5_LAUNCH_CHECK();
    S5_CUDA(cudaStreamSynchronize(st));
    std::memset(d.img_key, 0, sizeof(d.img_key));
    std::memcpy(d.img_key, &key, sizeof(Key));
    d.img_ok = true;
  }
  p->img = reinterpret_cast<const unsigned char*>(d.img);
  p->img_bytes = (uint32_t)bytes;
  return S5_OK;
}

// Builds the block kernel's image for the scales / per-column vectors the model will use, at model-create time (the
// forward then finds a matching key: no allocation, no synchronisation and no handle mutation on the hot path, which
// also keeps it capturable in a CUDA graph).  Only the image-relevant fields of `a` are read.
int tc_glu_prepare_image(const TcGluDev& d, const TcGluArgs& a) {
  if (!d.ok || !d.bpj_ok || !tc_glu_scan_ok(d) || !is_pow2(a.s_ut) || !is_pow2(a.s_yt) || !is_pow2(a.s_o_out) || !is_pow2(a.s_bu_re) ||
      !is_pow2(a.s_bu_im))
    return S5_OK;                                             // this layer never runs the block kernel
  TcGluParams p{};
  p.H = d.H; p.K1 = d.K1;
  p.Wc = reinterpret_cast<const __nv_bfloat16*>(d.Wc); p.Wo = reinterpret_cast<const signed char*>(d.Wo);
  p.xs_odd_mul = d.xs_odd_mul; p.c_scale = d.c_scale; p.o_col_scale = d.o_col_scale;
  p.bn_sub = a.bn_sub; p.bn_mul = a.bn_mul; p.bn_add = a.bn_add; p.D = a.D; p.o_bias = a.o_bias;
  p.ut_s = a.s_ut; p.ut_inv = 1.f / a.s_ut; p.yt_s = a.s_yt; p.yt_inv = 1.f / a.s_yt; p.oa_inv = 1.f / a.s_o_act;
  p.oo_s = a.s_o_out; p.oo_inv = 1.f / a.s_o_out; p.gl_s = a.s_gl; p.gl_inv = 1.f / a.s_gl; p.gr_s = a.s_gr; p.gr_inv = 1.f / a.s_gr;
  p.qs_min = a.qs_min; p.qs_max = a.qs_max; p.qn_min = a.qn_min; p.qn_max = a.qn_max; p.relu_out = a.relu_out;
  p.WB = reinterpret_cast<const signed char*>(d.WB); p.b_col_scale = d.b_col_scale;
  p.b_out_scale0 = a.s_bu_re; p.b_out_scale1 = a.s_bu_im; p.b_out_inv0 = 1.f / a.s_bu_re; p.b_out_inv1 = 1.f / a.s_bu_im;
  return tc_glu_ensure_image(d, &p, nullptr);
}

// fills the kernel parameters of one layer; *grid_out = grid size this layer alone would use
static int tc_glu_fill(const TcGluDev& d, const TcGluArgs& a, TcGluParams* out, int* grid_out, cudaStream_t st) {
  const float sc[] = {a.s_ut, a.s_yt, a.s_o_act, a.s_o_out, a.s_gl, a.s_gr};
  for (float s : sc)
    if (!is_pow2(s)) S5_FAIL(S5_E_BADARG, "dense-tile glu: scales must be powers of two");
  if (!d.ok) S5_FAIL(S5_E_BADARG, "dense-tile glu: not eligible");
  if (!a.fuse_bproj && !aligned16(a.XS)) S5_FAIL(S5_E_BADARG, "dense-tile glu: xs must be 16-byte aligned");
  TcGluParams p{};
  p.XS = a.XS; p.X = a.X; p.OUT = a.OUT; p.rows = a.rows; p.H = d.H; p.K1 = d.K1;
  p.Wc = reinterpret_cast<const __nv_bfloat16*>(d.Wc); p.Wo = reinterpret_cast<const signed char*>(d.Wo);
  p.xs_odd_mul = d.xs_odd_mul; p.c_scale = d.c_scale; p.o_col_scale = d.o_col_scale;
  p.bn_sub = a.bn_sub; p.bn_mul = a.bn_mul; p.bn_add = a.bn_add; p.D = a.D; p.o_bias = a.o_bias;
  p.ut_s = a.s_ut; p.ut_inv = 1.f / a.s_ut; p.yt_s = a.s_yt; p.yt_inv = 1.f / a.s_yt; p.oa_inv = 1.f / a.s_o_act;
  p.oo_s = a.s_o_out; p.oo_inv = 1.f / a.s_o_out; p.gl_s = a.s_gl; p.gl_inv = 1.f / a.s_gl; p.gr_s = a.s_gr; p.gr_inv = 1.f / a.s_gr;
  p.qs_min = a.qs_min; p.qs_max = a.qs_max; p.qn_min = a.qn_min; p.qn_max = a.qn_max; p.relu_out = a.relu_out;
  if (!a.lam) {
    // rows are split evenly over 2 x grid pipelines (balanced partition); at least 32 rows per pipeline
    *grid_out = (int)std::max<long long>(1, std::min<long long>((p.rows + 63) / 64, (long long)sm_count()));
    *out = p;
    return S5_OK;
  }
  // ---- fused recurrence: XS is Bu (unused with the fused B projection).  Tiles are R consecutive frames of one clip.
  const int P = d.K1 >> 1;
  if (!tc_glu_scan_ok(d) || a.B <= 0 || a.L <= 0 || (long long)a.B * a.L != a.rows) S5_FAIL(S5_E_BADARG, "dense-tile glu: fused recurrence not eligible");
  // Rows per tile depend on L only (never on B): the grouping of the floating-point recurrence, and with it every
  // output bit, is then independent of how clips are batched, chunked by the host pipeline or sharded over GPUs.
  // S5B200_SCAN_TILE_ROWS (<= 128) overrides the tile height; measured on B200 at 32 x 3751 frames: 128 -> 98.6 us,
  // 104 (a whole number of rounds over the 296 pipelines) -> 100.6 us, 96 -> 110.5 us, 64 -> 125 us.
  static const int tile_rows = [] { const char* e = getenv("S5B200_SCAN_TILE_ROWS"); return e ? std::max(8, std::min(kTcRows, atoi(e))) : kTcRows; }();
  const int nseg = kGluHalf / P, rmax = std::min(tile_rows, 32 * nseg);
  // Full-height tiles (every clip ends with one short tile) rather than balanced ones: with P = 64 every segment of a full
  // tile is 32 frames long and takes the kernel's straight-line recurrence path.  S5B200_SCAN_BALANCED=1: the old split.
  static const bool balanced = [] { const char* e = getenv("S5B200_SCAN_BALANCED"); return e && atoi(e) != 0; }();
  const int c = (a.L + rmax - 1) / rmax;
  p.R = balanced ? (a.L + c - 1) / c : std::min(rmax, a.L);
  p.tiles_per_clip = (a.L + p.R - 1) / p.R;
  const long long tiles = (long long)a.B * p.tiles_per_clip;
  *grid_out = (int)std::max<long long>(1, std::min<long long>((tiles + 1) / 2, (long long)sm_count()));
  p.lam = reinterpret_cast<const float2*>(a.lam); p.B = a.B; p.L = a.L;
  p.relu_scan = a.relu_scan;
  const size_t flags_bytes = round_up((size_t)a.B * p.tiles_per_clip * sizeof(uint32_t), 256);
  const size_t need = flags_bytes + (size_t)a.B * p.tiles_per_clip * 2 * P * sizeof(float2);
  if (!a.scan_ws || a.scan_ws_bytes < need || !aligned16(a.scan_ws)) S5_FAIL(S5_E_BADARG, "dense-tile glu: recurrence workspace %zu < %zu bytes", a.scan_ws_bytes, need);
  p.flags = reinterpret_cast<uint32_t*>(a.scan_ws);
  p.recs = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(a.scan_ws) + flags_bytes);
  S5_CUDA(cudaMemsetAsync(p.flags, 0, flags_bytes, st));
  if (a.fuse_bproj) {
    if (!d.bpj_ok || !is_pow2(a.s_bu_re) || !is_pow2(a.s_bu_im)) S5_FAIL(S5_E_BADARG, "dense-tile glu: fused B projection not eligible");
    p.WB = reinterpret_cast<const signed char*>(d.WB); p.b_col_scale = d.b_col_scale;
    p.b_out_scale0 = a.s_bu_re; p.b_out_scale1 = a.s_bu_im; p.b_out_inv0 = 1.f / a.s_bu_re; p.b_out_inv1 = 1.f / a.s_bu_im;
    static const bool use_img = [] { const char* e = getenv("S5B200_LAYER_IMAGE"); return !e || atoi(e) != 0; }();
    int rc;
    if (use_img && (rc = tc_glu_ensure_image(d, &p, st))) return rc;
  }
  *out = p;
  return S5_OK;
}

template <int HJ>
static int tc_gluw_go(const TcGluParams& p, size_t smem, int grid, cudaStream_t st) {
  auto kernel = tc_gluw_kernel<HJ>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(S5_K_GLU, st);
  kernel<<<grid, kGluwThreads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

int tc_glu_launch(const TcGluDev& d, const TcGluArgs& a, cudaStream_t st) {
  TcGluParams p{};
  int grid = 1, rc;
  if (!d.ok && d.wide_ok) {
    if (a.lam) S5_FAIL(S5_E_BADARG, "dense-tile glu (wide): the recurrence is not fused at this size");
    TcGluDev d1 = d;
    d1.ok = true;                                              // the parameter block is filled the same way
    if ((rc = tc_glu_fill(d1, a, &p, &grid, st))) return rc;
    grid = (int)std::max<long long>(1, std::min<long long>((p.rows + 63) / 64, (long long)sm_count()));
    switch (d.H >> 5) {
      case 1: return tc_gluw_go<1>(p, d.smem_wide, grid, st);
      case 2: return tc_gluw_go<2>(p, d.smem_wide, grid, st);
      case 3: return tc_gluw_go<3>(p, d.smem_wide, grid, st);
      case 4: return tc_gluw_go<4>(p, d.smem_wide, grid, st);
      case 5: return tc_gluw_go<5>(p, d.smem_wide, grid, st);
      case 6: return tc_gluw_go<6>(p, d.smem_wide, grid, st);
      default: S5_FAIL(S5_E_BADSHAPE, "dense-tile glu (wide): H = %d not covered", d.H);
    }
  }
  if ((rc = tc_glu_fill(d, a, &p, &grid, st))) return rc;
  const int HJ = d.H >> 5;
  if (!a.lam) {
    switch (HJ) {
      case 1: return tc_glu_go<1, false, false>(p, d.smem, grid, st);
      case 2: return tc_glu_go<2, false, false>(p, d.smem, grid, st);
      case 3: return tc_glu_go<3, false, false>(p, d.smem, grid, st);
      default: return tc_glu_go<4, false, false>(p, d.smem, grid, st);
    }
  }
  if (a.fuse_bproj) {
    switch (HJ) {
      case 1: return tc_glu_go<1, true, true>(p, d.smem_bpj, grid, st);
      case 2: return tc_glu_go<2, true, true>(p, d.smem_bpj, grid, st);
      case 3: return tc_glu_go<3, true, true>(p, d.smem_bpj, grid, st);
      default: return tc_glu_go<4, true, true>(p, d.smem_bpj, grid, st);
    }
  }
  switch (HJ) {
    case 1: return tc_glu_go<1, true, false>(p, d.smem_scan, grid, st);
    case 2: return tc_glu_go<2, true, false>(p, d.smem_scan, grid, st);
    case 3: return tc_glu_go<3, true, false>(p, d.smem_scan, grid, st);
    default: return tc_glu_go<4, true, false>(p, d.smem_scan, grid, st);
  }
}

// All blocks in one launch (every layer: block kernel with the same H, P, B, L).  `done` holds (n - 1) * tiles flags.
int tc_glu_launch_multi(const TcGluDev* const* devs, const TcGluArgs* args, int n, uint32_t* done, size_t done_bytes, cudaStream_t st) {
  if (n < 1 || n > kMaxFusedLayers) S5_FAIL(S5_E_BADARG, "dense-tile block kernel: %d layers in one launch not supported", n);
  TcGluMulti mp{};
  mp.nlayers = n;
  int grid = 1, rc;
  for (int l = 0; l < n; ++l) {
    if (!args[l].lam || !args[l].fuse_bproj || devs[l]->H != devs[0]->H || devs[l]->K1 != devs[0]->K1)
      S5_FAIL(S5_E_BADARG, "dense-tile block kernel: layer %d cannot join the single launch", l);
    if ((rc = tc_glu_fill(*devs[l], args[l], &mp.layer[l], &grid, st))) return rc;
  }
  const size_t tiles = (size_t)mp.layer[0].B * mp.layer[0].tiles_per_clip;
  if (n > 1) {
    if (!done || done_bytes < (size_t)(n - 1) * tiles * sizeof(uint32_t)) S5_FAIL(S5_E_BADARG, "dense-tile block kernel: done-flag workspace too small");
    S5_CUDA(cudaMemsetAsync(done, 0, (size_t)(n - 1) * tiles * sizeof(uint32_t), st));
    mp.done = done;
    // all layers' tiles share the pipelines
    grid = (int)std::max<long long>(1, std::min<long long>(((long long)tiles * n + 1) / 2, (long long)sm_count()));
  }
  const size_t smem = devs[0]->smem_bpj;
  switch (devs[0]->H >> 5) {
    case 1: return tc_glu_go_multi<1, true, true>(mp, smem, grid, st);
    case 2: return tc_glu_go_multi<2, true, true>(mp, smem, grid, st);
    case 3:
      // NDNS dim_scale 0.5: H 96, P 64, full-height tiles (every clip of at least 128 frames)
      if (devs[0]->K1 == 128 && mp.layer[0].R == kTcRows) return tc_glu_go_multi<3, true, true, 4>(mp, smem, grid, st);
      return tc_glu_go_multi<3, true, true>(mp, smem, grid, st);
    default: return tc_glu_go_multi<4, true, true>(mp, smem, grid, st);
  }
}

bool tc_glu_scan_ok(const TcGluDev& d) {
  const int P = d.K1 >> 1;
  return d.ok && (P == 32 || P == 64);   // whole warps of carry threads, <= 32 frames per segment
}

}  // namespace s5
