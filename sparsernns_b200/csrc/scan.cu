// scan.cu - diagonal linear recurrences of the S5 SSM.
//
//  (1) scan_c64: x_k = lam * x_{k-1} + Bu_k in complex64, replacing jax.lax.associative_scan /
//      jax.lax.scan (reference sparseRNNs/model/ssm.py:127, :150, :313).  Single pass over HBM:
//      each CTA owns one time chunk [TL][P] of one clip, brought in by one 1-D bulk async copy (TMA
//      engine, UBLKCP) and written back by one bulk store.  Inside the chunk every thread owns one
//      state p of one sub-chunk of T consecutive steps; sub-chunk and chunk carries are combined with
//      Horner steps on lam^T / lam^TL, the cross-chunk carry travels through a decoupled look-back
//      (aggregate / inclusive records + flags in a caller-provided workspace).
//  (2) scan_i32: the integer recurrence of fxpmodel.py:147-208.  It floors every product separately, so it
//      is not associative and must be sequential in time; one lane per (clip, state), register
//      prefetch ring over time.
#include <algorithm>

#include "scan.cuh"

namespace s5 {

constexpr int kScanT = 16;        // steps a thread scans sequentially (sub-chunk length)
constexpr int kScanThreads = 256;  // upper bound for nsub * P

struct ScanGeom {
  int nsub, threads, TL, nchunks;
  size_t flags_bytes, rec_floats2;  // per-(clip, chunk) record: agg[P] + incl[P] (double2 each)
  size_t smem_bytes;
};

static ScanGeom scan_geom(int L, int P) {
  ScanGeom g;
  g.nsub = kScanThreads / P;
  if (g.nsub < 1) g.nsub = 1;
  g.threads = g.nsub * P;
  g.TL = g.nsub * kScanT;
  g.nchunks = (L + g.TL - 1) / g.TL;
  g.rec_floats2 = 2 * (size_t)round_up((size_t)P, 16);
  g.smem_bytes = (size_t)g.TL * P * 8 + (size_t)g.nsub * P * 8 + (size_t)P * 16 + 16;
  return g;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * x + v
__device__ __forceinline__ float2 cfma(float2 a, float2 x, float2 v) {
  return make_float2(fmaf(a.x, x.x, fmaf(-a.y, x.y, v.x)), fmaf(a.x, x.y, fmaf(a.y, x.x, v.y)));
}
__device__ __forceinline__ float2 relu_lex(float2 z) {
  const bool keep = (z.x > 0.f) || (z.x == 0.f && z.y > 0.f);
  return keep ? z : make_float2(0.f, 0.f);
}

// double-precision complex helpers: chunk carries are combined in fp64 so that the powers lam^T, lam^TL,
// lam^(k*TL) do not accumulate an O(k * eps_f32) error (the states themselves stay fp32 like the reference).
__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 zfma(double2 a, double2 x, double2 v) {
  return make_double2(fma(a.x, x.x, fma(-a.y, x.y, v.x)), fma(a.x, x.y, fma(a.y, x.x, v.y)));
}
__device__ __forceinline__ double2 f2z(float2 v) { return make_double2((double)v.x, (double)v.y); }

// LB_FIXED: fixed-stride look-back (bit-reproducible, two round trips per chunk) for grids larger than one resident wave;
// otherwise the classic walk (lowest latency when a whole clip is co-resident, e.g. batch 1).
template <bool LB_FIXED>
__global__ void __launch_bounds__(kScanThreads, 4)
scan_c64_kernel(const float2* __restrict__ Bu, const float2* __restrict__ lam, float2* xs, int L, int P, int nsub,
                int nchunks, int relu, uint32_t* flags, double2* recs, int rec_stride) {
  constexpr bool lb_fixed = LB_FIXED;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TL = nsub * kScanT;
  float2* tile = reinterpret_cast<float2*>(smem_raw);        // [TL][P]
  float2* sub_agg = tile + (size_t)TL * P;                    // [nsub][P]
  double2* carry_s = reinterpret_cast<double2*>(sub_agg + (size_t)nsub * P);  // [P]
  uint64_t* bar = reinterpret_cast<uint64_t*>(carry_s + P);

  // Block order is clip-fastest: all clips' chunk c are dispatched before any chunk c+1, so by the time a chunk
  // starts, its predecessors are usually finished and the look-back finds an inclusive prefix after 1-2 steps.
  const int c = blockIdx.y, b = blockIdx.x;
  const int tid = threadIdx.x;
  const int p = tid % P, s = tid / P;
  const int t0 = c * TL;
  const int rows = min(TL, L - t0);
  const size_t base = ((size_t)b * L + t0) * P;
  const uint32_t bytes = (uint32_t)rows * P * 8u;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, bytes);
    bulk_g2s(tile, Bu + base, bytes, bar);
  }

  // lam^T and lam^TL (fp64) while the tile is in flight
  const float2 a = lam[p];
  double2 aT = f2z(a);
#pragma unroll
  for (int i = 1; i < kScanT; i <<= 1) aT = zmul(aT, aT);
  double2 aTL = aT;
  for (int i = 1; i < nsub; ++i) aTL = zmul(aTL, aT);

  mbar_wait(bar, 0);

  // pass 1: local recurrence from zero -> sub-chunk aggregate
  const int ts = s * kScanT;
  float2 x = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < kScanT; ++j) {
    const int t = ts + j;
    if (t < rows) x = cfma(a, x, tile[(size_t)t * P + p]);
  }
  sub_agg[s * P + p] = x;
  __syncthreads();

  // Only the first P threads (the "carry warps") take part in the aggregate / look-back / publish protocol;
  // they synchronise among themselves on named barrier 1 and the rest of the CTA waits at the barrier below.
  uint32_t* flag_row = flags + (size_t)b * nchunks;
  double2* rec_row = recs + (size_t)b * nchunks * rec_stride;
  const int half = rec_stride >> 1;
  const bool publish = (c + 1 < nchunks);  // nobody looks back at the last chunk
  const int carry_threads = (P + 31) & ~31;
  if (tid < carry_threads) {
    double2 e = make_double2(0., 0.);
    if (tid < P) {
      e = f2z(sub_agg[p]);
      for (int q = 1; q < nsub; ++q) e = zfma(aT, e, f2z(sub_agg[q * P + p]));
      if (publish) {
        double2* rec = rec_row + (size_t)c * rec_stride;
        __stcg(rec + (c == 0 ? half : 0) + p, e);  // chunk 0: aggregate == inclusive prefix
        __threadfence();
      }
      if (c == 0) carry_s[p] = make_double2(0., 0.);
    }
    if (publish) {
      asm volatile("bar.sync 1, %0;" ::"r"(carry_threads) : "memory");
      if (tid == 0) st_release_u32(flag_row + c, c == 0 ? 2u : 1u);
    }
    if constexpr (lb_fixed) if (c > 0) {
      // Fixed-stride look-back (as in the fused block kernel, tc_model.cu).  Every chunk publishes its aggregate (flag 1),
      // every kLbStride-th chunk (an "anchor") also its inclusive prefix (flag 2).  The carry-in of chunk c is ALWAYS
      //     incl[a], then Horner over agg[a + 1 .. c - 1],     a = the last anchor before c,
      // so (i) the floating-point association depends on c only - results are bit-reproducible - and (ii) all flags and
      // all records of the walk are requested together: two memory round trips per chunk instead of two per
      // predecessor, and only every kLbStride-th chunk sits on the serial anchor chain (118 chunks per 3751-frame clip
      // used to be 118 dependent hops: the kernel was latency bound at 32 clips).
      constexpr int kLbStride = 8;
      const int a0 = ((c - 1) / kLbStride) * kLbStride;
      const int n = c - a0;                                   // records to read: incl[a0], agg[a0 + 1 .. c - 1]
      const bool anchor = (c % kLbStride) == 0;
      double2 acc = make_double2(0., 0.);
      if (tid < P) {
        uint32_t f[kLbStride];
#pragma unroll
        for (int w = 0; w < kLbStride; ++w) f[w] = w < n ? ld_relaxed_u32(flag_row + a0 + w) : 2u;
        while (f[0] != 2u) f[0] = ld_relaxed_u32(flag_row + a0);
#pragma unroll
        for (int w = 1; w < kLbStride; ++w)
          while (f[w] == 0u) f[w] = ld_relaxed_u32(flag_row + a0 + w);
        fence_acquire_gpu();                                  // one acquire for the whole walk
        double2 rec[kLbStride];
#pragma unroll
        for (int w = 0; w < kLbStride; ++w)
          if (w < n) rec[w] = __ldcg(rec_row + (size_t)(a0 + w) * rec_stride + (w == 0 ? half : 0) + p);
        acc = rec[0];
#pragma unroll
        for (int w = 1; w < kLbStride; ++w)
          if (w < n) acc = zfma(aTL, acc, rec[w]);
        carry_s[p] = acc;
        if (publish && anchor) {
          __stcg(rec_row + (size_t)c * rec_stride + half + p, zfma(aTL, acc, e));
          __threadfence();
        }
      }
      if (publish && anchor) {
        asm volatile("bar.sync 1, %0;" ::"r"(carry_threads) : "memory");
        if (tid == 0) st_release_u32(flag_row + c, 2u);
      }
    }
    // classic decoupled look-back (whole clips co-resident, e.g. batch 1: a chunk may walk past predecessors on
    // aggregates alone instead of waiting for the anchor's inclusive prefix - lower latency, timing-dependent association)
    if constexpr (!lb_fixed) if (c > 0) {
      if (tid < P) {
        // decoupled look-back: walk predecessors until one has published its inclusive prefix
        double2 mult = make_double2(1., 0.);
        double2 acc = make_double2(0., 0.);
        for (int j = c - 1; j >= 0; --j) {
          uint32_t f;
          do {
            f = ld_acquire_u32(flag_row + j);
          } while (f == 0u);
          const double2* rec = rec_row + (size_t)j * rec_stride;
          const double2 v = __ldcg(rec + (f == 2u ? half : 0) + p);
          acc = zfma(mult, v, acc);
          if (f == 2u) break;
          mult = zmul(mult, aTL);
        }
        carry_s[p] = acc;
        if (publish) {
          __stcg(rec_row + (size_t)c * rec_stride + half + p, zfma(aTL, acc, e));
          __threadfence();
        }
      }
      if (publish) {
        asm volatile("bar.sync 1, %0;" ::"r"(carry_threads) : "memory");
        if (tid == 0) st_release_u32(flag_row + c, 2u);
      }
    }
  }
  __syncthreads();

  // pass 2: recurrence with the true carry-in, results in place
  double2 cin = carry_s[p];
  for (int q = 0; q < s; ++q) cin = zfma(aT, cin, f2z(sub_agg[q * P + p]));
  x = make_float2((float)cin.x, (float)cin.y);
#pragma unroll
  for (int j = 0; j < kScanT; ++j) {
    const int t = ts + j;
    if (t < rows) {
      x = cfma(a, x, tile[(size_t)t * P + p]);
      tile[(size_t)t * P + p] = relu ? relu_lex(x) : x;
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(xs + base, tile, bytes);
    bulk_commit();
    bulk_wait_read0();
  }
}

// ---------------------------------------------------------------------------------------------- int
// int32 product with two's-complement wrap-around (what jnp gives in 32-bit mode)
__device__ __forceinline__ int mulw(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

__device__ __forceinline__ int shift_align(int v, int e_from, int e_to) {
  // fxpmodel.py:158-167: >> when the source exponent is larger, << otherwise (wrapping)
  return (e_from > e_to) ? (v >> (e_from - e_to)) : (int)((unsigned)v << (e_to - e_from));
}

// fxp_relu on a complex value (fxpmodel.py:30-45): the reference builds complex64, so both parts take a
// float32 round trip (round-to-nearest-even in, truncation with saturation out).
__device__ __forceinline__ void fxp_relu_complex(int& re, int& im) {
  const float fr = (float)re, fi = (float)im;
  const bool keep = (fr > 0.f) || (fr == 0.f && fi > 0.f);
  re = keep ? __float2int_rz(fr) : 0;
  im = keep ? __float2int_rz(fi) : 0;
}


template <bool INTERLEAVED, int DEPTH>
__global__ void __launch_bounds__(128) scan_i32_kernel(ScanI32Params q) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= q.lanes) return;
  const int p = (int)(g % q.P);
  const long long b = g / q.P;
  const int ar = q.A_re[p], ai = q.A_im[p];
  const size_t stride = (size_t)q.P;
  size_t idx = (size_t)b * q.L * q.P + p;
  int xr = 0, xi = 0;
  int br[DEPTH], bi[DEPTH];
  auto load = [&](int t, int& r, int& i) {
    if (t < q.L) {
      const size_t o = idx + (size_t)t * stride;
      if (INTERLEAVED) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(q.Bu_re) + o);
        r = v.x; i = v.y;
      } else {
        r = __ldg(q.Bu_re + o); i = __ldg(q.Bu_im + o);
      }
    }
  };
#pragma unroll
  for (int k = 0; k < DEPTH; ++k) load(k, br[k], bi[k]);
  for (int t0 = 0; t0 < q.L; t0 += DEPTH) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const int t = t0 + k;
      if (t < q.L) {
        const int cr = shift_align(br[k], q.eBu_re, q.ex_re);
        const int ci = shift_align(bi[k], q.eBu_im, q.ex_im);
        load(t + DEPTH, br[k], bi[k]);
        // fxpmodel.py:155-156: each product is floored on its own; products and sums wrap in int32
        const int nr = (int)((unsigned)(mulw(ar, xr) >> q.eA_re) - (unsigned)(mulw(ai, xi) >> q.eA_re) + (unsigned)cr);
        const int ni = (int)((unsigned)(mulw(ar, xi) >> q.eA_im) + (unsigned)(mulw(ai, xr) >> q.eA_im) + (unsigned)ci);
        xr = nr; xi = ni;
        int orr = xr, oi = xi;
        if (q.relu) fxp_relu_complex(orr, oi);
        const size_t o = idx + (size_t)t * stride;
        if (INTERLEAVED) {
          reinterpret_cast<int2*>(q.xs_re)[o] = make_int2(orr, oi);
        } else {
          q.xs_re[o] = orr; q.xs_im[o] = oi;
        }
      }
    }
  }
}

// Ring variant of the integer recurrence for throughput AND latency: one CTA per clip, one thread per state, the
// clip's [L][P] stream flows through a ring of kRingStages shared-memory tiles filled by 1-D bulk async copies
// (TMA engine) stages-2 tiles ahead (deep enough to cover DRAM + TMA latency) and drained by bulk stores, so the sequential dependency chain never waits
// on DRAM: time per clip ~ L x (one IMAD -> shift -> add chain).
constexpr int kRingMaxStages = 10;
constexpr int kRingSteps = 32;

template <bool INTERLEAVED, int CP>
__global__ void __launch_bounds__(256) scan_i32_ring_kernel(ScanI32Params q, int kRingStages) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kRingMaxStages];
  const int P = CP ? CP : q.P, L = q.L;  // compile-time row stride for the usual state sizes: immediate offsets
  const int p = threadIdx.x;
  const long long b = blockIdx.x;
  const size_t tile_elems = (size_t)kRingSteps * P;                 // states per tile
  const size_t stage_bytes = tile_elems * 8;
  int* tiles = reinterpret_cast<int*>(smem_raw);
  const int ntiles = (L + kRingSteps - 1) / kRingSteps;
  const size_t clip_off = (size_t)b * L * P;
  if (p == 0) {
    for (int s = 0; s < kRingStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue_load = [&](int t) {
    const int s = t % kRingStages;
    const int rows = min(kRingSteps, L - t * kRingSteps);
    const uint32_t bytes = (uint32_t)rows * P * 4u;
    unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
    const size_t off = clip_off + (size_t)t * tile_elems;
    mbar_expect_tx(&full[s], 2 * bytes);
    if (INTERLEAVED) {
      bulk_g2s(dst, reinterpret_cast<const int2*>(q.Bu_re) + off, 2 * bytes, &full[s]);
    } else {
      bulk_g2s(dst, q.Bu_re + off, bytes, &full[s]);
      bulk_g2s(dst + tile_elems * 4, q.Bu_im + off, bytes, &full[s]);
    }
  };
  if (p == 0)
    for (int t = 0; t < min(kRingStages - 2, ntiles); ++t) issue_load(t);
  const bool active = p < P;
  const int ar = active ? q.A_re[p] : 0, ai = active ? q.A_im[p] : 0;
  const int sl_re = max(0, q.ex_re - q.eBu_re), sr_re = max(0, q.eBu_re - q.ex_re);
  const int sl_im = max(0, q.ex_im - q.eBu_im), sr_im = max(0, q.eBu_im - q.ex_im);
  int xr = 0, xi = 0;
  for (int i = 0; i < ntiles; ++i) {
    const int s = i % kRingStages;
    if (p == 0) {
      const int j = i + kRingStages - 2;
      if (j < ntiles) {
        bulk_wait_read1();  // the store that last used stage j % S (iteration i-2) has released it
        issue_load(j);
      }
    }
    mbar_wait(&full[s], (uint32_t)(i / kRingStages) & 1u);
    const int rows = min(kRingSteps, L - i * kRingSteps);
    int* tre = reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes);
    int* tim = tre + tile_elems;
    if (active) {
      // The whole tile column of this state goes through registers: (1) load and exponent-align all kRingSteps
      // inputs (independent, pipelined), (2) run the dependent chain - per step only 4 IMAD, 4 shifts and the adds -
      // (3) apply the complex ReLU and store (independent again).  With one warp per scheduler nothing else hides
      // ALU latency, so everything that is not on the recurrence chain is kept out of the chain loop.
      int vr[kRingSteps], vi[kRingSteps];
      const int2* t2 = reinterpret_cast<const int2*>(tre) + p;
      const int* tr1 = tre + p;
      const int* ti1 = tim + p;
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        int br = 0, bi = 0;
        if (t < rows) {
          if (INTERLEAVED) {
            const int2 v = t2[t * P];
            br = v.x; bi = v.y;
          } else {
            br = tr1[t * P]; bi = ti1[t * P];
          }
        }
        // fxpmodel.py:158-167 without a branch: one of the two shift counts is zero
        vr[t] = (int)((unsigned)br << sl_re) >> sr_re;
        vi[t] = (int)((unsigned)bi << sl_im) >> sr_im;
      }
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        // fxpmodel.py:155-156: each product is floored on its own; products and sums wrap in int32
        const int nr = (int)((unsigned)(mulw(ar, xr) >> q.eA_re) - (unsigned)(mulw(ai, xi) >> q.eA_re) + (unsigned)vr[t]);
        const int ni = (int)((unsigned)(mulw(ar, xi) >> q.eA_im) + (unsigned)(mulw(ai, xr) >> q.eA_im) + (unsigned)vi[t]);
        xr = nr; xi = ni;
        vr[t] = nr; vi[t] = ni;
      }
      int2* o2 = reinterpret_cast<int2*>(tre) + p;
      int* or1 = tre + p;
      int* oi1 = tim + p;
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        if (t < rows) {
          int orr = vr[t], oi = vi[t];
          if (q.relu) {
            // |v| < 2^24 survives the float32 round trip of fxp_relu unchanged: integer compare, no conversions
            const bool small = ((unsigned)(orr + (1 << 24)) < (1u << 25)) && ((unsigned)(oi + (1 << 24)) < (1u << 25));
            if (small) {
              const bool keep = orr > 0 || (orr == 0 && oi > 0);
              orr = keep ? orr : 0; oi = keep ? oi : 0;
            } else {
              fxp_relu_complex(orr, oi);
            }
          }
          if (INTERLEAVED) o2[t * P] = make_int2(orr, oi);
          else { or1[t * P] = orr; oi1[t * P] = oi; }
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (p == 0) {
      const uint32_t bytes = (uint32_t)rows * P * 4u;
      const size_t off = clip_off + (size_t)i * tile_elems;
      if (INTERLEAVED) {
        bulk_s2g(reinterpret_cast<int2*>(q.xs_re) + off, tre, 2 * bytes);
      } else {
        bulk_s2g(q.xs_re + off, tre, bytes);
        bulk_s2g(q.xs_im + off, tim, bytes);
      }
      bulk_commit();
    }
  }
  if (p == 0) bulk_wait0();
}

// Warp-specialised form of the ring kernel (P a multiple of 32).  The recurrence is one dependent chain per
// state, and with one warp per scheduler nothing hides ALU latency, so every instruction that is not on the chain
// is moved to other warps.  Per 32 states there are three warps working on different tiles of the ring:
//   pre   : wait for the TMA load, align Bu to the state exponent in place          (fxpmodel.py:158-167)
//   chain : column -> registers, the kRingSteps dependent steps, raw states back     (fxpmodel.py:155-156, :169)
//   post  : complex ReLU in place (fxpmodel.py:30-45), bulk store, refill of the ring
// handing tiles over through per-(stage, column-half) mbarriers.
constexpr int kWsMaxHalves = 8;

template <bool INTERLEAVED>
__global__ void __launch_bounds__(384) scan_i32_ws_kernel(ScanI32Params q, int S, int npost) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kRingMaxStages];
  __shared__ __align__(8) uint64_t aligned[kRingMaxStages][kWsMaxHalves];
  __shared__ __align__(8) uint64_t scanned[kRingMaxStages][kWsMaxHalves];
  const int P = q.P, L = q.L;
  const int nh = P >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = warp / nh, h = warp % nh;   // 0 pre, 1 chain, 2 .. 1 + npost post (post warp k owns the 8-row batches k, k + npost, ..)
  const int p = h * 32 + lane;
  const long long b = blockIdx.x;
  const size_t tile_elems = (size_t)kRingSteps * P;
  const size_t stage_bytes = tile_elems * 8;
  const int ntiles = (L + kRingSteps - 1) / kRingSteps;
  const size_t clip_off = (size_t)b * L * P;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      for (int j = 0; j < nh; ++j) { mbar_init(&aligned[s][j], 1); mbar_init(&scanned[s][j], 1); }
    }
    fence_mbar_init();
  }
  __syncthreads();
  auto issue_load = [&](int t) {
    const int s = t % S;
    const int rows = min(kRingSteps, L - t * kRingSteps);
    const uint32_t bytes = (uint32_t)rows * P * 4u;
    unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
    const size_t off = clip_off + (size_t)t * tile_elems;
    mbar_expect_tx(&full[s], 2 * bytes);
    if (INTERLEAVED) {
      bulk_g2s(dst, reinterpret_cast<const int2*>(q.Bu_re) + off, 2 * bytes, &full[s]);
    } else {
      bulk_g2s(dst, q.Bu_re + off, bytes, &full[s]);
      bulk_g2s(dst + tile_elems * 4, q.Bu_im + off, bytes, &full[s]);
    }
  };
  const bool leader = (role == 2 && h == 0 && lane == 0);
  if (leader)
    for (int t = 0; t < min(S, ntiles); ++t) issue_load(t);
  const int sl_re = max(0, q.ex_re - q.eBu_re), sr_re = max(0, q.eBu_re - q.ex_re);
  const int sl_im = max(0, q.ex_im - q.eBu_im), sr_im = max(0, q.eBu_im - q.ex_im);
  const int ar = q.A_re[p], ai = q.A_im[p];
  int xr = 0, xi = 0;
  for (int i = 0; i < ntiles; ++i) {
    const int s = i % S;
    const uint32_t par = (uint32_t)(i / S) & 1u;
    const int rows = min(kRingSteps, L - i * kRingSteps);
    int* tre = reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes);
    int* tim = tre + tile_elems;
    int2* c2 = reinterpret_cast<int2*>(tre) + p;
    if (role == 0) {
      mbar_wait(&full[s], par);
      // eight rows are read before any is written back: one shared-memory round trip per batch, not per row
      // (rows past the end of the clip are stale ring contents; shifting them is harmless and nobody reads them)
#pragma unroll
      for (int t0 = 0; t0 < kRingSteps; t0 += 8) {
        if (t0 >= rows) break;
        int a[8], c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const size_t o = (size_t)(t0 + j) * P;
          if (INTERLEAVED) { const int2 v = c2[o]; a[j] = v.x; c[j] = v.y; }
          else { a[j] = tre[o + p]; c[j] = tim[o + p]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] = (int)((unsigned)a[j] << sl_re) >> sr_re;
          c[j] = (int)((unsigned)c[j] << sl_im) >> sr_im;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const size_t o = (size_t)(t0 + j) * P;
          if (INTERLEAVED) c2[o] = make_int2(a[j], c[j]);
          else { tre[o + p] = a[j]; tim[o + p] = c[j]; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&aligned[s][h]);
    } else if (role == 1) {
      mbar_wait(&aligned[s][h], par);
      int vr[kRingSteps], vi[kRingSteps];
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        // unpredicated: rows past the end of the clip (last tile only) hold stale ring contents; the steps computed
        // from them come after the clip's last state and are never stored to global memory
        if (INTERLEAVED) { const int2 v = c2[(size_t)t * P]; vr[t] = v.x; vi[t] = v.y; }
        else { vr[t] = tre[(size_t)t * P + p]; vi[t] = tim[(size_t)t * P + p]; }
      }
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        const int nr = (int)((unsigned)(mulw(ar, xr) >> q.eA_re) - (unsigned)(mulw(ai, xi) >> q.eA_re) + (unsigned)vr[t]);
        const int ni = (int)((unsigned)(mulw(ar, xi) >> q.eA_im) + (unsigned)(mulw(ai, xr) >> q.eA_im) + (unsigned)vi[t]);
        xr = nr; xi = ni;
        vr[t] = nr; vi[t] = ni;
      }
#pragma unroll
      for (int t = 0; t < kRingSteps; ++t) {
        if (INTERLEAVED) c2[(size_t)t * P] = make_int2(vr[t], vi[t]);
        else { tre[(size_t)t * P + p] = vr[t]; tim[(size_t)t * P + p] = vi[t]; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&scanned[s][h]);
    } else {
      mbar_wait(&scanned[s][h], par);
      if (q.relu) {
        for (int t0 = 8 * (role - 2); t0 < rows; t0 += 8 * npost) {
          int a[8], c[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const size_t o = (size_t)(t0 + j) * P;
            if (INTERLEAVED) { const int2 v = c2[o]; a[j] = v.x; c[j] = v.y; }
            else { a[j] = tre[o + p]; c[j] = tim[o + p]; }
          }
          // Branch-free: the lexicographic test on the float32 images equals the test on the integers themselves (int ->
          // float keeps sign and zero-ness); only the VALUES of kept entries with |v| >= 2^24 change in the float round
          // trip, and a whole batch takes that exact path when any lane saw one (warp vote, practically never).
          unsigned big = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            big |= ((unsigned)(a[j] + (1 << 24)) | (unsigned)(c[j] + (1 << 24))) >> 25;
            const bool keep = a[j] > 0 || (a[j] == 0 && c[j] > 0);
            a[j] = keep ? a[j] : 0; c[j] = keep ? c[j] : 0;
          }
          if (__any_sync(0xffffffffu, big != 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) fxp_relu_complex(a[j], c[j]);   // idempotent on the already ReLU-ed pair
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const size_t o = (size_t)(t0 + j) * P;
            if (INTERLEAVED) c2[o] = make_int2(a[j], c[j]);
            else { tre[o + p] = a[j]; tim[o + p] = c[j]; }
          }
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"r"(nh * 32 * npost) : "memory");   // all post warps of this tile
      if (leader) {
        const uint32_t bytes = (uint32_t)rows * P * 4u;
        const size_t off = clip_off + (size_t)i * tile_elems;
        if (INTERLEAVED) {
          bulk_s2g(reinterpret_cast<int2*>(q.xs_re) + off, tre, 2 * bytes);
        } else {
          bulk_s2g(q.xs_re + off, tre, bytes);
          bulk_s2g(q.xs_im + off, tim, bytes);
        }
        bulk_commit();
        // the store of tile i-1 has released its stage: refill it with tile i-1+S
        if (i >= 1 && i - 1 + S < ntiles) {
          bulk_wait_read1();
          issue_load(i - 1 + S);
        }
      }
    }
  }
  if (leader) bulk_wait0();
}

int launch_scan_i32(const ScanI32Params& q, bool interleaved, cudaStream_t st) {
  const long long B = q.lanes / q.P;
  const bool ptr_ok = aligned16(q.Bu_re) && aligned16(q.xs_re) && (interleaved || (aligned16(q.Bu_im) && aligned16(q.xs_im)));
  const bool ring_ok = ptr_ok && q.P <= 256 && (interleaved ? (q.P % 2 == 0) : (q.P % 4 == 0)) && B <= 0x7fffffffLL;
  ProfScope prof(S5_K_SCAN_I32, st);
  if (ring_ok) {
    // ring depth: ~8 tiles (6 in flight = 192 steps of look-ahead, about 2 us of dependent-chain time) when there are
    // few clips; shallower when many clips share an SM so that several CTAs stay resident
    const size_t tile = (size_t)kRingSteps * q.P * 8;
    int stages = B >= 2LL * sm_count() ? 4 : 8;
    while (stages > 3 && stages * tile > 200 * 1024) --stages;
    stages = std::min(stages, kRingMaxStages);
    const size_t smem = (size_t)stages * tile;
    if (q.P % 32 == 0 && q.P / 32 <= kWsMaxHalves && 3 * q.P <= 384) {
      auto kernel = interleaved ? scan_i32_ws_kernel<true> : scan_i32_ws_kernel<false>;
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // the post-processing (complex ReLU) is the slowest role with one warp per 32 states: up to four post warps
      // share a tile's rows when the thread budget allows (P = 64: 2 + 4 roles x 64 threads)
      const int npost = q.relu ? std::max(1, std::min(4, 384 / q.P - 2)) : 1;
      kernel<<<(unsigned)B, (2 + npost) * q.P, smem, st>>>(q, stages, npost);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
    const int threads = (q.P + 31) & ~31;
    auto go = [&](auto kernel) -> int {
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<(unsigned)B, threads, smem, st>>>(q, stages);
      return S5_OK;
    };
    int rc;
    if (interleaved) {
      if (q.P == 64) rc = go(scan_i32_ring_kernel<true, 64>);
      else if (q.P == 128) rc = go(scan_i32_ring_kernel<true, 128>);
      else rc = go(scan_i32_ring_kernel<true, 0>);
    } else {
      if (q.P == 64) rc = go(scan_i32_ring_kernel<false, 64>);
      else if (q.P == 128) rc = go(scan_i32_ring_kernel<false, 128>);
      else rc = go(scan_i32_ring_kernel<false, 0>);
    }
    if (rc) return rc;
    S5_LAUNCH_CHECK();
    return S5_OK;
  }
  const int threads = 128;
  const long long blocks = (q.lanes + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) S5_FAIL(S5_E_BADSHAPE, "scan_i32: too many lanes");
  if (interleaved) scan_i32_kernel<true, 8><<<(unsigned)blocks, threads, 0, st>>>(q);
  else scan_i32_kernel<false, 8><<<(unsigned)blocks, threads, 0, st>>>(q);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

int launch_scan_c64(const float* Bu, const float* lam, float* xs, int B, int L, int P, int relu, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  if (B <= 0 || L <= 0) S5_FAIL(S5_E_BADSHAPE, "scan_c64: B=%d L=%d must be positive", B, L);
  if (P < 2 || P > kScanThreads || (P & 1)) S5_FAIL(S5_E_BADSHAPE, "scan_c64: P=%d must be even and in [2, %d]", P, kScanThreads);
  if (!Bu || !lam || !xs || !workspace) S5_FAIL(S5_E_BADARG, "scan_c64: null pointer");
  if (!aligned16(Bu) || !aligned16(xs) || !aligned16(workspace)) S5_FAIL(S5_E_BADARG, "scan_c64: pointers must be 16-byte aligned");
  const size_t need = s5b200_scan_c64_workspace_bytes(B, L, P);
  if (workspace_bytes < need) S5_FAIL(S5_E_BADARG, "scan_c64: workspace %zu < %zu bytes", workspace_bytes, need);
  const ScanGeom g = scan_geom(L, P);
  static bool attr_set = false;
  if (!attr_set) {
    S5_CUDA(cudaFuncSetAttribute(scan_c64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    S5_CUDA(cudaFuncSetAttribute(scan_c64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  const size_t flags_bytes = round_up((size_t)B * g.nchunks * sizeof(uint32_t), 256);
  uint32_t* flags = reinterpret_cast<uint32_t*>(workspace);
  double2* recs = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(workspace) + flags_bytes);
  S5_CUDA(cudaMemsetAsync(flags, 0, flags_bytes, st));
  if (g.nchunks > 65535) S5_FAIL(S5_E_BADSHAPE, "scan_c64: L=%d needs more than 65535 chunks", L);
  dim3 grid(B, g.nchunks);
  ProfScope prof(S5_K_SCAN_C64, st);
  auto kernel = (long long)B * g.nchunks > 4LL * sm_count() ? scan_c64_kernel<true> : scan_c64_kernel<false>;
  kernel<<<grid, g.threads, g.smem_bytes, st>>>(reinterpret_cast<const float2*>(Bu), reinterpret_cast<const float2*>(lam),
                                                reinterpret_cast<float2*>(xs), L, P, g.nsub, g.nchunks, relu, flags, recs,
                                                (int)g.rec_floats2);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

}  // namespace s5

extern "C" size_t s5b200_scan_c64_workspace_bytes(int B, int L, int P) {
  if (B <= 0 || L <= 0 || P <= 0 || P > s5::kScanThreads) return 0;
  const s5::ScanGeom g = s5::scan_geom(L, P);
  return s5::round_up((size_t)B * g.nchunks * sizeof(uint32_t), 256) + (size_t)B * g.nchunks * g.rec_floats2 * sizeof(double2);
}

extern "C" int s5b200_scan_c64(const float* Bu, const float* lam, float* xs, int B, int L, int P, int relu,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return s5::launch_scan_c64(Bu, lam, xs, B, L, P, relu, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int s5b200_scan_i32(const int32_t* Bu_re, const int32_t* Bu_im, const int32_t* A_re, const int32_t* A_im,
                               int32_t* xs_re, int32_t* xs_im, int B, int L, int P, int eA_re, int eA_im, int eBu_re,
                               int eBu_im, int ex_re, int ex_im, int relu, void* stream) {
  if (B <= 0 || L <= 0 || P <= 0) S5_FAIL(S5_E_BADSHAPE, "scan_i32: B=%d L=%d P=%d must be positive", B, L, P);
  if (!Bu_re || !Bu_im || !A_re || !A_im || !xs_re || !xs_im) S5_FAIL(S5_E_BADARG, "scan_i32: null pointer");
  if (eA_re < 0 || eA_re > 31 || eA_im < 0 || eA_im > 31) S5_FAIL(S5_E_BADARG, "scan_i32: A exponent outside [0,31]");
  if (eBu_re - ex_re > 31 || ex_re - eBu_re > 31 || eBu_im - ex_im > 31 || ex_im - eBu_im > 31)
    S5_FAIL(S5_E_BADARG, "scan_i32: Bu/x exponent gap above 31");
  s5::ScanI32Params q;
  q.Bu_re = Bu_re; q.Bu_im = Bu_im; q.A_re = A_re; q.A_im = A_im; q.xs_re = xs_re; q.xs_im = xs_im;
  q.L = L; q.P = P; q.lanes = (long long)B * P;
  q.eA_re = eA_re; q.eA_im = eA_im; q.eBu_re = eBu_re; q.eBu_im = eBu_im; q.ex_re = ex_re; q.ex_im = ex_im; q.relu = relu;
  return s5::launch_scan_i32(q, false, (cudaStream_t)stream);
}
