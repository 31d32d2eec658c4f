// scan.cu - diagonal linear recurrences of the S5 SSM.
//
//  (1) scan_c64: x_k = lam * x_{k-1} + Bu_k in complex64, replacing jax.lax.associative_scan /
//      jax.lax.scan (reference sparseRNNs/model/ssm.py:127, :150, :313).  One CTA per (clip, time
//      segment), one thread per state; the segment streams through a ring of shared-memory tiles
//      filled by 1-D bulk async copies (TMA engine, UBLKCP) and drained by bulk stores.  Throughput
//      batches are a pure single-pass stream (98-103 % of the measured HBM copy peak at batch 4096);
//      few clips are cut into segments that exchange ONE carry each through a decoupled look-back
//      (segment aggregates + flags in a caller-provided workspace, Horner on lam^len in fp64).
//  (2) scan_i32: the integer recurrence of fxpmodel.py:147-208.  It floors every product separately, so it
//      is not associative and must be sequential in time; one lane per (clip, state), register
//      prefetch ring over time.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "scan.cuh"

namespace s5 {

constexpr int kScanThreads = 256;   // one thread per state: P <= 256
constexpr int kSegRows = 16;         // time steps per shared-memory tile (one bulk copy of 16 * P * 8 bytes)
constexpr int kSegMaxStages = 12;
constexpr int kSegMaxCtas = 1024;    // upper bound on resident CTAs the plan may assume (sizes the workspace)

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

// double-precision complex helpers: segment carries are combined in fp64 so that the powers lam^len do not
// accumulate an O(n * eps_f32) error (the states themselves stay fp32 like the reference).
__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 zfma(double2 a, double2 x, double2 v) {
  return make_double2(fma(a.x, x.x, fma(-a.y, x.y, v.x)), fma(a.x, x.y, fma(a.y, x.x, v.y)));
}
__device__ __forceinline__ double2 f2z(float2 v) { return make_double2((double)v.x, (double)v.y); }

// NUMERICS (fixed by L alone - never by the batch, the grid or timing).  A clip is divided into BLOCKS of `vlen` steps,
// vlen = scan_vlen(L).  Inside a block the recurrence is the sequential fp32 chain of the reference's lax.scan; the state
// that enters block k is float(c_k) with  c_k = lam^vlen c_(k-1) + agg_(k-1)  in fp64 (Horner), agg_j = the fp32 chain of
// block j started from zero.  Every scheduling mode below evaluates exactly this expression, so a clip's bits do not
// depend on how many clips are in the batch, on host chunking or on sharding over GPUs.
//
// SCHEDULING.  One CTA per (clip, segment), a segment being `gblocks` consecutive blocks.  Throughput batches (B >= 96)
// are not cut (one segment per clip): a pure stream.  Few clips are cut so that all resident CTA slots have work; the
// segments of a clip then exchange their block aggregates ONCE (decoupled look-back), not once per tile.
constexpr int kSegMaxBlocks = 512;   // blocks per clip (bounds the look-back and the workspace)
static int scan_vlen(int L) {       // block length: >= 128 steps, a multiple of the tile height, at most 512 blocks per clip
  const int per = (L + kSegMaxBlocks - 1) / kSegMaxBlocks;
  return std::max(128, ((per + kSegRows - 1) / kSegRows) * kSegRows);
}
struct ScanPlan {
  int nseg, gblocks, vlen, nblocks, stages, threads;
  size_t tile_bytes, smem_bytes;
};
static int scan_nseg_bound(int B, int L) {
  if (B >= 96) return 1;
  const int nblocks = (L + scan_vlen(L) - 1) / scan_vlen(L);
  return std::max(1, std::min(nblocks, kSegMaxCtas / B));
}
// flags of a launch: [B][nseg] for the segmented ring kernel, [B][groups of 64 states][nblocks] for the resident form
static size_t scan_flag_count(int B, int L, int P) {
  const size_t nseg = (size_t)scan_nseg_bound(B, L);
  if (nseg <= 1) return (size_t)B;
  const size_t nblocks = (size_t)((L + scan_vlen(L) - 1) / scan_vlen(L));
  return (size_t)B * std::max(nseg, nblocks * (size_t)((P + 31) / 32));
}
static ScanPlan scan_plan(int B, int L, int P, int resident_ctas) {
  ScanPlan g;
  g.threads = (P + 31) & ~31;
  g.tile_bytes = (size_t)kSegRows * P * 8;
  // Small rings, many CTAs per SM: one warp per 32 states runs a dependent chain and cannot hide its own latencies, so
  // an SM needs many such warps (P = 64: 8 KB tiles x 4 stages = 32 KB per CTA, 7 CTAs = 14 warps per SM).
  static const int env_stages = [] { const char* e = getenv("S5B200_SCAN_STAGES"); return e ? atoi(e) : 0; }();
  g.stages = env_stages > 0 ? std::max(3, std::min(kSegMaxStages, env_stages)) : 4;
  g.smem_bytes = (size_t)g.stages * g.tile_bytes;
  g.vlen = scan_vlen(L);
  g.nblocks = (L + g.vlen - 1) / g.vlen;
  // Few clips: cut until the resident CTA slots are full (at most one segment per block).
  int nseg = std::min(scan_nseg_bound(B, L), std::max(1, resident_ctas / std::max(1, B)));
  static const int forced = [] { const char* e = getenv("S5B200_SCAN_NSEG"); return e ? atoi(e) : 0; }();
  if (forced > 0) nseg = std::min(forced, scan_nseg_bound(B, L));
  g.gblocks = (g.nblocks + nseg - 1) / nseg;
  g.nseg = (g.nblocks + g.gblocks - 1) / g.gblocks;          // no empty segment
  return g;
}

// One CTA per (clip, segment), one thread per state, the segment streamed through a ring of shared-memory tiles:
// every tile [16][P] is contiguous in the [B][L][P] layout, so ONE 1-D bulk async copy (TMA engine, SASS UBLKCP) brings
// it in `stages - 2` tiles ahead of the recurrence and ONE bulk store takes the finished tile back; the warps issue no
// global load or store for the data.  A thread pulls its state's 16 values of a tile into registers (independent,
// conflict-free 8-byte shared loads), runs the dependent chains - the true states and, beside them, the block aggregate
// from zero: two independent FMA chains, the second one free on an HBM-bound kernel - and writes the states back in place.
//   nseg == 1 (throughput batches): a pure stream, 16 bytes of HBM traffic per state-step, no inter-CTA protocol at all.
//   nseg  > 1 (few clips): segment 0 streams once and publishes its block aggregates as it goes; an inner segment first
//   streams its rows READ-ONLY to get its block aggregates, publishes them, combines the aggregates of all earlier
//   blocks (ONE exchange per CTA - the first version paid a flag / record round trip per 8 KB tile and was latency
//   bound at 36 % of the HBM roofline for 32 clips), and streams its rows a second time with the true carry-in; the last
//   segment publishes nothing and only waits.  The second read of an inner segment comes out of L2 when the batch is
//   small enough to be cut (the ring keeps requesting tiles across the phase change, so the exchange overlaps the
//   prefetch).  All CTAs of a cut grid must be co-resident: cooperative launch.
// The first version of this kernel (one CTA per 8 KB chunk, load -> scan -> store in series) reached 51-63 % of the
// measured HBM peak at batch 4096; this one 98-103 %.
#ifdef S5_SCAN_TRACE
#include <cstdio>
__device__ __forceinline__ unsigned long long scan_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SCAN_T(i) do { if (p == 0) tr[i] = scan_gtime(); } while (0)
#else
#define SCAN_T(i)
#endif
template <int CP, bool RELU>
__global__ void __launch_bounds__(kScanThreads)
scan_c64_seg_kernel(const float2* __restrict__ Bu, const float2* __restrict__ lam, float2* xs, int L, int P_rt, int nseg,
                    int gblocks, int vlen, int nblocks, uint32_t* flags, double2* recs, int stages) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kSegMaxStages];
  const int P = CP ? CP : P_rt;                               // compile-time row stride for the usual state sizes
  const int p = threadIdx.x;
#ifdef S5_SCAN_TRACE
  unsigned long long tr[6] = {0, 0, 0, 0, 0, 0};
#endif
  SCAN_T(0);
  const bool active = p < P;
  const int b = blockIdx.x / nseg, sidx = blockIdx.x - b * nseg;
  const int blk0 = sidx * gblocks;                             // first block of this segment
  const int t0 = blk0 * vlen, rows_total = min(L, t0 + gblocks * vlen) - t0;
  const int ntiles = (rows_total + kSegRows - 1) / kSegRows;
  const int last_rows = rows_total - (ntiles - 1) * kSegRows;
  const int tiles_per_block = vlen / kSegRows;
  const size_t tile_elems = (size_t)kSegRows * P;
  const size_t base = ((size_t)b * L + t0) * P;
  float2* tiles = reinterpret_cast<float2*>(smem_raw);
  const bool two_pass = sidx > 0 && sidx + 1 < nseg;          // inner segment: aggregate pass, then the real one
  const bool publishes = nseg > 1 && sidx + 1 < nseg;         // its block aggregates are read by later segments
  const int njobs = two_pass ? 2 * ntiles : ntiles;           // job i streams tile i % ntiles

  if (p == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // a stage is reloaded two jobs after it was read (its store, if any, has had a whole job to read the tile out)
  const int ahead = stages - 2;
  // producer state (thread 0): next job to request, its tile and stage - counters instead of divisions in the loop
  int lj = 0, lt = 0, ls = 0;
  auto issue_load = [&]() {                                    // thread 0 only
    const int rows = lt + 1 == ntiles ? last_rows : kSegRows;
    const uint32_t bytes = (uint32_t)rows * P * 8u;
    mbar_expect_tx(&full[ls], bytes);
    bulk_g2s(tiles + ls * tile_elems, Bu + base + (size_t)lt * tile_elems, bytes, &full[ls]);
    ++lj;
    if (++lt == ntiles) lt = 0;
    if (++ls == stages) ls = 0;
  };
  if (p == 0)
    for (int i = 0; i < min(ahead, njobs); ++i) issue_load();

  const float2 a = active ? lam[p] : make_float2(0.f, 0.f);
  // lam^vlen in fp64 by binary powers
  double2 aV = make_double2(1., 0.);
  {
    double2 sq = f2z(a);
    for (int n = vlen; n > 0; n >>= 1) {
      if (n & 1) aV = zmul(aV, sq);
      sq = zmul(sq, sq);
    }
  }

  uint32_t* flag_row = flags + (size_t)b * nseg;
  double2* rec_row = recs + (size_t)b * nblocks * P;           // [block][P] aggregates of this clip
  auto raise_flag = [&]() {                                    // all threads: the aggregates of this segment are complete
    __syncthreads();                                           // the release below is cumulative over the barrier
    if (p == 0) st_release_u32(flag_row + sidx, 1u);
  };
  auto look_back = [&]() -> double2 {                          // c entering block blk0: Horner over the blocks 0 .. blk0 - 1
    // every thread polls its own share of the segment flags (all requests in flight together), then the CTA meets
    for (int j0 = p; j0 < sidx; j0 += 4 * (int)blockDim.x) {
      uint32_t f[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) f[w] = j0 + w * (int)blockDim.x < sidx ? ld_relaxed_u32(flag_row + j0 + w * blockDim.x) : 1u;
#pragma unroll
      for (int w = 0; w < 4; ++w)
        while (f[w] == 0u) f[w] = ld_relaxed_u32(flag_row + j0 + w * blockDim.x);
    }
    __syncthreads();
    fence_acquire_gpu();
    double2 c = make_double2(0., 0.);
    if (active) {
      for (int j0 = 0; j0 < blk0; j0 += 16) {                  // 16 records in flight per thread, Horner in order
        double2 r[16];
#pragma unroll
        for (int w = 0; w < 16; ++w)
          if (j0 + w < blk0) r[w] = __ldcg(rec_row + (size_t)(j0 + w) * P + p);
#pragma unroll
        for (int w = 0; w < 16; ++w)
          if (j0 + w < blk0) c = zfma(aV, c, r[w]);
      }
    }
    return c;
  };

  double2 c = make_double2(0., 0.);                            // fp64 carry at the start of the current block
  if (sidx > 0 && !two_pass) c = look_back();                  // last segment: nothing to publish, wait for the carry
  float2 x = make_float2((float)c.x, (float)c.y);              // true state
  float2 xa = make_float2(0.f, 0.f);                           // block aggregate (the chain from zero)

  int st = 0, tix = 0, tib = 0, blk = blk0;                    // consumer state: stage, tile, tile within block, block
  uint32_t par = 0;
  for (int i = 0; i < njobs; ++i) {
    const bool write = !two_pass || i >= ntiles;
    if (two_pass && i == ntiles) {                             // phase change of an inner segment (tiles keep arriving meanwhile)
      SCAN_T(2);
      raise_flag();
      c = look_back();
      x = make_float2((float)c.x, (float)c.y);
      xa = make_float2(0.f, 0.f);
      tib = 0; blk = blk0;
      SCAN_T(3);
    }
    if (i == 1) SCAN_T(1);
    if (p == 0 && lj < njobs) {
      bulk_wait_read1();                                       // the store that last used stage `ls` (job i - 2) has read it
      issue_load();
    }
    mbar_wait(&full[st], par);
    const int rows = tix + 1 == ntiles ? last_rows : kSegRows;
    float2* tile = tiles + st * tile_elems;
    if (active) {
      float2 v[kSegRows];
      const float2* src = tile + p;
      if (rows == kSegRows) {                                  // full tile: no predicates
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) v[t] = src[t * P];
        if (write) {
#pragma unroll
          for (int t = 0; t < kSegRows; ++t) {
            xa = cfma(a, xa, v[t]);
            x = cfma(a, x, v[t]);
            v[t] = x;
          }
#pragma unroll
          for (int t = 0; t < kSegRows; ++t) tile[t * P + p] = RELU ? relu_lex(v[t]) : v[t];
        } else {
#pragma unroll
          for (int t = 0; t < kSegRows; ++t) xa = cfma(a, xa, v[t]);
        }
      } else {
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) v[t] = t < rows ? src[t * P] : make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < kSegRows; ++t)
          if (t < rows) {
            xa = cfma(a, xa, v[t]);
            if (write) {
              x = cfma(a, x, v[t]);
              tile[t * P + p] = RELU ? relu_lex(x) : x;
            }
          }
      }
      if (++tib == tiles_per_block) {                          // block boundary: the next block starts from float(c)
        tib = 0;
        // an aggregate is published once: by the read-only pass of an inner segment, by the only pass of segment 0
        if (publishes && (!write || !two_pass)) __stcg(rec_row + (size_t)blk * P + p, f2z(xa));
        if (write) {
          c = zfma(aV, c, f2z(xa));
          x = make_float2((float)c.x, (float)c.y);
        }
        xa = make_float2(0.f, 0.f);
        ++blk;
      }
    }
    if (write) fence_proxy_async_smem();
    __syncthreads();                                           // every thread is done with the stage
    if (write && p == 0) {
      bulk_s2g(xs + base + (size_t)tix * tile_elems, tile, (uint32_t)rows * P * 8u);
      bulk_commit();
    }
    if (++st == stages) { st = 0; par ^= 1u; }
    if (++tix == ntiles) tix = 0;
  }
  SCAN_T(4);
  if (sidx == 0 && nseg > 1) raise_flag();                     // segment 0 published its aggregates while streaming
  if (p == 0) bulk_wait0();
  SCAN_T(5);
#ifdef S5_SCAN_TRACE
  if (p == 0 && b == 1 && (sidx < 3 || sidx + 2 >= nseg || sidx == nseg / 2))
    printf("[scan trace] seg %d/%d tiles %d: first tile %llu, phase1 end %llu, exchange %llu, phase2 %llu, drain %llu ns\n", sidx, nseg, ntiles,
           tr[1] - tr[0], tr[2] ? tr[2] - tr[0] : 0ull, tr[3] ? tr[3] - tr[2] : 0ull, tr[4] - (tr[3] ? tr[3] : tr[0]), tr[5] - tr[4]);
#endif
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


// ---- state-split variant for FEW clips.  One CTA per clip leaves most of the GPU idle when there are fewer clips than
// SMs (32 clips: 32 of 148), and each of those SMs then has to stream its clip's whole Bu / xs (3.8 MB at L = 3751, P = 64)
// through ONE SM's memory path (~50 GB/s): measured 127 us against ~36 us for the dependent chain.  The states of a clip
// are independent, so a CTA takes 32 of them: P / 32 CTAs per clip.  Its tile is a [32 steps][32 states] column block of
// the interleaved [B][L][P] array - a strided box, moved by 3-D TENSOR-MAP bulk copies (cp.async.bulk.tensor, dims
// {2P ints, L, B}): rows past the end of a clip are zero-filled on load and clipped on store, so the last short tile
// needs no special case.  Same three roles (pre-shift, chain, post / ReLU + store), one warp each, up to four post warps.
constexpr int kSplitStates = 32;
#ifndef S5_SPLIT_STEPS
#define S5_SPLIT_STEPS 64
#endif
constexpr int kSplitSteps = S5_SPLIT_STEPS;   // time steps per tile: the per-tile hand-over (mbarrier wait + arrive, ~200 cycles) is amortised over them
constexpr int kSplitMaxStages = 12;   // 16 KB per stage (input tile + output tile)
__device__ __forceinline__ void tma3d_load(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma3d_store(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
               : "memory");
}

// FEW CLIPS, RESIDENT FORM (round 2, last session).  The ring kernel above, cut into segments, reads an inner segment twice
// (aggregate pass, then the real one) through a ring that keeps only two 8 KB tiles in flight per CTA: at 32 clips x 3751
// frames it was latency bound (ncu: 0.54 waves of 64-thread CTAs, 9 % of the warp slots, DRAM 71 MB in / 15 MB out in 37 us).
// Here ONE block of `vlen` steps x (up to) 64 states is a CTA's whole job and stays RESIDENT in shared memory (64 KB, three
// CTAs per SM): all of its bulk copies are issued at once, the aggregate pass runs as the tiles land, the block publishes its
// aggregate, combines the aggregates of the earlier blocks of its clip (the same in-order fp64 Horner as above - every bit is
// the same as in the other scheduling modes), and the second pass runs out of shared memory and leaves by bulk stores tile by
// tile.  HBM sees every byte once in, once out.  The grid is block-major (all clips' block k before any block k + 1) and is
// NOT co-resident: a CTA waits only on lower block indices, which were dispatched before it and publish without waiting on
// anything (the usual decoupled look-back assumption of in-order CTA dispatch).  P > 64 is split into groups of 64 states per
// CTA.  A [16 steps][64 states] tile is a box of the [B][L][P] array and moves by 3-D TENSOR-MAP bulk copies (dims {2P floats, L, B}):
// one instruction per tile whatever P is (512-byte row copies made the P = 128 case slower than the ring kernel), rows past the
// end of a clip and states past P are zero-filled on load and clipped on store.
constexpr int kResG = 64;            // states per CTA (at most; 32 when that saves a wave of CTAs)
constexpr int kResMaxTiles = 16;     // vlen <= 256
template <bool RELU, int G>
__global__ void __launch_bounds__(G)
scan_c64_res_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, const float2* __restrict__ lam, int L,
                    int P, int B, int groups, int gw, int vlen, int nblocks, uint32_t* flags, double2* recs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kResMaxTiles];
  const int p = threadIdx.x;
  const int per = B * groups;
  const int sidx = blockIdx.x / per, rem = blockIdx.x - sidx * per;   // block-major
  const int b = rem / groups, g = rem - b * groups;
  const int gp = g * G + p;                                // global state index; gw = states per tile row = min(P, 64)
  const bool active = p < gw && gp < P;
  const int t0 = sidx * vlen, rows_total = min(L - t0, vlen);
  const int ntiles = (rows_total + kSegRows - 1) / kSegRows;
  float2* tiles = reinterpret_cast<float2*>(smem_raw);         // [rows][gw]
  const int c0 = 2 * g * G;                                // first float of the group within a time step
  const bool first = sidx == 0, last = sidx + 1 == nblocks;

  if (p == 0) {
    for (int j = 0; j < ntiles; ++j) mbar_init(&full[j], 1);
    fence_mbar_init();
    for (int j = 0; j < ntiles; ++j) {                         // the whole box always arrives (zero fill past L and past P)
      mbar_expect_tx(&full[j], (uint32_t)kSegRows * gw * 8u);
      tma3d_load(tiles + (size_t)j * kSegRows * gw, &tin, c0, t0 + j * kSegRows, b, &full[j]);
    }
  }
  __syncthreads();                                             // the barriers are initialised before anybody waits on them

  const float2 a = active ? lam[gp] : make_float2(0.f, 0.f);
  double2 aV = make_double2(1., 0.);
  {
    double2 sq = f2z(a);
    for (int n = vlen; n > 0; n >>= 1) {
      if (n & 1) aV = zmul(aV, sq);
      sq = zmul(sq, sq);
    }
  }
  uint32_t* flag_row = flags + ((size_t)b * groups + g) * nblocks;
  double2* rec_row = recs + (size_t)b * nblocks * P;           // [block][P] aggregates of this clip

  float2 xa = make_float2(0.f, 0.f);
  if (!first && !last) {                                       // aggregate pass (inner blocks are always full)
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&full[j], 0);
      if (active) {
        float2 v[kSegRows];
        const float2* src = tiles + (size_t)j * kSegRows * gw + p;
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) v[t] = src[t * gw];
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) xa = cfma(a, xa, v[t]);
      }
    }
    if (active) __stcg(rec_row + (size_t)sidx * P + gp, f2z(xa));
    __syncthreads();                                           // the release below is cumulative over the barrier
    if (p == 0) st_release_u32(flag_row + sidx, 1u);
  }
  double2 c = make_double2(0., 0.);
  if (!first) {                                                // carry entering this block: Horner over the blocks 0 .. sidx - 1
    for (int j0 = p; j0 < sidx; j0 += 4 * G) {
      uint32_t f[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) f[w] = j0 + w * G < sidx ? ld_relaxed_u32(flag_row + j0 + w * G) : 1u;
#pragma unroll
      for (int w = 0; w < 4; ++w)
        while (f[w] == 0u) f[w] = ld_relaxed_u32(flag_row + j0 + w * G);
    }
    __syncthreads();
    fence_acquire_gpu();
    if (active) {
      for (int j0 = 0; j0 < sidx; j0 += 16) {
        double2 r[16];
#pragma unroll
        for (int w = 0; w < 16; ++w)
          if (j0 + w < sidx) r[w] = __ldcg(rec_row + (size_t)(j0 + w) * P + gp);
#pragma unroll
        for (int w = 0; w < 16; ++w)
          if (j0 + w < sidx) c = zfma(aV, c, r[w]);
      }
    }
  }
  float2 x = make_float2((float)c.x, (float)c.y);
  for (int j = 0; j < ntiles; ++j) {
    mbar_wait(&full[j], 0);                                    // returns at once when the aggregate pass has been here already
    const int rows = min(kSegRows, rows_total - j * kSegRows);
    float2* tile = tiles + (size_t)j * kSegRows * gw;
    if (active) {
      float2 v[kSegRows];
      const float2* src = tile + p;
      if (rows == kSegRows) {
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) v[t] = src[t * gw];
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) {
          x = cfma(a, x, v[t]);
          v[t] = x;
        }
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) tile[t * gw + p] = RELU ? relu_lex(v[t]) : v[t];
      } else {
#pragma unroll
        for (int t = 0; t < kSegRows; ++t) v[t] = t < rows ? src[t * gw] : make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < kSegRows; ++t)
          if (t < rows) {
            x = cfma(a, x, v[t]);
            tile[t * gw + p] = RELU ? relu_lex(x) : x;
          }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (p == 0) {
      tma3d_store(&tout, c0, t0 + j * kSegRows, b, tile);      // rows past L and states past P are clipped
      bulk_commit();
    }
  }
  if (first && !last) {                                        // block 0: its states from zero ARE its aggregate
    if (active) __stcg(rec_row + gp, f2z(x));
    __syncthreads();
    if (p == 0) st_release_u32(flag_row, 1u);
  }
  bulk_wait_read0();                                           // the stores have read the tile (shared memory goes away at exit)
}


#ifdef S5_SCAN_TRACE
#include <cstdio>
#define SPLIT_TRACE_DECL long long tr_wait = 0; const long long tr_t0 = clock64();
#define SPLIT_WAIT(stmt) do { const long long t_ = clock64(); stmt; tr_wait += clock64() - t_; } while (0)
#define SPLIT_TRACE_PRINT if (blockIdx.x == 5 && lane == 0) printf("[split role %d] total %lld waiting %lld cycles, %d tiles\n", role, clock64() - tr_t0, tr_wait, ntiles);
#else
#define SPLIT_TRACE_DECL
#define SPLIT_WAIT(stmt) stmt
#define SPLIT_TRACE_PRINT
#endif
// IL: interleaved (re, im) pairs in one array (the model's layout; tin / tout, box of 64 ints x steps); else planar arrays as at
// the s5b200_scan_i32 ABI (tin / tout = real parts, tin2 / tout2 = imaginary parts, boxes of 32 ints x steps, two per tile)
template <bool IL>
__global__ void __launch_bounds__(192) scan_i32_split_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                                                             const __grid_constant__ CUtensorMap tin2, const __grid_constant__ CUtensorMap tout2,
                                                             ScanI32Params q, int S, int npost, int nsplit) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kSplitMaxStages], aligned[kSplitMaxStages], scanned[kSplitMaxStages];
  const int L = q.L;
  // Roles: 0 pre, 1 chain, 2 .. 1 + npost post.  Warp w issues on scheduler w % 4: the chain warp (its ~12 dependent
  // ALU / FMA instructions per step are the critical path of the kernel) is warp 2, alone on its scheduler with up to six warps
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = wid == 2 ? 1 : (wid == 1 ? 2 : wid);    // warps 0, 1, 2, 3.. = pre, post 0, chain, post 1..
  const int b = blockIdx.x / nsplit, sub = blockIdx.x - b * nsplit;
  const int p = sub * kSplitStates + lane;
  // A stage is an input tile and a separate output tile: the chain warp loads its 32 inputs from one and stores its 32 states
  // to the other.  In place, the loads of later steps queue behind the stores of earlier ones: measured 36 cycles per step
  // against 20-22 for the bare dependent chain (profiles/scripts/chain_rate.cu).
  constexpr uint32_t tile_bytes = kSplitSteps * kSplitStates * 8, stage_bytes = 2 * tile_bytes;
  const int ntiles = (L + kSplitSteps - 1) / kSplitSteps;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&aligned[s], 1); mbar_init(&scanned[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  auto issue_load = [&](int t, int st) {                      // tile t into stage st (= t % S)
    mbar_expect_tx(&full[st], tile_bytes);
    if (IL) {
      tma3d_load(smem_raw + (size_t)st * stage_bytes, &tin, 2 * kSplitStates * sub, t * kSplitSteps, b, &full[st]);
    } else {
      tma3d_load(smem_raw + (size_t)st * stage_bytes, &tin, kSplitStates * sub, t * kSplitSteps, b, &full[st]);
      tma3d_load(smem_raw + (size_t)st * stage_bytes + tile_bytes / 2, &tin2, kSplitStates * sub, t * kSplitSteps, b, &full[st]);
    }
  };
  const bool leader = role == 2 && lane == 0;
  if (leader)
    for (int t = 0; t < min(S, ntiles); ++t) issue_load(t, t);
  const int sl_re = max(0, q.ex_re - q.eBu_re), sr_re = max(0, q.eBu_re - q.ex_re);
  const int sl_im = max(0, q.ex_im - q.eBu_im), sr_im = max(0, q.eBu_im - q.ex_im);
  const int ar = q.A_re[p], ai = q.A_im[p];
  int xr = 0, xi = 0;
  SPLIT_TRACE_DECL
  int s = 0;                                                  // stage and parity advance incrementally: S is a run-time
  uint32_t par = 0;                                           // value, and i % S, i / S cost ~40 instructions per tile and role
  for (int i = 0; i < ntiles; ++i, s = (s + 1 == S ? 0 : s + 1), par ^= (s == 0 ? 1u : 0u)) {
    const int rows = min(kSplitSteps, L - i * kSplitSteps);
    int* cbase = reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes);       // input tile, then the output tile
    int* obase = reinterpret_cast<int*>(smem_raw + (size_t)s * stage_bytes + tile_bytes);
    constexpr int kPlane = kSplitSteps * kSplitStates;        // ints per planar half-tile
    auto ld2 = [&](const int* base, int t) -> int2 {
      if (IL) return reinterpret_cast<const int2*>(base)[t * kSplitStates + lane];
      return make_int2(base[t * kSplitStates + lane], base[kPlane + t * kSplitStates + lane]);
    };
    auto st2 = [&](int* base, int t, int2 v) {
      if (IL) reinterpret_cast<int2*>(base)[t * kSplitStates + lane] = v;
      else { base[t * kSplitStates + lane] = v.x; base[kPlane + t * kSplitStates + lane] = v.y; }
    };
    if (role == 0) {
      SPLIT_WAIT(mbar_wait(&full[s], par));
#pragma unroll
      for (int t0 = 0; t0 < kSplitSteps; t0 += 8) {
        int2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ld2(cbase, t0 + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j].x = (int)((unsigned)v[j].x << sl_re) >> sr_re;
          v[j].y = (int)((unsigned)v[j].y << sl_im) >> sr_im;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) st2(cbase, t0 + j, v[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&aligned[s]);
    } else if (role == 1) {
      SPLIT_WAIT(mbar_wait(&aligned[s], par));
      int vr[kSplitSteps], vi[kSplitSteps];
#pragma unroll
      for (int t = 0; t < kSplitSteps; ++t) { const int2 v = ld2(cbase, t); vr[t] = v.x; vi[t] = v.y; }
#pragma unroll
      for (int t = 0; t < kSplitSteps; ++t) {
        const int nr = (int)((unsigned)(mulw(ar, xr) >> q.eA_re) - (unsigned)(mulw(ai, xi) >> q.eA_re) + (unsigned)vr[t]);
        const int ni = (int)((unsigned)(mulw(ar, xi) >> q.eA_im) + (unsigned)(mulw(ai, xr) >> q.eA_im) + (unsigned)vi[t]);
        xr = nr; xi = ni;
        st2(obase, t, make_int2(nr, ni));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&scanned[s]);
    } else {
      SPLIT_WAIT(mbar_wait(&scanned[s], par));
      if (q.relu) {
        for (int t0 = 8 * (role - 2); t0 < rows; t0 += 8 * npost) {
          int a[8], c[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { const int2 v = ld2(obase, t0 + j); a[j] = v.x; c[j] = v.y; }
          unsigned big = 0;   // see scan_i32_ws_kernel: only kept values with |v| >= 2^24 change in the float round trip
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            big |= ((unsigned)(a[j] + (1 << 24)) | (unsigned)(c[j] + (1 << 24))) >> 25;
            const bool keep = a[j] > 0 || (a[j] == 0 && c[j] > 0);
            a[j] = keep ? a[j] : 0; c[j] = keep ? c[j] : 0;
          }
          if (__any_sync(0xffffffffu, big != 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) fxp_relu_complex(a[j], c[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) st2(obase, t0 + j, make_int2(a[j], c[j]));
        }
      }
      fence_proxy_async_smem();
      SPLIT_WAIT(asm volatile("bar.sync 1, %0;" ::"r"(32 * npost) : "memory"));
      if (leader) {
        if (IL) {
          tma3d_store(&tout, 2 * kSplitStates * sub, i * kSplitSteps, b, obase);
        } else {
          tma3d_store(&tout, kSplitStates * sub, i * kSplitSteps, b, obase);
          tma3d_store(&tout2, kSplitStates * sub, i * kSplitSteps, b, obase + kPlane);
        }
        bulk_commit();
        if (i >= 1 && i - 1 + S < ntiles) {                    // the store of tile i-1 has released its stage
          SPLIT_WAIT(bulk_wait_read1());
          issue_load(i - 1 + S, s == 0 ? S - 1 : s - 1);
        }
      }
    }
  }
  SPLIT_TRACE_PRINT
  if (leader) bulk_wait0();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links the static runtime only)
static int make_tmap3d_i32(CUtensorMap* tm, const void* base, int P, int L, long long B, bool interleaved) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeFn>(f);
  }();
  if (!fn) return -1;
  const cuuint64_t W = interleaved ? 2ull * P : (cuuint64_t)P;   // ints per time step of one array
  const cuuint64_t dims[3] = {W, (cuuint64_t)L, (cuuint64_t)B};
  const cuuint64_t strides[2] = {W * 4, (cuuint64_t)L * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(interleaved ? 2 * kSplitStates : kSplitStates), kSplitSteps, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

// [B][L][P] complex64 as a 3-D tensor of floats {2P, L, B}; box = {2 * gw floats, kSegRows steps, 1 clip}
static int make_tmap3d_c64(CUtensorMap* tm, const void* base, int P, int L, long long B, int gw) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeFn>(f);
  }();
  if (!fn) return -1;
  const cuuint64_t W = 2ull * P;
  const cuuint64_t dims[3] = {W, (cuuint64_t)L, (cuuint64_t)B};
  const cuuint64_t strides[2] = {W * 4, (cuuint64_t)L * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * gw), (cuuint32_t)kSegRows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

int launch_scan_i32(const ScanI32Params& q, bool interleaved, cudaStream_t st) {
  const long long B = q.lanes / q.P;
  const bool ptr_ok = aligned16(q.Bu_re) && aligned16(q.xs_re) && (interleaved || (aligned16(q.Bu_im) && aligned16(q.xs_im)));
  const bool ring_ok = ptr_ok && q.P <= 256 && (interleaved ? (q.P % 2 == 0) : (q.P % 4 == 0)) && B <= 0x7fffffffLL;
  ProfScope prof(S5_K_SCAN_I32, st);
  // fewer clips than SMs: 32 states per CTA (S5B200_SCAN_I32_SPLIT=0: one CTA per clip)
  static const bool split_on = [] { const char* e = getenv("S5B200_SCAN_I32_SPLIT"); return !e || atoi(e) != 0; }();
  if (split_on && ring_ok && q.P % kSplitStates == 0 && q.P > kSplitStates && B < sm_count() && q.L < (1 << 30)) {
    CUtensorMap tin, tout, tin2, tout2;
    bool ok = make_tmap3d_i32(&tin, q.Bu_re, q.P, q.L, B, interleaved) == 0 && make_tmap3d_i32(&tout, q.xs_re, q.P, q.L, B, interleaved) == 0;
    if (ok && !interleaved) ok = make_tmap3d_i32(&tin2, q.Bu_im, q.P, q.L, B, false) == 0 && make_tmap3d_i32(&tout2, q.xs_im, q.P, q.L, B, false) == 0;
    if (ok) {
      if (interleaved) { tin2 = tin; tout2 = tout; }
      static const int split_stages = [] { const char* e = getenv("S5B200_SCAN_I32_STAGES"); return e ? std::max(3, std::min(kSplitMaxStages, atoi(e))) : (kSplitSteps >= 64 ? 6 : 10); }();
      const int nsplit = q.P / kSplitStates, npost = q.relu ? 4 : 1, stages = split_stages;
      const size_t smem = (size_t)stages * 2 * kSplitSteps * kSplitStates * 8;   // input + output tile per stage
      auto kernel = interleaved ? scan_i32_split_kernel<true> : scan_i32_split_kernel<false>;
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<(unsigned)(B * nsplit), 32 * (2 + npost), smem, st>>>(tin, tout, tin2, tout2, q, stages, npost, nsplit);
      S5_LAUNCH_CHECK();
      return S5_OK;
    }
  }
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
  // resident CTA slots for this tile size (the cut grid must be co-resident)
  using Kern = void (*)(const float2*, const float2*, float2*, int, int, int, int, int, int, uint32_t*, double2*, int);
  Kern kern;
  if (P == 64) kern = relu ? scan_c64_seg_kernel<64, true> : scan_c64_seg_kernel<64, false>;
  else if (P == 128) kern = relu ? scan_c64_seg_kernel<128, true> : scan_c64_seg_kernel<128, false>;
  else kern = relu ? scan_c64_seg_kernel<0, true> : scan_c64_seg_kernel<0, false>;
  ScanPlan g = scan_plan(B, L, P, 1);
  S5_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  int per_sm = 0;
  S5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.threads, g.smem_bytes));
  if (per_sm < 1) S5_FAIL(S5_E_CUDA, "scan_c64: no CTA fits an SM (%zu bytes of shared memory)", g.smem_bytes);
  g = scan_plan(B, L, P, std::min(kSegMaxCtas, per_sm * sm_count()));
  const size_t flags_bytes = round_up(scan_flag_count(B, L, P) * sizeof(uint32_t), 256);
  uint32_t* flags = reinterpret_cast<uint32_t*>(workspace);
  double2* recs = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(workspace) + flags_bytes);
  // Few clips: one resident block per CTA when a block fits shared memory (S5B200_SCAN_RESIDENT=0: the segmented ring kernel)
  const char* res_env = getenv("S5B200_SCAN_RESIDENT");
  // Measured (profiles/r02_scan_resident.md): 5-14 % faster than the ring kernel for P <= 64 and up to ~64 blocks per clip, slower for
  // very long single clips (every CTA walks the whole look-back); S5B200_SCAN_RESIDENT=1 forces it.
  // wider states: at par with the ring kernel (66.5 against 65.6 us at 32 x 3751 x 128), ahead of it while the grid is one wave
  const bool res_auto = g.nblocks <= 64 && (P <= kResG || (long long)B * ((P + kResG - 1) / kResG) * g.nblocks <= 3LL * sm_count());
  CUtensorMap tin, tout;
  // 32 states per CTA (32 KB blocks, seven CTAs per SM) were measured too: no faster than 64 at any shape (profiles/r02_scan_resident.md)
  constexpr int G = kResG;
  const int gw = std::min(P, G);
  if (g.nseg > 1 && g.vlen <= kResMaxTiles * kSegRows && (res_env ? res_env[0] != '0' : res_auto) && L < (1 << 30) &&
      make_tmap3d_c64(&tin, Bu, P, L, B, gw) == 0 && make_tmap3d_c64(&tout, xs, P, L, B, gw) == 0) {
    const int groups = (P + G - 1) / G;
    const long long grid = (long long)g.nblocks * B * groups;
    if (grid > 0x7fffffffLL) S5_FAIL(S5_E_BADSHAPE, "scan_c64: too many clips");
    const size_t smem = (size_t)g.vlen * gw * 8;
    using ResKern = void (*)(const CUtensorMap, const CUtensorMap, const float2*, int, int, int, int, int, int, int, uint32_t*, double2*);
    ResKern rk = relu ? scan_c64_res_kernel<true, G> : scan_c64_res_kernel<false, G>;
    S5_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    S5_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * groups * g.nblocks * sizeof(uint32_t), st));
    ProfScope prof(S5_K_SCAN_C64, st);
    rk<<<(unsigned)grid, G, smem, st>>>(tin, tout, reinterpret_cast<const float2*>(lam), L, P, B, groups, gw, g.vlen, g.nblocks, flags, recs);
    S5_LAUNCH_CHECK();
    return S5_OK;
  }
  if (g.nseg > 1) S5_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * g.nseg * sizeof(uint32_t), st));
  if ((long long)B * g.nseg > 0x7fffffffLL) S5_FAIL(S5_E_BADSHAPE, "scan_c64: too many clips");
  ProfScope prof(S5_K_SCAN_C64, st);
  const float2* bu2 = reinterpret_cast<const float2*>(Bu);
  const float2* lam2 = reinterpret_cast<const float2*>(lam);
  float2* xs2 = reinterpret_cast<float2*>(xs);
  if (g.nseg > 1) {
    void* args[] = {&bu2, &lam2, &xs2, &L, &P, &g.nseg, &g.gblocks, &g.vlen, &g.nblocks, &flags, &recs, &g.stages};
    S5_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(B * g.nseg), dim3(g.threads), args, g.smem_bytes, st));
  } else {
    kern<<<B, g.threads, g.smem_bytes, st>>>(bu2, lam2, xs2, L, P, 1, g.gblocks, g.vlen, g.nblocks, flags, recs, g.stages);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}

}  // namespace s5

extern "C" size_t s5b200_scan_c64_workspace_bytes(int B, int L, int P) {
  if (B <= 0 || L <= 0 || P <= 0 || P > s5::kScanThreads) return 0;
  const size_t nseg = (size_t)s5::scan_nseg_bound(B, L);      // flags + block aggregates [B][nblocks][P] double2 (cut grids only)
  const size_t nblocks = nseg > 1 ? (size_t)((L + s5::scan_vlen(L) - 1) / s5::scan_vlen(L)) : 0;
  return s5::round_up(s5::scan_flag_count(B, L, P) * sizeof(uint32_t), 256) + (size_t)B * nblocks * P * sizeof(double2);
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
