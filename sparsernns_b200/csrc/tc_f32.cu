// tc_f32.cu - dense-tile projections of the FLOAT flavour on the tcgen05 tensor cores (kind::tf32, 3 x TF32).
//
// The fp32 model (quantization = none: ssm.py:84-185, layers.py:180-241) multiplies arbitrary fp32 activations with
// arbitrary fp32 weights, so the exact-integer trick of the static flavour does not apply.  Each operand is split into
// hi = x rounded to the nearest tf32 number (19 bits) and lo = x - hi (exact in fp32; the tensor core reads its top 19 bits),
// and   x w  ~=  x_hi w_hi + x_lo w_hi + x_hi w_lo   is accumulated in fp32 in TMEM: the dropped terms are
// <= 2^-21 |x w| per product, the same order as the rounding of an fp32 dot product (tolerance of the flavour: 1e-5).
//
// TRANSPOSED formulation: D[n][row] = sum_k W[n][k] X[row][k] - the weights are the A operand (M = 128 output
// columns per M-block), a 128-row tile of activations the B operand.  The accumulator then has TMEM lane = output
// column, TMEM column = row, and the epilogue writes Y[row][n0 + lane]: 32 lanes store 32 consecutive floats, so
// results go from registers to global memory coalesced, with no shared-memory transpose (257-float rows included).
// The epilogue also does what follows the product in the reference, with coalesced side reads:
//   EPI_BIAS : + bias, optional ReLU            (encoder / decoder Dense, B projection, s5b200_spmm_f32)
//   EPI_SSM  : y = conj_mul acc + D u, u = BN(x); x1 = activation(y)                  (ssm.py:179-185, 618-622; layers.py:217)
//   EPI_GATE : out = x1 sigmoid(acc + bias) + skip, optional ReLU                     (layers.py:223-238)
// K is processed in chunks of 16 columns: the chunk's hi / lo planes of X and of W (streamed from L2 by one bulk copy,
// pre-arranged in the operand layout by the host) sit in one of two buffers; the tensor core accumulates chunk after
// chunk while the warps stage the next one.  Two CTAs of 8 warps share an SM when the tile has <= 256 output columns.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "tc.cuh"
#include "tc_f32.cuh"

namespace s5 {

constexpr int kF32Threads = 256;
constexpr int kF32Rows = 128;
constexpr int kF32KC = 16;                                    // fp32 columns per K chunk
constexpr uint32_t kF32Lbo = kF32Rows * 16 + 32;              // padded K-chunk stride of a [128][.] operand tile: with 32 bytes
                                                              // both staging patterns below are bank-conflict free
constexpr uint32_t kF32Plane = (kF32KC / 4) * kF32Lbo;        // bytes of one [128][16] fp32 plane

enum { EPI_BIAS = 0, EPI_SSM = 1, EPI_GATE = 2 };

struct TcF32Params {
  const float* X; long long rows; int K, x_stride;
  const float* in_sub; const float* in_mul; const float* in_add;   // BatchNorm on the input columns (nullable)
  const unsigned char* Wg; int N, MB, NC;                          // weights: [NC][MB][hi, lo] planes
  float* Y; int y_stride;
  int epi; const float* bias; int relu;
  const float* X2; const float* bn_sub; const float* bn_mul; const float* bn_add; const float* D; float conj_mul; int act;
  const float* X1; const float* SKIP;
};

__device__ __forceinline__ uint32_t tc_idesc_tf32(int M, int N) {   // A = B = TF32 (format 2), D = F32, K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// logistic function from the two SFU approximations (ex2: 2 ulp, rcp: 1 ulp - a few 1e-7 relative, the flavour's tolerance is 1e-5);
// exp overflow gives rcp(inf) = 0, underflow gives 1
__device__ __forceinline__ float f32_sigmoid_fast(float z) {
  float ex, sg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex));
  return sg;
}
__device__ __forceinline__ float f32_gelu_tanh(float x) {     // jax.nn.gelu(approximate=True), layers.py:116-124
  // 0.5 (1 + tanh z) = sigmoid(2 z), z = c (x + 0.044715 x^3): no tanhf (~50 instructions with its branch)
  const float k0 = -2.0f * 0.7978845608028654f * 1.4426950408889634f, k1 = k0 * 0.044715f;
  float ex, sg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x * fmaf(x * x, k1, k0)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.0f + ex));
  return x * sg;
}

// VEC4: input rows are 16-byte aligned and K % 4 == 0 - a lane stages four consecutive columns (one 16-byte load, one
// 16-byte store per plane) instead of one.
template <bool VEC4>
__global__ void __launch_bounds__(kF32Threads, 2) tc_f32_kernel(const __grid_constant__ TcF32Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int MB = p.MB, NC = p.NC;
  const uint32_t wbytes = (uint32_t)MB * 2u * kF32Plane;      // one chunk of W (all M blocks, hi + lo)
  const uint32_t xbytes = 2u * kF32Plane;                     // hi / lo planes of one X chunk
  unsigned char* Wring = smem;                                // three W slots: chunk g + 1 is requested while chunk g is staged
  unsigned char* Xbuf = smem + 3u * wbytes;                   // two X buffers
  __shared__ __align__(8) uint64_t full[3], mdone[2], tdone;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // The tensor core truncates when it adds a product group to the running fp32 sum (a bias that grows with the number of
  // accumulation steps), so with one M block there are TWO accumulators, one for the large hi x hi products and one for
  // the small cross terms, added in fp32 (round to nearest) in the epilogue: a third of the steps on the accumulator
  // that matters, still 256 TMEM columns = two CTAs per SM.
  const int NACC = MB <= 2 ? 2 : 1;                           // MB == 2 runs one CTA per SM (shared memory): all 512 columns are its own
  const uint32_t tmem_cols = MB == 1 ? 256u : 512u;
  if (warp == 0) tc_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&full[2], 1); mbar_init(&mdone[0], 1); mbar_init(&mdone[1], 1); mbar_init(&tdone, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t taddr = tmem_slot;
  const long long row_begin = p.rows * blockIdx.x / gridDim.x, row_end = p.rows * (blockIdx.x + 1) / gridDim.x;
  const long long last_row = p.rows - 1;
  // staging, scalar form: half a warp per row (16 columns of the chunk), 8 warps x 2 rows per pass, 8 passes per tile
  // VEC4 form: four lanes per row, 8 rows per warp, 64 rows per pass, 2 passes per tile
  const int srow = VEC4 ? (tid >> 2) : 2 * warp + (lane >> 4), l16 = VEC4 ? 4 * (lane & 3) : (lane & 15);
  const uint32_t soff = (uint32_t)(l16 >> 2) * kF32Lbo + 4u * (l16 & 3);
  constexpr int NV = VEC4 ? 2 : 8;                            // loads per lane and chunk
  constexpr int RSTEP = VEC4 ? 64 : 16;
  const uint32_t idesc = tc_idesc_tf32(kF32Rows, kF32Rows);
  const int q4 = warp & 3, grp = warp >> 2;                   // epilogue: TMEM lane quadrant, item group
  long long g = 0;                                            // running chunk counter (buffers alternate across tiles)
  const long long total_chunks = (row_end - row_begin + kF32Rows - 1) / kF32Rows * NC;
  uint32_t tile_par = 0;
  float4 v[NV];                                               // the NEXT chunk's raw values travel while the current one is multiplied (scalar form: .x only)
  auto load_chunk = [&](long long r0, int c) {
    const int kc = min(c * kF32KC + l16, p.K - (VEC4 ? 4 : 1));   // pad columns meet zero weights
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const float* src = p.X + (size_t)min(r0 + RSTEP * q + srow, last_row) * p.x_stride + kc;
      if constexpr (VEC4) v[q] = *reinterpret_cast<const float4*>(src);
      else v[q].x = *src;
    }
  };
  if (row_begin < row_end) load_chunk(row_begin, 0);
  for (long long row0 = row_begin; row0 < row_end; row0 += kF32Rows) {
    const int rv = (int)min((long long)kF32Rows, row_end - row0);
    for (int c = 0; c < NC; ++c, ++g) {
      const int b = (int)(g & 1);
      if (g >= 2) mbar_wait(&mdone[b], (uint32_t)((g - 2) >> 1) & 1u);   // the MMAs of chunk g - 2 are done: X buffer b and W slot (g + 1) % 3 are free
      if (tid == 0) {
        auto request = [&](long long gg) {                    // W chunk of running chunk gg -> slot gg % 3
          const int s3 = (int)(gg % 3);
          mbar_expect_tx(&full[s3], wbytes);
          bulk_g2s(Wring + (size_t)s3 * wbytes, p.Wg + (size_t)(gg % NC) * wbytes, wbytes, &full[s3]);
        };
        if (g == 0) request(0);
        if (g + 1 < total_chunks) request(g + 1);
      }
      const int k = c * kF32KC + l16;
      unsigned char* xh = Xbuf + (size_t)b * xbytes;
      if constexpr (VEC4) {
        float bsub[4], bmul[4], badd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool in = p.in_sub && k + e < p.K;
          const int ke = min(k + e, p.K - 1);
          bsub[e] = in ? p.in_sub[ke] : 0.f; bmul[e] = in ? p.in_mul[ke] : 1.f; badd[e] = in ? p.in_add[ke] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          float x[4] = {v[q].x, v[q].y, v[q].z, v[q].w}, hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.in_sub) x[e] = (x[e] - bsub[e]) * bmul[e] + badd[e];
            hi[e] = __uint_as_float((__float_as_uint(x[e]) + 0x1000u) & 0xFFFFE000u);   // nearest tf32: |lo| <= 2^-12 |x|
            lo[e] = x[e] - hi[e];
          }
          const uint32_t o = soff + (uint32_t)(RSTEP * q + srow) * 16u;
          *reinterpret_cast<float4*>(xh + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(xh + kF32Plane + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else {
        const int kc = min(k, p.K - 1);
        const float bsub = (p.in_sub && k < p.K) ? p.in_sub[kc] : 0.f;
        const float bmul = (p.in_sub && k < p.K) ? p.in_mul[kc] : 1.f;
        const float badd = (p.in_sub && k < p.K) ? p.in_add[kc] : 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          float x = v[q].x;
          if (p.in_sub) x = (x - bsub) * bmul + badd;
          const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);   // nearest tf32: |lo| <= 2^-12 |x|
          const uint32_t o = soff + (uint32_t)(RSTEP * q + srow) * 16u;
          *reinterpret_cast<float*>(xh + o) = hi;
          *reinterpret_cast<float*>(xh + kF32Plane + o) = x - hi;
        }
      }
      if (c + 1 < NC) load_chunk(row0, c + 1);
      else if (row0 + kF32Rows < row_end) load_chunk(row0 + kF32Rows, 0);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        mbar_wait(&full[g % 3], (uint32_t)(g / 3) & 1u);      // this chunk of W has landed (requested one iteration ago)
        const uint32_t x0 = smem_u32(xh), w0 = smem_u32(Wring + (size_t)(g % 3) * wbytes);
        for (int mb = 0; mb < MB; ++mb) {
          // one M block: the large products (hi x hi) go to accumulator 0, the two small cross terms to accumulator 1, issued
          // alternately so that back-to-back MMAs never target the same TMEM columns
          const uint32_t d0 = taddr + 128u * mb, d1 = NACC == 2 ? d0 + 128u * MB : d0;
          const uint32_t wh = w0 + (uint32_t)(2 * mb) * kF32Plane, wl = wh + kF32Plane;
#pragma unroll
          for (int j = 0; j < kF32KC / 8; ++j) {
            const uint32_t ko = (uint32_t)j * 2u * kF32Lbo;
            const uint64_t dwh = tc_smem_desc(wh + ko, kF32Lbo, 128), dwl = tc_smem_desc(wl + ko, kF32Lbo, 128);
            const uint64_t dxh = tc_smem_desc(x0 + ko, kF32Lbo, 128), dxl = tc_smem_desc(x0 + kF32Plane + ko, kF32Lbo, 128);
            const bool first = c == 0 && j == 0;
            tc_mma_tf32(d1, dwl, dxh, idesc, !first);          // small terms first (NACC == 1: same accumulator)
            tc_mma_tf32(d0, dwh, dxh, idesc, NACC == 2 ? !first : true);
            tc_mma_tf32(d1, dwh, dxl, idesc, true);
          }
        }
        tc_commit(&mdone[b]);
        if (c + 1 == NC) tc_commit(&tdone);
      }
    }
    mbar_wait(&tdone, tile_par);
    tile_par ^= 1u;
    tc_fence_after();
    // ---- epilogue: lane = output column, registers = 32 rows; coalesced stores (and side reads) straight from registers
    for (int item = grp; item < MB * 4; item += 2) {
      const int mb = item >> 2, rc = item & 3;
      const int n = 128 * mb + 32 * q4 + lane;
      if (32 * rc >= rv) continue;                            // warp-uniform
      uint32_t acc[32];
      tc_ld32(taddr + ((uint32_t)(32 * q4) << 16) + (uint32_t)(128 * mb + 32 * rc), acc);
      tc_wait_ld();
      if (NACC == 2) {                                        // second accumulator (the small cross terms)
        uint32_t acc2[32];
        tc_ld32(taddr + ((uint32_t)(32 * q4) << 16) + (uint32_t)(128 * MB + 128 * mb + 32 * rc), acc2);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(acc2[j]));
      }
      if (n < p.N) {
        const long long r0 = row0 + 32 * rc;
        const int nr = min(32, rv - 32 * rc);
        if (p.epi == EPI_BIAS) {
          const float bn = p.bias ? p.bias[n] : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nr) {
              const float y = __uint_as_float(acc[j]) + bn;
              p.Y[(size_t)(r0 + j) * p.y_stride + n] = p.relu ? fmaxf(y, 0.f) : y;
            }
        } else if (p.epi == EPI_SSM) {
          const float dn = p.D[n];
          const float bs = p.bn_sub ? p.bn_sub[n] : 0.f, bm = p.bn_sub ? p.bn_mul[n] : 1.f, ba = p.bn_sub ? p.bn_add[n] : 0.f;
          // side reads in batches of 16 rows BEFORE the batch's stores: a load behind a store it may alias with is not
          // hoisted by the compiler, and one exposed L2 round trip per row made this epilogue latency bound
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float u[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = p.X2[(size_t)min(r0 + 16 * h + j, last_row) * p.N + n];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (16 * h + j < nr) {
                const float uu = p.bn_sub ? (u[j] - bs) * bm + ba : u[j];
                const float y = fmaf(p.conj_mul, __uint_as_float(acc[16 * h + j]), dn * uu);
                p.Y[(size_t)(r0 + 16 * h + j) * p.y_stride + n] = p.act == S5_ACT_RELU ? fmaxf(y, 0.f) : f32_gelu_tanh(y);
              }
          }
        } else {
          const float bn = p.bias[n];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float x1[16], sk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const size_t o = (size_t)min(r0 + 16 * h + j, last_row) * p.N + n;
              x1[j] = p.X1[o]; sk[j] = p.SKIP[o];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (16 * h + j < nr) {
                const float sg = f32_sigmoid_fast(__uint_as_float(acc[16 * h + j]) + bn);
                const float y = fmaf(x1[j], sg, sk[j]);
                p.Y[(size_t)(r0 + 16 * h + j) * p.y_stride + n] = p.relu ? fmaxf(y, 0.f) : y;
              }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                          // every accumulator has been read: the next tile may overwrite TMEM
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, tmem_cols);
}

// ------------------------------------------------------------------------------------------- warp-specialised form
// Same product, same operand images, same accumulators and epilogues as tc_f32_kernel - but no phase of a tile waits for
// another one.  One CTA of 15 warps per SM:
//   warp 8        TMA producer: the X tile arrives K chunk by K chunk as [128 rows][16 columns] boxes of a 2-D tensor map
//                 (cp.async.bulk.tensor; columns past K and rows past the end are zero-filled by the copy engine) in a ring
//                 of raw slots - HBM reads never stop while tiles are converted, multiplied or written
//   warp 9        W producer: all K chunks of W once when they fit beside the rings (resident), else chunk by chunk from L2
//                 into a ring of three slots (bulk copies)
//   warp 10       MMA issuer: per chunk 2 K steps x 3 products x M blocks; tcgen05.commit frees the X unit (and the W slot)
//   warps 11..14  converters (one warp per chunk): raw slot -> (BatchNorm) -> hi / lo planes of a ring of X units, operand layout
//   warps 0..7    epilogue (TMEM lane quadrant w % 4, row half w / 4) as before, in batches of 16 rows; with one M block the
//                 accumulators are double buffered (2 x 2 x 128 TMEM columns), so tile t is written while tile t + 1 is
//                 multiplied
// linked by mbarriers only.  Needs 16-byte aligned rows (K % 4 == 0, aligned base): the tensor map's row pitch.
constexpr int kFwEpi = 8, kFwConv = 4;
constexpr int kFwThreads = 32 * (kFwEpi + 3 + kFwConv);
constexpr uint32_t kFwRaw = kF32Rows * kF32KC * 4;            // one raw slot: [128][16] fp32, row-major
constexpr int kFwSupCols = kF32KC + 4;                        // SUPER: box width (16 columns + up to 3 of lead-in)
constexpr uint32_t kFwRawSup = kF32Rows * kFwSupCols * 4;     // SUPER: [128][20] fp32
constexpr uint32_t kFwUnit = 2 * kF32Plane;                   // hi / lo planes of one X chunk
// tc_f32g_kernel's x1 units: an epilogue warp writes 32 consecutive columns = two units at the same row; kFwUnit is a multiple of
// 128 bytes, which put lanes l and l + 16 on one bank (ncu: 42 % of the kernel's shared-memory wavefronts were conflicts) - 16 bytes
// of padding per unit move the second unit by four banks.  The region is rounded up to 128 bytes for what follows it.
constexpr uint32_t kFgX1Unit = kFwUnit + 16;
__host__ __device__ constexpr size_t fg_x1_bytes(int nc2) { return ((size_t)nc2 * kFgX1Unit + 127) & ~(size_t)127; }
constexpr int kFwMaxRing = 8;
constexpr int kFwSubRows = 32;                                // rows per box of the second tensor map (ragged last tile of a CTA)

__device__ __forceinline__ void tma2d_load(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

struct FwRing {                                               // slot / phase of a running counter over a ring of n slots
  int slot = 0; uint32_t par = 0;
  __device__ __forceinline__ void next(int n) { if (++slot == n) { slot = 0; par ^= 1u; } }
};

// BNF: BatchNorm on the input columns (EPI_BIAS) / on the side input u (EPI_SSM); GELU: activation of the SSM tail
// SUPER: rows whose pitch is not a multiple of 16 bytes (K = 257: the model input) cannot be the rows of a tensor map, but
// q = 4 / gcd(K, 4) consecutive rows can: the maps describe [rows / q] super-rows of q K floats, and a chunk of a tile arrives as
// q boxes, one per row i of the super-rows.  A box must START on a 16-byte boundary (measured: an unaligned inner coordinate
// raises "illegal instruction"), so the boxes are 20 floats wide and begin at the multiple of 4 below i K + 16 c; the converter
// skips the (i K) % 4 leading floats and clears the columns past K (they hold the next row).  Raw / operand row rr of a group
// of G rows (G = 128, or 32 in a ragged tile) holds tile row G (rr / G) + q (rr % (G / q)) + (rr % G) / (G / q); the epilogue
// undoes the permutation.  lq = log2(q); p.rows is a multiple of q.
template <int EPI, bool BNF, bool GELU, bool SUPER>
__global__ void __launch_bounds__(kFwThreads, 1) tc_f32w_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmx_sub,
                                                                const __grid_constant__ TcF32Params p, int resident, int NW, int NU, int NRAW, int lq) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // tensor-map copies want 128-byte aligned boxes
  const int MB = p.MB, NC = p.NC;
  const uint32_t wbytes = (uint32_t)MB * 2u * kF32Plane;
  unsigned char* Wbuf = smem;                                 // resident: NC chunks, else NW slots
  unsigned char* Ubuf = Wbuf + (size_t)(resident ? NC : NW) * wbytes;
  unsigned char* Rbuf = Ubuf + (size_t)NU * kFwUnit;
  constexpr uint32_t raw_bytes = SUPER ? kFwRawSup : kFwRaw, raw_pitch = SUPER ? kFwSupCols * 4 : kF32KC * 4;
  float4* bnp = reinterpret_cast<float4*>(Rbuf + (size_t)NRAW * raw_bytes);   // input BatchNorm: [sub | mul | add][NC * 16 columns]
  __shared__ __align__(8) uint64_t raw_full[kFwMaxRing], raw_free[kFwMaxRing], unit_full[kFwMaxRing], unit_free[kFwMaxRing];
  __shared__ __align__(8) uint64_t w_full[kFwMaxRing], w_free[kFwMaxRing], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NACC = MB <= 2 ? 2 : 1, NBUF = MB == 1 ? 2 : 1;   // as tc_f32_kernel: separate accumulator for the small cross terms
  const uint32_t acc_cols = (uint32_t)(MB * NACC) * 128u;     // TMEM columns of one accumulator buffer
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < kFwMaxRing; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], 1); mbar_init(&unit_full[i], 1); mbar_init(&unit_free[i], 1); }
    for (int i = 0; i < kFwMaxRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], kFwEpi); }
    fence_mbar_init();
  }
  if constexpr (EPI == EPI_BIAS && BNF) {                      // pad columns: identity (they meet zero weights)
    float* b = reinterpret_cast<float*>(bnp);
    const int Kp = NC * kF32KC;
    for (int k = tid; k < Kp; k += kFwThreads) {
      const bool in = k < p.K;
      b[k] = in ? p.in_sub[k] : 0.f; b[Kp + k] = in ? p.in_mul[k] : 1.f; b[2 * Kp + k] = in ? p.in_add[k] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_slot;
  const long long qmask = SUPER ? ~(long long)((1 << lq) - 1) : ~0LL;   // CTA boundaries on whole super-rows
  const long long row_begin = (p.rows * blockIdx.x / gridDim.x) & qmask;
  const long long row_end = blockIdx.x + 1 == gridDim.x ? p.rows : (p.rows * (blockIdx.x + 1) / gridDim.x) & qmask;
  const int ntiles = (int)((row_end - row_begin + kF32Rows - 1) / kF32Rows);

  if (warp == kFwEpi) {
    // ---------------------------------------------------------------- TMA producer (X)
    if (lane == 0) {
      FwRing r;
      long long g = 0;
      for (int t = 0; t < ntiles; ++t) {
        const long long row0 = row_begin + (long long)t * kF32Rows;
        const int rv = (int)min((long long)kF32Rows, row_end - row0);
        for (int c = 0; c < NC; ++c, ++g, r.next(NRAW)) {
          if (g >= NRAW) mbar_wait(&raw_free[r.slot], r.par ^ 1u);
          unsigned char* dst = Rbuf + (size_t)r.slot * raw_bytes;
          if constexpr (SUPER) {
            const int q = 1 << lq;
            if (rv == kF32Rows) {
              mbar_expect_tx(&raw_full[r.slot], kFwRawSup);
              for (int i = 0; i < q; ++i)
                tma2d_load(dst + (size_t)i * (kF32Rows >> lq) * raw_pitch, &tmx, (i * p.K + c * kF32KC) & ~3, (int)(row0 >> lq), &raw_full[r.slot]);
            } else {
              const int nb = (rv + kFwSubRows - 1) / kFwSubRows;
              mbar_expect_tx(&raw_full[r.slot], (uint32_t)nb * kFwSubRows * raw_pitch);
              for (int b = 0; b < nb; ++b)
                for (int i = 0; i < q; ++i)
                  tma2d_load(dst + (size_t)(b * kFwSubRows + i * (kFwSubRows >> lq)) * raw_pitch, &tmx_sub, (i * p.K + c * kF32KC) & ~3,
                             (int)((row0 + b * kFwSubRows) >> lq), &raw_full[r.slot]);
            }
          } else if (rv == kF32Rows) {
            mbar_expect_tx(&raw_full[r.slot], kFwRaw);
            tma2d_load(dst, &tmx, c * kF32KC, (int)row0, &raw_full[r.slot]);
          } else {                                            // ragged last tile: 32-row boxes, nothing beyond the CTA's rows but < 32 rows
            const int nb = (rv + kFwSubRows - 1) / kFwSubRows;
            mbar_expect_tx(&raw_full[r.slot], (uint32_t)nb * kFwSubRows * kF32KC * 4u);
            for (int i = 0; i < nb; ++i)
              tma2d_load(dst + (size_t)i * kFwSubRows * kF32KC * 4, &tmx_sub, c * kF32KC, (int)row0 + i * kFwSubRows, &raw_full[r.slot]);
          }
        }
      }
    }
  } else if (warp == kFwEpi + 1) {
    // ---------------------------------------------------------------- W producer
    if (lane == 0) {
      if (resident) {
        mbar_expect_tx(&w_full[0], (uint32_t)NC * wbytes);
        for (int c = 0; c < NC; ++c) bulk_g2s(Wbuf + (size_t)c * wbytes, p.Wg + (size_t)c * wbytes, wbytes, &w_full[0]);
      } else {
        FwRing w;
        long long g = 0;
        for (int t = 0; t < ntiles; ++t)
          for (int c = 0; c < NC; ++c, ++g, w.next(NW)) {
            if (g >= NW) mbar_wait(&w_free[w.slot], w.par ^ 1u);
            mbar_expect_tx(&w_full[w.slot], wbytes);
            bulk_g2s(Wbuf + (size_t)w.slot * wbytes, p.Wg + (size_t)c * wbytes, wbytes, &w_full[w.slot]);
          }
      }
    }
  } else if (warp == kFwEpi + 2) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      FwRing u, w;
      if (resident) mbar_wait(&w_full[0], 0);
      for (int t = 0; t < ntiles; ++t) {
        const int ab = NBUF == 2 ? (t & 1) : 0, use = NBUF == 2 ? (t >> 1) : t;
        const int rv = (int)min((long long)kF32Rows, row_end - (row_begin + (long long)t * kF32Rows));
        const uint32_t idesc = tc_idesc_tf32(kF32Rows, SUPER ? (rv + 31) & ~31 : (rv + 15) & ~15);   // SUPER: whole 32-row groups (permuted rows)
        if (use >= 1) mbar_wait(&acc_free[ab], (uint32_t)(use - 1) & 1u);
        for (int c = 0; c < NC; ++c, u.next(NU)) {
          mbar_wait(&unit_full[u.slot], u.par);
          if (!resident) mbar_wait(&w_full[w.slot], w.par);
          tc_fence_after();
          const uint32_t x0 = smem_u32(Ubuf + (size_t)u.slot * kFwUnit), w0 = smem_u32(Wbuf + (size_t)(resident ? c : w.slot) * wbytes);
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t d0 = taddr + (uint32_t)ab * acc_cols + 128u * mb, d1 = NACC == 2 ? d0 + 128u * MB : d0;
            const uint32_t wh = w0 + (uint32_t)(2 * mb) * kF32Plane, wl = wh + kF32Plane;
#pragma unroll
            for (int j = 0; j < kF32KC / 8; ++j) {
              const uint32_t ko = (uint32_t)j * 2u * kF32Lbo;
              const uint64_t dwh = tc_smem_desc(wh + ko, kF32Lbo, 128), dwl = tc_smem_desc(wl + ko, kF32Lbo, 128);
              const uint64_t dxh = tc_smem_desc(x0 + ko, kF32Lbo, 128), dxl = tc_smem_desc(x0 + kF32Plane + ko, kF32Lbo, 128);
              const bool first = c == 0 && j == 0;
              tc_mma_tf32(d1, dwl, dxh, idesc, !first);        // same order as tc_f32_kernel: identical results
              tc_mma_tf32(d0, dwh, dxh, idesc, NACC == 2 ? !first : true);
              tc_mma_tf32(d1, dwh, dxl, idesc, true);
            }
          }
          tc_commit(&unit_free[u.slot]);
          if (!resident) { tc_commit(&w_free[w.slot]); w.next(NW); }
        }
        tc_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= kFwEpi + 3) {
    // ---------------------------------------------------------------- converters: ONE warp per chunk, warp cw owns the chunks
    // g = cw (mod NU), i.e. X unit cw and the raw slots = cw (mod NU) (NRAW is a multiple of NU): every barrier is waited on by one
    // warp in phase order (a parity wait two phases ahead would pass), up to NU chunks are converted at the same time, and a lane
    // has 16 independent 16-byte items per chunk.  Item f = 16-byte group f & 3 of row f >> 2: raw offset 16 f (row-major slot).
    const int cw = warp - (kFwEpi + 3);
    if (cw < NU) {
      const int grp = lane & 3;
      const long long total = (long long)ntiles * NC;
      uint32_t rpar = 0, upar = 0;
      int slot = cw, c = cw % NC, t = cw / NC;
      unsigned char* const xu = Ubuf + (size_t)cw * kFwUnit + (uint32_t)grp * kF32Lbo + (uint32_t)(lane >> 2) * 16u;
      for (long long g = cw; g < total; g += NU) {
        float4 bsub, bmul, badd;
        if constexpr (EPI == EPI_BIAS && BNF) {
          const float4* b = bnp + (c * (kF32KC / 4) + grp);
          bsub = b[0]; bmul = b[NC * (kF32KC / 4)]; badd = b[2 * NC * (kF32KC / 4)];
        }
        mbar_wait(&raw_full[slot], rpar);
        if (g >= NU) mbar_wait(&unit_free[cw], upar ^ 1u);
        const unsigned char* raw = Rbuf + (size_t)slot * raw_bytes + (SUPER ? 0u : 16u * lane);
        // a CTA's last tile may be ragged: only the 32-row batches the MMA reads (it rounds the tile up to 16 rows)
        const int rvt = t + 1 < ntiles ? kF32Rows : (int)(row_end - (row_begin + (long long)t * kF32Rows));
        const int nbq = (rvt + 31) >> 5;
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
          if (bq >= nbq) break;
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if constexpr (SUPER) {                            // row class of raw row rr -> lead-in of its box
              const int rr = bq * 32 + i * 8 + (lane >> 2);
              const int cls = rvt == kF32Rows ? rr >> (7 - lq) : (rr & 31) >> (5 - lq);
              const float* src = reinterpret_cast<const float*>(raw + (uint32_t)rr * raw_pitch + 16u * grp) + ((cls * p.K) & 3);
              // (four 4-byte loads whose addresses are all congruent mod 4 words: a 4-way bank conflict each.  Starting every lane at
              // word `lane & 3` and rotating the registers back by selects removes it - 4.7 M -> 1.2 M conflicts - but the converters
              // are bound by instruction issue: 30.3 M -> 34.1 M warp instructions, 82.9 -> 89.4 us.  Measured, not kept.)
              v[i] = make_float4(src[0], src[1], src[2], src[3]);
            } else {
              v[i] = *reinterpret_cast<const float4*>(raw + (bq * 128 + i * 32) * 16);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if constexpr (EPI == EPI_BIAS && BNF) {
              v[i].x = (v[i].x - bsub.x) * bmul.x + badd.x; v[i].y = (v[i].y - bsub.y) * bmul.y + badd.y;
              v[i].z = (v[i].z - bsub.z) * bmul.z + badd.z; v[i].w = (v[i].w - bsub.w) * bmul.w + badd.w;
            }
            if constexpr (SUPER) {                            // columns past K hold the next row, not zeros
              const int k = c * kF32KC + 4 * grp;
              if (k + 3 >= p.K) {
                if (k >= p.K) v[i].x = 0.f;
                if (k + 1 >= p.K) v[i].y = 0.f;
                if (k + 2 >= p.K) v[i].z = 0.f;
                v[i].w = 0.f;
              }
            }
            float4 hi, lo;                                    // nearest tf32: |lo| <= 2^-12 |x|
            hi.x = __uint_as_float((__float_as_uint(v[i].x) + 0x1000u) & 0xFFFFE000u); lo.x = v[i].x - hi.x;
            hi.y = __uint_as_float((__float_as_uint(v[i].y) + 0x1000u) & 0xFFFFE000u); lo.y = v[i].y - hi.y;
            hi.z = __uint_as_float((__float_as_uint(v[i].z) + 0x1000u) & 0xFFFFE000u); lo.z = v[i].z - hi.z;
            hi.w = __uint_as_float((__float_as_uint(v[i].w) + 0x1000u) & 0xFFFFE000u); lo.w = v[i].w - hi.w;
            unsigned char* dst = xu + (bq * 32 + i * 8) * 16;  // row (bq 32 + i 8 + lane / 4)
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + kF32Plane) = lo;
          }
        }
        fence_proxy_async_smem();                             // the planes are read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(&raw_free[slot]); mbar_arrive(&unit_full[cw]); }
        upar ^= 1u;
        slot += NU;
        if (slot >= NRAW) { slot -= NRAW; rpar ^= 1u; }
        c += NU;
        while (c >= NC) { c -= NC; ++t; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: lane = output column, registers = 16 rows
    const int q4 = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
    for (int t = 0; t < ntiles; ++t) {
      const int ab = NBUF == 2 ? (t & 1) : 0, use = NBUF == 2 ? (t >> 1) : t;
      const long long row0 = row_begin + (long long)t * kF32Rows;
      const int rv = (int)min((long long)kF32Rows, row_end - row0);
      mbar_wait(&acc_full[ab], (uint32_t)use & 1u);
      tc_fence_after();
      for (int mb = 0; mb < MB; ++mb) {
        const int n = 128 * mb + 32 * q4 + lane;
        const bool ncol = n < p.N;
        const int nc = min(n, p.N - 1);
        const uint32_t d0 = taddr + (uint32_t)ab * acc_cols + 128u * mb + lane_base;
        float cb = 0.f, dn = 0.f, bs = 0.f, bm = 1.f, ba = 0.f;
        if constexpr (EPI == EPI_SSM) {
          dn = p.D[nc];
          if constexpr (BNF) { bs = p.bn_sub[nc]; bm = p.bn_mul[nc]; ba = p.bn_add[nc]; }
        } else {
          cb = p.bias ? p.bias[nc] : 0.f;
        }
        const bool relu = p.relu != 0;
        if constexpr (SUPER) {                                // permuted rows (see above); bias epilogue only
          const int lg = rv == kF32Rows ? 7 : 5, G = 1 << lg, ncov = (rv + 31) & ~31;
          float* const yb = p.Y + (size_t)row0 * p.y_stride + n;
          for (int c0 = half * 64; c0 < min(ncov, half * 64 + 64); c0 += 16) {
            uint32_t acc[16], acc2[16];
            tc_ld16(d0 + (uint32_t)c0, acc);
            if (NACC == 2) tc_ld16(d0 + 128u * MB + (uint32_t)c0, acc2);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int rr = c0 + j, w = rr & (G - 1);
              const int real = (rr & ~(G - 1)) + ((w & ((G >> lq) - 1)) << lq) + (w >> (lg - lq));
              const float a = NACC == 2 ? __uint_as_float(acc[j]) + __uint_as_float(acc2[j]) : __uint_as_float(acc[j]);
              const float v = a + cb;
              if (ncol && real < rv) yb[(size_t)real * p.y_stride] = relu ? fmaxf(v, 0.f) : v;
            }
          }
          continue;
        }
        for (int c0 = half * 64; c0 < min(rv, half * 64 + 64); c0 += 16) {
          const long long r0 = row0 + c0;
          const int nr = min(16, rv - c0);
          uint32_t acc[16], acc2[16];
          float s1[16], s2[16];
          tc_ld16(d0 + (uint32_t)c0, acc);
          if (NACC == 2) tc_ld16(d0 + 128u * MB + (uint32_t)c0, acc2);
          // side reads of the whole batch BEFORE its stores (a load behind a store it may alias with is not hoisted);
          // one 64-bit base per batch, 32-bit row offsets (rows past the batch repeat its last row)
          if constexpr (EPI == EPI_SSM) {
            const float* sp = p.X2 + (size_t)r0 * p.N + nc;
#pragma unroll
            for (int j = 0; j < 16; ++j) s1[j] = sp[min(j, nr - 1) * p.N];
          } else if constexpr (EPI == EPI_GATE) {
            const float* sp1 = p.X1 + (size_t)r0 * p.N + nc;
            const float* sp2 = p.SKIP + (size_t)r0 * p.N + nc;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int o = min(j, nr - 1) * p.N;
              s1[j] = sp1[o]; s2[j] = sp2[o];
            }
          }
          tc_wait_ld();
          float y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = NACC == 2 ? __uint_as_float(acc[j]) + __uint_as_float(acc2[j]) : __uint_as_float(acc[j]);
            if constexpr (EPI == EPI_BIAS) {
              const float v = a + cb;
              y[j] = relu ? fmaxf(v, 0.f) : v;
            } else if constexpr (EPI == EPI_SSM) {
              const float uu = BNF ? (s1[j] - bs) * bm + ba : s1[j];
              const float v = fmaf(p.conj_mul, a, dn * uu);
              y[j] = GELU ? f32_gelu_tanh(v) : fmaxf(v, 0.f);
            } else {
              const float v = fmaf(s1[j], f32_sigmoid_fast(a + cb), s2[j]);
              y[j] = relu ? fmaxf(v, 0.f) : v;
            }
          }
          if (ncol) {
            float* q = p.Y + (size_t)r0 * p.y_stride + n;
            if (nr == 16) {
#pragma unroll
              for (int j = 0; j < 16; ++j, q += p.y_stride) *q = y[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j, q += p.y_stride)
                if (j < nr) *q = y[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- C projection + SSM tail + out2 + gate in ONE launch
// The block tail of the float flavour (ssm.py:179-185, 618-622; layers.py:217-238) as two chained products per 128-row tile:
//   acc1 = C~ xs            (K1 = 2P chunks: the scan output through the tensor-map ring / converters of tc_f32w_kernel)
//   E1   : x1 = act(conj acc1 + D BN(x))  -> hi / lo operand planes of x1 in shared memory (never written to HBM)
//   acc2 = W_out2 x1        (operands: the x1 planes)
//   E2   : out = x1 sigmoid(acc2 + b) + x -> global (x1 = hi + lo exactly, read back from the planes)
// Against the two launches it replaces this saves writing x1, reading it twice and reading the block input twice: 153 instead of
// 337 MB per block at the bench shape.  Both weight sets are streamed from L2 through one slot ring (C~ chunks, then out2 chunks, per
// tile: the x1 planes take the room a resident W would need).  TMEM: acc1 = columns 0..255, acc2 = 256..511 (two accumulators each,
// as everywhere in this file); the issuer alternates MMA1(t), MMA2(t), MMA1(t + 1) ... so E2(t) runs under MMA1(t + 1).
// Same MMA order and epilogue arithmetic as the two launches: bit-identical (tested).  One M block: H <= 128.
struct TcF32GluParams {
  long long rows; int K1, H, NC1, NC2;
  const unsigned char* Wc; const unsigned char* Wo;
  const float* X; const float* bn_sub; const float* bn_mul; const float* bn_add; const float* D; float conj_mul;
  const float* o_bias; int relu;
  float* Y;
};

// 16 epilogue warps (TMEM lane quadrant x 32 rows) and the two converter warps the X-unit ring of this kernel can feed.  Measured at the
// bench shape (us per launch; the two launches it replaces: 108): 8 epilogue warps with 16-row batches 92; + all 64 block-input values of a lane
// requested before the wait for acc1 90; 16 warps with 8-row batches 87 (kept); 8 warps, both epilogues fed from 64 prefetched values,
// 16-row TMEM round trips 103.  The critical path of a tile is MMA1 -> E1 -> MMA2 -> E2 with E1 of the next tile behind E2 (shared x1
// planes): a cycle-counter trace gives 20 700 cycles per tile of which the issuer waits 7 900 for E1 and 2 300 for X units.
constexpr int kFgEpi = 16, kFgConv = 2;
constexpr int kFgThreads = 32 * (kFgEpi + 3 + kFgConv);

template <bool BNF, bool GELU>
__global__ void __launch_bounds__(kFgThreads, 1) tc_f32g_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmx_sub,
                                                                const __grid_constant__ TcF32GluParams p, int NW, int NU, int NRAW) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int NC1 = p.NC1, NC2 = p.NC2, NCT = NC1 + NC2;
  constexpr uint32_t wbytes = 2u * kF32Plane;                 // one chunk of either weight set (one M block, hi + lo)
  unsigned char* X1buf = smem;                                // NC2 units: hi / lo planes of x1, chunk = 16 columns
  unsigned char* Wbuf = X1buf + fg_x1_bytes(NC2);
  unsigned char* Ubuf = Wbuf + (size_t)NW * wbytes;
  unsigned char* Rbuf = Ubuf + (size_t)NU * kFwUnit;
  __shared__ __align__(8) uint64_t raw_full[kFwMaxRing], raw_free[kFwMaxRing], unit_full[kFwMaxRing], unit_free[kFwMaxRing];
  __shared__ __align__(8) uint64_t w_full[kFwMaxRing], w_free[kFwMaxRing], acc1_full, acc1_free, acc2_full, acc2_free, x1_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < kFwMaxRing; ++i) {
      mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], 1); mbar_init(&unit_full[i], 1); mbar_init(&unit_free[i], 1);
      mbar_init(&w_full[i], 1); mbar_init(&w_free[i], 1);
    }
    mbar_init(&acc1_full, 1); mbar_init(&acc1_free, kFgEpi); mbar_init(&acc2_full, 1); mbar_init(&acc2_free, kFgEpi); mbar_init(&x1_full, kFgEpi);
    fence_mbar_init();
  }
  for (size_t i = (size_t)tid * 16; i < fg_x1_bytes(NC2); i += kFgThreads * 16) *reinterpret_cast<int4*>(X1buf + i) = make_int4(0, 0, 0, 0);   // pad columns stay zero
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#ifdef S5_F32G_TRACE
  long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long tstart = clock64();
#define FG_T(i, stmt) do { const long long t_ = clock64(); stmt; tw[i] += clock64() - t_; } while (0)
#else
#define FG_T(i, stmt) stmt
#endif
  const uint32_t taddr = tmem_slot;
  const long long row_begin = p.rows * blockIdx.x / gridDim.x, row_end = p.rows * (blockIdx.x + 1) / gridDim.x;
  const int ntiles = (int)((row_end - row_begin + kF32Rows - 1) / kF32Rows);

  if (warp == kFgEpi) {
    // ---------------------------------------------------------------- TMA producer (xs)
    if (lane == 0) {
      FwRing r;
      long long g = 0;
      for (int t = 0; t < ntiles; ++t) {
        const long long row0 = row_begin + (long long)t * kF32Rows;
        const int rv = (int)min((long long)kF32Rows, row_end - row0);
        for (int c = 0; c < NC1; ++c, ++g, r.next(NRAW)) {
          if (g >= NRAW) mbar_wait(&raw_free[r.slot], r.par ^ 1u);
          unsigned char* dst = Rbuf + (size_t)r.slot * kFwRaw;
          if (rv == kF32Rows) {
            mbar_expect_tx(&raw_full[r.slot], kFwRaw);
            tma2d_load(dst, &tmx, c * kF32KC, (int)row0, &raw_full[r.slot]);
          } else {
            const int nb = (rv + kFwSubRows - 1) / kFwSubRows;
            mbar_expect_tx(&raw_full[r.slot], (uint32_t)nb * kFwSubRows * kF32KC * 4u);
            for (int i = 0; i < nb; ++i)
              tma2d_load(dst + (size_t)i * kFwSubRows * kF32KC * 4, &tmx_sub, c * kF32KC, (int)row0 + i * kFwSubRows, &raw_full[r.slot]);
          }
        }
      }
    }
  } else if (warp == kFgEpi + 1) {
    // ---------------------------------------------------------------- W producer: per tile the chunks of C~, then those of out2
    if (lane == 0) {
      FwRing w;
      long long g = 0;
      for (int t = 0; t < ntiles; ++t)
        for (int c = 0; c < NCT; ++c, ++g, w.next(NW)) {
          if (g >= NW) mbar_wait(&w_free[w.slot], w.par ^ 1u);
          mbar_expect_tx(&w_full[w.slot], wbytes);
          const unsigned char* src = c < NC1 ? p.Wc + (size_t)c * wbytes : p.Wo + (size_t)(c - NC1) * wbytes;
          bulk_g2s(Wbuf + (size_t)w.slot * wbytes, src, wbytes, &w_full[w.slot]);
        }
    }
  } else if (warp == kFgEpi + 2) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      FwRing u, w;
      auto chunk = [&](uint32_t x0, uint32_t w0, uint32_t d0, uint32_t idesc, bool first_chunk) {
        const uint32_t d1 = d0 + 128u, wh = w0, wl = w0 + kF32Plane;
#pragma unroll
        for (int j = 0; j < kF32KC / 8; ++j) {
          const uint32_t ko = (uint32_t)j * 2u * kF32Lbo;
          const uint64_t dwh = tc_smem_desc(wh + ko, kF32Lbo, 128), dwl = tc_smem_desc(wl + ko, kF32Lbo, 128);
          const uint64_t dxh = tc_smem_desc(x0 + ko, kF32Lbo, 128), dxl = tc_smem_desc(x0 + kF32Plane + ko, kF32Lbo, 128);
          const bool first = first_chunk && j == 0;
          tc_mma_tf32(d1, dwl, dxh, idesc, !first);            // same order as tc_f32_kernel: identical results
          tc_mma_tf32(d0, dwh, dxh, idesc, !first);
          tc_mma_tf32(d1, dwh, dxl, idesc, true);
        }
      };
      for (int t = 0; t < ntiles; ++t) {
        const int rv = (int)min((long long)kF32Rows, row_end - (row_begin + (long long)t * kF32Rows));
        const uint32_t idesc = tc_idesc_tf32(kF32Rows, (rv + 15) & ~15);
        if (t >= 1) FG_T(0, mbar_wait(&acc1_free, (uint32_t)(t - 1) & 1u));
        for (int c = 0; c < NC1; ++c, u.next(NU), w.next(NW)) {
          FG_T(1, mbar_wait(&unit_full[u.slot], u.par));
          FG_T(2, mbar_wait(&w_full[w.slot], w.par));
          tc_fence_after();
          chunk(smem_u32(Ubuf + (size_t)u.slot * kFwUnit), smem_u32(Wbuf + (size_t)w.slot * wbytes), taddr, idesc, c == 0);
          tc_commit(&unit_free[u.slot]);
          tc_commit(&w_free[w.slot]);
        }
        tc_commit(&acc1_full);
        FG_T(3, mbar_wait(&x1_full, (uint32_t)t & 1u));                // E1(t) has written the x1 planes (and E2(t - 1) is over: same warps)
        if (t >= 1) FG_T(4, mbar_wait(&acc2_free, (uint32_t)(t - 1) & 1u));
        for (int c = 0; c < NC2; ++c, w.next(NW)) {
          FG_T(5, mbar_wait(&w_full[w.slot], w.par));
          tc_fence_after();
          chunk(smem_u32(X1buf + (size_t)c * kFgX1Unit), smem_u32(Wbuf + (size_t)w.slot * wbytes), taddr + 256u, idesc, c == 0);
          tc_commit(&w_free[w.slot]);
        }
        tc_commit(&acc2_full);
      }
    }
  } else if (warp >= kFgEpi + 3) {
    // ---------------------------------------------------------------- converters (xs chunks), as in tc_f32w_kernel
    const int cw = warp - (kFgEpi + 3);
    if (cw < NU) {
      const int grp = lane & 3;
      const long long total = (long long)ntiles * NC1;
      uint32_t rpar = 0, upar = 0;
      int slot = cw, c = cw % NC1, t = cw / NC1;
      unsigned char* const xu = Ubuf + (size_t)cw * kFwUnit + (uint32_t)grp * kF32Lbo + (uint32_t)(lane >> 2) * 16u;
      for (long long g = cw; g < total; g += NU) {
        mbar_wait(&raw_full[slot], rpar);
        if (g >= NU) mbar_wait(&unit_free[cw], upar ^ 1u);
        const unsigned char* raw = Rbuf + (size_t)slot * kFwRaw + 16u * lane;
        const int rvt = t + 1 < ntiles ? kF32Rows : (int)(row_end - (row_begin + (long long)t * kF32Rows));
        const int nbq = (rvt + 31) >> 5;
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
          if (bq >= nbq) break;
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(raw + (bq * 128 + i * 32) * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 hi, lo;                                    // nearest tf32: |lo| <= 2^-12 |x|
            hi.x = __uint_as_float((__float_as_uint(v[i].x) + 0x1000u) & 0xFFFFE000u); lo.x = v[i].x - hi.x;
            hi.y = __uint_as_float((__float_as_uint(v[i].y) + 0x1000u) & 0xFFFFE000u); lo.y = v[i].y - hi.y;
            hi.z = __uint_as_float((__float_as_uint(v[i].z) + 0x1000u) & 0xFFFFE000u); lo.z = v[i].z - hi.z;
            hi.w = __uint_as_float((__float_as_uint(v[i].w) + 0x1000u) & 0xFFFFE000u); lo.w = v[i].w - hi.w;
            unsigned char* dst = xu + (bq * 32 + i * 8) * 16;
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + kF32Plane) = lo;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&raw_free[slot]); mbar_arrive(&unit_full[cw]); }
        upar ^= 1u;
        slot += NU;
        if (slot >= NRAW) { slot -= NRAW; rpar ^= 1u; }
        c += NU;
        while (c >= NC1) { c -= NC1; ++t; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogues: lane = column n of the block, registers = 16 rows
    const int q4 = warp & 3, part = warp >> 2;                // TMEM lane quadrant, rows [32 part, 32 part + 32) of the tile
    const uint32_t d0 = taddr + ((uint32_t)(32 * q4) << 16);
    const int n = 32 * q4 + lane, H = p.H;
    const bool ncol = n < H;
    const int nc = min(n, H - 1);
    const float dn = p.D[nc], ob = p.o_bias[nc];
    float bs = 0.f, bm = 1.f, ba = 0.f;
    if constexpr (BNF) { bs = p.bn_sub[nc]; bm = p.bn_mul[nc]; ba = p.bn_add[nc]; }
    const bool relu = p.relu != 0;
    // element (row, column n) of the x1 planes: chunk n / 16, 16-byte group (n / 4) % 4, word n % 4
    unsigned char* const x1p = X1buf + (size_t)(nc >> 4) * kFgX1Unit + (uint32_t)((nc >> 2) & 3) * kF32Lbo + 4u * (nc & 3);
    for (int t = 0; t < ntiles; ++t) {
      const long long row0 = row_begin + (long long)t * kF32Rows;
      const int rv = (int)min((long long)kF32Rows, row_end - row0);
      const int rbeg = 32 * part, rend = min(rv, rbeg + 32);
      // ---- E1
      FG_T(0, mbar_wait(&acc1_full, (uint32_t)t & 1u));
      tc_fence_after();
#ifdef S5_F32G_TRACE
      const long long te1 = clock64();
#endif
      for (int c0 = rbeg; c0 < rend; c0 += 8) {
        const int nr = min(8, rv - c0);
        uint32_t acc[8], acc2[8];
        float s1[8];
        tc_ld8(d0 + (uint32_t)c0, acc);
        tc_ld8(d0 + 128u + (uint32_t)c0, acc2);
        const float* sp = p.X + (size_t)(row0 + c0) * H + nc;
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = sp[min(j, nr - 1) * H];
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = __uint_as_float(acc[j]) + __uint_as_float(acc2[j]);
          const float uu = BNF ? (s1[j] - bs) * bm + ba : s1[j];
          const float v = fmaf(p.conj_mul, a, dn * uu);
          const float x1 = GELU ? f32_gelu_tanh(v) : fmaxf(v, 0.f);
          const float hi = __uint_as_float((__float_as_uint(x1) + 0x1000u) & 0xFFFFE000u);
          if (ncol) {
            *reinterpret_cast<float*>(x1p + (uint32_t)(c0 + j) * 16u) = hi;
            *reinterpret_cast<float*>(x1p + kF32Plane + (uint32_t)(c0 + j) * 16u) = x1 - hi;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();                               // the planes are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive(&acc1_free); mbar_arrive(&x1_full); }
#ifdef S5_F32G_TRACE
      tw[1] += clock64() - te1;
#endif
      // ---- E2
      FG_T(2, mbar_wait(&acc2_full, (uint32_t)t & 1u));
      tc_fence_after();
#ifdef S5_F32G_TRACE
      const long long te2 = clock64();
#endif
      for (int c0 = rbeg; c0 < rend; c0 += 8) {
        const int nr = min(8, rv - c0);
        uint32_t acc[8], acc2[8];
        float s1[8], sk[8];
        tc_ld8(d0 + 256u + (uint32_t)c0, acc);
        tc_ld8(d0 + 384u + (uint32_t)c0, acc2);
        const float* sp = p.X + (size_t)(row0 + c0) * H + nc;
#pragma unroll
        for (int j = 0; j < 8; ++j) sk[j] = sp[min(j, nr - 1) * H];
#pragma unroll
        for (int j = 0; j < 8; ++j)                             // x1 = hi + lo, exact
          s1[j] = *reinterpret_cast<const float*>(x1p + (uint32_t)(c0 + j) * 16u) + *reinterpret_cast<const float*>(x1p + kF32Plane + (uint32_t)(c0 + j) * 16u);
        tc_wait_ld();
        if (ncol) {
          float* q = p.Y + (size_t)(row0 + c0) * H + n;
#pragma unroll
          for (int j = 0; j < 8; ++j, q += H) {
            const float a = __uint_as_float(acc[j]) + __uint_as_float(acc2[j]);
            const float v = fmaf(s1[j], f32_sigmoid_fast(a + ob), sk[j]);
            if (j < nr) *q = relu ? fmaxf(v, 0.f) : v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_free);
#ifdef S5_F32G_TRACE
      tw[3] += clock64() - te2;
#endif
    }
  }
#ifdef S5_F32G_TRACE
  if (blockIdx.x == 3 && lane == 0 && (warp == 0 || warp == 7 || warp == kFgEpi + 2))
    printf("[f32g trace] warp %d: total %lld cyc, tiles %d | %lld %lld %lld %lld %lld %lld\n", warp, clock64() - tstart, ntiles, tw[0], tw[1], tw[2], tw[3], tw[4], tw[5]);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem_slot, 512);
}

// ------------------------------------------------------------------------------------------- host side
int tc_f32_build(const float* W /*[K][N] row-major*/, int K, int N, TcF32Dev* out) {
  out->ok = false;
  if (K < 8 || N < 8 || N > 384 || K > 4096) return S5_OK;
  const int MB = (N + 127) / 128, NC = (K + kF32KC - 1) / kF32KC;
  const size_t per_chunk = (size_t)MB * 2 * kF32Plane;
  std::vector<unsigned char> img(per_chunk * NC, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = W[(size_t)k * N + n];
      uint32_t bits;
      std::memcpy(&bits, &w, 4);
      bits = (bits + 0x1000u) & 0xFFFFE000u;                 // nearest tf32
      float hi;
      std::memcpy(&hi, &bits, 4);
      const float lo = w - hi;
      const int c = k / kF32KC, kk = k % kF32KC, mb = n / 128, r = n % 128;
      const size_t base = (size_t)c * per_chunk + (size_t)(2 * mb) * kF32Plane + (size_t)(kk >> 2) * kF32Lbo + (size_t)r * 16 + (size_t)(kk & 3) * 4;
      std::memcpy(&img[base], &hi, 4);
      std::memcpy(&img[base + kF32Plane], &lo, 4);
    }
  S5_CUDA(cudaMalloc(&out->Wg, img.size()));
  S5_CUDA(cudaMemcpy(out->Wg, img.data(), img.size(), cudaMemcpyHostToDevice));
  out->K = K; out->N = N; out->MB = MB; out->NC = NC;
  out->ok = true;
  return S5_OK;
}

void tc_f32_free(TcF32Dev* d) {
  cudaFree(d->Wg);
  *d = TcF32Dev();
}

// 2-D tensor map over X [rows][K] fp32 with boxes of [box_rows][16 columns]; out-of-range elements read as zero
static int make_tmap2d_f32(CUtensorMap* tm, const float* base, long long K, long long rows, int box_rows, int box_cols = kF32KC) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeFn>(f);
  }();
  if (!fn) return -1;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

// ring sizes of the warp-specialised kernel for one weight set; false when nothing fits
static bool tc_f32w_plan(const TcF32Dev& d, size_t raw_bytes, int* resident, int* NW, int* NU, int* NRAW, size_t* smem) {
  const size_t budget = 227 * 1024 - 1536 - 128 - (size_t)3 * d.NC * kF32KC * 4;   // static shared memory (barriers), alignment slack, BatchNorm columns
  const size_t wchunk = (size_t)d.MB * 2 * kF32Plane;
  // X units (= converter warps) and raw slots (a multiple of the units): the most units that leave 48 KB of raw slots in flight
  // (profiles/r02_pipe_rates.md: 64 KB in flight per SM saturate the memory system, 32 KB give 80 %), else the most raw slots
  auto rings = [&](size_t left, int min_raw) {
    int best = 0;
    for (int nu = std::min(4, kFwConv); nu >= 2; --nu) {
      if (left < (size_t)nu * kFwUnit) continue;
      int nraw = (int)std::min<size_t>(kFwMaxRing, (left - (size_t)nu * kFwUnit) / raw_bytes);
      nraw -= nraw % nu;
      if (nraw >= 6) { *NU = nu; *NRAW = nraw; return true; }
      if (nraw >= min_raw && nraw > best) { *NU = nu; *NRAW = nraw; best = nraw; }
    }
    return best > 0;
  };
  static const int force_stream = [] { const char* e = getenv("S5B200_F32_W_STREAM"); return e ? atoi(e) : 0; }();
  const size_t wres = (size_t)d.NC * wchunk;
  if (!force_stream && wres < budget && wres < (1u << 20) && rings(budget - wres, 4)) { *resident = 1; *NW = 0; }
  else {
    *resident = 0;
    bool ok = false;
    // a W chunk takes ~1.5 us to arrive from L2 and is multiplied in 0.2 us per M block: enough slots to cover that
    // (measured, K = 257 -> 96: three slots 86 us, six slots see profiles/r03_f32w.md), as long as 48 KB of raw slots remain
    const int nw_want = d.MB == 1 ? 6 : d.MB == 2 ? 4 : 3;
    for (int nw = nw_want; nw >= 2 && !ok; --nw)
      if ((size_t)nw * wchunk < budget && rings(budget - (size_t)nw * wchunk, nw > 3 ? 6 : nw == 3 ? 4 : 3)) { *NW = nw; ok = true; }
    if (!ok) return false;
  }
  *smem = (size_t)(*resident ? d.NC : *NW) * wchunk + (size_t)*NU * kFwUnit + (size_t)*NRAW * raw_bytes + 128 + (size_t)3 * d.NC * kF32KC * 4;
  return true;
}

int tc_f32_launch(const TcF32Dev& d, const TcF32Args& a, int kclass, cudaStream_t st) {
  if (!d.ok || !a.X || !a.Y || a.rows <= 0) S5_FAIL(S5_E_BADARG, "dense-tile f32 linear: not eligible");
  TcF32Params p{};
  p.X = a.X; p.rows = a.rows; p.K = d.K; p.x_stride = d.K;
  p.in_sub = a.in_sub; p.in_mul = a.in_mul; p.in_add = a.in_add;
  p.Wg = reinterpret_cast<const unsigned char*>(d.Wg); p.N = d.N; p.MB = d.MB; p.NC = d.NC;
  p.Y = a.Y; p.y_stride = d.N;
  p.epi = a.epi; p.bias = a.bias; p.relu = a.relu;
  p.X2 = a.X2; p.bn_sub = a.bn_sub; p.bn_mul = a.bn_mul; p.bn_add = a.bn_add; p.D = a.D; p.conj_mul = a.conj_mul; p.act = a.act;
  p.X1 = a.X1; p.SKIP = a.SKIP;
  if (a.epi == EPI_SSM && (!a.X2 || !a.D)) S5_FAIL(S5_E_BADARG, "dense-tile f32 linear: SSM epilogue needs x and D");
  if (a.epi == EPI_GATE && (!a.X1 || !a.SKIP || !a.bias)) S5_FAIL(S5_E_BADARG, "dense-tile f32 linear: gate epilogue needs x1, skip and bias");
  const bool vec4 = (d.K & 3) == 0 && aligned16(a.X);
  // warp-specialised kernel (S5B200_F32_V2=0: the first form): rows with a 16-byte pitch - or, for the plain bias epilogue,
  // super-rows of q = 4 / gcd(K, 4) rows (K = 257); at most 2^31 rows.  The < q rows left over go to the first form.
  const char* v2_env = getenv("S5B200_F32_V2");
  const bool v2_on = !v2_env || atoi(v2_env) != 0;
  const int lq = (d.K & 3) == 0 ? 0 : (d.K & 1) ? 2 : 1, q = 1 << lq;
  const bool super_ok = lq == 0 || (a.epi == EPI_BIAS && !a.in_sub);
  long long rows_old = a.rows, row_off = 0;                   // rows the first form processes, starting at row_off
  if (v2_on && aligned16(a.X) && super_ok && d.K >= kF32KC && a.rows < (1LL << 31) && a.rows >= q && (a.epi == EPI_BIAS || !a.in_sub)) {
    int resident = 0, NW = 0, NU = 0, NRAW = 0;
    size_t smem = 0;
    CUtensorMap tmx, tmx_sub;
    const long long rows_main = a.rows & ~(long long)(q - 1);
    const int bc = lq ? kFwSupCols : kF32KC;
    if (tc_f32w_plan(d, lq ? kFwRawSup : kFwRaw, &resident, &NW, &NU, &NRAW, &smem) &&
        make_tmap2d_f32(&tmx, a.X, (long long)q * d.K, rows_main >> lq, kF32Rows >> lq, bc) == 0 &&
        make_tmap2d_f32(&tmx_sub, a.X, (long long)q * d.K, rows_main >> lq, kFwSubRows >> lq, bc) == 0) {
      const bool gelu = a.act != S5_ACT_RELU, bn2 = a.bn_sub != nullptr;
      auto kernel = a.epi == EPI_BIAS  ? (lq ? tc_f32w_kernel<EPI_BIAS, false, false, true>
                                             : a.in_sub ? tc_f32w_kernel<EPI_BIAS, true, false, false> : tc_f32w_kernel<EPI_BIAS, false, false, false>)
                    : a.epi == EPI_SSM ? (gelu ? (bn2 ? tc_f32w_kernel<EPI_SSM, true, true, false> : tc_f32w_kernel<EPI_SSM, false, true, false>)
                                               : (bn2 ? tc_f32w_kernel<EPI_SSM, true, false, false> : tc_f32w_kernel<EPI_SSM, false, false, false>))
                                       : tc_f32w_kernel<EPI_GATE, false, false, false>;
      S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int grid = (int)std::max<long long>(1, std::min<long long>((rows_main + kF32Rows - 1) / kF32Rows, (long long)sm_count()));
      TcF32Params pm = p;
      pm.rows = rows_main;
      {
        ProfScope prof(kclass, st);
        kernel<<<grid, kFwThreads, smem, st>>>(tmx, tmx_sub, pm, resident, NW, NU, NRAW, lq);
        S5_LAUNCH_CHECK();
      }
      if (rows_main == a.rows) return S5_OK;
      row_off = rows_main; rows_old = a.rows - rows_main;
    }
  }
  if (row_off) {                                              // bias epilogue only: no side inputs to shift
    p.X += (size_t)row_off * d.K; p.Y += (size_t)row_off * d.N; p.rows = rows_old;
  }
  const size_t smem = 3 * ((size_t)d.MB * 2 * kF32Plane) + 2 * (2 * (size_t)kF32Plane);   // W ring + X double buffer
  auto kernel = (vec4 && !row_off) ? tc_f32_kernel<true> : tc_f32_kernel<false>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int per_sm = (d.MB <= 2 && 2 * (smem + 2048) <= 227 * 1024) ? 2 : 1;
  const int grid = (int)std::max<long long>(1, std::min<long long>((p.rows + 63) / 64, (long long)per_sm * sm_count()));
  ProfScope prof(kclass, st);
  kernel<<<grid, kF32Threads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// C projection + SSM tail + out2 + gate of one float block in one launch (tc_f32g_kernel); S5_E_UNSUPPORTED (no error text) when the
// shapes do not qualify - the caller then runs the two launches
int tc_f32_glu_launch(const TcF32Dev& fc, const TcF32Dev& fo, const TcF32GluArgs& a, int kclass, cudaStream_t st) {
  const char* env = getenv("S5B200_F32_GLU_FUSED");
  if (env && atoi(env) == 0) return S5_E_UNSUPPORTED;
  const int H = fc.N;
  if (!fc.ok || !fo.ok || fc.MB != 1 || fo.MB != 1 || fo.K != H || fo.N != H || (fc.K & 3) || fc.K < kF32KC || !aligned16(a.XS) ||
      a.rows <= 0 || a.rows >= (1LL << 31) || !a.X || !a.Y || !a.D || !a.o_bias)
    return S5_E_UNSUPPORTED;
  const size_t budget = 227 * 1024 - 1536 - 128;
  const size_t x1 = fg_x1_bytes(fo.NC);
  int NW = 3, NU = 0, NRAW = 0;
  bool ok = false;
  for (; NW >= 2 && !ok; --NW) {
    const size_t fixed = x1 + (size_t)NW * kFwUnit;
    for (int nu = kFgConv; nu >= 2 && !ok; --nu) {
      if (budget < fixed + (size_t)nu * kFwUnit) continue;
      int nraw = (int)std::min<size_t>(kFwMaxRing, (budget - fixed - (size_t)nu * kFwUnit) / kFwRaw);
      nraw -= nraw % nu;
      if (nraw >= 4) { NU = nu; NRAW = nraw; ok = true; }
    }
    if (ok) break;
  }
  if (!ok) return S5_E_UNSUPPORTED;
  CUtensorMap tmx, tmx_sub;
  if (make_tmap2d_f32(&tmx, a.XS, fc.K, a.rows, kF32Rows) != 0 || make_tmap2d_f32(&tmx_sub, a.XS, fc.K, a.rows, kFwSubRows) != 0) return S5_E_UNSUPPORTED;
  TcF32GluParams p{};
  p.rows = a.rows; p.K1 = fc.K; p.H = H; p.NC1 = fc.NC; p.NC2 = fo.NC;
  p.Wc = reinterpret_cast<const unsigned char*>(fc.Wg); p.Wo = reinterpret_cast<const unsigned char*>(fo.Wg);
  p.X = a.X; p.bn_sub = a.bn_sub; p.bn_mul = a.bn_mul; p.bn_add = a.bn_add; p.D = a.D; p.conj_mul = a.conj_mul;
  p.o_bias = a.o_bias; p.relu = a.relu; p.Y = a.Y;
  const bool gelu = a.act != S5_ACT_RELU, bn = a.bn_sub != nullptr;
  auto kernel = gelu ? (bn ? tc_f32g_kernel<true, true> : tc_f32g_kernel<false, true>) : (bn ? tc_f32g_kernel<true, false> : tc_f32g_kernel<false, false>);
  const size_t smem = x1 + (size_t)NW * kFwUnit + (size_t)NU * kFwUnit + (size_t)NRAW * kFwRaw + 128;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<long long>(1, std::min<long long>((a.rows + kF32Rows - 1) / kF32Rows, (long long)sm_count()));
  ProfScope prof(kclass, st);
  kernel<<<grid, kFgThreads, smem, st>>>(tmx, tmx_sub, p, NW, NU, NRAW);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

}  // namespace s5
