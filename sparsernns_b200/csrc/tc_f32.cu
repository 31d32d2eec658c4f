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
#include <algorithm>
#include <cmath>
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

__device__ __forceinline__ float f32_gelu_tanh(float x) {     // jax.nn.gelu(approximate=True), layers.py:116-124
  const float c = 0.7978845608028654f;
  return x * (0.5f * (1.0f + tanhf(c * (x + 0.044715f * (x * x * x)))));
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
                const float y = p.conj_mul * __uint_as_float(acc[16 * h + j]) + dn * uu;
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
                const float sg = 1.0f / (1.0f + expf(-(__uint_as_float(acc[16 * h + j]) + bn)));
                const float y = x1[j] * sg + sk[j];
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
  const size_t smem = 3 * ((size_t)d.MB * 2 * kF32Plane) + 2 * (2 * (size_t)kF32Plane);   // W ring + X double buffer
  const bool vec4 = (d.K & 3) == 0 && aligned16(a.X);
  auto kernel = vec4 ? tc_f32_kernel<true> : tc_f32_kernel<false>;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int per_sm = (d.MB <= 2 && 2 * (smem + 2048) <= 227 * 1024) ? 2 : 1;
  const int grid = (int)std::max<long long>(1, std::min<long long>((a.rows + 63) / 64, (long long)per_sm * sm_count()));
  ProfScope prof(kclass, st);
  kernel<<<grid, kF32Threads, smem, st>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

}  // namespace s5
