// proj.cuh - row-tile projection machinery shared by the float, static-quant and fixed-point kernels.
//
// Y[rows][N] = f_out( f_in(X[rows][K]) @ W[K][N] ) with W resident in shared memory, either compressed
// (entries ordered by output column: (k, w) pairs, the "CSC by output" form of the pruned weights) or as a
// dense [N][K] tile.  A CTA stages TR rows of X transposed ([K][TR+1], conflict-free for both the
// lanes-along-k staging writes and the lanes-along-rows reads), each warp then walks whole output columns:
// the (k, w) stream is a warp-uniform broadcast read, every lane owns R = TR/32 rows and keeps R accumulators
// in registers.  Results go to a second transposed tile and leave through coalesced row-major stores.
#pragma once

#include <vector>

#include "common.cuh"
#include "tc_f32.cuh"

namespace s5 {

constexpr int kProjThreads = 256;
constexpr int kProjWarps = kProjThreads / 32;

struct SpEntry {
  int k;
  int wbits;  // float or int32 weight, bit pattern
};

// Immutable device copy of one weight matrix in both forms the kernels can consume.
struct SparseDev {
  int K = 0, N = 0, nnz = 0, dense = 0, dtype = S5_F32;
  int* colptr = nullptr;      // [N+1]           (compressed form)
  SpEntry* ent = nullptr;     // [nnz], (n, k) order
  int* wdense = nullptr;      // [N][K] bit patterns (dense form)
  std::vector<int> h_k, h_n;  // host copy of the index set for s5b200_sparse_export
  size_t smem_bytes() const { return dense ? (size_t)K * N * 4 : (size_t)nnz * 8 + (size_t)(N + 1) * 4; }
};

}  // namespace s5

// public opaque handle of s5b200_sparse_create
struct s5_sparse {
  s5::SparseDev d;
  s5::TcF32Dev tc;   // float weights: 3 x TF32 dense-tile form for throughput batches (tc_f32.cu)
};

namespace s5 {

int sparse_build(const void* W_host, int K, int N, int dtype, int dense_format, SparseDev* out);
void sparse_free(SparseDev* w);

// Weight view handed to kernels (pointers into global memory; copied to shared memory once per CTA).
struct WView {
  const int* colptr;
  const SpEntry* ent;
  const int* wdense;
  int K, N, nnz, dense;
};
inline WView wview(const SparseDev& s) { return WView{s.colptr, s.ent, s.wdense, s.K, s.N, s.nnz, s.dense}; }

#ifdef __CUDACC__

// Shared-memory image of a weight matrix.
struct WSmem {
  const int* colptr;
  const SpEntry* ent;
  const int* wdense;
};

__device__ __forceinline__ size_t wsmem_words(const WView& w) {
  return w.dense ? (size_t)w.K * w.N : (size_t)w.nnz * 2 + (size_t)(w.N + 1);
}

// Copies W into shared memory at `dst` (8-byte aligned) and returns the view.  All threads participate.
__device__ __forceinline__ WSmem wsmem_load(int* dst, const WView& w) {
  WSmem s{nullptr, nullptr, nullptr};
  if (w.dense) {
    for (int i = threadIdx.x; i < w.K * w.N; i += blockDim.x) dst[i] = w.wdense[i];
    s.wdense = dst;
  } else {
    int2* e = reinterpret_cast<int2*>(dst);
    const int2* src = reinterpret_cast<const int2*>(w.ent);
    for (int i = threadIdx.x; i < w.nnz; i += blockDim.x) e[i] = src[i];
    int* cp = dst + (size_t)w.nnz * 2;
    for (int i = threadIdx.x; i <= w.N; i += blockDim.x) cp[i] = w.colptr[i];
    s.ent = reinterpret_cast<const SpEntry*>(e);
    s.colptr = cp;
  }
  return s;
}

// acc += w * x; the weight arrives as a bit pattern.  The int form wraps modulo 2**32 (fxparray.py:662).
__device__ __forceinline__ void mac(float& acc, int wbits, float x) { acc = fmaf(__int_as_float(wbits), x, acc); }
__device__ __forceinline__ void mac(int& acc, int wbits, int x) { acc = (int)((unsigned)acc + (unsigned)wbits * (unsigned)x); }

// Zero-activation skipping on the CUDA-core path.  After a tile is staged, bit j of zmask[k] says whether input column k
// holds ANY non-zero among the 32 rows lane + 32 j of the tile (one ballot per 32-row group).  column_dot then skips the
// shared-memory load and the multiply-add of every (entry, row group) whose activations are all zero - a warp-uniform
// predicate, so it saves exactly the resource the compressed kernels are bound by (shared-memory reads).  ReLU-ified
// activations are sparse per element and whole groups vanish for rarely firing channels; per-ELEMENT skipping has no
// SIMT form with lanes along rows, and on the tensor-core path a zero costs nothing.
template <typename T, int R, int TR>
__device__ __forceinline__ void zero_group_mask(const T* __restrict__ xt, int K, uint32_t* zmask) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int k = warp; k < K; k += nw) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (__ballot_sync(0xffffffffu, xt[(size_t)k * (TR + 1) + lane + 32 * j] != T(0))) m |= 1u << j;
    if (lane == 0) zmask[k] = m;
  }
}

// One output column n for the R rows a lane owns.  T is float or int.  zmask: nullable, see zero_group_mask.
template <typename T, int R, int TR>
__device__ __forceinline__ void column_dot(const WSmem& w, int K, int n, const T* __restrict__ xt, int lane, T (&acc)[R],
                                           const uint32_t* __restrict__ zmask = nullptr) {
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = T(0);
  const T* xl = xt + lane;
  if (zmask) {
    if (w.wdense) {
      const int* wr = w.wdense + (size_t)n * K;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const uint32_t m = zmask[k];
        if (!m) continue;
        const int wb = wr[k];
        const T* xp = xl + (size_t)k * (TR + 1);
#pragma unroll
        for (int j = 0; j < R; ++j)
          if ((m >> j) & 1u) mac(acc[j], wb, xp[32 * j]);
      }
    } else {
      const int e0 = w.colptr[n], e1 = w.colptr[n + 1];
#pragma unroll 4
      for (int e = e0; e < e1; ++e) {
        const int2 en = *reinterpret_cast<const int2*>(w.ent + e);
        const uint32_t m = zmask[en.x];
        const T* xp = xl + (size_t)en.x * (TR + 1);
#pragma unroll
        for (int j = 0; j < R; ++j)
          if ((m >> j) & 1u) mac(acc[j], en.y, xp[32 * j]);
      }
    }
    return;
  }
  if (w.wdense) {
    const int* wr = w.wdense + (size_t)n * K;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const int wb = wr[k];
      const T* xp = xl + (size_t)k * (TR + 1);
#pragma unroll
      for (int j = 0; j < R; ++j) mac(acc[j], wb, xp[32 * j]);
    }
  } else {
    const int e0 = w.colptr[n], e1 = w.colptr[n + 1];
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
      const int2 en = *reinterpret_cast<const int2*>(w.ent + e);
      const T* xp = xl + (size_t)en.x * (TR + 1);
#pragma unroll
      for (int j = 0; j < R; ++j) mac(acc[j], en.y, xp[32 * j]);
    }
  }
}

// Stage TR rows of a row-major [rows][K] matrix into the transposed tile xt[K][TR+1]; rows past
// `rows_valid` are zero-filled.  f(value, k) is the per-element input transform.
// Every warp owns blocks of kStageRows consecutive rows and requests all of them before it converts and stores any:
// eight independent global loads in flight per lane instead of one DRAM round trip per row (the staging used to be
// the latency bottleneck of every CUDA-core kernel: time did not depend on how much arithmetic was skipped).
constexpr int kStageRows = 8;
template <typename Tin, typename T, int TR, typename F>
__device__ __forceinline__ void stage_rows(T* xt, const Tin* __restrict__ src, int rows_valid, int K, F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  for (int r0 = warp * kStageRows; r0 < TR; r0 += nw * kStageRows) {
    for (int k = lane; k < K; k += 32) {
      Tin v[kStageRows];
#pragma unroll
      for (int q = 0; q < kStageRows; ++q) v[q] = r0 + q < rows_valid ? src[(size_t)(r0 + q) * K + k] : Tin(0);
#pragma unroll
      for (int q = 0; q < kStageRows; ++q)
        if (r0 + q < TR) xt[(size_t)k * (TR + 1) + r0 + q] = r0 + q < rows_valid ? f(v[q], k) : T(0);
    }
  }
}

// Write the transposed tile yt[N][TR+1] to a row-major [rows][N] matrix; g(value, r, n) is the final
// per-element transform (e.g. residual add with a coalesced read of the skip tensor).  g is evaluated for a block of
// rows before any of them is stored, so the global reads inside g (skip connection) are not serialised behind the
// stores they cannot be proven independent of.
template <typename T, typename Tout, int TR, typename G>
__device__ __forceinline__ void unstage_rows(Tout* dst, const T* yt, int rows_valid, int N, G g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  for (int r0 = warp * kStageRows; r0 < rows_valid; r0 += nw * kStageRows) {
    for (int n = lane; n < N; n += 32) {
      Tout v[kStageRows];
#pragma unroll
      for (int q = 0; q < kStageRows; ++q)
        if (r0 + q < rows_valid) v[q] = g(yt[(size_t)n * (TR + 1) + r0 + q], r0 + q, n);
#pragma unroll
      for (int q = 0; q < kStageRows; ++q)
        if (r0 + q < rows_valid) dst[(size_t)(r0 + q) * N + n] = v[q];
    }
  }
}

// quantization.py:386-400 with zero_point 0: clip(round_half_even(x / s), qmin, qmax) * s
__device__ __forceinline__ float fake_quant(float x, float s, float qmin, float qmax) {
  float q = rintf(x / s);
  q = fminf(fmaxf(q, qmin), qmax);
  return q * s;
}

#endif  // __CUDACC__
}  // namespace s5
