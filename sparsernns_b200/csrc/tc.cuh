// tc.cuh - tcgen05 (5th-gen tensor core) building blocks for the dense-tile projection path, sm_100a.
//
// Operands live in shared memory in the canonical K-major, no-swizzle ("interleaved") UMMA layout: 8-row x 16-byte
// core matrices, element (row r, column k) of a [R][K] f16 tile at byte offset
//     (k / 8) * (R * 16) + r * 16 + (k % 8) * 2
// i.e. stride-byte-offset (next 8-row group) = 128 B, leading-byte-offset (next 16-byte K chunk) = R * 16 B.
// 8-bit operands (kind::i8) use the same byte geometry with 16 elements per 16-byte chunk:
//     (k / 16) * LBO + r * 16 + (k % 16);  one MMA consumes K = 32 elements = two chunks, as for f16.
// Accumulators live in TMEM (lane = tile row, column = output column) and are read back with tcgen05.ld.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace s5 {
#ifdef __CUDACC__

__device__ __forceinline__ size_t tc_off(int rows, int r, int k) {  // byte offset inside a [rows][K] f16 operand tile
  return (size_t)(k >> 3) * ((size_t)rows * 16) + (size_t)r * 16 + (size_t)(k & 7) * 2;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// the same descriptor `units` x 16 bytes further: only the 14-bit start-address field changes, and it WRAPS (the field holds
// address bits 4..17; shared-window addresses of a large CTA do cross 2^18, so a plain 64-bit add would carry into the
// reserved bits - measured: wrong products in the second tile pipeline of the block kernel)
__device__ __forceinline__ uint64_t tc_desc_advance(uint64_t desc, uint32_t units) {
  return (desc & ~0x3FFFull) | ((desc + units) & 0x3FFFull);
}

// instruction descriptor for kind::f16: A = B = F16, D = F32, both K-major, M x N tile (cute::UMMA::InstrDescriptor)
__host__ __device__ __forceinline__ uint32_t tc_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A = B = BF16 (a_format, b_format = 1)
__host__ __device__ __forceinline__ uint32_t tc_idesc_bf16(int M, int N) {
  return tc_idesc_f16(M, N) | (1u << 7) | (1u << 10);
}

// kind::i8: A, B 8-bit integers (format 1 = signed, 0 = unsigned), D = S32 (c_format 2), both K-major.  MMA K = 32.
__host__ __device__ __forceinline__ uint32_t tc_idesc_i8(int M, int N, bool a_signed, bool b_signed) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tc_alloc(uint32_t* smem_slot, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive columns: thread i of the warp gets row (lane base + i), registers = columns
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------- operand tiles shared by the dense-tile kernels
constexpr int kTcRows = 128;       // rows of an activation tile (MMA M)

// four int16 codes (low halves of b0..b3) -> word of the four low bytes, word of the four high bytes
__device__ __forceinline__ void split_planes(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t& lo, uint32_t& hi) {
  const uint32_t t01 = __byte_perm(b0, b1, 0x5140), t23 = __byte_perm(b2, b3, 0x5140);   // {lo0, lo1, hi0, hi1}
  lo = __byte_perm(t01, t23, 0x5410);
  hi = __byte_perm(t01, t23, 0x7632);
}

// Operand tiles use a padded leading-byte-offset (K-chunk stride = rows*16 + 16 bytes): with it the staging stores
// (lanes along k) are bank-conflict free.
__host__ __device__ __forceinline__ uint32_t tc_lbo(int rows) { return (uint32_t)rows * 16u + 16u; }
__host__ __device__ __forceinline__ size_t tc_tile_bytes(int rows, int K) { return (size_t)(K >> 3) * tc_lbo(rows); }   // f16 / bf16
__host__ __device__ __forceinline__ size_t tc_tile_bytes8(int rows, int K) { return (size_t)(K >> 4) * tc_lbo(rows); }  // 8-bit

// copy a plain row-major [R][K] f16 matrix into the canonical operand layout
__device__ __forceinline__ void tc_load_weights(unsigned char* dst, const __half* __restrict__ W, int R, int K) {
  const int kchunks = K >> 3;
  const uint32_t lbo = tc_lbo(R);
  for (int i = threadIdx.x; i < R * kchunks; i += blockDim.x) {
    const int r = i / kchunks, kc = i % kchunks;
    *reinterpret_cast<int4*>(dst + (size_t)kc * lbo + (size_t)r * 16) = *reinterpret_cast<const int4*>(W + (size_t)r * K + kc * 8);
  }
}
// same for a row-major [R][K] int8 matrix (16 codes per 16-byte chunk)
__device__ __forceinline__ void tc_load_weights8(unsigned char* dst, const signed char* __restrict__ W, int R, int K) {
  const int kchunks = K >> 4;
  const uint32_t lbo = tc_lbo(R);
  for (int i = threadIdx.x; i < R * kchunks; i += blockDim.x) {
    const int r = i / kchunks, kc = i % kchunks;
    *reinterpret_cast<int4*>(dst + (size_t)kc * lbo + (size_t)r * 16) = *reinterpret_cast<const int4*>(W + (size_t)r * K + kc * 16);
  }
}

// issue the K/16 MMAs of one [128 x K] x [K x N] product (A and B in the padded canonical layout)
__device__ __forceinline__ void tc_issue(uint32_t d_tmem, const void* A, const void* B, int K, int N, bool accumulate_first,
                                         bool bf16 = false) {
  const uint32_t la = tc_lbo(kTcRows), lb = tc_lbo(N);
  const uint32_t idesc = bf16 ? tc_idesc_bf16(kTcRows, N) : tc_idesc_f16(kTcRows, N);
  // The descriptors of consecutive K steps differ only in the start-address field (low 14 bits, units of 16 bytes):
  // build them once and add a constant per step - the issuing thread is alone on the critical path of the tile.
  const uint64_t a_desc = tc_smem_desc(smem_u32(A), la, 128), b_desc = tc_smem_desc(smem_u32(B), lb, 128);
  const uint32_t sa = (2 * la) >> 4, sb = (2 * lb) >> 4;
  for (int j = 0; j < (K >> 4); ++j)
    tc_mma_f16(d_tmem, tc_desc_advance(a_desc, j * sa), tc_desc_advance(b_desc, j * sb), idesc, accumulate_first || j > 0);
}
// int8 form: K/32 MMAs of [128 x K] (8-bit, signed or unsigned) x [K x N] (int8; `brows` rows in the whole B tile)
__device__ __forceinline__ void tc_issue8(uint32_t d_tmem, const void* A, const void* B, int K, int N, int brows, bool a_signed) {
  const uint32_t la = tc_lbo(kTcRows), lb = tc_lbo(brows);
  const uint32_t idesc = tc_idesc_i8(kTcRows, N, a_signed, true);
  const uint64_t a_desc = tc_smem_desc(smem_u32(A), la, 128), b_desc = tc_smem_desc(smem_u32(B), lb, 128);
  const uint32_t sa = (2 * la) >> 4, sb = (2 * lb) >> 4;
  for (int j = 0; j < (K >> 5); ++j) tc_mma_i8(d_tmem, tc_desc_advance(a_desc, j * sa), tc_desc_advance(b_desc, j * sb), idesc, j > 0);
}

#endif  // __CUDACC__
}  // namespace s5
