// scan.cuh - launch entry points of scan.cu shared with the model orchestration code.
#pragma once
#include "common.cuh"

namespace s5 {

struct ScanI32Params {
  const int* Bu_re; const int* Bu_im;  // planar [B][L][P]; interleaved (re, im) pairs in Bu_re when `interleaved`
  const int* A_re; const int* A_im;
  int* xs_re; int* xs_im;              // same convention
  int L, P; long long lanes;           // lanes = B * P
  int eA_re, eA_im, eBu_re, eBu_im, ex_re, ex_im, relu;
};

int launch_scan_i32(const ScanI32Params& q, bool interleaved, cudaStream_t st);
int launch_scan_c64(const float* Bu, const float* lam, float* xs, int B, int L, int P, int relu, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);

}  // namespace s5
