// tc_f32.cuh - host-side handle of the 3 x TF32 dense-tile projection kernel (tc_f32.cu), float flavour.
#pragma once
#include "common.cuh"

namespace s5 {

struct TcF32Dev {
  void* Wg = nullptr;   // [K chunks][M blocks][hi, lo] planes of W^T in the tcgen05 operand layout
  int K = 0, N = 0, MB = 0, NC = 0;
  bool ok = false;
};

struct TcF32Args {
  const float* X = nullptr; float* Y = nullptr; long long rows = 0;
  const float* in_sub = nullptr; const float* in_mul = nullptr; const float* in_add = nullptr;   // BatchNorm on the input
  int epi = 0;                         // 0 bias (+ ReLU), 1 SSM tail, 2 gate (tc_f32.cu)
  const float* bias = nullptr; int relu = 0;
  const float* X2 = nullptr; const float* bn_sub = nullptr; const float* bn_mul = nullptr; const float* bn_add = nullptr;
  const float* D = nullptr; float conj_mul = 1.f; int act = 0;
  const float* X1 = nullptr; const float* SKIP = nullptr;
};

// block tail of the float flavour in one launch: xs [rows][2P] -> out [rows][H] (tc_f32g_kernel)
struct TcF32GluArgs {
  const float* XS = nullptr; const float* X = nullptr; float* Y = nullptr; long long rows = 0;
  const float* bn_sub = nullptr; const float* bn_mul = nullptr; const float* bn_add = nullptr;   // BatchNorm of the block input (u = BN(x))
  const float* D = nullptr; float conj_mul = 1.f; int act = 0;
  const float* o_bias = nullptr; int relu = 0;
};

int tc_f32_build(const float* W, int K, int N, TcF32Dev* out);
void tc_f32_free(TcF32Dev* d);
int tc_f32_launch(const TcF32Dev& d, const TcF32Args& a, int kclass, cudaStream_t st);
int tc_f32_glu_launch(const TcF32Dev& fc, const TcF32Dev& fo, const TcF32GluArgs& a, int kclass, cudaStream_t st);   // S5_E_UNSUPPORTED: run the two launches

}  // namespace s5
