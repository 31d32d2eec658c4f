// tc_model.cuh - host-side handles of the dense-tile (tcgen05) kernels in tc_model.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "fxp_chain.cuh"

namespace s5 {

struct TcLinearDev {
  int Kfull = 0, Nfull = 0, Kp[2] = {0, 0}, Np = 0, KV = 0;
  signed char* W8[2] = {nullptr, nullptr};   // int8 codes [Np][Kp]: K in natural order / in the lane-interleaved order
  signed char* WN[2] = {nullptr, nullptr};   // codes of the extra output column (xn), same two orders
  bool xn = false;
  float* col_scale = nullptr;
  bool ok = false;
};

struct TcGluDev {
  int H = 0, K1 = 0; size_t smem = 0, smem_scan = 0, smem_bpj = 0;
  void* WB = nullptr; float* b_col_scale = nullptr; bool bpj_ok = false;   // fused B projection (int8 [2P][128], lane-interleaved K)
  void* Wc = nullptr; void* Wo = nullptr;   // bf16 [H][2P] and int8 [H][128] (lane-interleaved K) weight codes
  float xs_odd_mul = 1.f, c_scale = 1.f, o_col_scale = 1.f;
  bool ok = false;
  // wide form (H > 128 or 2P > 128, e.g. dim_scale 1.0): C in 64-column chunks of the operand layout, out2 in the wide K order
  bool wide_ok = false; size_t smem_wide = 0;
  // block kernel: the layer's shared-memory image in global memory, built by tc_glu_prepare_image at model-create time
  // (rebuilt by tc_glu_ensure_image only if a caller passes other scales / vectors than the ones prepared)
  mutable void* img = nullptr; mutable bool img_ok = false; mutable unsigned char img_key[160] = {};
};
struct TcGluArgs {
  const float* XS; const float* X; float* OUT; long long rows;
  const float* bn_sub; const float* bn_mul; const float* bn_add; const float* D; const float* o_bias;
  float s_ut, s_yt, s_o_act, s_o_out, s_gl, s_gr, qs_min, qs_max, qn_min, qn_max; int relu_out;
  // fused recurrence (optional): lam != nullptr means XS holds Bu [B][L][P] and the kernel runs the scan itself
  const float* lam = nullptr; int B = 0, L = 0, relu_scan = 0; void* scan_ws = nullptr; size_t scan_ws_bytes = 0;
  // fused B projection (optional, needs the fused recurrence): XS is then unused
  bool fuse_bproj = false; float s_bu_re = 0.f, s_bu_im = 0.f;
};

// returned by tc_glu_launch / tc_glu_launch_multi when the device refuses the cooperative launch of the fused
// recurrence (the error text is set); the caller then runs the unfused kernels
constexpr int kCoopRefused = -100;

bool is_pow2(float s);
float pow2_grid(const float* w, size_t n, size_t start, size_t stride);
int tc_glu_build(const float* C_tilde, const float* out2_kernel, int H, int P, float conj_mul, float s_o_act, const float* B_bar,
                 float s_ut, TcGluDev* out);
void tc_glu_free(TcGluDev* d);
int tc_glu_prepare_image(const TcGluDev& d, const TcGluArgs& a);
int tc_glu_launch(const TcGluDev& d, const TcGluArgs& a, cudaStream_t st);
bool tc_glu_scan_ok(const TcGluDev& d);
int tc_glu_launch_multi(const TcGluDev* const* devs, const TcGluArgs* args, int n, uint32_t* done, size_t done_bytes, cudaStream_t st);
size_t tc_glu_scan_workspace_bytes(int B, int L, int P);
int tc_enc_fx_launch(const signed char* W8, int Kfull, int Kp, int Np, int Nfull, int KV, const void* X, int x_is_float, int in_bits, int in_exp,
                     int* Y, long long rows, int rshift, int out_bits, const int* bias_aligned, int relu, unsigned* cmax, unsigned* cmin,
                     cudaStream_t st);
int tc_dec_fx_launch(const signed char* W8, const signed char* WN, int Kfull, int Kp, int Np, int Nfull, int KV, const int* X, int* Y, float* Yf,
                     long long rows, const int* ex_in, int inp_bits, int inp_exp, int w_exp, int out_exp, int out_bits, const int* bias_aligned,
                     float out_scale, int* status, cudaStream_t st);
int tc_linear_build(const float* W, int Kfull, int Nfull, float s_in, TcLinearDev* out);
void tc_linear_free(TcLinearDev* d);
int tc_linear_launch(const TcLinearDev& d, const float* X, float* Y, long long rows, const float* in_sub, const float* in_mul,
                     const float* in_add, float in_scale, float in_qmin, float in_qmax, const float* bias, float out_scale0,
                     float out_scale1, float out_qmin, float out_qmax, int relu, int kclass, cudaStream_t st);

}  // namespace s5
