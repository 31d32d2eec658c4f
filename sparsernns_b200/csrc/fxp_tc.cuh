// fxp_tc.cuh - host-side handles of the integer dense-tile (tcgen05 kind::i8) kernels in fxp_tc.cu.
//
// The fixed-point model multiplies 16-bit activation payloads with 8-bit weight payloads (w8a16, fxprun.py:302-308) and
// keeps the int32 sum, i.e. every projection is an exact integer GEMM: the activation is split into a signed high byte and
// an unsigned low byte, both planes go through tcgen05.mma.kind::i8, and 256 * acc_hi + acc_lo is the reference's int32
// `@` (fxparray.py:640-678) bit for bit (no plane sum can reach 2^31 for K <= 288).  What follows the product - shifts,
// saturation, bias, ReLU, sigmoid table, gate - stays integer arithmetic on the CUDA cores, as in fxp_model.cu.
#pragma once

#include "fxp_chain.cuh"

namespace s5 {

struct FxpTcLinearDev {
  int Kfull = 0, Nfull = 0, Kp[2] = {0, 0}, Np = 0, KV = 0;
  signed char* W8[2] = {nullptr, nullptr};   // int8 payloads [Np][Kp]: K in natural order / in the lane-interleaved order
  signed char* WN[2] = {nullptr, nullptr};   // payloads of the one output column beyond Np (decoder: 257 = 256 + 1)
  bool xn = false;
  bool ok = false;
};

struct FxpTcLinArgs {
  const void* X; int* Y; float* Yf; long long rows;
  int mode;                 // LIN_ENC / LIN_BPROJ / LIN_DEC (fxp_chain.cuh)
  int x_is_float, in_bits, in_exp;                       // ENC: fxp_from_fp FLOOR of a float input (fxprun.py:69-75)
  int rshift0, rshift1, out_bits;                        // fxp_matmul epilogue (even / odd output column), FLOOR
  const int* bias_aligned; int relu; float out_scale;    // FxpDense bias at out_exp; 2^-out_exp for Yf
  BnChain c; unsigned* fm; int* ex; int* status;         // BPROJ: BatchNorm chain of the block
  const int* next_mmean; int next_e_mean; unsigned* next_fm; int out_exp;   // ENC: maxima for block 0's first BatchNorm stage
  unsigned* next_cmax; unsigned* next_cmin;              // ENC: or per-column extremes of the output (fxp_bn_derive_kernel)
  unsigned short* U16;                                   // BPROJ (16-byte aligned rows): also store the 16-bit mixer input u [rows][K]
  int dec_inp_bits, dec_inp_exp, dec_w_exp, dec_out_exp; // DEC: dynamic input exponent against the static FxpDense config
};

// W: int32 payloads [K][N] (the layout sparse_build takes).  out->ok stays false when the shape or the payload range
// (int8) is not covered; the CUDA-core kernels of fxp_model.cu are used then.
int fxp_tc_linear_build(const int32_t* W, int K, int N, FxpTcLinearDev* out);
void fxp_tc_linear_free(FxpTcLinearDev* d);
int fxp_tc_linear_launch(const FxpTcLinearDev& d, const FxpTcLinArgs& a, int kclass, cudaStream_t st);

// ---- C projection + D feed-through + out2 + sigmoid + mult_gate (FxpSSM.forward tail, FxpSequenceLayer.forward)
struct FxpTcGluDev {
  int H = 0, K1 = 0;
  signed char* Wc = nullptr;   // [2H][K1] int8: rows 0..H-1 = C_re on the even k, rows H..2H-1 = C_im on the odd k
  signed char* Wo = nullptr;   // [H][128] int8 out2 kernel, K in the lane-interleaved order
  size_t smem = 0;
  bool ok = false;
};
struct FxpTcGluArgs {
  int* XS; const int* X; int* G; long long rows;
  int relu_in;                     // the scan left fxp_relu of the states to this kernel (bad tiles are written back ReLU-ed)
  const unsigned short* U16;       // the mixer input u as 16-bit payloads [rows][H], written by the B projection (else recomputed)
  BnChain c; unsigned* fm; int* ex; int* status;
  const int* D; int rs_cre, rs_cim, rs_D, y_bits, conj;
  int o_in_change, y_exp, o_inp_bits, o_inp_exp, o_rs, o_out_bits, o_out_exp; const int* o_bias_aligned;
  int sig_x_exp, sig_y_exp; int lut[8];
  int gl_bits, gl_exp, gr_bits, gr_exp, g_rs, g_bits, g_exp;
  int* bad_list; int* bad_count;   // 128-row tiles whose states do not fit 16 bits: left to the CUDA-core kernel
};
int fxp_tc_glu_build(const int32_t* C_re, const int32_t* C_im, const int32_t* out2_w, int H, int P, FxpTcGluDev* out);
void fxp_tc_glu_free(FxpTcGluDev* d);
int fxp_tc_glu_launch(const FxpTcGluDev& d, const FxpTcGluArgs& a, cudaStream_t st);

}  // namespace s5
