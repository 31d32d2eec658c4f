// fxp_ops.cu - granular fixed-point ops mirroring sparseRNNs/fxparray.py one to one (element-wise, int32
// payloads on the device).  The fused model in fxp_model.cu is built from the same primitives
// (fxp_prims.cuh); these entry points exist so that every primitive can be checked against the oracle.
#include "fxp_prims.cuh"

namespace s5 {

static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
#define EW_LOOP(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void k_from_fp(const float* x, int* y, long long n, int bits, int exp, int mode) {
  const float sc = exp2f((float)exp);
  EW_LOOP(i, n) {
    const float v = x[i] * sc;  // fxparray.py:297
    int q;
    if (mode == S5_ROUND_HALF_UP) q = __float2int_rn(v);        // np.round: half to even
    else if (mode == S5_ROUND_CEIL) q = __float2int_ru(v);
    else q = __float2int_rd(v);
    y[i] = fx_sat(q, bits);
  }
}
__global__ void k_change_cfg(const int* x, int* y, long long n, int bits, int exp, int nbits, int nexp, int mode) {
  EW_LOOP(i, n) y[i] = fx_change_cfg(x[i], bits, exp, nbits, nexp, mode);
}
__global__ void k_add(const int* a, const int* b, int* y, long long n, long long nb, int a_bits, int a_exp, int b_bits,
                      int b_exp, int rbits, int rexp, int mode) {
  EW_LOOP(i, n) {
    const int av = fx_change_exp(a[i], a_bits, a_exp, rexp, mode);   // fxparray.py:449-455
    const int bv = fx_change_exp(b[i % nb], b_bits, b_exp, rexp, mode);
    y[i] = fx_sat(fx_addw(av, bv), rbits);
  }
}
__global__ void k_mul(const int* a, const int* b, int* y, long long n, long long nb, int rshift, int rbits, int mode) {
  EW_LOOP(i, n) y[i] = fx_mul_apply(a[i], b[i % nb], rshift, mode, rbits);
}
__global__ void k_relu(const int* x, int* y, long long n) { EW_LOOP(i, n) y[i] = x[i] > 0 ? x[i] : 0; }
__global__ void k_sigmoid(const int* x, int* y, long long n, int bits, int exp, int x_exp, int y_exp, int l0, int l1, int l2,
                          int l3, int l4, int l5, int l6, int l7) {
  const int lut[8] = {l0, l1, l2, l3, l4, l5, l6, l7};
  EW_LOOP(i, n) y[i] = fx_sigmoid(fx_change_exp(x[i], bits, exp, x_exp, S5_ROUND_FLOOR), x_exp, y_exp, lut);
}

// compute_best: [0] max|fa + fb| (add) or max|fa * fb| (mul), [1] max|fa|, [2] max|fb|
__global__ void k_cb_reduce(const int* a, const int* b, long long n, long long nb, int a_exp, int b_exp, int is_mul,
                            unsigned* slots) {
  __shared__ float red[32];
  float m0 = 0.f, m1 = 0.f, m2 = 0.f;
  EW_LOOP(i, n) {
    const float fa = fx_to_float(a[i], a_exp), fb = fx_to_float(b[i % nb], b_exp);
    m0 = fmaxf(m0, fabsf(is_mul ? fa * fb : fa + fb));
    m1 = fmaxf(m1, fabsf(fa));
    m2 = fmaxf(m2, fabsf(fb));
  }
  block_max_to(slots + 0, m0, red);
  block_max_to(slots + 1, m1, red);
  block_max_to(slots + 2, m2, red);
}
// plan[0] = result_exp, [1] = agg_bits, [2] = agg_exp, [3] = error flag
__global__ void k_cb_plan(const unsigned* slots, int* plan, int a_exp, int b_exp, int rbits, int is_mul) {
  if (threadIdx.x || blockIdx.x) return;
  if (is_mul) {
    const int e = fx_plan_mul(__uint_as_float(slots[0]), rbits);
    const int rs = a_exp + b_exp - e;
    plan[0] = e; plan[1] = 0; plan[2] = 0; plan[3] = (rs < 0 || rs > 31) ? 1 : 0;
  } else {
    const AddPlan p = fx_plan_add(__uint_as_float(slots[0]), __uint_as_float(slots[1]), __uint_as_float(slots[2]), a_exp, b_exp, rbits);
    plan[0] = p.res_exp; plan[1] = p.agg_bits; plan[2] = p.agg_exp;
    plan[3] = (p.agg_bits > 32 || p.agg_exp - a_exp > 31 || p.agg_exp - b_exp > 31 || p.res_exp - p.agg_exp > 31 ||
               p.agg_exp - p.res_exp > 31) ? 1 : 0;
  }
}
__global__ void k_cb_apply(const int* a, const int* b, int* y, long long n, long long nb, int a_bits, int a_exp, int b_bits,
                           int b_exp, int rbits, int is_mul, const int* plan) {
  if (plan[3]) return;
  if (is_mul) {
    const int rs = a_exp + b_exp - plan[0];
    EW_LOOP(i, n) y[i] = fx_mul_apply(a[i], b[i % nb], rs, S5_ROUND_FLOOR, rbits);
  } else {
    AddPlan p; p.res_exp = plan[0]; p.agg_bits = plan[1]; p.agg_exp = plan[2];
    EW_LOOP(i, n) y[i] = fx_add_cb_apply(a[i], a_bits, a_exp, b[i % nb], b_bits, b_exp, p, rbits);
  }
}

static int compute_best(const int* a, const int* b, int* y, long long n, long long nb, int a_bits, int a_exp, int b_bits,
                        int b_exp, int rbits, int is_mul, int* result_exp_host, cudaStream_t st) {
  if (!a || !b || !y || !result_exp_host) S5_FAIL(S5_E_BADARG, "fxp compute_best: null pointer");
  if (n <= 0 || nb <= 0 || n % nb) S5_FAIL(S5_E_BADSHAPE, "fxp compute_best: n=%lld nb=%lld", n, nb);
  unsigned* slots = nullptr;
  S5_CUDA(cudaMalloc(&slots, 8 * sizeof(unsigned)));
  int* plan = reinterpret_cast<int*>(slots + 4);
  int rc = S5_OK;
  int hplan[4] = {0, 0, 0, 0};
  do {
    if (cudaMemsetAsync(slots, 0, 8 * sizeof(unsigned), st) != cudaSuccess) { rc = S5_E_CUDA; break; }
    k_cb_reduce<<<ew_blocks(n), 256, 0, st>>>(a, b, n, nb, a_exp, b_exp, is_mul, slots);
    launch_counter()++;
    k_cb_plan<<<1, 32, 0, st>>>(slots, plan, a_exp, b_exp, rbits, is_mul);
    launch_counter()++;
    k_cb_apply<<<ew_blocks(n), 256, 0, st>>>(a, b, y, n, nb, a_bits, a_exp, b_bits, b_exp, rbits, is_mul, plan);
    launch_counter()++;
    if (cudaMemcpyAsync(hplan, plan, sizeof(hplan), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = S5_E_CUDA; break; }
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = S5_E_CUDA; break; }
  } while (0);
  cudaFree(slots);
  if (rc) S5_FAIL(rc, "fxp compute_best: %s", cudaGetErrorString(cudaGetLastError()));
  if (hplan[3]) S5_FAIL(S5_E_UNSUPPORTED, "fxp compute_best: invalid result_exp %d (negative shift; the reference raises ValueError)", hplan[0]);
  *result_exp_host = hplan[0];
  return S5_OK;
}

}  // namespace s5

using namespace s5;

extern "C" int s5b200_fxp_from_fp(const float* x, int32_t* y, long long n, int bits, int exp, int round_mode, void* stream) {
  if (!x || !y) S5_FAIL(S5_E_BADARG, "fxp_from_fp: null pointer");
  if (n <= 0 || bits < 1 || bits > 32 || exp < 0 || exp > 30) S5_FAIL(S5_E_BADARG, "fxp_from_fp: n=%lld bits=%d exp=%d", n, bits, exp);
  k_from_fp<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, bits, exp, round_mode);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
extern "C" int s5b200_fxp_change_cfg(const int32_t* x, int32_t* y, long long n, int bits, int exp, int new_bits, int new_exp,
                                     int round_mode, void* stream) {
  if (!x || !y) S5_FAIL(S5_E_BADARG, "fxp_change_cfg: null pointer");
  if (n <= 0 || bits < 1 || bits > 32 || new_bits < 1 || new_bits > 32 || abs(exp - new_exp) > 31 ||
      (round_mode == S5_ROUND_HALF_UP && exp - new_exp == 0 && false))
    S5_FAIL(S5_E_BADARG, "fxp_change_cfg: invalid configuration");
  k_change_cfg<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, bits, exp, new_bits, new_exp, round_mode);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
extern "C" int s5b200_fxp_add(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_bits, int a_exp,
                              int b_bits, int b_exp, int result_bits, int result_exp, int round_mode, void* stream) {
  if (!a || !b || !y) S5_FAIL(S5_E_BADARG, "fxp_add: null pointer");
  if (n <= 0 || nb <= 0 || n % nb || abs(a_exp - result_exp) > 31 || abs(b_exp - result_exp) > 31)
    S5_FAIL(S5_E_BADARG, "fxp_add: invalid sizes or exponents");
  k_add<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, y, n, nb, a_bits, a_exp, b_bits, b_exp, result_bits, result_exp, round_mode);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
extern "C" int s5b200_fxp_mul(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_exp, int b_exp,
                              int result_bits, int result_exp, int round_mode, void* stream) {
  if (!a || !b || !y) S5_FAIL(S5_E_BADARG, "fxp_mul: null pointer");
  const int rs = a_exp + b_exp - result_exp;
  if (rs < 0) S5_FAIL(S5_E_UNSUPPORTED, "fxp_mul: invalid result_exp: %d", result_exp);  // fxparray.py:619-621
  if (n <= 0 || nb <= 0 || n % nb || rs > 31 || (round_mode == S5_ROUND_HALF_UP && rs < 1)) S5_FAIL(S5_E_BADARG, "fxp_mul: invalid sizes or shift");
  k_mul<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, y, n, nb, rs, result_bits, round_mode);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
extern "C" int s5b200_fxp_add_compute_best(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_bits,
                                           int a_exp, int b_bits, int b_exp, int result_bits, int* result_exp_host, void* stream) {
  return compute_best(a, b, y, n, nb, a_bits, a_exp, b_bits, b_exp, result_bits, 0, result_exp_host, (cudaStream_t)stream);
}
extern "C" int s5b200_fxp_mul_compute_best(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_exp,
                                           int b_exp, int result_bits, int* result_exp_host, void* stream) {
  return compute_best(a, b, y, n, nb, 0, a_exp, 0, b_exp, result_bits, 1, result_exp_host, (cudaStream_t)stream);
}
extern "C" int s5b200_fxp_relu(const int32_t* x, int32_t* y, long long n, void* stream) {
  if (!x || !y || n <= 0) S5_FAIL(S5_E_BADARG, "fxp_relu: bad arguments");
  k_relu<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
extern "C" int s5b200_fxp_sigmoid(const int32_t* x, int32_t* y, long long n, int bits, int exp, int x_exp, int y_exp,
                                  const int32_t* lut8_host, void* stream) {
  if (!x || !y || !lut8_host || n <= 0) S5_FAIL(S5_E_BADARG, "fxp_sigmoid: bad arguments");
  if (x_exp < 0 || x_exp > 15 || y_exp < 1 || y_exp > 30 || abs(exp - x_exp) > 31) S5_FAIL(S5_E_BADARG, "fxp_sigmoid: exponents out of range");
  const int32_t* l = lut8_host;
  k_sigmoid<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, bits, exp, x_exp, y_exp, l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7]);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
