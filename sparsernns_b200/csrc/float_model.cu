// float_model.cu - float32 and static fake-quant (w8a16) S5 / NDNS forward.
//
// Reference path (paths relative to the reference repo):
//   QRegressionModel.__call__        sparseRNNs/model/seq_model.py:439-456  (encoder Dense -> blocks -> decoder Dense)
//   QSequenceLayer.__call__          sparseRNNs/model/layers.py:180-241     (BN -> SSM -> act -> GLU half1 -> +skip -> ReLU)
//   qS5SSM.__call__ / apply_ssm      sparseRNNs/model/ssm.py:84-185, 194-343, 525-622
//   QuantizedDense / FakeQuant       sparseRNNs/utils/quantization.py:386-448, 684-816
//
// Kernels per forward: encoder linear, then per block {BN+B-projection linear, scan_c64, C-projection+GLU},
// decoder linear.  Bu/xs are the only SSM intermediates that touch HBM.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "proj.cuh"
#include "scan.cuh"
#include "tc_model.cuh"
#include "calib.cuh"

namespace s5 {

// ------------------------------------------------------------------------------------------- sparse
int sparse_build(const void* W_host, int K, int N, int dtype, int dense_format, SparseDev* out) {
  if (!W_host || !out) S5_FAIL(S5_E_BADARG, "sparse_create: null pointer");
  if (K <= 0 || N <= 0 || K > 4096 || N > 4096) S5_FAIL(S5_E_BADSHAPE, "sparse_create: K=%d N=%d outside [1,4096]", K, N);
  if (dtype != S5_F32 && dtype != S5_I32) S5_FAIL(S5_E_BADARG, "sparse_create: dtype %d", dtype);
  const int* wb = reinterpret_cast<const int*>(W_host);
  const float* wf = reinterpret_cast<const float*>(W_host);
  out->K = K; out->N = N; out->dense = dense_format ? 1 : 0; out->dtype = dtype;
  std::vector<int> colptr(N + 1, 0);
  std::vector<SpEntry> ent;
  out->h_k.clear(); out->h_n.clear();
  for (int n = 0; n < N; ++n) {
    for (int k = 0; k < K; ++k) {
      const size_t i = (size_t)k * N + n;
      const bool nz = dtype == S5_F32 ? (wf[i] != 0.0f) : (wb[i] != 0);
      if (nz) {
        ent.push_back(SpEntry{k, wb[i]});
        out->h_k.push_back(k);
        out->h_n.push_back(n);
      }
    }
    colptr[n + 1] = (int)ent.size();
  }
  out->nnz = (int)ent.size();
  S5_CUDA(cudaMalloc(&out->colptr, (N + 1) * sizeof(int)));
  S5_CUDA(cudaMemcpy(out->colptr, colptr.data(), (N + 1) * sizeof(int), cudaMemcpyHostToDevice));
  S5_CUDA(cudaMalloc(&out->ent, std::max<size_t>(1, ent.size()) * sizeof(SpEntry)));
  if (!ent.empty()) S5_CUDA(cudaMemcpy(out->ent, ent.data(), ent.size() * sizeof(SpEntry), cudaMemcpyHostToDevice));
  if (out->dense) {
    std::vector<int> wd((size_t)K * N);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) wd[(size_t)n * K + k] = wb[(size_t)k * N + n];
    S5_CUDA(cudaMalloc(&out->wdense, wd.size() * sizeof(int)));
    S5_CUDA(cudaMemcpy(out->wdense, wd.data(), wd.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  return S5_OK;
}

void sparse_free(SparseDev* w) {
  if (!w) return;
  cudaFree(w->colptr); cudaFree(w->ent); cudaFree(w->wdense);
  w->colptr = nullptr; w->ent = nullptr; w->wdense = nullptr;
}

static inline size_t wsmem_words_host(const WView& s) {
  size_t w = s.dense ? (size_t)s.K * s.N : (size_t)s.nnz * 2 + (size_t)(s.N + 1);
  return (w + 1) & ~(size_t)1;
}

constexpr size_t kSmemMax = 227 * 1024;

// ------------------------------------------------------------------------------------------- linear
struct LinearParams {
  const float* X; float* Y; long long rows; WView w;
  const float* in_sub; const float* in_mul; const float* in_add;  // u = (x - sub) * mul + add when in_sub != null
  int in_quant; float in_scale, in_qmin, in_qmax;
  const float* bias;
  int out_quant; float out_scale0, out_scale1, out_qmin, out_qmax;  // scale0 for even n, scale1 for odd n
  int relu;
  int skip_zero;   // zero-activation skipping per (input column, 32-row group), proj.cuh:zero_group_mask
};

template <int TR>
__global__ void __launch_bounds__(kProjThreads) linear_f32_kernel(LinearParams p) {
  constexpr int R = TR / 32;
  extern __shared__ __align__(16) int smem_i[];
  const int K = p.w.K, N = p.w.N;
  size_t ww = wsmem_words(p.w);
  ww = (ww + 1) & ~(size_t)1;
  const WSmem w = wsmem_load(smem_i, p.w);
  float* xt = reinterpret_cast<float*>(smem_i + ww);
  float* yt = xt + (size_t)K * (TR + 1);
  uint32_t* zmask = p.skip_zero ? reinterpret_cast<uint32_t*>(yt + (size_t)N * (TR + 1)) : nullptr;   // [K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (p.rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int rv = (int)min((long long)TR, p.rows - row0);
    stage_rows<float, float, TR>(xt, p.X + row0 * K, rv, K, [&](float v, int k) {
      if (p.in_sub) v = (v - p.in_sub[k]) * p.in_mul[k] + p.in_add[k];
      if (p.in_quant) v = fake_quant(v, p.in_scale, p.in_qmin, p.in_qmax);
      return v;
    });
    __syncthreads();
    if (zmask) {
      zero_group_mask<float, R, TR>(xt, K, zmask);
      __syncthreads();
    }
    for (int n = warp; n < N; n += kProjWarps) {
      float acc[R];
      column_dot<float, R, TR>(w, K, n, xt, lane, acc, zmask);
      const float bn = p.bias ? p.bias[n] : 0.f;
      const float sc = (n & 1) ? p.out_scale1 : p.out_scale0;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float v = acc[j] + bn;
        if (p.out_quant) v = fake_quant(v, sc, p.out_qmin, p.out_qmax);
        if (p.relu) v = fmaxf(v, 0.f);
        yt[(size_t)n * (TR + 1) + lane + 32 * j] = v;
      }
    }
    __syncthreads();
    unstage_rows<float, float, TR>(p.Y + row0 * N, yt, rv, N, [](float v, int, int) { return v; });
    // the next iteration's staging only writes xt; its __syncthreads orders yt reuse
  }
}

static int pick_tr(size_t fixed_bytes, size_t per_row_bytes) {
  // prefer two resident CTAs per SM, then anything that fits
  const int cands[3] = {128, 64, 32};
  for (int tr : cands)
    if (fixed_bytes + per_row_bytes * (tr + 1) <= kSmemMax / 2 - 1024) return tr;
  for (int tr : cands)
    if (fixed_bytes + per_row_bytes * (tr + 1) <= kSmemMax) return tr;
  return 0;
}

template <typename KernelT>
static int persistent_grid(KernelT kernel, size_t smem, long long ntiles, int* grid_out) {
  int per_sm = 0;
  S5_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  S5_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kProjThreads, smem));
  if (per_sm < 1) S5_FAIL(S5_E_BADSHAPE, "kernel does not fit on an SM with %zu bytes of shared memory", smem);
  const long long g = std::min<long long>(ntiles, (long long)per_sm * sm_count());
  *grid_out = (int)std::max<long long>(1, g);
  return S5_OK;
}

static int launch_linear(LinearParams p, cudaStream_t st, int kclass = S5_K_OTHER) {
  if (p.rows <= 0) S5_FAIL(S5_E_BADSHAPE, "linear: rows=%lld must be positive", p.rows);
  const size_t wbytes = ((wsmem_words_host(p.w)) * 4);
  const size_t zbytes = p.skip_zero ? (size_t)p.w.K * 4 : 0;
  const int tr = pick_tr(wbytes + zbytes, (size_t)(p.w.K + p.w.N) * 4);
  if (!tr) S5_FAIL(S5_E_BADSHAPE, "linear: K=%d N=%d weights (%zu B) do not fit in shared memory", p.w.K, p.w.N, wbytes);
  const size_t smem = wbytes + (size_t)(p.w.K + p.w.N) * (tr + 1) * 4 + zbytes;
  const long long ntiles = (p.rows + tr - 1) / tr;
  int grid = 1, rc;
  ProfScope prof(kclass, st);
  switch (tr) {
    case 128:
      if ((rc = persistent_grid(linear_f32_kernel<128>, smem, ntiles, &grid))) return rc;
      linear_f32_kernel<128><<<grid, kProjThreads, smem, st>>>(p);
      break;
    case 64:
      if ((rc = persistent_grid(linear_f32_kernel<64>, smem, ntiles, &grid))) return rc;
      linear_f32_kernel<64><<<grid, kProjThreads, smem, st>>>(p);
      break;
    default:
      if ((rc = persistent_grid(linear_f32_kernel<32>, smem, ntiles, &grid))) return rc;
      linear_f32_kernel<32><<<grid, kProjThreads, smem, st>>>(p);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// ------------------------------------------------------------------------------------------- C-projection + GLU
struct GluParams {
  const float* XS;   // [rows][2P]  scan output (ReLU already applied when relufication)
  const float* X;    // [rows][H]   block input (skip connection and BatchNorm source); u itself when !bn
  float* OUT;        // [rows][H]   block output, or y of the SSM when ssm_only
  long long rows;
  WView wc, wo;      // C: K = 2P, N = H ; out2: K = H, N = H
  const float* bn_sub; const float* bn_mul; const float* bn_add;  // null: X already holds u
  const float* D; const float* out2_bias;
  float conj_mul;
  int static_q; float s_ut, s_yt, s_o_act, s_o_out, s_gl, s_gr;
  float qs_min, qs_max, qn_min, qn_max;  // SSM-activation / non-SSM-activation integer ranges
  int act, relu_out, ssm_only;
  int skip_zero;   // zero-activation skipping for the C projection (ReLU-ified states) and out2 (x1 >= 0)
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // jax.nn.gelu(approximate=True)
  const float c = 0.7978845608028654f;
  return x * (0.5f * (1.0f + tanhf(c * (x + 0.044715f * (x * x * x)))));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int TR>
__global__ void __launch_bounds__(kProjThreads) glu_f32_kernel(GluParams p) {
  constexpr int R = TR / 32;
  constexpr int LD = TR + 1;
  extern __shared__ __align__(16) int smem_i[];
  const int K1 = p.wc.K, H = p.wc.N;
  size_t w1 = (wsmem_words(p.wc) + 1) & ~(size_t)1;
  size_t w2 = p.ssm_only ? 0 : ((wsmem_words(p.wo) + 1) & ~(size_t)1);
  const WSmem wc = wsmem_load(smem_i, p.wc);
  WSmem wo{nullptr, nullptr, nullptr};
  if (!p.ssm_only) wo = wsmem_load(smem_i + w1, p.wo);
  const int rowsA = K1 > H ? K1 : H;
  float* bufA = reinterpret_cast<float*>(smem_i + w1 + w2);  // xs tile, later g
  float* bufB = bufA + (size_t)rowsA * LD;                   // x tile (u source), later FQ(x1) for out2
  float* bufC = bufB + (size_t)H * LD;                       // x1
  uint32_t* zmask = p.skip_zero ? reinterpret_cast<uint32_t*>(bufC + (size_t)H * LD) : nullptr;   // [max(K1, H)]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (p.rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int rv = (int)min((long long)TR, p.rows - row0);
    __syncthreads();  // previous tile's write-out reads bufA
    stage_rows<float, float, TR>(bufA, p.XS + row0 * K1, rv, K1, [](float v, int) { return v; });
    stage_rows<float, float, TR>(bufB, p.X + row0 * H, rv, H, [&](float v, int k) {
      if (p.bn_sub) v = (v - p.bn_sub[k]) * p.bn_mul[k] + p.bn_add[k];
      if (p.static_q) v = fake_quant(v, p.s_ut, p.qs_min, p.qs_max);
      return v;  // u (u_q for the static flavour), ssm.py:230
    });
    __syncthreads();
    if (zmask) {
      zero_group_mask<float, R, TR>(bufA, K1, zmask);
      __syncthreads();
    }
    // y = conj_mul * Re(C x) + D * u  [-> FQ]  -> x1 = act(y)
    for (int n = warp; n < H; n += kProjWarps) {
      float acc[R];
      column_dot<float, R, TR>(wc, K1, n, bufA, lane, acc, zmask);
      const float dn = p.D[n];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const size_t o = (size_t)n * LD + lane + 32 * j;
        float y = p.conj_mul * acc[j] + dn * bufB[o];
        if (p.static_q) y = fake_quant(y, p.s_yt, p.qs_min, p.qs_max);
        if (p.ssm_only) {
          bufC[o] = y;
        } else {
          const float x1 = p.act == S5_ACT_RELU ? fmaxf(y, 0.f) : gelu_tanh_f(y);
          bufC[o] = x1;
          bufB[o] = p.static_q ? fake_quant(x1, p.s_o_act, p.qn_min, p.qn_max) : x1;  // QuantizedDense input
        }
      }
    }
    __syncthreads();
    if (p.ssm_only) {
      unstage_rows<float, float, TR>(p.OUT + row0 * H, bufC, rv, H, [](float v, int, int) { return v; });
      continue;
    }
    if (zmask) {
      zero_group_mask<float, R, TR>(bufB, H, zmask);
      __syncthreads();
    }
    // o = out2(x1) ; g = FQ(x1) * FQ(sigmoid(o))
    for (int n = warp; n < H; n += kProjWarps) {
      float acc[R];
      column_dot<float, R, TR>(wo, H, n, bufB, lane, acc, zmask);
      const float bn = p.out2_bias[n];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const size_t o = (size_t)n * LD + lane + 32 * j;
        float ov = acc[j] + bn;
        if (p.static_q) ov = fake_quant(ov, p.s_o_out, p.qn_min, p.qn_max);
        float sg = sigmoid_f(ov);
        float l = bufC[o];
        if (p.static_q) {
          sg = fake_quant(sg, p.s_gr, p.qn_min, p.qn_max);
          l = fake_quant(l, p.s_gl, p.qn_min, p.qn_max);
        }
        bufA[o] = l * sg;
      }
    }
    __syncthreads();
    const float* skip = p.X + row0 * H;
    const int relu_out = p.relu_out;
    unstage_rows<float, float, TR>(p.OUT + row0 * H, bufA, rv, H, [=](float g, int r, int n) {
      const float v = g + skip[(size_t)r * H + n];
      return relu_out ? fmaxf(v, 0.f) : v;
    });
  }
}

static int launch_glu(GluParams p, cudaStream_t st) {
  if (p.rows <= 0) S5_FAIL(S5_E_BADSHAPE, "glu: rows=%lld must be positive", p.rows);
  const int K1 = p.wc.K, H = p.wc.N;
  const size_t wbytes = (wsmem_words_host(p.wc) + (p.ssm_only ? 0 : wsmem_words_host(p.wo))) * 4;
  const size_t per_row = (size_t)(std::max(K1, H) + 2 * H) * 4;
  const size_t zbytes = p.skip_zero ? (size_t)std::max(K1, H) * 4 : 0;
  const int tr = pick_tr(wbytes + zbytes, per_row);
  if (!tr) S5_FAIL(S5_E_BADSHAPE, "glu: H=%d 2P=%d weights (%zu B) do not fit in shared memory", H, K1, wbytes);
  const size_t smem = wbytes + per_row * (tr + 1) + zbytes;
  const long long ntiles = (p.rows + tr - 1) / tr;
  int grid = 1, rc;
  ProfScope prof(S5_K_GLU, st);
  switch (tr) {
    case 128:
      if ((rc = persistent_grid(glu_f32_kernel<128>, smem, ntiles, &grid))) return rc;
      glu_f32_kernel<128><<<grid, kProjThreads, smem, st>>>(p);
      break;
    case 64:
      if ((rc = persistent_grid(glu_f32_kernel<64>, smem, ntiles, &grid))) return rc;
      glu_f32_kernel<64><<<grid, kProjThreads, smem, st>>>(p);
      break;
    default:
      if ((rc = persistent_grid(glu_f32_kernel<32>, smem, ntiles, &grid))) return rc;
      glu_f32_kernel<32><<<grid, kProjThreads, smem, st>>>(p);
  }
  S5_LAUNCH_CHECK();
  return S5_OK;
}

}  // namespace s5

// ------------------------------------------------------------------------------------------- handles
namespace s5 {

struct LayerDev {
  SparseDev wb, wc, wo;
  TcLinearDev tb;  // dense-tile form of the B projection (static flavour)
  TcGluDev tg;     // dense-tile form of the C projection + GLU
  TcF32Dev fb, fc, fo;   // float flavour: 3 x TF32 dense-tile forms of B_bar, C_tilde (real form) and the out2 kernel
  float *lam = nullptr, *D = nullptr, *bn_mean = nullptr, *bn_mul = nullptr, *bn_bias = nullptr, *o_bias = nullptr;
  s5_layer_desc sc;  // scales only (pointers unused)
};

static int upload(const float* h, size_t n, float** d) {
  if (!h) S5_FAIL(S5_E_BADARG, "model_create: null parameter array");
  S5_CUDA(cudaMalloc(d, std::max<size_t>(1, n) * sizeof(float)));
  S5_CUDA(cudaMemcpy(*d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return S5_OK;
}

}  // namespace s5

struct s5_model {
  s5_model_desc desc;
  s5::SparseDev enc, dec;
  s5::TcLinearDev enc_tc, dec_tc;
  s5::TcF32Dev enc_f32, dec_f32;   // float flavour
  long long tc_min_rows = 512;   // below this the compressed CUDA-core kernels are used (one 3751-frame clip: 141 us dense-tile, 312 us CUDA-core)
  bool fuse_scan = true;         // static dense-tile path: run the recurrence inside the GLU kernel (S5B200_FUSE_SCAN=0 disables)
  bool fuse_bproj = true;        // ... and the B projection as well: one kernel per block (S5B200_FUSE_BPROJ=0 disables)
  bool fuse_layers = true;       // ... and all blocks in one launch (S5B200_FUSE_LAYERS=0 disables)
  int f32_tc = 3;                // float flavour on 3 x TF32 dense tiles: bit 0 encoder / decoder, bit 1 the blocks (S5B200_F32_TC)
  bool skip_zero = false;        // CUDA-core kernels skip all-zero (input column, 32-row group) pairs of ReLU-ified inputs (S5B200_SKIP_ZERO=1
                                 // enables; exact, but measured slower on B200 - the kernels are issue bound, not multiply bound)
  float *enc_bias = nullptr, *dec_bias = nullptr;
  std::vector<s5::LayerDev> layers;
  // host-buffer entry point: ring of device buffers, three streams (H2D | forward | D2H) and their events
  static constexpr int kSlots = 5;
  std::mutex mu;
  float* hx[kSlots] = {}; float* hy[kSlots] = {};
  void* hws[kSlots] = {};
  void* ha[kSlots] = {};              // audio mode: audio in | magnitude | angle | audio out of the slot's chunk
  size_t hx_cap[kSlots] = {}, hy_cap[kSlots] = {}, hws_cap[kSlots] = {}, ha_cap[kSlots] = {};
  cudaStream_t hs[4] = {nullptr, nullptr, nullptr, nullptr};   // H2D, D2H, two forward streams
  cudaEvent_t ev_h2d[kSlots] = {}, ev_cmp[kSlots] = {}, ev_d2h[kSlots] = {}, ev_user = nullptr, ev_done = nullptr;
  unsigned long long chunk_seq = 0;
  void* last_user = nullptr;          // caller stream of the previous forward_host call
  bool done_recorded = false;
};

namespace s5 {

static void qrange(int bits, float* lo, float* hi) {
  *lo = -(float)(1 << (bits - 1));
  *hi = (float)((1 << (bits - 1)) - 1);
}

struct WsLayout {
  size_t off_a, off_b, off_bu, off_scan, scan_bytes, total;
};
static WsLayout ws_layout(const s5_model* m, int B, int L) {
  const size_t rows = (size_t)B * L;
  WsLayout w;
  const size_t act = round_up(rows * m->desc.H * sizeof(float), 256);
  w.off_a = 0;
  w.off_b = act;
  w.off_bu = 2 * act;
  // + 32 rows of slack: the fused recurrence reads whole 32-frame segments without clamping at the end of the buffer
  w.off_scan = w.off_bu + round_up((rows + 32) * m->desc.P * 2 * sizeof(float), 256);
  // one look-back region per layer (all blocks may run in one launch) + the cross-layer done flags
  const size_t glu_ws = round_up(tc_glu_scan_workspace_bytes(B, L, m->desc.P), 256);
  w.scan_bytes = std::max(s5b200_scan_c64_workspace_bytes(B, L, m->desc.P), (size_t)(m->desc.n_layers + 1) * glu_ws);
  w.total = w.off_scan + round_up(w.scan_bytes, 256);
  return w;
}

static int check_model_args(const s5_model* m, const void* a, const void* b, int B, int L, void* ws, size_t wsb,
                            const char* what) {
  if (!m || !a || !b || !ws) S5_FAIL(S5_E_BADARG, "%s: null pointer", what);
  if (B <= 0 || L <= 0 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "%s: B=%d L=%d outside the supported range", what, B, L);
  const size_t need = ws_layout(m, B, L).total;
  if (wsb < need) S5_FAIL(S5_E_BADARG, "%s: workspace %zu < %zu bytes", what, wsb, need);
  if (!aligned16(ws)) S5_FAIL(S5_E_BADARG, "%s: workspace must be 16-byte aligned", what);
  return S5_OK;
}

static LinearParams bproj_params(const s5_model* m, const LayerDev& ld, const float* x, float* bu, long long rows, bool bn) {
  const bool sq = m->desc.flavour == S5_STATIC;
  LinearParams p{};
  p.X = x; p.Y = bu; p.rows = rows; p.w = wview(ld.wb);
  if (bn) { p.in_sub = ld.bn_mean; p.in_mul = ld.bn_mul; p.in_add = ld.bn_bias; }
  p.in_quant = sq; p.in_scale = ld.sc.s_ut;
  qrange(m->desc.ssm_act_bits, &p.in_qmin, &p.in_qmax);
  p.bias = nullptr;
  p.out_quant = sq; p.out_scale0 = ld.sc.s_But_re; p.out_scale1 = ld.sc.s_But_im;
  qrange(m->desc.ssm_act_bits, &p.out_qmin, &p.out_qmax);
  p.relu = 0;
  return p;
}

static GluParams glu_params(const s5_model* m, const LayerDev& ld, const float* xs, const float* x, float* out,
                            long long rows, bool bn, bool ssm_only) {
  const bool sq = m->desc.flavour == S5_STATIC;
  GluParams p{};
  p.XS = xs; p.X = x; p.OUT = out; p.rows = rows; p.wc = wview(ld.wc); p.wo = wview(ld.wo);
  if (bn) { p.bn_sub = ld.bn_mean; p.bn_mul = ld.bn_mul; p.bn_add = ld.bn_bias; }
  p.D = ld.D; p.out2_bias = ld.o_bias;
  p.conj_mul = m->desc.conj_sym ? 2.f : 1.f;
  p.static_q = sq;
  p.s_ut = ld.sc.s_ut; p.s_yt = ld.sc.s_yt; p.s_o_act = ld.sc.s_out2_act; p.s_o_out = ld.sc.s_out2_out;
  p.s_gl = ld.sc.s_gate_left; p.s_gr = ld.sc.s_gate_right;
  qrange(m->desc.ssm_act_bits, &p.qs_min, &p.qs_max);
  qrange(m->desc.non_ssm_act_bits, &p.qn_min, &p.qn_max);
  p.act = m->desc.activation; p.relu_out = m->desc.relufication; p.ssm_only = ssm_only;
  // the states are ReLU-ified only with relufication; x1 >= 0 is sparse only behind a ReLU activation
  p.skip_zero = m->skip_zero && m->desc.relufication && m->desc.activation == S5_ACT_RELU;
  return p;
}

struct BlockPlan {
  bool glu_tc, fuse_scan, fuse_bproj;
};
static BlockPlan block_plan(const s5_model* m, const LayerDev& ld, long long rows, bool ssm_only, bool want_xs) {
  const bool sq = m->desc.flavour == S5_STATIC;
  BlockPlan bp;
  bp.glu_tc = sq && (ld.tg.ok || ld.tg.wide_ok) && !ssm_only && rows >= m->tc_min_rows && is_pow2(ld.sc.s_ut) && is_pow2(ld.sc.s_yt) &&
              is_pow2(ld.sc.s_out2_act) && is_pow2(ld.sc.s_out2_out) && is_pow2(ld.sc.s_gate_left) && is_pow2(ld.sc.s_gate_right);
  // the recurrence runs inside the GLU kernel (states never leave the SM) unless the caller wants to see them,
  // and so does the B projection: the whole block is then one kernel
  bp.fuse_scan = bp.glu_tc && m->fuse_scan && !want_xs && tc_glu_scan_ok(ld.tg);
  bp.fuse_bproj = bp.fuse_scan && m->fuse_bproj && ld.tg.bpj_ok && is_pow2(ld.sc.s_But_re) && is_pow2(ld.sc.s_But_im);
  return bp;
}
static TcGluArgs glu_args(const s5_model* m, const LayerDev& ld, const BlockPlan& bp, const float* x, float* out, float* bu, int B,
                          int L, void* scan_ws, size_t scan_ws_bytes, bool bn) {
  TcGluArgs a{};
  a.XS = bu; a.X = x; a.OUT = out; a.rows = (long long)B * L;
  if (bp.fuse_scan) {
    a.lam = ld.lam; a.B = B; a.L = L; a.relu_scan = m->desc.relufication; a.scan_ws = scan_ws; a.scan_ws_bytes = scan_ws_bytes;
    a.fuse_bproj = bp.fuse_bproj; a.s_bu_re = ld.sc.s_But_re; a.s_bu_im = ld.sc.s_But_im;
  }
  a.bn_sub = bn ? ld.bn_mean : nullptr; a.bn_mul = bn ? ld.bn_mul : nullptr; a.bn_add = bn ? ld.bn_bias : nullptr;
  a.D = ld.D; a.o_bias = ld.o_bias;
  a.s_ut = ld.sc.s_ut; a.s_yt = ld.sc.s_yt; a.s_o_act = ld.sc.s_out2_act; a.s_o_out = ld.sc.s_out2_out;
  a.s_gl = ld.sc.s_gate_left; a.s_gr = ld.sc.s_gate_right;
  qrange(m->desc.ssm_act_bits, &a.qs_min, &a.qs_max);
  qrange(m->desc.non_ssm_act_bits, &a.qn_min, &a.qn_max);
  a.relu_out = m->desc.relufication;
  return a;
}

// One S5 block: x [rows][H] -> out [rows][H]; bu is the [rows][P] complex scratch (Bu, then xs in place).
static int run_block(const s5_model* m, int li, const float* x, float* out, float* bu, int B, int L, void* scan_ws,
                     size_t scan_ws_bytes, bool bn, bool ssm_only, cudaStream_t st, bool want_xs = false, bool unfused = false) {
  const LayerDev& ld = m->layers[li];
  const long long rows = (long long)B * L;
  int rc;
  const bool sq = m->desc.flavour == S5_STATIC;
  if (!sq && !ssm_only && (m->f32_tc & 2) && rows >= m->tc_min_rows && ld.fb.ok && ld.fc.ok && ld.fo.ok) {
    // float flavour, throughput batch: B projection | scan | C projection + SSM tail | out2 + gate, the three products
    // as 3 x TF32 dense tiles (tc_f32.cu); the gate kernel reads x1 and writes the block output in place
    TcF32Args a;
    a.X = x; a.Y = bu; a.rows = rows;
    if (bn) { a.in_sub = ld.bn_mean; a.in_mul = ld.bn_mul; a.in_add = ld.bn_bias; }
    if ((rc = tc_f32_launch(ld.fb, a, S5_K_BPROJ, st))) return rc;
    if ((rc = launch_scan_c64(bu, ld.lam, bu, B, L, m->desc.P, m->desc.relufication, scan_ws, scan_ws_bytes, st))) return rc;
    {   // C projection + SSM tail + out2 + gate in one launch when one M block holds H (tc_f32g_kernel)
      TcF32GluArgs g;
      g.XS = bu; g.X = x; g.Y = out; g.rows = rows; g.D = ld.D; g.o_bias = ld.o_bias; g.relu = m->desc.relufication;
      if (bn) { g.bn_sub = ld.bn_mean; g.bn_mul = ld.bn_mul; g.bn_add = ld.bn_bias; }
      g.conj_mul = m->desc.conj_sym ? 2.f : 1.f; g.act = m->desc.activation;
      rc = tc_f32_glu_launch(ld.fc, ld.fo, g, S5_K_GLU, st);
      if (rc != S5_E_UNSUPPORTED) return rc;
    }
    TcF32Args c;
    c.X = bu; c.Y = out; c.rows = rows; c.epi = 1; c.X2 = x; c.D = ld.D;
    if (bn) { c.bn_sub = ld.bn_mean; c.bn_mul = ld.bn_mul; c.bn_add = ld.bn_bias; }
    c.conj_mul = m->desc.conj_sym ? 2.f : 1.f; c.act = m->desc.activation;
    if ((rc = tc_f32_launch(ld.fc, c, S5_K_GLU, st))) return rc;
    TcF32Args o;
    o.X = out; o.Y = out; o.rows = rows; o.epi = 2; o.bias = ld.o_bias; o.X1 = out; o.SKIP = x; o.relu = m->desc.relufication;
    return tc_f32_launch(ld.fo, o, S5_K_GLU, st);
  }
  BlockPlan bp = block_plan(m, ld, rows, ssm_only, want_xs);
  if (unfused) bp.fuse_scan = bp.fuse_bproj = false;          // retry after a refused cooperative launch
  if (bp.fuse_bproj) {
    // nothing to launch here
  } else if (sq && ld.tb.ok && rows >= m->tc_min_rows && is_pow2(ld.sc.s_ut) && is_pow2(ld.sc.s_But_re) && is_pow2(ld.sc.s_But_im)) {
    float qlo, qhi;
    qrange(m->desc.ssm_act_bits, &qlo, &qhi);
    if ((rc = tc_linear_launch(ld.tb, x, bu, rows, bn ? ld.bn_mean : nullptr, bn ? ld.bn_mul : nullptr, bn ? ld.bn_bias : nullptr,
                               ld.sc.s_ut, qlo, qhi, nullptr, ld.sc.s_But_re, ld.sc.s_But_im, qlo, qhi, 0, S5_K_BPROJ, st)))
      return rc;
  } else if ((rc = launch_linear(bproj_params(m, ld, x, bu, rows, bn), st, S5_K_BPROJ))) return rc;
  if (!bp.fuse_scan && (rc = launch_scan_c64(bu, ld.lam, bu, B, L, m->desc.P, m->desc.relufication, scan_ws, scan_ws_bytes, st))) return rc;
  if (bp.glu_tc) {
    rc = tc_glu_launch(ld.tg, glu_args(m, ld, bp, x, out, bu, B, L, scan_ws, scan_ws_bytes, bn), st);
    // the fused recurrence needs a cooperative launch; where the device refuses it (SM limit under MPS, ...) the
    // separate B projection | scan_c64 | GLU kernels compute the same block
    if (rc == kCoopRefused && bp.fuse_scan && !unfused)
      return run_block(m, li, x, out, bu, B, L, scan_ws, scan_ws_bytes, bn, ssm_only, st, want_xs, true);
    return rc;
  }
  return launch_glu(glu_params(m, ld, bu, x, out, rows, bn, ssm_only), st);
}

static LinearParams dense_params(const s5_model* m, const SparseDev& w, const float* bias, const float* x, float* y,
                                 long long rows, float s_act, float s_out, int relu) {
  const bool sq = m->desc.flavour == S5_STATIC;
  LinearParams p{};
  p.X = x; p.Y = y; p.rows = rows; p.w = wview(w);
  p.in_quant = sq; p.in_scale = s_act;
  qrange(m->desc.non_ssm_act_bits, &p.in_qmin, &p.in_qmax);
  p.bias = bias;
  p.out_quant = sq; p.out_scale0 = p.out_scale1 = s_out;
  qrange(m->desc.non_ssm_act_bits, &p.out_qmin, &p.out_qmax);
  p.relu = relu;
  p.skip_zero = m->skip_zero && m->desc.relufication && relu == 0;   // decoder input: the ReLU-ified block output
  return p;
}

}  // namespace s5

extern "C" int s5b200_sparse_create(const void* W_host, int K, int N, int dtype, int dense_format, s5_sparse** out) {
  if (!out) S5_FAIL(S5_E_BADARG, "sparse_create: null out");
  s5_sparse* h = new s5_sparse();
  int rc = s5::sparse_build(W_host, K, N, dtype, dense_format, &h->d);
  if (!rc && dtype == S5_F32) rc = s5::tc_f32_build(reinterpret_cast<const float*>(W_host), K, N, &h->tc);
  if (rc) { s5::sparse_free(&h->d); s5::tc_f32_free(&h->tc); delete h; return rc; }
  *out = h;
  return S5_OK;
}
extern "C" int s5b200_sparse_destroy(s5_sparse* w) {
  if (!w) return S5_OK;
  s5::sparse_free(&w->d);
  s5::tc_f32_free(&w->tc);
  delete w;
  return S5_OK;
}
extern "C" int s5b200_sparse_nnz(const s5_sparse* w) { return w ? w->d.nnz : 0; }
extern "C" int s5b200_sparse_export(const s5_sparse* w, int32_t* k_idx_host, int32_t* n_idx_host) {
  if (!w || !k_idx_host || !n_idx_host) S5_FAIL(S5_E_BADARG, "sparse_export: null pointer");
  std::copy(w->d.h_k.begin(), w->d.h_k.end(), k_idx_host);
  std::copy(w->d.h_n.begin(), w->d.h_n.end(), n_idx_host);
  return S5_OK;
}
extern "C" int s5b200_spmm_f32(const s5_sparse* w, const float* X, const float* bias, float* Y, long long rows,
                               int skip_zero_cols, void* stream) {
  if (!w || !X || !Y) S5_FAIL(S5_E_BADARG, "spmm_f32: null pointer");
  if (w->d.dtype != S5_F32) S5_FAIL(S5_E_BADARG, "spmm_f32: handle holds int32 weights");
  // throughput batches: the projection is a genuine dense contraction -> 3 x TF32 dense tiles on the tensor cores
  // (S5B200_TC_MIN_ROWS, default 512); below that, and when zero skipping is asked for, the compressed CUDA-core kernel
  static const long long tc_min = [] { const char* e = getenv("S5B200_TC_MIN_ROWS"); return e ? atoll(e) : 512LL; }();
  if (w->tc.ok && rows >= tc_min && !skip_zero_cols) {
    s5::TcF32Args a;
    a.X = X; a.Y = Y; a.rows = rows; a.bias = bias;
    return s5::tc_f32_launch(w->tc, a, S5_K_OTHER, (cudaStream_t)stream);
  }
  s5::LinearParams p{};
  p.X = X; p.Y = Y; p.rows = rows; p.w = s5::wview(w->d); p.bias = bias; p.skip_zero = skip_zero_cols != 0;
  return s5::launch_linear(p, (cudaStream_t)stream);
}

extern "C" int s5b200_model_create(const s5_model_desc* d, s5_model** out) {
  using namespace s5;
  if (!d || !out || !d->layers) S5_FAIL(S5_E_BADARG, "model_create: null pointer");
  if (d->flavour != S5_FLOAT && d->flavour != S5_STATIC) S5_FAIL(S5_E_BADARG, "model_create: flavour %d", d->flavour);
  if (d->n_layers < 1 || d->n_layers > 64 || d->H < 1 || d->H > 1024 || d->P < 2 || d->P > 256 || (d->P & 1) ||
      d->d_input < 1 || d->d_input > 4096 || d->d_output < 1 || d->d_output > 4096)
    S5_FAIL(S5_E_BADSHAPE, "model_create: n_layers=%d H=%d P=%d d_input=%d d_output=%d unsupported", d->n_layers, d->H,
            d->P, d->d_input, d->d_output);
  if (d->flavour == S5_STATIC && (d->ssm_act_bits < 2 || d->ssm_act_bits > 24 || d->non_ssm_act_bits < 2 || d->non_ssm_act_bits > 24))
    S5_FAIL(S5_E_BADARG, "model_create: activation bit widths must be in [2, 24]");
  s5_model* m = new s5_model();
  m->desc = *d;
  m->desc.layers = nullptr;
  // S5B200_TC_MIN_ROWS: row count from which the dense-tile (tcgen05) kernels replace the compressed ones
  // (1 forces them, a huge value disables them); default 4096.
  if (const char* e = getenv("S5B200_TC_MIN_ROWS")) m->tc_min_rows = atoll(e);
  if (const char* e = getenv("S5B200_FUSE_SCAN")) m->fuse_scan = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FUSE_BPROJ")) m->fuse_bproj = atoi(e) != 0;
  if (const char* e = getenv("S5B200_FUSE_LAYERS")) m->fuse_layers = atoi(e) != 0;
  if (const char* e = getenv("S5B200_SKIP_ZERO")) m->skip_zero = atoi(e) != 0;
  if (const char* e = getenv("S5B200_F32_TC")) m->f32_tc = atoi(e);
  const int H = d->H, P = d->P, dense = d->use_sparse ? 0 : 1;
  int rc = S5_OK;
  auto fail = [&](int code) { s5b200_model_destroy(m); return code; };
  if ((rc = sparse_build(d->enc_kernel_host, d->d_input, H, S5_F32, dense, &m->enc))) return fail(rc);
  if ((rc = sparse_build(d->dec_kernel_host, H, d->d_output, S5_F32, dense, &m->dec))) return fail(rc);
  if ((rc = upload(d->enc_bias_host, H, &m->enc_bias))) return fail(rc);
  if ((rc = upload(d->dec_bias_host, d->d_output, &m->dec_bias))) return fail(rc);
  if (d->flavour == S5_STATIC) {
    if ((rc = tc_linear_build(d->enc_kernel_host, d->d_input, H, d->s_enc_act, &m->enc_tc))) return fail(rc);
    if ((rc = tc_linear_build(d->dec_kernel_host, H, d->d_output, d->s_dec_act, &m->dec_tc))) return fail(rc);
  } else {
    if ((rc = tc_f32_build(d->enc_kernel_host, d->d_input, H, &m->enc_f32))) return fail(rc);
    if ((rc = tc_f32_build(d->dec_kernel_host, H, d->d_output, &m->dec_f32))) return fail(rc);
  }
  m->layers.resize(d->n_layers);
  for (int i = 0; i < d->n_layers; ++i) {
    const s5_layer_desc& l = d->layers[i];
    LayerDev& ld = m->layers[i];
    ld.sc = l;
    if (!l.B_bar_host || !l.C_tilde_host) return fail((set_error("model_create: null SSM matrix"), S5_E_BADARG));
    // B: [P][H][2] -> W[h][2p + c];  C: [H][P][2] -> W[2p][h] = C_re, W[2p+1][h] = -C_im  (Re(C x))
    std::vector<float> wb((size_t)H * 2 * P), wc((size_t)2 * P * H);
    for (int p = 0; p < P; ++p)
      for (int h = 0; h < H; ++h) {
        wb[(size_t)h * 2 * P + 2 * p] = l.B_bar_host[((size_t)p * H + h) * 2];
        wb[(size_t)h * 2 * P + 2 * p + 1] = l.B_bar_host[((size_t)p * H + h) * 2 + 1];
        wc[(size_t)(2 * p) * H + h] = l.C_tilde_host[((size_t)h * P + p) * 2];
        wc[(size_t)(2 * p + 1) * H + h] = -l.C_tilde_host[((size_t)h * P + p) * 2 + 1];
      }
    if ((rc = sparse_build(wb.data(), H, 2 * P, S5_F32, dense, &ld.wb))) return fail(rc);
    if (d->flavour == S5_STATIC && (rc = tc_linear_build(wb.data(), H, 2 * P, l.s_ut, &ld.tb))) return fail(rc);
    if (d->flavour == S5_STATIC && d->activation == S5_ACT_RELU && l.out2_kernel_host &&
        (rc = tc_glu_build(l.C_tilde_host, l.out2_kernel_host, H, P, d->conj_sym ? 2.f : 1.f, l.s_out2_act, wb.data(), l.s_ut, &ld.tg)))
      return fail(rc);
    if ((rc = sparse_build(wc.data(), 2 * P, H, S5_F32, dense, &ld.wc))) return fail(rc);
    if ((rc = sparse_build(l.out2_kernel_host, H, H, S5_F32, dense, &ld.wo))) return fail(rc);
    if (d->flavour == S5_FLOAT) {
      if ((rc = tc_f32_build(wb.data(), H, 2 * P, &ld.fb))) return fail(rc);
      if ((rc = tc_f32_build(wc.data(), 2 * P, H, &ld.fc))) return fail(rc);
      if (l.out2_kernel_host && (rc = tc_f32_build(l.out2_kernel_host, H, H, &ld.fo))) return fail(rc);
    }
    if ((rc = upload(l.lambda_bar_host, 2 * P, &ld.lam))) return fail(rc);
    if ((rc = upload(l.D_host, H, &ld.D))) return fail(rc);
    if ((rc = upload(l.bn_mean_host, H, &ld.bn_mean))) return fail(rc);
    if ((rc = upload(l.bn_mul_host, H, &ld.bn_mul))) return fail(rc);
    if ((rc = upload(l.bn_bias_host, H, &ld.bn_bias))) return fail(rc);
    if ((rc = upload(l.out2_bias_host, H, &ld.o_bias))) return fail(rc);
    if (d->flavour == S5_STATIC && ld.tg.ok) {
      // the block kernel's per-layer image is built here, once, so that the forward neither allocates, synchronises
      // nor touches the handle (run_block and the single launch pass exactly these vectors and scales)
      const BlockPlan all{true, true, true};
      if ((rc = tc_glu_prepare_image(ld.tg, glu_args(m, ld, all, nullptr, nullptr, nullptr, 1, 1, nullptr, 0, true)))) return fail(rc);
    }
  }
  *out = m;
  return S5_OK;
}

extern "C" int s5b200_model_destroy(s5_model* m) {
  if (!m) return S5_OK;
  s5::sparse_free(&m->enc); s5::sparse_free(&m->dec);
  s5::tc_linear_free(&m->enc_tc); s5::tc_linear_free(&m->dec_tc);
  s5::tc_f32_free(&m->enc_f32); s5::tc_f32_free(&m->dec_f32);
  cudaFree(m->enc_bias); cudaFree(m->dec_bias);
  for (auto& l : m->layers) {
    s5::sparse_free(&l.wb); s5::sparse_free(&l.wc); s5::sparse_free(&l.wo);
    s5::tc_linear_free(&l.tb);
    s5::tc_glu_free(&l.tg);
    s5::tc_f32_free(&l.fb); s5::tc_f32_free(&l.fc); s5::tc_f32_free(&l.fo);
    cudaFree(l.lam); cudaFree(l.D); cudaFree(l.bn_mean); cudaFree(l.bn_mul); cudaFree(l.bn_bias); cudaFree(l.o_bias);
  }
  for (int i = 0; i < s5_model::kSlots; ++i) {
    cudaFree(m->hx[i]); cudaFree(m->hy[i]); cudaFree(m->hws[i]); cudaFree(m->ha[i]);
    if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
    if (m->ev_cmp[i]) cudaEventDestroy(m->ev_cmp[i]);
    if (m->ev_d2h[i]) cudaEventDestroy(m->ev_d2h[i]);
  }
  if (m->ev_user) cudaEventDestroy(m->ev_user);
  if (m->ev_done) cudaEventDestroy(m->ev_done);
  for (int i = 0; i < 4; ++i)
    if (m->hs[i]) cudaStreamDestroy(m->hs[i]);
  delete m;
  return S5_OK;
}

extern "C" size_t s5b200_model_workspace_bytes(const s5_model* m, int B, int L) {
  if (!m || B <= 0 || L <= 0) return 0;
  return s5::ws_layout(m, B, L).total;
}

extern "C" int s5b200_model_forward(const s5_model* m, const float* x, float* y, int B, int L, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  using namespace s5;
  int rc;
  if ((rc = check_model_args(m, x, y, B, L, workspace, workspace_bytes, "model_forward"))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const WsLayout w = ws_layout(m, B, L);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* a = reinterpret_cast<float*>(base + w.off_a);
  float* b = reinterpret_cast<float*>(base + w.off_b);
  float* bu = reinterpret_cast<float*>(base + w.off_bu);
  const long long rows = (long long)B * L;
  const s5_model_desc& d = m->desc;
  const bool tc = d.flavour == S5_STATIC && rows >= m->tc_min_rows;
  float nlo = 0.f, nhi = 0.f;
  if (tc) qrange(d.non_ssm_act_bits, &nlo, &nhi);
  if (tc && m->enc_tc.ok && is_pow2(d.s_enc_act) && is_pow2(d.s_enc_out)) {
    if ((rc = tc_linear_launch(m->enc_tc, x, a, rows, nullptr, nullptr, nullptr, d.s_enc_act, nlo, nhi, m->enc_bias, d.s_enc_out,
                               d.s_enc_out, nlo, nhi, d.relufication, S5_K_ENC, st))) return rc;
  } else if (d.flavour == S5_FLOAT && (m->f32_tc & 1) && rows >= m->tc_min_rows && m->enc_f32.ok) {
    TcF32Args e;
    e.X = x; e.Y = a; e.rows = rows; e.bias = m->enc_bias; e.relu = d.relufication;
    if ((rc = tc_f32_launch(m->enc_f32, e, S5_K_ENC, st))) return rc;
  } else if ((rc = launch_linear(dense_params(m, m->enc, m->enc_bias, x, a, rows, d.s_enc_act, d.s_enc_out, d.relufication), st, S5_K_ENC))) return rc;
  // all blocks in ONE launch when every layer runs the block kernel (S5B200_FUSE_LAYERS=0: one launch per block)
  bool one_launch = m->fuse_layers && d.n_layers >= 2 && d.n_layers <= 4;
  for (int i = 0; i < d.n_layers && one_launch; ++i)
    one_launch = block_plan(m, m->layers[i], rows, false, false).fuse_bproj && m->layers[i].tg.H == m->layers[0].tg.H &&
                 m->layers[i].tg.K1 == m->layers[0].tg.K1;
  if (one_launch) {
    const size_t glu_ws = round_up(tc_glu_scan_workspace_bytes(B, L, d.P), 256);
    const TcGluDev* devs[4];
    TcGluArgs args[4];
    float *pa = a, *pb = b;
    for (int i = 0; i < d.n_layers; ++i) {
      const LayerDev& ld = m->layers[i];
      devs[i] = &ld.tg;
      args[i] = glu_args(m, ld, block_plan(m, ld, rows, false, false), pa, pb, bu, B, L, base + w.off_scan + (size_t)i * glu_ws, glu_ws, true);
      std::swap(pa, pb);
    }
    rc = tc_glu_launch_multi(devs, args, d.n_layers, reinterpret_cast<uint32_t*>(base + w.off_scan + (size_t)d.n_layers * glu_ws),
                             glu_ws, st);
    if (rc == kCoopRefused) one_launch = false;               // per-block launches below (each falls back further on its own)
    else if (rc) return rc;
    else { a = pa; b = pb; }
  }
  if (!one_launch) {
    for (int i = 0; i < d.n_layers; ++i) {
      if ((rc = run_block(m, i, a, b, bu, B, L, base + w.off_scan, w.scan_bytes, true, false, st))) return rc;
      std::swap(a, b);
    }
  }
  if (tc && m->dec_tc.ok && is_pow2(d.s_dec_act) && is_pow2(d.s_dec_out))
    return tc_linear_launch(m->dec_tc, a, y, rows, nullptr, nullptr, nullptr, d.s_dec_act, nlo, nhi, m->dec_bias, d.s_dec_out,
                            d.s_dec_out, nlo, nhi, 0, S5_K_DEC, st);
  if (d.flavour == S5_FLOAT && (m->f32_tc & 1) && rows >= m->tc_min_rows && m->dec_f32.ok) {
    TcF32Args e;
    e.X = a; e.Y = y; e.rows = rows; e.bias = m->dec_bias;
    return tc_f32_launch(m->dec_f32, e, S5_K_DEC, st);
  }
  return launch_linear(dense_params(m, m->dec, m->dec_bias, a, y, rows, d.s_dec_act, d.s_dec_out, 0), st, S5_K_DEC);
}

extern "C" int s5b200_layer_forward(const s5_model* m, int layer, const float* x, float* y, float* xs_out, int B, int L,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  using namespace s5;
  int rc;
  if ((rc = check_model_args(m, x, y, B, L, workspace, workspace_bytes, "layer_forward"))) return rc;
  if (layer < 0 || layer >= m->desc.n_layers) S5_FAIL(S5_E_BADARG, "layer_forward: layer %d out of range", layer);
  const WsLayout w = ws_layout(m, B, L);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* bu = xs_out ? xs_out : reinterpret_cast<float*>(base + w.off_bu);
  if (!aligned16(bu)) S5_FAIL(S5_E_BADARG, "layer_forward: xs_out must be 16-byte aligned");
  return run_block(m, layer, x, y, bu, B, L, base + w.off_scan, w.scan_bytes, true, false, (cudaStream_t)stream, xs_out != nullptr);
}

extern "C" int s5b200_ssm_forward(const s5_model* m, int layer, const float* u, float* y, float* xs_out, int B, int L,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  using namespace s5;
  int rc;
  if ((rc = check_model_args(m, u, y, B, L, workspace, workspace_bytes, "ssm_forward"))) return rc;
  if (layer < 0 || layer >= m->desc.n_layers) S5_FAIL(S5_E_BADARG, "ssm_forward: layer %d out of range", layer);
  const WsLayout w = ws_layout(m, B, L);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* bu = xs_out ? xs_out : reinterpret_cast<float*>(base + w.off_bu);
  if (!aligned16(bu)) S5_FAIL(S5_E_BADARG, "ssm_forward: xs_out must be 16-byte aligned");
  return run_block(m, layer, u, y, bu, B, L, base + w.off_scan, w.scan_bytes, false, true, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------- calibration pass
extern "C" int s5b200_calibrate_sites(const s5_model* m) { return m ? 4 + 9 * m->desc.n_layers : 0; }

extern "C" int s5b200_model_calibrate(const s5_model* m, const float* x, float* y, float* minmax, int B, int L, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  using namespace s5;
  int rc;
  if ((rc = check_model_args(m, x, y, B, L, workspace, workspace_bytes, "model_calibrate"))) return rc;
  if (!minmax) S5_FAIL(S5_E_BADARG, "model_calibrate: null minmax");
  const s5_model_desc& d = m->desc;
  if (d.flavour != S5_FLOAT) S5_FAIL(S5_E_BADARG, "model_calibrate: needs a float-flavour handle (the calibrating graph passes values through unchanged)");
  if (2 * d.P < d.H) S5_FAIL(S5_E_UNSUPPORTED, "model_calibrate: needs 2P >= H (scratch reuse)");
  cudaStream_t st = (cudaStream_t)stream;
  const WsLayout w = ws_layout(m, B, L);
  unsigned char* base = reinterpret_cast<unsigned char*>(workspace);
  float* a = reinterpret_cast<float*>(base + w.off_a);
  float* b = reinterpret_cast<float*>(base + w.off_b);
  float* bu = reinterpret_cast<float*>(base + w.off_bu);
  const long long rows = (long long)B * L;
  const int sites = s5b200_calibrate_sites(m);
  uint32_t* mm = reinterpret_cast<uint32_t*>(minmax);
  const int grid = (int)std::min<long long>((rows * 32 + 255) / 256, (long long)sm_count() * 8);
  ProfScope prof(S5_K_OTHER, st);
  calib_init_kernel<<<(sites + 127) / 128, 128, 0, st>>>(mm, sites);
  S5_LAUNCH_CHECK();
  auto observe = [&](int site, const float* t, long long n, int stride, int offset, const LayerDev* bn) -> int {
    calib_minmax_kernel<<<grid, 256, 0, st>>>(t, n, stride, offset, d.H, bn ? bn->bn_mean : nullptr, bn ? bn->bn_mul : nullptr,
                                              bn ? bn->bn_bias : nullptr, mm + 2 * site);
    S5_LAUNCH_CHECK();
    return S5_OK;
  };
  auto ew = [&](int mode, float* t, const float* u, const float* v, long long n) -> int {
    calib_ew_kernel<<<grid, 256, 0, st>>>(mode, t, u, v, n);
    S5_LAUNCH_CHECK();
    return S5_OK;
  };
  // encoder QuantizedDense (quantization.py:730-745, 793-811: input and output observers), then ReLU (seq_model.py:100-115)
  if ((rc = observe(0, x, rows * d.d_input, 1, 0, nullptr))) return rc;
  if ((rc = launch_linear(dense_params(m, m->enc, m->enc_bias, x, a, rows, 1.f, 1.f, 0), st, S5_K_ENC))) return rc;
  if ((rc = observe(1, a, rows * d.H, 1, 0, nullptr))) return rc;
  if (d.relufication && (rc = ew(CAL_RELU, a, nullptr, nullptr, rows * d.H))) return rc;
  for (int i = 0; i < d.n_layers; ++i) {
    const LayerDev& ld = m->layers[i];
    const int s0 = 2 + 9 * i;
    // u = BN(h) (quant_ut); Bu (quant_But); states before the ReLU (quant_xt); y = ys + D u (quant_yt)      ssm.py:230-337, 594-601
    if ((rc = observe(s0 + 0, a, rows * d.H, 1, 0, &ld))) return rc;
    if ((rc = launch_linear(bproj_params(m, ld, a, bu, rows, true), st, S5_K_BPROJ))) return rc;
    if ((rc = observe(s0 + 1, bu, rows * d.P, 2, 0, nullptr)) || (rc = observe(s0 + 2, bu, rows * d.P, 2, 1, nullptr))) return rc;
    if ((rc = launch_scan_c64(bu, ld.lam, bu, B, L, d.P, 0, base + w.off_scan, w.scan_bytes, st))) return rc;
    if ((rc = observe(s0 + 3, bu, rows * d.P, 2, 0, nullptr)) || (rc = observe(s0 + 4, bu, rows * d.P, 2, 1, nullptr))) return rc;
    if (d.relufication && (rc = ew(CAL_CRELU, bu, nullptr, nullptr, rows * d.P))) return rc;
    if ((rc = launch_glu(glu_params(m, ld, bu, a, b, rows, true, true), st))) return rc;          // ssm_only: y before the activation
    if ((rc = observe(s0 + 5, b, rows * d.H, 1, 0, nullptr))) return rc;
    if ((rc = ew(d.activation == S5_ACT_RELU ? CAL_RELU : CAL_GELU, b, nullptr, nullptr, rows * d.H))) return rc;
    // out2 QuantizedDense on x1, sigmoid, mult_gate observers (layers.py:217-238)
    if ((rc = observe(s0 + 6, b, rows * d.H, 1, 0, nullptr))) return rc;
    LinearParams o{};
    o.X = b; o.Y = bu; o.rows = rows; o.w = wview(ld.wo); o.bias = ld.o_bias;
    if ((rc = launch_linear(o, st, S5_K_GLU))) return rc;
    if ((rc = observe(s0 + 7, bu, rows * d.H, 1, 0, nullptr))) return rc;
    if ((rc = ew(CAL_SIGMOID, bu, nullptr, nullptr, rows * d.H))) return rc;
    if ((rc = observe(s0 + 8, bu, rows * d.H, 1, 0, nullptr))) return rc;
    if ((rc = ew(d.relufication ? CAL_GATE_RELU : CAL_GATE, b, bu, a, rows * d.H))) return rc;      // x1 * sg + skip (+ ReLU)
    std::swap(a, b);
  }
  if ((rc = observe(sites - 2, a, rows * d.H, 1, 0, nullptr))) return rc;
  if ((rc = launch_linear(dense_params(m, m->dec, m->dec_bias, a, y, rows, 1.f, 1.f, 0), st, S5_K_DEC))) return rc;
  if ((rc = observe(sites - 1, y, rows * d.d_output, 1, 0, nullptr))) return rc;
  calib_decode_kernel<<<(2 * sites + 127) / 128, 128, 0, st>>>(mm, 2 * sites);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

// NDNS pre / post-processing (ndns.cu)
extern "C" int s5b200_stft_frames(int T);
extern "C" int s5b200_stft(const float* audio, int B, int T, float* mag, float* phase, int layout_blf, float* model_in, float mag_sub, void* stream);
extern "C" int s5b200_istft(const float* mag, const float* phase, int layout_blf, const float* mask_blf, int B, int L, float* audio_out,
                            float* cleaned_mag_out, void* stream);

// audio_T > 0: AUDIO mode (s5b200_ndns_denoise_host) - x_host is noisy audio [B][audio_T], y_host the cleaned audio [B][(L - 1) * 128],
// L = s5b200_stft_frames(audio_T); a chunk's forward is STFT -> model -> mask apply + iSTFT and spectra never cross the bus.
static int forward_host_impl(const s5_model* mc, const float* x_host, float* y_host, int B, int L, void* stream, bool sync, int audio_T = 0,
                             float mag_sub = 0.f) {
  // End-to-end entry point: the batch is cut into clip chunks that flow through a three-stage pipeline
  // (H2D copy | forward | D2H copy) on four internal streams (two alternating forward streams) with a ring of kSlots device buffers, so that
  // the two PCIe directions and the kernels overlap.  Clips are independent, so chunking does not change results.
  using namespace s5;
  s5_model* m = const_cast<s5_model*>(mc);
  if (!m || !x_host || !y_host) S5_FAIL(S5_E_BADARG, "model_forward_host: null pointer");
  if (B <= 0 || L <= 0 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "model_forward_host: B=%d L=%d outside the supported range", B, L);
  cudaStream_t user = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(m->mu);
  constexpr int kSlots = s5_model::kSlots;
  // Chunk schedule (S5B200_E2E_CHUNKS=n forces n equal chunks, "a,b,c" explicit clip counts).
  std::vector<int> sched;
  const char* e = getenv("S5B200_E2E_CHUNKS");
  if (e && !*e) e = nullptr;                     // an empty value means "not set"
  if (e && strchr(e, ',')) {                    // explicit clip counts, e.g. "6,6,6,6,4,4"; the rest goes into a last chunk
    int used = 0;
    for (const char* q = e; *q && used < B;) {
      const int v = std::max(1, std::min(B - used, atoi(q)));
      sched.push_back(v);
      used += v;
      q = strchr(q, ',');
      if (!q) break;
      ++q;
    }
    if (used < B) sched.push_back(B - used);
  } else if (e) {
    const int n = std::max(1, std::min(B, atoi(e)));
    const int cbu = (B + n - 1) / n;
    for (int b0 = 0; b0 < B; b0 += cbu) sched.push_back(std::min(cbu, B - b0));
  } else if (!sync && m->done_recorded && m->last_user != stream && cudaEventQuery(m->ev_done) == cudaErrorNotReady) {
    // The pipeline is already busy with the previous call, queued on ANOTHER caller stream (batches in flight on alternating
    // streams): this call's first copy overlaps that call's tail, so the ramp-up below has nothing to hide and only costs
    // small copies and small forwards.  Chunks of about 64 MB of input (16 NDNS clips) then run closest to the copy-only
    // ceiling: 2.68-2.73 ms per 32-clip batch against 2.99 with the ramp and 2.59 copy-only (profiles/r02_e2e_pipelined.md).
    const size_t clip_bytes = (size_t)L * m->desc.d_input * sizeof(float);
    const int per = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)64 << 20) / std::max<size_t>(1, clip_bytes)));
    const int n = (B + per - 1) / per;
    int rem = B;
    for (int i = 0; i < n; ++i) {
      const int take = (rem + (n - i) - 1) / (n - i);
      sched.push_back(take);
      rem -= take;
    }
  } else {
    // The device-to-host direction is the bottleneck of the pipeline (measured: ~42 GB/s under bidirectional load),
    // so it should start as early as possible: ramp up 1, 1, 2, 4 clips, then chunks of about 6 that amortise the
    // fixed latency of a forward (11 launches, ~0.3 ms) while two forward streams keep the GPU busy.
    int rem = B;
    for (int t : {1, 1, 2, 4})
      if (rem > 2 * t) { sched.push_back(t); rem -= t; }
    const int nbody = std::max(1, (rem + 5) / 6);
    for (int i = 0; i < nbody; ++i) {
      const int take = (rem + (nbody - i) - 1) / (nbody - i);
      sched.push_back(take);
      rem -= take;
    }
  }
  const int nchunks = (int)sched.size();
  const int cb = *std::max_element(sched.begin(), sched.end());  // largest chunk sizes the slot buffers
  const size_t xb = (size_t)cb * L * m->desc.d_input * sizeof(float), yb = (size_t)cb * L * m->desc.d_output * sizeof(float);
  const size_t wsb = ws_layout(m, cb, L).total;
  auto grow = [&](void** p, size_t* cap, size_t need) -> int {
    if (*cap >= need) return S5_OK;
    if (*p) S5_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    S5_CUDA(cudaMalloc(p, need));
    *cap = need;
    return S5_OK;
  };
  int rc;
  if (!m->hs[0]) {
    for (int i = 0; i < 4; ++i) S5_CUDA(cudaStreamCreateWithFlags(&m->hs[i], cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
      S5_CUDA(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
      S5_CUDA(cudaEventCreateWithFlags(&m->ev_cmp[i], cudaEventDisableTiming));
      S5_CUDA(cudaEventCreateWithFlags(&m->ev_d2h[i], cudaEventDisableTiming));
    }
    S5_CUDA(cudaEventCreateWithFlags(&m->ev_user, cudaEventDisableTiming));
    S5_CUDA(cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
  }
  const size_t out_len = audio_T > 0 ? (size_t)(L - 1) * 128 : 0;
  // audio-mode scratch of a slot: audio in | magnitude | angle | audio out, each rounded up to 256 bytes
  const size_t a_in_b = round_up((size_t)cb * (size_t)std::max(audio_T, 0) * sizeof(float), 256), a_sp_b = round_up(xb, 256);
  const size_t a_out_b = round_up((size_t)cb * out_len * sizeof(float), 256);
  for (int i = 0; i < kSlots; ++i) {
    if ((rc = grow(reinterpret_cast<void**>(&m->hx[i]), &m->hx_cap[i], xb))) return rc;
    if ((rc = grow(reinterpret_cast<void**>(&m->hy[i]), &m->hy_cap[i], yb))) return rc;
    if ((rc = grow(&m->hws[i], &m->hws_cap[i], wsb))) return rc;
    if (audio_T > 0 && (rc = grow(&m->ha[i], &m->ha_cap[i], a_in_b + 2 * a_sp_b + a_out_b))) return rc;
  }
  cudaStream_t s_in = m->hs[0], s_out = m->hs[1];
  S5_CUDA(cudaEventRecord(m->ev_user, user));  // order after whatever the caller queued before
  S5_CUDA(cudaStreamWaitEvent(s_in, m->ev_user, 0));
  // S5B200_E2E_TRACE=1: print the device timeline of the pipeline (debug aid; adds timing events)
  const bool trace = getenv("S5B200_E2E_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  std::vector<double> thost;
  const auto host_t0 = std::chrono::steady_clock::now();
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    tev.push_back(ev);
    thost.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
  };
  mark(s_in);
  int b0 = 0;
  for (int c = 0; c < nchunks; b0 += sched[c], ++c) {
    const int nb = sched[c];
    const unsigned long long seq = m->chunk_seq++;   // slots and forward streams rotate across calls, too
    const int slot = (int)(seq % kSlots);
    cudaStream_t s_cmp = m->hs[2 + (seq & 1)];   // consecutive chunks run concurrently: a chunk alone does not fill the GPU
    const size_t xoff = (size_t)b0 * L * m->desc.d_input, yoff = (size_t)b0 * L * m->desc.d_output;
    const size_t xbytes = (size_t)nb * L * m->desc.d_input * sizeof(float), ybytes = (size_t)nb * L * m->desc.d_output * sizeof(float);
    if (seq >= (unsigned long long)kSlots) S5_CUDA(cudaStreamWaitEvent(s_in, m->ev_d2h[slot], 0));  // slot buffers are free again
    float *a_in = nullptr, *a_mag = nullptr, *a_ph = nullptr, *a_out = nullptr;
    if (audio_T > 0) {
      unsigned char* ab = reinterpret_cast<unsigned char*>(m->ha[slot]);
      a_in = reinterpret_cast<float*>(ab);
      a_mag = reinterpret_cast<float*>(ab + a_in_b);
      a_ph = reinterpret_cast<float*>(ab + a_in_b + a_sp_b);
      a_out = reinterpret_cast<float*>(ab + a_in_b + 2 * a_sp_b);
    }
    if (audio_T > 0) {
      S5_CUDA(cudaMemcpyAsync(a_in, x_host + (size_t)b0 * audio_T, (size_t)nb * audio_T * sizeof(float), cudaMemcpyHostToDevice, s_in));
    } else {
      S5_CUDA(cudaMemcpyAsync(m->hx[slot], x_host + xoff, xbytes, cudaMemcpyHostToDevice, s_in));
    }
    S5_CUDA(cudaEventRecord(m->ev_h2d[slot], s_in));
    mark(s_in);
    S5_CUDA(cudaStreamWaitEvent(s_cmp, m->ev_h2d[slot], 0));
    mark(s_cmp);
    if (audio_T > 0 && (rc = s5b200_stft(a_in, nb, audio_T, a_mag, a_ph, 1, m->hx[slot], mag_sub, s_cmp))) return rc;
    if (!(trace && getenv("S5B200_E2E_TRACE_COPY_ONLY")) &&   // debug: measure the copy pipeline alone
        (rc = s5b200_model_forward(m, m->hx[slot], m->hy[slot], nb, L, m->hws[slot], m->hws_cap[slot], s_cmp))) return rc;
    if (audio_T > 0 && (rc = s5b200_istft(a_mag, a_ph, 1, m->hy[slot], nb, L, a_out, nullptr, s_cmp))) return rc;
    S5_CUDA(cudaEventRecord(m->ev_cmp[slot], s_cmp));
    mark(s_cmp);
    S5_CUDA(cudaStreamWaitEvent(s_out, m->ev_cmp[slot], 0));
    if (audio_T > 0) {
      S5_CUDA(cudaMemcpyAsync(y_host + (size_t)b0 * out_len, a_out, (size_t)nb * out_len * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    } else {
      S5_CUDA(cudaMemcpyAsync(y_host + yoff, m->hy[slot], ybytes, cudaMemcpyDeviceToHost, s_out));
    }
    S5_CUDA(cudaEventRecord(m->ev_d2h[slot], s_out));
    mark(s_out);
  }
  // the caller's stream continues after the last device-to-host copy (which is after everything else)
  S5_CUDA(cudaEventRecord(m->ev_done, s_out));
  S5_CUDA(cudaStreamWaitEvent(user, m->ev_done, 0));
  m->done_recorded = true;
  m->last_user = stream;
  if (sync || trace) {
    S5_CUDA(cudaStreamSynchronize(s_out));
    S5_CUDA(cudaStreamSynchronize(m->hs[2]));
    S5_CUDA(cudaStreamSynchronize(m->hs[3]));
    S5_CUDA(cudaStreamSynchronize(s_in));
  }
  if (trace) {
    const double host_end = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    fprintf(stderr, "[s5b200 e2e trace] chunk clips | h2d done | fwd start..done | d2h done (ms, device) | host enqueue done (ms)\n");
    for (int c = 0; c < nchunks; ++c) {
      float t[4];
      for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[1 + 4 * c + j]);
      fprintf(stderr, "[s5b200 e2e trace] %2d %3d | %7.3f | %7.3f .. %7.3f | %7.3f | %7.3f\n", c, sched[c], t[0], t[1], t[2], t[3], thost[4 + 4 * c]);
    }
    fprintf(stderr, "[s5b200 e2e trace] host returned at %.3f ms\n", host_end);
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
  }
  return S5_OK;
}

extern "C" int s5b200_model_forward_host(const s5_model* m, const float* x_host, float* y_host, int B, int L, void* stream) {
  return forward_host_impl(m, x_host, y_host, B, L, stream, true);
}

extern "C" int s5b200_model_forward_host_async(const s5_model* m, const float* x_host, float* y_host, int B, int L,
                                               void* stream) {
  return forward_host_impl(m, x_host, y_host, B, L, stream, false);
}

static int denoise_host(const s5_model* m, const float* noisy_host, float* cleaned_host, int B, int T, float mag_sub, void* stream, bool sync) {
  if (!m) S5_FAIL(S5_E_BADARG, "ndns_denoise_host: null model");
  if (m->desc.d_input != 257 || m->desc.d_output != 257)
    S5_FAIL(S5_E_BADSHAPE, "ndns_denoise_host: the model maps %d -> %d features, the NDNS spectra have 257 bins", m->desc.d_input, m->desc.d_output);
  if (T <= 0) S5_FAIL(S5_E_BADSHAPE, "ndns_denoise_host: T=%d must be positive", T);
  const int L = s5b200_stft_frames(T);                        // >= 2 for every positive T (boundary padding of nfft / 2 either side)
  if (L < 2) S5_FAIL(S5_E_BADSHAPE, "ndns_denoise_host: T=%d gives %d frame(s)", T, L);
  return forward_host_impl(m, noisy_host, cleaned_host, B, L, stream, sync, T, mag_sub);
}
extern "C" int s5b200_ndns_denoise_host(const s5_model* m, const float* noisy_host, float* cleaned_host, int B, int T, float mag_sub,
                                        void* stream) {
  return denoise_host(m, noisy_host, cleaned_host, B, T, mag_sub, stream, true);
}
extern "C" int s5b200_ndns_denoise_host_async(const s5_model* m, const float* noisy_host, float* cleaned_host, int B, int T, float mag_sub,
                                              void* stream) {
  return denoise_host(m, noisy_host, cleaned_host, B, T, mag_sub, stream, false);
}
