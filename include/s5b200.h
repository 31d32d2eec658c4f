/*
 * s5b200.h - C ABI of libs5b200.so: B200-native (sm_100a) S5 / NDNS forward path.
 *
 * The reference (stevenabreu7/SparseRNNs) has no FFI or plugin interface: the seams of this path are
 * Python callables.  Each entry point below names the reference callable it
 * replaces (paths relative to the reference repository root) and is what a `jax.ffi` custom call /
 * ctypes stub binds (INTEGRATION.md shows both bindings).
 *
 * Conventions
 *   - plain C, no C++/torch types; all tensors are caller-owned, row-major, DEVICE pointers unless the
 *     argument name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, never synchronises
 *     the device and never allocates on the hot path (scratch comes from caller-provided workspaces or
 *     from the handle);
 *   - return value: 0 on success, negative S5_E_* otherwise; `s5b200_last_error()` returns a
 *     thread-local message for the last failing call on this thread;
 *   - complex values are interleaved (re, im) pairs, i.e. the memory layout of complex64;
 *   - handles are immutable after creation and may be used concurrently from several streams.
 */
#ifndef S5B200_H_
#define S5B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define S5B200_VERSION 100 /* major*10000 + minor*100 + patch */

enum s5_status {
  S5_OK = 0,
  S5_E_BADSHAPE = -1,    /* dimension out of the supported range / inconsistent sizes */
  S5_E_BADARG = -2,      /* null pointer, misaligned pointer, bad enum value, workspace too small */
  S5_E_CUDA = -3,        /* a CUDA runtime call failed (message carries cudaGetErrorString) */
  S5_E_UNSUPPORTED = -4  /* configuration the reference cannot reach either (e.g. bidirectional) */
};

/* fxparray.py:13-17 RoundingMode */
enum s5_round_mode { S5_ROUND_FLOOR = 0, S5_ROUND_CEIL = 1, S5_ROUND_HALF_UP = 2 };

/* numerics flavour of a model handle */
enum s5_flavour {
  S5_FLOAT = 0,  /* quantization=none: build_apply_ssm, nn.Dense (ssm.py:80-187) */
  S5_STATIC = 1  /* static fake-quant, frozen scales: build_apply_ssm_static_quant (ssm.py:190-345) */
};

enum s5_activation { S5_ACT_GELU_TANH = 0, S5_ACT_RELU = 1 };

int s5b200_version(void);
const char* s5b200_last_error(void);
/* Number of kernels launched by this library on the calling thread since the last reset (bench.py's
 * `gpu_launches`). */
long long s5b200_launch_count(void);
void s5b200_reset_launch_count(void);

/* Per-kernel-class device timing (CUDA events recorded on the launching stream around every kernel this
 * library launches on the calling thread).  begin() clears and enables, end() synchronises the recorded
 * events, accumulates milliseconds and launch counts per class into the S5_K_COUNT-long arrays, and
 * disables.  Meant for bench.py's roofline figures; adds two event records per launch while enabled. */
enum s5_kernel_class {
  S5_K_ENC = 0, S5_K_BPROJ = 1, S5_K_SCAN_C64 = 2, S5_K_GLU = 3, S5_K_DEC = 4,
  S5_K_FXP_ENC = 5, S5_K_FXP_BN = 6, S5_K_FXP_BPROJ = 7, S5_K_SCAN_I32 = 8, S5_K_FXP_GLU = 9,
  S5_K_FXP_RESID = 10, S5_K_FXP_DEC = 11, S5_K_OTHER = 12,
  S5_K_GLU_SCAN = 13,  /* recurrence + C projection + GLU in one kernel */
  S5_K_BLOCK = 14,     /* B projection + recurrence + C projection + GLU: one kernel per S5 block */
  S5_K_COUNT = 15
};
int s5b200_profile_begin(void);
int s5b200_profile_end(double* ms_per_class, long long* launches_per_class);

/* ------------------------------------------------------------------------------------------------
 * (1) Diagonal complex linear recurrence  x_k = lam * x_{k-1} + Bu_k,  x_{-1} = 0.
 * Replaces jax.lax.associative_scan / jax.lax.scan in sparseRNNs/model/ssm.py:127, :150, :313.
 *   Bu, xs : [B][L][P] complex64 (xs may alias Bu);  lam : [P] complex64.
 *   relu != 0 applies the lexicographic complex ReLU of ssm.py:159/:325 to the OUTPUT only.
 * Chunked single-pass scan with decoupled look-back; `workspace` holds the chunk carries/flags.
 * ---------------------------------------------------------------------------------------------- */
size_t s5b200_scan_c64_workspace_bytes(int B, int L, int P);
int s5b200_scan_c64(const float* Bu, const float* lam, float* xs, int B, int L, int P, int relu,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3a) Integer recurrence, bit-exact with make_ssm_step_fn / recurrent_loop
 * (sparseRNNs/fxpmodel.py:147-208): per-product arithmetic >>, wrapping int32 state, no saturation.
 *   Bu_re, Bu_im, xs_re, xs_im : [B][L][P] int32 (planar, as the reference's FxpArray payloads);
 *   A_re, A_im : [P] int32.  relu != 0 applies fxp_relu (fxpmodel.py:27-52) to the output.
 * ---------------------------------------------------------------------------------------------- */
int s5b200_scan_i32(const int32_t* Bu_re, const int32_t* Bu_im, const int32_t* A_re, const int32_t* A_im,
                    int32_t* xs_re, int32_t* xs_im, int B, int L, int P, int eA_re, int eA_im,
                    int eBu_re, int eBu_im, int ex_re, int ex_im, int relu, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Sparse projection  Y[rows][N] = X[rows][K] @ W[K][N]  with W in compressed form.
 * Replaces jnp.dot / lax.dot_general in ssm.py:116-124, :179-185, quantization.py:772-776 and
 * `op1.data @ op2.data` in fxparray.py:662.  The non-zero index set is taken from the matrix that
 * is handed in (the already-quantised weights), so it is bit-exact by construction;
 * s5b200_sparse_export returns it for checking.
 * ---------------------------------------------------------------------------------------------- */
typedef struct s5_sparse s5_sparse;
enum s5_dtype { S5_F32 = 0, S5_I32 = 1 };
/* W_host: dense [K][N] row-major, float32 or int32.  dense_format != 0 keeps every entry. */
int s5b200_sparse_create(const void* W_host, int K, int N, int dtype, int dense_format, s5_sparse** out);
int s5b200_sparse_destroy(s5_sparse* w);
int s5b200_sparse_nnz(const s5_sparse* w);
/* k_idx, n_idx : host arrays of s5b200_sparse_nnz() entries, ordered by (n, k). */
int s5b200_sparse_export(const s5_sparse* w, int32_t* k_idx_host, int32_t* n_idx_host);
/* Y = X @ W (+ bias[N] if non-null).  skip_zero_cols != 0 enables tile-level zero-activation skipping. */
int s5b200_spmm_f32(const s5_sparse* w, const float* X, const float* bias, float* Y, long long rows,
                    int skip_zero_cols, void* stream);
/* fxp_matmul (fxparray.py:640-678): int32 accumulate (wraps), rshift-round, saturate to `bits`. */
int s5b200_spmm_i32(const s5_sparse* w, const int32_t* X, int32_t* Y, long long rows, int rshift,
                    int round_mode, int bits, int skip_zero_cols, void* stream);

/* Dense-tile projection primitive on the tcgen05 tensor cores (throughput batches, where the projection is a
 * genuine dense contraction): Y[rows][N] (f32) = A[rows][K] (f16) * W[N][K]^T (f16), fp32 accumulation in TMEM.
 * K and N multiples of 16, N <= 256.  Integer codes of up to 11 bits are exact in f16, so int8 weights times
 * hi/lo-split int16 activations give exact integer sums ("dense-tile path"). */
int s5b200_tc_gemm_f16(const void* A, const void* W, float* Y, long long rows, int K, int N, void* stream);
/* The integer form the model kernels use (tcgen05 kind::i8, int32 accumulation in TMEM):
 * Y[rows][N] (int32) = A[rows][K] (int8 if a_signed else uint8) * W[N][K]^T (int8).  K multiple of 32, N of 16, N <= 256. */
int s5b200_tc_gemm_i8(const void* A, const void* W, int32_t* Y, long long rows, int K, int N, int a_signed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (1)+(2) Float / static-quant model.  Host code (Python) keeps the reference's parameter pytree,
 * performs the per-weight-set work (ZOH discretisation ssm.py:37-50, weight fake-quant
 * ssm.py:231-233, BatchNorm constant folding of running stats) and hands flat arrays over.
 * ---------------------------------------------------------------------------------------------- */
typedef struct s5_layer_desc {
  /* SSM (ssm.py:525-622).  For S5_STATIC these are the fake-quantised values. */
  const float* lambda_bar_host; /* [P][2]    */
  const float* B_bar_host;      /* [P][H][2] */
  const float* C_tilde_host;    /* [H][P][2] */
  const float* D_host;          /* [H]       */
  /* BatchNorm, running statistics (layers.py:146-152, :210-211): u = (x - mean) * mul + bias */
  const float* bn_mean_host;    /* [H] */
  const float* bn_mul_host;     /* [H] rsqrt(var + eps) * gamma */
  const float* bn_bias_host;    /* [H] beta (zeros if absent) */
  /* GLU "half1" (layers.py:223-225): x1 * sigmoid(out2(x1)) */
  const float* out2_kernel_host; /* [H][H] (in, out); fake-quantised for S5_STATIC */
  const float* out2_bias_host;   /* [H]; fake-quantised for S5_STATIC */
  /* frozen power-of-two scales, S5_STATIC only (ssm.py:408-421, quantization.py:699-702, layers.py:168-174) */
  float s_ut, s_But_re, s_But_im, s_yt;
  float s_out2_act, s_out2_out, s_gate_left, s_gate_right;
} s5_layer_desc;

typedef struct s5_model_desc {
  int flavour;       /* enum s5_flavour */
  int n_layers, H, P, d_input, d_output;
  int activation;    /* enum s5_activation: GELU(tanh) if relufication is off, ReLU if on */
  int relufication;  /* ReLU after encoder, on xs (complex) and on the block output */
  int conj_sym;      /* y = 2*Re(C x) */
  int ssm_act_bits, non_ssm_act_bits; /* S5_STATIC activation widths (16 for w8a16) */
  int use_sparse;    /* build compressed weights from the non-zeros; 0 keeps dense tiles */
  const float* enc_kernel_host; /* [d_input][H] */
  const float* enc_bias_host;   /* [H] */
  float s_enc_act, s_enc_out;   /* S5_STATIC */
  const float* dec_kernel_host; /* [H][d_output] */
  const float* dec_bias_host;   /* [d_output] */
  float s_dec_act, s_dec_out;   /* S5_STATIC */
  const s5_layer_desc* layers;  /* [n_layers] */
} s5_model_desc;

typedef struct s5_model s5_model;
int s5b200_model_create(const s5_model_desc* desc, s5_model** out);
int s5b200_model_destroy(s5_model* m);
size_t s5b200_model_workspace_bytes(const s5_model* m, int B, int L);
/* QBatchRegressionModel.apply (seq_model.py:439-473): x [B][L][d_input] -> y [B][L][d_output]. */
int s5b200_model_forward(const s5_model* m, const float* x, float* y, int B, int L, void* workspace,
                         size_t workspace_bytes, void* stream);
/* QSequenceLayer.__call__ (layers.py:180-241) for block `layer`: x [B][L][H] -> y [B][L][H];
 * xs_out (optional, may be NULL) receives pre_C = xs [B][L][P] complex64 (after ReLU if relufication). */
int s5b200_layer_forward(const s5_model* m, int layer, const float* x, float* y, float* xs_out, int B,
                         int L, void* workspace, size_t workspace_bytes, void* stream);
/* qS5SSM.__call__ (ssm.py:525-622) for block `layer`: u [B][L][H] (already normalised) ->
 * y = ys + D*u [B][L][H] (quantised for S5_STATIC) and xs [B][L][P] complex64. */
int s5b200_ssm_forward(const s5_model* m, int layer, const float* u, float* y, float* xs_out, int B, int L,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Same as s5b200_model_forward with HOST buffers (pinned for full speed): the batch flows in clip chunks through an
 * internal H2D | forward | D2H pipeline.  Host-synchronous: y_host is complete on return (bench.py
 * "e2e.host_synchronous_value"). */
int s5b200_model_forward_host(const s5_model* m, const float* x_host, float* y_host, int B, int L,
                              void* stream);
/* Stream-ordered form of the same call: returns after enqueueing; `stream` is made to wait for the last
 * device-to-host copy, so y_host is complete once work queued on `stream` after this call has run (e.g. after
 * cudaStreamSynchronize(stream)).  x_host / y_host must stay valid until then.  Calls on ONE stream are ordered
 * like any CUDA work (batch i + 1 is copied in after batch i has been copied out); to keep both PCIe directions
 * busy across batches, queue consecutive batches on two alternating streams - the library then sees that the
 * previous call is still in flight and switches from its ramp-up chunk schedule to large chunks.  This is the loop the
 * end-to-end benchmark times (bench.py "e2e.value": 44 M frames/s against 32 M host-synchronous on one B200). */
int s5b200_model_forward_host_async(const s5_model* m, const float* x_host, float* y_host, int B, int L,
                                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Fixed-point model, bit-exact with sparseRNNs/fxpmodel.py driven as in fxprun.py:25-95.
 * All integer parameters are produced on the host exactly as FxpDense.setup / FxpSSM.setup /
 * FxpBatchNorm.setup do (fxp_from_fp ROUND at the qconfig exponents) and passed as int32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct s5_fxp_dense_desc {
  const int32_t* weight_host; /* [in][out] */
  const int32_t* bias_host;   /* [out] */
  int w_exp, b_exp, b_bits, inp_bits, inp_exp, out_bits, out_exp;
} s5_fxp_dense_desc;

typedef struct s5_fxp_layer_desc {
  /* FxpBatchNorm (fxpmodel.py:850-944); scale/bias may be NULL ("simple" BatchNorm) */
  const int32_t* bn_minus_mean_host; int bn_mean_exp, bn_mean_bits;
  const int32_t* bn_invsq_var_host;  int bn_invsq_exp, bn_invsq_bits;
  const int32_t* bn_scale_host;      int bn_scale_exp, bn_scale_bits;
  const int32_t* bn_bias_host;       int bn_bias_exp, bn_bias_bits;
  /* FxpSSM (fxpmodel.py:396-794) */
  const int32_t* A_re_host; const int32_t* A_im_host; /* [P] */
  const int32_t* B_re_host; const int32_t* B_im_host; /* [P][H] */
  const int32_t* C_re_host; const int32_t* C_im_host; /* [H][P] */
  const int32_t* D_host;                              /* [H] */
  int eA_re, eA_im, eB_re, eB_im, eC_re, eC_im, eD;
  int u_bits, u_exp, Bu_re_bits, Bu_re_exp, Bu_im_bits, Bu_im_exp, x_re_exp, x_im_exp, y_bits, y_exp;
  /* out2 + sigmoid LUT + mult_gate (fxpmodel.py:1075-1104) */
  s5_fxp_dense_desc out2;
  int sig_x_exp, sig_y_exp; int32_t sig_lut[8];
  int gate_l_bits, gate_l_exp, gate_r_bits, gate_r_exp, gate_res_bits, gate_res_exp;
} s5_fxp_layer_desc;

typedef struct s5_fxp_model_desc {
  int n_layers, H, P, d_input, d_output;
  int use_sparse;
  s5_fxp_dense_desc encoder, decoder;
  const s5_fxp_layer_desc* layers;
} s5_fxp_model_desc;

typedef struct s5_fxp_model s5_fxp_model;
int s5b200_fxp_model_create(const s5_fxp_model_desc* desc, s5_fxp_model** out);
int s5b200_fxp_model_destroy(s5_fxp_model* m);
size_t s5b200_fxp_model_workspace_bytes(const s5_fxp_model* m, int B, int L);
/* run_validation's hot part (fxprun.py:69-77): x float32 [B][L][d_input] is quantised with
 * fxp_from_fp(FLOOR, encoder inp bits/exp), the integer model runs, and
 *   y_int [B][L][d_output] int32 (payload, exponent = decoder out_exp) and/or
 *   y_float [B][L][d_output] = to_float() are written (either may be NULL).
 * status_dev (optional int32[4] device): [0] != 0 if an op the reference would raise on was hit
 * (negative shift in fxp_mul, fxparray.py:619-621). */
int s5b200_fxp_model_forward(const s5_fxp_model* m, const float* x, int32_t* y_int, float* y_float, int B,
                             int L, void* workspace, size_t workspace_bytes, int32_t* status_dev, void* stream);
/* Same, taking an already quantised int32 input (model(fxp_x), fxprun.py:76).  x_int must hold valid payloads of the encoder's
 * input format (inp_bits wide, i.e. what fxp_from_fp produces): the dense-tile encoder splits them into two byte planes. */
int s5b200_fxp_model_forward_int(const s5_fxp_model* m, const int32_t* x_int, int32_t* y_int, float* y_float,
                                 int B, int L, void* workspace, size_t workspace_bytes, int32_t* status_dev,
                                 void* stream);
/* Global-batch parity when clips are sharded over GPUs (option B).  `compute_best` exponents are functions
 * of float maxima over the WHOLE tensor a model sees (fxparray.py:420-448, 601-609), so a shard-local run is bit-exact with the
 * reference run on that shard only.  With a sync function installed, every kernel that produced running maxima is followed by
 * fn(user, maxima_dev, count, stream): the caller max-all-reduces `count` uint32 words in place (bit patterns of non-negative
 * floats, so an integer max is the float max), stream-ordered - e.g. ncclAllReduce(ncclMax) / torch.distributed.all_reduce.
 * 1 + 5 n_layers calls per forward (16 for NDNS); every rank then computes exactly what one process would on the union batch. */
typedef void (*s5_fxp_sync_fn)(void* user, uint32_t* maxima_dev, int count, void* stream);
int s5b200_fxp_model_set_maxima_sync(s5_fxp_model* m, s5_fxp_sync_fn fn, void* user);

/* Intermediate tap after a forward: copies the named per-layer tensor out of the workspace for
 * verification (names follow the `sow` keys of fxpmodel.py).  tap ids: */
enum s5_fxp_tap {
  S5_TAP_LAYER_INPUT = 0,  /* "ssm_input"  int32 [B][L][H]  + exponent */
  S5_TAP_PRE_S5 = 1,       /* the mixer input u = change_cfg("pre_s5") (fxpmodel.py:620-624) int32 [B][L][H] + exponent */
  S5_TAP_XS = 2,           /* "xs_relu"  int32 [B][L][P][2] interleaved (re, im) */
  S5_TAP_POST_GLU = 3,     /* "post_GLU" int32 [B][L][H] */
  S5_TAP_OUTPUT = 4        /* "output"   int32 [B][L][H] + exponent */
};
/* Runs the forward up to and including block `layer` and copies tap `tap` of that block to `dst`
 * (device); `exp_out_host` receives the tensor's (possibly data-dependent) exponent.  Synchronises. */
int s5b200_fxp_model_tap(const s5_fxp_model* m, const int32_t* x_int, int layer, int tap, int32_t* dst,
                         int* exp_out_host, int B, int L, void* workspace, size_t workspace_bytes, void* stream);

/* Granular fixed-point ops mirroring sparseRNNs/fxparray.py (element-wise, device int32 payloads).
 * `b` is broadcast along the last axis when nb < n (nb must divide n). */
int s5b200_fxp_from_fp(const float* x, int32_t* y, long long n, int bits, int exp, int round_mode, void* stream);          /* fxparray.py:287-307 */
int s5b200_fxp_change_cfg(const int32_t* x, int32_t* y, long long n, int bits, int exp, int new_bits, int new_exp,
                          int round_mode, void* stream);                                                                   /* fxparray.py:232-271 */
int s5b200_fxp_add(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_bits, int a_exp,
                   int b_bits, int b_exp, int result_bits, int result_exp, int round_mode, void* stream);                  /* fxparray.py:449-455 */
int s5b200_fxp_mul(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_exp, int b_exp,
                   int result_bits, int result_exp, int round_mode, void* stream);                                         /* fxparray.py:573-637 */
/* result_exp="compute_best" variants: the exponent is derived on the device from max|.| in float32
 * (fxparray.py:420-448, :601-609) and returned through result_exp_host.  Synchronises the stream. */
int s5b200_fxp_add_compute_best(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_bits,
                                int a_exp, int b_bits, int b_exp, int result_bits, int* result_exp_host, void* stream);
int s5b200_fxp_mul_compute_best(const int32_t* a, const int32_t* b, int32_t* y, long long n, long long nb, int a_exp,
                                int b_exp, int result_bits, int* result_exp_host, void* stream);
int s5b200_fxp_relu(const int32_t* x, int32_t* y, long long n, void* stream);                                              /* fxpmodel.py:53-63 */
int s5b200_fxp_sigmoid(const int32_t* x, int32_t* y, long long n, int bits, int exp, int x_exp, int y_exp,
                       const int32_t* lut8_host, void* stream);                                                            /* fxpmodel.py:70-144 */

/* ------------------------------------------------------------------------------------------------
 * (5) NDNS pre/post-processing either side of the model forward.
 * Boxcar STFT with nperseg = nfft = 512, hop 128, one-sided (257 bins), zero boundary extension,
 * scaling 1/512 - the jax.scipy.signal.stft / istft calls of sparseRNNs/train_helpers.py.
 *   layout_blf != 0: spectra are (B, L, 257) (the model's layout); 0: (B, 257, L) (what the reference's
 *   stft_splitter returns and stft_mixer takes).
 * ---------------------------------------------------------------------------------------------- */
/* frames L of the transform for T audio samples (480000 -> 3751) */
int s5b200_stft_frames(int T);
/* stft_splitter, train_helpers.py:1381-1396: audio [B][T] -> mag, phase (each nullable).  model_in (nullable,
 * always (B, L, 257)) receives mag - mag_sub, the model input of train_helpers.py:1570-1573. */
int s5b200_stft(const float* audio, int B, int T, float* mag, float* phase, int layout_blf, float* model_in,
                float mag_sub, void* stream);
/* stft_mixer, train_helpers.py:1399-1412, with the mask apply of :1605-1606 fused in: mask_blf (nullable) is the model
 * output (B, L, 257) and the spectrum magnitude becomes mag * (1 + mask); audio_out [B][(L - 1) * 128];
 * cleaned_mag_out (nullable, layout of mag, must not alias an input) receives mag * (1 + mask). */
int s5b200_istft(const float* mag, const float* phase, int layout_blf, const float* mask_blf, int B, int L,
                 float* audio_out, float* cleaned_mag_out, void* stream);
/* Forward of the reference's NDNS evaluation step (train_helpers.py:1570-1606) on HOST audio: noisy audio [B][T] (pinned for full
 * speed) -> H2D -> stft_splitter -> model input mag - mag_sub (0.0007 in the reference) -> model -> mag * (1 + mask) -> stft_mixer -> D2H
 * of the cleaned audio [B][(s5b200_stft_frames(T) - 1) * 128].  Same clip-chunk pipeline, chunk schedule and stream semantics as
 * s5b200_model_forward_host / _async; spectra never cross the bus (a quarter of the PCIe bytes of the tensor-level call).  The model
 * must map 257 -> 257 features.  Results equal s5b200_stft + s5b200_model_forward + s5b200_istft on the whole batch bit for bit. */
int s5b200_ndns_denoise_host(const s5_model* m, const float* noisy_host, float* cleaned_host, int B, int T, float mag_sub,
                             void* stream);
int s5b200_ndns_denoise_host_async(const s5_model* m, const float* noisy_host, float* cleaned_host, int B, int T,
                                   float mag_sub, void* stream);
/* workspace of the two per-clip reductions below */
size_t s5b200_reduce_workspace_bytes(int B, long long n_per_clip);
/* si_snr_jax, train_helpers.py:16-53: target, estimate [B][T] -> out_db [B] (sums accumulated in float64, fixed order) */
int s5b200_si_snr(const float* target, const float* estimate, int B, long long T, float* out_db, void* workspace,
                  size_t workspace_bytes, void* stream);
/* mean((a - b)^2) over the n_per_clip values of every clip: the magnitude term of the loss, train_helpers.py:1609-1611 */
int s5b200_mse_clips(const float* a, const float* b, int B, long long n_per_clip, float* out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (6) Static-quant calibration pass: the `calibrating=True` graph of the reference
 * (quantization.py:428-436, 730-745, 793-811; ssm.py:230-337; layers.py:217-238; driven by convert.py:881-971) on one batch.
 * Every FakeQuant / QuantizedDense site passes its value through unchanged and records min / max; `m` must be a
 * FLOAT-flavour handle of the (pruned) float weights.  minmax (device) receives s5b200_calibrate_sites(m) pairs (min, max):
 *   0 encoder input, 1 encoder output (before the ReLU);
 *   2 + 9 i + {0: u = BN(h), 1: Re Bu, 2: Im Bu, 3: Re x, 4: Im x (states before the ReLU), 5: y = ys + D u,
 *              6: x1 (out2 input = left gate), 7: out2 output, 8: sigmoid (right gate)}  for block i;
 *   sites - 2: decoder input, sites - 1: decoder output.
 * y [B][L][d_output] receives the float model output.  First-call semantics: no in-loop re-quantisation of the states (the
 * `quant_xt` observer does not exist yet when the loop is built, ssm.py:254, :314).  Scales are derived on the host
 * (quantization.py:352-370; sparsernns_b200/calibrate.py).
 * ---------------------------------------------------------------------------------------------- */
int s5b200_calibrate_sites(const s5_model* m);
int s5b200_model_calibrate(const s5_model* m, const float* x, float* y, float* minmax, int B, int L, void* workspace,
                           size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* S5B200_H_ */
