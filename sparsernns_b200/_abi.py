"""ctypes binding of libs5b200.so (the C ABI declared in include/s5b200.h).

There is no CPU fallback: if the shared library is missing or a call fails this module raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libs5b200.so")

S5_OK = 0
ROUND_FLOOR, ROUND_CEIL, ROUND_HALF_UP = 0, 1, 2
FLOAT, STATIC = 0, 1
ACT_GELU_TANH, ACT_RELU = 0, 1
F32, I32 = 0, 1
KERNEL_CLASSES = ["enc", "bproj", "scan_c64", "glu", "dec", "fxp_enc", "fxp_bn", "fxp_bproj", "scan_i32", "fxp_glu",
                  "fxp_resid", "fxp_dec", "other", "glu_scan", "block"]

FXP_SYNC_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)   # s5_fxp_sync_fn
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class S5Error(RuntimeError):
    pass


class LayerDesc(C.Structure):
    _fields_ = [
        ("lambda_bar_host", c_float_p), ("B_bar_host", c_float_p), ("C_tilde_host", c_float_p), ("D_host", c_float_p),
        ("bn_mean_host", c_float_p), ("bn_mul_host", c_float_p), ("bn_bias_host", c_float_p),
        ("out2_kernel_host", c_float_p), ("out2_bias_host", c_float_p),
        ("s_ut", C.c_float), ("s_But_re", C.c_float), ("s_But_im", C.c_float), ("s_yt", C.c_float),
        ("s_out2_act", C.c_float), ("s_out2_out", C.c_float), ("s_gate_left", C.c_float), ("s_gate_right", C.c_float),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("flavour", C.c_int), ("n_layers", C.c_int), ("H", C.c_int), ("P", C.c_int), ("d_input", C.c_int),
        ("d_output", C.c_int), ("activation", C.c_int), ("relufication", C.c_int), ("conj_sym", C.c_int),
        ("ssm_act_bits", C.c_int), ("non_ssm_act_bits", C.c_int), ("use_sparse", C.c_int),
        ("enc_kernel_host", c_float_p), ("enc_bias_host", c_float_p), ("s_enc_act", C.c_float), ("s_enc_out", C.c_float),
        ("dec_kernel_host", c_float_p), ("dec_bias_host", c_float_p), ("s_dec_act", C.c_float), ("s_dec_out", C.c_float),
        ("layers", C.POINTER(LayerDesc)),
    ]


class FxpDenseDesc(C.Structure):
    _fields_ = [
        ("weight_host", c_int32_p), ("bias_host", c_int32_p),
        ("w_exp", C.c_int), ("b_exp", C.c_int), ("b_bits", C.c_int), ("inp_bits", C.c_int), ("inp_exp", C.c_int),
        ("out_bits", C.c_int), ("out_exp", C.c_int),
    ]


class FxpLayerDesc(C.Structure):
    _fields_ = [
        ("bn_minus_mean_host", c_int32_p), ("bn_mean_exp", C.c_int), ("bn_mean_bits", C.c_int),
        ("bn_invsq_var_host", c_int32_p), ("bn_invsq_exp", C.c_int), ("bn_invsq_bits", C.c_int),
        ("bn_scale_host", c_int32_p), ("bn_scale_exp", C.c_int), ("bn_scale_bits", C.c_int),
        ("bn_bias_host", c_int32_p), ("bn_bias_exp", C.c_int), ("bn_bias_bits", C.c_int),
        ("A_re_host", c_int32_p), ("A_im_host", c_int32_p), ("B_re_host", c_int32_p), ("B_im_host", c_int32_p),
        ("C_re_host", c_int32_p), ("C_im_host", c_int32_p), ("D_host", c_int32_p),
        ("eA_re", C.c_int), ("eA_im", C.c_int), ("eB_re", C.c_int), ("eB_im", C.c_int), ("eC_re", C.c_int),
        ("eC_im", C.c_int), ("eD", C.c_int),
        ("u_bits", C.c_int), ("u_exp", C.c_int), ("Bu_re_bits", C.c_int), ("Bu_re_exp", C.c_int),
        ("Bu_im_bits", C.c_int), ("Bu_im_exp", C.c_int), ("x_re_exp", C.c_int), ("x_im_exp", C.c_int),
        ("y_bits", C.c_int), ("y_exp", C.c_int),
        ("out2", FxpDenseDesc),
        ("sig_x_exp", C.c_int), ("sig_y_exp", C.c_int), ("sig_lut", C.c_int32 * 8),
        ("gate_l_bits", C.c_int), ("gate_l_exp", C.c_int), ("gate_r_bits", C.c_int), ("gate_r_exp", C.c_int),
        ("gate_res_bits", C.c_int), ("gate_res_exp", C.c_int),
    ]


class FxpModelDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int), ("H", C.c_int), ("P", C.c_int), ("d_input", C.c_int), ("d_output", C.c_int),
        ("use_sparse", C.c_int),
        ("encoder", FxpDenseDesc), ("decoder", FxpDenseDesc),
        ("layers", C.POINTER(FxpLayerDesc)),
    ]


_V = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/s5b200.h declares
PROTOTYPES = {
    "s5b200_version": (_I, []),
    "s5b200_last_error": (C.c_char_p, []),
    "s5b200_launch_count": (_LL, []),
    "s5b200_reset_launch_count": (None, []),
    "s5b200_profile_begin": (_I, []),
    "s5b200_profile_end": (_I, [C.POINTER(C.c_double), C.POINTER(_LL)]),
    "s5b200_scan_c64_workspace_bytes": (_SZ, [_I, _I, _I]),
    "s5b200_scan_c64": (_I, [_V, _V, _V, _I, _I, _I, _I, _V, _SZ, _V]),
    "s5b200_scan_i32": (_I, [_V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V]),
    "s5b200_sparse_create": (_I, [_V, _I, _I, _I, _I, C.POINTER(_V)]),
    "s5b200_sparse_destroy": (_I, [_V]),
    "s5b200_sparse_nnz": (_I, [_V]),
    "s5b200_sparse_export": (_I, [_V, _V, _V]),
    "s5b200_spmm_f32": (_I, [_V, _V, _V, _V, _LL, _I, _V]),
    "s5b200_spmm_i32": (_I, [_V, _V, _V, _LL, _I, _I, _I, _I, _V]),
    "s5b200_tc_gemm_f16": (_I, [_V, _V, _V, _LL, _I, _I, _V]),
    "s5b200_tc_gemm_i8": (_I, [_V, _V, _V, _LL, _I, _I, _I, _V]),
    "s5b200_model_create": (_I, [C.POINTER(ModelDesc), C.POINTER(_V)]),
    "s5b200_model_destroy": (_I, [_V]),
    "s5b200_model_workspace_bytes": (_SZ, [_V, _I, _I]),
    "s5b200_model_forward": (_I, [_V, _V, _V, _I, _I, _V, _SZ, _V]),
    "s5b200_layer_forward": (_I, [_V, _I, _V, _V, _V, _I, _I, _V, _SZ, _V]),
    "s5b200_ssm_forward": (_I, [_V, _I, _V, _V, _V, _I, _I, _V, _SZ, _V]),
    "s5b200_model_forward_host": (_I, [_V, _V, _V, _I, _I, _V]),
    "s5b200_model_forward_host_async": (_I, [_V, _V, _V, _I, _I, _V]),
    "s5b200_fxp_model_create": (_I, [C.POINTER(FxpModelDesc), C.POINTER(_V)]),
    "s5b200_fxp_model_destroy": (_I, [_V]),
    "s5b200_fxp_model_workspace_bytes": (_SZ, [_V, _I, _I]),
    "s5b200_fxp_model_forward": (_I, [_V, _V, _V, _V, _I, _I, _V, _SZ, _V, _V]),
    "s5b200_fxp_model_forward_int": (_I, [_V, _V, _V, _V, _I, _I, _V, _SZ, _V, _V]),
    "s5b200_fxp_model_tap": (_I, [_V, _V, _I, _I, _V, C.POINTER(_I), _I, _I, _V, _SZ, _V]),
    "s5b200_fxp_from_fp": (_I, [_V, _V, _LL, _I, _I, _I, _V]),
    "s5b200_fxp_change_cfg": (_I, [_V, _V, _LL, _I, _I, _I, _I, _I, _V]),
    "s5b200_fxp_add": (_I, [_V, _V, _V, _LL, _LL, _I, _I, _I, _I, _I, _I, _I, _V]),
    "s5b200_fxp_mul": (_I, [_V, _V, _V, _LL, _LL, _I, _I, _I, _I, _I, _V]),
    "s5b200_fxp_add_compute_best": (_I, [_V, _V, _V, _LL, _LL, _I, _I, _I, _I, _I, C.POINTER(_I), _V]),
    "s5b200_fxp_mul_compute_best": (_I, [_V, _V, _V, _LL, _LL, _I, _I, _I, C.POINTER(_I), _V]),
    "s5b200_fxp_relu": (_I, [_V, _V, _LL, _V]),
    "s5b200_fxp_sigmoid": (_I, [_V, _V, _LL, _I, _I, _I, _I, _V, _V]),
    "s5b200_stft_frames": (_I, [_I]),
    "s5b200_stft": (_I, [_V, _I, _I, _V, _V, _I, _V, C.c_float, _V]),
    "s5b200_istft": (_I, [_V, _V, _I, _V, _I, _I, _V, _V, _V]),
    "s5b200_ndns_denoise_host": (_I, [_V, _V, _V, _I, _I, C.c_float, _V]),
    "s5b200_ndns_denoise_host_async": (_I, [_V, _V, _V, _I, _I, C.c_float, _V]),
    "s5b200_reduce_workspace_bytes": (_SZ, [_I, _LL]),
    "s5b200_si_snr": (_I, [_V, _V, _I, _LL, _V, _V, _SZ, _V]),
    "s5b200_mse_clips": (_I, [_V, _V, _I, _LL, _V, _V, _SZ, _V]),
    "s5b200_fxp_model_set_maxima_sync": (_I, [_V, _V, _V]),
    "s5b200_calibrate_sites": (_I, [_V]),
    "s5b200_model_calibrate": (_I, [_V, _V, _V, _V, _I, _I, _V, _SZ, _V]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libs5b200.so (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S5Error(f"{LIB_PATH} is missing: run `python -m sparsernns_b200.build` (needs nvcc); "
                          "there is no CPU fallback")
        h = C.CDLL(LIB_PATH)
        missing = [n for n in PROTOTYPES if not hasattr(h, n)]
        if missing:
            raise S5Error(f"{LIB_PATH} does not export {missing}: header and library disagree, rebuild it")
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status: int, what: str = "") -> None:
    if status != S5_OK:
        msg = lib().s5b200_last_error().decode("utf-8", "replace")
        raise S5Error(f"{what or 'libs5b200'} failed with status {status}: {msg}")
