"""Host-side mirror of the reference's float / static-quant S5 NDNS model (inference forward).

Keeps the reference's call structure and parameter pytree and routes the hot path to libs5b200:

  reference                                            here
  ---------------------------------------------------  -----------------------------------------
  discretize_zoh / discretize_bilinear (ssm.py:19-50)   same names, NumPy, once per weight set
  qS5SSM.__call__ / apply_ssm (ssm.py:84-345, 525-622)  BoundModel.apply_ssm(layer, u) -> (y, xs)
  QSequenceLayer.__call__ (layers.py:180-241)           BoundModel.layer(layer, x)     -> x
  QBatchRegressionModel.apply (seq_model.py:439-473)    BoundModel(x) / BoundModel.forward_host(x)

`variables` is the Flax tree ``{"params": ..., "batch_stats": ...}``; for the
static flavour ``params`` carries the frozen scales exactly where ``sc_calibrated_params.pkl`` has
them.  Tensors are torch CUDA tensors (device memory + stream plumbing only); there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi

F32 = np.float32

# quantization.py:96-177, the recipes reachable on the NDNS path
QUANT_RECIPES = {
    "none": None,
    "w8a16": dict(a_precision=16, b_precision=8, c_precision=8, d_precision=8, non_ssm_precision=8,
                  ssm_act_precision=16, non_ssm_act_precision=16),
    "w8a8": dict(a_precision=16, b_precision=8, c_precision=8, d_precision=8, non_ssm_precision=8,
                 ssm_act_precision=8, non_ssm_act_precision=8),
}


@dataclass
class S5Config:
    n_layers: int = 3
    d_model: int = 96
    ssm_size: int = 64
    d_input: int = 257
    d_output: int = 257
    relufication: bool = False
    quantization: str = "none"       # "none" -> float flavour, otherwise static fake-quant with that recipe
    clip_eigs: bool = True
    conj_sym: bool = True
    discretization: str = "zoh"
    step_rescale: float = 1.0
    bn_eps: float = 1e-5
    glu_variant: str = "half1"
    prenorm: bool = True
    batchnorm: bool = True
    bidirectional: bool = False
    use_sparse: bool = True          # compressed weights from the non-zeros; False keeps dense tiles
    recipe: Optional[dict] = field(default=None, repr=False)

    def __post_init__(self):
        if self.quantization not in QUANT_RECIPES:
            raise ValueError(f"unknown quantization recipe {self.quantization!r}")
        self.recipe = QUANT_RECIPES[self.quantization]
        if self.glu_variant != "half1" or not (self.prenorm and self.batchnorm) or self.bidirectional:
            raise NotImplementedError("only the NDNS recipe structure is implemented: glu_variant='half1', prenorm, "
                                      "batchnorm, unidirectional (recipes/ndns.json)")

    @property
    def static_quant(self) -> bool:
        return self.recipe is not None


def discretize_zoh(Lambda, B_tilde, Delta):
    """ssm.py:37-50."""
    Lambda_bar = np.exp(Lambda * Delta)
    B_bar = (1 / Lambda * (Lambda_bar - 1))[..., None] * B_tilde
    return Lambda_bar, B_bar


def discretize_bilinear(Lambda, B_tilde, Delta):
    """ssm.py:19-34."""
    BL = 1 / (1 - (Delta / 2.0) * Lambda)
    return BL * (1 + (Delta / 2.0) * Lambda), (BL * Delta)[..., None] * B_tilde


def fake_quant(x, scale, bits: int):
    """quantization.py:386-400 (zero point 0), float32."""
    x = np.asarray(x, dtype=F32)
    s = F32(scale)
    q = np.clip(np.round(x / s), -(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
    return (q * s).astype(F32)


def ssm_discrete(mixer: dict, cfg: S5Config):
    """qS5SSM.setup (ssm.py:436-523): Lambda_bar (P,), B_bar (P,H), C_tilde (H,P) complex64, D (H,)."""
    lre = np.asarray(mixer["Lambda_re"], dtype=F32)
    lim = np.asarray(mixer["Lambda_im"], dtype=F32)
    if cfg.clip_eigs:
        lre = np.minimum(lre, F32(-1e-4))
    Lambda = (lre + 1j * lim).astype(np.complex64)
    B = np.asarray(mixer["B"], dtype=F32)
    Cm = np.asarray(mixer["C"], dtype=F32)
    B_tilde = (B[..., 0] + 1j * B[..., 1]).astype(np.complex64)
    C_tilde = (Cm[..., 0] + 1j * Cm[..., 1]).astype(np.complex64)
    step = (F32(cfg.step_rescale) * np.exp(np.asarray(mixer["log_step"], dtype=F32)[:, 0])).astype(F32)
    if cfg.discretization == "zoh":
        Lb, Bb = discretize_zoh(Lambda, B_tilde, step)
    elif cfg.discretization == "bilinear":
        Lb, Bb = discretize_bilinear(Lambda, B_tilde, step)
    else:
        raise NotImplementedError(cfg.discretization)
    return Lb.astype(np.complex64), Bb.astype(np.complex64), C_tilde, np.asarray(mixer["D"], dtype=F32)


def _c2f(z) -> np.ndarray:
    """complex array -> float32 array with a trailing (re, im) axis, C-contiguous."""
    z = np.asarray(z)
    return np.ascontiguousarray(np.stack([z.real, z.imag], axis=-1).astype(F32))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_abi.c_float_p)


class BoundModel:
    """A model handle bound to one weight set (the product of ``model.apply`` with fixed variables).

    The handle owns ONE device workspace (activations, look-back flags and records), sized on demand: calls on the same
    instance must be stream-ordered.  For forwards in flight on several streams use one ``BoundModel`` per stream (the C
    handle itself is immutable and could be shared; the workspace cannot)."""

    def __init__(self, cfg: S5Config, variables: dict):
        import torch  # device memory / stream plumbing only

        if not torch.cuda.is_available():
            raise _abi.S5Error("sparsernns_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._torch = torch
        self.cfg = cfg
        self._lib = _abi.lib()
        self._keep = []   # host arrays referenced by the descriptor during create
        self._handle = C.c_void_p()
        self._ws = None
        desc = self._describe(variables)
        _abi.check(self._lib.s5b200_model_create(C.byref(desc), C.byref(self._handle)), "s5b200_model_create")
        self._keep = []

    # ------------------------------------------------------------------ descriptor
    def _arr(self, a) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(a, dtype=F32))
        self._keep.append(a)
        return a

    def _describe(self, variables: dict) -> _abi.ModelDesc:
        cfg = self.cfg
        P, BS = variables["params"], variables["batch_stats"]
        q = cfg.recipe
        sq = cfg.static_quant
        enc, dec = P["encoder"]["encoder"], P["decoder"]

        def dense(p, w_bits, a_bits):
            k = np.asarray(p["kernel"], dtype=F32)
            b = np.asarray(p["bias"], dtype=F32)
            if sq:  # quantization.py:769, :788 (bias uses the activation scale)
                k = fake_quant(k, p["weight_scale"], w_bits)
                b = fake_quant(b, p["act_scale"], a_bits)
            return self._arr(k), self._arr(b)

        layers = (_abi.LayerDesc * cfg.n_layers)()
        for i in range(cfg.n_layers):
            lp = P["encoder"][f"layers_{i}"]
            ls = BS["encoder"][f"layers_{i}"]
            mx = lp["mixer"]
            Lb, Bb, Ct, D = ssm_discrete(mx, cfg)
            d = layers[i]
            if sq:
                g = lambda *k: float(F32(_get(mx, k)["scale"]))
                Lb = fake_quant(Lb.real, g("quant_A", "quant_real"), q["a_precision"]) + 1j * fake_quant(
                    Lb.imag, g("quant_A", "quant_imag"), q["a_precision"])
                Bb = fake_quant(Bb.real, g("quant_B", "quant_real"), q["b_precision"]) + 1j * fake_quant(
                    Bb.imag, g("quant_B", "quant_imag"), q["b_precision"])
                Ct = fake_quant(Ct.real, g("quant_C", "quant_real"), q["c_precision"]) + 1j * fake_quant(
                    Ct.imag, g("quant_C", "quant_imag"), q["c_precision"])
                D = fake_quant(D, g("quant_D"), q["d_precision"])
                d.s_ut, d.s_yt = g("quant_ut"), g("quant_yt")
                d.s_But_re, d.s_But_im = g("quant_But", "quant_real"), g("quant_But", "quant_imag")
                d.s_out2_act, d.s_out2_out = float(lp["out2"]["act_scale"]), float(lp["out2"]["out_scale"])
                d.s_gate_left = float(lp["mult_gate"]["quant_left"]["scale"])
                d.s_gate_right = float(lp["mult_gate"]["quant_right"]["scale"])
            d.lambda_bar_host = _fp(self._arr(_c2f(Lb)))
            d.B_bar_host = _fp(self._arr(_c2f(Bb)))
            d.C_tilde_host = _fp(self._arr(_c2f(Ct)))
            d.D_host = _fp(self._arr(D))
            norm_p = lp.get("norm", {})
            mean = np.asarray(ls["norm"]["mean"], dtype=F32)
            var = np.asarray(ls["norm"]["var"], dtype=F32)
            mul = (F32(1.0) / np.sqrt(var + F32(cfg.bn_eps))).astype(F32)   # flax BatchNorm, running stats
            if "scale" in norm_p:
                mul = mul * np.asarray(norm_p["scale"], dtype=F32)
            beta = np.asarray(norm_p["bias"], dtype=F32) if "bias" in norm_p else np.zeros_like(mean)
            d.bn_mean_host, d.bn_mul_host, d.bn_bias_host = _fp(self._arr(mean)), _fp(self._arr(mul)), _fp(self._arr(beta))
            ok, ob = dense(lp["out2"], q["non_ssm_precision"] if sq else 0, q["non_ssm_act_precision"] if sq else 0)
            d.out2_kernel_host, d.out2_bias_host = _fp(ok), _fp(ob)
        self._keep.append(layers)

        desc = _abi.ModelDesc()
        desc.flavour = _abi.STATIC if sq else _abi.FLOAT
        desc.n_layers, desc.H, desc.P = cfg.n_layers, cfg.d_model, cfg.ssm_size
        desc.d_input, desc.d_output = cfg.d_input, cfg.d_output
        desc.activation = _abi.ACT_RELU if cfg.relufication else _abi.ACT_GELU_TANH
        desc.relufication = int(cfg.relufication)
        desc.conj_sym = int(cfg.conj_sym)
        desc.ssm_act_bits = q["ssm_act_precision"] if sq else 0
        desc.non_ssm_act_bits = q["non_ssm_act_precision"] if sq else 0
        desc.use_sparse = int(cfg.use_sparse)
        ek, eb = dense(enc, q["non_ssm_precision"] if sq else 0, q["non_ssm_act_precision"] if sq else 0)
        dk, db = dense(dec, q["non_ssm_precision"] if sq else 0, q["non_ssm_act_precision"] if sq else 0)
        desc.enc_kernel_host, desc.enc_bias_host = _fp(ek), _fp(eb)
        desc.dec_kernel_host, desc.dec_bias_host = _fp(dk), _fp(db)
        if sq:
            desc.s_enc_act, desc.s_enc_out = float(enc["act_scale"]), float(enc["out_scale"])
            desc.s_dec_act, desc.s_dec_out = float(dec["act_scale"]), float(dec["out_scale"])
        desc.layers = layers
        return desc

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if self._handle:
            self._lib.s5b200_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream().cuda_stream)

    def _workspace(self, B: int, L: int):
        need = self._lib.s5b200_model_workspace_bytes(self._handle, B, L)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self._torch.empty(need, dtype=self._torch.uint8, device="cuda")
        return self._ws, need

    def _check_in(self, x, feat: int):
        t = self._torch
        if not (isinstance(x, t.Tensor) and x.is_cuda and x.dtype == t.float32 and x.dim() == 3 and x.shape[-1] == feat):
            raise ValueError(f"expected a float32 CUDA tensor of shape (B, L, {feat})")
        return x.contiguous()

    # ------------------------------------------------------------------ public calls
    def __call__(self, x):
        """QBatchRegressionModel.apply: (B, L, d_input) -> (B, L, d_output)."""
        x = self._check_in(x, self.cfg.d_input)
        B, L, _ = x.shape
        y = self._torch.empty((B, L, self.cfg.d_output), dtype=self._torch.float32, device=x.device)
        ws, need = self._workspace(B, L)
        _abi.check(self._lib.s5b200_model_forward(self._handle, x.data_ptr(), y.data_ptr(), B, L, ws.data_ptr(), need,
                                                  self._stream()), "s5b200_model_forward")
        return y

    def layer(self, layer: int, x, return_xs: bool = False):
        """QSequenceLayer.__call__ for block ``layer``: (B, L, H) -> (B, L, H) [, xs (B, L, P) complex64]."""
        x = self._check_in(x, self.cfg.d_model)
        B, L, _ = x.shape
        t = self._torch
        y = t.empty_like(x)
        xs = t.empty((B, L, self.cfg.ssm_size, 2), dtype=t.float32, device=x.device) if return_xs else None
        ws, need = self._workspace(B, L)
        _abi.check(self._lib.s5b200_layer_forward(self._handle, layer, x.data_ptr(), y.data_ptr(),
                                                  xs.data_ptr() if return_xs else None, B, L, ws.data_ptr(), need,
                                                  self._stream()), "s5b200_layer_forward")
        return (y, t.view_as_complex(xs)) if return_xs else y

    def apply_ssm(self, layer: int, u):
        """qS5SSM.__call__ for block ``layer`` on an already normalised input: returns (ys + D*u, xs)."""
        u = self._check_in(u, self.cfg.d_model)
        B, L, _ = u.shape
        t = self._torch
        y = t.empty_like(u)
        xs = t.empty((B, L, self.cfg.ssm_size, 2), dtype=t.float32, device=u.device)
        ws, need = self._workspace(B, L)
        _abi.check(self._lib.s5b200_ssm_forward(self._handle, layer, u.data_ptr(), y.data_ptr(), xs.data_ptr(), B, L,
                                                ws.data_ptr(), need, self._stream()), "s5b200_ssm_forward")
        return y, t.view_as_complex(xs)

    def forward_host(self, x_host, out=None, sync=True):
        """End-to-end call on HOST buffers (CPU tensors, ideally pinned): H2D, forward, D2H.

        ``sync=True`` returns when ``out`` is complete.  ``sync=False`` is stream-ordered like any CUDA call: it returns
        after enqueueing and ``out`` is complete after the current stream has been synchronised; keep ``x_host`` and
        ``out`` alive (and unmodified) until then.  Back-to-back calls then overlap across batches."""
        t = self._torch
        cfg = self.cfg
        if not (isinstance(x_host, t.Tensor) and not x_host.is_cuda and x_host.dtype == t.float32 and x_host.dim() == 3
                and x_host.shape[-1] == cfg.d_input):
            raise ValueError(f"expected a float32 CPU tensor of shape (B, L, {cfg.d_input})")
        if not x_host.is_contiguous():
            if not sync:
                # a temporary contiguous copy would be freed while the H2D copies are still queued
                raise ValueError("forward_host(sync=False) needs a contiguous input (it is read after the call returns)")
            x_host = x_host.contiguous()
        B, L, _ = x_host.shape
        if out is None:
            out = t.empty((B, L, cfg.d_output), dtype=t.float32, pin_memory=True)
        elif not (isinstance(out, t.Tensor) and not out.is_cuda and out.dtype == t.float32 and out.is_contiguous()
                  and tuple(out.shape) == (B, L, cfg.d_output)):
            raise ValueError(f"out must be a contiguous float32 CPU tensor of shape ({B}, {L}, {cfg.d_output})")
        if not sync and not (x_host.is_pinned() and out.is_pinned()):
            # cudaMemcpyAsync from / to pageable memory is staged synchronously by the driver: still correct, but the call
            # then blocks; say so instead of silently losing the overlap
            import warnings
            warnings.warn("forward_host(sync=False) with pageable host buffers: copies are not asynchronous; pin both "
                          "buffers (tensor.pin_memory()) for stream-ordered overlap", RuntimeWarning, stacklevel=2)
        fn = self._lib.s5b200_model_forward_host if sync else self._lib.s5b200_model_forward_host_async
        _abi.check(fn(self._handle, x_host.data_ptr(), out.data_ptr(), B, L, self._stream()), "s5b200_model_forward_host")
        return out


def _get(d, keys):
    for k in keys:
        d = d[k]
    return d


def scan_c64(Bu, lam, relu: bool = False, out=None):
    """x_k = lam * x_{k-1} + Bu_k over axis -2 (replaces lax.associative_scan / lax.scan, ssm.py:127, :150).

    Bu: complex64 CUDA tensor (B, L, P) or (L, P); lam: complex64 (P,).
    """
    import torch

    lib = _abi.lib()
    squeeze = Bu.dim() == 2
    if squeeze:
        Bu = Bu.unsqueeze(0)
    if not (Bu.is_cuda and Bu.dtype == torch.complex64 and Bu.dim() == 3):
        raise ValueError("Bu must be a complex64 CUDA tensor of shape (B, L, P)")
    Bu = Bu.contiguous()
    lam = lam.to(device=Bu.device, dtype=torch.complex64).contiguous()
    B, L, P = Bu.shape
    if out is None:
        out = torch.empty_like(Bu)
    need = lib.s5b200_scan_c64_workspace_bytes(B, L, P)
    if need == 0:
        raise _abi.S5Error(f"scan_c64: unsupported shape B={B} L={L} P={P}")
    ws = torch.empty(need, dtype=torch.uint8, device=Bu.device)
    _abi.check(lib.s5b200_scan_c64(Bu.data_ptr(), lam.data_ptr(), out.data_ptr(), B, L, P, int(relu), ws.data_ptr(),
                                   need, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s5b200_scan_c64")
    return out[0] if squeeze else out
