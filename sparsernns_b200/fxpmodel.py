"""Host-side mirror of the reference's fixed-point model (sparseRNNs/fxpmodel.py, fxprun.py).

``FxpRegressionModel(modeldict, fxp_qconfig, n_layers)`` takes exactly what the reference's constructor
takes (fxprun.py:437-457): the merged parameter/statistics dictionary produced by ``load_modeldict`` and
the bits/exponent configuration produced by ``create_fxp_qconfig`` + ``add_target_bits_exp``.  The
``setup()`` work of the reference modules - ZOH discretisation in float, ``fxp_from_fp(ROUND)`` of every
parameter at its configured exponent (fxpmodel.py:313-329, 568-608, 853-888), the sigmoid LUT
(fxpmodel.py:89-95) - is done here once with NumPy; the forward runs entirely in libs5b200.

Only the configuration reachable in the reference is supported: prenorm BatchNorm (not fused),
relufication, glu_variant "half1", conj_sym, ZOH, use_lax_scan.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .model import S5Config, ssm_discrete

F32 = np.float32
I32 = np.int32

TAP_LAYER_INPUT, TAP_PRE_S5, TAP_XS, TAP_POST_GLU, TAP_OUTPUT = range(5)


def from_fp_round(x, bits: int, exp: int) -> np.ndarray:
    """fxp_from_fp(..., round_mode=ROUND) on the host (fxparray.py:287-307): x*2^exp, half-to-even, saturate."""
    v = np.round(np.asarray(x, dtype=F32) * F32(1 << exp)).astype(np.float64)
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    return np.ascontiguousarray(np.clip(v, lo, hi).astype(np.int64).astype(I32))


def sigmoid_lut(x_exp: int, y_exp: int, x_extra: int = 3, n_exp: int = 3) -> np.ndarray:
    """FxpSigmoid.__init__ (fxpmodel.py:89-95), float32 like the reference."""
    x = np.linspace(0, 1 << (x_exp + x_extra), (1 << n_exp) + 1, dtype=F32)[:-1].astype(np.int64)
    xf = (x.astype(F32) / F32(1 << x_exp)).astype(F32)
    sig = (F32(1.0) / (F32(1.0) + np.exp(-xf))).astype(F32)
    return (np.round(sig * F32(1 << y_exp)) - F32(1 << (y_exp - 1))).astype(np.int64).astype(I32)


@dataclass
class FxpTensor:
    """Device FxpArray: int32 payload (torch CUDA tensor) + (bits, exp), signed."""
    data: object
    bits: int
    exp: int

    def to_float(self):
        return self.data.float() / float(1 << self.exp)  # FxpArray.to_float, fxparray.py:72-73


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_abi.c_int32_p)


class FxpRegressionModel:
    """fxpmodel.py:1380-1439 bound to one weight set; ``__call__`` is ``model(fxp_x)`` of fxprun.py:76."""

    def __init__(self, modeldict: dict, fxp_qconfig: dict, n_layers: int, step_rescale: float = 1.0,
                 use_sparse: bool = True, global_batch: bool = False, group=None):
        import torch

        if not torch.cuda.is_available():
            raise _abi.S5Error("sparsernns_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._torch = torch
        self._lib = _abi.lib()
        self.q = fxp_qconfig
        self.n_layers = n_layers
        enc = modeldict["encoder"]
        self.d_input, self.H = np.asarray(enc["encoder"]["kernel"]).shape
        self.d_output = np.asarray(modeldict["decoder"]["kernel"]).shape[-1]
        self.P = np.asarray(enc["layers_0"]["mixer"]["Lambda_re"]).shape[0]
        self._keep = []
        self._handle = C.c_void_p()
        self._ws = None
        # the fxp model never clips eigenvalues (fxpmodel.py:434, :440-441)
        self._cfg = S5Config(n_layers=n_layers, d_model=self.H, ssm_size=self.P, d_input=self.d_input,
                             d_output=self.d_output, relufication=True, clip_eigs=False, step_rescale=step_rescale)
        desc = self._describe(modeldict, use_sparse)
        _abi.check(self._lib.s5b200_fxp_model_create(C.byref(desc), C.byref(self._handle)), "s5b200_fxp_model_create")
        self._keep = []
        self.out_bits = fxp_qconfig["decoder"]["out_bits"]
        self.out_exp = fxp_qconfig["decoder"]["out_exp"]
        self.inp_bits = fxp_qconfig["encoder"]["inp_bits"]
        self.inp_exp = fxp_qconfig["encoder"]["inp_exp"]
        self._sync_cb = None
        if global_batch:
            self._install_global_batch_sync(group)

    def _install_global_batch_sync(self, group):
        """Sharded clips, result bit-exact with ONE process on the union batch (SURVEY section 8e option B): the running float
        maxima behind the ``compute_best`` exponents are max-all-reduced (NCCL) after every kernel that produces them."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise _abi.S5Error("global_batch=True needs an initialised torch.distributed process group")
        t = self._torch

        def cb(_user, ptr, count, _stream):
            ws = self._ws                                    # the maxima live in the caller-provided workspace
            off = int(ptr) - ws.data_ptr()
            view = ws[off:off + 4 * count].view(t.int32)      # bit patterns of non-negative floats: integer max == float max
            dist.all_reduce(view, op=dist.ReduceOp.MAX, group=group)

        self._sync_cb = _abi.FXP_SYNC_FN(cb)                  # keep the thunk alive as long as the handle
        _abi.check(self._lib.s5b200_fxp_model_set_maxima_sync(self._handle, C.cast(self._sync_cb, C.c_void_p), None),
                   "s5b200_fxp_model_set_maxima_sync")

    # ------------------------------------------------------------------ descriptor
    def _k(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=I32)
        self._keep.append(a)
        return a

    def _dense(self, md: dict, q: dict, d: _abi.FxpDenseDesc):
        """FxpDense.setup (fxpmodel.py:298-329)."""
        d.weight_host = _ip(self._k(from_fp_round(md["kernel"], q["w_bits"], q["w_exp"])))
        d.bias_host = _ip(self._k(from_fp_round(md["bias"], q["b_bits"], q["b_exp"])))
        d.w_exp, d.b_exp, d.b_bits = q["w_exp"], q["b_exp"], q["b_bits"]
        d.inp_bits, d.inp_exp, d.out_bits, d.out_exp = q["inp_bits"], q["inp_exp"], q["out_bits"], q["out_exp"]
        # FxpDense.export (fxpmodel.py:368-393)
        return dict(
            params=dict(weight=self._keep[-2].copy(), bias=self._keep[-1].copy()),
            qconfig=dict(weight_exp=q["w_exp"], weight_bits=q["w_bits"], weight_signed=True, bias_exp=q["b_exp"],
                         bias_bits=q["b_bits"], bias_signed=True, inp_bits=q["inp_bits"], inp_exp=q["inp_exp"],
                         out_bits=q["out_bits"], out_exp=q["out_exp"]))

    def _describe(self, md: dict, use_sparse: bool) -> _abi.FxpModelDesc:
        q = self.q
        layers = (_abi.FxpLayerDesc * self.n_layers)()
        exp_params = {"encoder": {}, "decoder": None}
        exp_qcfg = {"encoder": {}, "decoder": None}
        for i in range(self.n_layers):
            lq = q["blocks"]
            if any(k.startswith("layers_") for k in lq):  # per-layer exponents (fxpmodel.py:988-993)
                lq = lq[f"layers_{i}"]
            lm = md["encoder"][f"layers_{i}"]
            d = layers[i]
            nq, nm = lq["norm"], lm["norm"]
            # FxpBatchNorm.setup (fxpmodel.py:852-888)
            d.bn_minus_mean_host = _ip(self._k(from_fp_round(F32(-1) * np.asarray(nm["mean"], F32), nq["mean"]["bits"], nq["mean"]["exp"])))
            d.bn_mean_exp, d.bn_mean_bits = nq["mean"]["exp"], nq["mean"]["bits"]
            invsq = F32(1.0) / np.sqrt(np.asarray(nm["var"], F32) + F32(1e-5))
            d.bn_invsq_var_host = _ip(self._k(from_fp_round(invsq, nq["invsq_var"]["bits"], nq["invsq_var"]["exp"])))
            d.bn_invsq_exp, d.bn_invsq_bits = nq["invsq_var"]["exp"], nq["invsq_var"]["bits"]
            if "scale" in nm:
                d.bn_scale_host = _ip(self._k(from_fp_round(nm["scale"], nq["scale"]["bits"], nq["scale"]["exp"])))
                d.bn_scale_exp, d.bn_scale_bits = nq["scale"]["exp"], nq["scale"]["bits"]
            if "bias" in nm:
                d.bn_bias_host = _ip(self._k(from_fp_round(nm["bias"], nq["bias"]["bits"], nq["bias"]["exp"])))
                d.bn_bias_exp, d.bn_bias_bits = nq["bias"]["exp"], nq["bias"]["bits"]
            # FxpSSM.setup (fxpmodel.py:436-451, 562-608)
            mixer = lm["seq"] if "seq" in lm else lm["mixer"]
            Lb, Bb, Ct, D = ssm_discrete(mixer, self._cfg)
            w, a = lq["ssm"]["weights"], lq["ssm"]["activations"]
            ff = lambda v, k: self._k(from_fp_round(v, w[k]["bits"], w[k]["exp"]))
            d.A_re_host, d.A_im_host = _ip(ff(Lb.real, "A_re")), _ip(ff(Lb.imag, "A_im"))
            d.B_re_host, d.B_im_host = _ip(ff(Bb.real, "B_re")), _ip(ff(Bb.imag, "B_im"))
            d.C_re_host, d.C_im_host = _ip(ff(Ct.real, "C_re")), _ip(ff(Ct.imag, "C_im"))
            d.D_host = _ip(ff(D, "D"))
            d.eA_re, d.eA_im, d.eB_re, d.eB_im = w["A_re"]["exp"], w["A_im"]["exp"], w["B_re"]["exp"], w["B_im"]["exp"]
            d.eC_re, d.eC_im, d.eD = w["C_re"]["exp"], w["C_im"]["exp"], w["D"]["exp"]
            d.u_bits, d.u_exp = a["u"]["bits"], a["u"]["exp"]
            d.Bu_re_bits, d.Bu_re_exp = a["Bu_re"]["bits"], a["Bu_re"]["exp"]
            d.Bu_im_bits, d.Bu_im_exp = a["Bu_im"]["bits"], a["Bu_im"]["exp"]
            d.x_re_exp, d.x_im_exp = a["x_re"]["exp"], a["x_im"]["exp"]
            d.y_bits, d.y_exp = a["y"]["bits"], a["y"]["exp"]
            # FxpSSM.export (fxpmodel.py:819-847) / FxpBatchNorm.export (:946-967): integer payloads and their formats
            g = lambda ptr_arr: ptr_arr.copy()
            names = ("A_real", "A_imag", "B_real", "B_imag", "C_real", "C_imag", "D")
            keys = ("A_re", "A_im", "B_re", "B_im", "C_re", "C_im", "D")
            vals = (Lb.real, Lb.imag, Bb.real, Bb.imag, Ct.real, Ct.imag, D)
            mp, mq = {}, {}
            for nme, k, v in zip(names, keys, vals):
                mp[nme] = from_fp_round(v, w[k]["bits"], w[k]["exp"])
                mq[f"{nme}_bits"], mq[f"{nme}_exp"], mq[f"{nme}_signed"] = w[k]["bits"], w[k]["exp"], True
            for k in ("u", "Bu_re", "Bu_im", "x_re", "x_im", "y"):
                mq[f"{k}_bits"], mq[f"{k}_exp"] = a[k]["bits"], a[k]["exp"]
            np_, nq_ = dict(mean=-from_fp_round(F32(-1) * np.asarray(nm["mean"], F32), nq["mean"]["bits"], nq["mean"]["exp"]),
                            invsq_var=from_fp_round(invsq, nq["invsq_var"]["bits"], nq["invsq_var"]["exp"])), \
                dict(mean_bits=nq["mean"]["bits"], mean_exp=nq["mean"]["exp"], invsq_var_bits=nq["invsq_var"]["bits"],
                     invsq_var_exp=nq["invsq_var"]["exp"])
            for k in ("bias", "scale"):
                if k in nm:
                    np_[k] = from_fp_round(nm[k], nq[k]["bits"], nq[k]["exp"])
                    nq_[f"{k}_bits"], nq_[f"{k}_exp"] = nq[k]["bits"], nq[k]["exp"]
            o2 = self._dense(lm["out2"], lq["out2"], d.out2)
            mgq = lq["multgate"]
            exp_params["encoder"][f"layers_{i}"] = dict(mixer=mp, out2=o2["params"], norm=np_)
            exp_qcfg["encoder"][f"layers_{i}"] = dict(
                mixer=mq, out2=o2["qconfig"], norm=nq_,
                multgate={k: mgq[k] for k in ("l_bits", "l_exp", "r_bits", "r_exp", "res_bits", "res_exp")},
                sigmoid=dict(x_exp=min(lq["out2"]["out_exp"], 6), y_exp=lq["out2"]["out_bits"] - 2, x_extra=3, n_exp=3))
            # sigmoid + mult_gate (fxpmodel.py:1075-1108)
            d.sig_x_exp = min(lq["out2"]["out_exp"], 6)
            d.sig_y_exp = lq["out2"]["out_bits"] - 2
            lut = sigmoid_lut(d.sig_x_exp, d.sig_y_exp)
            for j in range(8):
                d.sig_lut[j] = int(lut[j])
            mg = lq["multgate"]
            d.gate_l_bits, d.gate_l_exp, d.gate_r_bits, d.gate_r_exp = mg["l_bits"], mg["l_exp"], mg["r_bits"], mg["r_exp"]
            d.gate_res_bits, d.gate_res_exp = mg["res_bits"], mg["res_exp"]
        self._keep.append(layers)
        desc = _abi.FxpModelDesc()
        desc.n_layers, desc.H, desc.P = self.n_layers, self.H, self.P
        desc.d_input, desc.d_output, desc.use_sparse = self.d_input, self.d_output, int(use_sparse)
        e = self._dense(md["encoder"]["encoder"], q["encoder"], desc.encoder)
        dd = self._dense(md["decoder"], q["decoder"], desc.decoder)
        exp_params["encoder"]["encoder"], exp_qcfg["encoder"]["encoder"] = e["params"], e["qconfig"]
        exp_params["decoder"], exp_qcfg["decoder"] = dd["params"], dd["qconfig"]
        self._export = dict(params=exp_params, qconfig=exp_qcfg)
        self.modeldict = md
        desc.layers = layers
        return desc

    # ------------------------------------------------------------------ export
    def export(self, fxp_x=None) -> dict:
        """``FxpRegressionModel.export()`` (fxpmodel.py:1441-1458): ``dict(params, qconfig, intermediates)`` with the
        reference's nesting (encoder / layers_i / mixer | out2 | norm, decoder).  Integer payloads are NumPy int32 arrays.
        ``intermediates`` holds what the library can tap after a forward on ``fxp_x``: per block ``ssm_input``,
        ``post_GLU``, ``output``, ``mixer/xs_relu`` and ``mixer/u`` (the mixer's input AFTER its change_cfg to the ``u``
        format, fxpmodel.py:620-624 - the reference sows the value before that step as ``pre_s5`` and has no key for this
        one), plus ``encoder/pre_encoder`` and ``decoder/__call__``; every entry is
        ``dict(data, bits, exp, signed)`` (the reference pickles ``FxpArray`` objects, whose class lives in its own
        package; the fields are the same)."""
        out = dict(params=self._export["params"], qconfig=self._export["qconfig"], intermediates={})
        if fxp_x is not None:
            fa = lambda t: dict(data=t.data.cpu().numpy(), bits=t.bits, exp=t.exp, signed=True)
            enc = {"encoder": {"pre_encoder": fa(fxp_x)}}
            for i in range(self.n_layers):
                enc[f"layers_{i}"] = dict(ssm_input=fa(self.tap(fxp_x, i, TAP_LAYER_INPUT)),
                                          post_GLU=fa(self.tap(fxp_x, i, TAP_POST_GLU)), output=fa(self.tap(fxp_x, i, TAP_OUTPUT)),
                                          mixer=dict(xs_relu=fa(self.tap(fxp_x, i, TAP_XS)), u=fa(self.tap(fxp_x, i, TAP_PRE_S5))))
            out["intermediates"] = dict(encoder=enc, decoder={"__call__": fa(self(fxp_x))})
        return out

    def export_files(self, folder: str, fxp_x) -> None:
        """The five pickles ``fxprun.py --export`` writes (fxprun.py:476-495), same file names and top-level keys."""
        import os
        import pickle

        data = self.export(fxp_x)
        os.makedirs(folder, exist_ok=True)
        dump = lambda name, obj: pickle.dump(obj, open(os.path.join(folder, name), "wb"))
        dump("fxpmodel_modeldict.pkl", self.modeldict)
        dump("fxpmodel_fxp_qconfig.pkl", self.q)
        dump("fxpmodel_io.pkl", dict(input=data["intermediates"]["encoder"]["encoder"]["pre_encoder"],
                                     output=data["intermediates"]["decoder"]["__call__"]))
        dump("fxpmodel.pkl", {k: v for k, v in data.items() if k != "intermediates"})
        dump("fxpmodel_activations.pkl", {"intermediates": data["intermediates"]})

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if self._handle:
            self._lib.s5b200_fxp_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream().cuda_stream)

    def _workspace(self, B, L):
        need = self._lib.s5b200_fxp_model_workspace_bytes(self._handle, B, L)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self._torch.empty(need, dtype=self._torch.uint8, device="cuda")
        return self._ws, need

    def _run(self, x, is_float: bool, want_float: bool):
        t = self._torch
        want_dtype = t.float32 if is_float else t.int32
        if not (isinstance(x, t.Tensor) and x.is_cuda and x.dtype == want_dtype and x.dim() == 3 and x.shape[-1] == self.d_input):
            raise ValueError(f"expected a {want_dtype} CUDA tensor of shape (B, L, {self.d_input})")
        x = x.contiguous()
        B, L, _ = x.shape
        y = t.empty((B, L, self.d_output), dtype=t.int32, device=x.device)
        yf = t.empty((B, L, self.d_output), dtype=t.float32, device=x.device) if want_float else None
        status = t.zeros(4, dtype=t.int32, device=x.device)
        ws, need = self._workspace(B, L)
        fn = self._lib.s5b200_fxp_model_forward if is_float else self._lib.s5b200_fxp_model_forward_int
        _abi.check(fn(self._handle, x.data_ptr(), y.data_ptr(), yf.data_ptr() if want_float else None, B, L,
                      ws.data_ptr(), need, status.data_ptr(), self._stream()), "s5b200_fxp_model_forward")
        return y, yf, status

    # ------------------------------------------------------------------ public calls
    def quantize_input(self, x):
        """fxprun.py:69-75: fxp_from_fp(x, enc inp bits/exp, FLOOR) on the device."""
        t = self._torch
        x = x.contiguous()
        y = t.empty(x.shape, dtype=t.int32, device=x.device)
        _abi.check(self._lib.s5b200_fxp_from_fp(x.data_ptr(), y.data_ptr(), x.numel(), self.inp_bits, self.inp_exp,
                                                _abi.ROUND_FLOOR, self._stream()), "s5b200_fxp_from_fp")
        return FxpTensor(y, self.inp_bits, self.inp_exp)

    def __call__(self, fxp_x, check: bool = True) -> FxpTensor:
        """model(fxp_x) -> FxpTensor (decoder output)."""
        x = fxp_x.data if isinstance(fxp_x, FxpTensor) else fxp_x
        y, _, status = self._run(x, False, False)
        if check and int(status[0].item()) != 0:
            raise _abi.S5Error("fixed-point forward hit an op the reference raises on (negative shift in fxp_mul)")
        return FxpTensor(y, self.out_bits, self.out_exp)

    def forward_float(self, x, check: bool = False):
        """run_validation's hot part on a float input: returns (y_int FxpTensor, mask float32) (fxprun.py:69-77)."""
        y, yf, status = self._run(x, True, True)
        if check and int(status[0].item()) != 0:
            raise _abi.S5Error("fixed-point forward hit an op the reference raises on (negative shift in fxp_mul)")
        return FxpTensor(y, self.out_bits, self.out_exp), yf

    def tap(self, fxp_x, layer: int, tap: int) -> FxpTensor:
        """Intermediate of block ``layer`` (names of the reference's ``sow`` keys, see s5b200.h)."""
        t = self._torch
        x = (fxp_x.data if isinstance(fxp_x, FxpTensor) else fxp_x).contiguous()
        B, L, _ = x.shape
        shape = (B, L, self.P, 2) if tap == TAP_XS else (B, L, self.H)
        dst = t.empty(shape, dtype=t.int32, device=x.device)
        e = C.c_int(0)
        ws, need = self._workspace(B, L)
        _abi.check(self._lib.s5b200_fxp_model_tap(self._handle, x.data_ptr(), layer, tap, dst.data_ptr(), C.byref(e), B, L,
                                                  ws.data_ptr(), need, self._stream()), "s5b200_fxp_model_tap")
        return FxpTensor(dst, 16, int(e.value))
