"""Recipe-shaped synthetic NDNS model and input (no dataset, no checkpoint).

Host-side, NumPy only.  Produces the Flax-layout variable tree the reference's
``QBatchRegressionModel`` owns, so that the same tree can be fed to the
CUDA path, to the oracle and (on a box with JAX) to the reference itself.

Follows, for shapes and distributions only:
  * ``recipes/ndns.json`` + ``main.py:480-485`` (dim_scale arithmetic)
  * ``sparseRNNs/train.py:40-111`` (block-diagonal HiPPO-LegS init, conj_sym halving)
  * ``sparseRNNs/model/ssm_init.py:8-75, 103-175`` (make_DPLR_HiPPO, init_VinvB, init_CV, log-step init)
  * ``sparseRNNs/model/ssm.py:436-512`` (parameter names / shapes)
  * ``sparseRNNs/utils/pruning.py:7-19`` (magnitude pruning, ERK distribution, target 0.9)
JAX's PRNG cannot be reproduced without JAX; a NumPy Generator with a fixed seed stands in.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEQ_LENGTH = 3751   # dataloaders/dataloading.py:132-135
IN_DIM = 257
N_CLASSES = 257
STFT_MAG_MEAN = 0.0007  # train_helpers.py:1570


@dataclass(frozen=True)
class NdnsDims:
    n_layers: int
    d_model: int      # H
    ssm_size: int     # P (complex states after conj_sym halving)
    blocks: int
    d_input: int = IN_DIM
    d_output: int = N_CLASSES


def ndns_dims(dim_scale: float = 0.5) -> NdnsDims:
    """recipes/ndns.json through main.py:480-485 and train.py:99-101."""
    n_layers, d_model, ssm_size_base, blocks = 3, 192, 256, 16
    d_model_block = d_model / blocks
    ssm_size_block = ssm_size_base / blocks
    blocks = int(blocks * dim_scale)
    d_model = int(blocks * d_model_block)
    ssm_size_base = int(blocks * ssm_size_block)
    return NdnsDims(n_layers=n_layers, d_model=d_model, ssm_size=ssm_size_base // 2, blocks=blocks)


# --------------------------------------------------------------------------- init
def _make_dplr_hippo(n: int):
    """HiPPO-LegS normal-plus-low-rank diagonalisation (ssm_init.py:8-75)."""
    p = np.sqrt(1 + 2 * np.arange(n))
    a = p[:, None] * p[None, :]
    a = -(np.tril(a) - np.diag(np.arange(n)))
    pl = np.sqrt(np.arange(n) + 0.5)
    s = a + pl[:, None] * pl[None, :]
    lam_real = np.mean(np.diagonal(s)) * np.ones(n)
    lam_imag, v = np.linalg.eigh(s * -1j)
    return lam_real + 1j * lam_imag, v


def _lecun_normal(rng: np.random.Generator, shape, fan_in: int):
    """Truncated normal (+-2 sigma) with variance 1/fan_in (jax.nn.initializers.lecun_normal)."""
    std = np.sqrt(1.0 / fan_in) / 0.87962566103423978
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * std


def init_variables(dims: NdnsDims, seed: int = 1919, *, bn_affine: bool = True,
                   dt_min: float = 1e-3, dt_max: float = 1e-1) -> dict:
    """Random-init variable tree in recipe shapes (float32 leaves).

    Dense biases and BatchNorm affine parameters get small non-zero values ("trained-like"): the
    reference's fixed-point configuration takes log2 of their absmax (fxputils.py:263-265), which
    is undefined for the all-zero initial values.  BatchNorm running statistics are filled in by
    ``set_batchnorm_stats``.
    """
    rng = np.random.default_rng(seed)
    H, P, nb = dims.d_model, dims.ssm_size, dims.blocks
    block = 2 * P // nb                    # train.py:46
    lam, v = _make_dplr_hippo(block)       # train.py:97
    half = block // 2                      # train.py:99-101 (conj_sym)
    lam = lam[:half]
    v = v[:, :half]
    vc = v.conj().T
    lam_full = np.tile(lam, nb)            # train.py:109
    V = np.zeros((block * nb, half * nb), dtype=np.complex128)     # block_diag, train.py:110-111
    Vinv = np.zeros((half * nb, block * nb), dtype=np.complex128)
    for b in range(nb):
        V[b * block:(b + 1) * block, b * half:(b + 1) * half] = v
        Vinv[b * half:(b + 1) * half, b * block:(b + 1) * block] = vc

    def dense(n_in, n_out, bias_std=0.02):
        return {"kernel": _lecun_normal(rng, (n_in, n_out), n_in).astype(np.float32),
                "bias": (bias_std * rng.standard_normal(n_out)).astype(np.float32)}

    # The encoder bias is kept inside the range the reference's static path can represent: it
    # quantises the bias with the *input* scale (quantization.py:788), and STFT magnitudes are ~1e-2.
    params = {"encoder": {"encoder": dense(dims.d_input, H, bias_std=0.001)}, "decoder": None}
    stats = {"encoder": {}}
    for i in range(dims.n_layers):
        Bm = _lecun_normal(rng, (2 * P, H), 2 * P)                 # ssm.py:448-454, init_VinvB
        VinvB = Vinv @ Bm
        Cm = _lecun_normal(rng, (H, 2 * P, 2), 2 * P * H)          # ssm.py:461-463, init_CV
        CV = (Cm[..., 0] + 1j * Cm[..., 1]) @ V
        log_step = rng.uniform(size=(P, 1)) * (np.log(dt_max) - np.log(dt_min)) + np.log(dt_min)
        mixer = {
            "Lambda_re": lam_full.real.astype(np.float32),
            "Lambda_im": lam_full.imag.astype(np.float32),
            "B": np.stack([VinvB.real, VinvB.imag], axis=-1).astype(np.float32),
            "C": np.stack([CV.real, CV.imag], axis=-1).astype(np.float32),
            "D": rng.standard_normal(H).astype(np.float32),
            "log_step": log_step.astype(np.float32),
        }
        layer = {"mixer": mixer, "out2": dense(H, H)}
        if bn_affine:
            layer["norm"] = {"scale": rng.uniform(0.8, 1.2, H).astype(np.float32),
                             "bias": (0.1 * rng.standard_normal(H)).astype(np.float32)}
        else:
            layer["norm"] = {}
        params["encoder"][f"layers_{i}"] = layer
        stats["encoder"][f"layers_{i}"] = {"norm": {"mean": np.zeros(H, np.float32),
                                                    "var": np.ones(H, np.float32)}}
    params["decoder"] = dense(H, dims.d_output)
    return {"params": params, "batch_stats": stats}


# --------------------------------------------------------------------------- pruning
def erk_sparsities(shapes: dict, target: float) -> dict:
    """Erdos-Renyi-Kernel per-tensor sparsities averaging ``target`` (pruning.py:17, dist_type 'erk').

    Restates the published RigL / jaxpruner ERK rule: density_t proportional to sum(shape)/prod(shape),
    tensors whose density would exceed 1 are kept dense and the rest re-balanced.
    """
    dense = set()
    while True:
        divisor, rhs, raw = 0.0, 0.0, {}
        for name, shape in shapes.items():
            n = float(np.prod(shape))
            n_zeros = n * target
            if name in dense:
                rhs -= n_zeros
            else:
                rhs += n - n_zeros
                raw[name] = float(np.sum(shape)) / n
                divisor += raw[name] * n
        eps = rhs / divisor
        mx = max(raw.values())
        if mx * eps > 1:
            dense |= {k for k, p in raw.items() if p == mx}
        else:
            break
    return {k: (0.0 if k in dense else 1.0 - eps * raw[k]) for k in shapes}


def _prunable(params: dict, prefix=()):
    for k, v in params.items():
        if isinstance(v, dict):
            yield from _prunable(v, prefix + (k,))
        elif np.ndim(v) > 1:       # jaxpruner default filter: every parameter with ndim > 1
            yield prefix + (k,), v


def prune_magnitude(variables: dict, target: float = 0.9) -> dict:
    """Per-tensor magnitude top-k masks at ERK sparsities; returns a new tree with zeros applied."""
    import copy

    out = copy.deepcopy(variables)
    tensors = dict(_prunable(out["params"]))
    spars = erk_sparsities({k: v.shape for k, v in tensors.items()}, target)
    for path, w in tensors.items():
        s = spars[path]
        n_keep = int(round(w.size * (1.0 - s)))
        if n_keep >= w.size:
            continue
        flat = np.abs(w).ravel()
        thresh_idx = np.argpartition(flat, w.size - n_keep)[w.size - n_keep:]
        mask = np.zeros(w.size, dtype=bool)
        mask[thresh_idx] = True
        node = out["params"]
        for k in path[:-1]:
            node = node[k]
        node[path[-1]] = (w * mask.reshape(w.shape)).astype(np.float32)
    return out


# --------------------------------------------------------------------------- inputs
def synthetic_stft_batch(n_clips: int, n_frames: int = SEQ_LENGTH, seed: int = 0,
                         audio_std: float = 0.05, exact_stft: bool = False) -> np.ndarray:
    """Model input ``(n_clips, n_frames, 257)`` = |STFT(noise)| - 0.0007 (train_helpers.py:1381-1396, 1570-1573).

    ``exact_stft=True`` frames Gaussian audio with the reference's boxcar window (nfft 512, hop 128,
    zero boundary padding, 1/nfft scaling).  The default draws the bin magnitudes directly from the
    distribution white noise has under that transform (Rayleigh; DC and Nyquist bins half-normal),
    which is indistinguishable for benchmarking and far cheaper for thousands of clips.
    """
    rng = np.random.default_rng(seed)
    nfft, hop = 512, 128
    if exact_stft:
        n_samples = (n_frames - 1) * hop
        audio = (audio_std * rng.standard_normal((n_clips, n_samples))).astype(np.float32)
        pad = nfft // 2
        audio = np.pad(audio, ((0, 0), (pad, pad)))
        idx = np.arange(n_frames)[:, None] * hop + np.arange(nfft)[None, :]
        spec = np.fft.rfft(audio[:, idx], axis=-1) / nfft
        mag = np.abs(spec)
    else:
        sigma = audio_std * np.sqrt(nfft / 2.0) / nfft
        re = rng.standard_normal((n_clips, n_frames, nfft // 2 + 1), dtype=np.float32)
        im = rng.standard_normal((n_clips, n_frames, nfft // 2 + 1), dtype=np.float32)
        im[..., 0] = 0
        im[..., -1] = 0
        re[..., 0] *= np.sqrt(2.0)
        re[..., -1] *= np.sqrt(2.0)
        mag = sigma * np.sqrt(re * re + im * im)
    return (mag - STFT_MAG_MEAN).astype(np.float32)


def set_batchnorm_stats(variables: dict, per_layer_inputs: list) -> None:
    """Fill BatchNorm running mean/var from the given per-layer block inputs ((..., H) arrays)."""
    for i, h in enumerate(per_layer_inputs):
        flat = np.asarray(h, dtype=np.float64).reshape(-1, h.shape[-1])
        node = variables["batch_stats"]["encoder"][f"layers_{i}"]["norm"]
        node["mean"] = flat.mean(0).astype(np.float32)
        node["var"] = flat.var(0).astype(np.float32)
