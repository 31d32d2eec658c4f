"""Oracle (TEST INFRASTRUCTURE ONLY): NDNS pre/post-processing either side of the model forward.

Restates, in NumPy float64 unless told otherwise,

* ``stft_splitter``  - sparseRNNs/train_helpers.py:1381-1396  (``jax.scipy.signal.stft``: boxcar window, nperseg = nfft
                       = 512, noverlap = 384, one-sided, defaults ``boundary='zeros'``, ``padded=True``,
                       ``scaling='spectrum'``, i.e. 256 zeros either side, tail zero-padded to a whole number of
                       hops, spectrum divided by ``sum(window) = 512``); returns magnitude and angle, (B, 257, L)
* ``stft_mixer``     - sparseRNNs/train_helpers.py:1399-1412  (``jax.scipy.signal.istft``: un-normalised inverse
                       real DFT of every frame (``irfft * sum(window)``), overlap-add, division by the overlap-added
                       squared window, 256 samples trimmed either side)
* ``si_snr``         - sparseRNNs/train_helpers.py:16-53
* ``ndns_eval_tail`` - sparseRNNs/train_helpers.py:1570-1573, 1605-1614: model input ``mag - 0.0007`` transposed to
                       (B, L, 257); ``cleaned_mag = mag * (1 + mask)``; ``cleaned = stft_mixer(cleaned_mag, phase)``;
                       ``loss = 0.001 * mean((cleaned_mag - clean_mag)^2) + (100 - si_snr)``

``jax.scipy.signal`` is third-party code that is absent from /root/reference; it documents itself as the JAX
implementation of ``scipy.signal.stft`` / ``istft``.  PIN: ``tests/test_oracle_ndns.py`` checks this restatement
against SciPy's own ``scipy.signal.stft`` / ``istft`` (installed here) with the reference's exact keyword arguments.
"""

from __future__ import annotations

import numpy as np

NFFT = 512
HOP = 128
STFT_MAG_MEAN = 0.0007   # train_helpers.py:1570
LAMBDA_MSE = 0.001       # train_helpers.py:1609
EPS = 1e-8               # train_helpers.py:37


def n_frames(n_samples: int) -> int:
    """Frames of the reference STFT for ``n_samples`` audio samples (480 000 -> 3751)."""
    ext = n_samples + NFFT                                   # boundary='zeros': nperseg // 2 either side
    pad = (-(ext - NFFT)) % HOP                              # padded=True
    return (ext + pad - NFFT) // HOP + 1


def stft_complex(audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """(B, T) -> complex (B, 257, L), scaled by 1 / 512 (train_helpers.py:1381-1393)."""
    audio = np.asarray(audio, dtype)
    B, T = audio.shape
    L = n_frames(T)
    ext = np.zeros((B, (L - 1) * HOP + NFFT), dtype)
    ext[:, NFFT // 2:NFFT // 2 + T] = audio
    idx = np.arange(L)[:, None] * HOP + np.arange(NFFT)[None, :]
    spec = np.fft.rfft(ext[:, idx], axis=-1) / NFFT          # (B, L, 257)
    return np.swapaxes(spec, 1, 2)


def stft_splitter(audio: np.ndarray, dtype=np.float64):
    """train_helpers.py:1381-1396: returns (mag, phase), each (B, 257, L)."""
    spec = stft_complex(audio, dtype)
    return np.abs(spec), np.angle(spec)


def stft_mixer(mag: np.ndarray, phase: np.ndarray, dtype=np.float64) -> np.ndarray:
    """train_helpers.py:1399-1412: (B, 257, L) magnitude and angle -> audio (B, (L - 1) * 128)."""
    mag = np.asarray(mag, dtype)
    phase = np.asarray(phase, dtype)
    spec = mag * np.exp(1j * phase)
    B, F, L = spec.shape
    assert F == NFFT // 2 + 1
    frames = np.fft.irfft(np.swapaxes(spec, 1, 2), n=NFFT, axis=-1) * NFFT   # (B, L, 512): irfft * sum(window)
    out_len = (L - 1) * HOP + NFFT
    x = np.zeros((B, out_len), dtype)
    norm = np.zeros(out_len, dtype)
    for k in range(L):                                       # overlap-add in frame order
        x[:, k * HOP:k * HOP + NFFT] += frames[:, k]
        norm[k * HOP:k * HOP + NFFT] += 1.0                  # boxcar: window^2 = 1
    x = np.where(norm > 1e-10, x / np.where(norm > 1e-10, norm, 1.0), x)
    return x[:, NFFT // 2:out_len - NFFT // 2]


def si_snr(target: np.ndarray, estimate: np.ndarray, dtype=np.float64) -> np.ndarray:
    """train_helpers.py:16-53, last axis = time; returns dB per clip."""
    target = np.asarray(target, dtype)
    estimate = np.asarray(estimate, dtype)
    s_t = target - target.mean(-1, keepdims=True)
    s_e = estimate - estimate.mean(-1, keepdims=True)
    dot = np.sum(s_t * s_e, -1, keepdims=True)
    proj = dot * s_t / np.sum(s_t ** 2, -1, keepdims=True)
    noise = s_e - proj
    sdr = np.sum(proj ** 2, -1) / (np.sum(noise ** 2, -1) + EPS)
    return 10 * np.log10(sdr + EPS)


def model_input(noisy_mag: np.ndarray) -> np.ndarray:
    """train_helpers.py:1570-1573: (B, 257, L) magnitude -> model input (B, L, 257) float32."""
    return np.swapaxes(np.asarray(noisy_mag) - STFT_MAG_MEAN, 1, 2).astype(np.float32)


def ndns_eval_tail(mask_blf: np.ndarray, noisy_mag: np.ndarray, noisy_phase: np.ndarray, clean_audio: np.ndarray,
                   clean_mag: np.ndarray, dtype=np.float64):
    """train_helpers.py:1605-1614 with the model's mask (B, L, 257): returns (losses, si_snr, cleaned audio)."""
    mask = np.swapaxes(np.asarray(mask_blf, dtype), 1, 2)
    cleaned_mag = np.asarray(noisy_mag, dtype) * (1.0 + mask)
    cleaned = stft_mixer(cleaned_mag, noisy_phase, dtype)
    score = si_snr(clean_audio, cleaned, dtype)
    losses = LAMBDA_MSE * np.mean((cleaned_mag - np.asarray(clean_mag, dtype)) ** 2, axis=(1, 2)) + (100 - score)
    return losses, score, cleaned
