"""NDNS pre/post-processing on the GPU, mirroring the reference's callables (sparseRNNs/train_helpers.py):

* ``stft_splitter(audio) -> (mag, phase)``              train_helpers.py:1381-1396
* ``stft_mixer(mag, phase) -> audio``                    train_helpers.py:1399-1412
* ``si_snr(target, estimate) -> dB``                     train_helpers.py:16-53 (``si_snr_jax``)
* ``denoise_eval(model, noisy, clean) -> (losses, si_snr, cleaned)``
                                                         the forward of ``ndns_eval_step`` / ``train_step_ndns``,
                                                         train_helpers.py:1570-1573, 1605-1614

All tensors are float32 CUDA tensors; spectra use the reference's (B, 257, L) layout unless ``blf=True``.
The kernels live in csrc/ndns.cu behind the C ABI (include/s5b200.h section 5); there is no CPU path.
"""

from __future__ import annotations

import ctypes as C

from . import _abi

NFFT, HOP, BINS = 512, 128, 257
STFT_MAG_MEAN = 0.0007   # train_helpers.py:1570
LAMBDA_MSE = 0.001       # train_helpers.py:1609


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_audio(audio):
    import torch

    if not (audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2):
        raise ValueError("audio must be a float32 CUDA tensor of shape (B, T)")
    return audio.contiguous()


def stft_splitter(audio, *, blf: bool = False, model_input: bool = False):
    """(B, T) -> (mag, phase) [, model_in]; ``model_input=True`` also returns ``mag - 0.0007`` as (B, L, 257)."""
    import torch

    lib = _abi.lib()
    audio = _check_audio(audio)
    B, T = audio.shape
    L = lib.s5b200_stft_frames(T)
    shape = (B, L, BINS) if blf else (B, BINS, L)
    mag = torch.empty(shape, dtype=torch.float32, device=audio.device)
    phase = torch.empty(shape, dtype=torch.float32, device=audio.device)
    xin = torch.empty((B, L, BINS), dtype=torch.float32, device=audio.device) if model_input else None
    _abi.check(lib.s5b200_stft(audio.data_ptr(), B, T, mag.data_ptr(), phase.data_ptr(), int(blf),
                               xin.data_ptr() if model_input else None, STFT_MAG_MEAN, _stream()), "s5b200_stft")
    return (mag, phase, xin) if model_input else (mag, phase)


def stft_mixer(mag, phase, *, blf: bool = False, mask=None, return_cleaned_mag: bool = False):
    """(mag, phase) -> audio (B, (L - 1) * 128); ``mask`` (B, L, 257) applies ``mag * (1 + mask)`` first."""
    import torch

    lib = _abi.lib()
    if not (mag.is_cuda and mag.dtype == torch.float32 and mag.dim() == 3 and phase.shape == mag.shape and phase.dtype == torch.float32):
        raise ValueError("mag / phase must be float32 CUDA tensors of the same 3-d shape")
    mag, phase = mag.contiguous(), phase.contiguous()
    B = mag.shape[0]
    L = mag.shape[1] if blf else mag.shape[2]
    if (mag.shape[2] if blf else mag.shape[1]) != BINS:
        raise ValueError(f"expected {BINS} frequency bins")
    if mask is not None:
        if tuple(mask.shape) != (B, L, BINS) or mask.dtype != torch.float32:
            raise ValueError("mask must be the model output (B, L, 257) float32")
        mask = mask.contiguous()
    audio = torch.empty((B, (L - 1) * HOP), dtype=torch.float32, device=mag.device)
    cleaned = torch.empty_like(mag) if return_cleaned_mag else None
    _abi.check(lib.s5b200_istft(mag.data_ptr(), phase.data_ptr(), int(blf), mask.data_ptr() if mask is not None else None,
                                B, L, audio.data_ptr(), cleaned.data_ptr() if cleaned is not None else None, _stream()),
               "s5b200_istft")
    return (audio, cleaned) if return_cleaned_mag else audio


def _reduce(fn_name, a, b):
    import torch

    lib = _abi.lib()
    if not (a.is_cuda and a.dtype == torch.float32 and b.shape == a.shape and b.dtype == torch.float32):
        raise ValueError("operands must be float32 CUDA tensors of the same shape")
    a, b = a.contiguous(), b.contiguous()
    B = a.shape[0]
    n = a[0].numel()
    need = lib.s5b200_reduce_workspace_bytes(B, n)
    ws = torch.empty(max(need, 8), dtype=torch.uint8, device=a.device)
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    _abi.check(getattr(lib, fn_name)(a.data_ptr(), b.data_ptr(), B, n, out.data_ptr(), ws.data_ptr(), need, _stream()), fn_name)
    return out


def si_snr(target, estimate):
    """SI-SNR in dB per clip, last axis = time (si_snr_jax)."""
    return _reduce("s5b200_si_snr", target, estimate)


def mse_clips(a, b):
    """mean((a - b)^2) per clip."""
    return _reduce("s5b200_mse_clips", a, b)


def denoise_host(model, noisy_host, out=None, sync=True):
    """Forward of the NDNS evaluation step on HOST audio (CPU tensors, ideally pinned): (B, T) noisy -> (B, (L - 1) * 128) cleaned.

    One call of ``s5b200_ndns_denoise_host``: the batch flows in clip chunks through H2D | STFT -> model -> mask apply + iSTFT | D2H
    inside the library (``model`` is a float / static ``BoundModel``).  ``sync=False`` is stream-ordered like
    ``BoundModel.forward_host(sync=False)``: keep both buffers alive until the current stream has been synchronised; consecutive
    batches on two alternating streams keep both PCIe directions busy."""
    import torch

    lib = _abi.lib()
    if not (isinstance(noisy_host, torch.Tensor) and not noisy_host.is_cuda and noisy_host.dtype == torch.float32 and noisy_host.dim() == 2
            and noisy_host.is_contiguous()):
        raise ValueError("noisy_host must be a contiguous float32 CPU tensor of shape (B, T)")
    B, T = noisy_host.shape
    L = lib.s5b200_stft_frames(T)
    if out is None:
        out = torch.empty((B, (L - 1) * HOP), dtype=torch.float32)
        if noisy_host.is_pinned():
            out = out.pin_memory()
    if not (isinstance(out, torch.Tensor) and not out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
            and tuple(out.shape) == (B, (L - 1) * HOP)):
        raise ValueError("out must be a contiguous float32 CPU tensor of shape (B, (L - 1) * 128)")
    if not sync and not (noisy_host.is_pinned() and out.is_pinned()):
        import warnings
        warnings.warn("denoise_host(sync=False) with pageable host buffers: copies are not asynchronous; pin both buffers")
    fn = lib.s5b200_ndns_denoise_host if sync else lib.s5b200_ndns_denoise_host_async
    _abi.check(fn(model._handle, noisy_host.data_ptr(), out.data_ptr(), B, T, STFT_MAG_MEAN, _stream()), "s5b200_ndns_denoise_host")
    return out


def denoise_eval(model, noisy_audio, clean_audio):
    """Forward of the reference's NDNS evaluation step: returns (losses, si_snr, cleaned audio).

    ``model`` maps the (B, L, 257) input to the (B, L, 257) mask (a ``BoundModel`` or ``FxpRegressionModel.forward_float``-like
    callable).  Everything stays on the device: STFT -> model -> mask apply + iSTFT -> SI-SNR / loss.
    """
    nm, nph, xin = stft_splitter(noisy_audio, blf=True, model_input=True)
    cm, _ = stft_splitter(clean_audio, blf=True)
    mask = model(xin)
    cleaned, cleaned_mag = stft_mixer(nm, nph, blf=True, mask=mask, return_cleaned_mag=True)
    score = si_snr(_check_audio(clean_audio), cleaned)
    losses = LAMBDA_MSE * mse_clips(cleaned_mag, cm) + (100.0 - score)
    return losses, score, cleaned
