"""GPU parity: NDNS pre/post-processing (STFT, mask apply + iSTFT, SI-SNR, loss) through the C ABI vs the oracle
(oracle/ndns.py, pinned against SciPy).  Tolerances: float32 transforms, 1e-5 relative to the largest value."""
import numpy as np
import pytest

from oracle import ndns as ON

pytestmark = pytest.mark.gpu
TOL = 1e-5


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("n_samples", [128 * 40, 1000, 128 * 33 + 77, 128 * 31])
@pytest.mark.parametrize("blf", [False, True])
def test_stft_matches_oracle(cuda_device, n_samples, blf):
    import torch
    from sparsernns_b200 import ndns

    rng = np.random.default_rng(n_samples)
    audio = (0.05 * rng.standard_normal((3, n_samples))).astype(np.float32)
    ref = ON.stft_complex(audio)                                # (B, 257, L) complex128
    mag, ph, xin = ndns.stft_splitter(torch.from_numpy(audio).to(cuda_device), blf=blf, model_input=True)
    mag, ph, xin = mag.cpu().numpy(), ph.cpu().numpy(), xin.cpu().numpy()
    if blf:
        mag, ph = mag.swapaxes(1, 2), ph.swapaxes(1, 2)
    assert mag.shape == ref.shape
    assert relerr(mag, np.abs(ref)) <= TOL
    assert relerr(mag * np.exp(1j * ph), ref) <= TOL            # the angle, compared where it is well conditioned
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    dphi = np.angle(np.exp(1j * (ph - np.angle(ref))))
    assert np.abs(dphi[big]).max() <= 1e-3
    assert np.array_equal(xin, (mag - np.float32(ON.STFT_MAG_MEAN)).swapaxes(1, 2))
    assert np.array_equal(xin.astype(np.float32), xin)


@pytest.mark.parametrize("L", [2, 5, 33, 70, 97])
@pytest.mark.parametrize("blf", [False, True])
def test_istft_matches_oracle(cuda_device, L, blf):
    import torch
    from sparsernns_b200 import ndns

    rng = np.random.default_rng(L)
    mag = np.abs(rng.standard_normal((2, 257, L))).astype(np.float32)
    ph = rng.uniform(-np.pi, np.pi, (2, 257, L)).astype(np.float32)
    mask = (0.3 * rng.standard_normal((2, L, 257))).astype(np.float32)
    ref = ON.stft_mixer(mag.astype(np.float64) * (1 + mask.swapaxes(1, 2).astype(np.float64)), ph)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a.swapaxes(1, 2) if blf else a)).to(cuda_device)
    audio, cleaned = ndns.stft_mixer(to(mag), to(ph), blf=blf, mask=torch.from_numpy(mask).to(cuda_device), return_cleaned_mag=True)
    assert audio.shape == ref.shape
    assert relerr(audio.cpu().numpy(), ref) <= TOL
    cl = cleaned.cpu().numpy()
    cl = cl.swapaxes(1, 2) if blf else cl
    assert np.array_equal(cl, mag * (np.float32(1) + mask.swapaxes(1, 2)))
    # no mask: plain stft_mixer
    ref0 = ON.stft_mixer(mag, ph)
    assert relerr(ndns.stft_mixer(to(mag), to(ph), blf=blf).cpu().numpy(), ref0) <= TOL


def test_round_trip_full_length(cuda_device):
    """A full 30 s clip (480 000 samples, L = 3751): istft(stft(x)) == x, the size-independent property."""
    import torch
    from sparsernns_b200 import ndns

    g = torch.Generator(device="cpu").manual_seed(0)
    audio = (0.05 * torch.randn(4, 480000, generator=g)).to(cuda_device)
    mag, ph = ndns.stft_splitter(audio)
    assert mag.shape == (4, 257, 3751)
    back = ndns.stft_mixer(mag, ph)
    assert back.shape == audio.shape
    assert float((back - audio).abs().max()) <= 2e-5 * float(audio.abs().max())


def test_si_snr_and_loss_match_oracle(cuda_device):
    import torch
    from sparsernns_b200 import ndns

    rng = np.random.default_rng(7)
    clean = (0.05 * rng.standard_normal((3, 128 * 150 + 5))).astype(np.float32)
    for noise in (0.5, 0.02, 1e-4):
        est = (1.7 * clean + noise * 0.05 * rng.standard_normal(clean.shape) + 0.01).astype(np.float32)
        got = ndns.si_snr(torch.from_numpy(clean).to(cuda_device), torch.from_numpy(est).to(cuda_device)).cpu().numpy()
        ref = ON.si_snr(clean, est)
        assert np.abs(got - ref).max() <= 1e-3, (noise, got, ref)   # dB
    a = rng.standard_normal((3, 257, 40)).astype(np.float32)
    b = rng.standard_normal((3, 257, 40)).astype(np.float32)
    got = ndns.mse_clips(torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device)).cpu().numpy()
    assert np.allclose(got, ((a.astype(np.float64) - b) ** 2).mean((1, 2)), rtol=1e-6)


def test_denoise_eval_matches_oracle(cuda_device):
    """audio -> STFT -> (a fixed random mask standing in for the model) -> iSTFT -> SI-SNR / loss, vs the oracle chain."""
    import torch
    from sparsernns_b200 import ndns

    rng = np.random.default_rng(11)
    clean = (0.05 * rng.standard_normal((2, 128 * 60))).astype(np.float32)
    noisy = (clean + 0.02 * rng.standard_normal(clean.shape)).astype(np.float32)
    L = ON.n_frames(clean.shape[1])
    mask_np = (0.1 * rng.standard_normal((2, L, 257))).astype(np.float32)
    seen = {}

    def model(xin):
        seen["x"] = xin.cpu().numpy()
        return torch.from_numpy(mask_np).to(xin.device)

    losses, score, cleaned = ndns.denoise_eval(model, torch.from_numpy(noisy).to(cuda_device), torch.from_numpy(clean).to(cuda_device))
    nm, nph = ON.stft_splitter(noisy)
    cm, _ = ON.stft_splitter(clean)
    assert relerr(seen["x"], ON.model_input(nm)) <= TOL
    l_ref, s_ref, c_ref = ON.ndns_eval_tail(mask_np, nm, nph, clean, cm)
    assert relerr(cleaned.cpu().numpy(), c_ref) <= TOL
    assert np.abs(score.cpu().numpy() - s_ref).max() <= 1e-3
    assert np.abs(losses.cpu().numpy() - l_ref).max() <= 1e-3
