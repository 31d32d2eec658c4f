"""CPU: the NDNS pre/post-processing oracle (oracle/ndns.py) pinned against SciPy's stft / istft, which
jax.scipy.signal.stft / istft (the calls at train_helpers.py:1385-1393, 1404-1411) implement."""
import numpy as np
import pytest

from oracle import ndns as ON

scipy_signal = pytest.importorskip("scipy.signal")


@pytest.mark.parametrize("n_samples", [128 * 7, 1000, 4096 + 37])
def test_stft_matches_scipy(n_samples):
    rng = np.random.default_rng(1)
    audio = 0.05 * rng.standard_normal((3, n_samples))
    _, _, ref = scipy_signal.stft(audio, nperseg=512, nfft=512, noverlap=384, window="boxcar", return_onesided=True)
    got = ON.stft_complex(audio)
    assert got.shape == ref.shape == (3, 257, ON.n_frames(n_samples))
    assert np.abs(got - ref).max() <= 1e-15
    mag, ph = ON.stft_splitter(audio)
    assert np.array_equal(mag, np.abs(got)) and np.array_equal(ph, np.angle(got))


def test_reference_shapes():
    assert ON.n_frames(480000) == 3751            # the NDNS clip: 30 s at 16 kHz -> L = 3751


@pytest.mark.parametrize("L", [5, 40])
def test_istft_matches_scipy(L):
    rng = np.random.default_rng(2)
    spec = rng.standard_normal((2, 257, L)) + 1j * rng.standard_normal((2, 257, L))
    spec[:, 0].imag = 0
    spec[:, -1].imag = 0
    mag, ph = np.abs(spec), np.angle(spec)
    _, ref = scipy_signal.istft(mag * np.exp(1j * ph), nperseg=512, nfft=512, window="boxcar", noverlap=384, input_onesided=True)
    got = ON.stft_mixer(mag, ph)
    assert got.shape == ref.shape == (2, (L - 1) * 128)
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


def test_round_trip_and_si_snr():
    rng = np.random.default_rng(3)
    audio = 0.05 * rng.standard_normal((2, 128 * 30))
    mag, ph = ON.stft_splitter(audio)
    back = ON.stft_mixer(mag, ph)
    assert np.abs(back - audio).max() <= 1e-12
    # SI-SNR known answers: identical signals -> 10 log10(|s|^2 / 1e-8 + 1e-8); scale invariance; orthogonal noise
    t = rng.standard_normal((1, 4000))
    t -= t.mean()
    n = rng.standard_normal((1, 4000))
    n -= n.mean()
    n -= (n * t).sum() / (t * t).sum() * t
    est = 3.0 * t + n
    want = 10 * np.log10(9 * (t * t).sum() / ((n * n).sum() + 1e-8) + 1e-8)
    assert abs(ON.si_snr(t, est)[0] - want) <= 1e-9
    assert abs(ON.si_snr(t, 0.5 * est)[0] - 10 * np.log10(2.25 * (t * t).sum() / (0.25 * (n * n).sum() + 1e-8) + 1e-8)) <= 1e-9


def test_eval_tail_consistency():
    rng = np.random.default_rng(4)
    clean = 0.05 * rng.standard_normal((2, 128 * 20))
    noisy = clean + 0.02 * rng.standard_normal(clean.shape)
    nm, nph = ON.stft_splitter(noisy)
    cm, _ = ON.stft_splitter(clean)
    x = ON.model_input(nm)
    assert x.shape == (2, 21, 257) and x.dtype == np.float32
    zero_mask = np.zeros_like(x)
    losses, score, cleaned = ON.ndns_eval_tail(zero_mask, nm, nph, clean, cm)
    assert np.abs(cleaned - noisy).max() <= 1e-12           # mask 0: the noisy audio comes back
    assert np.allclose(score, ON.si_snr(clean, noisy))
    assert np.allclose(losses, 0.001 * np.mean((nm - cm) ** 2, axis=(1, 2)) + 100 - score)
