"""Known-answer tests that pin oracle/psd.py (no GPU)."""
import numpy as np

from oracle import psd as P


def test_parseval_variance():
    rng = np.random.default_rng(1)
    fs, n = 1.25e6, 4096
    tr = rng.standard_normal((64, n)) * 3e-9
    f, psd = P.calc_psd(tr, fs)
    df = fs / n
    # sum(psd) * df == mean over traces of mean(x^2)
    assert np.isclose(psd.sum() * df, np.mean(tr ** 2), rtol=1e-12)
    assert np.allclose(f, np.fft.fftfreq(n, 1 / fs))
    # two-sided and even in f
    assert np.allclose(psd[1:n // 2], psd[:n // 2:-1])


def test_white_noise_level_and_cut():
    rng = np.random.default_rng(2)
    fs, n, sig = 1.0e6, 1024, 2e-10
    tr = rng.standard_normal((2000, n)) * sig
    cut = rng.random(2000) < 0.7
    _, psd = P.calc_psd(tr, fs, cut)
    assert abs(np.mean(psd) / (sig ** 2 / fs) - 1) < 0.01
    s, c = P.periodogram_sums(tr, cut)
    assert c == int(cut.sum())
    assert np.allclose(s / (c * n * fs), psd[:n // 2 + 1], rtol=1e-12)


def test_sine_line():
    fs, n = 1.0e6, 2048
    k0 = 37
    t = np.arange(n) / fs
    x = 5e-9 * np.sin(2 * np.pi * k0 * fs / n * t)[None, :]
    _, psd = P.calc_psd(x, fs)
    # all power in bins +-k0: A^2/4 * N / fs each
    assert np.isclose(psd[k0], (5e-9) ** 2 / 4 * n / fs, rtol=1e-9)
    assert np.isclose(psd[n - k0], psd[k0], rtol=1e-12)
    assert psd[np.r_[0:k0 - 1, k0 + 2:n - k0 - 1]].max() < 1e-20 * psd[k0]


def test_offset_matches_numpy():
    rng = np.random.default_rng(3)
    tr = rng.standard_normal((10, 100)) + 4.0
    assert P.offset(tr) == np.average(np.median(tr, axis=-1))
