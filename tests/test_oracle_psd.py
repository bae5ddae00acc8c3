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


def test_csd_oracle_diagonal_is_the_psd_and_matches_the_generating_csd():
    """calc_csd: hermitian in (a, b), conjugate-symmetric in k, diagonal == calc_psd, and its expectation is the CSD the
    synthetic traces were drawn from (complex off-diagonal terms included)."""
    from detprocess_b200.synth import SynthNxM
    from oracle.psd import calc_csd
    S = SynthNxM(2048, 3, 1)
    x = S.traces(600, np.random.default_rng(3), pulse_fraction=0.0)
    f, csd = calc_csd(x, S.fs)
    assert csd.shape == (3, 3, 2048)
    assert np.allclose(csd, np.conj(np.transpose(csd, (1, 0, 2))))
    assert np.allclose(csd[:, :, 1:], np.conj(csd[:, :, :0:-1]))
    for a in range(3):
        assert np.allclose(csd[a, a].real, P.calc_psd(x[:, a], S.fs)[1], rtol=1e-12)
    band = slice(4, 200)
    for a, b in [(0, 1), (0, 2), (1, 2)]:
        r = np.mean(csd[a, b, band] / S.csd[a, b, band])
        assert abs(r - 1) < 0.1
    cut = np.arange(600) % 3 != 0
    assert np.allclose(calc_csd(x, S.fs, cut)[1], calc_csd(x[cut], S.fs)[1])
