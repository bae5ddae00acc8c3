"""
Cross-check of oracle/of1x1.py that shares NOTHING with its FFT formulation (VERDICT r1, weak #1 iii).

The optimum filter is the generalised-least-squares fit of `x = A * s_d + n` with stationary noise n: with the noise
covariance matrix C (circulant, autocovariance R[m] = sum_k J(f_k) df cos(2 pi k m / N): Wiener-Khinchin, J the two-sided
PSD in A^2/Hz) and the template shifted by d samples,

    amp(d)  = s_d^T C^-1 x / s_d^T C^-1 s_d          chi2(d) = (x - amp s_d)^T C^-1 (x - amp s_d)

'AC' coupling (J[0] = inf in the frequency-domain form) = the constant direction carries no weight = data and template
are projected onto the zero-mean subspace.  Everything below is dense real linear algebra on explicit N x N matrices
(cosine sums, `np.linalg.solve`): no FFT, no 1/N/df convention, no Parseval.  It pins, to 1e-9, the amplitude at every
delay, the arg-min delay inside / outside a window and -- the part no self-consistency test can give -- the ABSOLUTE
scale of chi2 (and with it ampres = 1/sqrt(s^T C^-1 s)), which is where the oracle's `_fft_norm` convention enters.
"""
import numpy as np
import pytest

from detprocess_b200.synth import make_psd, make_template, make_traces
from oracle.of1x1 import of1x1_batch, OFBaseOracle, OF1x1Oracle


def _covariance(psd, fs):
    """circulant noise covariance from the two-sided PSD by an explicit cosine sum (no FFT)"""
    n = len(psd)
    df = fs / n
    k = np.arange(n)
    m = np.arange(n)
    J = np.array(psd, dtype=np.float64).copy()
    J[0] = J[1]                                    # any finite value: the constant direction is projected out below
    R = (J[None, :] * df * np.cos(2.0 * np.pi * np.outer(m, k) / n)).sum(axis=1)
    idx = (np.arange(n)[:, None] - np.arange(n)[None, :]) % n
    return R[idx]


def _gls(traces, template, psd, fs):
    n = len(template)
    C = _covariance(psd, fs)
    P = np.eye(n) - np.ones((n, n)) / n            # zero-mean projector ('AC' coupling)
    S = np.stack([np.roll(template, d) for d in range(n)], axis=1)      # column d = template delayed by d samples
    W = P @ np.linalg.solve(C, P)                  # C^-1 on the zero-mean subspace
    WS = W @ S
    sWs = np.einsum('nd,nd->d', S, WS)
    amps = (traces @ WS) / sWs[None, :]            # [B, delay]
    xWx = np.einsum('bn,nm,bm->b', traces, W, traces)
    chi2 = xWx[:, None] - amps ** 2 * sWs[None, :]
    return amps, chi2, xWx, sWs


@pytest.mark.parametrize('n', [256, 512, 1024])
def test_oracle_equals_time_domain_gls(n):
    fs = 1.25e6
    rng = np.random.default_rng(100 + n)
    pre = n // 2 - 7                               # pretrigger not at the centre: the roll convention is checked too
    template = make_template(n, fs, nb_pretrigger=pre, tau_rise=4e-6, tau_fall=30e-6)
    psd = make_psd(n, fs)
    traces = make_traces(12, template, psd, fs, rng, max_delay=n // 8, amp_max=2e-7)
    amps, chi2, xWx, sWs = _gls(traces, template, psd, fs)
    # column d of the GLS scan = template delayed by d relative to ITSELF; the oracle's rolled index is pre + d
    roll = lambda a: np.roll(a, pre, axis=-1)      # noqa: E731
    amps_r, chi2_r = roll(amps), roll(chi2)
    wins = [(None, None, False), (pre - n // 16, pre + n // 16, False), (pre, pre + 1, False), (pre - 5, pre + 40, True)]
    o = of1x1_batch(traces, template, psd, fs, pre, windows=wins)
    assert np.allclose(o['chi0'], xWx, rtol=1e-9, atol=0)                       # absolute chi2 scale
    assert o['norm'] == pytest.approx(sWs[0], rel=1e-9)                         # => ampres = 1/sqrt(s^T C^-1 s)
    assert np.allclose(sWs, sWs[0], rtol=1e-9)                                  # circularity of the setup itself
    rows = np.arange(len(traces))
    for iw, (lo, hi, outside) in enumerate(wins):
        mask = np.zeros(n, dtype=bool)
        mask[(0 if lo is None else lo):(n if hi is None else hi)] = True
        if outside:
            mask = ~mask
        ind = np.argmin(np.where(mask[None, :], chi2_r, np.inf), axis=1)
        assert np.array_equal(o['ind'][iw], ind), f'window {iw}'
        assert np.allclose(o['amp'][iw], amps_r[rows, ind], rtol=1e-9, atol=1e-9 / np.sqrt(sWs[0]))
        assert np.allclose(o['chi2'][iw], chi2_r[rows, ind], rtol=1e-9, atol=0)
    # amplitude at EVERY delay, through the object API the per-event loop of the reference drives
    ofb = OFBaseOracle(fs)
    ofb.set_csd('c', psd, coupling='AC')
    ofb.add_template('c', template, 'default', pretrigger_samples=pre)
    ofb.calc_phi('c', 'default')
    ofb.update_signal('c', traces[0])
    of = OF1x1Oracle(ofb, 'c', 'default')
    a_all, c_all, _ = of._arrays()
    scale = np.max(np.abs(amps_r[0]))
    assert np.max(np.abs(a_all - amps_r[0])) < 1e-9 * scale
    assert np.allclose(c_all, chi2_r[0], rtol=1e-9, atol=0)


def test_gls_recovers_an_injected_pulse_and_whitens_the_noise():
    """the cross-check itself is sane: noiseless delayed pulse -> exact amplitude / delay, chi2 = 0; noise-only ensemble
    -> E[x^T C^-1 x] = N - 1 (one degree of freedom removed with the mean)"""
    n, fs = 256, 1.25e6
    template = make_template(n, fs, tau_rise=4e-6, tau_fall=30e-6)
    psd = make_psd(n, fs)
    x = 3.5e-8 * np.roll(template, 19)[None, :]
    amps, chi2, _, _ = _gls(x, template, psd, fs)
    d = int(np.argmin(chi2[0]))
    assert d == 19 and amps[0, d] == pytest.approx(3.5e-8, rel=1e-9)
    assert abs(chi2[0, d]) < 1e-6
    noise = make_traces(3000, template, psd, fs, np.random.default_rng(4), pulse_fraction=0.0)
    _, _, xWx, _ = _gls(noise, template, psd, fs)
    assert np.mean(xWx) == pytest.approx(n - 1, rel=0.02)
