"""CPU: invariants of the NxM optimal-filter oracle (oracle/ofnxm.py).  QETpy is absent, so the oracle is pinned on
what does not depend on QETpy's conventions: it reduces to the 1x1 oracle, recovers injected amplitudes and delays
exactly on noiseless data, and its chi2 is the residual of the fitted model."""
import numpy as np
import pytest

from detprocess_b200.synth import SynthNxM
from oracle.of1x1 import of1x1_batch, _fft_norm
from oracle.ofnxm import ofnxm_setup, ofnxm_batch


def test_nxm_oracle_reduces_to_1x1():
    S = SynthNxM(4096, 1, 1)
    st = ofnxm_setup(S.templates, S.csd, S.fs, S.nb_pretrigger)
    x = S.traces(20, np.random.default_rng(3))
    pre = S.nb_pretrigger
    for win in [(None, None, False), (pre - 100, pre + 200, False), (pre - 100, pre + 200, True)]:
        o = ofnxm_batch(x, st, win)
        r = of1x1_batch(x[:, 0], S.templates[0, 0], np.real(S.csd[0, 0]), S.fs, pre, windows=[win])
        assert np.array_equal(o['ind'], r['ind'][0])
        assert np.allclose(o['amps'][:, 0], r['amp'][0], rtol=1e-12, atol=0)
        assert np.allclose(o['chi2'], r['chi2'][0], rtol=1e-12)
        assert np.allclose(o['chi2_0'], r['chi2_0'], rtol=1e-12)
        assert np.allclose(o['chi0'], r['chi0'], rtol=1e-12)


@pytest.mark.parametrize('n,m', [(2, 2), (3, 2), (4, 3)])
def test_nxm_oracle_recovers_noiseless_amplitudes(n, m):
    S = SynthNxM(4096, n, m)
    pre = S.nb_pretrigger
    st = ofnxm_setup(S.templates, S.csd, S.fs, pre)
    rng = np.random.default_rng(4)
    amps = rng.uniform(-3, 3, size=(6, m))
    delays = np.array([0, 5, -17, 250, -300, 1])
    x = np.stack([sum(a[i] * np.roll(S.templates[:, i], d, axis=-1) for i in range(m)) for a, d in zip(amps, delays)])
    o = ofnxm_batch(x, st, (None, None, False))
    assert np.array_equal(o['ind'] - pre, delays)
    assert np.allclose(o['amps'], amps, rtol=1e-9)
    assert np.all(np.abs(o['chi2']) < 1e-8 * o['chi0'])
    # zero delay: the no-delay fit is exact as well
    assert np.allclose(o['amps0'][0], amps[0], rtol=1e-9)


def test_nxm_oracle_chi2_is_the_residual_of_the_fit():
    S = SynthNxM(2048, 2, 2)
    pre = S.nb_pretrigger
    st = ofnxm_setup(S.templates, S.csd, S.fs, pre)
    x = S.traces(5, np.random.default_rng(6), max_delay=100)
    o = ofnxm_batch(x, st, (pre - 200, pre + 200, False))
    df = S.fs / S.nb_samples
    for e in range(5):
        model = sum(o['amps'][e, i] * np.roll(S.templates[:, i], o['ind'][e] - pre, axis=-1) for i in range(2))
        R = _fft_norm(x[e] - model, S.fs)                      # [n, N]
        chi2 = np.real(np.einsum('ak,kab,bk->', np.conj(R), st['iS'], R)) * df
        assert chi2 == pytest.approx(o['chi2'][e], rel=1e-9)


def test_against_qetpy_golden():
    """Upstream parity of the NxM oracle and the csd convention, where someone has run oracle/dump_golden.py with real
    QETpy (the file cannot be produced in this environment and is not committed)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ofnxm_qetpy.npz')
    if not os.path.exists(path):
        pytest.skip('tests/golden/ofnxm_qetpy.npz not present: QETpy is not installable in this environment')
    from oracle.psd import calc_csd
    g = np.load(path)
    S = SynthNxM(int(g['nb_samples']), 2, 2)
    pre = S.nb_pretrigger
    x = S.traces(int(g['n_events']), np.random.default_rng(int(g['seed'])))
    o = ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, S.fs, pre), (pre - 500, pre + 500, False))
    assert np.allclose(o['amps'], g['amps'], rtol=1e-9)
    assert np.allclose(o['t0'], g['t0'], rtol=0, atol=1e-12)
    assert np.allclose(o['chi2'], g['chi2'], rtol=1e-9)
    assert np.allclose(o['amps0'], g['amps0'], rtol=1e-9)
    noise = S.traces(64, np.random.default_rng(int(g['seed']) + 1), pulse_fraction=0.0)
    assert np.allclose(calc_csd(noise, S.fs)[1], g['csd_of_noise'], rtol=1e-9)
