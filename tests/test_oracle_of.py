"""
Pins the OF1x1 oracle with convention-independent known-answer tests (the reference
ships no golden vectors for this path -- SURVEY.md F2/F3, 8(c)).
"""
import numpy as np
import pytest

from detprocess_b200.synth import SynthSetup, make_noise, make_traces
from oracle.of1x1 import OFBaseOracle, OF1x1Oracle, of1x1_batch


@pytest.fixture(scope='module')
def setup():
    return SynthSetup(4096)


def test_noiseless_pulse_recovered(setup):
    S = setup
    for A, d in [(3.0, 17), (-2.5, -211), (1e-7, 0)]:
        x = A * np.roll(S.template, d)
        o = of1x1_batch(x[None], S.template, S.psd, S.fs, S.nb_pretrigger, windows=[(None, None, False)])
        assert o['amp'][0, 0] == pytest.approx(A, rel=1e-10)
        assert o['ind'][0, 0] - S.nb_pretrigger == d
        assert o['t0'][0, 0] == d / S.fs
        assert abs(o['chi2'][0, 0]) < 1e-6 * o['chi0'][0]


def test_white_ensemble_statistics(setup):
    S = setup
    rng = np.random.default_rng(7)
    x = make_noise(4000, S.psd, S.fs, rng)
    o = of1x1_batch(x, S.template, S.psd, S.fs, S.nb_pretrigger, windows=[(None, None, False)])
    n = S.nb_samples
    # AC coupling drops the DC bin: E[chi0] = N - 1
    assert o['chi0'].mean() == pytest.approx(n - 1, rel=0.01)
    # no-delay amplitude is unbiased with std = ampres
    assert abs(o['amp0'].mean()) < 4 * o['ampres'] / np.sqrt(len(x))
    assert o['amp0'].std() == pytest.approx(o['ampres'], rel=0.05)
    assert o['chi2_0'].mean() == pytest.approx(n - 2, rel=0.01)


def test_parseval_chi0(setup):
    S = setup
    rng = np.random.default_rng(3)
    x = make_noise(4, S.psd, S.fs, rng)
    flat = np.full(S.nb_samples, 2.5e-22)
    o = of1x1_batch(x, S.template, flat, S.fs, S.nb_pretrigger, coupling='DC')
    # flat PSD: chi0 = sum(x^2) / (J fs)   (Parseval)
    assert np.allclose(o['chi0'], (x ** 2).sum(-1) / (2.5e-22 * S.fs), rtol=1e-12)


def test_class_path_matches_batch(setup):
    S = setup
    tr = make_traces(5, S.template, S.psd, S.fs, np.random.default_rng(11))
    pre = S.nb_pretrigger
    ob = of1x1_batch(tr, S.template, S.psd, S.fs, pre,
                     windows=[(pre - 500, pre + 500, False), (None, None, False)])
    ofb = OFBaseOracle(S.fs)
    ofb.set_csd('a', S.psd, coupling='AC')
    ofb.add_template('a', S.template, 'default', pretrigger_samples=pre)
    ofb.calc_phi('a', 'default')
    for i in range(len(tr)):
        ofb.clear_signal()
        assert not ofb.is_signal_stored('a')
        ofb.update_signal('a', tr[i], calc_fft=True)
        ofb.calc_signal_filt('a')
        ofb.calc_signal_filt_td('a')
        OF = OF1x1Oracle(ofb, 'a', 'default')
        OF.calc(window_min_index=pre - 500, window_max_index=pre + 500, lgc_fit_nodelay=True)
        a, t0, c, lc = OF.get_result_withdelay()
        assert a == pytest.approx(ob['amp'][0, i], rel=1e-12)
        assert t0 == ob['t0'][0, i]
        assert c == pytest.approx(ob['chi2'][0, i], rel=1e-10)
        assert lc == pytest.approx(ob['lowchi2'][0, i], rel=1e-10)
        a0, _, c0, lc0 = OF.get_result_nodelay()
        assert a0 == pytest.approx(ob['amp0'][i], rel=1e-12)
        assert c0 == pytest.approx(ob['chi2_0'][i], rel=1e-10)
        assert OF.get_chisq_nopulse() == pytest.approx(ob['chi0'][i], rel=1e-12)
        assert OF.get_energy_resolution() == pytest.approx(ob['ampres'], rel=1e-12)
        assert OF.get_time_resolution() == pytest.approx(ob['timeres'][0, i], rel=1e-10)
        OF2 = OF1x1Oracle(ofb, 'a', 'default')
        OF2.calc(lgc_fit_nodelay=False)
        assert OF2.get_result_withdelay()[0] == pytest.approx(ob['amp'][1, i], rel=1e-12)


def test_usec_window_wins_over_index(setup):
    S = setup
    tr = make_traces(2, S.template, S.psd, S.fs, np.random.default_rng(5))
    pre = S.nb_pretrigger
    ofb = OFBaseOracle(S.fs)
    ofb.set_csd('a', S.psd)
    ofb.add_template('a', S.template, 'default', pretrigger_samples=pre)
    ofb.update_signal('a', tr[0])
    OF = OF1x1Oracle(ofb, 'a')
    OF.calc(window_min_from_trig_usec=-400, window_max_from_trig_usec=400,
            window_min_index=0, window_max_index=10)
    r1 = OF.get_result_withdelay()
    OF.calc(window_min_index=pre - 500, window_max_index=pre + 500)
    assert r1 == OF.get_result_withdelay()


def test_outside_window_and_integralnorm(setup):
    S = setup
    pre = S.nb_pretrigger
    x = 2.0 * np.roll(S.template, 40)
    o = of1x1_batch(x[None], S.template, S.psd, S.fs, pre, windows=[(pre - 10, pre + 100, True)])
    assert not (pre - 10 <= o['ind'][0, 0] < pre + 100)
    o2 = of1x1_batch(x[None], S.template, S.psd, S.fs, pre, windows=[(None, None, False)], integralnorm=True)
    # integralnorm rescales the template by 1/s[0]: amplitude scales by s[0]
    s0 = np.fft.fft(S.template)[0] / S.nb_samples / (S.fs / S.nb_samples)
    assert o2['amp'][0, 0] == pytest.approx(2.0 * s0.real, rel=1e-9)


def test_against_qetpy_golden():
    """Upstream parity, where someone has run oracle/dump_golden.py with real QETpy (file not committed here)."""
    import os
    import pytest
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'of1x1_qetpy.npz')
    if not os.path.exists(path):
        pytest.skip('tests/golden/of1x1_qetpy.npz not present: QETpy is not installable in this environment')
    g = np.load(path)
    from detprocess_b200.synth import SynthSetup, make_traces
    from oracle.of1x1 import of1x1_batch
    S = SynthSetup(int(g['nb_samples']))
    pre = S.nb_pretrigger
    tr = make_traces(int(g['n_events']), S.template, S.psd, S.fs, np.random.default_rng(int(g['seed'])))
    o = of1x1_batch(tr, S.template, S.psd, S.fs, pre, windows=[(pre, pre + 1, False), (None, None, False), (pre - 500, pre + 500, False)])
    assert np.allclose(o['amp'][0], g['amp_nodelay'], rtol=1e-9)
    assert np.allclose(o['amp'][1], g['amp_un'], rtol=1e-9)
    assert np.allclose((o['ind'][1] - pre) / S.fs, g['t0_un'], rtol=0, atol=1e-12)
    assert np.allclose(o['amp'][2], g['amp_con'], rtol=1e-9)
    assert np.allclose(o['chi2'][1], g['chi2_un'], rtol=1e-9)
