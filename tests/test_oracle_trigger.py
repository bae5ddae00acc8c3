"""Known-answer tests that pin oracle/trigger.py (no GPU)."""
import numpy as np

from detprocess_b200.synth import make_template, make_psd, make_continuous
from oracle import trigger as T
from oracle.of1x1 import OFBaseOracle


def test_chi2_threshold_sigma():
    # one amplitude: chi2 with 1 dof, P(chi2 > s^2) = 2 sf(s)
    for s in (1.0, 3.0, 5.0, 10.0):
        assert np.isclose(T.chi2_threshold(s), s * s, rtol=1e-9)
    assert T.chi2_threshold(30.0) == 900.0


def test_grouping_ranges():
    x = np.array([3, 4, 5, 20, 21, 50, 51, 52, 53, 200])
    r = T.getchangeslessthanthresh(x, 10)
    assert [tuple(a) for a in r] == [(0, 3), (3, 5), (5, 9), (9, 10)]
    r0 = T.getchangeslessthanthresh(x, 0)        # default pile-up window: every index is its own group
    assert len(r0) == len(x)
    assert len(T.getchangeslessthanthresh(np.array([], dtype=int), 5)) == 1   # one empty range, skipped by the caller


def _setup(n=4096, fs=1.25e6):
    template = make_template(n, fs)
    psd = make_psd(n, fs)
    ofb = OFBaseOracle(fs)
    ofb.set_csd('c', psd, coupling='AC')
    ofb.add_template('c', template, 'default', pretrigger_samples=n // 2)
    ofb.calc_phi('c', 'default')
    return template, psd, ofb


def test_injected_pulses_are_found_with_their_amplitude():
    n, fs = 4096, 1.25e6
    template, psd, ofb = _setup(n, fs)
    phi = ofb.phi('c', 'default')
    norm = ofb.norm('c', 'default')
    phi_td = T.phi_td_from_phi_fd(phi)
    rng = np.random.default_rng(11)
    x, t0, amps = make_continuous(200_000, template, psd, fs, rng, pulse_rate_hz=60.0, return_truth=True)
    filtered, dchi2 = T.filter_trace(x, phi_td, 1.0 / norm, norm)
    assert np.all(dchi2[:n] == 0) and np.all(dchi2[-(n - 1):] == 0) and dchi2[-n] != 0
    out = T.find_triggers_once(dchi2, filtered, T.chi2_threshold(8.0), int(1.0 * fs / 1000), 0, fs)
    # every isolated injected pulse away from the edges is found within 2 samples with its amplitude
    sig = 1.0 / np.sqrt(norm)
    found = 0
    for t, a in zip(t0, amps):
        if t < 2 * n or t > len(x) - 2 * n or a < 12 * sig:
            continue
        if np.min(np.abs(t0[t0 != t] - t), initial=10 ** 9) < 3000:
            continue
        d = np.abs(out['trigger_index'] - t)
        j = int(np.argmin(d))
        assert d[j] <= 2, (t, out['trigger_index'][j])
        assert abs(out['trigger_amplitude'][j] / a - 1) < 0.2
        assert np.isclose(out['trigger_delta_chi2'][j], out['trigger_amplitude'][j] ** 2 * norm)
        found += 1
    assert found >= 3
