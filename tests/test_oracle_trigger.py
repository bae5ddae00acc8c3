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


def test_dynamic_ranges_product_loop_equals_the_reference_loop():
    """host logic of dynamic=True: the product's running-maximum loop == the restated reference loop (oftrigger.py:78-141)"""
    from detprocess_b200.core.oftrigger import _dynamic_ranges
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = int(rng.integers(0, 400))
        x = np.cumsum(rng.integers(1, 60, size=n))
        vals = rng.exponential(50.0, size=n) + 25.0
        fn = lambda c: 5.0 + 0.5 * c    # noqa: E731
        a = _dynamic_ranges(x, vals, fn)
        b = T.getchangeslessthandynamicthresh(x, vals, fn)
        assert [tuple(r) for r in a] == [tuple(int(v) for v in r) for r in b]
    # constant window == the static grouping
    x = np.array([3, 4, 5, 20, 21, 50, 51, 52, 53, 200])
    assert [tuple(r) for r in _dynamic_ranges(x, np.ones(len(x)), lambda c: 10)] == [(0, 3), (3, 5), (5, 9), (9, 10)]


def test_residual_pass_recovers_a_pulse_hidden_in_the_tail_of_a_large_one():
    n, fs = 4096, 1.25e6
    template, psd, ofb = _setup(n, fs)
    phi_td = T.phi_td_from_phi_fd(ofb.phi('c', 'default'))
    norm = ofb.norm('c', 'default')
    rng = np.random.default_rng(5)
    x = make_continuous(100_000, template, psd, fs, rng, pulse_rate_hz=0.0)
    sig = 1.0 / np.sqrt(norm)
    t_big, t_small = 40_000, 40_000 + 600
    x[t_big - n // 2:t_big + n // 2] += 300 * sig * template
    x[t_small - n // 2:t_small + n // 2] += 60 * sig * template
    filtered, dchi2 = T.filter_trace(x, phi_td, 1.0 / norm, norm)
    thr, win = T.chi2_threshold(6.0), 1250
    first = T.find_triggers_once(dchi2, filtered, thr, win, 0, fs)
    assert len(first['trigger_index']) == 1 and abs(first['trigger_index'][0] - t_big) <= 1
    res = T.residual_delta_chi2(dchi2, filtered, first['trigger_index'], template, phi_td, 1.0 / norm, norm)
    assert np.all(res <= dchi2 + 1e-9 * dchi2.max())        # the subtracted shapes are non-negative
    second = T.find_triggers_once(res, filtered, thr, win, 0, fs)
    # the subtraction happens in delta-chi2 space (cross terms of the two pulses stay), so the second pass sees an excess
    # after the large pulse -- a new trigger in its tail -- not a clean copy of the small pulse
    new = np.setdiff1d(second['trigger_index'], first['trigger_index'])
    assert len(new) >= 1 and np.all((new > t_big) & (new < t_big + n // 2))
    both = T.combine_triggers(first, second)
    assert len(both['trigger_index']) == 1 + np.sum(~np.isin(second['trigger_index'], first['trigger_index']))
    # a saturated first-pass trigger is left alone
    flags = T.saturated_flags(x, first['trigger_index'], n, 100 * sig)
    assert flags.all()
    assert np.array_equal(T.residual_delta_chi2(dchi2, filtered, first['trigger_index'], template, phi_td, 1.0 / norm, norm,
                                                saturated=flags), dchi2)


def test_saturation_lowpass_on_a_window_equals_the_whole_trace_filter():
    """the product low-passes a window with a 4096-sample margin around each trigger instead of the whole stream
    (core/oftrigger.py::_saturated): a first-order Butterworth run forward and backward forgets its past within a few
    hundred samples, so both give the same values where the reference looks (oftrigger.py:776-786)"""
    from scipy.signal import butter, filtfilt
    fs, L, nt = 1.25e6, 400_000, 4096
    rng = np.random.default_rng(8)
    x = np.cumsum(rng.standard_normal(L)) * 1e-9 + rng.standard_normal(L) * 1e-8     # drifting baseline + noise
    b, a = butter(1, 50e3 / (0.5 * fs))
    full = filtfilt(b, a, x, padtype='even')
    q, margin = int(nt / 4), 4096
    for t in (nt, 5000, 123_456, L - nt, L - 1500):
        lo, hi = max(t - q, 0), min(t + q, L)
        wlo, whi = max(lo - margin, 0), min(hi + margin, L)
        seg = filtfilt(b, a, x[wlo:whi], padtype='even')[lo - wlo:hi - wlo]
        assert np.max(np.abs(seg - full[lo:hi])) < 1e-12 * np.max(np.abs(full))
