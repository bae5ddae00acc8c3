"""GPU parity of ``OptimumFilterTrigger.find_triggers`` with residual=True / dynamic=True (reference
core/oftrigger.py:78-141, 682-845) against oracle/trigger.py, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from detprocess_b200.synth import make_template, make_psd, make_continuous  # noqa: E402
from oracle import trigger as T  # noqa: E402


def _pileup_stream(nt, L, seed, pre):
    """noise + isolated pulses + small pulses riding on the tail of large ones (what the residual pass is for)"""
    fs = 1.25e6
    template = make_template(nt, fs, nb_pretrigger=pre)
    psd = make_psd(nt, fs)
    rng = np.random.default_rng(seed)
    x = make_continuous(L, template, psd, fs, rng, pulse_rate_hz=15.0)
    sig = np.std(x[:nt])
    for t0 in np.arange(3 * nt, L - 3 * nt, 5 * nt):
        big = 60 * sig
        x[t0 - pre:t0 - pre + nt] += big * template
        d = int(rng.integers(nt // 16, nt // 6))
        x[t0 + d - pre:t0 + d - pre + nt] += 0.25 * big * template
    return fs, template, psd, x


def _oracle_two_pass(trig, x, thresh, window, dynamic_fn=None, sat=None, positive=True):
    filtered, dchi2 = T.filter_trace(x, trig._phi_td, trig._iw_matrix, trig._w_matrix)
    thr = T.chi2_threshold(thresh)
    once = (lambda d: T.find_triggers_once_dynamic(d, filtered, thr, dynamic_fn, trig._trigger_index_shift, trig._fs)) if dynamic_fn \
        else (lambda d: T.find_triggers_once(d, filtered, thr, window, trig._trigger_index_shift, trig._fs))
    first = once(dchi2)
    flags = None
    if sat is not None:
        from scipy.signal import butter, filtfilt
        b, a = butter(1, 50e3 / (0.5 * trig._fs))
        flags = T.saturated_flags(filtfilt(b, a, x, padtype='even'), first['trigger_index'], len(trig._template), sat, positive)
    res = T.residual_delta_chi2(dchi2, filtered, first['trigger_index'], trig._template, trig._phi_td, trig._iw_matrix,
                                trig._w_matrix, saturated=flags)
    second = once(res)
    return first, second, T.combine_triggers(first, second), flags


@pytest.mark.parametrize('nt,pre,window', [(4096, 2048, 1250), (4096, 1500, 0), (8192, 4096, 300)])
def test_residual_retrigger_matches_oracle(nt, pre, window):
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    L = 400_000
    fs, template, psd, x = _pileup_stream(nt, L, 7, pre)
    trig = OptimumFilterTrigger('ch', fs, template, psd, pre, max_samples=L)
    trig.update_trace(torch.from_numpy(x).cuda())
    ret = trig.find_triggers(6.0, pileup_window_samples=window, residual=True, return_trigger_data=True, max_triggers=200_000)
    first, second, combined, _ = _oracle_two_pass(trig, x, 6.0, window)
    d1, l1, d2, l2 = ret
    assert np.array_equal(np.asarray(d1['ch']['trigger_index']), first['trigger_index'])
    assert len(second['trigger_index']) > 0
    # the second pass must find pulses the first pass merged away
    assert len(np.setdiff1d(second['trigger_index'], first['trigger_index'])) > 0 or window == 0
    assert np.array_equal(np.asarray(d2['ch']['trigger_index']), second['trigger_index'])
    assert np.allclose(d2['ch']['trigger_amplitude'], second['trigger_amplitude'], rtol=1e-9, atol=0)
    assert np.allclose(d2['ch']['trigger_delta_chi2'], second['trigger_delta_chi2'], rtol=1e-7, atol=0)
    d = trig.get_trigger_data()['ch']
    assert np.array_equal(np.asarray(d['trigger_index']), combined['trigger_index'])
    assert np.allclose(d['trigger_amplitude'], combined['trigger_amplitude'], rtol=1e-9, atol=0)
    assert np.allclose(d['trigger_delta_chi2'], combined['trigger_delta_chi2'], rtol=1e-7, atol=0)
    assert d['trigger_index_ch'] == d['trigger_index'] and len(d['trigger_channel']) == len(d['trigger_index'])
    # the sparse lists: every sample above threshold, before and after the subtraction
    filtered, dchi2 = T.filter_trace(x, trig._phi_td, trig._iw_matrix, trig._w_matrix)
    assert np.array_equal(l1['index'], np.where(dchi2 > T.chi2_threshold(6.0))[0])
    res = T.residual_delta_chi2(dchi2, filtered, first['trigger_index'], template, trig._phi_td, trig._iw_matrix, trig._w_matrix)
    assert np.array_equal(l2['index'], np.where(res > T.chi2_threshold(6.0))[0])
    assert np.allclose(l2['delta_chi2'], res[l2['index']], rtol=1e-7, atol=0)


def test_residual_skips_saturated_pulses_and_negative_polarity():
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    nt, pre, L = 4096, 2048, 300_000
    fs, template, psd, x = _pileup_stream(nt, L, 9, pre)
    trig = OptimumFilterTrigger('ch', fs, template, psd, pre, max_samples=L)
    trig.update_trace(torch.from_numpy(x).cuda())
    sat = 0.5 * np.max(x)
    trig.find_triggers(6.0, pileup_window_msec=1.0, residual=True, saturation_amplitudes_LPF_50kHz=[sat])
    _, _, combined, flags = _oracle_two_pass(trig, x, 6.0, int(1.0 * fs / 1000), sat=sat)
    assert flags.any() and not flags.all()
    assert np.array_equal(np.asarray(trig.get_trigger_data()['ch']['trigger_index']), combined['trigger_index'])
    # residual with nothing to subtract from: a quiet stream keeps an empty table
    quiet = np.random.default_rng(1).standard_normal(L) * 1e-12
    trig.update_trace(torch.from_numpy(quiet).cuda())
    trig.find_triggers(50.0, pileup_window_msec=1.0, residual=True)
    assert len(trig.get_trigger_data()['ch']['trigger_index']) == 0 and trig.get_trigger_data_df() is None


def test_dynamic_pileup_window_matches_oracle():
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    nt, pre, L = 4096, 2048, 300_000
    fs, template, psd, x = _pileup_stream(nt, L, 11, pre)
    trig = OptimumFilterTrigger('ch', fs, template, psd, pre, max_samples=L)
    trig.update_trace(torch.from_numpy(x).cuda())
    fn = lambda chi2: 20.0 + 40.0 * np.log10(chi2)   # noqa: E731  window (samples) grows with the pulse
    d = trig.find_triggers_once(5.0, dynamic=True, dynamic_threshold_function=fn)['ch']
    filtered, dchi2 = T.filter_trace(x, trig._phi_td, trig._iw_matrix, trig._w_matrix)
    o = T.find_triggers_once_dynamic(dchi2, filtered, T.chi2_threshold(5.0), fn, trig._trigger_index_shift, fs)
    assert len(o['trigger_index']) > 5
    assert np.array_equal(np.asarray(d['trigger_index']), o['trigger_index'])
    assert np.allclose(d['trigger_delta_chi2'], o['trigger_delta_chi2'], rtol=1e-9, atol=0)
    # a constant window function reproduces the static grouping
    const = trig.find_triggers_once(5.0, dynamic=True, dynamic_threshold_function=lambda c: 300)['ch']['trigger_index']
    static = trig.find_triggers_once(5.0, pileup_window_samples=300)['ch']['trigger_index']
    assert list(const) == list(static)
    # dynamic + residual
    trig.find_triggers(5.0, dynamic=True, dynamic_threshold_function=fn, residual=True)
    _, _, combined, _ = _oracle_two_pass(trig, x, 5.0, 0, dynamic_fn=fn)
    assert np.array_equal(np.asarray(trig.get_trigger_data()['ch']['trigger_index']), combined['trigger_index'])
    with pytest.raises(ValueError):
        trig.find_triggers_once(5.0, dynamic=True)


def test_filtered_at_equals_the_filtered_trace():
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    nt, L = 4095, 100_001
    fs = 1.25e6
    template, psd = make_template(nt, fs), make_psd(nt, fs)
    x = make_continuous(L, template, psd, fs, np.random.default_rng(2), pulse_rate_hz=50.0)
    trig = OptimumFilterTrigger('ch', fs, template, psd, nt // 2, max_samples=L)
    filtered, _ = T.filter_trace(x, trig._phi_td, trig._iw_matrix, trig._w_matrix)
    idx = np.array([0, 1, 17, nt // 2, nt, 50_000, L - nt, L - 2, L - 1], dtype=np.int64)
    for dt in (np.float64, np.float32):
        xs = torch.from_numpy(x.astype(dt)).cuda()
        got = trig._plan.filtered_at(xs, torch.from_numpy(idx)).cpu().numpy()
        ref = filtered if dt is np.float64 else T.filter_trace(x.astype(dt).astype(np.float64), trig._phi_td, trig._iw_matrix, trig._w_matrix)[0]
        assert np.allclose(got, ref[idx], rtol=1e-9, atol=1e-12 * np.max(np.abs(ref)))


def test_yaml_run_residual_adds_the_second_pass_triggers(tmp_path):
    """`run_residual: True` / `sat_amps_50kHz` of the YAML trigger section (reference process/triggers.py:742-751) reach
    find_triggers: the table holds the first-pass triggers and the new second-pass ones, like the oracle's two passes"""
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    from detprocess_b200.process import TriggerProcessing
    nt, pre, L = 4096, 2048, 300_000
    fs, template, psd, x = _pileup_stream(nt, L, 21, pre)
    text = '''
trigger:
    chanA:
        run: True
        threshold_sigma: 6
        pileup_window_msec: 1
        run_residual: %s
'''
    fd = FilterData()
    fd.set_psd('chanA', psd, sample_rate=fs)
    fd.set_template('chanA', template, sample_rate=fs, pretrigger_length_samples=pre)
    tables = {}
    for flag in ('False', 'True'):
        yml = tmp_path / f'trig_{flag}.yaml'
        yml.write_text(text % flag)
        tp = TriggerProcessing({'traces': torch.from_numpy(x[None, None, :]), 'channels': ['chanA'], 'sample_rate': fs,
                                'admin': [{'event_time': 1_700_000_000, 'series_num': 1, 'event_num': 1, 'dump_num': 1}]},
                               str(yml), filter_data=fd, processing_id='unit', verbose=False)
        tables[flag] = tp.process()
    trig = OptimumFilterTrigger('chanA', fs, template, psd, pre, max_samples=L)
    first, second, combined, _ = _oracle_two_pass(trig, x, 6.0, int(1.0 * fs / 1000))
    assert sorted(tables['False']['trigger_index']) == sorted(first['trigger_index'])
    assert sorted(tables['True']['trigger_index']) == sorted(combined['trigger_index'])
    assert len(tables['True']) > len(tables['False'])
