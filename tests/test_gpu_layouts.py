"""GPU tests of the input layouts behind dp_of1x1_batch_ex / dp_window_reduce_batch_ex / dp_channel_combine: reader batches
consumed in place, windows of continuous multi-channel int16 streams at trigger indices (reference
processing_data.py:643-688), weighted channel algebra (processing_data.py:1033-1047), per-fit lowchi2_fcutoff."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from detprocess_b200.synth import SynthSetup, make_traces, make_continuous  # noqa: E402
from oracle.of1x1 import of1x1_batch  # noqa: E402
from oracle import reductions as R  # noqa: E402


def _plan(S, precision, n_chan, adc=None, fcuts=(None,)):
    from detprocess_b200.core.plans import OFPlan
    pre = S.nb_pretrigger
    plan = OFPlan(S.nb_samples, S.fs, n_chan, precision)
    fits = []
    for c in range(n_chan):
        plan.set_psd(c, S.psd * (1 + c))
        if adc is not None:
            plan.set_adc_conversion(c, *adc[c])
        t = plan.add_template(c, S.template, pre)
        fits.append([plan.add_fit(c, t, pre - 500, pre + 500, lowchi2_fcutoff=f) for f in fcuts])
    plan.finalize()
    return plan, fits


@pytest.mark.parametrize('precision', ['f64', 'f32'])
@pytest.mark.parametrize('nb_samples', [32768, 4096])
def test_reader_batch_consumed_in_place(precision, nb_samples):
    """plan channels = rows (3, 1) of a 4-channel reader batch: identical numbers to the gathered [B, 2, N] batch"""
    S = SynthSetup(nb_samples)
    rng = np.random.default_rng(3)
    full = np.stack([make_traces(37, S.template, S.psd * (1 + (c % 2)), S.fs, rng) for c in range(4)], axis=1)
    plan, _ = _plan(S, precision, 2)
    x = torch.from_numpy(full).cuda()
    got = plan.run_layout(x, [3, 1]).cpu().numpy()
    ref = plan.run(x[:, [3, 1]].contiguous()).cpu().numpy()
    assert np.array_equal(got, ref)
    with pytest.raises(ValueError):
        plan.run_layout(x, [3, 4])


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_windows_of_two_channel_int16_streams(precision):
    """events = windows of continuous 2-channel int16 ADC streams at arbitrary (odd and even) start samples == the oracle on
    the windows gathered and converted on the host (adc * gain + offset, as H5Reader(adctoamp=True) hands them over)"""
    S = SynthSetup(16384)
    n, pre, fs = S.nb_samples, S.nb_pretrigger, S.fs
    L = 400_001
    gains, offs = (1.0e-11, 1.7e-11), (-2.0e-9, 4.0e-9)
    amps = np.stack([make_continuous(L, S.template, S.psd * (1 + c), fs, np.random.default_rng(50 + c), pulse_rate_hz=30.0) for c in range(2)])
    adc = np.stack([np.clip(np.round((amps[c] - offs[c]) / gains[c]), -32768, 32767).astype(np.int16) for c in range(2)])
    starts = np.array([0, 1, 12345, 77777, L - n, L - n - 1, 200_000, 33_333, -5, L - n + 1], dtype=np.int64)
    plan, fits = _plan(S, precision, 2, adc=[(gains[c], offs[c]) for c in range(2)])
    out = plan.run_layout(torch.from_numpy(adc).cuda(), [0, 1], torch.from_numpy(starts).cuda()).cpu().numpy()
    ok = (starts >= 0) & (starts + n <= L)
    assert np.all(out[~ok] == -999999.0) and ok.sum() == 8
    tol = dict(f64=(1e-9, 1e-9), f32=(1e-5, 1e-4))[precision]
    for c in range(2):
        win = np.stack([adc[c, s:s + n].astype(np.float64) * gains[c] + offs[c] for s in starts[ok]])
        o = of1x1_batch(win, S.template, S.psd * (1 + c), fs, pre, windows=[(pre - 500, pre + 500, False)])
        off = plan.fit_offset(c, fits[c][0])
        g = out[ok]
        same = g[:, off + 1].astype(np.int64) == o['ind'][0]
        ref = o if same.all() else None
        if ref is None:                 # fp32 near-tie: compare with the oracle at the kernel's index (and bound the tie)
            assert precision == 'f32'
            at = o['at'](g[:, off + 1].astype(np.int64))
            bound = 2 * tol[0] * np.maximum(o['amp'][0] ** 2, (5 * o['ampres']) ** 2) * o['norm']
            assert np.all(np.abs(at['chi2'] - o['chi2'][0]) <= bound)
            amp_ref, chi_ref = at['amp'], at['chi2']
        else:
            amp_ref, chi_ref = o['amp'][0], o['chi2'][0]
        assert np.max(np.abs(g[:, off] - amp_ref) / np.maximum(np.abs(amp_ref), 5 * o['ampres'])) < tol[0]
        assert np.max(np.abs(g[:, off + 2] / chi_ref - 1)) < tol[1]
        assert np.max(np.abs(g[:, plan.chi0_offset(c)] / o['chi0'] - 1)) < tol[1]


def test_window_reductions_at_start_offsets_bit_exact():
    from detprocess_b200.core.plans import ReducePlan
    n, L = 16384, 250_001
    rng = np.random.default_rng(8)
    gains, offs = (1.0e-11, 3.0e-11), (1.0e-9, -2.0e-9)
    adc = rng.integers(-30000, 30000, size=(2, L)).astype(np.int16)
    starts = np.array([0, 7, 12345, L - n, L - n + 1, -1, 99_999], dtype=np.int64)
    red = ReducePlan(n, 1.25e6, 2)
    h = []
    for c in range(2):
        red.set_adc_conversion(c, gains[c], offs[c])
        h.append((red.add(c, 'baseline', 0, 6000), red.add(c, 'integral', 7000, 9001), red.add(c, 'maximum'), red.add(c, 'minimum', 5, 16000)))
    red.finalize()
    out = red.run_layout(torch.from_numpy(adc).cuda(), [0, 1], torch.from_numpy(starts).cuda()).cpu().numpy()
    ok = (starts >= 0) & (starts + n <= L)
    assert np.all(out[~ok] == -999999.0)
    for c in range(2):
        win = np.stack([adc[c, s:s + n].astype(np.float64) * gains[c] + offs[c] for s in starts[ok]])
        assert np.array_equal(out[ok][:, red.column(h[c][0])], R.baseline_batch(win, 0, 6000))
        assert np.array_equal(out[ok][:, red.column(h[c][1])], R.integral_batch(win, 1.25e6, 7000, 9001))
        assert np.array_equal(out[ok][:, red.column(h[c][2])], R.maximum_batch(win, 0, n - 1))
        assert np.array_equal(out[ok][:, red.column(h[c][3])], R.minimum_batch(win, 5, 16000))
    # float64 batch in place: rows (2, 0) of a 3-channel batch
    x = rng.standard_normal((11, 3, n))
    red2 = ReducePlan(n, 1.25e6, 2)
    a = red2.add(0, 'baseline', 100, 9000)
    b = red2.add(1, 'integral', 0, n - 1)
    red2.finalize()
    o2 = red2.run_layout(torch.from_numpy(x).cuda(), [2, 0]).cpu().numpy()
    assert np.array_equal(o2[:, red2.column(a)], R.baseline_batch(x[:, 2], 100, 9000))
    assert np.array_equal(o2[:, red2.column(b)], R.integral_batch(x[:, 0], 1.25e6, 0, n - 1))


@pytest.mark.parametrize('dtype', ['float64', 'int16'])
def test_channel_algebra_bit_identical_to_numpy(dtype):
    from detprocess_b200.core.plans import combine_channels
    rng = np.random.default_rng(12)
    nb, n = 9, 4096
    gains, offs = (1.1e-11, 2.3e-11, 0.7e-11), (1e-9, -3e-9, 0.0)
    if dtype == 'int16':
        raw = rng.integers(-30000, 30000, size=(nb, 3, n)).astype(np.int16)
        conv = np.stack([raw[:, c].astype(np.float64) * gains[c] + offs[c] for c in range(3)], axis=1)
        adc = {c: (gains[c], offs[c]) for c in range(3)}
    else:
        raw = rng.standard_normal((nb, 3, n)) * 1e-8
        conv, adc = raw, None
    wa, wb, wc = 0.73, 1.9, -0.25
    terms = [[(0, None, 1.0), (2, None, 1.0)],              # a+c
             [(1, None, 1.0), (0, None, -1.0)],             # b-a
             [(0, wa, 1.0), (1, wb, 1.0), (2, wc, 1.0)],    # weighted a+b+c
             [(2, wb, 1.0), (1, wa, -1.0)]]                 # weighted c-b
    got = combine_channels(torch.from_numpy(raw).cuda(), terms, adc=adc).cpu().numpy()
    assert np.array_equal(got[:, 0], conv[:, 0] + conv[:, 2])
    assert np.array_equal(got[:, 1], conv[:, 1] - conv[:, 0])
    assert np.array_equal(got[:, 2], wa * conv[:, 0] + wb * conv[:, 1] + wc * conv[:, 2])
    assert np.array_equal(got[:, 3], wb * conv[:, 2] - wa * conv[:, 1])


@pytest.mark.parametrize('nb_samples', [32768, 8192])
def test_two_fits_with_their_own_lowchi2_cutoff_share_one_launch(nb_samples):
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    traces = make_traces(40, S.template, S.psd, S.fs, np.random.default_rng(4))
    plan, fits = _plan(S, 'f64', 1, fcuts=(10000.0, 2500.0, None))
    out = plan.run(torch.from_numpy(traces).cuda()).cpu().numpy()
    assert plan.launch_count() == 1
    for fi, fcut in zip(fits[0], (10000.0, 2500.0, 10000.0)):
        o = of1x1_batch(traces, S.template, S.psd, S.fs, pre, windows=[(pre - 500, pre + 500, False)], lowchi2_fcutoff=fcut)
        off = plan.fit_offset(0, fi)
        assert np.array_equal(out[:, off + 1].astype(np.int64), o['ind'][0])
        assert np.max(np.abs(out[:, off + 3] / o['lowchi2'][0] - 1)) < 1e-9
        assert np.max(np.abs(out[:, off] / o['amp'][0] - 1)[np.abs(o['amp'][0]) > 5 * o['ampres']]) < 1e-9


def test_stream_to_triggers_to_features_pipeline(tmp_path):
    """continuous int16 2-channel streams (raw ADC file) -> TriggerProcessing table -> FeatureProcessing(trigger_dataframe=...)
    == the oracle on the windows cut and converted on the host; the trigger rows come back verbatim next to the features"""
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.io import RawBinaryReader, write_raw_binary
    from detprocess_b200.process import TriggerProcessing, FeatureProcessing
    S = SynthSetup(16384)
    pre, fs, n = S.nb_pretrigger, S.fs, S.nb_samples
    L = 700_000
    gains, offs = [2.0e-11, 1.5e-11], [5.0e-9, -1.0e-9]
    ev = np.zeros((2, 2, L))
    for e in range(2):
        for c in range(2):
            ev[e, c] = make_continuous(L, S.template, S.psd, fs, np.random.default_rng(70 + 2 * e + c), pulse_rate_hz=0.0)
        for k, t in enumerate(60_001 + 71_003 * np.arange(8)):
            ev[e, 0, t:t + n - pre] += (1.0e-7 + 1e-8 * k) * S.template[pre:]
            ev[e, 1, t:t + n - pre] += 0.6e-7 * S.template[pre:]
    adc = np.stack([np.stack([np.clip(np.round((ev[e, c] - offs[c]) / gains[c]), -32768, 32767).astype(np.int16) for c in range(2)]) for e in range(2)])
    base = str(tmp_path / 'cont')
    write_raw_binary(base, adc, ['chanA', 'chanB'], fs, adc_gain=gains, adc_offset=offs,
                     admin={'event_time': [1_700_000_000, 1_700_000_010], 'series_num': [3, 3], 'event_num': [1, 2]})
    fd = FilterData()
    for c in ('chanA', 'chanB'):
        fd.set_psd(c, S.psd, sample_rate=fs)
        fd.set_template(c, S.template, sample_rate=fs, pretrigger_length_samples=pre)
    ytrig = tmp_path / 'trig.yaml'
    ytrig.write_text('trigger:\n    chanA:\n        run: True\n        threshold_sigma: 10\n        pileup_window_msec: 2\n')
    trig = TriggerProcessing(RawBinaryReader(base), str(ytrig), filter_data=fd, verbose=False).process()
    assert len(trig) == 16
    yfeat = tmp_path / 'feat.yaml'
    yfeat.write_text(f'''
global:
    trace_length_samples: {n}
    pretrigger_length_samples: {pre}
chanA,chanB:
    of1x1_constrained:
        run: True
        template_tag: default
        window_min_from_trig_usec: -400
        window_max_from_trig_usec: 400
    of1x1_nodelay:
        run: True
        template_tag: default
        lowchi2_fcutoff: 5000
    baseline:
        run: True
        window_min_from_start_usec: 0
        window_max_from_trig_usec: -1000
    maximum:
        run: True
''')
    fp = FeatureProcessing(RawBinaryReader(base), str(yfeat), filter_data=fd, verbose=False, trigger_dataframe=trig)
    df = fp.process()
    assert len(df) == 16 and np.array_equal(df['trigger_index'].values, trig['trigger_index'].values)
    assert np.array_equal(df['event_number'].values, trig['event_number'].values)
    conv = np.stack([[adc[e, c].astype(np.float64) * gains[c] + offs[c] for c in range(2)] for e in range(2)])
    for ci, c in enumerate(('chanA', 'chanB')):
        win = np.stack([conv[int(en) - 1, ci, int(ti) - pre:int(ti) - pre + n] for en, ti in zip(trig['event_number'], trig['trigger_index'])])
        o = of1x1_batch(win, S.template, S.psd, fs, pre, windows=[(pre - 500, pre + 500, False)])
        o5 = of1x1_batch(win, S.template, S.psd, fs, pre, windows=[(pre, pre + 1, False)], lowchi2_fcutoff=5000)
        assert np.allclose(df[f'amp_of1x1_constrained_{c}'], o['amp'][0], rtol=1e-9, atol=0)
        assert np.allclose(df[f'chi2_of1x1_constrained_{c}'], o['chi2'][0], rtol=1e-9, atol=0)
        assert np.allclose(df[f't0_of1x1_constrained_{c}'], o['t0'][0], rtol=0, atol=1e-12)
        assert np.allclose(df[f'lowchi2_of1x1_nodelay_{c}'], o5['lowchi2'][0], rtol=1e-9, atol=0)
        assert np.array_equal(df[f'baseline_{c}'].values, R.baseline_batch(win, 0, pre - 1250))
        assert np.array_equal(df[f'maximum_{c}'].values, R.maximum_batch(win, 0, n - 1))
    # the pulse sits at the trigger: the constrained fit finds it within a few samples of zero delay
    assert np.all(np.abs(df['t0_of1x1_constrained_chanA']) < 20 / fs)


def test_yaml_pipeline_with_weighted_channel_sum_from_int16_reader(tmp_path):
    """'chanA+chanB' with weights from an int16 reader: the channel-algebra kernel feeds the OF fit and the reductions;
    == oracle on w_a * a + w_b * b of the host-converted traces"""
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.io import ArrayReader
    from detprocess_b200.process import FeatureProcessing
    S = SynthSetup(4096)
    pre, fs, n = S.nb_pretrigger, S.fs, S.nb_samples
    gains, offs = [2.0e-11, 1.5e-11], [5.0e-9, -1.0e-9]
    amps = np.stack([make_traces(50, S.template, S.psd, fs, np.random.default_rng(60 + c)) for c in range(2)], axis=1)
    adc = np.stack([np.clip(np.round((amps[:, c] - offs[c]) / gains[c]), -32768, 32767).astype(np.int16) for c in range(2)], axis=1)
    conv = np.stack([adc[:, c].astype(np.float64) * gains[c] + offs[c] for c in range(2)], axis=1)
    wa, wb = 0.8, 1.3
    fd = FilterData()
    for c in ('chanA', 'chanB', 'chanA+chanB'):
        fd.set_psd(c, S.psd * (3 if '+' in c else 1), sample_rate=fs)
        fd.set_template(c, S.template, sample_rate=fs, pretrigger_length_samples=pre)
    y = tmp_path / 'f.yaml'
    y.write_text(f'''
global:
    trace_length_samples: {n}
    pretrigger_length_samples: {pre}
chanA:
    of1x1_nodelay:
        run: True
        template_tag: default
    baseline:
        run: True
chanA+chanB:
    weight_chanA: {wa}
    weight_chanB: {wb}
    of1x1_constrained:
        run: True
        template_tag: default
        window_min_from_trig_usec: -100
        window_max_from_trig_usec: 100
    integral:
        run: True
        window_min_from_trig_usec: -500
        window_max_from_trig_usec: 500
''')
    fp = FeatureProcessing(ArrayReader(torch.from_numpy(adc), ['chanA', 'chanB'], fs, adc_gain=gains, adc_offset=offs), str(y),
                           filter_data=fd, verbose=False)
    df = fp.process(batch_size=32)
    s = wa * conv[:, 0] + wb * conv[:, 1]
    o = of1x1_batch(s, S.template, S.psd * 3, fs, pre, windows=[(pre - 125, pre + 125, False)])
    assert np.allclose(df['amp_of1x1_constrained_chanA+chanB'], o['amp'][0], rtol=1e-9, atol=1e-9 * o['ampres'])
    assert np.allclose(df['chi2_of1x1_constrained_chanA+chanB'], o['chi2'][0], rtol=1e-9)
    assert np.array_equal(df['integral_chanA+chanB'].values, R.integral_batch(s, fs, pre - 625, pre + 625))
    o0 = of1x1_batch(conv[:, 0], S.template, S.psd, fs, pre, windows=[(pre, pre + 1, False)])
    assert np.allclose(df['amp_of1x1_nodelay_chanA'], o0['amp'][0], rtol=1e-9, atol=1e-9 * o0['ampres'])
    assert np.array_equal(df['baseline_chanA'].values, R.baseline_batch(conv[:, 0], 0, n - 1))


@pytest.mark.parametrize('nb_samples,precision', [(25000, 'f64'), (12500, 'f64'), (25000, 'f32'), (20000, 'f64')])
def test_non_power_of_two_trace_lengths(nb_samples, precision):
    """25000 / 12500 samples (20 ms / 10 ms at 1.25 MHz) are the reference's own example configuration
    (examples/processing/process_example.yaml:93-94); served by the mixed-radix kernel (dp_ofg_kernel.cuh)"""
    from detprocess_b200.core.plans import OFPlan
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    traces = make_traces(64, S.template, S.psd, S.fs, np.random.default_rng(11), offset=(3e-7 if precision == 'f64' else 0.0))
    plan = OFPlan(nb_samples, S.fs, 1, precision)
    plan.set_psd(0, S.psd)
    t0 = plan.add_template(0, S.template, pre)
    t1 = plan.add_template(0, S.template_glitch, pre)
    wins0 = [(None, None, False), (pre - 500, pre + 500, False), (pre, pre + 1, False)]
    wins1 = [(pre - 100, pre + 300, True)]
    f0 = [plan.add_fit(0, t0, lo, hi, outside) for lo, hi, outside in wins0]
    f1 = [plan.add_fit(0, t1, lo, hi, outside) for lo, hi, outside in wins1]
    plan.finalize()
    out = plan.run(torch.from_numpy(traces).cuda()).cpu().numpy()
    o0 = of1x1_batch(traces, S.template, S.psd, S.fs, pre, windows=wins0)
    o1 = of1x1_batch(traces, S.template_glitch, S.psd, S.fs, pre, windows=wins1)
    tol = dict(f64=(1e-9, 1e-9, 1e-9), f32=(1e-5, 1e-4, 1e-3))[precision]
    assert np.max(np.abs(out[:, plan.chi0_offset(0)] / o0['chi0'] - 1)) < tol[1]
    for fits, o in ((f0, o0), (f1, o1)):
        for iw, fi in enumerate(fits):
            off = plan.fit_offset(0, fi)
            ind = out[:, off + 1].astype(np.int64)
            ref = {k: o[k][iw] for k in ('amp', 'chi2', 'lowchi2')}
            if precision == 'f64':
                assert np.array_equal(ind, o['ind'][iw])
            elif not np.array_equal(ind, o['ind'][iw]):      # fp32 near-ties, same rule as tests/test_gpu_of1x1.py
                at = o['at'](ind)
                bound = 2 * tol[0] * np.maximum(o['amp'][iw] ** 2, (5 * o['ampres']) ** 2) * o['norm']
                assert np.all(np.abs(at['chi2'] - o['chi2'][iw]) <= bound)
                ref = {k: at[k] for k in ref}
            assert np.max(np.abs(out[:, off] - ref['amp']) / np.maximum(np.abs(ref['amp']), 5 * o['ampres'])) < tol[0]
            assert np.max(np.abs(out[:, off + 2] / ref['chi2'] - 1)) < tol[1]
            assert np.max(np.abs(out[:, off + 3] / ref['lowchi2'] - 1)) < tol[2]
    # int16 windows of a continuous stream through the same kernel
    if precision == 'f64':
        L = 5 * nb_samples + 3
        gain, offs = 1.0e-11, 2.0e-9
        stream = make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(12), pulse_rate_hz=100.0)
        adc = np.clip(np.round((stream - offs) / gain), -32768, 32767).astype(np.int16)
        plan2 = OFPlan(nb_samples, S.fs, 1, precision)
        plan2.set_psd(0, S.psd)
        plan2.set_adc_conversion(0, gain, offs)
        fi = plan2.add_fit(0, plan2.add_template(0, S.template, pre), pre - 500, pre + 500)
        plan2.finalize()
        starts = np.array([0, 1, nb_samples + 7, L - nb_samples, L - nb_samples + 1], dtype=np.int64)
        got = plan2.run_layout(torch.from_numpy(adc[None, :]).cuda(), [0], torch.from_numpy(starts).cuda()).cpu().numpy()
        assert np.all(got[-1] == -999999.0)
        win = np.stack([adc[s:s + nb_samples].astype(np.float64) * gain + offs for s in starts[:-1]])
        o = of1x1_batch(win, S.template, S.psd, S.fs, pre, windows=[(pre - 500, pre + 500, False)])
        off = plan2.fit_offset(0, fi)
        assert np.array_equal(got[:-1, off + 1].astype(np.int64), o['ind'][0])
        assert np.max(np.abs(got[:-1, off + 2] / o['chi2'][0] - 1)) < 1e-9


def test_reference_example_yaml_lengths_end_to_end(tmp_path):
    """the feature section of the reference's example YAML asks for 20 ms traces with 10 ms pretrigger at 1.25 MHz
    (25000 / 12500 samples): the pipeline processes them (OF blocks on the mixed-radix kernel, window features)"""
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.process import FeatureProcessing
    fs = 1.25e6
    S = SynthSetup(25000, fs, nb_pretrigger=12500)
    n, pre = S.nb_samples, S.nb_pretrigger
    traces = make_traces(40, S.template, S.psd, fs, np.random.default_rng(3))
    fd = FilterData()
    fd.set_psd('chan1', S.psd, sample_rate=fs)
    fd.set_template('chan1', S.template, sample_rate=fs, pretrigger_length_samples=pre)
    y = tmp_path / 'ex.yaml'
    y.write_text('''
global:
    trace_length_msec: 20
    pretrigger_length_msec: 10
chan1:
    of1x1_nodelay:
        run: True
        template_tag: default
    of1x1_constrained:
        run: True
        template_tag: default
        window_min_from_trig_usec: -400
        window_max_from_trig_usec: 400
    baseline:
        run: True
        window_min_from_start_usec: 0
        window_max_from_trig_usec: -1000
    integral:
        run: True
        window_min_from_trig_usec: -500
        window_max_from_trig_usec: 500
''')
    df = FeatureProcessing({'traces': traces[:, None, :], 'channels': ['chan1'], 'sample_rate': fs}, str(y), filter_data=fd, verbose=False).process()
    o = of1x1_batch(traces, S.template, S.psd, fs, pre, windows=[(pre - 500, pre + 500, False), (pre, pre + 1, False)])
    assert np.allclose(df['amp_of1x1_constrained_chan1'], o['amp'][0], rtol=1e-9, atol=1e-9 * o['ampres'])
    assert np.allclose(df['chi2_of1x1_constrained_chan1'], o['chi2'][0], rtol=1e-9)
    assert np.allclose(df['t0_of1x1_constrained_chan1'], o['t0'][0], rtol=0, atol=1e-12)
    assert np.allclose(df['amp_of1x1_nodelay_chan1'], o['amp'][1], rtol=1e-9, atol=1e-9 * o['ampres'])
    assert np.array_equal(df['baseline_chan1'].values, R.baseline_batch(traces, 0, pre - 1250))
    assert np.array_equal(df['integral_chan1'].values, R.integral_batch(traces, fs, pre - 625, pre + 625))


@pytest.mark.parametrize('nb_samples,precision', [(32768, 'f64'), (16384, 'f64'), (65536, 'f64'), (32768, 'f32'), (25000, 'f64')])
def test_interpolate_t0_parabola(nb_samples, precision):
    """interpolate / interpolate_t0 (reference algorithms.py:415, 543 -> qp.OF1x1.calc(interpolate_t0=True)): the kernel
    reports the amplitudes one sample before / after the best delay, the parabola (oracle.interpolate_parabola convention)
    refines amp, t0 and chi2.  Sub-sample delays are recovered better than the sample grid allows."""
    from detprocess_b200.core.ofbase import OFBaseBatch
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    from detprocess_b200.synth import make_template
    S = SynthSetup(nb_samples)
    pre, fs = S.nb_pretrigger, S.fs
    rng = np.random.default_rng(21)
    traces = make_traces(48, S.template, S.psd, fs, rng, amp_max=2e-7, max_delay=200)
    # a noiseless pulse delayed by a fraction of a sample: only the interpolated fit can see the fraction
    frac = 0.37
    t = (np.arange(nb_samples) - pre - frac) / fs
    tp = np.where(t > 0, t, 0.0)
    shifted = np.where(t > 0, np.exp(-tp / 200e-6) - np.exp(-tp / 20e-6), 0.0)
    traces[0] = 1.5e-7 * shifted / shifted.max()
    ofb = OFBaseBatch(fs, precision=precision)
    ofb.set_csd('c', S.psd)
    ofb.add_template('c', S.template, template_tag='default', pretrigger_samples=pre)
    ofb.update_signal('c', torch.from_numpy(traces).cuda())
    got_c = FE.of1x1_constrained('c', ofb, template_tag='default', window_min_from_trig_usec=-400, window_max_from_trig_usec=400,
                                 interpolate=True, feature_base_name='con')
    got_u = FE.of1x1_unconstrained('c', ofb, template_tag='default', interpolate=True, feature_base_name='unc')
    plain = FE.of1x1_constrained('c', ofb, template_tag='default', window_min_from_trig_usec=-400, window_max_from_trig_usec=400,
                                 feature_base_name='con')
    oc = of1x1_batch(traces, S.template, S.psd, fs, pre, windows=[(pre - 500, pre + 500, False)], interpolate=True)
    ou = of1x1_batch(traces, S.template, S.psd, fs, pre, windows=[(None, None, False)], interpolate=True)
    tol = dict(f64=(1e-9, 1e-9, 1e-12), f32=(2e-5, 2e-4, 2e-9))[precision]
    for got, o, name in ((got_c, oc, 'con'), (got_u, ou, 'unc')):
        same = np.abs(got[f't0_{name}'] - o['t0'][0]) < 0.6 / fs       # fp32: a near-tie may sit on the neighbouring sample
        assert same.mean() > (0.999 if precision == 'f64' else 0.95)
        den = np.maximum(np.abs(o['amp'][0]), 5 * o['ampres'])
        assert np.max((np.abs(got[f'amp_{name}'] - o['amp'][0]) / den)[same]) < tol[0]
        # (the noiseless event's chi2 is ~1e-7 of chi0: compare on the scale the subtraction chi0 - amp^2 norm works on)
        cden = np.maximum(np.abs(o['chi2'][0]), 1e-3 * o['chi0'])
        assert np.max((np.abs(got[f'chi2_{name}'] - o['chi2'][0]) / cden)[same]) < tol[1]
        big = same & (np.abs(o['amp'][0]) > 20 * o['ampres'])       # the vertex of a flat parabola is ill-conditioned
        assert np.max(np.abs(got[f't0_{name}'] - o['t0'][0])[big]) < tol[2]
    # sub-sample recovery on the noiseless event
    assert abs(got_c['t0_con'][0] * fs - frac) < 0.05 and abs(plain['t0_con'][0] * fs - frac) > 0.3
    # the plain fit of the same base is untouched by the interpolating one
    op = of1x1_batch(traces, S.template, S.psd, fs, pre, windows=[(pre - 500, pre + 500, False)])
    assert np.allclose(plain['t0_con'], op['t0'][0], atol=1e-15)
