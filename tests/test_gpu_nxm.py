"""GPU: fused NxM optimal filter (dp_nxm_kernel.cuh through dp_ofnxm_batch) == oracle/ofnxm.py.
Tolerances: float64 mode amplitudes and chi2 to 1e-9 relative with identical delay index; float32 mode amplitudes to
1e-5 of the largest amplitude and chi2 to 1e-4, index identical except on near-ties (BASELINE.json north_star)."""
import numpy as np
import pytest

from detprocess_b200.synth import SynthNxM
from oracle.ofnxm import ofnxm_setup, ofnxm_batch

pytestmark = pytest.mark.gpu

TOL = {'f64': dict(amp=1e-9, chi2=1e-9), 'f32': dict(amp=1e-5, chi2=1e-4)}


def _plan(S, precision):
    from detprocess_b200.core.plans import NxMPlan
    plan = NxMPlan(S.nb_samples, S.fs, S.n_chan, S.n_templ, precision)
    plan.set_filter(S.templates, S.csd, S.nb_pretrigger, 'AC')
    plan.finalize()
    return plan


def _check(out, o, m, precision):
    tol = TOL[precision]
    same = out[:, 2].astype(np.int64) == o['ind']
    if precision == 'f64':
        assert same.all()
    else:
        assert same.mean() > 0.98
    scale = np.max(np.abs(o['amps']))
    assert np.max(np.abs(out[:, 0] / o['chi0'] - 1)) < tol['chi2']
    assert np.max(np.abs(out[:, 1] / o['chi2'] - 1)[same]) < tol['chi2']
    assert np.max(np.abs(out[:, 3:3 + m] - o['amps'])[same]) < tol['amp'] * scale
    assert np.max(np.abs(out[:, 3 + m] / o['chi2_0'] - 1)) < tol['chi2']
    assert np.max(np.abs(out[:, 4 + m:4 + 2 * m] - o['amps0'])) < tol['amp'] * scale


@pytest.mark.parametrize('precision,nb_samples,n,m', [('f64', 32768, 2, 2), ('f32', 32768, 2, 2), ('f64', 16384, 3, 3),
                                                      ('f32', 16384, 4, 1), ('f64', 65536, 2, 1), ('f32', 65536, 1, 2),
                                                      ('f64', 16384, 4, 3)])
def test_ofnxm_parity(precision, nb_samples, n, m):
    import torch
    S = SynthNxM(nb_samples, n, m)
    pre = S.nb_pretrigger
    st = ofnxm_setup(S.templates, S.csd, S.fs, pre)
    x = S.traces(300 if nb_samples < 65536 else 150, np.random.default_rng(12345))
    plan = _plan(S, precision)
    P, Pinv = plan.p_matrix()
    assert np.allclose(P, st['P'], rtol=1e-10) and np.allclose(Pinv, st['Pinv'], rtol=1e-8)
    xd = torch.from_numpy(x).cuda()
    for win in [(pre - 500, pre + 500, False), (None, None, False), (pre - 100, pre + 50, True)]:
        plan.set_window(*win)
        out = plan.run(xd).cpu().numpy()
        _check(out, ofnxm_batch(x, st, win), m, precision)


def test_ofnxm_noiseless_recovery_and_empty_window():
    """Injected amplitudes / delay come back exactly (convention independent); an empty window gives the sentinel for
    the delay fit and leaves the no-delay fit intact."""
    import torch
    S = SynthNxM(16384, 2, 2)
    pre = S.nb_pretrigger
    amps = np.array([[1.5, -0.5], [0.2, 3.0], [-2.0, 1e-3]])
    delays = [0, 123, -250]
    x = np.stack([sum(a[i] * np.roll(S.templates[:, i], d, axis=-1) for i in range(2)) for a, d in zip(amps, delays)])
    plan = _plan(S, 'f64')
    out = plan.run(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(out[:, 2].astype(int) - pre, delays)
    assert np.allclose(out[:, 3:5], amps, rtol=1e-9)
    assert np.all(np.abs(out[:, 1]) < 1e-8 * out[:, 0])
    plan.set_window(pre, pre, False)
    out2 = plan.run(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.all(out2[:, 1:5] == -999999.0)
    assert np.allclose(out2[:, 5:], out[:, 5:], rtol=1e-12)


def test_ofnxm_extractor_mirrors_reference_keys():
    """FeatureExtractors.ofnxm(channel='a|b', of_base, template_tag=...) returns the reference's keys
    (algorithms.py:229-272) with the oracle's numbers."""
    from detprocess_b200.core import FeatureExtractors, OFBaseBatch
    S = SynthNxM(16384, 2, 2)
    pre = S.nb_pretrigger
    x = S.traces(32, np.random.default_rng(7))
    ofb = OFBaseBatch(S.fs)
    ofb.set_csd('a|b', S.csd, coupling='AC')
    ofb.add_template('a|b', S.templates, template_tag='shared', pretrigger_samples=pre)
    ofb.calc_phi('a|b', template_tag='shared')
    empty = FeatureExtractors.ofnxm('a|b', ofb, template_tag='shared', amplitude_names=['phonon', 'glitch'])
    assert empty['chi2_ofnxm_constrained'] == -999999.0 and empty['glitch_ofnxm_nodelay'] == -999999.0
    ofb.update_signal('a|b', x)
    r = FeatureExtractors.ofnxm('a|b', ofb, template_tag='shared', amplitude_names=['phonon', 'glitch'],
                                window_min_from_trig_usec=-400, window_max_from_trig_usec=400)
    assert set(r) == {'chi2_ofnxm_constrained', 't0_ofnxm_constrained', 'phonon_ofnxm_constrained', 'glitch_ofnxm_constrained',
                      'chi2_ofnxm_nodelay', 'phonon_ofnxm_nodelay', 'glitch_ofnxm_nodelay'}
    o = ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, S.fs, pre), (pre - 500, pre + 500, False))
    assert np.allclose(r['t0_ofnxm_constrained'], o['t0'], rtol=0, atol=1e-12)
    assert np.allclose(r['phonon_ofnxm_constrained'], o['amps'][:, 0], rtol=1e-8, atol=1e-9 * np.max(np.abs(o['amps'])))
    assert np.allclose(r['chi2_ofnxm_nodelay'], o['chi2_0'], rtol=1e-9)
    with pytest.raises(ValueError):
        FeatureExtractors.ofnxm('a|b', ofb)                                   # template_tag is mandatory
    with pytest.raises(ValueError):
        FeatureExtractors.ofnxm('a|b', ofb, template_tag='shared', amplitude_names=['one'])
    one = OFBaseBatch(S.fs)
    one.set_csd('a|b', S.csd)
    one.add_template('a|b', S.templates, template_tag='shared', pretrigger_samples=pre)
    one.update_signal('a|b', x[0])
    r1 = FeatureExtractors.ofnxm('a|b', one, template_tag='shared')
    assert np.isscalar(r1['amp1_ofnxm_constrained']) or r1['amp1_ofnxm_constrained'].ndim == 0


def test_yaml_pipeline_with_joint_channel_ofnxm_block(tmp_path):
    """A YAML feature block on channel 'chanA|chanB' with base_algorithm: ofnxm (reference
    examples/salting/run46_salting_test.yaml:182-197) through FeatureProcessing: csd / templates come from FilterData
    (get_csd / get_template of the joint channel), the columns carry the feature_channel rename."""
    import torch
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.process.features import FeatureProcessing
    S = SynthNxM(16384, 2, 2)
    pre = S.nb_pretrigger
    x = S.traces(48, np.random.default_rng(11))
    yml = tmp_path / 'nxm.yaml'
    yml.write_text('''
global:
    trace_length_samples: 16384
    pretrigger_length_samples: 8192
chanA:
    of1x1_nodelay:
        run: True
        template_tag: default
    baseline:
        run: True
        window_min_from_start_usec: 0
        window_max_from_trig_usec: -1000
chanA|chanB:
    feature_channel: pair
    of2x2_test:
        run: True
        base_algorithm: ofnxm
        window_min_from_trig_usec: -100
        window_max_from_trig_usec: 100
        noise_tag: default
        template_tag: shared
        amplitude_names: [phonon, glitch]
''')
    fd = FilterData()
    fd.set_psd('chanA', np.real(S.csd[0, 0]), sample_rate=S.fs)
    fd.set_template('chanA', S.templates[0, 0] / S.templates[0, 0].max(), sample_rate=S.fs, pretrigger_length_samples=pre)
    fd.set_csd('chanA|chanB', S.csd, sample_rate=S.fs)
    fd.set_template('chanA|chanB', S.templates, sample_rate=S.fs, pretrigger_length_samples=pre, tag='shared')
    fp = FeatureProcessing({'traces': torch.from_numpy(x), 'channels': ['chanA', 'chanB'], 'sample_rate': S.fs},
                           str(yml), filter_data=fd, verbose=False)
    df = fp.process()
    lo, hi = pre - 125, pre + 125
    o = ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, S.fs, pre), (lo, hi, False))
    assert np.allclose(df['phonon_of2x2_test_constrained_pair'], o['amps'][:, 0], rtol=1e-8, atol=1e-9 * np.max(np.abs(o['amps'])))
    assert np.allclose(df['glitch_of2x2_test_nodelay_pair'], o['amps0'][:, 1], rtol=1e-8, atol=1e-9 * np.max(np.abs(o['amps'])))
    assert np.allclose(df['t0_of2x2_test_constrained_pair'], o['t0'], atol=1e-12)
    assert np.allclose(df['chi2_of2x2_test_constrained_pair'], o['chi2'], rtol=1e-9)
    assert 'amp_of1x1_nodelay_chanA' in df and 'baseline_chanA' in df
    assert np.allclose(df['baseline_chanA'], x[:, 0, :pre - 1250].mean(axis=1), rtol=1e-12)


def test_estimated_csd_feeds_the_nxm_filter():
    """Noise randoms -> NoiseCSD (device) -> NxM filter built from the estimate -> amplitudes of injected pulses come
    back within their resolution; the oracle run on the same estimated csd agrees to 1e-9."""
    import torch
    from detprocess_b200.core.noise import NoiseCSD
    S = SynthNxM(16384, 2, 2)
    pre = S.nb_pretrigger
    noise = S.traces(400, np.random.default_rng(50), pulse_fraction=0.0)
    est = NoiseCSD(S.nb_samples, S.fs, 2)
    est.update(torch.from_numpy(noise).cuda())
    _, csd = est.finalize()
    x, amps, delays = S.traces(64, np.random.default_rng(51), pulse_fraction=1.0, return_truth=True)
    from detprocess_b200.core.plans import NxMPlan
    plan = NxMPlan(S.nb_samples, S.fs, 2, 2, 'f64')
    plan.set_filter(S.templates, csd, pre, 'AC')
    plan.set_window(pre - 400, pre + 400)
    plan.finalize()
    out = plan.run(torch.from_numpy(x).cuda()).cpu().numpy()
    o = ofnxm_batch(x, ofnxm_setup(S.templates, csd, S.fs, pre), (pre - 400, pre + 400, False))
    assert np.array_equal(out[:, 2].astype(np.int64), o['ind'])
    assert np.max(np.abs(out[:, 3:5] - o['amps'])) < 1e-9 * np.max(np.abs(o['amps']))
    P, Pinv = plan.p_matrix()
    sigma = np.sqrt(np.diag(Pinv))
    assert np.all(np.abs(out[:, 3:5] - amps) < 6 * sigma)
    assert np.mean(np.abs(out[:, 2] - pre - delays) <= 2) > 0.9


def test_ofnxm_integralnorm_and_ignored_frequency_peaks():
    """OFBase.add_template(..., integralnorm=True) and OFBase.set_csd(..., ignored_frequency_peaks=, ignore_harmonics=)
    for joint channels (reference processing_data.py:321-326, 369-376) through OFBaseBatch == the oracle with the same
    options"""
    import torch
    from detprocess_b200.core.ofbase import OFBaseBatch
    S = SynthNxM(16384, 2, 2)
    pre, fs, N = S.nb_pretrigger, S.fs, S.nb_samples
    x = S.traces(60, np.random.default_rng(5))
    peaks = [60.0e3]
    ofb = OFBaseBatch(fs)
    ofb.set_csd('a|b', S.csd, coupling='AC', ignored_frequency_peaks=peaks, ignore_harmonics=True)
    ofb.add_template('a|b', S.templates, template_tag='t', pretrigger_samples=pre, integralnorm=True)
    ofb.update_signal('a|b', torch.from_numpy(x).cuda())
    r = ofb.nxm_results('a|b', 't', pre - 300, pre + 300)
    f = np.abs(np.fft.fftfreq(N, 1 / fs))
    csd = np.array(S.csd, dtype=np.complex128)
    for fpk in np.arange(peaks[0], fs / 2, peaks[0]):
        sel = np.abs(f - fpk) <= fs / N / 2
        for a in range(2):
            csd[a, a, sel] = np.inf
    assert np.isinf(csd.real).sum() > 10
    st = ofnxm_setup(S.templates, csd, fs, pre, integralnorm=True)
    o = ofnxm_batch(x, st, (pre - 300, pre + 300, False))
    assert np.array_equal(r['ind'], o['ind'])
    assert np.allclose(r['chi0'], o['chi0'], rtol=1e-9) and np.allclose(r['chi2'], o['chi2'], rtol=1e-9)
    assert np.max(np.abs(r['amps'] - o['amps'])) < 1e-9 * np.max(np.abs(o['amps']))
    # and the options matter: without them the numbers differ
    o_plain = ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, fs, pre), (pre - 300, pre + 300, False))
    assert np.max(np.abs(o_plain['chi0'] / o['chi0'] - 1)) > 1e-6
    assert np.max(np.abs(o_plain['amps'] - o['amps'])) > 1e-6 * np.max(np.abs(o['amps']))
