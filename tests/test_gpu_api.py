"""Drop-in API on the GPU: FeatureExtractors static methods and the YAML-driven
FeatureProcessing against the CPU oracle driven the way the reference drives QETpy."""
import textwrap
import warnings

import numpy as np
import pytest

from detprocess_b200.synth import SynthSetup, make_traces
from oracle.of1x1 import OFBaseOracle, OF1x1Oracle
from oracle import reductions as R
from oracle.windows import get_window_indices

pytestmark = pytest.mark.gpu


def _oracle_rows(S, traces, tags, algos):
    """Per-event loop shaped like the reference (features.py:533-851)."""
    ofb = OFBaseOracle(S.fs)
    ofb.set_csd('ch', S.psd, coupling='AC')
    for tag, tmpl in tags.items():
        ofb.add_template('ch', tmpl, tag, pretrigger_samples=S.nb_pretrigger)
        ofb.calc_phi('ch', tag)
    rows = []
    for x in traces:
        ofb.clear_signal()
        ofb.update_signal('ch', x, calc_fft=True)
        ofb.calc_signal_filt('ch')
        ofb.calc_signal_filt_td('ch')
        row = {}
        for name, (kind, tag, kw) in algos.items():
            OF = OF1x1Oracle(ofb, 'ch', tag)
            if kind == 'nodelay':
                OF.calc(lgc_fit_withdelay=False, lgc_fit_nodelay=True)
                a, _, c, lc = OF.get_result_nodelay()
                row.update({f'amp_{name}': a, f'chi2_{name}': c, f'lowchi2_{name}': lc})
            else:
                OF.calc(lgc_fit_withdelay=True, lgc_fit_nodelay=False, **kw)
                a, t0, c, lc = OF.get_result_withdelay()
                row.update({f'amp_{name}': a, f't0_{name}': t0, f'chi2_{name}': c, f'lowchi2_{name}': lc})
                if kind == 'constrained':
                    row.update({f'chi2nopulse_{name}': OF.get_chisq_nopulse(), f'ampres_{name}': OF.get_energy_resolution(),
                                f'timeres_{name}': OF.get_time_resolution()})
        rows.append(row)
    return rows


def _close(got, want, key, ampres):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    if key.startswith('t0_'):
        assert np.array_equal(got, want), key
    elif key.startswith('amp_'):
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 5 * ampres)) < 1e-9, key
    elif key.startswith('timeres_'):
        pass  # 1/|amp|: compared through amp
    else:
        assert np.max(np.abs(got / want - 1)) < 1e-9, key


def test_feature_extractors_drop_in():
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    from detprocess_b200.core.ofbase import OFBaseBatch
    S = SynthSetup(8192)
    pre = S.nb_pretrigger
    traces = make_traces(40, S.template, S.psd, S.fs, np.random.default_rng(7))
    ofb = OFBaseBatch(S.fs)
    ofb.set_csd('ch', S.psd, coupling='AC')
    ofb.add_template('ch', S.template, template_tag='default', pretrigger_samples=pre, overwrite=True)
    ofb.add_template('ch', S.template_glitch, template_tag='glitch', pretrigger_samples=pre, overwrite=True)
    ofb.calc_phi('ch', 'default')
    assert ofb.phi('ch', 'default').shape == (S.nb_samples,)
    ofb.clear_signal()
    assert not ofb.is_signal_stored('ch')
    ofb.update_signal('ch', traces, calc_fft=True)
    ofb.calc_signal_filt('ch')
    ofb.calc_signal_filt_td('ch')
    wmin, wmax = get_window_indices(S.nb_samples, pre, S.fs, window_min_from_trig_usec=-400, window_max_from_trig_usec=400)
    got = {}
    got.update(FE.of1x1_nodelay('ch', ofb, template_tag='default', feature_base_name='nd'))
    got.update(FE.of1x1_unconstrained('ch', ofb, template_tag='default', feature_base_name='un',
                                      window_min_index=3, some_unknown_kwarg=1))     # swallowed like the reference
    got.update(FE.of1x1_constrained('ch', ofb, template_tag='default', window_min_index=wmin, window_max_index=wmax,
                                    feature_base_name='co'))
    got.update(FE.of1x1_constrained('ch', ofb, template_tag='glitch', window_min_from_trig_usec=-400,
                                    window_max_from_trig_usec=400, window_min_index=0, window_max_index=10,
                                    feature_base_name='gl'))
    algos = {'nd': ('nodelay', 'default', {}), 'un': ('unconstrained', 'default', {}),
             'co': ('constrained', 'default', dict(window_min_index=wmin, window_max_index=wmax)),
             'gl': ('constrained', 'glitch', dict(window_min_from_trig_usec=-400, window_max_from_trig_usec=400,
                                                  window_min_index=0, window_max_index=10))}
    rows = _oracle_rows(S, traces, {'default': S.template, 'glitch': S.template_glitch}, algos)
    ampres = rows[0]['ampres_co']
    assert set(got) == set(rows[0])
    for key in rows[0]:
        _close(got[key], [r[key] for r in rows], key, ampres)
    # single-trace call returns scalars, like the reference
    ofb.clear_signal()
    ofb.update_signal('ch', traces[3])
    one = FE.of1x1_constrained('ch', ofb, template_tag='default', window_min_index=wmin, window_max_index=wmax,
                               feature_base_name='co')
    assert np.isscalar(one['amp_co']) and one['t0_co'] == rows[3]['t0_co']
    # trace extractors: bit-exact
    assert np.array_equal(FE.baseline(traces, 0, 1000)['baseline'], R.baseline_batch(traces, 0, 1000))
    assert FE.integral(traces[5], S.fs, 10, 500)['integral'] == R.integral(traces[5], S.fs, 10, 500)['integral']
    assert np.array_equal(FE.maximum(traces)['maximum'], R.maximum_batch(traces, 0, S.nb_samples - 1))
    assert np.array_equal(FE.minimum(traces, feature_base_name='mn')['mn'], R.minimum_batch(traces, 0, S.nb_samples - 1))
    with pytest.raises(ValueError):
        FE.maximum(traces, 7, 7)


def test_feature_processing_yaml_pipeline(tmp_path):
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.process.features import FeatureProcessing
    S = SynthSetup(4096)
    pre, fs, n = S.nb_pretrigger, S.fs, S.nb_samples
    rng = np.random.default_rng(3)
    tr_a = make_traces(50, S.template, S.psd, fs, rng)
    tr_b = make_traces(50, S.template, 2 * S.psd, fs, rng)
    traces = np.stack([tr_a, tr_b], axis=1)
    fd = FilterData()
    for ch, psd in (('chanA', S.psd), ('chanB', 2 * S.psd)):
        fd.set_psd(ch, psd, sample_rate=fs)
        fd.set_template(ch, S.template, sample_rate=fs, pretrigger_length_samples=pre)
        fd.set_template(ch, S.template_glitch, sample_rate=fs, pretrigger_length_samples=pre, tag='glitch')
    fd.save(str(tmp_path / 'filter.npz'))
    fd2 = FilterData()
    fd2.load(str(tmp_path / 'filter.npz'))
    yaml_text = textwrap.dedent(f'''
        global:
            trace_length_samples: {n}
            pretrigger_length_samples: {pre}
        chanA,chanB:
            of1x1_nodelay:
                run: True
                template_tag: default
            of1x1_constrained:
                run: True
                template_tag: default
                window_min_from_trig_usec: -400
                window_max_from_trig_usec: 400
            of1x1_glitch:
                run: True
                base_algorithm: of1x1_constrained
                template_tag: glitch
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
            energyabsorbed:
                run: False
        chanA:
            feature_channel: A
            of1x1_nodelay:
                run: True
                template_tag: default
            maximum:
                run: True
            user_rms:
                run: True
        chanA-chanB:
            weight_chanA: 2.0
            weight_chanB: 0.5
            minimum:
                run: True
        ''')
    cfg = tmp_path / 'proc.yaml'
    cfg.write_text(yaml_text)
    ext = tmp_path / 'features_user.py'
    ext.write_text(textwrap.dedent('''
        import numpy as np
        class FeatureExtractors:
            @staticmethod
            def user_rms(trace, feature_base_name='user_rms', **kwargs):
                return {feature_base_name: float(np.std(trace))}
        '''))
    fp = FeatureProcessing({'traces': traces, 'channels': ['chanA', 'chanB'], 'sample_rate': fs},
                           str(cfg), filter_data=fd2, external_file=str(ext))
    df = fp.process(batch_size=32)      # two batches, the second one ragged
    assert len(df) == 50
    # chanA was overwritten by its own block: only nodelay + maximum + user_rms, named with feature_channel "A"
    assert 'amp_of1x1_nodelay_A' in df and 'maximum_A' in df and 'amp_of1x1_constrained_chanA' not in df
    # chanB keeps the shared block
    algos = {'of1x1_nodelay': ('nodelay', 'default', {}),
             'of1x1_constrained': ('constrained', 'default', dict(window_min_from_trig_usec=-400, window_max_from_trig_usec=400)),
             'of1x1_glitch': ('constrained', 'glitch', dict(window_min_from_trig_usec=-400, window_max_from_trig_usec=400))}
    S2 = SynthSetup(4096)
    S2.psd = 2 * S.psd
    rows = _oracle_rows(S2, tr_b, {'default': S.template, 'glitch': S.template_glitch}, algos)
    ampres = rows[0]['ampres_of1x1_constrained']
    for key in rows[0]:
        _close(df[f'{key}_chanB'].to_numpy(), [r[key] for r in rows], key, ampres)
    rows_a = _oracle_rows(S, tr_a, {'default': S.template}, {'of1x1_nodelay': ('nodelay', 'default', {})})
    _close(df['amp_of1x1_nodelay_A'].to_numpy(), [r['amp_of1x1_nodelay'] for r in rows_a], 'amp_', ampres)
    # windows: from_start 0 -> from_trig -1000us ; from_trig +-500us
    a, b = get_window_indices(n, pre, fs, window_min_from_start_usec=0, window_max_from_trig_usec=-1000)
    assert np.array_equal(df['baseline_chanB'].to_numpy(), R.baseline_batch(tr_b, a, b))
    a, b = get_window_indices(n, pre, fs, window_min_from_trig_usec=-500, window_max_from_trig_usec=500)
    assert np.array_equal(df['integral_chanB'].to_numpy(), R.integral_batch(tr_b, fs, a, b))
    assert np.array_equal(df['maximum_A'].to_numpy(), R.maximum_batch(tr_a, 0, n - 1))
    # weighted difference channel (reference get_channel_trace :1033-1047)
    diff = tr_a * 2.0 - tr_b * 0.5
    assert np.array_equal(df['minimum_chanA-chanB'].to_numpy(), R.minimum_batch(diff, 0, n - 1))
    assert np.allclose(df['user_rms_A'].to_numpy(), tr_a.std(axis=1))


def test_pipeline_from_raw_adc_file_to_feature_dumps(tmp_path):
    """File -> features -> file without HDF5: int16 ADC events in a raw-binary container (detprocess_b200.io) go through
    the YAML pipeline (ADC -> amps on the device) and come out as parquet dumps; the numbers equal the pipeline run on the
    float64 amps the reference's reader would have produced (processing_data.py:674-684, adctoamp=True)."""
    import pandas as pd
    import torch
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.io import RawBinaryReader, write_raw_binary
    from detprocess_b200.process.features import FeatureProcessing
    from detprocess_b200.synth import SynthSetup, make_traces
    S = SynthSetup(16384)
    pre = S.nb_pretrigger
    gain, off = [1.0e-11, 1.3e-11], [2.0e-9, -1.0e-9]
    amps = np.stack([make_traces(300, S.template, S.psd, S.fs, np.random.default_rng(60 + c)) for c in range(2)], axis=1)
    adc = np.stack([np.clip(np.round((amps[:, c] - off[c]) / gain[c]), -32768, 32767) for c in range(2)], axis=1).astype(np.int16)
    base = str(tmp_path / 'raw_series')
    write_raw_binary(base, adc, ['chanA', 'chanB'], S.fs, adc_gain=gain, adc_offset=off,
                     admin={'event_number': np.arange(300) + 1, 'series_number': np.full(300, 42)})
    yml = tmp_path / 'cfg.yaml'
    yml.write_text('''
global:
    trace_length_samples: 16384
    pretrigger_length_samples: 8192
chanA,chanB:
    of1x1_constrained:
        run: True
        template_tag: default
        window_min_from_trig_usec: -400
        window_max_from_trig_usec: 400
    baseline:
        run: True
        window_min_from_start_usec: 0
        window_max_from_trig_usec: -1000
    maximum:
        run: True
''')
    fd = FilterData()
    for c in ('chanA', 'chanB'):
        fd.set_psd(c, S.psd, sample_rate=S.fs)
        fd.set_template(c, S.template, sample_rate=S.fs, pretrigger_length_samples=pre)
    fp = FeatureProcessing(RawBinaryReader(base), str(yml), filter_data=fd, processing_id='unit', verbose=False)
    df = fp.process(lgc_save=True, save_path=str(tmp_path / 'out'), batch_size=128, memory_limit=2e-5)
    assert len(df) == 300 and list(df['event_number']) == list(range(1, 301)) and (df['series_number'] == 42).all()
    assert len(fp.output_files) >= 2
    back = pd.concat([pd.read_parquet(f) for f in fp.output_files], ignore_index=True)
    assert back.equals(df)
    conv = np.stack([adc[:, c].astype(np.float64) * gain[c] + off[c] for c in range(2)], axis=1)
    ref = FeatureProcessing({'traces': torch.from_numpy(conv), 'channels': ['chanA', 'chanB'], 'sample_rate': S.fs},
                            str(yml), filter_data=fd, verbose=False).process()
    # the int16 events stay int16 on the device: the reductions are bit-identical to the float64 run (numpy's two-step
    # conversion in the load), the OF kernel converts with one fused multiply-add (differs in the last bit of a sample)
    assert fp._adc is not None
    for col in ('baseline_chanA', 'maximum_chanB', 'baseline_chanB', 'maximum_chanA'):
        assert np.array_equal(df[col].values, ref[col].values), col
    for col in ('amp_of1x1_constrained_chanA', 'chi2_of1x1_constrained_chanB', 'lowchi2_of1x1_constrained_chanA'):
        assert np.allclose(df[col].values, ref[col].values, rtol=1e-10, atol=0), col
    assert np.array_equal(df['t0_of1x1_constrained_chanB'].values, ref['t0_of1x1_constrained_chanB'].values)
