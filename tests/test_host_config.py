"""Host logic (no GPU): YamlConfig against the reference's saved output, channel algebra,
window indices, extractor sentinels / errors."""
import ast
import json
import os

import numpy as np
import pytest

from detprocess_b200.process.config import YamlConfig
from detprocess_b200.utils import utils
from oracle.windows import get_window_indices as oracle_window

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_yamlconfig_matches_reference_notebook_output():
    inp = json.load(open(os.path.join(GOLD, 'yaml_config_input.json')))
    expected = ast.literal_eval(open(os.path.join(GOLD, 'yaml_config_expected.pyl')).read())
    cfg = YamlConfig(inp['yaml'], inp['available_channels'], sample_rate=inp['sample_rate']).get_config()
    assert cfg['feature'] == expected['feature']
    assert cfg['global'] == expected['global']
    # spot checks of what the golden pins
    f = cfg['feature']
    assert f['traces_config'] == {(25000, 12500): ['Melange1pc1ch', 'Melange025pcLeft', 'Melange025pcRight', 'Melange4pc1ch']}
    assert 'energyabsorbed' not in f['channels']['Melange1pc1ch']          # run: False dropped
    assert f['channels']['Melange1pc1ch']['baseline']['nb_samples'] == 25000


def _write(tmp_path, text, name='c.yaml'):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_yamlconfig_semantics(tmp_path):
    chans = ['A', 'B', 'C']
    # duplicate keys are an error
    with pytest.raises(ValueError, match='Duplicate key'):
        YamlConfig(_write(tmp_path, 'A:\n  baseline:\n    run: True\n  baseline:\n    run: True\n'), chans, 1.25e6)
    # missing "run"
    with pytest.raises(ValueError, match='Missing "run"'):
        YamlConfig(_write(tmp_path, 'A:\n  baseline:\n    window_min_index: 3\n'), chans, 1.25e6)
    # include + obsolete keys + "all" + comma list + disable + channel-level lengths + weights
    inc = _write(tmp_path, 'filter_file: /x/filter.hdf5\n', 'inc.yaml')
    y = f'''
include: {inc}
global:
  trace_length_msec: 20
  pretrigger_length_msec: 10
all:
  baseline:
    run: True
A,B:
  of1x1_nodelay:
    run: True
    template_tag: default
    psd_tag: mypsd
    nb_samples: 4096
    nb_pretrigger_samples: 2048
C:
  disable: True
  maximum:
    run: True
A+B:
  weight_A: 0.5
  weight_B: 2
  integral:
    run: True
'''
    f = YamlConfig(_write(tmp_path, y), chans, 1.25e6).get_config('feature')
    assert f['overall']['filter_file'] == '/x/filter.hdf5'
    # "all" expands first, "A,B" overwrites A and B; a disabled block is skipped, so C keeps
    # what "all" gave it (reference config.py:226-248)
    assert set(f['channels']) == {'A', 'B', 'C', 'A+B'}
    assert list(f['channels']['C']) == ['baseline']
    a = f['channels']['A']['of1x1_nodelay']
    assert a['csd_tag'] == 'mypsd' and 'psd_tag' not in a    # obsolete key renamed
    assert (a['nb_samples'], a['nb_pretrigger_samples']) == (4096, 2048)
    assert f['channels']['A+B']['integral']['nb_samples'] == 25000
    assert f['weights'] == {'A+B': {'weight_A': 0.5, 'weight_B': 2}}
    assert f['traces_config'] == {(4096, 2048): ['A', 'B'], (25000, 12500): ['C', 'A', 'B']}
    assert f['channel_list'] == ['A', 'B', 'C']
    # msec lengths need a sample rate
    with pytest.raises(ValueError, match='sample rate'):
        YamlConfig(_write(tmp_path, 'global:\n  trace_length_msec: 20\n  pretrigger_length_msec: 10\nA:\n  baseline:\n    run: True\n'), chans)
    with pytest.raises(ValueError, match='pretrigger_length_samples'):
        YamlConfig(_write(tmp_path, 'A:\n  trace_length_samples: 100\n  baseline:\n    run: True\n'), chans, 1.25e6)


def test_split_channel_name():
    av = ['chanA', 'chanB', 'chanC']
    assert utils.split_channel_name('chanA', av) == (['chanA'], None)
    assert utils.split_channel_name('chanA+chanB', av) == (['chanA', 'chanB'], '+')
    assert utils.split_channel_name('chanA-chanB', av) == (['chanA', 'chanB'], '-')
    assert utils.split_channel_name('chanA|chanB|chanC', av) == (['chanA', 'chanB', 'chanC'], '|')
    assert utils.split_channel_name('chanA, chanB', av, separator=',') == (['chanA', 'chanB'], ',')
    assert utils.split_channel_name('chanA+chanB', av, separator=',') == (['chanA+chanB'], None)
    with pytest.raises(ValueError):
        utils.split_channel_name('chanA+nope', av)
    with pytest.raises(ValueError):
        utils.split_channel_name('a-b', None, separator='-')
    assert utils.split_channel_name('a|b', None, separator='|') == (['a', 'b'], '|')


def test_window_indices_match_oracle():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.choice([1000, 25000, 32768]))
        pre = int(rng.integers(0, n))
        kw = {}
        for side in ('min', 'max'):
            form = rng.choice(['from_start', 'to_end', 'from_trig', 'none'])
            if form != 'none':
                kw[f'window_{side}_{form}_usec'] = float(rng.uniform(-3e4, 3e4))
        try:
            want = oracle_window(n, pre, 1.25e6, **kw)
        except ValueError:
            with pytest.raises(ValueError):
                utils.get_window_indices(n, pre, 1.25e6, **kw)
            continue
        assert utils.get_window_indices(n, pre, 1.25e6, **kw) == want


def test_extractor_sentinels_and_errors_without_gpu():
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    from detprocess_b200.core.ofbase import OFBaseBatch
    # only the in-scope names are public (the pipeline treats every public name as an algorithm)
    assert sorted(m for m in dir(FE) if not m.startswith('_')) == sorted(
        ['of1x1_nodelay', 'of1x1_unconstrained', 'of1x1_constrained', 'ofnxm', 'baseline', 'integral', 'maximum', 'minimum',
         'psd_amp'])
    assert FE.baseline(None) == {'baseline': -999999.0}
    assert FE.integral(np.array([]), 1.25e6, feature_base_name='x') == {'x': -999999.0}
    assert FE.maximum(None, feature_base_name='m') == {'m': -999999.0}
    assert FE.minimum(None) == {'minimum': -999999.0}
    ofb = OFBaseBatch(1.25e6)
    with pytest.raises(ValueError, match='Template tag required'):
        FE.of1x1_nodelay('A', ofb)
    assert FE.of1x1_nodelay('A', ofb, template_tag='default', feature_base_name='nd') == {
        'amp_nd': -999999.0, 'chi2_nd': -999999.0, 'lowchi2_nd': -999999.0}
    assert set(FE.of1x1_unconstrained('A', ofb)) == {
        'amp_of1x1_unconstrained', 't0_of1x1_unconstrained', 'chi2_of1x1_unconstrained', 'lowchi2_of1x1_unconstrained'}
    r = FE.of1x1_constrained('A', ofb, feature_base_name='c')
    assert set(r) == {f'{k}_c' for k in ('amp', 't0', 'chi2', 'lowchi2', 'chi2nopulse', 'ampres', 'timeres')}
    assert all(v == -999999.0 for v in r.values())


def test_trigger_section_of_the_yaml(tmp_path):
    """Trigger section semantics of the reference's YamlConfig._configure_triggers (process/config.py:324-408)."""
    from detprocess_b200.process.config import YamlConfig
    yml = tmp_path / 'trig.yaml'
    yml.write_text('''
trigger:
    coincident_window_msec: 0.1
    chanA:
        run: True
        threshold_sigma: 10
        pileup_window_msec: 2
    chanB:
        trigger_name: B
        fast:
            run: True
            threshold_sigma: 8
            pileup_window_samples: 2500
        slow:
            run: False
            threshold_sigma: 8
    chanA+chanB:
        run: False
        threshold_sigma: 5
chanA:
    baseline:
        run: True
''')
    cfg = YamlConfig(str(yml), ['chanA', 'chanB'], sample_rate=1.25e6, verbose=False).get_config('trigger')
    assert list(cfg['channels']) == ['chanA', 'fast_B']
    assert cfg['channels']['fast_B']['channel_name'] == 'chanB'
    assert cfg['channels']['chanA'] == {'run': True, 'threshold_sigma': 10, 'pileup_window_msec': 2, 'channel_name': 'chanA'}
    assert cfg['overall']['coincident_window_msec'] == 0.1
    bad = tmp_path / 'bad.yaml'
    bad.write_text('trigger:\n    chanA:\n        fast:\n            threshold_sigma: 8\n')
    with pytest.raises(ValueError):
        YamlConfig(str(bad), ['chanA'], sample_rate=1.25e6, verbose=False)


def test_admin_columns_follow_the_reference_set_and_dtypes():
    """ProcessingData.get_event_admin (reference processing_data.py:811-887): fixed names and dtypes"""
    from detprocess_b200.process.features import standard_admin
    raw = {'event_num': [200003, 200004], 'series_num': [220230510153045] * 2, 'dump_num': [2, 2], 'event_index': [3, 4],
           'event_id': [3, 4], 'event_time': [1_700_000_000, 1_700_000_001], 'run_type': [1, 1], 'data_mode': ['cont', 'rand'],
           'fridge_run': [24, 24], 'series_start': [1_699_999_000] * 2, 'my_extra': [0.5, 1.5]}
    a = standard_admin(raw, 2, group_name='grp')
    names = ['event_number', 'event_index', 'dump_number', 'series_number', 'event_id', 'event_time', 'run_type', 'data_type',
             'group_name', 'trigger_type', 'trigger_amplitude', 'trigger_time', 'fridge_run_number', 'fridge_run_start_time',
             'series_start_time', 'group_start_time']
    assert list(a)[:len(names)] == names and list(a)[-1] == 'my_extra'
    assert a['event_number'].dtype == np.int64 and a['event_index'].dtype == np.int32 and a['dump_number'].dtype == np.int16
    assert a['series_number'].dtype == np.int64 and a['event_id'].dtype == np.int32 and a['event_time'].dtype == np.int64
    assert list(a['trigger_type']) == [1, 3] and np.isnan(a['trigger_amplitude']).all() and np.isnan(a['group_start_time']).all()
    assert list(a['run_type']) == ['1', '1'] and list(a['data_type']) == ['1', '1'] and list(a['group_name']) == ['grp', 'grp']
    assert a['fridge_run_number'].dtype == np.int64 and list(a['series_start_time']) == [1_699_999_000] * 2
    # a reader that only numbers its events: the DAQ's numbering convention fills the rest
    b = standard_admin({'event_number': np.array([300007])}, 1)
    assert b['event_index'][0] == 7 and b['dump_number'][0] == 3 and np.isnan(b['group_name'][0]) and np.isnan(b['trigger_type'][0])
