"""CPU: EventBuilder (host mirror of the reference's core/eventbuilder.py) -- coincidence merge rules
(:336-497) and event metadata / ids (:178-333) on hand-built trigger tables."""
import numpy as np
import pandas as pd
import pytest

from detprocess_b200.core.eventbuilder import EventBuilder


def _table(chan, idx, dchi2, fs=1.25e6):
    n = len(idx)
    d = {'trigger_index': np.asarray(idx, dtype=np.int64), 'trigger_time': np.asarray(idx) / fs,
         'trigger_delta_chi2': np.asarray(dchi2, dtype=float), 'trigger_amplitude': np.asarray(dchi2, dtype=float) ** 0.5,
         'trigger_channel': [chan] * n}
    for k in ('trigger_index', 'trigger_delta_chi2', 'trigger_amplitude'):
        d[f'{k}_{chan}'] = d[k]
    return pd.DataFrame(d)


def test_coincident_triggers_merge_into_the_largest_delta_chi2():
    eb = EventBuilder()
    eb.add_trigger_data('A', _table('A', [1000, 50000, 90000], [30.0, 80.0, 25.0]))
    eb.add_trigger_data('B', _table('B', [1010, 70000, 90020], [60.0, 40.0, 20.0]))
    eb.build_event({'sample_rate': 1.25e6, 'nb_samples': 125000, 'event_time': 1000, 'series_num': 7, 'event_num': 3,
                    'run_type': 1}, coincident_window_samples=100)
    df = eb.get_event_df()
    assert list(df['trigger_index']) == [1010, 50000, 70000, 90000]
    assert list(df['trigger_channel']) == ['B', 'A', 'B', 'A']
    # the merged rows carry the other channel's own columns
    assert df.loc[0, 'trigger_index_A'] == 1000 and df.loc[0, 'trigger_delta_chi2_A'] == 30.0
    assert df.loc[3, 'trigger_index_B'] == 90020
    assert np.isnan(df.loc[1, 'trigger_index_B'])
    assert list(df['trigger_prod_id']) == [1, 2, 3, 4]
    assert list(df['series_number']) == [7] * 4 and list(df['event_number']) == [3] * 4
    assert list(df['data_type']) == ['1'] * 4
    assert list(df['event_time']) == [1000] * 4


def test_same_channel_neighbours_are_pileups_and_mixed_runs_split():
    eb = EventBuilder()
    #            pile-up pair (A,A)      mixed run A B A B: two coincidences
    eb.add_trigger_data('A', _table('A', [100, 150, 5000, 5080], [10., 11., 50., 20.]))
    eb.add_trigger_data('B', _table('B', [5040, 5120], [30., 90.]))
    eb.build_event({'sample_rate': 1.25e6, 'nb_samples': 10000}, coincident_window_samples=60)
    df = eb.get_event_df()
    assert list(df['trigger_index']) == [100, 150, 5000, 5120]
    assert df.loc[2, 'trigger_index_B'] == 5040 and df.loc[3, 'trigger_index_A'] == 5080
    assert list(df['event_time']) == [-1] * 4          # no event_time in the metadata


def test_trigger_ids_continue_across_events_and_window_zero_keeps_all():
    eb = EventBuilder()
    eb.add_trigger_data('A', _table('A', [10, 20], [1., 2.]))
    eb.add_trigger_data('B', _table('B', [12], [3.]))
    eb.build_event({'sample_rate': 1.0e3, 'nb_samples': 1000})
    assert len(eb.get_event_df()) == 3
    eb.clear_event()
    eb.add_trigger_data('A', _table('A', [5], [1.]))
    eb.build_event({'sample_rate': 1.0e3, 'nb_samples': 1000}, nb_trigger_channels=1)
    assert list(eb.get_event_df()['trigger_prod_id']) == [4]
    with pytest.raises(ValueError):
        eb.add_trigger_data('A', _table('A', [6], [1.]))
    with pytest.raises(ValueError):
        EventBuilder().acquire_triggers('nope', np.zeros(10), 5.0)
