"""
``EventBuilder``: host mirror of the reference's ``detprocess/core/eventbuilder.py`` -- the step between the per-channel
stream triggers (``OptimumFilterTrigger``, device) and the feature extraction at the trigger indices
(``OFPlan.run_windows``, device).  It concatenates the per-channel trigger tables sorted by ``trigger_index``
(:126-175), merges triggers of different channels that fall within the coincidence window into the row with the
largest ``trigger_delta_chi2`` (:336-497; same-channel neighbours are pile-ups and stay separate; mixed runs are
split where a channel repeats) and attaches the event metadata / ids (:178-333).

Tables are pandas DataFrames (the reference uses vaex, which is not available here; the columns are the same).
The work is O(number of triggers) bookkeeping and stays on the host, as SURVEY.md 8(f) rank 2 prescribes.
"""
import numpy as np
import pandas as pd

__all__ = ['EventBuilder']


class EventBuilder:
    def __init__(self):
        self._event_df = None
        self._trigger_objects = None
        self._trigger_names = None
        self._current_event_time = 0
        self._current_trigger_id = 0
        self._current_nb_samples = None

    def clear_event(self):
        self._event_df = None
        self._trigger_names = None

    def get_event_df(self):
        return self._event_df

    def add_trigger_object(self, trigger_name, trigger_object):
        if self._trigger_objects is None:
            self._trigger_objects = dict()
        if trigger_name in self._trigger_objects:
            raise ValueError('ERROR: Trigger object "' + trigger_name + ' already stored!')
        self._trigger_objects[trigger_name] = trigger_object
        if self._trigger_names is None:
            self._trigger_names = list()
        self._trigger_names.append(trigger_name)

    def get_trigger_object(self, trigger_name):
        if self._trigger_objects is None or trigger_name not in self._trigger_objects:
            raise ValueError('ERROR: Trigger object "' + trigger_name + ' does not exist!')
        return self._trigger_objects[trigger_name]

    def _append(self, df):
        if df is None or len(df) == 0:
            return
        self._event_df = df if self._event_df is None else pd.concat([self._event_df, df], ignore_index=True)
        self._event_df = self._event_df.sort_values('trigger_index', kind='stable').reset_index(drop=True)

    def add_trigger_data(self, trigger_name, trigger_data):
        """trigger_data: DataFrame (or dict of columns) of one trigger channel."""
        if self._trigger_names is None:
            self._trigger_names = list()
        if trigger_name in self._trigger_names:
            raise ValueError('ERROR: Trigger data for channel ' + trigger_name + ' already added!')
        self._trigger_names.append(trigger_name)
        self._append(pd.DataFrame(trigger_data))

    def acquire_triggers(self, trigger_name, trace, thresh, pileup_window_msec=None, pileup_window_samples=None,
                         positive_pulses=True, run_residual=False, sat_amps_50kHz=None, edge_exclusion_msec=None,
                         livetime=None):
        if self._trigger_objects is None or trigger_name not in self._trigger_objects:
            raise ValueError('ERROR: Trigger object ' + trigger_name + ' not found!')
        obj = self._trigger_objects[trigger_name]
        obj.update_trace(trace)
        self._current_nb_samples = trace.shape[-1]
        obj.find_triggers(thresh, pileup_window_msec=pileup_window_msec, pileup_window_samples=pileup_window_samples,
                          positive_pulses=positive_pulses, residual=run_residual,
                          saturation_amplitudes_LPF_50kHz=sat_amps_50kHz, edge_exclusion_msec=edge_exclusion_msec,
                          livetime=livetime)
        self._append(obj.get_trigger_data_df())

    def build_event(self, event_metadata=None, fs=None, coincident_window_msec=None, coincident_window_samples=None,
                    nb_trigger_channels=None, trace_length_continuous_sec=None):
        if event_metadata is None:
            event_metadata = dict()
        if fs is None and 'sample_rate' in event_metadata:
            fs = event_metadata['sample_rate']
        if fs is None and coincident_window_msec is not None:
            raise ValueError('ERROR: sample rate required ("fs")')
        if trace_length_continuous_sec is None:
            if self._current_nb_samples is None and 'nb_samples' in event_metadata:
                self._current_nb_samples = event_metadata['nb_samples']
            if self._current_nb_samples is None or fs is None:
                raise ValueError('ERROR: "trace_length_continuous_sec" argument required!')
            trace_length_continuous_sec = self._current_nb_samples / fs
        event_time_start = np.nan
        event_time_end = np.nan
        if 'event_time' in event_metadata:
            event_time_data = event_metadata['event_time']
            event_time_start = event_time_data if event_time_data >= self._current_event_time else self._current_event_time
            event_time_end = int(event_time_start + trace_length_continuous_sec)
        self._current_event_time = event_time_end
        if self._event_df is None or len(self._event_df) == 0:
            return
        if nb_trigger_channels is None or nb_trigger_channels > 1:
            self._merge_coincident_triggers(fs=fs, coincident_window_msec=coincident_window_msec,
                                            coincident_window_samples=coincident_window_samples)
        df = self._event_df
        nb = len(df)
        strings = {'processing_id': None, 'data_type': None, 'group_name': None}
        for key in strings:
            if key in event_metadata:
                strings[key] = str(event_metadata[key]).replace('\0', '')
        if 'run_type' in event_metadata:
            strings['data_type'] = str(event_metadata['run_type']).replace('\0', '')
        for key, val in strings.items():
            df[key] = pd.array([val] * nb, dtype='string')
        ints = {k: np.full(nb, -1, dtype=np.int64) for k in
                ('series_number', 'event_number', 'dump_number', 'series_start_time', 'group_start_time',
                 'fridge_run_start_time', 'fridge_run_number')}
        for key in list(ints):
            if key in event_metadata:
                ints[key] = np.full(nb, np.int64(event_metadata[key]))
        for src, dst in (('series_num', 'series_number'), ('event_num', 'event_number'), ('dump_num', 'dump_number'),
                         ('fridge_run', 'fridge_run_number')):
            if src in event_metadata:
                ints[dst] = np.full(nb, np.int64(event_metadata[src]))
        event_times = df['trigger_time'].values + event_time_start
        if np.all(np.isfinite(event_times)):
            event_times_int = np.int64(np.around(event_times))
        else:   # no 'event_time' in the metadata: the reference casts NaN to int64 (undefined); kept as -1
            event_times_int = np.full(nb, -1, dtype=np.int64)
        ints['event_time'] = event_times_int
        for key in ('series_start_time', 'group_start_time', 'fridge_run_start_time'):
            ints[key] = event_times_int - ints[key]
        ints['trigger_prod_id'] = np.arange(nb, dtype=np.int64) + np.int64(self._current_trigger_id) + 1
        self._current_trigger_id = int(ints['trigger_prod_id'][-1])
        for key, val in ints.items():
            df[key] = val
        self._event_df = df

    def _merge_coincident_triggers(self, fs=None, coincident_window_msec=None, coincident_window_samples=None):
        if self._event_df is None or len(self._event_df) == 0:
            raise ValueError('ERROR: No trigger data available')
        merge_window = 0
        if coincident_window_msec is not None:
            if fs is None:
                raise ValueError('ERROR: sample rate "fs" needs to be provided!')
            merge_window = int(coincident_window_msec * fs / 1000)
        elif coincident_window_samples is not None:
            merge_window = coincident_window_samples
        if merge_window == 0:
            return
        df = self._event_df.reset_index(drop=True)
        idx = np.asarray(df['trigger_index'].values)
        dchi2 = np.asarray(df['trigger_delta_chi2'].values)
        names = np.asarray(df['trigger_channel'].values)
        # runs of neighbours closer than the window
        close = np.concatenate(([0], np.diff(idx) < merge_window, [0])).astype(int)
        ranges = np.where(np.abs(np.diff(close)) == 1)[0].reshape(-1, 2)
        groups = []
        for lo, hi in ranges:
            members = np.arange(lo, hi + 1)
            chans = names[members]
            nuniq = len(np.unique(chans))
            if nuniq == 1:
                continue                                    # one channel only: pile-up, not a coincidence
            if nuniq == len(chans):
                groups.append(members)
                continue
            # pile-ups and coincidences mixed: cut the run where a channel repeats
            cur_ch, cur = [], []
            for ch, k in zip(chans, members):
                if ch in cur_ch:
                    if len(cur) > 1:
                        groups.append(np.asarray(cur))
                    cur_ch, cur = [], []
                cur_ch.append(ch)
                cur.append(k)
            if len(cur) > 1:
                groups.append(np.asarray(cur))
        drop = []
        for members in groups:
            primary = int(members[np.argmax(dchi2[members])])
            for other in members[members != primary]:
                other = int(other)
                ch = str(names[other])
                row = df.iloc[other]
                for col in df.columns[row.notnull().values]:
                    if ch in str(col):                      # the other channel's own columns move to the primary row
                        df.at[primary, col] = row[col]
                drop.append(other)
        if drop:
            df = df.drop(index=drop).reset_index(drop=True)
        self._event_df = df
