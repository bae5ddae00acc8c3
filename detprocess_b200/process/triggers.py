"""
``TriggerProcessing``: YAML-driven trigger driver -- the in-memory counterpart of the reference's
``detprocess/process/triggers.py`` (``TriggerProcessing.process`` :228, trigger loop :714-800): per continuous event,
every configured trigger channel runs through ``EventBuilder.acquire_triggers`` (``OptimumFilterTrigger`` on the
device), coincident triggers are merged and the event metadata attached; the per-event tables are concatenated.

The reference reads continuous events from pytesdaq HDF5 files and writes vaex-HDF5 dumps; neither library exists in
this image, so the raw data arrives as a dict of arrays (like ``FeatureProcessing``) and the result is a pandas
DataFrame (optionally saved as parquet).  Continuous events are sharded over ranks (one process per GPU); the merge
is per event, so no collective is needed beyond the final gather of the small tables.
"""
import copy
import os

import numpy as np

from ..core.eventbuilder import EventBuilder
from ..core.oftrigger import OptimumFilterTrigger
from .config import YamlConfig
from .features import dist_info, shard_range, gather_frames
from .. import utils

__all__ = ['TriggerProcessing']


class TriggerProcessing:
    def __init__(self, raw_data, config_file, filter_data=None, processing_id=None, edge_exclusion_msec=None,
                 livetime=None, precision='f64', device=None, verbose=True):
        """
        raw_data : a ``detprocess_b200.io.EventReader`` of continuous events, or a dict with
            'traces'      ndarray / torch tensor [n_events, n_chan, L] continuous events (float64 / float32 / int16)
            'channels'    list of channel names (length n_chan)
            'sample_rate' float
            'admin'       optional list of per-event metadata dicts (event_time, series_num, event_num, dump_num, ...)
        config_file : YAML path with a ``trigger:`` section (reference process/config.py:324-408)
        filter_data : FilterData with the templates / PSDs the YAML refers to
        """
        self._verbose = verbose
        from ..io.readers import EventReader, ArrayReader
        if isinstance(raw_data, EventReader):
            self._reader = raw_data
            self._admin = None
        else:
            tr = raw_data['traces']
            if tr.ndim == 2:          # one continuous event [n_chan, L]
                tr = tr[None]
            self._reader = ArrayReader(tr, raw_data['channels'], raw_data['sample_rate'],
                                       adc_gain=raw_data.get('adc_gain'), adc_offset=raw_data.get('adc_offset'))
            self._admin = raw_data.get('admin')
        self._channels = self._reader.channels
        self._fs = self._reader.sample_rate
        self._filter_data = filter_data
        self._processing_id = processing_id
        self._edge_exclusion_msec = edge_exclusion_msec
        self._livetime = livetime
        self._precision = precision
        self._device = device
        cfg = YamlConfig(config_file, self._channels, sample_rate=self._fs, verbose=verbose).get_config('trigger')
        self._trigger_config = cfg['channels']
        if not self._trigger_config:
            raise ValueError('ERROR: no trigger channel enabled in the yaml file')
        self._evtbuilder_config = cfg['overall']
        if filter_data is None:
            raise ValueError('ERROR: filter data (templates and noise spectra) required')

    def _channel_trace(self, traces, channel):
        """one continuous event [n_chan, L] -> the trigger channel's stream (sums / differences like the reference's
        get_channel_trace, processing_data.py:941-1049)"""
        parts, sep = utils.split_channel_name(channel, available_channels=self._channels)
        idx = [self._channels.index(c) for c in parts]
        if sep is None:
            return traces[idx[0]]
        if sep == '+':
            out = traces[idx[0]]
            for i in idx[1:]:
                out = out + traces[i]
            return out
        if sep == '-':
            return traces[idx[0]] - traces[idx[1]]
        raise NotImplementedError(f'channel operator "{sep}" is outside the built trigger path (1x1 triggers)')

    def _build_triggers(self, max_samples):
        eb = EventBuilder()
        for trig_chan, td in self._trigger_config.items():
            chan = td['channel_name']
            ttag = td.get('template_tag', 'default')
            template, _, tmeta = self._filter_data.get_template(chan, tag=ttag, return_metadata=True)
            pre = None
            for key in ('nb_pretrigger_samples', 'pretrigger_length_samples', 'pretrigger_samples'):
                if key in tmeta:
                    pre = int(tmeta[key])
                    break
            if pre is None:
                raise ValueError('ERROR: Template metadata needs to contain "nb_pretrigger_samples" value')
            csd, _, _ = self._filter_data.get_psd(chan, tag=td.get('csd_tag', 'default'), return_metadata=True)
            peaks = td.get('ignored_frequency_peaks')
            if peaks is not None and not isinstance(peaks, list):
                peaks = [peaks]
            eb.add_trigger_object(trig_chan, OptimumFilterTrigger(
                chan, self._fs, template, csd, pre, ignored_frequency_peaks=peaks,
                ignore_harmonics=td.get('ignore_harmonics', False) if peaks is not None else False,
                trigger_name=trig_chan, precision=self._precision, max_samples=max_samples, device=self._device))
        return eb

    def process(self, ntriggers=-1, lgc_output=True, lgc_save=False, save_path=None, ncores=1, gather=True, **kwargs):
        import pandas as pd
        import torch
        reader = self._reader
        n_events, L = len(reader), int(reader.metadata['nb_samples'])
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        eb = self._build_triggers(L)
        admin = self._admin
        rank, world = dist_info()
        lo, hi = shard_range(n_events, rank, world)
        frames = []
        ntrig = 0
        for ev in range(lo, hi):
            eb.clear_event()
            x = reader.to_amps(reader.upload(ev, ev + 1, dev))[0]   # [n_chan, L] amps on the device
            for trig_chan, td in self._trigger_config.items():
                if 'threshold_sigma' in td:
                    thr = float(td['threshold_sigma'])
                elif 'threshold' in td:
                    thr = float(td['threshold'])
                else:
                    raise ValueError('ERROR: "treshold_sigma" missing in yaml configuration file')
                run_residual = bool(td.get('run_residual', False))      # process/triggers.py:742-751
                sat_amps = [float(a) for a in td['sat_amps_50kHz']] if run_residual and 'sat_amps_50kHz' in td else None
                eb.acquire_triggers(trig_chan, self._channel_trace(x, td['channel_name']), thr,
                                    pileup_window_msec=(float(td['pileup_window_msec']) if 'pileup_window_msec' in td else None),
                                    pileup_window_samples=(int(td['pileup_window_samples']) if 'pileup_window_samples' in td else None),
                                    positive_pulses=td.get('positive_pulses', True),
                                    run_residual=run_residual, sat_amps_50kHz=sat_amps,
                                    edge_exclusion_msec=self._edge_exclusion_msec, livetime=self._livetime)
            if admin is not None:
                info = copy.deepcopy(admin[ev])
            else:      # per-event columns of the reader (event_time, series_num, event_num, dump_num, ...)
                info = {k: (v[0].item() if hasattr(v[0], 'item') else v[0]) for k, v in reader.admin(ev, ev + 1).items()}
                info.setdefault('event_num', info.get('event_number', ev + 1))
            info.setdefault('sample_rate', self._fs)
            info.setdefault('nb_samples', L)
            if self._processing_id is not None:
                info['processing_id'] = self._processing_id
            eb.build_event(info, fs=self._fs,
                           coincident_window_msec=self._evtbuilder_config.get('coincident_window_msec'),
                           coincident_window_samples=self._evtbuilder_config.get('coincident_window_samples'),
                           nb_trigger_channels=len(self._trigger_config))
            df = eb.get_event_df()
            if df is not None and len(df):
                frames.append(df)
                ntrig += len(df)
            if ntriggers is not None and ntriggers > 0 and ntrig >= ntriggers:
                break
        df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
        if ntriggers is not None and ntriggers > 0:
            df = df.iloc[:ntriggers]
        if gather:
            df = gather_frames(df)
        if lgc_save and rank == 0:
            save_path = save_path or '.'
            os.makedirs(save_path, exist_ok=True)
            df.to_parquet(os.path.join(save_path, f'{self._processing_id or "threshtrig"}_F0001.parquet'))
        return df if lgc_output else None
