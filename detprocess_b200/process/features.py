"""
``FeatureProcessing``: YAML-driven batched feature extraction, the B200 counterpart of the
reference's ``detprocess/process/features.py`` event loop (:533-851).

The reference reads ONE event, pushes it through every ``qp.OFBase`` and calls one
extractor per (channel, algorithm), building a dict row.  Here a batch of B events
``[B, n_chan, N]`` is processed at once: per OF key ``(nb_samples, nb_pretrigger, tag)``
one ``OFBaseBatch`` (one fused kernel launch for all its fits), one ``ReducePlan`` launch
for all trace-window features; the extractors of ``FeatureExtractors`` return whole
columns.  Column naming, window arithmetic, key/tag construction and the resolution of
``base_algorithm`` follow the reference (:729-846, :1243-1344).

Input is an in-memory batch source (no HDF5 library exists in this image; raw pytesdaq
reading is the "next" row of SURVEY.md 8(f)).
"""
import os
import importlib.util

import numpy as np

from ..core.algorithms import FeatureExtractors as FE
from ..core.ofbase import OFBaseBatch
from ..core.plans import ReducePlan
from ..utils import utils
from .config import YamlConfig

__all__ = ['FeatureProcessing']

_OF_PREFIXES = ('of1x', 'ofnx', 'psd_amp', 'psd_peaks', 'phase')
_TRACE_OPS = ('baseline', 'integral', 'maximum', 'minimum')


def dist_info():
    """(rank, world) of the torch.distributed job, (0, 1) when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_events, rank, world):
    """Contiguous block [lo, hi) of events owned by `rank` (rank r gets [r*B/G, (r+1)*B/G);
    SURVEY.md 8(e)).  Blocks are disjoint, ordered and cover all events."""
    return rank * n_events // world, (rank + 1) * n_events // world


def gather_frames(df):
    """All ranks receive the concatenation (rank order == event order) of the per-rank
    feature tables.  The tables are small (~16 doubles / event); no trace ever crosses NVLink."""
    import pandas as pd
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return df
    parts = [None] * world
    dist.all_gather_object(parts, df)
    parts = [p for p in parts if len(p)]
    return pd.concat(parts, ignore_index=True) if parts else df


def _public_algorithms(cls):
    return [m for m in dir(cls) if not m.startswith('_')]


class FeatureProcessing:
    def __init__(self, raw_data, config_file, filter_data=None, external_file=None,
                 processing_id=None, precision='f64', device=None, verbose=True):
        """
        raw_data : dict with
            'traces'      ndarray / torch tensor [B, n_chan, N] (float64 amps)
            'channels'    list of channel names (length n_chan)
            'sample_rate' float
            'admin'       optional dict of per-event columns (event_number, series_number, ...)
        config_file : YAML path (or an already-loaded dict)
        filter_data : FilterData with the templates / PSDs the YAML refers to
        """
        self._verbose = verbose
        from ..io.readers import EventReader, ArrayReader
        if isinstance(raw_data, EventReader):
            self._reader = raw_data
        else:
            self._reader = ArrayReader(raw_data['traces'], raw_data['channels'], raw_data['sample_rate'],
                                       admin=raw_data.get('admin'), adc_gain=raw_data.get('adc_gain'),
                                       adc_offset=raw_data.get('adc_offset'))
        self._channels = self._reader.channels
        self._fs = self._reader.sample_rate
        self._filter_data = filter_data
        self._precision = precision
        self._device = device
        self._processing_id = processing_id
        cfg = YamlConfig(config_file, self._channels, sample_rate=self._fs, verbose=verbose)
        fcfg = cfg.get_config('feature')
        self._processing_config = fcfg['channels']
        self._weights = fcfg['weights']
        self._traces_config = fcfg['traces_config']
        self._algorithm_list = _public_algorithms(FE)
        self._ext = None
        self._ext_algorithm_list = []
        if external_file is not None:
            self._ext = self._load_external_extractors(external_file)
            self._ext_algorithm_list = _public_algorithms(self._ext)
            dup = set(self._ext_algorithm_list) & set(self._algorithm_list)
            if dup:
                raise ValueError(f'ERROR: External feature extractor(s) {sorted(dup)} duplicate internal names')
        self._of_bases = {}     # key_tuple -> {'OF': OFBaseBatch, 'channels': [...], 'algorithms': [...]}
        # int16 ADC events stay int16 on the device (the fused kernels convert in their loads) when every configured
        # channel is a plain channel and every algorithm is built in; channel algebra / NxM / external extractors take
        # the float64 path (ADC -> amps by a torch kernel first)
        meta = self._reader.metadata
        self._adc = None
        if meta.get('dtype') == 'int16' and 'adc_gain' in meta:
            plain = all(utils.split_channel_name(c, available_channels=self._channels)[1] is None
                        for c, cc in self._processing_config.items() if isinstance(cc, dict))
            builtin = all(params.get('base_algorithm', algo) in self._algorithm_list
                          for cc in self._processing_config.values() if isinstance(cc, dict)
                          for algo, params in cc.items() if isinstance(params, dict) and params.get('run'))
            if plain and builtin:
                self._adc = {c: (float(meta['adc_gain'][i]), float(meta['adc_offset'][i])) for i, c in enumerate(self._channels)}
        self._instantiate_of_bases()

    # ------------------------------------------------------------------ setup
    @staticmethod
    def _load_external_extractors(path):
        if not os.path.isfile(path):
            raise ValueError(f'ERROR: External feature extractors file "{path}" not found!')
        spec = importlib.util.spec_from_file_location('detprocess_b200_external', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if not hasattr(mod, 'FeatureExtractors'):
            raise ValueError('ERROR: external file must define a class named "FeatureExtractors"')
        return mod.FeatureExtractors

    @staticmethod
    def _of_key(params):
        """(nb_samples, nb_pretrigger, '<csd_tag>_<coupling>[_freqs[_harmonics]]') -- reference
        processing_data.py:239-275 == features.py:794-820"""
        tag = params.get('csd_tag', 'default')
        tag = f'{tag}_{params.get("coupling", "AC")}'
        if 'ignored_frequency_peaks' in params:
            freqs = params['ignored_frequency_peaks']
            freqs = freqs if isinstance(freqs, list) else [freqs]
            s = '_'.join(map(str, freqs))
            if params.get('ignore_harmonics'):
                s += '_harmonics'
            tag = f'{tag}_{s}'
        return (params['nb_samples'], params['nb_pretrigger_samples'], tag)

    def _instantiate_of_bases(self):
        if self._filter_data is None:
            return
        for chan, chan_config in self._processing_config.items():
            if not isinstance(chan_config, dict):
                continue
            for algo, params in chan_config.items():
                if not isinstance(params, dict) or not params['run']:
                    continue
                base = params.get('base_algorithm', algo)
                if not any(p in base for p in _OF_PREFIXES):
                    continue
                if not (base.startswith('of1x1') or base == 'ofnxm'):
                    raise NotImplementedError(f'algorithm "{base}" is outside the built hot path')
                key = self._of_key(params)
                entry = self._of_bases.setdefault(key, {
                    'OF': OFBaseBatch(self._fs, precision=self._precision, device=self._device),
                    'channels': [], 'algorithms': []})
                entry['algorithms'].append(algo)
                if chan not in entry['channels']:
                    entry['channels'].append(chan)
                ofb = entry['OF']
                if self._adc is not None and chan in self._adc:
                    ofb.set_adc_conversion(chan, *self._adc[chan])
                csd_tag = params.get('csd_tag', 'default')
                if base == 'ofnxm':
                    psd, _, meta = self._filter_data.get_csd(chan, tag=csd_tag, return_metadata=True)
                else:
                    psd, _, meta = self._filter_data.get_psd(chan, tag=csd_tag, return_metadata=True)
                if meta.get('sample_rate', self._fs) != self._fs:
                    raise ValueError(f'Sample rate is not consistent between raw data and csd for channel {chan}!')
                if params['nb_samples'] != psd.shape[-1]:
                    raise ValueError(f'Number of samples is not consistent between raw data (={params["nb_samples"]}) '
                                     f'and csd (={psd.shape[-1]})for channel {chan}, algorithm {algo}!')
                if ofb.csd(chan) is None:
                    peaks = params.get('ignored_frequency_peaks')
                    if peaks is not None and not isinstance(peaks, list):
                        peaks = [peaks]
                    ofb.set_csd(chan, psd, coupling=params.get('coupling', 'AC'),
                                ignored_frequency_peaks=peaks,
                                ignore_harmonics=params.get('ignore_harmonics', False))
                if 'template_tag' not in params:
                    raise ValueError(f'ERROR: a "template_tag" in yaml file is required for channel {chan}, '
                                     f'algorithm "{algo}" !')
                ttag = params['template_tag']
                tmpl, _, tmeta = self._filter_data.get_template(chan, tag=ttag, return_metadata=True)
                if params['nb_samples'] != tmpl.shape[-1]:
                    raise ValueError(f'Number of samples is not consistent between raw data and template '
                                     f'("{ttag}") for channel {chan}, algorithm {algo}!')
                # the template's own pretrigger metadata wins over the YAML (reference :360-366)
                pre = int(tmeta.get('nb_pretrigger_samples', params['nb_pretrigger_samples']))
                ofb.add_template(chan, tmpl, template_tag=ttag, pretrigger_samples=pre,
                                 integralnorm=params.get('integralnorm', False), overwrite=True)

    # ------------------------------------------------------------------ traces
    def _channel_trace(self, traces, channel):
        """Weighted channel algebra of ProcessingData.get_channel_trace (reference :941-1049)."""
        parts, sep = utils.split_channel_name(channel, available_channels=self._channels)
        idx = [self._channels.index(c) for c in parts]
        w = None
        if channel in self._weights:
            wd = self._weights[channel]
            w = []
            for c in parts:
                if f'weight_{c}' not in wd:
                    raise ValueError(f'ERROR: Missing parameter weight weight_{c} for channel {channel}!')
                w.append(float(wd[f'weight_{c}']))
        if sep == '+':
            cols = [traces[:, i, :] * w[j] if w is not None else traces[:, i, :] for j, i in enumerate(idx)]
            out = cols[0]
            for c in cols[1:]:
                out = out + c
            return out
        if sep == '-':
            if w is not None:
                return traces[:, idx[0], :] * w[0] - traces[:, idx[1], :] * w[1]
            return traces[:, idx[0], :] - traces[:, idx[1], :]
        if sep is None:
            return traces[:, idx[0], :]
        if sep == '|':
            return traces[:, idx, :]            # [B, n, N] in the listed order (NxM filter)
        raise NotImplementedError(f'channel operator "{sep}" is outside the built hot path')

    # ------------------------------------------------------------------ process
    def process(self, nevents=-1, lgc_save=False, lgc_output=True, save_path=None, ncores=1,
                batch_size=8192, gather=True, memory_limit=2.0, **kwargs):
        """Events are read batch by batch through the reader (int16 ADC counts cross PCIe as they are stored and become
        amps on the device), sharded over ranks in contiguous blocks.  lgc_save: every rank writes its own dumps
        ``<prefix>_rank<r>_F000k.parquet`` when ``memory_limit`` (GB) of rows is queued (reference features.py:584-629:
        one file set per worker, no merge); lgc_output: the gathered table is returned."""
        import pandas as pd
        reader = self._reader
        nev_total = len(reader) if nevents is None or nevents < 0 else min(nevents, len(reader))
        # ---- shard events over ranks (one process per GPU), contiguous blocks -----------------
        rank, world = dist_info()
        lo, hi = shard_range(nev_total, rank, world)
        writer = None
        if lgc_save:
            from ..io.writers import FeatureWriter
            writer = FeatureWriter(save_path or '.', prefix=self._processing_id or 'feature',
                                   series_name=(f'rank{rank}' if world > 1 else None), memory_limit_gb=memory_limit)
        frames = []
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        copy_stream = torch.cuda.Stream(dev)

        def fetch(b0):
            # the next batch is uploaded on a side stream while the current one is processed
            b1 = min(b0 + batch_size, hi)
            t = reader.upload(b0, b1, dev, stream=copy_stream)   # the reader orders its staging reuse after this copy
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            return t, ev, b0, b1

        nxt = fetch(lo) if lo < hi else None
        while nxt is not None:
            t, ev, b0, b1 = nxt
            nxt = fetch(b1) if b1 < hi else None
            torch.cuda.current_stream(dev).wait_event(ev)
            t.record_stream(torch.cuda.current_stream(dev))
            df = self._process_batch(t, b0, b1)
            if writer is not None:
                writer.add(df)
            if lgc_output:
                frames.append(df)
        self.output_files = writer.close() if writer is not None else []
        if not lgc_output:
            return None
        df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
        return gather_frames(df) if gather else df

    def _process_batch(self, traces, ev0, ev1):
        import pandas as pd
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        traces = traces.to(dev, non_blocking=True)
        if self._adc is None:
            traces = self._reader.to_amps(traces)                             # ADC -> amps on the device (torch)
        nb, _, n = traces.shape
        cols = {k: np.asarray(v) for k, v in self._reader.admin(ev0, ev1).items()}
        if self._processing_id is not None:
            cols['processing_id'] = np.full(nb, self._processing_id)

        # ---- push the batch into every OF base (reference update_signal_OF, :712-772) --------
        for key, entry in self._of_bases.items():
            ofb = entry['OF']
            ofb.clear_signal()
            if key[0] != n:
                raise ValueError(f'ERROR: trace length {n} != configured nb_samples {key[0]} '
                                 '(window extraction from continuous data is not built)')
            for chan in entry['channels']:
                ofb.update_signal(chan, self._channel_trace(traces, chan), calc_fft=True)

        # ---- trace-window features: ONE reduction launch for all of them -----------------------
        red = ReducePlan(n, self._fs, 1)
        red_jobs = []          # (handle, feature column name, channel)
        red_inputs = {}        # channel -> row index in the stacked input
        external_jobs = []
        for channel, algorithms in self._processing_config.items():
            if not isinstance(algorithms, dict):
                continue
            feature_channel = algorithms.get('feature_channel', channel)
            for algorithm, params in algorithms.items():
                if not isinstance(params, dict) or not params['run']:
                    continue
                base = params.get('base_algorithm', algorithm)
                if base in self._algorithm_list:
                    extractor = getattr(FE, base)
                elif base in self._ext_algorithm_list:
                    extractor = getattr(self._ext, base)
                else:
                    raise ValueError(f'ERROR: Cannot find algorithm "{base}" anywhere. '
                                     f'Check feature extractor exists!')
                kw = {k: v for k, v in params.items() if k != 'run'}
                kw['fs'] = self._fs
                kw.setdefault('nb_samples', n)
                kw.setdefault('nb_pretrigger_samples', n // 2)
                wmin, wmax = utils.get_window_indices(**kw)
                kw['window_min_index'], kw['window_max_index'] = wmin, wmax
                kw['feature_base_name'] = algorithm
                entry = self._of_bases.get(self._of_key(params)) if (base.startswith('of1x1') or base == 'ofnxm') else None
                if entry is not None and algorithm in entry['algorithms']:
                    feats = extractor(channel, entry['OF'], **kw)
                    for name, val in feats.items():
                        cols[f'{name}_{feature_channel}'] = val
                elif base in _TRACE_OPS and extractor is getattr(FE, base):
                    red_jobs.append((channel, base, wmin, wmax, f'{algorithm}_{feature_channel}'))
                else:
                    external_jobs.append((channel, extractor, kw, feature_channel))
        if red_jobs:
            chans = utils.unique_list([j[0] for j in red_jobs])
            red = ReducePlan(n, self._fs, len(chans))
            if self._adc is not None:
                for i, c in enumerate(chans):
                    red.set_adc_conversion(i, *self._adc[c])
            handles = [red.add(chans.index(c), op, a, b) for c, op, a, b, _ in red_jobs]
            red.finalize(dev)
            x = torch.stack([self._channel_trace(traces, c) for c in chans], dim=1).contiguous()
            out = red.run(x).cpu().numpy()
            for h, job in zip(handles, red_jobs):
                cols[job[4]] = out[:, red.column(h)]
        # user-supplied extractors keep the reference's per-event numpy calling convention
        for channel, extractor, kw, feature_channel in external_jobs:
            tr = self._channel_trace(traces, channel).cpu().numpy()
            rows = [extractor(tr[i], **kw) for i in range(nb)]
            for name in rows[0]:
                cols[f'{name}_{feature_channel}'] = np.array([r[name] for r in rows])
        return pd.DataFrame(cols)
