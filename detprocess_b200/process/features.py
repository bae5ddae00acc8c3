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
from ..utils.utils import columns_to_frame
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
    """All ranks receive the concatenation (rank order == event order) of the per-rank feature tables -- the "gather of
    the small per-event feature tables" of the multi-GPU design.  The numeric columns travel as ONE
    float64 tensor per rank through ``all_gather`` (NCCL over NVLink when the job runs on GPUs; ~16 doubles per event),
    padded to the longest shard; only columns that are not numbers (strings such as data_type) go through the pickled
    object gather.  No trace ever crosses NVLink."""
    import pandas as pd
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return df
    on_gpu = dist.get_backend() == 'nccl'
    dev = torch.device('cuda', torch.cuda.current_device()) if on_gpu else torch.device('cpu')
    numeric = [c for c in df.columns if pd.api.types.is_numeric_dtype(df[c]) or pd.api.types.is_bool_dtype(df[c])]
    other = [c for c in df.columns if c not in numeric]
    # every rank runs the same configuration, hence the same columns; a rank without events has none of them yet
    meta = [None] * world
    dist.all_gather_object(meta, (len(df), list(df.columns), {c: str(df[c].dtype) for c in numeric}))
    ref = next((m for m in meta if m[0] > 0), meta[0])
    columns, dtypes = ref[1], ref[2]
    numeric = [c for c in columns if c in dtypes]
    other = [c for c in columns if c not in dtypes]
    nmax = max(m[0] for m in meta)
    if nmax == 0:
        return df
    block = torch.zeros((nmax, max(len(numeric), 1)), dtype=torch.float64, device=dev)
    if len(df) and numeric:
        block[:len(df), :len(numeric)] = torch.from_numpy(df[numeric].to_numpy(dtype=np.float64, copy=True)).to(dev)
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block)
    host = [p[:m[0], :len(numeric)].cpu().numpy() for p, m in zip(parts, meta)]
    out = pd.DataFrame(np.concatenate(host, axis=0), columns=numeric)
    for c in numeric:                       # integers and booleans come back as what they were
        if not dtypes[c].startswith('float'):
            out[c] = out[c].astype(dtypes[c])
    if other:
        objs = [None] * world
        dist.all_gather_object(objs, df[other] if len(df) else None)
        objs = [o for o in objs if o is not None and len(o)]
        extra = pd.concat(objs, ignore_index=True)
        for c in other:
            out[c] = extra[c].to_numpy()
    return out[columns]


def standard_admin(admin, nb, group_name=None):
    """The reference's fixed admin column set and dtypes (``ProcessingData.get_event_admin``,
    detprocess/process/processing_data.py:811-887) from whatever per-event columns a reader supplies.  Reader names as
    pytesio reports them (event_num, series_num, dump_num, event_index, event_id, event_time, run_type, data_mode,
    trigger_type, trigger_amplitude, trigger_time, fridge_run, fridge_run_start, series_start, group_start) and their
    already-renamed forms are both accepted.  What the reader does not know is derived the way the DAQ defines it
    (event_number = dump_number * 100000 + event_index) or left as NaN like the reference does; columns the reference
    has no name for pass through unchanged after the standard ones."""
    alias = {'event_num': 'event_number', 'series_num': 'series_number', 'dump_num': 'dump_number',
             'fridge_run': 'fridge_run_number', 'fridge_run_start': 'fridge_run_start_time',
             'series_start': 'series_start_time', 'group_start': 'group_start_time'}
    src = {alias.get(k, k): np.asarray(v) for k, v in admin.items()}
    out = {}
    ev = src['event_number'].astype(np.int64) if 'event_number' in src else np.arange(nb, dtype=np.int64)
    out['event_number'] = ev
    out['event_index'] = (src['event_index'] if 'event_index' in src else ev % 100000).astype(np.int32)
    out['dump_number'] = (src['dump_number'] if 'dump_number' in src else ev // 100000).astype(np.int16)
    out['series_number'] = (src['series_number'] if 'series_number' in src else np.zeros(nb)).astype(np.int64)
    out['event_id'] = (src['event_id'] if 'event_id' in src else out['event_index']).astype(np.int32)
    out['event_time'] = (src['event_time'] if 'event_time' in src else np.zeros(nb)).astype(np.int64)
    run_type = src['run_type'].astype(str) if 'run_type' in src else np.full(nb, 'nan')
    out['run_type'] = run_type
    out['data_type'] = run_type.copy()            # the reference fills data_type from run_type (:819)
    out['group_name'] = np.full(nb, group_name if group_name is not None else np.nan, dtype=object if group_name is not None else float)
    if 'trigger_type' in src:
        out['trigger_type'] = src['trigger_type'].astype(np.int16)
    elif 'data_mode' in src:
        modes = ['cont', 'trig-ext', 'rand', 'threshold']
        out['trigger_type'] = np.array([modes.index(m) + 1 if m in modes else np.nan for m in src['data_mode'].astype(str)])
    else:
        out['trigger_type'] = np.full(nb, np.nan)
    for k in ('trigger_amplitude', 'trigger_time'):
        out[k] = src[k].astype(np.float64) if k in src else np.full(nb, np.nan)
    for k in ('fridge_run_number', 'fridge_run_start_time', 'series_start_time', 'group_start_time'):
        out[k] = src[k].astype(np.int64) if k in src else np.full(nb, np.nan)
    known = set(out) | {'data_mode'}
    for k, v in src.items():
        if k not in known:
            out[k] = v
    return out


def _public_algorithms(cls):
    return [m for m in dir(cls) if not m.startswith('_')]


class FeatureProcessing:
    def __init__(self, raw_data, config_file, filter_data=None, external_file=None,
                 processing_id=None, precision='f64', device=None, verbose=True,
                 trigger_dataframe=None, trigger_dataframe_path=None):
        """
        raw_data : an ``EventReader`` or a dict with
            'traces'      ndarray / torch tensor [B, n_chan, N] (float64 amps, or int16 ADC counts + 'adc_gain')
            'channels'    list of channel names (length n_chan)
            'sample_rate' float
            'admin'       optional dict of per-event columns (event_number, series_number, ...)
        config_file : YAML path (or an already-loaded dict)
        filter_data : FilterData with the templates / PSDs the YAML refers to
        trigger_dataframe(_path) : table with one row per trigger (``trigger_index`` + the ``event_number`` [and
            ``series_number``] of the continuous event it was found in, as ``TriggerProcessing`` writes it; reference
            features.py:59, 235-238).  The reader then holds CONTINUOUS events and every feature is computed on the
            window ``[trigger_index - nb_pretrigger_samples, ... + nb_samples)`` of its event, cut inside the kernels'
            loads (reference ProcessingData.read_next_event -> read_single_event(trigger_index, trace_length_samples,
            pretrigger_length_samples), processing_data.py:643-688).
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
        if trigger_dataframe is None and trigger_dataframe_path is not None:
            import pandas as pd
            trigger_dataframe = pd.read_parquet(trigger_dataframe_path)
        self._triggers = trigger_dataframe
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
        # int16 ADC events stay int16 on the device: the fused kernels (OF, window reductions, channel algebra) convert in
        # their loads with the reader's per-channel gain / offset
        meta = self._reader.metadata
        self._adc = None
        if meta.get('dtype') == 'int16' and 'adc_gain' in meta:
            self._adc = {c: (float(meta['adc_gain'][i]), float(meta['adc_offset'][i])) for i, c in enumerate(self._channels)}
        self._instantiate_of_bases()
        self._compile_jobs()

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

    def _is_plain(self, channel):
        return utils.split_channel_name(channel, available_channels=self._channels)[1] is None

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
                if base == 'psd_amp':
                    continue        # no template / csd (reference processing_data.py:288-289): a band job of _compile_jobs
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

    def _compile_jobs(self):
        """Everything the YAML implies is resolved ONCE: per (channel, algorithm) the extractor, its kwargs (windows as
        indices) and the kind of job; every OF fit is registered with its OFBaseBatch before the first batch (one plan, one
        table build, one launch per batch whatever the blocks' windows and ``lowchi2_fcutoff``); the window reductions
        share cached ReducePlans; the combined channels ('a+b', 'a-b') are listed for the channel-algebra kernel."""
        self._of_jobs, self._ext_jobs = [], []
        self._band_jobs = []    # psd_amp blocks: dicts (channel, columns, bin ranges, nb_samples), one BandPlan each
        red_jobs = {}           # (nb_samples, nb_pretrigger) -> list of (channel, op, lo, hi, column)
        self._combined = []     # combined channel names, in the order of the combine kernel's output rows
        n_reader = int(self._reader.metadata['nb_samples'])
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
                kw.setdefault('nb_samples', n_reader)
                kw.setdefault('nb_pretrigger_samples', kw['nb_samples'] // 2)
                wmin, wmax = utils.get_window_indices(**kw)
                kw['window_min_index'], kw['window_max_index'] = wmin, wmax
                kw['feature_base_name'] = algorithm
                geom = (int(kw['nb_samples']), int(kw['nb_pretrigger_samples']))
                sep = utils.split_channel_name(channel, available_channels=self._channels)[1]
                if sep in ('+', '-') and channel not in self._combined:
                    self._combined.append(channel)
                entry = self._of_bases.get(self._of_key(params)) if (base.startswith('of1x1') or base == 'ofnxm') else None
                if entry is not None and algorithm in entry['algorithms']:
                    self._of_jobs.append((channel, extractor, entry, kw, feature_channel))
                    spec = FE._of_fit_spec(base, channel, entry['OF'], **{k: v for k, v in kw.items() if k != 'base_algorithm'})
                    if spec is not None:
                        entry['OF'].request_fit(channel, *spec)
                elif base in _TRACE_OPS and extractor is getattr(FE, base):
                    red_jobs.setdefault(geom, []).append((channel, base, wmin, wmax, f'{algorithm}_{feature_channel}'))
                elif base == 'psd_amp' and extractor is FE.psd_amp:
                    if not self._is_plain(channel):
                        raise NotImplementedError(f'psd_amp on the combined / joint channel "{channel}" is outside the built hot path')
                    names, bins = FE._psd_amp_ranges(geom[0], self._fs, params.get('f_lims', []))
                    self._band_jobs.append({'channel': channel, 'bins': bins, 'nb_samples': geom[0], 'plan': None,
                                            'columns': [f'{algorithm}_{nm}_{feature_channel}' for nm in names]})
                else:
                    self._ext_jobs.append((channel, extractor, kw, feature_channel, geom))
        # cached reduction plans: per trace geometry one plan over the plain channels (read in place from the reader
        # batch, ADC conversion in the load) and one over the combined channels (output of the channel-algebra kernel)
        self._red_plans = {}
        for geom, jobs in red_jobs.items():
            plans = {}
            for kind in ('plain', 'comb'):
                sel = [j for j in jobs if (self._is_plain(j[0]) if kind == 'plain' else j[0] in self._combined)]
                bad = [j[0] for j in jobs if not self._is_plain(j[0]) and j[0] not in self._combined]
                if bad:
                    raise NotImplementedError(f'window features on channel(s) {bad} are outside the built hot path')
                if not sel:
                    continue
                chans = utils.unique_list([j[0] for j in sel])
                plan = ReducePlan(geom[0], self._fs, len(chans))
                if kind == 'plain' and self._adc is not None:
                    for i, c in enumerate(chans):
                        plan.set_adc_conversion(i, *self._adc[c])
                handles = [plan.add(chans.index(c), op, a, b) for c, op, a, b, _ in sel]
                plans[kind] = {'plan': plan, 'chans': chans, 'handles': handles, 'columns': [j[4] for j in sel]}
            self._red_plans[geom] = plans
        # combine-kernel terms of the combined channels
        self._combine_terms = []
        for channel in self._combined:
            parts, sep = utils.split_channel_name(channel, available_channels=self._channels)
            rows = [self._channels.index(c) for c in parts]
            w = None
            if channel in self._weights:
                wd = self._weights[channel]
                w = []
                for c in parts:
                    if f'weight_{c}' not in wd:
                        raise ValueError(f'ERROR: Missing parameter weight weight_{c} for channel {channel}!')
                    w.append(float(wd[f'weight_{c}']))
            if sep == '-' and len(rows) != 2:
                raise ValueError(f'ERROR: "{channel}": a difference takes two channels')
            self._combine_terms.append([(r, None if w is None else w[i], -1.0 if (sep == '-' and i == 1) else 1.0)
                                        for i, r in enumerate(rows)])

    # ------------------------------------------------------------------ traces
    def _channel_trace(self, traces, channel):
        """Weighted channel algebra of ProcessingData.get_channel_trace (reference :941-1049) on a float64 batch (used for
        joint 'a|b' channels and external extractors; plain and 'a+b' / 'a-b' channels never come here)."""
        parts, sep = utils.split_channel_name(channel, available_channels=self._channels)
        idx = [self._channels.index(c) for c in parts]
        if sep is None:
            return traces[:, idx[0], :]
        if sep == '|':
            return traces[:, idx, :]            # [B, n, N] in the listed order (NxM filter)
        raise NotImplementedError(f'channel operator "{sep}" is outside the built hot path')

    # ------------------------------------------------------------------ process
    def process(self, nevents=-1, lgc_save=False, lgc_output=True, save_path=None, ncores=1,
                batch_size=8192, gather=True, memory_limit=2.0, **kwargs):
        """Events are read batch by batch through the reader (int16 ADC counts cross PCIe as they are stored and become
        amps inside the kernels), sharded over ranks in contiguous blocks.  lgc_save: every rank writes its own dumps
        ``<prefix>_rank<r>_F000k.parquet`` when ``memory_limit`` (GB) of rows is queued (reference features.py:584-629:
        one file set per worker, no merge); lgc_output: the gathered table is returned.  With a trigger dataframe the
        reader's events are continuous streams and the rows of the table are the events (see ``__init__``).

        Three batches are in flight: while batch k's kernels run, batch k+1 crosses PCIe into the pipeline's own device
        staging (three reused buffers: no allocator call per batch) and the host turns the results of batch k-1 -- copied
        to pinned memory behind their kernels -- into table columns.  Columns stay numpy arrays until the table (or a
        dump) is assembled: one DataFrame construction per table, not per batch."""
        import torch
        rank, world = dist_info()
        writer = None
        if lgc_save:
            from ..io.writers import FeatureWriter
            writer = FeatureWriter(save_path or '.', prefix=self._processing_id or 'feature',
                                   series_name=(f'rank{rank}' if world > 1 else None), memory_limit_gb=memory_limit)
        parts = []
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)

        def emit(cols):
            if writer is not None:
                writer.add(cols)
            if lgc_output:
                parts.append(cols)

        if self._triggers is not None:
            self._process_triggers(nevents, dev, rank, world, emit)
        else:
            reader = self._reader
            nev_total = len(reader) if nevents is None or nevents < 0 else min(nevents, len(reader))
            lo, hi = shard_range(nev_total, rank, world)      # contiguous blocks, one process per GPU
            copy_stream = torch.cuda.Stream(dev)
            main = torch.cuda.current_stream(dev)
            meta = reader.metadata
            shape = (min(batch_size, max(hi - lo, 1)), len(meta['channels']), int(meta['nb_samples']))
            key = (shape, meta['dtype'], str(dev))
            if getattr(self, '_stage_key', None) != key:
                self._stage = [torch.empty(shape, dtype=getattr(torch, meta['dtype']), device=dev) for _ in range(3)]
                self._stage_free = [None] * 3       # event on the compute stream after the last kernel that read the buffer
                self._stage_key = key
            slot = 0

            def fetch(b0):
                nonlocal slot
                b1 = min(b0 + batch_size, hi)
                k, slot = slot, (slot + 1) % 3
                if self._stage_free[k] is not None:
                    copy_stream.wait_event(self._stage_free[k])
                t = reader.upload(b0, b1, dev, stream=copy_stream, out=self._stage[k])
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                return t, ev, b0, b1, k

            nxt = fetch(lo) if lo < hi else None
            pending = None
            while nxt is not None:
                t, ev, b0, b1, k = nxt
                nxt = fetch(b1) if b1 < hi else None
                main.wait_event(ev)
                ticket = self._launch_batch(t, b0, b1)
                done = torch.cuda.Event()
                done.record(main)
                self._stage_free[k] = done
                if pending is not None:
                    emit(self._collect(pending))
                pending = ticket
            if pending is not None:
                emit(self._collect(pending))
        self.output_files = writer.close() if writer is not None else []
        if not lgc_output:
            return None
        df = columns_to_frame(parts)
        return gather_frames(df) if gather else df

    def _process_triggers(self, nevents, dev, rank, world, emit):
        """trigger-dataframe mode: rows are grouped by the continuous event they point into; each stream is uploaded once
        and all of its windows are processed by one launch per plan.  Streams (not rows) are sharded over the ranks."""
        import torch
        trig = self._triggers
        if nevents is not None and nevents >= 0:
            trig = trig.iloc[:nevents]
        if 'trigger_index' not in trig.columns:
            raise ValueError('ERROR: the trigger dataframe needs a "trigger_index" column')
        reader = self._reader
        alias = {'event_num': 'event_number', 'series_num': 'series_number', 'dump_num': 'dump_number'}
        adm = {alias.get(k, k): np.asarray(v) for k, v in reader.admin(0, len(reader)).items()}
        keys = [k for k in ('series_number', 'event_number') if k in adm and k in trig.columns]
        if 'event_number' not in keys:
            raise ValueError('ERROR: trigger dataframe and reader need an "event_number" column to find the continuous event')
        index_of = {tuple(int(adm[k][i]) for k in keys): i for i in range(len(reader))}
        trig_keys = np.stack([trig[k].to_numpy(dtype=np.int64) for k in keys], axis=1)
        stream_of_row = np.array([index_of.get(tuple(r), -1) for r in trig_keys.tolist()], dtype=np.int64)
        if (stream_of_row < 0).any():
            raise ValueError('ERROR: trigger dataframe rows point to events the raw data does not hold')
        streams = np.unique(stream_of_row)
        lo, hi = shard_range(len(streams), rank, world)
        tindex = trig['trigger_index'].to_numpy(dtype=np.int64)
        for si in streams[lo:hi]:
            rows = np.nonzero(stream_of_row == si)[0]
            batch = reader.upload(int(si), int(si) + 1, dev)[0]              # [n_chan, L] as stored
            tidx = torch.from_numpy(tindex[rows]).to(dev)
            cols = self._collect(self._launch_batch(batch, None, None, trigger_index=tidx))
            tab = trig.iloc[rows]                                             # the trigger rows verbatim (:795-804)
            out = {c: tab[c].to_numpy() for c in tab.columns}
            out.update(cols)
            emit(out)

    def _process_batch(self, batch, ev0, ev1, trigger_index=None):
        """one batch, start to finish -> DataFrame (see ``_launch_batch`` for the arguments)"""
        return columns_to_frame([self._collect(self._launch_batch(batch, ev0, ev1, trigger_index=trigger_index))])

    def _result_staging(self, n_doubles):
        """pinned host block for the results of one batch (three rotate: a block is reused two batches later, after its
        columns were copied out by ``_collect``)"""
        import torch
        st = getattr(self, '_res_stage', None)
        if st is None:
            st = self._res_stage = {'bufs': [None] * 3, 'k': 0}
        k = st['k']
        st['k'] = (k + 1) % 3
        buf = st['bufs'][k]
        if buf is None or buf.numel() < n_doubles:
            buf = st['bufs'][k] = torch.empty(max(n_doubles, 1), dtype=torch.float64).pin_memory()
        return buf

    def _launch_batch(self, batch, ev0, ev1, trigger_index=None):
        """Queue everything batch-shaped on the device and return a ticket for ``_collect``.
        batch mode: ``batch`` [B, n_chan, N] events ev0..ev1 of the reader; window mode: ``batch`` [n_chan, L] one
        continuous event and ``trigger_index`` int64 [B] (device)."""
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        batch = batch.to(dev, non_blocking=True)
        window_mode = trigger_index is not None
        rows = {c: i for i, c in enumerate(self._channels)}
        cols = {}
        if window_mode:
            nb, n = int(trigger_index.shape[0]), None
            if self._combined or self._ext_jobs or self._band_jobs:
                raise NotImplementedError('combined channels / external extractors / psd_amp in trigger-dataframe mode are not built')
        else:
            nb, _, n = batch.shape
        if self._processing_id is not None:
            cols['processing_id'] = np.full(nb, self._processing_id)
        if not window_mode:
            # admin columns (reference get_event_admin :811-887) and detector settings (get_channel_settings :891-937)
            cols.update(standard_admin(self._reader.admin(ev0, ev1), nb, group_name=self._reader.metadata.get('group_name')))
            det = self._reader.metadata.get('detector_config')
            if det:
                for channel in self._processing_config:
                    for c in utils.split_channel_name(channel, available_channels=self._channels)[0]:
                        if c in det:
                            cols[f'tes_bias_{c}'] = np.full(nb, det[c].get('tes_bias', np.nan))
                            cols[f'output_gain_{c}'] = np.full(nb, det[c].get('output_gain', np.nan))

        # ---- combined channels 'a+b' / 'a-b': one channel-algebra launch straight from the reader's buffer ---------
        comb = None
        if self._combined:
            from ..core.plans import combine_channels
            adc_rows = None if self._adc is None else {rows[c]: v for c, v in self._adc.items()}
            comb = combine_channels(batch, self._combine_terms, adc=adc_rows)
        amps = None          # float64 amps of the whole batch: only joint channels / external extractors need them

        def float_batch():
            nonlocal amps
            if amps is None:
                amps = self._reader.to_amps(batch)
            return amps

        # ---- hand the batch to every OF base (reference update_signal_OF, :712-772): plain channels are read where the
        # reader put them, combined ones from the channel-algebra output; the fused kernel of every base is queued
        blocks = []          # (kind, object, device result block)
        for key, entry in self._of_bases.items():
            ofb = entry['OF']
            ofb.clear_signal()
            if not window_mode and key[0] != n:
                raise ValueError(f'ERROR: trace length {n} != configured nb_samples {key[0]}: continuous data need a '
                                 'trigger dataframe (trigger_dataframe=...)')
            plain = {c: rows[c] for c in entry['channels'] if c in rows}
            starts = None if not window_mode else trigger_index - int(key[1])
            if plain:
                ofb.update_batch(batch, plain, start_index=starts)
            for chan in entry['channels']:
                if chan in plain:
                    continue
                if chan in self._combined:
                    ofb.update_signal(chan, comb[:, self._combined.index(chan), :], calc_fft=True)
                else:
                    ofb.update_signal(chan, self._channel_trace(float_batch(), chan), calc_fft=True)
            if ofb.has_of1x1:
                blocks.append(('of', ofb, ofb.launch()))
        # joint-channel (NxM) fits run when their extractor asks (one launch per fit): evaluated now, while the signals
        # of THIS batch are stored
        eager = {}
        for j, (channel, extractor, entry, kw, feature_channel) in enumerate(self._of_jobs):
            if '|' in channel:
                eager[j] = extractor(channel, entry['OF'], **kw)

        # ---- trace-window features: cached plans, inputs consumed in place --------------------------------------------
        for geom, plans in self._red_plans.items():
            if not window_mode and geom[0] != n:
                raise ValueError(f'ERROR: trace length {n} != configured nb_samples {geom[0]}')
            for kind, d in plans.items():
                plan = d['plan']
                if not plan.finalized:
                    plan.finalize(dev)
                if kind == 'plain':
                    starts = None if not window_mode else trigger_index - int(geom[1])
                    out = plan.run_layout(batch, [rows[c] for c in d['chans']], starts)
                else:
                    out = plan.run_layout(comb, [self._combined.index(c) for c in d['chans']])
                blocks.append(('red', d, out))
        # ---- psd_amp: band amplitudes of the plain channels, read where the reader put them --------------------------
        for job in self._band_jobs:
            if job['nb_samples'] != n:
                raise ValueError(f'ERROR: trace length {n} != configured nb_samples {job["nb_samples"]}')
            if job['plan'] is None:
                from ..core.plans import BandPlan
                job['plan'] = BandPlan(n, self._fs, job['bins'], device=dev)
            adc = self._adc.get(job['channel']) if (self._adc is not None and batch.dtype == torch.int16) else None
            blocks.append(('band', job, job['plan'].run(batch[:, rows[job['channel']], :], adc=adc)))
        # ---- all result blocks of the batch -> one pinned host block, behind the kernels
        total = sum(int(b[2].numel()) for b in blocks)
        host = self._result_staging(total)
        views, o = [], 0
        for _, _, t in blocks:
            v = host[o:o + t.numel()].view(t.shape)
            v.copy_(t, non_blocking=True)
            views.append(v)
            o += t.numel()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        # user-supplied extractors keep the reference's per-event numpy calling convention
        ext = {}
        for channel, extractor, kw, feature_channel, geom in self._ext_jobs:
            if channel in self._combined:
                tr = comb[:, self._combined.index(channel), :].cpu().numpy()
            else:
                tr = self._channel_trace(float_batch(), channel).cpu().numpy()
            rows_out = [extractor(tr[i], **kw) for i in range(nb)]
            for name in rows_out[0]:
                ext[f'{name}_{feature_channel}'] = np.array([r[name] for r in rows_out])
        return {'cols': cols, 'blocks': blocks, 'views': views, 'ready': ready, 'eager': eager, 'ext': ext, 'keep': (batch, comb)}

    def _collect(self, ticket):
        """wait for the batch's results and turn them into table columns (dict name -> ndarray [B])"""
        ticket['ready'].synchronize()
        cols = ticket['cols']
        reds = []
        for (kind, obj, _), v in zip(ticket['blocks'], ticket['views']):
            arr = v.numpy().copy()          # the pinned block is reused two batches later
            if kind == 'of':
                obj.set_results(arr)
            elif kind == 'band':
                for i, name in enumerate(obj['columns']):
                    cols[name] = arr[:, i]
            else:
                reds.append((obj, arr))
        for j, (channel, extractor, entry, kw, feature_channel) in enumerate(self._of_jobs):
            res = ticket['eager'][j] if j in ticket['eager'] else extractor(channel, entry['OF'], **kw)
            for name, val in res.items():
                cols[f'{name}_{feature_channel}'] = val
        for d, arr in reds:
            plan = d['plan']
            for h, name in zip(d['handles'], d['columns']):
                cols[name] = arr[:, plan.column(h)]
        cols.update(ticket['ext'])
        ticket['keep'] = None
        return cols
