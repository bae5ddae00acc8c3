"""
Raw-event readers behind ONE interface.

The reference reads events through pytesio's ``H5Reader`` -- ``read_single_event(event_index, trigger_index=...,
trace_length_samples=..., pretrigger_length_samples=..., adctoamp=True)`` per event
(``detprocess/process/processing_data.py:643-688``) and ``read_many_events(output_format=2, ...)`` for the noise randoms
(``core/noise.py:671-682``).  pytesio and HDF5 are not available in this image, so the pipeline talks to the small
``EventReader`` interface below; a pytesio-backed reader is one more subclass.  What the interface guarantees is what the
device path needs: events arrive in batches ``[B, n_chan, N]`` as the samples are stored (int16 ADC counts when the
file holds them -- a quarter of the PCIe bytes of float64 amps) in pinned host memory, with the per-channel linear
ADC -> amps conversion in the metadata so that it can run on the device.

``RawBinaryReader`` is a concrete, dependency-free container for that: ``<name>.json`` (metadata + per-event admin
columns) next to ``<name>.bin`` (the samples, C order ``[n_events, n_chan, nb_samples]``, memory mapped).
"""
import json
import os

import numpy as np

__all__ = ['EventReader', 'ArrayReader', 'RawBinaryReader', 'write_raw_binary']


class EventReader:
    """Interface.  ``metadata``: dict with 'sample_rate', 'channels', 'nb_samples', 'dtype' ('int16' | 'float32' |
    'float64') and, for ADC data, 'adc_gain' / 'adc_offset' lists (sample = adc * gain + offset, per channel)."""

    metadata = None

    def __len__(self):
        raise NotImplementedError

    def read_batch(self, i0, i1, pinned=True):
        """events [i0, i1) as stored: torch tensor [B, n_chan, N] on the host (pinned when requested and possible)."""
        raise NotImplementedError

    def upload(self, i0, i1, device, stream=None, out=None):
        """events [i0, i1) on ``device`` (asynchronous copy from pinned memory on ``stream`` / the current stream).
        This, not ``read_batch(...).to(device, non_blocking=True)``, is what the pipelines call: a reader that reuses
        its host staging orders the next overwrite of a buffer after the copy that still reads it.  ``out``: a device
        tensor [>= B, n_chan, N] of the stored dtype to copy into (the pipeline's own staging: no allocator call per
        batch); events that already live on the device are returned as they are."""
        import torch
        host = self.read_batch(i0, i1)
        return _to_device(host, device, stream, out)

    def admin(self, i0, i1):
        """dict of per-event columns (event_number, series_number, trigger_index, ...) for events [i0, i1)"""
        return {'event_number': np.arange(i0, i1, dtype=np.int64)}

    # ---- shared helpers -------------------------------------------------------------------------------------------
    @property
    def channels(self):
        return list(self.metadata['channels'])

    @property
    def sample_rate(self):
        return float(self.metadata['sample_rate'])

    def to_amps(self, batch):
        """ADC counts -> float64 amps on whatever device ``batch`` lives on (pytesio's ``adctoamp=True``,
        processing_data.py:674-684); float data pass through."""
        import torch
        if batch.dtype != torch.int16:
            return batch.to(torch.float64)
        gain = torch.tensor(self.metadata.get('adc_gain', [1.0] * batch.shape[1]), dtype=torch.float64, device=batch.device)
        off = torch.tensor(self.metadata.get('adc_offset', [0.0] * batch.shape[1]), dtype=torch.float64, device=batch.device)
        return batch.to(torch.float64) * gain[None, :, None] + off[None, :, None]


def _to_device(host, device, stream, out):
    import torch
    if host.is_cuda:
        return host
    st = torch.cuda.current_stream(device) if stream is None else stream
    with torch.cuda.stream(st):
        if out is not None and out.dtype == host.dtype and out.shape[0] >= host.shape[0] and out.shape[1:] == host.shape[1:]:
            dst = out[:host.shape[0]]
            dst.copy_(host, non_blocking=True)
            return dst
        return host.to(device, non_blocking=True)


def _pin(t, pinned):
    import torch
    if pinned and torch.cuda.is_available():
        return t.pin_memory()
    return t


class ArrayReader(EventReader):
    """In-memory arrays behind the reader interface."""

    def __init__(self, traces, channels, sample_rate, admin=None, adc_gain=None, adc_offset=None):
        import torch
        if isinstance(traces, np.ndarray):
            traces = torch.from_numpy(traces)
        if traces.ndim == 2:
            traces = traces[:, None, :]
        if traces.shape[1] != len(channels):
            raise ValueError('traces must be [n_events, n_chan, nb_samples] with one row per channel')
        self._traces = traces
        self._admin = admin
        self.metadata = {'sample_rate': float(sample_rate), 'channels': list(channels), 'nb_samples': int(traces.shape[-1]),
                         'dtype': str(traces.dtype).replace('torch.', '')}
        if adc_gain is not None:
            self.metadata['adc_gain'] = [float(g) for g in adc_gain]
            self.metadata['adc_offset'] = [0.0] * len(channels) if adc_offset is None else [float(o) for o in adc_offset]

    def __len__(self):
        return int(self._traces.shape[0])

    def read_batch(self, i0, i1, pinned=True):
        return self._traces[i0:i1]

    def admin(self, i0, i1):
        if self._admin is None:
            return super().admin(i0, i1)
        return {k: np.asarray(v)[i0:i1] for k, v in self._admin.items()}


def write_raw_binary(path_base, samples, channels, sample_rate, adc_gain=None, adc_offset=None, admin=None, extra=None):
    """Write ``<path_base>.bin`` + ``<path_base>.json``.  samples: [n_events, n_chan, nb_samples] int16 / float32 /
    float64; admin: dict of per-event columns (lists / arrays of length n_events)."""
    samples = np.ascontiguousarray(samples)
    if samples.ndim != 3 or samples.shape[1] != len(channels):
        raise ValueError('samples must be [n_events, n_chan, nb_samples] with one row per channel')
    if samples.dtype not in (np.dtype('int16'), np.dtype('float32'), np.dtype('float64')):
        raise ValueError(f'unsupported sample dtype {samples.dtype}')
    meta = {'format': 'detprocess_b200.raw/1', 'n_events': int(samples.shape[0]), 'channels': list(channels),
            'nb_samples': int(samples.shape[2]), 'sample_rate': float(sample_rate), 'dtype': str(samples.dtype)}
    if adc_gain is not None:
        meta['adc_gain'] = [float(g) for g in adc_gain]
        meta['adc_offset'] = [0.0] * len(channels) if adc_offset is None else [float(o) for o in adc_offset]
    if admin is not None:
        for k, v in admin.items():
            if len(v) != samples.shape[0]:
                raise ValueError(f'admin column "{k}" has {len(v)} entries for {samples.shape[0]} events')
        meta['admin'] = {k: np.asarray(v).tolist() for k, v in admin.items()}
    if extra:
        meta['extra'] = extra
    samples.tofile(path_base + '.bin')
    with open(path_base + '.json', 'w') as f:
        json.dump(meta, f)
    return path_base


class RawBinaryReader(EventReader):
    """Memory-mapped ``<name>.bin`` + ``<name>.json`` written by ``write_raw_binary``; batches are copied into a
    reused pinned staging buffer (two of them, so that the copy of one batch can overlap the device work on the
    previous one)."""

    def __init__(self, path_base):
        with open(path_base + '.json') as f:
            meta = json.load(f)
        if meta.get('format') != 'detprocess_b200.raw/1':
            raise ValueError(f'{path_base}.json: unknown format {meta.get("format")!r}')
        self._admin = {k: np.asarray(v) for k, v in meta.pop('admin', {}).items()}
        self.metadata = meta
        shape = (meta['n_events'], len(meta['channels']), meta['nb_samples'])
        expected = int(np.prod(shape)) * np.dtype(meta['dtype']).itemsize
        if os.path.getsize(path_base + '.bin') != expected:
            raise ValueError(f'{path_base}.bin: size does not match the metadata ({expected} bytes expected)')
        self._mm = np.memmap(path_base + '.bin', dtype=np.dtype(meta['dtype']), mode='r', shape=shape)
        self._stage = [None, None]
        self._copied = [None, None]   # CUDA event recorded after the last asynchronous copy out of each staging buffer
        self._flip = 0
        self._last = None

    def __len__(self):
        return int(self.metadata['n_events'])

    def read_batch(self, i0, i1, pinned=True):
        import torch
        nb = i1 - i0
        k = self._flip
        self._flip ^= 1
        self._last = k
        if self._copied[k] is not None:      # an upload() of this buffer may still be in flight: wait for it
            self._copied[k].synchronize()
            self._copied[k] = None
        buf = self._stage[k]
        if buf is None or buf.shape[0] < nb:
            buf = _pin(torch.empty((nb,) + self._mm.shape[1:], dtype=getattr(torch, self.metadata['dtype'])), pinned)
            self._stage[k] = buf
        out = buf[:nb]
        out.numpy()[...] = self._mm[i0:i1]
        return out

    def upload(self, i0, i1, device, stream=None, out=None):
        import torch
        host = self.read_batch(i0, i1)
        k = self._last
        st = torch.cuda.current_stream(device) if stream is None else stream
        dev = _to_device(host, device, st, out)
        ev = torch.cuda.Event()
        ev.record(st)
        self._copied[k] = ev
        return dev

    def admin(self, i0, i1):
        if not self._admin:
            return super().admin(i0, i1)
        return {k: v[i0:i1] for k, v in self._admin.items()}
