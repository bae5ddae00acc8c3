"""
Thin Python objects over the C-ABI plans (include/detprocess_b200.h).

``OFPlan``     <-> ``dp_of_plan``      (one per reference ``qp.OFBase`` object)
``ReducePlan`` <-> ``dp_reduce_plan``  (the trace-window features of one config)

PyTorch is used only as the owner of device memory and streams: tensors are passed
to the library as raw device pointers.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import lib, check

_PREC = {'f64': _lib.DP_PREC_F64, 'fp64': _lib.DP_PREC_F64, 'float64': _lib.DP_PREC_F64,
         'f32': _lib.DP_PREC_F32, 'fp32': _lib.DP_PREC_F32, 'float32': _lib.DP_PREC_F32}
_OPS = {'baseline': _lib.DP_OP_BASELINE, 'integral': _lib.DP_OP_INTEGRAL,
        'maximum': _lib.DP_OP_MAXIMUM, 'minimum': _lib.DP_OP_MINIMUM}


def _torch():
    import torch
    return torch


def _in_dtype_of(t):
    torch = _torch()
    if isinstance(t, np.ndarray):
        m = {np.dtype('float64'): _lib.DP_IN_F64, np.dtype('float32'): _lib.DP_IN_F32,
             np.dtype('int16'): _lib.DP_IN_I16}
        if t.dtype not in m:
            raise ValueError(f'unsupported trace dtype {t.dtype}')
        return m[t.dtype]
    m = {torch.float64: _lib.DP_IN_F64, torch.float32: _lib.DP_IN_F32, torch.int16: _lib.DP_IN_I16}
    if t.dtype not in m:
        raise ValueError(f'unsupported trace dtype {t.dtype}')
    return m[t.dtype]


def _stream_ptr(device):
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _layout_args(base, n_chan, nb_samples, chan_rows, start_index):
    """Arguments of the ``*_batch_ex`` entry points for a reader-shaped tensor.

    batch mode  (start_index None): ``base`` [B, n_rows, N] contiguous; plan channel c = row ``chan_rows[c]``.
    window mode (start_index given): ``base`` [n_rows, L] contiguous continuous streams; event i = the N samples that
    start at ``start_index[i]`` of every stream."""
    torch = _torch()
    if not base.is_cuda:
        raise ValueError('device tensors only')
    base = base.contiguous()
    rows = list(range(n_chan)) if chan_rows is None else [int(r) for r in chan_rows]
    if len(rows) != n_chan:
        raise ValueError(f'expected {n_chan} channel rows, got {len(rows)}')
    if start_index is None:
        if base.ndim == 2:
            base = base[:, None, :]
        if base.ndim != 3 or base.shape[-1] != nb_samples:
            raise ValueError('batch must be [n_events, n_rows, nb_samples]')
        nev, nrows, row_len = base.shape
        starts, n_stream = None, 0
        event_stride = nrows * row_len
    else:
        if base.ndim == 1:
            base = base[None, :]
        if base.ndim != 2:
            raise ValueError('streams must be [n_rows, n_stream_samples]')
        nrows, row_len = base.shape
        starts = start_index.to(device=base.device, dtype=torch.int64).contiguous()
        nev, n_stream, event_stride = int(starts.shape[0]), row_len, 0
    if any(r < 0 or r >= nrows for r in rows):
        raise ValueError('channel row out of range')
    offs = (C.c_longlong * n_chan)(*[r * row_len for r in rows])
    return base, nev, event_stride, offs, row_len, starts, n_stream


class OFPlan:
    """Batched OF1x1 plan: PSD + templates + delay-search fits for ``n_chan`` channels."""

    def __init__(self, nb_samples, sample_rate, n_chan=1, precision='f64'):
        if precision not in _PREC:
            raise ValueError(f'unknown precision "{precision}"')
        self.nb_samples = int(nb_samples)
        self.sample_rate = float(sample_rate)
        self.n_chan = int(n_chan)
        self.precision = 'f32' if _PREC[precision] == _lib.DP_PREC_F32 else 'f64'
        self._h = C.c_void_p()
        check(lib.dp_of_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate,
                                    self.n_chan, _PREC[precision]))
        self.finalized = False
        self.device = None
        self.pretriggers = {}   # (chan, templ) -> pretrigger
        self.fits = {}          # (chan, fit) -> (templ, lo, hi, outside)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_of_plan_destroy(h)
            self._h = C.c_void_p()

    # ---- setup ---------------------------------------------------------------
    def set_psd(self, chan, psd, coupling='AC'):
        psd = np.ascontiguousarray(np.real(np.asarray(psd)).reshape(-1), dtype=np.float64)
        if psd.shape[0] != self.nb_samples:
            raise ValueError(f'Number of samples is not consistent between raw data '
                             f'(={self.nb_samples}) and csd (={psd.shape[0]})')
        check(lib.dp_of_plan_set_psd(self._h, int(chan), psd.ctypes.data, 1 if coupling == 'AC' else 0))

    def add_template(self, chan, template, pretrigger_samples=None, integralnorm=False):
        template = np.ascontiguousarray(np.asarray(template).reshape(-1), dtype=np.float64)
        if template.shape[0] != self.nb_samples:
            raise ValueError('Number of samples is not consistent between raw data and template')
        if pretrigger_samples is None:
            pretrigger_samples = self.nb_samples // 2
        idx = C.c_int(-1)
        check(lib.dp_of_plan_add_template(self._h, int(chan), template.ctypes.data,
                                          int(pretrigger_samples), int(bool(integralnorm)), C.byref(idx)))
        self.pretriggers[(int(chan), idx.value)] = int(pretrigger_samples)
        return idx.value

    def add_fit(self, chan, templ, window_lo=None, window_hi=None, outside=False, lowchi2_fcutoff=None):
        """``lowchi2_fcutoff``: this fit's own cutoff (Hz); None = the plan default (``set_lowchi2_fcutoff``)."""
        lo = 0 if window_lo is None else int(window_lo)
        hi = self.nb_samples if window_hi is None else int(window_hi)
        idx = C.c_int(-1)
        check(lib.dp_of_plan_add_fit_ex(self._h, int(chan), int(templ), lo, hi, int(bool(outside)),
                                        -1.0 if lowchi2_fcutoff is None else float(lowchi2_fcutoff), C.byref(idx)))
        self.fits[(int(chan), idx.value)] = (int(templ), lo, hi, bool(outside))
        return idx.value

    def add_fit_nodelay(self, chan, templ):
        pre = self.pretriggers[(int(chan), int(templ))]
        return self.add_fit(chan, templ, pre, pre + 1)

    def set_lowchi2_fcutoff(self, fcutoff):
        check(lib.dp_of_plan_set_lowchi2_fcutoff(self._h, float(fcutoff)))

    def set_adc_conversion(self, chan, gain, offset=0.0):
        """int16 traces of this channel are raw ADC counts: sample = adc * gain + offset, converted in the kernel's
        load (what H5Reader.read_single_event(adctoamp=True) does on the host, reference processing_data.py:674-684)."""
        check(lib.dp_of_plan_set_adc_conversion(self._h, int(chan), float(gain), float(offset)))

    def set_neighbours(self, on=True):
        """every fit also reports the amplitude one sample before / after its best delay (for ``interpolate_t0``)"""
        check(lib.dp_of_plan_set_neighbours(self._h, int(bool(on))))
        self.neighbours = bool(on)

    def neighbour_offset(self, chan, fit):
        o = C.c_int()
        check(lib.dp_of_plan_neighbour_offset(self._h, int(chan), int(fit), C.byref(o)))
        return o.value

    def finalize(self, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        check(lib.dp_of_plan_finalize(self._h, device.index or 0))
        self.device = device
        self.finalized = True
        n = C.c_int()
        check(lib.dp_of_plan_n_out(self._h, C.byref(n)))
        self.n_out = n.value
        return self

    # ---- layout ----------------------------------------------------------------
    def fit_offset(self, chan, fit):
        o = C.c_int()
        check(lib.dp_of_plan_fit_offset(self._h, int(chan), int(fit), C.byref(o)))
        return o.value

    def chi0_offset(self, chan):
        o = C.c_int()
        check(lib.dp_of_plan_chi0_offset(self._h, int(chan), C.byref(o)))
        return o.value

    def phi(self, chan, templ):
        out = np.empty((self.nb_samples, 2), dtype=np.float64)
        check(lib.dp_of_plan_get_phi(self._h, int(chan), int(templ), out.ctypes.data))
        return out[:, 0] + 1j * out[:, 1]

    def template_fft(self, chan, templ):
        out = np.empty((self.nb_samples, 2), dtype=np.float64)
        check(lib.dp_of_plan_get_template_fft(self._h, int(chan), int(templ), out.ctypes.data))
        return out[:, 0] + 1j * out[:, 1]

    def norm(self, chan, templ):
        v = C.c_double()
        check(lib.dp_of_plan_get_norm(self._h, int(chan), int(templ), C.byref(v)))
        return v.value

    # ---- hot calls -------------------------------------------------------------
    def _shape(self, traces):
        if traces.ndim == 2:
            if self.n_chan != 1:
                raise ValueError('expected traces [n_events, n_chan, nb_samples]')
            nev = traces.shape[0]
        elif traces.ndim == 3:
            if traces.shape[1] != self.n_chan:
                raise ValueError(f'expected {self.n_chan} channels, got {traces.shape[1]}')
            nev = traces.shape[0]
        else:
            raise ValueError('traces must be [n_events, nb_samples] or [n_events, n_chan, nb_samples]')
        if traces.shape[-1] != self.nb_samples:
            raise ValueError('ERROR: signal length != template/psd length')
        return nev

    def run(self, traces, out=None):
        """traces: CUDA tensor [B, (C,) N] (f64 / f32 / i16).  Returns CUDA f64 [B, n_out]. Async."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if not traces.is_cuda:
            raise ValueError('run() takes device tensors; use run_host() for host arrays')
        nev = self._shape(traces)
        traces = traces.contiguous()
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=traces.device)
        check(lib.dp_of1x1_batch(self._h, C.c_void_p(traces.data_ptr()), _in_dtype_of(traces), nev,
                                 self.nb_samples, C.c_void_p(out.data_ptr()), _stream_ptr(traces.device)))
        return out

    def run_layout(self, base, chan_rows=None, start_index=None, out=None):
        """The batch consumed where the reader put it (``dp_of1x1_batch_ex``): ``base`` is the reader's device batch
        [B, n_file_chan, N] (f64 / f32 / i16) and plan channel c reads row ``chan_rows[c]`` of every event -- no
        gather / stack copy.  With ``start_index`` (int64 [n_events]) ``base`` is [n_file_chan, L] continuous streams
        and event i is the window starting at sample ``start_index[i]`` of the plan's channel rows (what the reference
        reads per trigger, processing_data.py:643-688); windows that leave the stream give -999999.0.  Async."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        base, nev, event_stride, offs, row_len, starts, n_stream = _layout_args(base, self.n_chan, self.nb_samples, chan_rows, start_index)
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=base.device)
        check(lib.dp_of1x1_batch_ex(self._h, C.c_void_p(base.data_ptr()), _in_dtype_of(base), nev, event_stride, offs, row_len,
                                    C.c_void_p(starts.data_ptr()) if starts is not None else None, n_stream,
                                    C.c_void_p(out.data_ptr()), _stream_ptr(base.device)))
        return out

    def run_windows(self, stream, start_index, out=None):
        """Features of the windows ``stream[s : s + nb_samples]`` for every ``s`` in ``start_index`` (CUDA int64),
        read straight from the continuous CUDA float64 ``stream`` -- no staging copy.  ``s = trigger_index -
        nb_pretrigger_samples``.  Windows leaving the stream get -999999.0 in every column.  Async."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if not stream.is_cuda or stream.dtype != torch.float64 or stream.ndim != 1:
            raise ValueError('run_windows() takes a 1-D float64 CUDA stream')
        start_index = start_index.to(device=stream.device, dtype=torch.int64).contiguous()
        stream = stream.contiguous()
        nev = start_index.shape[0]
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=stream.device)
        check(lib.dp_of1x1_windows(self._h, C.c_void_p(stream.data_ptr()), stream.shape[0],
                                   C.c_void_p(start_index.data_ptr()), nev, C.c_void_p(out.data_ptr()),
                                   _stream_ptr(stream.device)))
        return out

    def run_host(self, traces, out=None):
        """traces: host ndarray or (pinned) CPU tensor.  Returns ndarray [B, n_out].  Synchronous."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if isinstance(traces, np.ndarray):
            traces = np.ascontiguousarray(traces)
            ptr = traces.ctypes.data
        else:
            if traces.is_cuda:
                raise ValueError('run_host() takes host buffers')
            traces = traces.contiguous()
            ptr = traces.data_ptr()
        nev = self._shape(traces)
        if out is None:
            out = np.empty((nev, self.n_out), dtype=np.float64)
        optr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        check(lib.dp_of1x1_batch_host(self._h, C.c_void_p(ptr), _in_dtype_of(traces), nev,
                                      self.nb_samples, C.c_void_p(optr)))
        return out

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib.dp_of_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_longlong()
        check(lib.dp_of_plan_launch_count(self._h, C.byref(n)))
        return n.value


class NxMPlan:
    """Batched NxM optimal filter (``qp.OFnxm``): n channels with an [n, n, N] cross-spectral density, m templates
    [n, m, N] sharing one time delay.  Output row: chi0, chi2, index, amps[m], chi2_nodelay, amps_nodelay[m]."""

    def __init__(self, nb_samples, sample_rate, n_chan, n_templ, precision='f64'):
        if precision not in _PREC:
            raise ValueError(f'unknown precision "{precision}"')
        self.nb_samples, self.sample_rate = int(nb_samples), float(sample_rate)
        self.n_chan, self.n_templ = int(n_chan), int(n_templ)
        self._h = C.c_void_p()
        check(lib.dp_nxm_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate, self.n_chan, self.n_templ,
                                     _PREC[precision]))
        self.finalized = False
        self.n_out = 4 + 2 * self.n_templ
        self.pretrigger = None

    def __del__(self):
        try:
            if self._h:
                lib.dp_nxm_plan_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def set_filter(self, templates, csd, pretrigger_samples=None, coupling='AC'):
        templates = np.ascontiguousarray(templates, dtype=np.float64)
        csd = np.ascontiguousarray(csd, dtype=np.complex128)
        if templates.shape != (self.n_chan, self.n_templ, self.nb_samples):
            raise ValueError(f'templates must be [n_chan, n_templ, nb_samples], got {templates.shape}')
        if csd.shape != (self.n_chan, self.n_chan, self.nb_samples):
            raise ValueError(f'csd must be [n_chan, n_chan, nb_samples], got {csd.shape}')
        pre = self.nb_samples // 2 if pretrigger_samples is None else int(pretrigger_samples)
        check(lib.dp_nxm_plan_set_filter(self._h, C.c_void_p(templates.ctypes.data), C.c_void_p(csd.ctypes.data), pre,
                                         int(coupling == 'AC')))
        self.pretrigger = pre

    def set_window(self, window_lo=None, window_hi=None, outside=False):
        lo = 0 if window_lo is None else int(window_lo)
        hi = self.nb_samples if window_hi is None else int(window_hi)
        check(lib.dp_nxm_plan_set_window(self._h, lo, hi, int(bool(outside))))

    def finalize(self, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        check(lib.dp_nxm_plan_finalize(self._h, device.index or 0))
        self.device = device
        self.finalized = True
        return self

    def p_matrix(self):
        P = np.empty((self.n_templ, self.n_templ))
        Pinv = np.empty_like(P)
        check(lib.dp_nxm_plan_get_p_matrix(self._h, C.c_void_p(P.ctypes.data), C.c_void_p(Pinv.ctypes.data)))
        return P, Pinv

    def run(self, traces, out=None):
        """traces: CUDA float64 [B, n_chan, N].  Returns CUDA float64 [B, n_out].  Async."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if not traces.is_cuda or traces.dtype != torch.float64:
            raise ValueError('run() takes CUDA float64 tensors')
        if traces.ndim != 3 or traces.shape[1] != self.n_chan or traces.shape[2] != self.nb_samples:
            raise ValueError(f'traces must be [B, {self.n_chan}, {self.nb_samples}], got {tuple(traces.shape)}')
        traces = traces.contiguous()
        nev = traces.shape[0]
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=traces.device)
        check(lib.dp_ofnxm_batch(self._h, C.c_void_p(traces.data_ptr()), nev, self.n_chan * self.nb_samples, self.nb_samples,
                                 C.c_void_p(out.data_ptr()), _stream_ptr(traces.device)))
        return out

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib.dp_nxm_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class ReducePlan:
    """Bit-exact windowed baseline / integral / maximum / minimum for ``n_chan`` channels."""

    def __init__(self, nb_samples, sample_rate, n_chan=1):
        self.nb_samples = int(nb_samples)
        self.sample_rate = float(sample_rate)
        self.n_chan = int(n_chan)
        self._h = C.c_void_p()
        check(lib.dp_reduce_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate, self.n_chan))
        self.finalized = False
        self._added = []  # (chan, feat_index)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_reduce_plan_destroy(h)
            self._h = C.c_void_p()

    def add(self, chan, op, window_min_index=None, window_max_index=None):
        """Same defaults as the reference extractors: a=0, b=len-1 (algorithms.py:691-696)."""
        if isinstance(op, str):
            op = _OPS[op]
        a = 0 if window_min_index is None else int(window_min_index)
        b = self.nb_samples - 1 if window_max_index is None else int(window_max_index)
        idx = C.c_int(-1)
        check(lib.dp_reduce_plan_add(self._h, int(chan), int(op), a, b, C.byref(idx)))
        self._added.append((int(chan), idx.value))
        return len(self._added) - 1

    def set_adc_conversion(self, chan, gain, offset=0.0):
        """int16 traces of this channel are raw ADC counts: sample = adc * gain + offset, rounded like numpy's
        ``adc.astype(float64) * gain + offset``, so the reductions stay bit-identical to numpy on the converted trace."""
        check(lib.dp_reduce_plan_set_adc_conversion(self._h, int(chan), float(gain), float(offset)))

    def finalize(self, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        check(lib.dp_reduce_plan_finalize(self._h, device.index or 0))
        self.device = device
        n = C.c_int()
        check(lib.dp_reduce_plan_n_out(self._h, C.byref(n)))
        self.n_out = n.value
        self.columns = []
        for chan, fi in self._added:
            c = C.c_int()
            check(lib.dp_reduce_plan_column(self._h, chan, fi, C.byref(c)))
            self.columns.append(c.value)
        self.finalized = True
        return self

    def column(self, handle):
        return self.columns[handle]

    def run(self, traces, out=None):
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if not traces.is_cuda or traces.dtype not in (torch.float64, torch.int16):
            raise ValueError('run() takes float64 or int16 CUDA tensors')
        if traces.ndim == 2 and self.n_chan == 1:
            nev = traces.shape[0]
        elif traces.ndim == 3 and traces.shape[1] == self.n_chan:
            nev = traces.shape[0]
        else:
            raise ValueError('traces must be [n_events, n_chan, nb_samples]')
        if traces.shape[-1] != self.nb_samples:
            raise ValueError('trace length != plan nb_samples')
        traces = traces.contiguous()
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=traces.device)
        check(lib.dp_window_reduce_batch_raw(self._h, C.c_void_p(traces.data_ptr()), _in_dtype_of(traces), nev, self.nb_samples,
                                             C.c_void_p(out.data_ptr()), _stream_ptr(traces.device)))
        return out

    def run_layout(self, base, chan_rows=None, start_index=None, out=None):
        """``OFPlan.run_layout`` for the window reductions (``dp_window_reduce_batch_ex``): float64 or int16 reader batch
        [B, n_file_chan, N] consumed in place, or windows of continuous streams [n_file_chan, L] at ``start_index``."""
        torch = _torch()
        if not self.finalized:
            raise _lib.DetprocessB200Error('plan not finalized')
        if base.dtype not in (torch.float64, torch.int16):
            raise ValueError('run_layout() takes float64 or int16 CUDA tensors')
        base, nev, event_stride, offs, row_len, starts, n_stream = _layout_args(base, self.n_chan, self.nb_samples, chan_rows, start_index)
        if out is None:
            out = torch.empty((nev, self.n_out), dtype=torch.float64, device=base.device)
        check(lib.dp_window_reduce_batch_ex(self._h, C.c_void_p(base.data_ptr()), _in_dtype_of(base), nev, event_stride, offs, row_len,
                                            C.c_void_p(starts.data_ptr()) if starts is not None else None, n_stream,
                                            C.c_void_p(out.data_ptr()), _stream_ptr(base.device)))
        return out

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib.dp_reduce_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class BandPlan:
    """``dp_band_plan``: averages of sqrt(folded event PSD) over one-sided bin ranges -- ``FeatureExtractors.psd_amp``
    (reference core/algorithms.py:953-1042).  ``bin_ranges``: [(lo, hi)] with the DC bin at 0 (the reference's indices into
    the DC-less folded spectrum + 1)."""

    def __init__(self, nb_samples, sample_rate, bin_ranges, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        self.nb_samples, self.sample_rate = int(nb_samples), float(sample_rate)
        lo = np.ascontiguousarray([int(a) for a, _ in bin_ranges], dtype=np.int32)
        hi = np.ascontiguousarray([int(b) for _, b in bin_ranges], dtype=np.int32)
        self.n_bands = len(lo)
        self._h = C.c_void_p()
        check(lib.dp_band_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate, lo.ctypes.data, hi.ctypes.data,
                                      self.n_bands, self.device.index or 0))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_band_plan_destroy(h)
            self._h = C.c_void_p()

    def run(self, x, adc=None, out=None):
        """x: CUDA tensor [B, N] (f64 / f32 / i16), rows may be strided (e.g. ``batch[:, row, :]`` of a reader batch);
        adc = (gain, offset) of int16 samples.  Returns float64 [B, n_bands] on the device."""
        torch = _torch()
        if not x.is_cuda or x.ndim != 2 or x.shape[1] != self.nb_samples or x.stride(1) != 1:
            raise ValueError('run() takes a CUDA tensor [B, nb_samples] with contiguous samples')
        nb = int(x.shape[0])
        if out is None:
            out = torch.empty((nb, self.n_bands), dtype=torch.float64, device=x.device)
        gain, offset = (1.0, 0.0) if adc is None else (float(adc[0]), float(adc[1]))
        stride = int(x.stride(0)) if nb > 1 else self.nb_samples
        check(lib.dp_band_amplitudes(self._h, C.c_void_p(x.data_ptr()), _in_dtype_of(x), nb, stride, gain, offset,
                                     C.c_void_p(out.data_ptr()), _stream_ptr(x.device)))
        return out


def combine_channels(batch, terms, adc=None, out=None):
    """Weighted channel algebra on the device in ONE launch (``dp_channel_combine``; reference
    ``ProcessingData.get_channel_trace``, processing_data.py:1033-1047).

    batch : CUDA tensor [B, n_file_chan, N] (f64 / f32 / i16) as the reader delivered it
    terms : list (one entry per combined channel) of lists of ``(row, weight)``; weight None = plain ``a + b`` /
            ``a - b`` (pass ``(row, None, sign)`` with sign -1 for the subtrahend)
    adc   : {row: (gain, offset)} for int16 batches
    Returns float64 [B, len(terms), N], bit-identical to numpy on the host-converted traces."""
    torch = _torch()
    if not batch.is_cuda or batch.ndim != 3:
        raise ValueError('combine_channels() takes a CUDA batch [B, n_chan, N]')
    batch = batch.contiguous()
    nb, nrows, n = batch.shape
    n_out = len(terms)
    if not 1 <= n_out <= 8:
        raise ValueError('1..8 combined channels per call')
    nt = (C.c_int * n_out)()
    offs = (C.c_longlong * (4 * n_out))()
    w = (C.c_double * (4 * n_out))()
    weighted = (C.c_int * n_out)()
    gain = (C.c_double * (4 * n_out))(*([1.0] * (4 * n_out)))
    aoff = (C.c_double * (4 * n_out))()
    for j, tl in enumerate(terms):
        if not 1 <= len(tl) <= 4:
            raise ValueError('1..4 terms per combined channel')
        nt[j] = len(tl)
        weighted[j] = int(all(t[1] is not None for t in tl))
        for k, t in enumerate(tl):
            row = int(t[0])
            sign = float(t[2]) if len(t) > 2 else 1.0
            offs[4 * j + k] = row * n
            w[4 * j + k] = float(t[1]) * sign if t[1] is not None else sign
            if adc is not None and row in adc:
                gain[4 * j + k], aoff[4 * j + k] = float(adc[row][0]), float(adc[row][1])
    if out is None:
        out = torch.empty((nb, n_out, n), dtype=torch.float64, device=batch.device)
    check(lib.dp_channel_combine(C.c_void_p(batch.data_ptr()), _in_dtype_of(batch), nb, nrows * n, n, n_out, nt, offs, w, weighted,
                                 gain, aoff, C.c_void_p(out.data_ptr()), _stream_ptr(batch.device)))
    return out


class PSDPlan:
    """Per-GPU accumulation of sum_traces |fft(x)_k|^2, k = 0..N/2 (``dp_psd_plan``).

    ``accumulate`` may be called any number of times (C5 streams 262 GB through it);
    ``sums`` returns the per-GPU sums and accepted-trace count as CUDA tensors so the host
    layer can all-reduce them over NCCL before forming the PSD.
    """

    def __init__(self, nb_samples, sample_rate, precision='f64', device=None):
        torch = _torch()
        if precision not in _PREC:
            raise ValueError(f'unknown precision "{precision}"')
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        self.device = device
        self.nb_samples = int(nb_samples)
        self.sample_rate = float(sample_rate)
        self.precision = 'f32' if _PREC[precision] == _lib.DP_PREC_F32 else 'f64'
        self._h = C.c_void_p()
        check(lib.dp_psd_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate, _PREC[precision],
                                     device.index or 0))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_psd_plan_destroy(h)
            self._h = C.c_void_p()

    def set_scale(self, typical_rms):
        check(lib.dp_psd_plan_set_scale(self._h, float(typical_rms)))

    def reset(self):
        check(lib.dp_psd_reset(self._h, _stream_ptr(self.device)))

    def accumulate(self, traces, mask=None):
        """traces: CUDA float64 / float32 / int16 [n, N]; mask: optional CUDA bool/uint8 [n] (True = keep)."""
        torch = _torch()
        if not traces.is_cuda:
            raise ValueError('accumulate() takes CUDA tensors')
        if traces.ndim != 2 or traces.shape[1] != self.nb_samples:
            raise ValueError('traces must be [n_traces, nb_samples]')
        traces = traces.contiguous()
        mptr = C.c_void_p(0)
        if mask is not None:
            mask = mask.to(device=traces.device, dtype=torch.uint8).contiguous()
            if mask.shape != (traces.shape[0],):
                raise ValueError('mask must be [n_traces]')
            mptr = C.c_void_p(mask.data_ptr())
        check(lib.dp_psd_accumulate(self._h, C.c_void_p(traces.data_ptr()), _in_dtype_of(traces), traces.shape[0],
                                    self.nb_samples, mptr, _stream_ptr(traces.device)))

    def sums(self):
        """(sums [N/2+1] float64, count [1] int64) CUDA tensors of everything accumulated so far."""
        torch = _torch()
        sums = torch.empty(self.nb_samples // 2 + 1, dtype=torch.float64, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        check(lib.dp_psd_get_sums(self._h, C.c_void_p(sums.data_ptr()), C.c_void_p(count.data_ptr()),
                                  _stream_ptr(self.device)))
        return sums, count

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib.dp_psd_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class CSDPlan:
    """Per-GPU accumulation of sum_events X_a[k] conj(X_b[k]), k = 0..N/2, for n_chan channels (``dp_csd_plan``).
    ``sums`` returns [n_chan * n_chan, N/2 + 1] float64 (diagonal rows, then (re, im) rows per pair a < b) and the
    accepted-event count, as CUDA tensors, so that the host layer can all-reduce them before forming the CSD."""

    def __init__(self, nb_samples, sample_rate, n_chan, precision='f64', device=None):
        torch = _torch()
        if precision not in _PREC:
            raise ValueError(f'unknown precision "{precision}"')
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        self.device = device
        self.nb_samples, self.sample_rate, self.n_chan = int(nb_samples), float(sample_rate), int(n_chan)
        self._h = C.c_void_p()
        check(lib.dp_csd_plan_create(C.byref(self._h), self.nb_samples, self.sample_rate, self.n_chan, _PREC[precision],
                                     device.index or 0))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_csd_plan_destroy(h)
            self._h = C.c_void_p()

    def set_scale(self, typical_rms):
        check(lib.dp_csd_plan_set_scale(self._h, float(typical_rms)))

    def reset(self):
        check(lib.dp_csd_reset(self._h, _stream_ptr(self.device)))

    def accumulate(self, traces, mask=None):
        """traces: CUDA float64 [n_events, n_chan, N]; mask: optional CUDA bool/uint8 [n_events] (True = keep)."""
        torch = _torch()
        if not traces.is_cuda or traces.dtype != torch.float64:
            raise ValueError('accumulate() takes CUDA float64 tensors')
        if traces.ndim != 3 or traces.shape[1] != self.n_chan or traces.shape[2] != self.nb_samples:
            raise ValueError('traces must be [n_events, n_chan, nb_samples]')
        traces = traces.contiguous()
        mptr = C.c_void_p(0)
        if mask is not None:
            mask = mask.to(device=traces.device, dtype=torch.uint8).contiguous()
            if mask.shape != (traces.shape[0],):
                raise ValueError('mask must be [n_events]')
            mptr = C.c_void_p(mask.data_ptr())
        check(lib.dp_csd_accumulate(self._h, C.c_void_p(traces.data_ptr()), traces.shape[0], self.n_chan * self.nb_samples,
                                    self.nb_samples, mptr, _stream_ptr(traces.device)))

    def sums(self):
        torch = _torch()
        sums = torch.empty((self.n_chan * self.n_chan, self.nb_samples // 2 + 1), dtype=torch.float64, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        check(lib.dp_csd_get_sums(self._h, C.c_void_p(sums.data_ptr()), C.c_void_p(count.data_ptr()), _stream_ptr(self.device)))
        return sums, count

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib.dp_csd_plan_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value
