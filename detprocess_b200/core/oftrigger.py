"""
Continuous-stream optimal-filter trigger on the GPU -- the hot path of
``OptimumFilterTrigger`` (reference detprocess/core/oftrigger.py), BASELINE.json config C4.

Same construction / call sequence as the reference for one trigger channel and one amplitude:

    oftrigger = OptimumFilterTrigger(trigger_channel, fs, template, noisecsd, pretrigger_samples)
    oftrigger.update_trace(trace)                      # CUDA float64 tensor [L] (or [1, L])
    oftrigger.find_triggers_once(thresh=5, pileup_window_msec=1)
    data = oftrigger.get_trigger_data()                # {trigger_name: {trigger_index: [...], ...}}

The FIR filter (``oaconvolve(raw, phi_td, 'same')``, oftrigger.py:659-666), delta-chi2, edge
padding, threshold and pile-up grouping all run in ``dp_trigger_run``; the filtered trace is
never written to memory, so the threshold is needed at filter time: ``update_trace`` only
registers the device buffer and ``find_triggers_once`` launches the fused kernels.
``find_triggers(residual=True)`` (oftrigger.py:752-845) and the dynamic pile-up window (:78-141,
:975-979) work on the list of samples above threshold that the filter kernel leaves on the device.

phi is conj(s)/J under the OF conventions of ``oracle/of1x1.py`` (QETpy is not in the reference
tree; parity unpinned).  Defaults ``w = norm``, ``iw = 1/norm`` make
``filtered`` the OF amplitude and ``delta_chi2 = amp^2/sigma_amp^2``; pass ``iw`` / ``w`` to use
the values of a QETpy ``OFBase`` instead.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import lib, check
from .plans import _PREC, _stream_ptr, _in_dtype_of


def _torch():
    import torch
    return torch


def chi2_threshold_from_sigma(thresh, m_amplitudes=1):
    """sigma -> chi2 threshold exactly as oftrigger.py:961-965 (host scalar)."""
    from scipy import special, stats
    if thresh < 25:
        survival_fraction = stats.norm.sf(thresh) * 2
        return float(special.gammainccinv(m_amplitudes / 2, survival_fraction) * 2)
    return float(thresh) ** 2


class TriggerPlan:
    """``dp_trigger_plan``: FIR taps + weights -> fused filter / threshold / grouping launches."""

    def __init__(self, phi_td, iw, w, precision='f64', max_samples=12_500_000, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.DetprocessB200Error('no CUDA device: detprocess_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device) if isinstance(device, int) else torch.device(device)
        phi_td = np.ascontiguousarray(phi_td, dtype=np.float64)
        self.nb_filter = phi_td.shape[0]
        self.max_samples = int(max_samples)
        self._h = C.c_void_p()
        check(lib.dp_trigger_plan_create(C.byref(self._h), phi_td.ctypes.data, self.nb_filter, float(iw), float(w),
                                         _PREC[precision], self.max_samples, self.device.index or 0))
        f, h = C.c_int(), C.c_int()
        check(lib.dp_trigger_plan_geometry(self._h, C.byref(f), C.byref(h)))
        self.fft_size, self.hop = f.value, h.value

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.dp_trigger_plan_destroy(h)
            self._h = C.c_void_p()

    def set_scale(self, typical_rms):
        check(lib.dp_trigger_plan_set_scale(self._h, float(typical_rms)))

    def run(self, trace, chi2_threshold, pileup_window_samples=0, index_shift=0, padding=True, max_triggers=65536):
        """trace: CUDA float64 / float32 / int16 [L].  Returns (index int64 [n], amplitude [n], delta_chi2 [n]) CUDA tensors."""
        torch = _torch()
        if not trace.is_cuda or trace.ndim != 1:
            raise ValueError('run() takes a 1-D CUDA tensor')
        trace = trace.contiguous()
        # the reference returns EVERY trigger (oftrigger.py:996-1019): when the stream holds more groups than the output
        # buffers the run is repeated once with buffers sized to the count the kernel reported -- never truncated
        for attempt in range(2):
            idx = torch.empty(max_triggers, dtype=torch.int64, device=trace.device)
            amp = torch.empty(max_triggers, dtype=torch.float64, device=trace.device)
            dchi2 = torch.empty(max_triggers, dtype=torch.float64, device=trace.device)
            n = torch.zeros(1, dtype=torch.int32, device=trace.device)
            check(lib.dp_trigger_run_raw(self._h, C.c_void_p(trace.data_ptr()), _in_dtype_of(trace), trace.shape[0], float(chi2_threshold),
                                     int(pileup_window_samples), int(index_shift), int(bool(padding)),
                                     C.c_void_p(idx.data_ptr()), C.c_void_p(amp.data_ptr()), C.c_void_p(dchi2.data_ptr()),
                                     int(max_triggers), C.c_void_p(n.data_ptr()), _stream_ptr(trace.device)))
            nt = int(n.item())
            self.n_found = nt
            if nt <= max_triggers:
                return idx[:nt], amp[:nt], dchi2[:nt]
            max_triggers = nt
        raise RuntimeError(f'trigger count changed between two runs on the same stream ({nt} > {max_triggers})')

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        check(lib.dp_trigger_plan_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def candidates(self):
        """The samples above threshold of the last run (after ``residual_run``: the survivors of the second pass), ordered by
        stream index: (index int64 [K], filtered amplitude [K], delta chi2 [K]) CUDA tensors; indices are unshifted."""
        torch = _torch()
        cap = 1 << 16
        for attempt in range(2):
            idx = torch.empty(cap, dtype=torch.int64, device=self.device)
            amp = torch.empty(cap, dtype=torch.float64, device=self.device)
            val = torch.empty(cap, dtype=torch.float64, device=self.device)
            n = torch.zeros(1, dtype=torch.int64, device=self.device)
            check(lib.dp_trigger_candidates(self._h, C.c_void_p(idx.data_ptr()), C.c_void_p(amp.data_ptr()), C.c_void_p(val.data_ptr()),
                                            cap, C.c_void_p(n.data_ptr()), _stream_ptr(self.device)))
            k = int(n.item())
            if k <= cap:
                return idx[:k], amp[:k], val[:k]
            cap = k
        raise RuntimeError('candidate count changed between two reads of the same list')

    def filtered_at(self, trace, index):
        """``iw * oaconvolve(trace, phi_td, 'same')[index]`` for a few stream indices (int64 CUDA tensor) -> float64 CUDA tensor."""
        torch = _torch()
        index = index.to(device=trace.device, dtype=torch.int64).contiguous()
        out = torch.empty(index.shape[0], dtype=torch.float64, device=trace.device)
        check(lib.dp_trigger_filtered_at(self._h, C.c_void_p(trace.data_ptr()), _in_dtype_of(trace), trace.shape[0],
                                         C.c_void_p(index.data_ptr()), index.shape[0], C.c_void_p(out.data_ptr()), _stream_ptr(trace.device)))
        return out

    def residual_run(self, pulse_start, pulse_amp2, shape, chi2_threshold, pileup_window_samples=0, index_shift=0, max_triggers=65536):
        """Second pass of ``residual=True`` on the candidate list of the last ``run`` (see include/detprocess_b200.h).
        pulse_start (int64, ascending) / pulse_amp2 / shape: CUDA tensors.  Returns (index, amplitude, residual delta chi2)."""
        torch = _torch()
        dev = shape.device
        for attempt in range(2):
            idx = torch.empty(max_triggers, dtype=torch.int64, device=dev)
            amp = torch.empty(max_triggers, dtype=torch.float64, device=dev)
            dchi2 = torch.empty(max_triggers, dtype=torch.float64, device=dev)
            n = torch.zeros(1, dtype=torch.int32, device=dev)
            check(lib.dp_trigger_residual_run(self._h, C.c_void_p(pulse_start.data_ptr()), C.c_void_p(pulse_amp2.data_ptr()),
                                              int(pulse_start.shape[0]), C.c_void_p(shape.data_ptr()), int(shape.shape[0]),
                                              float(chi2_threshold), int(pileup_window_samples), int(index_shift),
                                              C.c_void_p(idx.data_ptr()), C.c_void_p(amp.data_ptr()), C.c_void_p(dchi2.data_ptr()),
                                              int(max_triggers), C.c_void_p(n.data_ptr()), _stream_ptr(dev)))
            nt = int(n.item())
            if nt <= max_triggers:
                return idx[:nt], amp[:nt], dchi2[:nt]
            # the list is already compacted: a second call with an empty pulse list only regroups it into larger buffers
            max_triggers = nt
            pulse_start, pulse_amp2 = pulse_start[:0], pulse_amp2[:0]
        raise RuntimeError('trigger count changed between two runs on the same list')


def _dynamic_ranges(x, vals, threshold_function):
    """Ranges of the ordered candidate indices ``x`` under the amplitude-dependent pile-up window of
    ``_getchangeslessthandynamicthresh`` (reference core/oftrigger.py:78-141): the gap to the next candidate is compared
    with ``threshold_function(largest delta chi2 of the open range, the next candidate included)``.  Same result as the
    reference's loop; the running maximum replaces its ``np.max`` over the whole open range and the callable is only
    re-evaluated when that maximum changes."""
    starts, ends = [], []
    current_start = 0
    n = len(x)
    if n:
        run = vals[0]
        cached_for, window = None, None
        for i in range(1, n):
            if vals[i] > run:
                run = vals[i]
            if cached_for != run:
                window = threshold_function(run)
                cached_for = run
            if (x[i] - x[i - 1]) > window:
                starts.append(current_start)
                ends.append(i)
                current_start = i
                run = vals[i]
    starts.append(current_start)
    ends.append(n)
    return list(zip(starts, ends))


class OptimumFilterTrigger:
    """1x1 mirror of the reference class (oftrigger.py:255-499, 588-679, 884-1034)."""

    def __init__(self, trigger_channel, fs, template, noisecsd, pretrigger_samples, trigger_name=None,
                 coupling='AC', iw=None, w=None, precision='f64', max_samples=12_500_000, device=None,
                 ignored_frequency_peaks=None, ignore_harmonics=False):
        template = np.asarray(template, dtype=np.float64).reshape(-1)
        noisecsd = np.real(np.asarray(noisecsd)).reshape(-1)
        if template.shape[0] != noisecsd.shape[0]:
            raise ValueError('ERROR: template and noise csd must have the same number of samples')
        self._trigger_channel = trigger_channel
        self._trigger_name = trigger_channel if trigger_name is None else trigger_name
        self._fs = float(fs)
        self._template = template
        self._nb_samples = template.shape[0]
        self._pretrigger_samples = int(pretrigger_samples)
        # oftrigger.py:456
        self._trigger_index_shift = self._pretrigger_samples - self._nb_samples // 2
        # OF pre-calculations (reference: qp.OFBase add_template / set_csd / calc_phi, oftrigger.py:468-485)
        # one-time host setup, any trace length: s = fft(template)/N/df, phi = conj(s)/J, norm = Re sum(phi s) df
        # (same conventions as csrc/dp_plan.hpp::finalize_template and oracle/of1x1.py)
        df = self._fs / self._nb_samples
        J = np.array(noisecsd, dtype=np.float64)
        if np.any(~(J > 0)):
            raise ValueError('psd must be strictly positive')
        if coupling == 'AC':
            J[0] = np.inf
        if ignored_frequency_peaks is not None:         # bins on an ignored peak carry no weight (OFBase.set_csd)
            f = np.abs(np.fft.fftfreq(self._nb_samples, d=1.0 / self._fs))
            for pk in np.atleast_1d(np.asarray(ignored_frequency_peaks, dtype=float)):
                for fpk in (np.arange(pk, self._fs / 2, pk) if ignore_harmonics else [pk]):
                    J[np.abs(f - fpk) <= df / 2] = np.inf
        s_fd = np.fft.fft(template) / self._nb_samples / df
        self._phi_fd = np.conj(s_fd) / J
        norm = float(np.real(np.sum(self._phi_fd * s_fd)) * df)
        self._w_matrix = float(norm if w is None else w)
        self._iw_matrix = float(1.0 / norm if iw is None else iw)
        phi_fd = self._phi_fd.copy()
        phi_fd[0] = 0                                   # oftrigger.py:492
        self._phi_td = np.fft.ifft(phi_fd).real         # oftrigger.py:493
        self._norm = float(np.dot(self._phi_td, template))   # oftrigger.py:496
        self._resolution = float(np.sqrt(1.0 / self._w_matrix))
        self._plan = TriggerPlan(self._phi_td, self._iw_matrix, self._w_matrix, precision=precision,
                                 max_samples=max_samples, device=device)
        self._trace = None
        self._padding = True
        self._trigger_data = None
        self._shape_dev = None
        self._shape_j = 0
        self.chi2_threshold = None

    def get_phi(self):
        return self._phi_td

    def get_resolution(self):
        return self._resolution

    def update_trace(self, trace=None, filtered_trace=None, padding=True):
        """Register the continuous trace (CUDA float64 tensor or numpy array, [L] or [1, L])."""
        torch = _torch()
        if filtered_trace is not None:
            raise NotImplementedError('pre-filtered traces are not supported by the fused trigger')
        if trace is None:
            raise ValueError('ERROR: "trace" or "filtered_trace required!')
        if isinstance(trace, np.ndarray):
            trace = torch.from_numpy(np.ascontiguousarray(trace, dtype=np.float64)).to(self._plan.device)
        if trace.ndim == 2:
            if trace.shape[0] != 1:
                raise ValueError(f'ERROR: "trace" has shape {tuple(trace.shape)}, but we have 1 channels!')
            trace = trace[0]
        self._trace = trace
        self._padding = bool(padding)

    def _trigger_dict(self, idx, amp, dchi2, thresh, pileup_window):
        """The reference's trigger dictionary (oftrigger.py:918-1034) from the arg-max arrays."""
        n = len(idx)
        d = {'trigger_delta_chi2': list(dchi2), 'trigger_time': list(idx / self._fs), 'trigger_index': list(idx),
             'trigger_pileup_window': [pileup_window] * n, 'trigger_threshold_sigma': [thresh] * n,
             'trigger_type': [4] * n, 'trigger_amplitude_0': list(amp), 'trigger_amplitude': list(amp)}
        if n > 0:
            d['trigger_channel'] = [str(self._trigger_name)] * n
        for key, val in list(d.items()):        # oftrigger.py:1029-1031
            d[key + '_' + self._trigger_name] = val
        return {self._trigger_name: d}

    @staticmethod
    def _pileup_window(fs, pileup_window_msec, pileup_window_samples):
        if pileup_window_msec is not None:
            return int(pileup_window_msec * fs / 1000)      # oftrigger.py:941
        if pileup_window_samples is not None:
            return pileup_window_samples
        return 0

    def _group_dynamic(self, thresh, dynamic_threshold_function):
        """dynamic=True (oftrigger.py:975-979): the candidate list of the last device pass, grouped on the host -- the
        window is a Python callable of the running delta-chi2 maximum, so this is host work in the reference's design as
        well; only the candidates (not the stream) come back."""
        if dynamic_threshold_function is None:
            raise ValueError('ERROR: dynamic=True requires "dynamic_threshold_function"')
        ci, ca, cv = (t.cpu().numpy() for t in self._plan.candidates())
        idx, amp, dchi2 = [], [], []
        for lo, hi in _dynamic_ranges(ci, cv, dynamic_threshold_function):
            if hi > lo:
                j = lo + int(np.argmax(cv[lo:hi]))
                idx.append(ci[j] + self._trigger_index_shift)
                amp.append(ca[j])
                dchi2.append(cv[j])
        return np.asarray(idx, dtype=np.int64), np.asarray(amp, dtype=np.float64), np.asarray(dchi2, dtype=np.float64)

    def find_triggers_once(self, thresh, pileup_window_msec=None, pileup_window_samples=None, dynamic=False,
                           dynamic_threshold_function=None, max_triggers=65536):
        if self._trace is None:
            raise ValueError('ERROR: Filter trace not available.  Use "update_trace" first!')
        pileup_window = self._pileup_window(self._fs, pileup_window_msec, pileup_window_samples)
        self.chi2_threshold = chi2_threshold_from_sigma(thresh)
        idx, amp, dchi2 = self._plan.run(self._trace, self.chi2_threshold, pileup_window, self._trigger_index_shift,
                                         self._padding, max_triggers)
        if dynamic:
            idx, amp, dchi2 = self._group_dynamic(thresh, dynamic_threshold_function)
        else:
            idx, amp, dchi2 = idx.cpu().numpy(), amp.cpu().numpy(), dchi2.cpu().numpy()
        self._trigger_data = self._trigger_dict(idx, amp, dchi2, thresh, pileup_window)
        return self._trigger_data

    def _unit_pulse_shape(self):
        """Delta-chi2 trace of a unit-amplitude template pulse, D = w (iw oaconvolve(template, phi_td, 'same'))^2, and the
        position of its maximum (oftrigger.py:796-815 with trigger_amplitudes = 1).  One-time host setup, kept on the device."""
        if self._shape_dev is None:
            from scipy.signal import oaconvolve
            f = self._iw_matrix * oaconvolve(self._template[None, :], self._phi_td[None, :], mode='same', axes=-1)[0]
            d = f * self._w_matrix * f
            self._shape_j = int(np.argmax(d))
            self._shape_dev = _torch().from_numpy(d).to(self._plan.device)
        return self._shape_dev, self._shape_j

    def _saturated(self, trigger_index, saturation_amplitudes_LPF_50kHz, positive_pulses):
        """oftrigger.py:776-786.  The reference low-passes the whole raw trace (``qp.utils.lowpassfilter``: first-order
        Butterworth, 50 kHz, ``filtfilt``) and looks nt/4 samples either side of the trigger.  A first-order IIR forgets
        its past within a few hundred samples, so the same values come from filtering a window with a 4096-sample margin
        around each trigger (host, only when a finite saturation amplitude is configured)."""
        sat = float(np.atleast_1d(np.asarray(saturation_amplitudes_LPF_50kHz, dtype=float))[0])
        n = len(trigger_index)
        flags = np.zeros(n, dtype=bool)
        if not np.isfinite(sat):
            return flags
        from scipy.signal import butter, filtfilt
        b, a = butter(1, 50e3 / (0.5 * self._fs))
        q, margin, L = int(self._nb_samples / 4), 4096, self._trace.shape[-1]
        for k, t in enumerate(trigger_index):
            lo, hi = max(int(t) - q, 0), min(int(t) + q, L)
            wlo, whi = max(lo - margin, 0), min(hi + margin, L)
            if hi <= lo:
                continue
            seg = filtfilt(b, a, self._trace[wlo:whi].double().cpu().numpy(), padtype='even')[lo - wlo:hi - wlo]
            flags[k] = bool(np.any(seg > sat)) if positive_pulses else bool(np.any(seg < -1 * sat))
        return flags

    def find_triggers(self, thresh, pileup_window_msec=None, pileup_window_samples=None, positive_pulses=True,
                      dynamic=False, dynamic_threshold_function=None, residual=False,
                      saturation_amplitudes_LPF_50kHz=None, edge_exclusion_msec=None, livetime=None,
                      return_trigger_data=False, max_triggers=65536):
        """``OptimumFilterTrigger.find_triggers`` (reference core/oftrigger.py:682-883): one ``find_triggers_once`` pass;
        with ``residual=True`` the best-fit pulse of every unsaturated first-pass trigger is subtracted in delta-chi2 space
        and the residual is triggered again (:752-845; on the device, on the candidate list -- the residual of a sample
        below threshold stays below threshold); then the edge exclusion and the livetime / edge-exclusion columns (:854-883).
        ``return_trigger_data=True`` returns the two passes' dictionaries; the delta-chi2 traces of the reference's tuple are
        never materialised here, their places hold the sparse lists ``{'index', 'delta_chi2'}`` of the samples above threshold."""
        torch = _torch()
        ret = None
        if residual:
            if saturation_amplitudes_LPF_50kHz is None:
                saturation_amplitudes_LPF_50kHz = [np.inf if positive_pulses else -np.inf]
            self.find_triggers_once(thresh, pileup_window_msec, pileup_window_samples, dynamic, dynamic_threshold_function,
                                    max_triggers=max_triggers)
            first = self._trigger_data
            first_index = np.asarray(first[self._trigger_name]['trigger_index'], dtype=np.int64)
            first_list = None
            if return_trigger_data:
                ci, _, cv = self._plan.candidates()
                first_list = {'index': ci.cpu().numpy(), 'delta_chi2': cv.cpu().numpy()}
            keep = ~self._saturated(first_index, saturation_amplitudes_LPF_50kHz, positive_pulses)
            shape, j = self._unit_pulse_shape()
            dev = self._plan.device
            pulses = torch.from_numpy(first_index[keep]).to(dev)
            # the reference reads the filtered trace at the trigger index it stored, i.e. the shifted one (:794)
            amp2 = self._plan.filtered_at(self._trace, pulses) ** 2 if len(pulses) else torch.zeros(0, dtype=torch.float64, device=dev)
            pileup_window = self._pileup_window(self._fs, pileup_window_msec, pileup_window_samples)
            idx, amp, dchi2 = self._plan.residual_run(pulses - j, amp2, shape, self.chi2_threshold, pileup_window,
                                                      self._trigger_index_shift, max_triggers)
            if dynamic:
                idx, amp, dchi2 = self._group_dynamic(thresh, dynamic_threshold_function)
            else:
                idx, amp, dchi2 = idx.cpu().numpy(), amp.cpu().numpy(), dchi2.cpu().numpy()
            second = self._trigger_dict(idx, amp, dchi2, thresh, pileup_window)
            # combine_trigger_data (oftrigger.py:262-320): second-pass triggers at new indices are appended
            seen = set(first_index.tolist())
            fresh = [k for k, t in enumerate(idx) if int(t) not in seen]
            combined = {}
            for key, val in first[self._trigger_name].items():
                combined[key] = list(val) + [second[self._trigger_name][key][k] for k in fresh]
            self._trigger_data = {self._trigger_name: combined}
            if return_trigger_data:
                ci, _, cv = self._plan.candidates()
                ret = (first, first_list, second, {'index': ci.cpu().numpy(), 'delta_chi2': cv.cpu().numpy()})
        else:
            self.find_triggers_once(thresh, pileup_window_msec, pileup_window_samples, dynamic, dynamic_threshold_function,
                                    max_triggers=max_triggers)
        if edge_exclusion_msec is not None:
            tmin = edge_exclusion_msec * 1e-3
            tmax = (self._trace.shape[-1] / self._fs) - edge_exclusion_msec * 1e-3
            for chan, data in list(self._trigger_data.items()):
                times = data['trigger_time']
                if len(times) == 0:
                    continue
                keep = [i for i, t in enumerate(times) if tmin < t < tmax]
                kept = {k: [v[i] for i in keep] for k, v in data.items()}
                kept[f'trigger_edge_exclusion_time_{chan}'] = [edge_exclusion_msec * 1e-3] * len(keep)
                if livetime is not None:
                    kept[f'trigger_livetime_{chan}'] = [livetime] * len(keep)
                self._trigger_data[chan] = kept
        return ret

    def get_trigger_data(self):
        return self._trigger_data

    def get_trigger_data_df(self):
        """The trigger table of the last ``find_triggers`` call as a pandas DataFrame (the reference returns a vaex frame
        of the same columns, core/oftrigger.py:318-321 / get_trigger_data_df); None without triggers."""
        import pandas as pd
        if not self._trigger_data:
            return None
        data = self._trigger_data[self._trigger_name]
        if len(data.get('trigger_index', [])) == 0:
            return None
        return pd.DataFrame({k: np.asarray(v) for k, v in data.items()})
