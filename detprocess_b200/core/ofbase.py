"""
``OFBaseBatch``: the object passed as ``of_base`` to the ``FeatureExtractors`` OF methods.

It exposes the ``qp.OFBase`` calls detprocess makes (reference
``process/processing_data.py:278-381, 731-772``): ``set_csd``, ``add_template``,
``calc_phi``, ``phi``, ``csd``, ``clear_signal``, ``update_signal``, ``is_signal_stored``,
``calc_signal_filt``, ``calc_signal_filt_td`` -- but inverts the ownership: instead of
one trace mutated per event it holds a ``[B, N]`` batch per channel on the device, runs
the fused CUDA kernel once for every requested fit and lets the extractors index into
the result columns.  The kernel launch is lazy: the first extractor call after
``update_signal`` (or after a new fit was requested) triggers it.
"""
import numpy as np

from .plans import OFPlan, NxMPlan

__all__ = ['OFBaseBatch']


class OFBaseBatch:
    def __init__(self, sample_rate, verbose=False, precision='f64', device=None):
        self._fs = float(sample_rate)
        self._verbose = verbose
        self._precision = precision
        self._device = device
        self._nbins = None
        self._chans = []       # channel names, dense order
        self._psd = {}         # chan -> (psd, coupling)
        self._templates = {}   # chan -> {tag: (template, pretrigger, integralnorm)}
        self._fits = {}        # (chan, tag, lo, hi, outside) -> None
        self._fcut = 10000.0
        self._interpolate = False
        self._plan = None
        self._plan_key = None
        self._handles = {}
        self._signals = {}     # chan -> tensor [B, N]
        self._batch = None     # reader batch [B, n_file_chan, N] (or streams [n_file_chan, L]) consumed in place
        self._batch_rows = {}  # chan -> row of the batch
        self._batch_starts = None   # window mode: int64 [B] first sample of every event in the streams
        self._out = None       # host ndarray [B, n_out]
        # joint channels 'a|b|c' (NxM filter): csd [n, n, N], templates {tag: ([n, m, N], pretrigger)}
        self._nxm_csd = {}
        self._nxm_templates = {}
        self._nxm_plans = {}   # (chan, tag) -> NxMPlan
        self._nxm_out = {}     # (chan, tag, lo, hi, outside) -> host ndarray
        self._adc = {}         # chan -> (gain, offset): int16 signals of the channel are raw ADC counts

    # ---- reference-shaped setup -------------------------------------------------
    def sample_rate(self):
        return self._fs

    def nb_samples(self):
        return self._nbins

    def fft_freqs(self):
        return np.fft.fftfreq(self._nbins, d=1.0 / self._fs)

    def _check_n(self, n):
        if self._nbins is None:
            self._nbins = int(n)
        elif self._nbins != int(n):
            raise ValueError('ERROR: inconsistent number of samples')

    def set_csd(self, channel, csd, coupling='AC', ignored_frequency_peaks=None, ignore_harmonics=False):
        if '|' in channel:
            csd = np.asarray(csd, dtype=np.complex128)
            if csd.ndim != 3 or csd.shape[0] != csd.shape[1] or csd.shape[0] != len(channel.split('|')):
                raise ValueError(f'ERROR: csd of "{channel}" must be [n, n, N]')
            self._check_n(csd.shape[-1])
            if ignored_frequency_peaks is not None:
                # bins on an ignored peak carry no weight: the csd diagonal is marked inf there (both signs of f) and the
                # plan builder zeroes the inverse covariance of the bin
                csd = csd.copy()
                f = np.abs(self.fft_freqs())
                df = self._fs / self._nbins
                for pk in np.atleast_1d(np.asarray(ignored_frequency_peaks, dtype=float)):
                    lines = np.arange(pk, self._fs / 2, pk) if ignore_harmonics else [pk]
                    for fpk in lines:
                        sel = np.abs(f - fpk) <= df / 2
                        for a in range(csd.shape[0]):
                            csd[a, a, sel] = np.inf
            self._nxm_csd[channel] = (csd, coupling)
            self._nxm_plans = {k: v for k, v in self._nxm_plans.items() if k[0] != channel}
            return
        psd = np.array(np.real(np.asarray(csd)), dtype=np.float64).reshape(-1)
        self._check_n(psd.shape[-1])
        if ignored_frequency_peaks is not None:
            # bins on an ignored peak carry no weight: J = inf  (handled on the host: 1/J = 0)
            f = np.abs(self.fft_freqs())
            df = self._fs / self._nbins
            for pk in np.atleast_1d(np.asarray(ignored_frequency_peaks, dtype=float)):
                lines = np.arange(pk, self._fs / 2, pk) if ignore_harmonics else [pk]
                for fpk in lines:
                    psd[np.abs(f - fpk) <= df / 2] = np.inf
        if channel not in self._chans:
            self._chans.append(channel)
        self._psd[channel] = (psd, coupling)
        self._plan = None

    def csd(self, channel):
        if '|' in channel:
            return self._nxm_csd[channel][0] if channel in self._nxm_csd else None
        return self._psd[channel][0] if channel in self._psd else None

    def add_template(self, channel, template, template_tag='default', pretrigger_samples=None,
                     integralnorm=False, overwrite=False, **kwargs):
        if '|' in channel:
            template = np.asarray(template, dtype=np.float64)
            if template.ndim != 3 or template.shape[0] != len(channel.split('|')):
                raise ValueError(f'ERROR: template of "{channel}" must be [n_chan, n_templ, N]')
            self._check_n(template.shape[-1])
            if integralnorm:
                # s /= s[0] with s = fft(template) / N / df: s[0] = sum(template) / fs, a scale of the time-domain template
                template = template / (template.sum(axis=-1, keepdims=True) / self._fs)
            tags = self._nxm_templates.setdefault(channel, {})
            if template_tag in tags and not overwrite:
                raise ValueError(f'ERROR: template "{template_tag}" already exists (use overwrite=True)')
            tags[template_tag] = (template, self._nbins // 2 if pretrigger_samples is None else int(pretrigger_samples))
            self._nxm_plans.pop((channel, template_tag), None)
            return
        template = np.asarray(template, dtype=np.float64).reshape(-1)
        self._check_n(template.shape[-1])
        tags = self._templates.setdefault(channel, {})
        if template_tag in tags and not overwrite:
            raise ValueError(f'ERROR: template "{template_tag}" already exists (use overwrite=True)')
        if pretrigger_samples is None:
            pretrigger_samples = self._nbins // 2
        tags[template_tag] = (template, int(pretrigger_samples), bool(integralnorm))
        if channel not in self._chans:
            self._chans.append(channel)
        self._plan = None

    def template(self, channel, template_tag='default'):
        if '|' in channel:
            t = self._nxm_templates.get(channel, {}).get(template_tag)
            return None if t is None else t[0]
        t = self._templates.get(channel, {}).get(template_tag)
        return None if t is None else t[0]

    def template_tags(self, channel):
        return list(self._templates.get(channel, {}).keys())

    def pretrigger_samples(self, channel, template_tag='default'):
        if '|' in channel:
            return self._nxm_templates[channel][template_tag][1]
        return self._templates[channel][template_tag][1]

    def calc_phi(self, channel, template_tag='default'):
        if '|' in channel:
            self._nxm_plan(channel, template_tag)
            return
        self._ensure_plan()

    def phi(self, channel, template_tag='default'):
        if channel not in self._templates or template_tag not in self._templates[channel]:
            return None
        if channel not in self._psd:
            return None
        self._ensure_plan(finalize=False)
        c, t = self._handles[('templ', channel, template_tag)]
        return self._plan.phi(c, t)

    def norm(self, channel, template_tag='default'):
        self._ensure_plan(finalize=False)
        c, t = self._handles[('templ', channel, template_tag)]
        return self._plan.norm(c, t)

    def set_adc_conversion(self, channel, gain, offset=0.0):
        """int16 signals of ``channel`` are raw ADC counts, converted in the kernel's load: sample = adc * gain + offset
        (what the reference's reader does on the host with adctoamp=True, processing_data.py:674-684)."""
        if self._adc.get(channel) != (float(gain), float(offset)):
            self._adc[channel] = (float(gain), float(offset))
            self._plan = None

    def set_lowchi2_fcutoff(self, fcutoff):
        """default cutoff of fits requested WITHOUT their own ``lowchi2_fcutoff`` (every fit carries its own: two YAML
        blocks with different cutoffs share one plan and one launch)"""
        self._fcut = float(fcutoff)

    # ---- fits (one per YAML OF algorithm block) ---------------------------------
    def request_fit(self, channel, template_tag, lo, hi, outside=False, lowchi2_fcutoff=None, interpolate=False):
        if interpolate and not self._interpolate:
            self._interpolate = True       # the plan must report the neighbour amplitudes: rebuild it once
            self._plan = None
        fcut = self._fcut if lowchi2_fcutoff is None else float(lowchi2_fcutoff)
        key = (channel, template_tag, None if lo is None else int(lo), None if hi is None else int(hi), bool(outside), fcut)
        if key not in self._fits:
            if channel not in self._templates or template_tag not in self._templates[channel]:
                raise ValueError(f'ERROR: no template "{template_tag}" for channel {channel}')
            self._fits[key] = None
            self._plan = None
        return key

    # ---- plan -------------------------------------------------------------------
    def _ensure_plan(self, finalize=True):
        if self._plan is not None and (self._plan.finalized or not finalize):
            return
        if self._plan is None:
            chans = [c for c in self._chans if c in self._psd and c in self._templates]
            if not chans:
                raise ValueError('ERROR: no channel has both a csd and a template')
            plan = OFPlan(self._nbins, self._fs, len(chans), self._precision)
            if self._interpolate:
                plan.set_neighbours(True)
            handles = {}
            for ci, chan in enumerate(chans):
                psd, coupling = self._psd[chan]
                plan.set_psd(ci, psd, coupling)
                if chan in self._adc:
                    plan.set_adc_conversion(ci, *self._adc[chan])
                for tag, (tmpl, pre, inorm) in self._templates[chan].items():
                    handles[('templ', chan, tag)] = (ci, plan.add_template(ci, tmpl, pre, inorm))
            for key in self._fits:
                chan, tag, lo, hi, outside, fcut = key
                ci, ti = handles[('templ', chan, tag)]
                handles[('fit',) + key] = (ci, plan.add_fit(ci, ti, lo, hi, outside, lowchi2_fcutoff=fcut))
            self._plan, self._handles, self._plan_chans = plan, handles, chans
            self._out = None
        if finalize and not self._plan.finalized:
            self._plan.finalize(self._device)

    # ---- per batch --------------------------------------------------------------
    def clear_signal(self):
        self._signals = {}
        self._batch, self._batch_rows, self._batch_starts = None, {}, None
        self._out = None
        self._nxm_out = {}

    def is_signal_stored(self, channel):
        return channel in self._signals or channel in self._batch_rows

    def update_batch(self, batch, rows, start_index=None):
        """The batched form of ``update_signal`` for plain channels: ``batch`` is the reader's device tensor
        [B, n_file_chan, N] (f64 / f32 / i16) and ``rows`` maps channel name -> its row; the kernel reads the rows where
        they are (``dp_of1x1_batch_ex``).  With ``start_index`` (int64 [B]) ``batch`` holds continuous streams
        [n_file_chan, L] and event i is the window that starts at sample ``start_index[i]`` -- the reference's
        ``read_single_event(trigger_index, trace_length_samples, pretrigger_length_samples)``
        (processing_data.py:643-688) without the staging copy."""
        self._batch, self._batch_rows, self._batch_starts = batch, dict(rows), start_index
        self._single = False
        self._out = None

    def update_signal(self, channel, signal, calc_fft=True, **kwargs):
        """signal: [N] or [B, N]; ndarray, CPU tensor or CUDA tensor (f64 / f32 / i16)."""
        import torch
        if isinstance(signal, np.ndarray):
            signal = torch.from_numpy(np.ascontiguousarray(signal))
        if '|' in channel:
            self._single = signal.ndim == 2
            if signal.ndim == 2:
                signal = signal[None]
            if signal.ndim != 3 or signal.shape[1] != len(channel.split('|')) or signal.shape[-1] != self._nbins:
                raise ValueError(f'ERROR: signal of "{channel}" must be [B, n_chan, N]')
            self._signals[channel] = signal
            self._nxm_out = {k: v for k, v in self._nxm_out.items() if k[0] != channel}
            return
        self._single = signal.ndim == 1
        if signal.ndim == 1:
            signal = signal[None, :]
        if signal.shape[-1] != self._nbins:
            raise ValueError('ERROR: signal length != template/psd length')
        self._signals[channel] = signal
        self._out = None

    def signal(self, channel):
        return self._signals.get(channel)

    def band_amplitudes(self, channel, bin_ranges):
        """sqrt(folded PSD) of the stored signal of ``channel`` averaged over one-sided bin ranges (``psd_amp``): ndarray
        [B, n_bands].  The event spectrum is not kept by the fused OF kernel; the few bins of the bands are evaluated
        directly by ``dp_band_amplitudes``."""
        from .plans import BandPlan
        if channel in self._batch_rows and self._batch is not None and self._batch_starts is None:
            x = self._batch[:, self._batch_rows[channel], :]
        elif channel in self._signals:
            x = self._signals[channel]
        else:
            raise ValueError(f'ERROR: no signal stored for channel {channel}')
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)
        x = x.to(dev)
        if x.stride(-1) != 1:
            x = x.contiguous()
        key = (int(x.shape[-1]), tuple((int(a), int(b)) for a, b in bin_ranges))
        cache = self.__dict__.setdefault('_band_plans', {})
        if key not in cache:
            cache[key] = BandPlan(key[0], self._fs, key[1], device=dev)
        adc = self._adc.get(channel) if x.dtype == torch.int16 else None
        return cache[key].run(x, adc=adc).cpu().numpy()

    def calc_signal_filt(self, channel, template_tag=None):
        return None   # fused into the kernel (phi * v / norm)

    def calc_signal_filt_td(self, channel, template_tag=None):
        return None   # fused into the kernel (inverse FFT)

    def launch(self):
        """Start the fused kernel for every requested fit of the stored batch and return the device result block
        [B, n_out] (float64) WITHOUT waiting for it: the pipeline copies it to pinned host memory behind the kernel and
        hands it back with ``set_results`` once the copy has completed (the next batch is launched in between)."""
        import torch
        self._ensure_plan()
        chans = self._plan_chans
        if self._batch is not None and all(c in self._batch_rows for c in chans):
            return self._plan.run_layout(self._batch, [self._batch_rows[c] for c in chans], self._batch_starts)
        if self._batch is not None and self._batch_starts is None:
            for c in chans:       # mixed case: a plain channel next to a combined (float64) one -> float64 amps
                if c not in self._signals and c in self._batch_rows:
                    x = self._batch[:, self._batch_rows[c], :]
                    if x.dtype == torch.int16 and c in self._adc:
                        x = x.to(torch.float64) * self._adc[c][0] + self._adc[c][1]     # numpy's two roundings
                    self._signals[c] = x.to(torch.float64)
        missing = [c for c in chans if c not in self._signals]
        if missing:
            raise ValueError(f'ERROR: no signal stored for channel(s) {missing}')
        dev = self._plan.device
        cols = [self._signals[c].to(dev, non_blocking=True) for c in chans]
        x = cols[0] if len(cols) == 1 else torch.stack(cols, dim=1)
        return self._plan.run(x.contiguous())

    def set_results(self, out):
        """host ndarray [B, n_out] of a ``launch()``: what the extractors index into"""
        self._out = out

    @property
    def has_of1x1(self):
        return bool(self._fits)

    def _run(self):
        self._out = self.launch().cpu().numpy()

    @staticmethod
    def _parabola(v_prev, v_best, v_next, delta, t_interp=None):
        """three-point parabola of QETpy's interpolate_t0 (same single-function convention as
        oracle/of1x1.py::interpolate_parabola; QETpy is not in the reference tree: unpinned)"""
        sf = 1.0 / (v_best * 100.0)
        a = sf * (v_next - 2.0 * v_best + v_prev) / (2.0 * delta ** 2)
        b = sf * (v_next - v_prev) / (2.0 * delta)
        c = sf * v_best
        if t_interp is None:
            t_interp = -b / (2.0 * a)
        return t_interp, (a * t_interp ** 2 + b * t_interp + c) / sf

    def results(self, fit_key, interpolate=False):
        """dict of arrays for one fit: amp, ind, t0, chi2, lowchi2, timeres, chi2nopulse, ampres"""
        if fit_key not in self._fits:
            raise ValueError('ERROR: unknown fit (call request_fit first)')
        if self._plan is None or self._out is None:
            self._run()
        chan, tag = fit_key[0], fit_key[1]
        ci, fi = self._handles[('fit',) + fit_key]
        off = self._plan.fit_offset(ci, fi)
        o = self._out
        pre = self._templates[chan][tag][1]
        ind = o[:, off + 1].astype(np.int64)
        res = {'amp': o[:, off], 'ind': ind, 't0': (ind - pre) / self._fs, 'chi2': o[:, off + 2],
               'lowchi2': o[:, off + 3], 'timeres': o[:, off + 4],
               'chi2nopulse': o[:, self._plan.chi0_offset(ci)],
               'ampres': 1.0 / np.sqrt(self.norm(chan, tag))}
        if interpolate:
            # qp.OF1x1.calc(interpolate_t0=True): parabola through the chi2 at the best delay and its two neighbours ->
            # refined time and chi2; the amplitude parabola is evaluated at that time (the kernel reports the two neighbour
            # amplitudes, chi2 = chi0 - amp^2 norm); events whose best delay is the first / last sample stay as they are
            no = self._plan.neighbour_offset(ci, fi)
            norm = self.norm(chan, tag)
            chi0 = res['chi2nopulse']
            ap, an = o[:, no], o[:, no + 1]
            inner = np.isfinite(ap) & np.isfinite(an)
            with np.errstate(all='ignore'):
                dt, c2 = self._parabola(chi0 - ap ** 2 * norm, res['chi2'], chi0 - an ** 2 * norm, 1.0 / self._fs)
                _, a2 = self._parabola(ap, res['amp'], an, 1.0 / self._fs, t_interp=dt)
            res['amp'] = np.where(inner, a2, res['amp'])
            res['chi2'] = np.where(inner, c2, res['chi2'])
            res['t0'] = np.where(inner, res['t0'] + dt, res['t0'])
        return res

    # ---- joint channels: NxM filter --------------------------------------------
    def _nxm_plan(self, channel, template_tag):
        plan = self._nxm_plans.get((channel, template_tag))
        if plan is None:
            if channel not in self._nxm_csd:
                raise ValueError(f'ERROR: no csd for channel {channel}')
            templ, pre = self._nxm_templates[channel][template_tag]
            csd, coupling = self._nxm_csd[channel]
            plan = NxMPlan(self._nbins, self._fs, templ.shape[0], templ.shape[1], self._precision)
            plan.set_filter(templ, csd, pre, coupling)
            plan.finalize(self._device)
            self._nxm_plans[(channel, template_tag)] = plan
        return plan

    def nxm_results(self, channel, template_tag, lo, hi, outside=False):
        """dict of arrays for one NxM fit: chi0, chi2, ind, t0, amps [B, m], chi2_nodelay, amps_nodelay [B, m]"""
        import torch
        key = (channel, template_tag, lo, hi, bool(outside))
        o = self._nxm_out.get(key)
        plan = self._nxm_plan(channel, template_tag)
        if o is None:
            plan.set_window(lo, hi, outside)
            x = self._signals[channel].to(device=plan.device, dtype=torch.float64)
            o = self._nxm_out[key] = plan.run(x).cpu().numpy()
        m = plan.n_templ
        ind = o[:, 2].astype(np.int64)
        return {'chi0': o[:, 0], 'chi2': o[:, 1], 'ind': ind, 't0': np.where(ind >= 0, (ind - plan.pretrigger) / self._fs, -999999.0),
                'amps': o[:, 3:3 + m], 'chi2_nodelay': o[:, 3 + m], 'amps_nodelay': o[:, 4 + m:4 + 2 * m]}

    @property
    def single(self):
        return getattr(self, '_single', False)
