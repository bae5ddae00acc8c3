"""
``FeatureExtractors``: drop-in for the in-scope static methods of the reference's
``detprocess/core/algorithms.py`` (same names, argument meaning, returned keys and
``-999999.0`` sentinels):

    of1x1_nodelay :278   of1x1_unconstrained :355   of1x1_constrained :436
    baseline :651        integral :709              maximum :771        minimum :830
    ofnxm :141 (joint channels 'a|b', templates [n, m, N], csd [n, n, N])

OF methods take ``(channel, of_base, ...)`` where ``of_base`` is an ``OFBaseBatch`` holding a
batch of B events on the device; trace methods take ``(trace, ...)`` where ``trace`` is
``[N]`` or ``[B, N]`` (ndarray, CPU tensor or CUDA tensor).  For a batch every returned
value is an array of length B; for a single trace it is a scalar, as in the reference.
All arithmetic runs in the CUDA library; there is no CPU path.

The class deliberately exposes no other public attribute: the reference's pipeline
treats every non-underscore name as an algorithm (``process/features.py:1112-1116``).
"""
import numpy as np

from .plans import ReducePlan

__all__ = ['FeatureExtractors']

_SENTINEL = -999999.0
# Whether window_max_index itself is a candidate delay of the constrained fit.  The
# reference forwards the index to QETpy, which slices [min:max) (see oracle/of1x1.py).
_OF_WINDOW_MAX_INCLUSIVE = False
_reduce_cache = {}


def _scalarize(of_base, d):
    if getattr(of_base, 'single', False):
        return {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    return d


def _of_window(of_base, channel, template_tag, window_min_from_trig_usec, window_max_from_trig_usec,
               window_min_index, window_max_index):
    """usec form wins over the index form (QETpy OF1x1.calc); returns rolled [lo, hi)."""
    n = of_base.nb_samples()
    fs = of_base.sample_rate()
    pre = of_base.pretrigger_samples(channel, template_tag)
    lo = hi = None
    if window_min_from_trig_usec is not None:
        lo = int(np.floor(pre + window_min_from_trig_usec * fs * 1e-6))
    elif window_min_index is not None:
        lo = int(window_min_index)
    if window_max_from_trig_usec is not None:
        hi = int(np.ceil(pre + window_max_from_trig_usec * fs * 1e-6))
    elif window_max_index is not None:
        hi = int(window_max_index)
    lo = 0 if lo is None else min(max(lo, 0), n)
    hi = n if hi is None else min(max(hi + (1 if _OF_WINDOW_MAX_INCLUSIVE else 0), 0), n)
    return lo, hi


def _reduce(trace, op, window_min_index, window_max_index, fs):
    """One windowed reduction on the device; returns ndarray [B] (or scalar for a 1-D trace)."""
    import torch
    single = trace.ndim == 1
    if isinstance(trace, np.ndarray):
        trace = torch.from_numpy(np.ascontiguousarray(trace, dtype=np.float64))
    if trace.ndim == 1:
        trace = trace[None, :]
    n = trace.shape[-1]
    key = (n, float(fs), op, window_min_index, window_max_index)
    plan = _reduce_cache.get(key)
    if plan is None:
        plan = ReducePlan(n, fs, 1)
        plan.add(0, op, window_min_index, window_max_index)
        plan.finalize()
        if len(_reduce_cache) > 256:
            _reduce_cache.clear()
        _reduce_cache[key] = plan
    x = trace.to(device=plan.device, dtype=torch.float64)
    out = plan.run(x).cpu().numpy()[:, 0]
    return out[0] if single else out


def _empty(trace):
    if trace is None:
        return True
    size = trace.size if isinstance(trace, np.ndarray) else trace.numel()
    return size == 0


class FeatureExtractors:

    @staticmethod
    def _of_fit_spec(base_algorithm, channel, of_base, template_tag=None, lowchi2_fcutoff=10000,
                     window_min_from_trig_usec=None, window_max_from_trig_usec=None, window_min_index=None,
                     window_max_index=None, lgc_outside_window=False, interpolate=False, **kwargs):
        """(template_tag, lo, hi, outside, lowchi2_fcutoff, interpolate) of the delay search an of1x1_* block will request -- lets the
        pipeline register every fit before the first batch (underscore name: not an algorithm, reference
        process/features.py:1112-1116)"""
        if base_algorithm == 'of1x1_nodelay':
            pre = of_base.pretrigger_samples(channel, template_tag)
            return template_tag, pre, pre + 1, False, lowchi2_fcutoff
        if base_algorithm == 'of1x1_unconstrained':
            return ('default' if template_tag is None else template_tag), None, None, False, lowchi2_fcutoff, bool(interpolate)
        if base_algorithm == 'of1x1_constrained':
            tag = 'default' if template_tag is None else template_tag
            lo, hi = _of_window(of_base, channel, tag, window_min_from_trig_usec, window_max_from_trig_usec,
                                window_min_index, window_max_index)
            return tag, lo, hi, bool(lgc_outside_window), lowchi2_fcutoff, bool(interpolate)
        return None

    @staticmethod
    def of1x1_nodelay(channel, of_base, template_tag=None, lowchi2_fcutoff=10000,
                      feature_base_name='of1x1_nodelay', **kwargs):
        if template_tag is None:
            raise ValueError('ERROR: Template tag required for OF 1x1')
        names = ('amp', 'chi2', 'lowchi2')
        if not of_base.is_signal_stored(channel):
            return {f'{k}_{feature_base_name}': _SENTINEL for k in names}
        pre = of_base.pretrigger_samples(channel, template_tag)
        r = of_base.results(of_base.request_fit(channel, template_tag, pre, pre + 1, lowchi2_fcutoff=lowchi2_fcutoff))
        return _scalarize(of_base, {f'{k}_{feature_base_name}': r[k] for k in names})

    @staticmethod
    def of1x1_unconstrained(channel, of_base, template_tag='default', interpolate=False,
                            lowchi2_fcutoff=10000, feature_base_name='of1x1_unconstrained', **kwargs):
        names = ('amp', 't0', 'chi2', 'lowchi2')
        if not of_base.is_signal_stored(channel):
            return {f'{k}_{feature_base_name}': _SENTINEL for k in names}
        r = of_base.results(of_base.request_fit(channel, template_tag, None, None, lowchi2_fcutoff=lowchi2_fcutoff,
                                                interpolate=interpolate), interpolate=interpolate)
        return _scalarize(of_base, {f'{k}_{feature_base_name}': r[k] for k in names})

    @staticmethod
    def of1x1_constrained(channel, of_base, template_tag='default',
                          window_min_from_trig_usec=None, window_max_from_trig_usec=None,
                          window_min_index=None, window_max_index=None, lgc_outside_window=False,
                          interpolate=False, lowchi2_fcutoff=10000,
                          feature_base_name='of1x1_constrained', **kwargs):
        names = ('amp', 't0', 'chi2', 'lowchi2', 'chi2nopulse', 'ampres', 'timeres')
        if not of_base.is_signal_stored(channel):
            return {f'{k}_{feature_base_name}': _SENTINEL for k in names}
        lo, hi = _of_window(of_base, channel, template_tag, window_min_from_trig_usec,
                            window_max_from_trig_usec, window_min_index, window_max_index)
        r = of_base.results(of_base.request_fit(channel, template_tag, lo, hi, lgc_outside_window, lowchi2_fcutoff=lowchi2_fcutoff,
                                                interpolate=interpolate), interpolate=interpolate)
        return _scalarize(of_base, {f'{k}_{feature_base_name}': r[k] for k in names})

    @staticmethod
    def ofnxm(channel, of_base, available_channels=None, feature_base_name='ofnxm', template_tag=None,
              amplitude_names=None, window_min_from_trig_usec=None, window_max_from_trig_usec=None,
              window_min_index=None, window_max_index=None, lgc_outside_window=False, lowchi2_fcutoff=10000,
              interpolate_t0=False, **kwargs):
        if template_tag is None:
            raise ValueError(f'ERROR: Missing "template_tag" argument for channel {channel}, '
                             f'algorithm "{feature_base_name}"')
        template = of_base.template(channel, template_tag=template_tag)
        if template is None:
            raise ValueError(f'ERROR: Missing template for channel {channel}, tag "{template_tag}", '
                             f'algorithm "{feature_base_name}"')
        ntmps = template.shape[1]
        if amplitude_names is None:
            amplitude_names = [f'amp{i + 1}' for i in range(ntmps)]
        else:
            if isinstance(amplitude_names, str):
                amplitude_names = [amplitude_names]
            if len(amplitude_names) != ntmps:
                raise ValueError(f'ERROR: Wrong length for "amplitude_names" argument. Expecting {ntmps} name '
                                 f'for  channel {channel}, algorithm "{feature_base_name}"')
        retdict = {f'chi2_{feature_base_name}_constrained': _SENTINEL, f't0_{feature_base_name}_constrained': _SENTINEL}
        for name in amplitude_names:
            retdict[f'{name}_{feature_base_name}_constrained'] = _SENTINEL
        retdict[f'chi2_{feature_base_name}_nodelay'] = _SENTINEL
        for name in amplitude_names:
            retdict[f'{name}_{feature_base_name}_nodelay'] = _SENTINEL
        if not of_base.is_signal_stored(channel):
            return retdict
        if interpolate_t0:
            raise NotImplementedError('interpolate_t0=True is not built')
        lo, hi = _of_window(of_base, channel, template_tag, window_min_from_trig_usec, window_max_from_trig_usec,
                            window_min_index, window_max_index)
        r = of_base.nxm_results(channel, template_tag, lo, hi, lgc_outside_window)
        retdict[f'chi2_{feature_base_name}_constrained'] = r['chi2']
        retdict[f't0_{feature_base_name}_constrained'] = r['t0']
        for i, name in enumerate(amplitude_names):
            retdict[f'{name}_{feature_base_name}_constrained'] = r['amps'][:, i]
        retdict[f'chi2_{feature_base_name}_nodelay'] = r['chi2_nodelay']
        for i, name in enumerate(amplitude_names):
            retdict[f'{name}_{feature_base_name}_nodelay'] = r['amps_nodelay'][:, i]
        return _scalarize(of_base, retdict)

    @staticmethod
    def _psd_amp_ranges(nb_samples, fs, f_lims):
        """(feature-name suffixes, one-sided bin ranges with the DC bin at 0) of a ``psd_amp`` block: the reference's
        ``cleanup_freq_ranges`` / ``get_ind_freq_ranges`` on the DC-less folded frequencies (core/algorithms.py:993-1021)"""
        from ..utils import utils
        if not f_lims:
            raise ValueError('ERROR: "f_lims" required for algorithm psd_amps')
        freq_ranges, names = utils.cleanup_freq_ranges(f_lims)
        freqs = np.fft.rfftfreq(int(nb_samples), d=1.0 / fs)[1:]
        return names, [(lo + 1, hi + 1) for lo, hi in utils.get_ind_freq_ranges(freq_ranges, freqs)]

    @staticmethod
    def psd_amp(channel, of_base, f_lims=[], feature_base_name='psd_amp', **kwargs):
        """Average of sqrt(folded PSD) of the event over frequency ranges (reference core/algorithms.py:953-1042)."""
        fs = kwargs.get('fs', of_base.sample_rate())
        names, bins = FeatureExtractors._psd_amp_ranges(of_base.nb_samples(), fs, f_lims)
        if not of_base.is_signal_stored(channel):
            return {f'{feature_base_name}_{n}': _SENTINEL for n in names}
        amps = of_base.band_amplitudes(channel, bins)
        return _scalarize(of_base, {f'{feature_base_name}_{n}': amps[:, i] for i, n in enumerate(names)})

    @staticmethod
    def baseline(trace, window_min_index=None, window_max_index=None,
                 feature_base_name='baseline', **kwargs):
        if _empty(trace):
            return {feature_base_name: _SENTINEL}
        return {feature_base_name: _reduce(trace, 'baseline', window_min_index, window_max_index,
                                           kwargs.get('fs', 1.0))}

    @staticmethod
    def integral(trace, fs, window_min_index=None, window_max_index=None,
                 feature_base_name='integral', **kwargs):
        if _empty(trace):
            return {feature_base_name: _SENTINEL}
        return {feature_base_name: _reduce(trace, 'integral', window_min_index, window_max_index, fs)}

    @staticmethod
    def maximum(trace, window_min_index=None, window_max_index=None,
                feature_base_name='maximum', **kwargs):
        if _empty(trace):
            return {feature_base_name: _SENTINEL}
        return {feature_base_name: _reduce(trace, 'maximum', window_min_index, window_max_index,
                                           kwargs.get('fs', 1.0))}

    @staticmethod
    def minimum(trace, window_min_index=None, window_max_index=None,
                feature_base_name='minimum', **kwargs):
        if _empty(trace):
            return {feature_base_name: _SENTINEL}
        return {feature_base_name: _reduce(trace, 'minimum', window_min_index, window_max_index,
                                           kwargs.get('fs', 1.0))}
