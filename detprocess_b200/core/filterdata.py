"""
``FilterData``: in-memory store of templates and two-sided PSDs per channel and tag, with the
getter / setter surface of the reference's ``detprocess/core/filterdata.py`` that the feature
path uses (``set_template`` :539, ``set_psd`` :636, ``get_template`` :450, ``get_psd`` :304,
``get_csd`` :380).  Layout ``{channel: {'template_<tag>': ..., 'psd_<tag>': ..., '<x>_metadata'}}``.
The reference persists this dict through pytesio's HDF5 ``FilterH5IO``; there is no HDF5
library in this image, so ``save`` / ``load`` use ``.npz`` behind the same two calls.
"""
import numpy as np

__all__ = ['FilterData']


class FilterData:
    def __init__(self, verbose=True, filter_data=None):
        self._verbose = verbose
        self._filter_data = filter_data if filter_data is not None else {}

    # ---- setters ----------------------------------------------------------------
    def set_template(self, channels, template, sample_rate=None, pretrigger_length_msec=None,
                     pretrigger_length_samples=None, metadata=None, tag='default'):
        if not isinstance(template, np.ndarray):
            raise ValueError('ERROR: "template" argument should be a numpy array!')
        nchan = len(channels.split('|')) if isinstance(channels, str) else 1
        if nchan > 1:
            # joint channels 'a|b': [n_chan, n_templ, N] (reference filterdata.py:539-633, NxM templates)
            if template.ndim != 3 or template.shape[0] != nchan:
                raise ValueError('ERROR: For multi-channels, expecting a 3D array [nchans, ntemplates, nsamples]')
        elif template.ndim != 1:
            raise ValueError('ERROR: For single channel, expecting and 1D array ')
        if sample_rate is None:
            raise ValueError('ERROR: "sample_rate" argument required!')
        if pretrigger_length_msec is None and pretrigger_length_samples is None:
            raise ValueError('ERROR: pretrigger length (samples or msec) required!')
        if pretrigger_length_msec is not None:
            pretrigger_length_samples = int(round(pretrigger_length_msec * sample_rate * 1e-3))
        meta = dict(metadata or {})
        meta.update({'sample_rate': sample_rate, 'nb_samples': template.shape[-1],
                     'nb_pretrigger_samples': int(pretrigger_length_samples), 'channel': channels})
        d = self._filter_data.setdefault(channels, {})
        d[f'template_{tag}'] = np.array(template, dtype=np.float64)
        d[f'template_{tag}_metadata'] = meta

    def set_psd(self, channels, psd, psd_freqs=None, sample_rate=None, metadata=None, tag='default'):
        if not isinstance(psd, np.ndarray) or psd.ndim != 1:
            raise ValueError('ERROR: Expecting a 1D "psd" numpy array (two-sided)')
        if sample_rate is None and psd_freqs is not None:
            sample_rate = 2.0 * float(np.max(np.abs(psd_freqs)))
        if sample_rate is None:
            raise ValueError('ERROR: "sample_rate" argument required!')
        meta = dict(metadata or {})
        meta.update({'sample_rate': sample_rate, 'nb_samples': psd.shape[-1], 'channel': channels})
        d = self._filter_data.setdefault(channels, {})
        d[f'psd_{tag}'] = np.array(psd, dtype=np.float64)
        d[f'psd_{tag}_metadata'] = meta

    def set_csd(self, channels, csd, csd_freqs=None, sample_rate=None, metadata=None, tag='default'):
        """Two-sided cross-spectral density [n, n, N] (complex) of joint channels 'a|b|...' (reference
        filterdata.py:754-828)."""
        csd = np.asarray(csd)
        nchan = len(channels.split('|'))
        if csd.ndim != 3 or csd.shape[0] != nchan or csd.shape[1] != nchan:
            raise ValueError('ERROR: Expecting a "csd" array [nchans, nchans, nsamples] matching the channel list')
        if sample_rate is None and csd_freqs is not None:
            sample_rate = 2.0 * float(np.max(np.abs(csd_freqs)))
        if sample_rate is None:
            raise ValueError('ERROR: "sample_rate" argument required!')
        meta = dict(metadata or {})
        meta.update({'sample_rate': sample_rate, 'nb_samples': csd.shape[-1], 'channel': channels})
        d = self._filter_data.setdefault(channels, {})
        d[f'csd_{tag}'] = np.array(csd, dtype=np.complex128)
        d[f'csd_{tag}_metadata'] = meta

    # ---- getters ----------------------------------------------------------------
    def _get(self, channel, name, tag):
        key = f'{name}_{tag}'
        if channel not in self._filter_data or key not in self._filter_data[channel]:
            raise ValueError(f'ERROR: No {name} with tag "{tag}" found for channel {channel}!')
        return self._filter_data[channel][key], self._filter_data[channel].get(key + '_metadata', {})

    def get_template(self, channel, tag='default', return_metadata=False):
        arr, meta = self._get(channel, 'template', tag)
        t = np.arange(arr.shape[-1]) / meta['sample_rate']
        return (arr, t, meta) if return_metadata else (arr, t)

    def get_psd(self, channels, tag='default', fold=False, return_metadata=False):
        arr, meta = self._get(channels, 'psd', tag)
        n, fs = arr.shape[-1], meta['sample_rate']
        f = np.fft.fftfreq(n, d=1.0 / fs)
        if fold:
            nh = n // 2 + 1
            folded = arr[:nh].copy()
            folded[1:nh - (1 if n % 2 == 0 else 0)] *= 2.0
            arr, f = folded, np.abs(f[:nh])
        return (arr, f, meta) if return_metadata else (arr, f)

    def get_csd(self, channels, tag='default', fold=False, return_metadata=False):
        """Joint channels: the stored [n, n, N] array; single channel: the PSD as a [1, 1, N] array."""
        if isinstance(channels, str) and '|' in channels:
            if fold:
                raise NotImplementedError('fold=True for a multi-channel csd is not built')
            arr, meta = self._get(channels, 'csd', tag)
            f = np.fft.fftfreq(arr.shape[-1], d=1.0 / meta['sample_rate'])
            return (arr, f, meta) if return_metadata else (arr, f)
        out = self.get_psd(channels, tag=tag, fold=fold, return_metadata=return_metadata)
        return (out[0][None, None, :],) + tuple(out[1:])

    # ---- persistence ------------------------------------------------------------
    def save(self, file_name):
        flat = {}
        for chan, d in self._filter_data.items():
            for key, val in d.items():
                if key.endswith('_metadata'):
                    for mk, mv in val.items():
                        flat[f'{chan}//{key}//{mk}'] = np.asarray(mv)
                else:
                    flat[f'{chan}//{key}'] = val
        np.savez(file_name, **flat)

    def load(self, file_name, overwrite=True):
        data = np.load(file_name, allow_pickle=False)
        if overwrite:
            self._filter_data = {}
        for name in data.files:
            parts = name.split('//')
            d = self._filter_data.setdefault(parts[0], {})
            if len(parts) == 2:
                d[parts[1]] = data[name]
            else:
                v = data[name]
                d.setdefault(parts[1], {})[parts[2]] = v.item() if v.ndim == 0 else v
