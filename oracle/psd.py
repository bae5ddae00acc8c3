"""
CPU oracle of the noise PSD (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates ``qp.calc_psd(traces[cut], fs=fs, folded_over=False)`` as called from
``Noise.calc_psd`` (reference detprocess/core/noise.py:331-349).  QETpy is not in the
reference tree (parity unpinned); the published definition is the
two-sided periodogram averaged over traces,

    psd[k] = mean_traces |fft(x)[k]|^2 / (N * fs),      f = fftfreq(N, 1/fs)

which integrates (sum * df) to the trace variance -- that Parseval identity is the KAT
``tests/test_oracle_psd.py`` pins it with.  ``offset`` follows noise.py:349 verbatim.
"""
import numpy as np


def calc_psd(traces, fs, cut=None):
    """Returns (freqs [N], psd [N]) two-sided, float64."""
    traces = np.asarray(traces, dtype=np.float64)
    if cut is not None:
        traces = traces[np.asarray(cut, dtype=bool)]
    n = traces.shape[-1]
    acc = np.zeros(n, dtype=np.float64)
    for x in traces:                       # one trace at a time: C5 does not fit in RAM at once
        X = np.fft.fft(x)
        acc += X.real ** 2 + X.imag ** 2
    psd = acc / (traces.shape[0] * n * fs)
    return np.fft.fftfreq(n, 1.0 / fs), psd


def periodogram_sums(traces, cut=None):
    """sum_traces |fft(x)_k|^2 for k = 0..N/2 and the number of traces used (what one GPU accumulates)."""
    traces = np.asarray(traces, dtype=np.float64)
    if cut is not None:
        traces = traces[np.asarray(cut, dtype=bool)]
    n = traces.shape[-1]
    acc = np.zeros(n // 2 + 1, dtype=np.float64)
    for x in traces:
        X = np.fft.rfft(x)
        acc += X.real ** 2 + X.imag ** 2
    return acc, traces.shape[0]


def offset(traces, cut=None):
    """noise.py:349: np.average(np.median(traces[cut], axis=-1))."""
    traces = np.asarray(traces, dtype=np.float64)
    if cut is not None:
        traces = traces[np.asarray(cut, dtype=bool)]
    return float(np.average(np.median(traces, axis=-1)))


def calc_csd(traces, fs, cut=None):
    """Cross-spectral density of [n_events, n_chan, N] traces, two-sided complex [n, n, N] in fftfreq order -- restates
    ``qp.calc_csd(traces[cut], fs=fs, folded_over=False)`` of ``Noise.calc_csd`` (reference core/noise.py:374-470):

        csd[a, b, k] = mean_events fft(x_a)[k] conj(fft(x_b)[k]) / (N fs)

    i.e. the noise covariance E[X_a conj(X_b)] that ``oracle/ofnxm.py`` inverts; its diagonal is ``calc_psd``.
    PARITY UNPINNED: scipy.signal.csd (which QETpy builds on) defines P_xy = conj(X) Y, the complex conjugate of this;
    the choice is isolated here and in ``detprocess_b200/core/noise.py::csd_from_sums``."""
    traces = np.asarray(traces, dtype=np.float64)
    if cut is not None:
        traces = traces[np.asarray(cut, dtype=bool)]
    ne, n, N = traces.shape
    acc = np.zeros((n, n, N), dtype=np.complex128)
    for x in traces:
        X = np.fft.fft(x, axis=-1)
        acc += X[:, None, :] * np.conj(X[None, :, :])
    return np.fft.fftfreq(N, 1.0 / fs), acc / (ne * N * fs)


# ---- psd_amp (detprocess/core/algorithms.py:953-1042) -------------------------------------------------------------
def fold_spectrum(spectrum, fs):
    """``qp.utils.fold_spectrum`` as recalled (QETpy is not in the tree: parity unpinned): the one-sided half of a two-sided
    spectrum with every bin but DC (and Nyquist, even length) doubled; returns (rfftfreq, folded)."""
    n = spectrum.shape[-1]
    f = np.fft.rfftfreq(n, d=1.0 / fs)
    out = np.array(spectrum[..., :len(f)], dtype=np.float64)
    out[..., 1:n // 2 + n % 2] *= 2.0
    return f, out


def cleanup_freq_ranges(f_lims):
    """utils/utils.py:437-470."""
    if not isinstance(f_lims, list):
        f_lims = [f_lims]
    freq_ranges, range_names = [], []
    for freq_range in f_lims:
        if isinstance(freq_range, (float, int)):
            freq_range = [freq_range]
        f_low = abs(freq_range[0])
        if len(freq_range) == 2:
            f_high = abs(freq_range[1])
            if f_low > f_high:
                f_low, f_high = f_high, f_low
            name = f'{round(f_low)}_{round(f_high)}'
            if name not in range_names:
                freq_ranges.append([f_low, f_high])
                range_names.append(name)
        else:
            name = f'{round(f_low)}'
            if name not in range_names:
                freq_ranges.append([f_low])
                range_names.append(name)
    return freq_ranges, range_names


def get_ind_freq_ranges(freq_ranges, freqs):
    """utils/utils.py:475-505."""
    idx_ranges = []
    for freq_range in freq_ranges:
        ind_low = int(np.argmin(np.abs(freqs - abs(freq_range[0]))))
        ind_high = ind_low + 1
        if len(freq_range) == 2:
            ind_high = int(np.argmin(np.abs(freqs - abs(freq_range[1]))))
        if ind_low > ind_high:
            ind_low, ind_high = ind_high, ind_low
        if ind_low == ind_high:
            if ind_high < len(freqs) - 1:
                ind_high += 1
            elif ind_low > 0:
                ind_low -= 1
            else:
                raise ValueError('Frequency range too narrow or outside bounds.')
        idx_ranges.append([ind_low, ind_high])
    return idx_ranges


def psd_amp(trace, fs, f_lims, feature_base_name='psd_amp'):
    """algorithms.py:993-1040 on one trace: dict feature name -> average of sqrt(folded psd) over each range."""
    trace = np.asarray(trace, dtype=np.float64)
    nbins = trace.shape[-1]
    df = fs / nbins
    trace_fft = np.fft.fft(trace) / nbins / df            # OFBase.signal_fft (oracle/of1x1.py::_fft_norm)
    psd = (np.abs(trace_fft) ** 2.0) * nbins / fs
    freqs_fold, psd_fold = fold_spectrum(psd, fs)
    psd_fold = np.sqrt(psd_fold[1:])
    freqs_fold = freqs_fold[1:]
    freq_ranges, range_names = cleanup_freq_ranges(f_lims)
    out = {}
    for it, (lo, hi) in enumerate(get_ind_freq_ranges(freq_ranges, freqs_fold)):
        out[f'{feature_base_name}_{range_names[it]}'] = float(np.average(psd_fold[lo:hi]))
    return out
