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
