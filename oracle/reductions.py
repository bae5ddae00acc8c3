"""
CPU oracle: windowed trace reductions.  TEST INFRASTRUCTURE ONLY.

Restates ``detprocess/core/algorithms.py``:
  baseline :651-704  ``np.mean(trace[a:b])``           (slice end-exclusive, :698)
  integral :709-765  ``np.trapz(trace[a:b]) / fs``     (:759)
  maximum  :771-824  ``np.amax(trace[a:b])``           (:818)
  minimum  :830-885  ``np.amin(trace[a:b])``           (:879)
with the same defaults (a=0, b=len-1) and the same -999999.0 sentinel for a
missing/empty trace (:683-688).

Pinned: these call numpy itself -- the library the reference calls -- so the
float64 results ARE the reference's results bit for bit.  ``np.trapz`` was
removed in numpy 2 (this image has numpy 2.3); ``_trapz`` restates its body
(numpy/lib/function_base.py: ``(d * (y[1:] + y[:-1]) / 2.0).sum(axis)`` with
d = 1.0) and is checked against ``np.trapezoid`` in tests/test_oracle_reductions.py.

``pairwise_sum`` is a pure-python restatement of numpy's float64 add.reduce
(numpy/core/src/umath/loops_utils.h.src ``pairwise_sum_DOUBLE``: 8 accumulators,
128-element leaves, halves rounded down to a multiple of 8).  It is the order
of operations the CUDA kernel must reproduce to be bit-exact and is itself
pinned against ``np.add.reduce`` in the tests.
"""

import numpy as np

SENTINEL = -999999.0
PW_BLOCKSIZE = 128


def _trapz(y):
    y = np.asarray(y)
    return (1.0 * (y[1:] + y[:-1]) / 2.0).sum(-1)


def _defaults(trace, window_min_index, window_max_index):
    if window_min_index is None:
        window_min_index = 0
    if window_max_index is None:
        window_max_index = trace.shape[-1] - 1
    return window_min_index, window_max_index


def baseline(trace, window_min_index=None, window_max_index=None,
             feature_base_name='baseline', **kwargs):
    if trace is None or trace.size == 0:
        return {feature_base_name: SENTINEL}
    a, b = _defaults(trace, window_min_index, window_max_index)
    return {feature_base_name: np.mean(trace[a:b])}


def integral(trace, fs, window_min_index=None, window_max_index=None,
             feature_base_name='integral', **kwargs):
    if trace is None or trace.size == 0:
        return {feature_base_name: SENTINEL}
    a, b = _defaults(trace, window_min_index, window_max_index)
    return {feature_base_name: _trapz(trace[a:b]) / fs}


def maximum(trace, window_min_index=None, window_max_index=None,
            feature_base_name='maximum', **kwargs):
    if trace is None or trace.size == 0:
        return {feature_base_name: SENTINEL}
    a, b = _defaults(trace, window_min_index, window_max_index)
    return {feature_base_name: np.amax(trace[a:b])}


def minimum(trace, window_min_index=None, window_max_index=None,
            feature_base_name='minimum', **kwargs):
    if trace is None or trace.size == 0:
        return {feature_base_name: SENTINEL}
    a, b = _defaults(trace, window_min_index, window_max_index)
    return {feature_base_name: np.amin(trace[a:b])}


# ---- batch forms used by the parity tests (same numpy calls, axis=-1) ----------
def baseline_batch(traces, a, b):
    return np.mean(traces[..., a:b], axis=-1)


def integral_batch(traces, fs, a, b):
    y = traces[..., a:b]
    return (1.0 * (y[..., 1:] + y[..., :-1]) / 2.0).sum(-1) / fs


def maximum_batch(traces, a, b):
    return np.amax(traces[..., a:b], axis=-1)


def minimum_batch(traces, a, b):
    return np.amin(traces[..., a:b], axis=-1)


# ---- restatement of numpy's pairwise summation ----------------------------------
def pairwise_sum(a):
    """Pure-python float64 restatement of numpy's pairwise_sum_DOUBLE."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        res = np.float64(0.0)
        for i in range(n):
            res = res + a[i]
        return res
    if n <= PW_BLOCKSIZE:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])
