"""Pins the reduction / window oracle against numpy itself and the reference's arithmetic."""
import numpy as np
import pytest

from oracle import reductions as R
from oracle.windows import get_window_indices


def test_pairwise_restatement_is_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(0, 300)) + [1000, 1023, 1024, 1025, 4097, 15134, 32767, 65536]:
        a = rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)
        assert R.pairwise_sum(a).tobytes() == np.add.reduce(a).tobytes(), n


def test_trapz_restatement_is_numpy():
    rng = np.random.default_rng(1)
    for n in [0, 1, 2, 3, 9, 130, 1250, 32768]:
        y = rng.standard_normal(n)
        assert np.array_equal(R._trapz(y), np.trapezoid(y)) or (n == 0)


def test_extractor_semantics():
    rng = np.random.default_rng(2)
    tr = rng.standard_normal(1000)
    # defaults: a = 0, b = len - 1, end exclusive  (algorithms.py:691-698)
    assert R.baseline(tr)['baseline'] == np.mean(tr[0:999])
    assert R.maximum(tr, 10, 20, feature_base_name='m')['m'] == tr[10:20].max()
    assert R.minimum(tr, 10, 20)['minimum'] == tr[10:20].min()
    assert R.integral(tr, 1.25e6, 5, 50)['integral'] == np.trapezoid(tr[5:50]) / 1.25e6
    # sentinel for missing / empty trace (algorithms.py:683-688)
    assert R.baseline(None)['baseline'] == -999999.0
    assert R.integral(np.array([]), 1.0)['integral'] == -999999.0
    # NaN propagates through amax like numpy
    tr2 = tr.copy()
    tr2[15] = np.nan
    assert np.isnan(R.maximum(tr2, 10, 20)['maximum'])


def test_window_indices():
    fs, n, pre = 1.25e6, 32768, 16384
    # README example windows (reference README.md:87-96)
    assert get_window_indices(n, pre, fs, window_min_from_trig_usec=-500, window_max_from_trig_usec=500) == (15759, 17009)
    assert get_window_indices(n, pre, fs, window_min_from_start_usec=0, window_max_from_trig_usec=-1000) == (0, 15134)
    assert get_window_indices(n, pre, fs) == (0, n - 1)
    # to_end form and clamping
    assert get_window_indices(n, pre, fs, window_min_to_end_usec=1000, window_max_to_end_usec=0) == (n - 1250 - 1, n - 1)
    assert get_window_indices(n, pre, fs, window_min_from_trig_usec=-1e9, window_max_from_trig_usec=1e9) == (0, n - 1)
    # priority: from_start beats from_trig
    assert get_window_indices(n, pre, fs, window_min_from_start_usec=8, window_min_from_trig_usec=-5)[0] == 10
    # int() truncates toward zero
    assert get_window_indices(n, pre, fs, window_min_from_trig_usec=-0.9)[0] == pre - 1
    with pytest.raises(ValueError):
        get_window_indices(n, pre, fs, window_min_from_trig_usec=10, window_max_from_trig_usec=-10)
