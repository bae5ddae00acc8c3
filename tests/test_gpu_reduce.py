"""GPU window reductions through the C ABI: bit-exact against numpy float64 (the library
the reference calls: detprocess/core/algorithms.py:698, 759, 818, 879)."""
import warnings

import numpy as np
import pytest

from oracle import reductions as R

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nan = np.isnan(b)
    return np.array_equal(np.isnan(a), nan) and np.array_equal(a[~nan].view(np.uint64), b[~nan].view(np.uint64))


def _oracle(op, traces, fs, a, b):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return {'baseline': R.baseline_batch, 'maximum': R.maximum_batch, 'minimum': R.minimum_batch,
                'integral': lambda t, a, b: R.integral_batch(t, fs, a, b)}[op](traces, a, b)


@pytest.mark.parametrize('nb_samples', [1000, 16384, 32768, 65536])
def test_reductions_bit_exact(nb_samples):
    import torch
    from detprocess_b200.core.plans import ReducePlan
    rng = np.random.default_rng(12345)
    n, fs = nb_samples, 1.25e6
    traces = rng.standard_normal((257, n)) * 1e-8 + 3e-7
    traces[7, n // 3] = np.nan
    traces[9, :] = 0.0
    feats = [('baseline', 0, n // 2 - 1250), ('integral', n // 2 - 625, n // 2 + 625), ('maximum', None, None),
             ('minimum', None, None), ('baseline', None, None), ('baseline', 100, 105), ('integral', 7, 8),
             ('baseline', 3, 3 + 129), ('integral', 1, min(n, 1026)), ('baseline', 0, n), ('maximum', n // 3 + 1, n // 2),
             ('baseline', 17, 25), ('integral', 40, 50), ('baseline', 9, 9), ('integral', 0, n), ('minimum', 5, 6)]
    plan = ReducePlan(n, fs, 1)
    handles = [plan.add(0, op, a, b) for op, a, b in feats]
    plan.finalize()
    out = plan.run(torch.from_numpy(traces).cuda()).cpu().numpy()
    for h, (op, a, b) in zip(handles, feats):
        aa = 0 if a is None else a
        bb = n - 1 if b is None else b
        ref = _oracle(op, traces, fs, aa, bb)
        assert _bits_equal(out[:, plan.column(h)], ref), (op, a, b)


def test_reductions_multichannel_layout():
    import torch
    from detprocess_b200.core.plans import ReducePlan
    rng = np.random.default_rng(4)
    n, fs, nch = 4096, 1.25e6, 3
    traces = rng.standard_normal((33, nch, n))
    plan = ReducePlan(n, fs, nch)
    hs = {}
    for c in (2, 0, 1):   # added out of order on purpose
        hs[(c, 'b')] = plan.add(c, 'baseline', 0, 1000 + c)
        hs[(c, 'i')] = plan.add(c, 'integral', 10 * c, 3000)
    plan.finalize()
    out = plan.run(torch.from_numpy(traces).cuda()).cpu().numpy()
    for c in range(nch):
        assert _bits_equal(out[:, plan.column(hs[(c, 'b')])], R.baseline_batch(traces[:, c], 0, 1000 + c))
        assert _bits_equal(out[:, plan.column(hs[(c, 'i')])], R.integral_batch(traces[:, c], fs, 10 * c, 3000))


@pytest.mark.parametrize('nb_samples', [1000, 32768])
def test_reductions_on_adc_counts_bit_exact(nb_samples):
    """int16 ADC counts + per-channel conversion (dp_reduce_plan_set_adc_conversion) == numpy on the trace the reference's
    reader converts on the host (adc.astype(float64) * gain + offset): bit for bit, two channels with different gains."""
    import torch
    from detprocess_b200.core.plans import ReducePlan
    rng = np.random.default_rng(5)
    n, fs = nb_samples, 1.25e6
    adc = rng.integers(-32768, 32767, size=(129, 2, n), dtype=np.int16)
    gains, offs = (3.0517578125e-11 * 1.7, 1.234e-11), (-2.5e-9, 7.7e-8)
    conv = np.stack([adc[:, c].astype(np.float64) * gains[c] + offs[c] for c in range(2)], axis=1)
    feats = [('baseline', 0, n // 2 - 100), ('integral', n // 2 - 62, n // 2 + 62), ('maximum', None, None), ('minimum', None, None),
             ('baseline', 3, 3 + 129), ('integral', 1, min(n, 1026)), ('baseline', 0, n), ('maximum', 5, 70)]
    plan = ReducePlan(n, fs, 2)
    hs = []
    for c in range(2):
        plan.set_adc_conversion(c, gains[c], offs[c])
        for op, a, b in feats:
            hs.append((c, op, a, b, plan.add(c, op, a, b)))
    plan.finalize()
    out = plan.run(torch.from_numpy(adc).cuda()).cpu().numpy()
    ref = plan.run(torch.from_numpy(conv).cuda()).cpu().numpy()          # the float64 path of the same plan
    assert _bits_equal(out, ref)
    for c, op, a, b, h in hs:
        lo = 0 if a is None else a
        hi = n - 1 if b is None else b
        assert _bits_equal(out[:, plan.column(h)], _oracle(op, conv[:, c], fs, lo, hi)), (c, op, a, b)
