"""Window mode: OF features read straight from a continuous stream at the trigger
indices == the same features on gathered [B, N] windows, and the trigger -> features chain on device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from detprocess_b200.synth import SynthSetup, make_continuous  # noqa: E402


def _plan(S, prec, two_templates=True):
    from detprocess_b200.core.plans import OFPlan
    pre = S.nb_pretrigger
    plan = OFPlan(S.nb_samples, S.fs, 1, prec)
    plan.set_psd(0, S.psd, 'AC')
    t0 = plan.add_template(0, S.template, pre)
    plan.add_fit(0, t0, pre - 500, pre + 500)
    plan.add_fit(0, t0, None, None)
    if two_templates:
        t1 = plan.add_template(0, S.template_glitch, pre)
        plan.add_fit(0, t1, pre - 500, pre + 500)
    return plan.finalize(0)


@pytest.mark.parametrize('n,prec,two', [(16384, 'f64', True), (32768, 'f64', False), (32768, 'f32', True), (65536, 'f64', False)])
def test_windows_equal_gathered_batches(n, prec, two):
    S = SynthSetup(n)
    L = 12 * n
    x = make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(4), pulse_rate_hz=300.0, offset=1e-7)
    xs = torch.from_numpy(x).cuda()
    rng = np.random.default_rng(5)
    starts = np.concatenate([rng.integers(0, L - n, 37), [0, L - n, 1, L - n - 1]])   # even and odd, both ends
    bad = np.array([-1, L - n + 1, -n - 5, L])
    all_starts = np.concatenate([starts, bad])
    plan = _plan(S, prec, two)
    out = plan.run_windows(xs, torch.from_numpy(all_starts).cuda()).cpu().numpy()
    gathered = torch.stack([xs[s:s + n] for s in starts])
    ref = plan.run(gathered).cpu().numpy()
    assert np.array_equal(out[:len(starts)], ref)          # same arithmetic, bit for bit
    assert np.all(out[len(starts):] == -999999.0)


def test_trigger_then_features_on_device():
    """continuous stream -> trigger indices (device) -> OF features of the triggered windows, no host copy of traces"""
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    n = 16384
    S = SynthSetup(n)
    L = 40 * n
    x, t0, amps = make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(6), pulse_rate_hz=150.0,
                                  amp_range=(1e-7, 2e-7), return_truth=True)
    xs = torch.from_numpy(x).cuda()
    trig = OptimumFilterTrigger('ch', S.fs, S.template, S.psd, S.nb_pretrigger, max_samples=L)
    idx, amp, _ = trig._plan.run(xs, 50.0 ** 2, pileup_window_samples=int(2e-3 * S.fs),
                                 index_shift=trig._trigger_index_shift)
    assert idx.shape[0] > 10
    plan = _plan(S, 'f64', two_templates=False)
    feats = plan.run_windows(xs, idx - S.nb_pretrigger).cpu().numpy()
    ok = feats[:, 0] != -999999.0
    off = plan.fit_offset(0, 0)
    # the constrained fit of each triggered window finds the pulse near the pretrigger sample with the trigger amplitude
    d = feats[ok, off + 1] - S.nb_pretrigger
    assert np.max(np.abs(d)) <= 3
    assert np.allclose(feats[ok, off], amp.cpu().numpy()[ok], rtol=0.05)


@pytest.mark.parametrize('n,prec', [(16384, 'f64'), (32768, 'f32')])
def test_windows_against_the_oracle_directly(n, prec):
    """window mode meets the CPU oracle itself (not only the gathered GPU batch): odd and even starts, both stream ends,
    and the sentinel rows of windows that leave the stream"""
    from oracle.of1x1 import of1x1_batch
    S = SynthSetup(n)
    pre = S.nb_pretrigger
    L = 9 * n + 3
    x = make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(14), pulse_rate_hz=400.0, offset=-2e-7)
    starts = np.array([0, 1, 2, 12345, 12346, 3 * n + 77, L - n - 1, L - n, -1, L - n + 1, L + 5])
    good = (starts >= 0) & (starts + n <= L)
    plan = _plan(S, prec, two_templates=True)
    out = plan.run_windows(torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda()).cpu().numpy()
    assert np.all(out[~good] == -999999.0)
    traces = np.stack([x[s:s + n] for s in starts[good]])
    tol_amp, tol_chi = (1e-9, 1e-9) if prec == 'f64' else (1e-5, 1e-4)
    wins = [(pre - 500, pre + 500, False), (None, None, False)]
    for templ, fits in ((S.template, (0, 1)), (S.template_glitch, (2,))):
        o = of1x1_batch(traces, templ, S.psd, S.fs, pre, windows=wins[:len(fits)])
        for iw, f in enumerate(fits):
            off = plan.fit_offset(0, f)
            g = out[good]
            assert np.array_equal(g[:, off + 1].astype(np.int64), o['ind'][iw])
            assert np.max(np.abs(g[:, off] - o['amp'][iw]) / np.maximum(np.abs(o['amp'][iw]), 5 * o['ampres'])) < tol_amp
            assert np.max(np.abs(g[:, off + 2] / o['chi2'][iw] - 1)) < tol_chi
