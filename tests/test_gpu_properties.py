"""Size-independent properties at BASELINE.json's full trace lengths (no oracle in the loop): linearity, circular
shift equivariance of the unconstrained fit, Parseval for the PSD, chunk-boundary invariance of the trigger."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from detprocess_b200.synth import SynthSetup, make_traces, make_continuous  # noqa: E402


def _plan(S, prec):
    from detprocess_b200.core.plans import OFPlan
    pre = S.nb_pretrigger
    plan = OFPlan(S.nb_samples, S.fs, 1, prec)
    plan.set_psd(0, S.psd, 'AC')
    t0 = plan.add_template(0, S.template, pre)
    f_un = plan.add_fit(0, t0, None, None)
    f_con = plan.add_fit(0, t0, pre - 500, pre + 500)
    t1 = plan.add_template(0, S.template_glitch, pre)
    f_gl = plan.add_fit(0, t1, pre - 500, pre + 500)
    return plan.finalize(0), (f_un, f_con, f_gl)


@pytest.mark.parametrize('n', [32768, 16384])
def test_of_linearity_and_shift_fp64(n):
    S = SynthSetup(n)
    plan, (f_un, f_con, f_gl) = _plan(S, 'f64')
    x = make_traces(1024, S.template, S.psd, S.fs, np.random.default_rng(31), amp_max=3e-7)
    xs = torch.from_numpy(x).cuda()
    a = plan.run(xs).cpu().numpy()
    b = plan.run(2.0 * xs).cpu().numpy()          # exact power-of-two scaling: same rounding everywhere
    for f in (f_un, f_con, f_gl):
        o = plan.fit_offset(0, f)
        assert np.array_equal(b[:, o + 1], a[:, o + 1])                     # same delay
        assert np.array_equal(b[:, o], 2.0 * a[:, o])                       # amp scales exactly
        assert np.allclose(b[:, o + 2], 4.0 * a[:, o + 2], rtol=1e-12)      # chi2 scales (difference of two exact scalings)
    assert np.array_equal(b[:, plan.chi0_offset(0)], 4.0 * a[:, plan.chi0_offset(0)])
    # circular shift by d samples moves the unconstrained delay by d and leaves amp / chi2 alone
    d = 137
    c = plan.run(torch.roll(xs, d, dims=1)).cpu().numpy()
    o = plan.fit_offset(0, f_un)
    assert np.array_equal((a[:, o + 1] + d) % n, c[:, o + 1])
    assert np.allclose(c[:, o], a[:, o], rtol=1e-9, atol=1e-20)
    assert np.allclose(c[:, o + 2], a[:, o + 2], rtol=1e-9)


def test_fast_mode_tracks_fp64_on_a_large_batch():
    """north_star fp32 tolerances on 8192 events of 32768 samples, against this library's own fp64 mode."""
    S = SynthSetup(32768)
    p64, fits = _plan(S, 'f64')
    p32, _ = _plan(S, 'f32')
    xs = torch.from_numpy(make_traces(2048, S.template, S.psd, S.fs, np.random.default_rng(32))).cuda().repeat(4, 1)
    a = p64.run(xs).cpu().numpy()
    b = p32.run(xs).cpu().numpy()
    ampres = 1.0 / np.sqrt(p64.norm(0, 0))
    for f in fits[:2]:
        o = p64.fit_offset(0, f)
        same = a[:, o + 1] == b[:, o + 1]
        assert same.mean() > 0.99
        den = np.maximum(np.abs(a[:, o]), 5 * ampres)
        assert np.max((np.abs(a[:, o] - b[:, o]) / den)[same]) < 1e-5
        assert np.max(np.abs(b[:, o + 2] / a[:, o + 2] - 1)[same]) < 1e-4


def test_psd_parseval_65536():
    from detprocess_b200.core.noise import NoisePSD
    n, fs = 65536, 1.25e6
    x = torch.randn((512, n), dtype=torch.float64, device='cuda') * 3e-10 + 1e-8
    est = NoisePSD(n, fs, device=0)
    est.update(x[:200])
    est.update(x[200:])
    _, psd = est.finalize()
    assert np.isclose(psd.sum() * fs / n, float((x ** 2).mean().item()), rtol=1e-12)
    assert np.allclose(psd[1:n // 2], psd[:n // 2:-1], rtol=0, atol=0)      # two-sided, even


def test_trigger_independent_of_stream_offset():
    """The same pulses trigger at the same stream positions wherever the overlap-save chunk boundaries fall."""
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    nt = 16384
    S = SynthSetup(nt)
    L = 30 * nt
    x = make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(33), pulse_rate_hz=120.0)
    trig = OptimumFilterTrigger('ch', S.fs, S.template, S.psd, S.nb_pretrigger, max_samples=L)
    xs = torch.from_numpy(x).cuda()
    ref = None
    for cut in (0, 2, 5000, 16385):
        idx, amp, _ = trig._plan.run(xs[cut:].clone(), 36.0, pileup_window_samples=1250, index_shift=0)
        idx = idx.cpu().numpy() + cut
        amp = amp.cpu().numpy()
        keep = (idx > 3 * nt) & (idx < L - 3 * nt)       # away from the (moving) zeroed edges
        if ref is None:
            ref = (idx[keep], amp[keep])
            assert len(ref[0]) > 5
        else:
            assert np.array_equal(idx[keep], ref[0])
            assert np.allclose(amp[keep], ref[1], rtol=1e-9)


def test_nxm_linearity_shift_and_shard_invariance():
    """NxM filter at 32768 samples, 3 channels x 2 templates, 1024 events: power-of-two scaling is exact, a circular
    shift of every channel moves the unconstrained delay, halves processed separately == the whole batch."""
    from detprocess_b200.core.plans import NxMPlan
    from detprocess_b200.synth import SynthNxM
    S = SynthNxM(32768, 3, 2)
    n, pre = S.nb_samples, S.nb_pretrigger
    plan = NxMPlan(n, S.fs, 3, 2, 'f64')
    plan.set_filter(S.templates, S.csd, pre, 'AC')
    plan.finalize(0)
    x = torch.from_numpy(S.traces(1024, np.random.default_rng(41), amp_max=3e-7)).cuda()
    a = plan.run(x).cpu().numpy()
    b = plan.run(4.0 * x).cpu().numpy()
    assert np.array_equal(b[:, 2], a[:, 2])                                  # same delay
    assert np.array_equal(b[:, 3:5], 4.0 * a[:, 3:5]) and np.array_equal(b[:, 6:8], 4.0 * a[:, 6:8])   # amplitudes
    assert np.array_equal(b[:, 0], 16.0 * a[:, 0])                           # chi0
    assert np.allclose(b[:, 1], 16.0 * a[:, 1], rtol=1e-12)
    d = 211
    c = plan.run(torch.roll(x, d, dims=2)).cpu().numpy()
    assert np.array_equal((a[:, 2] + d) % n, c[:, 2])
    assert np.allclose(c[:, 3:5], a[:, 3:5], rtol=1e-9, atol=1e-9 * np.max(np.abs(a[:, 3:5])))
    assert np.allclose(c[:, 1], a[:, 1], rtol=1e-9)
    lo = plan.run(x[:500]).cpu().numpy()
    hi = plan.run(x[500:]).cpu().numpy()
    assert np.array_equal(np.concatenate([lo, hi]), a)                       # bitwise: events are independent


def test_nxm_single_channel_equals_the_1x1_kernel():
    """n = m = 1: the NxM kernel and the OF1x1 kernel are two implementations of the same fit (4096 events, 32768
    samples, constrained window + no-delay)."""
    from detprocess_b200.core.plans import NxMPlan, OFPlan
    S = SynthSetup(32768)
    n, pre = S.nb_samples, S.nb_pretrigger
    x = torch.from_numpy(make_traces(4096, S.template, S.psd, S.fs, np.random.default_rng(42))).cuda()
    one = OFPlan(n, S.fs, 1, 'f64')
    one.set_psd(0, S.psd, 'AC')
    t = one.add_template(0, S.template, pre)
    f_nd = one.add_fit_nodelay(0, t)
    f_c = one.add_fit(0, t, pre - 500, pre + 500)
    one.finalize(0)
    o1 = one.run(x).cpu().numpy()
    nxm = NxMPlan(n, S.fs, 1, 1, 'f64')
    nxm.set_filter(S.template[None, None, :], S.psd[None, None, :].astype(np.complex128), pre, 'AC')
    nxm.set_window(pre - 500, pre + 500)
    nxm.finalize(0)
    o2 = nxm.run(x[:, None, :]).cpu().numpy()
    oc, on = one.fit_offset(0, f_c), one.fit_offset(0, f_nd)
    scale = np.max(np.abs(o1[:, oc]))
    assert np.array_equal(o2[:, 2], o1[:, oc + 1])
    assert np.max(np.abs(o2[:, 3] - o1[:, oc])) < 1e-10 * scale
    assert np.allclose(o2[:, 1], o1[:, oc + 2], rtol=1e-9)
    assert np.max(np.abs(o2[:, 5] - o1[:, on])) < 1e-10 * scale
    assert np.allclose(o2[:, 4], o1[:, on + 2], rtol=1e-9)
    assert np.allclose(o2[:, 0], o1[:, one.chi0_offset(0)], rtol=1e-11)


def test_csd_parseval_and_psd_consistency_65536():
    """sum_k csd[a, a, k] df == mean square of channel a (Parseval); the diagonal equals the PSD estimator's result;
    csd[a, b] integrates to the mean product of the two channels."""
    from detprocess_b200.core.noise import NoiseCSD, NoisePSD
    n, fs = 65536, 1.25e6
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn((256, 2, n), generator=g, device='cuda', dtype=torch.float64) * 1e-10
    x[:, 1] += 0.4 * torch.roll(x[:, 0], 3, dims=-1)
    est = NoiseCSD(n, fs, 2)
    est.update(x)
    f, csd = est.finalize()
    df = fs / n
    xn = x.cpu().numpy()
    for a in range(2):
        assert np.sum(csd[a, a].real) * df == pytest.approx(np.mean(xn[:, a] ** 2), rel=1e-11)
        psd = NoisePSD(n, fs)
        psd.update(x[:, a].contiguous())
        assert np.allclose(csd[a, a].real, psd.finalize()[1], rtol=1e-11)
    assert np.sum(csd[0, 1]).real * df == pytest.approx(np.mean(xn[:, 0] * xn[:, 1]), rel=1e-9)
    assert abs(np.sum(csd[0, 1]).imag) < 1e-12 * abs(np.sum(csd[0, 0]).real)
