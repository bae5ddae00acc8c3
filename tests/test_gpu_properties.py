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
