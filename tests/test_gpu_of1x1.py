"""
GPU parity of the fused OF1x1 kernel (through the C ABI) against the float64 CPU
oracle on the same seeded synthetic traces.

Tolerances are BASELINE.json's: fp64 mode amp/chi2 1e-9 relative with identical t0
sample index; fp32 fast mode amp 1e-5, chi2 1e-4 relative with t0 identical except
on near-ties (checked: when the index differs the oracle's own chi2 at the two
indices must agree to fp32 resolution).
"""
import numpy as np
import pytest

from detprocess_b200.synth import SynthSetup, make_traces
from oracle.of1x1 import of1x1_batch

pytestmark = pytest.mark.gpu

TOL = {'f64': dict(amp=1e-9, chi2=1e-9, low=1e-9, tres=1e-9),
       'f32': dict(amp=1e-5, chi2=1e-4, low=1e-3, tres=1e-5)}


def _windows(S):
    pre = S.nb_pretrigger
    return [(None, None, False),                 # unconstrained
            (pre - 500, pre + 500, False),       # constrained +-400us @1.25MHz
            (pre, pre + 1, False),               # no delay
            (pre - 100, pre + 300, True)]        # outside window


def _run_plan(S, traces, precision, templates, windows_per_template, in_dtype=None, fcut=10000.0):
    import torch
    from detprocess_b200.core.plans import OFPlan
    plan = OFPlan(S.nb_samples, S.fs, 1, precision)
    plan.set_psd(0, S.psd, 'AC')
    plan.set_lowchi2_fcutoff(fcut)
    fits = []
    for tpl, wins in zip(templates, windows_per_template):
        ti = plan.add_template(0, tpl, S.nb_pretrigger)
        for (lo, hi, outside) in wins:
            fits.append((ti, plan.add_fit(0, ti, lo, hi, outside)))
    plan.finalize()
    x = torch.from_numpy(traces)
    if in_dtype is not None:
        x = x.to(in_dtype)
    out = plan.run(x.cuda())
    torch.cuda.synchronize()
    return plan, fits, out.cpu().numpy()


NEAR_TIES = []      # (test id, differing events, events) of every fp32 comparison: printed by the last test of the module


def _compare(plan, fit, out, o, iw, tol, amp_floor, chan=0, max_diff_frac=0.01):
    """GPU fit `fit` against window `iw` of the oracle result `o`.

    fp64: identical delay index, everything else to 1e-9.
    fp32 ("t0 index identical except on documented near-ties"): an event may come out at another index ONLY IF the
    float64 oracle, evaluated at that index, is within the fp32 amplitude tolerance of its own optimum,
        |chi2_oracle[i_gpu] - chi2_oracle[i_oracle]| <= 2 * tol_amp * max(amp^2, (5 ampres)^2) * norm
    (chi2 = chi0 - amp^2 norm, so this is |amp_gpu_candidate - amp_best| <~ tol_amp * |amp|): the two delays are a tie
    at the resolution the mode promises.  Such events stay in every check below, compared with the oracle AT THE
    GPU's index; their number is bounded and recorded."""
    off = plan.fit_offset(chan, fit)
    amp, ind, chi2, low, tres = (out[:, off + i] for i in range(5))
    ind = ind.astype(np.int64)
    same = ind == o['ind'][iw]
    ref = {k: np.array(o[k][iw], dtype=np.float64) for k in ('amp', 'chi2', 'lowchi2', 'timeres')}
    if tol is TOL['f64']:
        assert same.all(), 't0 index must be identical in fp64 mode'
    elif not same.all():
        d = ~same
        at = o['at'](ind)
        bound = 2.0 * tol['amp'] * np.maximum(ref['amp'] ** 2, amp_floor ** 2) * o['norm']
        dchi = np.abs(at['chi2'] - ref['chi2'])
        assert np.all(dchi[d] <= bound[d]), f'index differs on a non-tie: dchi2 {dchi[d].max():.3e} > {bound[d].min():.3e}'
        assert d.mean() <= max_diff_frac, f'{d.sum()} of {d.size} events on near-ties'
        for k in ref:
            ref[k][d] = at[k][d]
    if tol is not TOL['f64']:
        NEAR_TIES.append((int((~same).sum()), int(same.size)))
    denom = np.maximum(np.abs(ref['amp']), amp_floor)
    assert np.max(np.abs(amp - ref['amp']) / denom) < tol['amp']
    big = np.abs(ref['amp']) > amp_floor
    if big.any():       # pure relative error where the pulse stands above the noise (north_star: 1e-9 / 1e-5 relative)
        assert np.max(np.abs(amp / ref['amp'] - 1)[big]) < tol['amp']
        assert np.max(np.abs(tres / ref['timeres'] - 1)[big]) < max(tol['tres'], tol['amp'] * 2)
    assert np.max(np.abs(chi2 / ref['chi2'] - 1)) < tol['chi2']
    assert np.max(np.abs(low / ref['lowchi2'] - 1)) < tol['low']


@pytest.mark.parametrize('precision,nb_samples', [('f64', 2048), ('f64', 4096), ('f64', 16384), ('f64', 32768),
                                                  ('f64', 65536), ('f64', 8192),
                                                  ('f32', 2048), ('f32', 8192), ('f32', 16384), ('f32', 32768),
                                                  ('f32', 65536)])
def test_of1x1_parity_single_template(precision, nb_samples):
    S = SynthSetup(nb_samples)
    nev = 300
    traces = make_traces(nev, S.template, S.psd, S.fs, np.random.default_rng(12345), offset=2.5e-7)
    wins = _windows(S)
    plan, fits, out = _run_plan(S, traces, precision, [S.template], [wins])
    o = of1x1_batch(traces, S.template, S.psd, S.fs, S.nb_pretrigger, windows=wins)
    tol = TOL[precision]
    assert np.max(np.abs(out[:, plan.chi0_offset(0)] / o['chi0'] - 1)) < tol['chi2']
    for iw, (_, fit) in enumerate(fits):
        _compare(plan, fit, out, o, iw, tol, amp_floor=5 * o['ampres'])


@pytest.mark.parametrize('precision,nb_samples', [('f64', 16384), ('f64', 32768), ('f32', 32768), ('f32', 65536)])
def test_of1x1_parity_glitch_template_variant(precision, nb_samples):
    """C2 shape: constrained fit with the default and the glitch template on the same events."""
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    traces = make_traces(200, S.template, S.psd, S.fs, np.random.default_rng(12346))
    w_def = [(pre - 500, pre + 500, False), (pre, pre + 1, False)]
    w_gl = [(pre - 500, pre + 500, False)]
    plan, fits, out = _run_plan(S, traces, precision, [S.template, S.template_glitch], [w_def, w_gl])
    tol = TOL[precision]
    o1 = of1x1_batch(traces, S.template, S.psd, S.fs, pre, windows=w_def)
    o2 = of1x1_batch(traces, S.template_glitch, S.psd, S.fs, pre, windows=w_gl)
    _compare(plan, fits[0][1], out, o1, 0, tol, 5 * o1['ampres'])
    _compare(plan, fits[1][1], out, o1, 1, tol, 5 * o1['ampres'])
    _compare(plan, fits[2][1], out, o2, 0, tol, 5 * o2['ampres'])


def test_of1x1_known_answer_on_gpu():
    """Noiseless template*A shifted by d: amp == A, delay == d, chi2 ~ 0 (convention independent)."""
    S = SynthSetup(8192)
    pre = S.nb_pretrigger
    truth = [(3.0, 17), (-2.5, -211), (1e-7, 0), (7.0, 300)]
    traces = np.stack([A * np.roll(S.template, d) for A, d in truth])
    plan, fits, out = _run_plan(S, traces, 'f64', [S.template], [[(None, None, False)]])
    off = plan.fit_offset(0, fits[0][1])
    for i, (A, d) in enumerate(truth):
        assert out[i, off] == pytest.approx(A, rel=1e-10)
        assert int(out[i, off + 1]) - pre == d
        assert abs(out[i, off + 2]) < 1e-6 * out[i, plan.chi0_offset(0)]


def test_of1x1_float32_and_int16_inputs():
    """Same numbers whether the trace buffer is f64, f32 or i16 (values exactly representable)."""
    import torch
    S = SynthSetup(4096)
    rng = np.random.default_rng(5)
    adc = rng.integers(-3000, 3000, size=(64, S.nb_samples)).astype(np.int16)
    adc[:, :] += (800 * S.template[None, :]).astype(np.int16)
    psd = np.full(S.nb_samples, 1e-3)
    S.psd = psd
    wins = [(None, None, False)]
    ref = _run_plan(S, adc.astype(np.float64), 'f64', [S.template], [wins])[2]
    for dt in (torch.float32, torch.int16):
        got = _run_plan(S, adc.astype(np.float64), 'f64', [S.template], [wins], in_dtype=dt)[2]
        assert np.array_equal(got, ref)


@pytest.mark.parametrize('nb_samples', [32768, 4096, 8192])
@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_of1x1_adc_counts_with_channel_conversion(precision, nb_samples):
    """int16 ADC counts + per-channel gain / offset (set_adc_conversion) == the oracle on the traces converted on
    the host the way H5Reader(adctoamp=True) hands them to the reference (adc * gain + offset, float64).
    4096 / 8192 samples run on the first-generation kernel (dp_of_kernel), 32768 on dp_of2_kernel."""
    import torch
    from detprocess_b200.core.plans import OFPlan
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    gains, offs = (1.0e-11, 1.3e-11), (-3.0e-9, 1.7e-8)
    amps = [make_traces(40, S.template, S.psd, S.fs, np.random.default_rng(20 + c)) for c in range(2)]
    adc = np.stack([np.clip(np.round((amps[c] - offs[c]) / gains[c]), -32768, 32767).astype(np.int16) for c in range(2)], axis=1)
    conv = [adc[:, c].astype(np.float64) * gains[c] + offs[c] for c in range(2)]
    plan = OFPlan(S.nb_samples, S.fs, 2, precision)
    fits = []
    for c in range(2):
        plan.set_psd(c, S.psd)
        plan.set_adc_conversion(c, gains[c], offs[c])
        t = plan.add_template(c, S.template, pre)
        fits.append(plan.add_fit(c, t, pre - 500, pre + 500))
    plan.finalize()
    out = plan.run(torch.from_numpy(adc).cuda()).cpu().numpy()
    for c in range(2):
        o = of1x1_batch(conv[c], S.template, S.psd, S.fs, pre, windows=[(pre - 500, pre + 500, False)])
        _compare(plan, fits[c], out, o, 0, TOL[precision], 5 * o['ampres'], chan=c)


def test_of1x1_two_channels_sharded_equals_single():
    """[B, 2, N] batch: each channel has its own PSD/template; equals two 1-channel runs."""
    import torch
    from detprocess_b200.core.plans import OFPlan
    S = SynthSetup(4096)
    pre = S.nb_pretrigger
    tr0 = make_traces(50, S.template, S.psd, S.fs, np.random.default_rng(1))
    tr1 = make_traces(50, S.template_glitch, 2 * S.psd, S.fs, np.random.default_rng(2))
    both = np.stack([tr0, tr1], axis=1)
    plan = OFPlan(S.nb_samples, S.fs, 2, 'f64')
    plan.set_psd(0, S.psd)
    plan.set_psd(1, 2 * S.psd)
    t0 = plan.add_template(0, S.template, pre)
    t1 = plan.add_template(1, S.template_glitch, pre)
    f0 = plan.add_fit(0, t0, pre - 500, pre + 500)
    f1 = plan.add_fit(1, t1, None, None)
    plan.finalize()
    out = plan.run(torch.from_numpy(both).cuda()).cpu().numpy()
    o0 = of1x1_batch(tr0, S.template, S.psd, S.fs, pre, windows=[(pre - 500, pre + 500, False)])
    o1 = of1x1_batch(tr1, S.template_glitch, 2 * S.psd, S.fs, pre, windows=[(None, None, False)])
    _compare(plan, f0, out[:, :], o0, 0, TOL['f64'], 5 * o0['ampres'])
    off1 = plan.fit_offset(1, f1)
    assert np.array_equal(out[:, off1 + 1].astype(int), o1['ind'][0])
    assert np.max(np.abs(out[:, off1] - o1['amp'][0]) / np.maximum(np.abs(o1['amp'][0]), 5 * o1['ampres'])) < 1e-9


@pytest.mark.parametrize('precision', ['f64', 'f32'])
def test_of1x1_c3_shape_eight_channels_16384(precision):
    """BASELINE config C3: 8 channels x 16384 samples, every channel with its own PSD / templates / fits;
    float32 and int16 trace buffers give the same numbers as float64 ones (v2 kernels)."""
    import torch
    from detprocess_b200.core.plans import OFPlan
    n, nch, nev = 16384, 8, 40
    S = SynthSetup(n)
    pre = S.nb_pretrigger
    rng = np.random.default_rng(77)
    traces = np.stack([make_traces(nev, S.template if c % 2 == 0 else S.template_glitch, S.psd * (1 + c), S.fs, rng)
                       for c in range(nch)], axis=1)
    plan = OFPlan(n, S.fs, nch, precision)
    fits = []
    for c in range(nch):
        plan.set_psd(c, S.psd * (1 + c))
        t0 = plan.add_template(c, S.template, pre)
        f = [plan.add_fit(c, t0, pre - 500, pre + 500), plan.add_fit_nodelay(c, t0)]
        if c % 2:
            t1 = plan.add_template(c, S.template_glitch, pre)
            f.append(plan.add_fit(c, t1, None, None))
        fits.append(f)
    plan.finalize()
    out = plan.run(torch.from_numpy(traces).cuda()).cpu().numpy()
    tol = TOL[precision]
    for c in range(nch):
        o = of1x1_batch(traces[:, c], S.template, S.psd * (1 + c), S.fs, pre, windows=[(pre - 500, pre + 500, False), (pre, pre + 1, False)])
        assert np.max(np.abs(out[:, plan.chi0_offset(c)] / o['chi0'] - 1)) < tol['chi2']
        _compare(plan, fits[c][0], out, o, 0, tol, 5 * o['ampres'], chan=c)
        _compare(plan, fits[c][1], out, o, 1, tol, 5 * o['ampres'], chan=c)
        if c % 2:
            # unconstrained fit of the glitch template on default-template pulses: many shallow optima, the near-tie
            # rule is what decides (every differing index is checked against it), not a fraction of identical indices
            og = of1x1_batch(traces[:, c], S.template_glitch, S.psd * (1 + c), S.fs, pre, windows=[(None, None, False)])
            _compare(plan, fits[c][2], out, og, 0, tol, 5 * og['ampres'], chan=c, max_diff_frac=0.05)
    # exactly representable samples: f32 and i16 buffers == f64 buffer
    adc = rng.integers(-2000, 2000, size=(6, nch, n)).astype(np.int16)
    ref = plan.run(torch.from_numpy(adc.astype(np.float64)).cuda()).cpu().numpy()
    for dt in (torch.float32, torch.int16):
        got = plan.run(torch.from_numpy(adc.astype(np.float64)).to(dt).cuda()).cpu().numpy()
        assert np.array_equal(got, ref)


def test_of1x1_empty_and_ragged_batches():
    import torch
    S = SynthSetup(2048)
    traces = make_traces(149, S.template, S.psd, S.fs, np.random.default_rng(3))   # not a multiple of the grid
    plan, fits, out = _run_plan(S, traces, 'f64', [S.template], [[(None, None, False)]])
    o = of1x1_batch(traces, S.template, S.psd, S.fs, S.nb_pretrigger, windows=[(None, None, False)])
    _compare(plan, fits[0][1], out, o, 0, TOL['f64'], 5 * o['ampres'])
    empty = plan.run(torch.empty((0, S.nb_samples), dtype=torch.float64, device='cuda'))
    assert empty.shape == (0, plan.n_out)
    with pytest.raises(ValueError):
        plan.run(torch.zeros((4, S.nb_samples + 2), dtype=torch.float64, device='cuda'))


def test_host_buffer_path_matches_device_path():
    import torch
    S = SynthSetup(8192)
    traces = make_traces(700, S.template, S.psd, S.fs, np.random.default_rng(9))
    plan, fits, out = _run_plan(S, traces, 'f32', [S.template], [[(None, None, False)]])
    host = plan.run_host(traces)
    assert np.array_equal(host, out)
    pinned = torch.from_numpy(traces).pin_memory()
    assert np.array_equal(plan.run_host(pinned), out)


def test_zz_report_near_ties():
    """not a check: prints how many fp32 events of this module sat on a documented near-tie (see _compare)"""
    n = sum(a for a, _ in NEAR_TIES)
    tot = sum(b for _, b in NEAR_TIES)
    print(f'\nfp32 near-ties: {n} of {tot} event-fits came out at another delay index (each within the tie bound)')
