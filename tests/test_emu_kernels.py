"""
The kernel SOURCE (detprocess_b200/csrc/*.cuh) compiled with g++ against a host-thread
CTA emulator (tests/emu): index maths, barrier placement, table layouts and the numpy
pairwise-summation order are checked here without a GPU.  The emulator is test
infrastructure only -- the product has no CPU path.
"""
import os
import struct
import subprocess
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'emu'))
import run_emu  # noqa: E402

from detprocess_b200.synth import SynthSetup, make_traces  # noqa: E402
from oracle.of1x1 import of1x1_batch  # noqa: E402
from oracle import reductions as R  # noqa: E402


def _check(out, o, fits_off, tol_amp, tol_chi2, tol_low):
    for iw, off in enumerate(fits_off):
        b = out[:, off:off + 5]
        assert np.array_equal(b[:, 1].astype(np.int64), o['ind'][iw])
        assert np.max(np.abs(b[:, 0] / o['amp'][iw] - 1)) < tol_amp
        assert np.max(np.abs(b[:, 2] / o['chi2'][iw] - 1)) < tol_chi2
        assert np.max(np.abs(b[:, 3] / o['lowchi2'][iw] - 1)) < tol_low
        assert np.max(np.abs(b[:, 4] / o['timeres'][iw] - 1)) < 4 * tol_amp


@pytest.mark.parametrize('nb_samples,precision,force_p2', [
    (2048, 'f64', False), (4096, 'f32', False), (8192, 'f64', False),
    (4096, 'f64', True), (8192, 'f32', True)])
def test_of_kernel_emulated(nb_samples, precision, force_p2):
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    tr = make_traces(3, S.template, S.psd, S.fs, np.random.default_rng(5),
                     offset=(1e-6 if precision == 'f64' else 0.0), amp_max=2e-7)
    w_def = [(None, None, False), (pre - 500, pre + 500, False), (pre, pre + 1, False)]
    w_gl = [(pre - 100, pre + 300, True)]
    fits = [(0, 0 if w[0] is None else w[0], nb_samples if w[1] is None else w[1], int(w[2])) for w in w_def]
    fits += [(1, w[0], w[1], int(w[2])) for w in w_gl]
    out = run_emu.run(tr, S.psd, [(S.template, pre, False), (S.template_glitch, pre, False)], fits, S.fs,
                      precision=precision, subtract_first=(precision == 'f32'),
                      scale=(2.0 ** 26 if precision == 'f32' else 1.0), force_p2=force_p2)
    o1 = of1x1_batch(tr, S.template, S.psd, S.fs, pre, windows=w_def)
    o2 = of1x1_batch(tr, S.template_glitch, S.psd, S.fs, pre, windows=w_gl)
    tol = (1e-11, 1e-11, 1e-11) if precision == 'f64' else (1e-5, 1e-4, 1e-4)
    assert np.max(np.abs(out[:, 0] / o1['chi0'] - 1)) < tol[1]
    _check(out, o1, [1, 6, 11], *tol)
    _check(out, o2, [16], *tol)


@pytest.mark.parametrize('nb_samples,precision', [
    (16384, 'f64'), (16384, 'f32'), (32768, 'f32'), (32768, 'f64'), (65536, 'f32'), (65536, 'f64')])
def test_of_v2_kernel_emulated(nb_samples, precision):
    """v2 kernels (dp_of2_kernel.cuh): packed-fp32 / fp64, 1, 2 and 4 phases, two templates,
    unconstrained + constrained + nodelay + outside-window fits."""
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    tr = make_traces(3, S.template, S.psd, S.fs, np.random.default_rng(5),
                     offset=(1e-6 if precision == 'f64' else 0.0), amp_max=2e-7)
    tr[2] = 0.0     # all-zero trace: every delay ties, numpy argmin takes the first candidate
    w_def = [(None, None, False), (pre - 500, pre + 500, False), (pre, pre + 1, False)]
    w_gl = [(pre - 100, pre + 300, True)]
    fits = [(0, 0 if w[0] is None else w[0], nb_samples if w[1] is None else w[1], int(w[2])) for w in w_def]
    fits += [(1, w[0], w[1], int(w[2])) for w in w_gl]
    out = run_emu.run(tr, S.psd, [(S.template, pre, False), (S.template_glitch, pre, False)], fits, S.fs,
                      precision=precision, subtract_first=(precision == 'f32'),
                      scale=(2.0 ** 26 if precision == 'f32' else 1.0), v2=True)
    o1 = of1x1_batch(tr[:2], S.template, S.psd, S.fs, pre, windows=w_def)
    o2 = of1x1_batch(tr[:2], S.template_glitch, S.psd, S.fs, pre, windows=w_gl)
    # tie-breaking on the all-zero trace: first candidate index of each window
    assert list(out[2, [2, 7, 12, 17]].astype(int)) == [0, pre - 500, pre, 0]
    assert np.all(out[2, [1, 6, 11, 16]] == 0.0)
    tol = (1e-11, 1e-11, 1e-11) if precision == 'f64' else (1e-5, 1e-4, 1e-4)
    out = out[:2]
    assert np.max(np.abs(out[:, 0] / o1['chi0'] - 1)) < tol[1]
    _check(out, o1, [1, 6, 11], *tol)
    _check(out, o2, [16], *tol)


@pytest.mark.parametrize('nb_samples,precision', [(16384, 'f32'), (32768, 'f64'), (65536, 'f32')])
def test_of_v2_kernel_emulated_constrained_only(nb_samples, precision):
    """C2 shape (only constrained fits): exercises the pruned inverse (pass 2' evaluated for the one or two
    outputs a +-500-sample window needs, pass 1' for the columns it touches)."""
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    tr = make_traces(3, S.template, S.psd, S.fs, np.random.default_rng(6),
                     offset=(1e-6 if precision == 'f64' else 0.0), amp_max=2e-7)
    w_def = [(pre - 500, pre + 500, False), (pre, pre + 1, False)]
    w_gl = [(pre - 300, pre + 200, False)]
    fits = [(0, w[0], w[1], 0) for w in w_def] + [(1, w[0], w[1], 0) for w in w_gl]
    out = run_emu.run(tr, S.psd, [(S.template, pre, False), (S.template_glitch, pre, False)], fits, S.fs,
                      precision=precision, subtract_first=(precision == 'f32'),
                      scale=(2.0 ** 26 if precision == 'f32' else 1.0), v2=True)
    o1 = of1x1_batch(tr, S.template, S.psd, S.fs, pre, windows=w_def)
    o2 = of1x1_batch(tr, S.template_glitch, S.psd, S.fs, pre, windows=w_gl)
    tol = (1e-11, 1e-11, 1e-11) if precision == 'f64' else (1e-5, 1e-4, 1e-4)
    _check(out, o1, [1, 6], *tol)
    _check(out, o2, [11], *tol)


@pytest.mark.parametrize('nb_samples,precision', [(1000, 'f64'), (2500, 'f64'), (1200, 'f32'), (25000, 'f64')])
def test_of_mixed_radix_kernel_emulated(nb_samples, precision):
    """mixed-radix kernel (dp_ofg_kernel.cuh) for trace lengths that are not powers of two -- 25000 = 2^3 5^5 is the
    reference's example configuration (process_example.yaml:93-94): radix 2/3/4/5 passes, digit-reversed pair tables,
    Bluestein template spectrum on the host, two templates, all window kinds"""
    S = SynthSetup(nb_samples)
    pre = S.nb_pretrigger
    w = max(20, nb_samples // 60)
    tr = make_traces(3, S.template, S.psd, S.fs, np.random.default_rng(7), offset=(1e-6 if precision == 'f64' else 0.0),
                     amp_max=2e-7, max_delay=w // 2)
    w_def = [(None, None, False), (pre - w, pre + w, False), (pre, pre + 1, False)]
    w_gl = [(pre - w // 3, pre + w, True)]
    fits = [(0, 0 if x[0] is None else x[0], nb_samples if x[1] is None else x[1], int(x[2])) for x in w_def]
    fits += [(1, x[0], x[1], int(x[2])) for x in w_gl]
    fcut = 10000.0 if nb_samples >= 8000 else 40000.0
    out = run_emu.run(tr, S.psd, [(S.template, pre, False), (S.template_glitch, pre, False)], fits, S.fs, fcut=fcut,
                      precision=precision, subtract_first=(precision == 'f32'),
                      scale=(2.0 ** 26 if precision == 'f32' else 1.0), generic=True)
    o1 = of1x1_batch(tr, S.template, S.psd, S.fs, pre, windows=w_def, lowchi2_fcutoff=fcut)
    o2 = of1x1_batch(tr, S.template_glitch, S.psd, S.fs, pre, windows=w_gl, lowchi2_fcutoff=fcut)
    tol = (1e-10, 1e-10, 1e-10) if precision == 'f64' else (1e-5, 1e-4, 1e-4)
    assert np.max(np.abs(out[:, 0] / o1['chi0'] - 1)) < tol[1]
    _check(out, o1, [1, 6, 11], *tol)
    _check(out, o2, [16], *tol)


def test_of_v2_kernel_emulated_address_sanitizer():
    """The kernel source under AddressSanitizer (compute-sanitizer is closed on the GPU pool): every shared-memory,
    scratch and table access of a two-phase, two-template fp64 run stays inside its buffer."""
    S = SynthSetup(32768)
    pre = S.nb_pretrigger
    tr = make_traces(2, S.template, S.psd, S.fs, np.random.default_rng(8))
    fits = [(0, 0, S.nb_samples, 0), (0, pre - 500, pre + 500, 0), (1, pre - 100, pre + 300, 1)]
    out = run_emu.run(tr, S.psd, [(S.template, pre, False), (S.template_glitch, pre, False)], fits, S.fs,
                      precision='f64', v2=True, asan=True)     # a finding aborts the emulator: CalledProcessError
    o1 = of1x1_batch(tr, S.template, S.psd, S.fs, pre, windows=[(None, None, False)])
    assert np.array_equal(out[:, 2].astype(np.int64), o1['ind'][0])


@pytest.mark.parametrize('precision,nb_samples,n,m', [('f64', 32768, 2, 2), ('f32', 16384, 3, 2), ('f64', 16384, 1, 3)])
def test_nxm_kernel_emulated(precision, nb_samples, n, m):
    """The NxM kernel source on host threads == the NxM oracle (tables, channel / template bookkeeping, barriers)."""
    from detprocess_b200.synth import SynthNxM
    from oracle.ofnxm import ofnxm_setup, ofnxm_batch
    S = SynthNxM(nb_samples, n, m)
    pre = S.nb_pretrigger
    st = ofnxm_setup(S.templates, S.csd, S.fs, pre)
    x = S.traces(3, np.random.default_rng(1))
    tol = 1e-10 if precision == 'f64' else 2e-5
    for win in [(pre - 500, pre + 500, False), (None, None, False), (pre - 100, pre + 50, True)]:
        out = run_emu.run_nxm(x, S.templates, S.csd, S.fs, pre, win, precision=precision)
        o = ofnxm_batch(x, st, win)
        assert np.array_equal(out[:, 2].astype(np.int64), o['ind'])
        assert np.max(np.abs(out[:, 0] / o['chi0'] - 1)) < tol
        assert np.max(np.abs(out[:, 1] / o['chi2'] - 1)) < tol * 10
        assert np.max(np.abs(out[:, 3:3 + m] - o['amps'])) < tol * np.max(np.abs(o['amps']))
        assert np.max(np.abs(out[:, 3 + m] / o['chi2_0'] - 1)) < tol * 10
        assert np.max(np.abs(out[:, 4 + m:] - o['amps0'])) < tol * np.max(np.abs(o['amps0']))


def test_nxm_kernel_emulated_address_sanitizer():
    """The NxM kernel source under AddressSanitizer: every shared-memory, scratch (X columns, parked blocks, parked
    q~ series) and table access of a two-phase fp64 run with 3 channels x 2 templates stays inside its buffer."""
    from detprocess_b200.synth import SynthNxM
    from oracle.ofnxm import ofnxm_setup, ofnxm_batch
    S = SynthNxM(32768, 3, 2)
    pre = S.nb_pretrigger
    x = S.traces(3, np.random.default_rng(2))
    out = run_emu.run_nxm(x, S.templates, S.csd, S.fs, pre, (pre - 300, pre + 700, False), precision='f64', asan=True)
    o = ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, S.fs, pre), (pre - 300, pre + 700, False))
    assert np.array_equal(out[:, 2].astype(np.int64), o['ind'])
    out = run_emu.run_nxm(x, S.templates, S.csd, S.fs, pre, (None, None, False), precision='f32', asan=True)
    assert np.mean(out[:, 2].astype(np.int64) == ofnxm_batch(x, ofnxm_setup(S.templates, S.csd, S.fs, pre))['ind']) == 1.0


@pytest.mark.parametrize('precision,nb_samples,n', [('f64', 32768, 2), ('f32', 16384, 3)])
def test_csd_kernel_emulated(precision, nb_samples, n):
    """The noise-CSD kernel source on host threads (under AddressSanitizer) + the host fold == oracle calc_csd."""
    from detprocess_b200.core.noise import csd_from_sums
    from detprocess_b200.synth import SynthNxM
    from oracle.psd import calc_csd
    S = SynthNxM(nb_samples, n, 1)
    x = S.traces(5, np.random.default_rng(4), pulse_fraction=0.0) + 2e-8
    mask = np.array([1, 1, 0, 1, 1], dtype=np.uint8)
    sums, count = run_emu.run_csd(x, S.fs, mask, precision=precision, asan=True)
    assert count == 4
    csd = csd_from_sums(sums, count, n, nb_samples, S.fs)
    ref = calc_csd(x, S.fs, mask)[1]
    d = np.arange(n)
    scale = np.sqrt(np.abs(ref[d, d])[:, None, :] * np.abs(ref[d, d])[None, :, :])
    assert np.max(np.abs(csd - ref) / scale) < (1e-11 if precision == 'f64' else 3e-5)


def test_reduce_kernel_emulated_bit_exact():
    exe = os.path.join(HERE, 'emu', '_build', 'emu_reduce')
    src = os.path.join(HERE, 'emu', 'emu_reduce.cpp')
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(['g++', '-std=c++20', '-O1', '-pthread', '-o', exe, src])
    rng = np.random.default_rng(3)
    n, nev, fs = 8192, 3, 1.25e6
    tr = rng.standard_normal((nev, n)) * 1e-8 + 3e-7
    tr[1, 500] = np.nan
    feats = [(0, 0, 3000), (1, 3500, 4750), (2, 0, n - 1), (3, 0, n - 1), (0, 100, 105), (1, 7, 8), (0, 3, 3 + 129),
             (1, 1, 1 + 1025), (0, 0, n), (2, 501, 600), (0, 17, 17 + 8), (1, 40, 40 + 10), (0, 9, 9),
             (3, 501, 600), (3, 10, 20), (2, 490, 510), (3, 0, n - 1), (3, 490, 510), (2, 490, 510), (3, 490, 510)]
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, 'i.bin'), os.path.join(td, 'o.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('<3i', n, nev, len(feats)))
            f.write(struct.pack('<d', fs))
            for ft in feats:
                f.write(struct.pack('<3i', *ft))
            f.write(tr.tobytes())
        subprocess.check_call([exe, fin, fout])
        out = np.fromfile(fout).reshape(nev, len(feats))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for i, (op, a, b) in enumerate(feats):
            ref = [R.baseline_batch, lambda t, a, b: R.integral_batch(t, fs, a, b), R.maximum_batch,
                   R.minimum_batch][op](tr, a, b)
            nan = np.isnan(ref)
            assert np.array_equal(np.isnan(out[:, i]), nan)
            assert np.array_equal(out[:, i][~nan].view(np.uint64), ref[~nan].view(np.uint64)), (op, a, b)
