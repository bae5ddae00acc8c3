"""GPU parity of the noise PSD accumulation (C5 row a13) against oracle/psd.py, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from oracle import psd as P  # noqa: E402


def _traces(n_tr, n, seed, offset=0.0):
    rng = np.random.default_rng(seed)
    # coloured: white + random walk component + a line
    w = rng.standard_normal((n_tr, n)) * 2e-11
    t = np.arange(n)
    line = 5e-11 * np.sin(2 * np.pi * 60.0 * t / n + rng.uniform(0, 6.28, (n_tr, 1)))
    return w + line + offset


@pytest.mark.parametrize('n,prec,tol', [(65536, 'f64', 1e-11), (32768, 'f64', 1e-11), (16384, 'f64', 1e-11),
                                        (4096, 'f64', 1e-11), (65536, 'f32', 2e-5), (32768, 'f32', 2e-5),
                                        (16384, 'f32', 2e-5), (8192, 'f32', 2e-5)])
def test_psd_matches_oracle(n, prec, tol):
    from detprocess_b200.core.noise import NoisePSD
    fs = 1.25e6
    tr = _traces(300, n, 7, offset=(3e-9 if prec == 'f64' else 3e-9))
    cut = np.random.default_rng(8).random(300) < 0.8
    dev = torch.device('cuda', 0)
    est = NoisePSD(n, fs, precision=prec, device=dev, typical_rms=2e-11)
    x = torch.from_numpy(tr).to(dev)
    c = torch.from_numpy(cut).to(dev)
    est.update(x[:128], c[:128])
    est.update(x[128:], c[128:])
    freqs, psd = est.finalize()
    f0, p0 = P.calc_psd(tr, fs, cut)
    assert est.count == int(cut.sum())
    assert np.array_equal(freqs, f0)
    rel = np.abs(psd / p0 - 1)
    assert rel.max() < tol, (rel.max(), int(rel.argmax()))


def test_psd_offset_and_reset():
    from detprocess_b200.core.noise import NoisePSD
    n, fs = 4096, 1.25e6
    tr = _traces(50, n, 9, offset=1e-9)
    dev = torch.device('cuda', 0)
    est = NoisePSD(n, fs, device=dev)
    est.update(torch.from_numpy(tr).to(dev), None, with_offset=True)
    _, psd = est.finalize()
    assert np.isclose(est.offset, P.offset(tr), rtol=1e-12)
    est.plan.reset()
    est.update(torch.from_numpy(tr[:10]).to(dev))
    _, psd2 = est.finalize()
    _, p0 = P.calc_psd(tr[:10], fs)
    assert np.abs(psd2 / p0 - 1).max() < 1e-11


def test_psd_int16_and_float32_traces():
    """Raw ADC (int16) and float32 buffers give the PSD of the same numbers (exactly representable samples)."""
    from detprocess_b200.core.noise import NoisePSD
    n, fs = 16384, 1.25e6
    adc = np.random.default_rng(10).integers(-2000, 2000, size=(64, n)).astype(np.int16)
    dev = torch.device('cuda', 0)
    ref = NoisePSD(n, fs, device=dev)
    ref.update(torch.from_numpy(adc.astype(np.float64)).to(dev))
    _, p_ref = ref.finalize()
    _, p0 = P.calc_psd(adc.astype(np.float64), fs)
    assert np.abs(p_ref / p0 - 1).max() < 1e-11
    for dt in (torch.int16, torch.float32):
        est = NoisePSD(n, fs, device=dev)
        est.update(torch.from_numpy(adc).to(dev).to(dt))
        _, p = est.finalize()
        assert np.array_equal(p, p_ref)


@pytest.mark.parametrize('precision,nb_samples,n_chan', [('f64', 16384, 2), ('f32', 32768, 3), ('f64', 65536, 2), ('f32', 16384, 4),
                                                         ('f64', 32768, 4)])
def test_csd_parity(precision, nb_samples, n_chan):
    """dp_csd_accumulate / dp_csd_get_sums == oracle calc_csd (Noise.calc_csd, noise.py:374-470): full two-sided
    [n, n, N] array, event mask, two accumulate calls."""
    import torch
    from detprocess_b200.core.noise import NoiseCSD
    from detprocess_b200.synth import SynthNxM
    from oracle.psd import calc_csd
    S = SynthNxM(nb_samples, n_chan, 1)
    x = S.traces(96, np.random.default_rng(9), pulse_fraction=0.0) + 3e-8     # DC offset: exercises the DC bin
    cut = np.arange(96) % 5 != 0
    est = NoiseCSD(nb_samples, S.fs, n_chan, precision=precision, typical_rms=1e-8)
    xd = torch.from_numpy(x).cuda()
    cd = torch.from_numpy(cut).cuda()
    est.update(xd[:40], cd[:40])
    est.update(xd[40:], cd[40:])
    f, csd = est.finalize()
    fo, ref = calc_csd(x, S.fs, cut)
    assert est.count == int(cut.sum())
    assert np.array_equal(f, fo)
    tol = 1e-11 if precision == 'f64' else 3e-5
    scale = np.sqrt(np.abs(ref[np.arange(n_chan), np.arange(n_chan)])[:, None, :] * np.abs(ref[np.arange(n_chan), np.arange(n_chan)])[None, :, :])
    assert np.max(np.abs(csd - ref) / scale) < tol
    # its inverse drives the NxM filter: the estimate is a valid (hermitian, positive) csd at every bin
    assert np.allclose(csd, np.conj(np.transpose(csd, (1, 0, 2))))


def test_psd_and_csd_from_a_raw_adc_file(tmp_path):
    """Noise randoms stored as int16 ADC counts in the raw-binary container -> PSD / CSD estimators through the reader
    interface == the oracle on the host-converted traces."""
    from detprocess_b200.core.noise import calc_psd_from_reader, calc_csd_from_reader
    from detprocess_b200.io import RawBinaryReader, write_raw_binary
    from detprocess_b200.synth import SynthNxM
    from oracle.psd import calc_csd
    S = SynthNxM(16384, 2, 1)
    x = S.traces(150, np.random.default_rng(13), pulse_fraction=0.0)
    gain, off = [2.0e-12, 3.0e-12], [1.0e-9, -4.0e-9]
    adc = np.stack([np.clip(np.round((x[:, c] - off[c]) / gain[c]), -32768, 32767) for c in range(2)], axis=1).astype(np.int16)
    base = str(tmp_path / 'randoms')
    write_raw_binary(base, adc, ['chanA', 'chanB'], S.fs, adc_gain=gain, adc_offset=off)
    conv = np.stack([adc[:, c].astype(np.float64) * gain[c] + off[c] for c in range(2)], axis=1)
    cut = np.arange(150) % 7 != 0
    r = RawBinaryReader(base)
    f, psd = calc_psd_from_reader(r, 'chanB', cut=cut, batch=64)
    assert np.allclose(psd, P.calc_psd(conv[:, 1], S.fs, cut)[1], rtol=1e-11)
    f, csd = calc_csd_from_reader(r, 'chanA|chanB', cut=cut, batch=64)
    ref = calc_csd(conv, S.fs, cut)[1]
    scale = np.sqrt(np.abs(ref[0, 0]) * np.abs(ref[1, 1]))
    assert np.max(np.abs(csd - ref) / scale[None, None, :]) < 1e-11
