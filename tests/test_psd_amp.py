"""psd_amp (reference core/algorithms.py:953-1042): oracle known answers, the product's host logic (frequency ranges -> bins,
feature names) and the band-amplitude kernel SOURCE run on the host-thread emulator (no GPU).  The kernel was written after
this round's GPU budget was spent: its B200 run is the first thing to do next round (tools/check_band_gpu.py)."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import psd as P

HERE = os.path.dirname(os.path.abspath(__file__))
F_LIMS = [[45.0, 75.0], [300.0, 500.0], [350.0, 450.0], [150, 250], [250, 350]]     # the reference's example YAML


def test_oracle_psd_amp_known_answers():
    fs, n = 1.25e6, 25000
    rng = np.random.default_rng(1)
    # The reference applies `* nbins / fs` to the OF-normalised spectrum fft(x) / N / df (algorithms.py:1006-1016), i.e. its
    # "psd" is the physical two-sided PSD times (N / fs)^2; restated as written.
    # white noise of variance s^2: physical level s^2 / fs  ->  sqrt(folded) ~ sqrt(2 s^2 / fs) * N / fs, averaged over many bins
    s = 3e-9
    x = rng.standard_normal(n) * s
    out = P.psd_amp(x, fs, [[1000.0, 400000.0]])
    assert abs(out['psd_amp_1000_400000'] / (np.sqrt(2 * s * s / fs) * (n / fs) * np.sqrt(np.pi) / 2) - 1) < 0.03   # E|z|, complex Gaussian
    # a sine on a bin centre: all of its power in that bin
    k, a = 12, 2e-8
    t = np.arange(n) / fs
    y = a * np.sin(2 * np.pi * k * (fs / n) * t)
    o = P.psd_amp(y, fs, [k * fs / n])
    assert abs(o[f'psd_amp_{round(k * fs / n)}'] / (np.sqrt(2 * (a / 2) ** 2 * n / fs) * (n / fs)) - 1) < 1e-9
    # names, ordering, duplicates (utils.py:437-470)
    r, names = P.cleanup_freq_ranges([[75.0, 45.0], 60, [-45.0, 75.0], [60.0]])
    assert names == ['45_75', '60'] and r == [[45.0, 75.0], [60]]


def test_product_ranges_equal_the_oracle():
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    from detprocess_b200.utils import utils
    for n, fs in ((25000, 1.25e6), (32768, 1.25e6), (4096, 625e3)):
        freqs = np.fft.rfftfreq(n, d=1.0 / fs)[1:]
        for lims in (F_LIMS, [[10.0, 20.0]], [60], [[fs / 2, fs / 2 - 1]], [[1e9, 2e9]]):
            a, na = utils.cleanup_freq_ranges(lims)
            b, nb = P.cleanup_freq_ranges(lims)
            assert (a, na) == (b, nb)
            assert utils.get_ind_freq_ranges(a, freqs) == P.get_ind_freq_ranges(b, freqs)
            names, bins = FE._psd_amp_ranges(n, fs, lims)
            assert names == nb and bins == [(lo + 1, hi + 1) for lo, hi in P.get_ind_freq_ranges(b, freqs)]
            assert all(1 <= lo < hi <= n // 2 + 1 for lo, hi in bins)
    with pytest.raises(ValueError):
        FE._psd_amp_ranges(4096, 1.25e6, [])


def _emulate(traces, fs, bins, in_dtype=0, gain=1.0, offset=0.0, stride=None):
    exe = os.path.join(HERE, 'emu', '_build', 'emu_band')
    src = os.path.join(HERE, 'emu', 'emu_band.cpp')
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(['g++', '-std=c++20', '-O1', '-pthread', '-ffp-contract=off', '-o', exe, src])
    nev, n = traces.shape
    stride = n if stride is None else stride
    buf = np.zeros((nev, stride), dtype=traces.dtype)
    buf[:, :n] = traces
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, 'i.bin'), os.path.join(td, 'o.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('<5i', n, nev, len(bins), in_dtype, stride))
            f.write(struct.pack('<3d', fs, gain, offset))
            f.write(np.asarray([b[0] for b in bins], dtype=np.int32).tobytes())
            f.write(np.asarray([b[1] for b in bins], dtype=np.int32).tobytes())
            f.write(buf.tobytes())
        subprocess.check_call([exe, fin, fout])
        return np.fromfile(fout).reshape(nev, len(bins))


@pytest.mark.parametrize('n', [25000, 4096])
def test_band_kernel_emulated_matches_the_oracle(n):
    """the kernel source on host threads: float64 traces, the example YAML's ranges + a wide band + Nyquist + a single bin"""
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    from detprocess_b200.synth import make_psd, make_template, make_traces
    fs = 1.25e6
    lims = F_LIMS + [[2000.0, 9000.0], [fs / 2, fs / 2 - 200.0], 1234.0]
    names, bins = FE._psd_amp_ranges(n, fs, lims)
    tr = make_traces(3, make_template(n, fs), make_psd(n, fs), fs, np.random.default_rng(2), offset=2e-7)
    out = _emulate(tr, fs, bins, stride=n + 8)          # rows further apart than one trace, like a reader batch
    for i in range(tr.shape[0]):
        ref = P.psd_amp(tr[i], fs, lims)
        for j, nm in enumerate(names):
            assert abs(out[i, j] / ref[f'psd_amp_{nm}'] - 1) < 1e-9, (i, nm)


def test_band_kernel_emulated_int16_and_float32():
    from detprocess_b200.core.algorithms import FeatureExtractors as FE
    n, fs = 4096, 1.25e6
    names, bins = FE._psd_amp_ranges(n, fs, [[300.0, 3000.0], 50000.0])
    rng = np.random.default_rng(4)
    adc = rng.integers(-3000, 3000, size=(2, n)).astype(np.int16)
    gain, offset = 3.1e-10, -2e-7
    out = _emulate(adc, fs, bins, in_dtype=2, gain=gain, offset=offset)
    amps = adc.astype(np.float64) * gain + offset
    for i in range(2):
        ref = P.psd_amp(amps[i], fs, [[300.0, 3000.0], 50000.0])
        assert all(abs(out[i, j] / ref[f'psd_amp_{nm}'] - 1) < 1e-9 for j, nm in enumerate(names))
    x32 = (rng.standard_normal((2, n)) * 1e-8).astype(np.float32)
    out = _emulate(x32, fs, bins, in_dtype=1)
    for i in range(2):
        ref = P.psd_amp(x32[i].astype(np.float64), fs, [[300.0, 3000.0], 50000.0])
        assert all(abs(out[i, j] / ref[f'psd_amp_{nm}'] - 1) < 1e-9 for j, nm in enumerate(names))


def test_yaml_psd_amp_blocks_become_band_jobs(tmp_path):
    """the pipeline resolves psd_amp blocks without a filter file (reference processing_data.py:288-289) into band jobs with the
    reference's column names <algorithm>_<range>_<feature_channel>; no device work happens before process()"""
    import textwrap
    import torch
    from detprocess_b200.process import FeatureProcessing
    n, fs = 4096, 1.25e6
    ev = np.zeros((4, 2, n))
    cfg = tmp_path / 'c.yaml'
    cfg.write_text(textwrap.dedent(f'''
        global:
            trace_length_samples: {n}
            pretrigger_length_samples: {n // 2}
        chanA:
            psd_amp:
                run: True
                f_lims: [[300.0, 3000.0], 50000.0]
            lines:
                run: True
                base_algorithm: psd_amp
                f_lims: [60]
    '''))
    fp = FeatureProcessing({'traces': torch.from_numpy(ev), 'channels': ['chanA', 'chanB'], 'sample_rate': fs}, str(cfg), verbose=False)
    jobs = fp._band_jobs
    assert [j['columns'] for j in jobs] == [['psd_amp_300_3000_chanA', 'psd_amp_50000_chanA'], ['lines_60_chanA']]
    freqs = np.fft.rfftfreq(n, d=1.0 / fs)[1:]
    ref = P.get_ind_freq_ranges(P.cleanup_freq_ranges([[300.0, 3000.0], 50000.0])[0], freqs)
    assert jobs[0]['bins'] == [(a + 1, b + 1) for a, b in ref] and jobs[0]['channel'] == 'chanA'


def test_fold_spectrum_against_qetpy_golden():
    """pins oracle/psd.py::fold_spectrum (and the recalled low-pass filter of the trigger's saturation test) to upstream QETpy
    where ``oracle/dump_golden.py`` has been run; skipped otherwise (QETpy is not installable in this image)"""
    path = os.path.join(HERE, 'golden', 'utils_qetpy.npz')
    if not os.path.exists(path):
        pytest.skip('tests/golden/utils_qetpy.npz is generated by oracle/dump_golden.py where QETpy is installed')
    g = np.load(path)
    for spec, f, fold in ((g['spec'], g['f_even'], g['fold_even']), (g['spec_odd'], g['f_odd'], g['fold_odd'])):
        fo, fold_o = P.fold_spectrum(spec, 1.25e6)
        assert np.allclose(fo, f) and np.allclose(fold_o, fold, rtol=1e-14)
    from scipy.signal import butter, filtfilt
    b, a = butter(1, 50e3 / (0.5 * 1.25e6))
    assert np.allclose(filtfilt(b, a, g['lpf_in'], padtype='even'), g['lpf_out'], rtol=1e-10, atol=1e-12)
