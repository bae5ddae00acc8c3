"""psd_amp on a B200 against the CPU oracle (development aid -- the band-amplitude kernel was written after the round-2 GPU
budget was spent and has only run on the host-thread emulator, tests/test_psd_amp.py):

    python tools/check_band_gpu.py
"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.core.algorithms import FeatureExtractors as FE
from detprocess_b200.core.plans import BandPlan
from detprocess_b200.synth import make_psd, make_template, make_traces
from oracle import psd as P

fs = 1.25e6
lims = [[45.0, 75.0], [300.0, 500.0], [350.0, 450.0], [150, 250], [250, 350], [2000.0, 9000.0], 1234.0]
for n in (25000, 32768, 4096):
    names, bins = FE._psd_amp_ranges(n, fs, lims)
    tr = make_traces(64, make_template(n, fs), make_psd(n, fs), fs, np.random.default_rng(2), offset=2e-7)
    x = torch.from_numpy(np.stack([tr, 2 * tr], axis=1)).cuda()          # [B, 2, N]: rows strided like a reader batch
    plan = BandPlan(n, fs, bins)
    for row, scale in ((0, 1.0), (1, 2.0)):
        out = plan.run(x[:, row, :]).cpu().numpy()
        ref = np.array([[P.psd_amp(scale * tr[i], fs, lims)[f'psd_amp_{nm}'] for nm in names] for i in range(len(tr))])
        print(f'N={n} row {row}: max rel err {np.max(np.abs(out / ref - 1)):.2e}')
        assert np.max(np.abs(out / ref - 1)) < 1e-9
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big = x[:, 0, :].repeat(32, 1).contiguous()
    plan.run(big)
    t0.record()
    plan.run(big)
    t1.record()
    torch.cuda.synchronize()
    print(f'N={n}: {big.shape[0] / t0.elapsed_time(t1) / 1e3:.2f} M events/s for {sum(b - a for a, b in bins)} bins')
print('psd_amp matches the oracle')
