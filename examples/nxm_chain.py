"""
Multi-channel chain on synthetic data (needs a B200): noise randoms -> cross-spectral density -> NxM optimal filter ->
`ofnxm` features through the YAML-driven pipeline.

    python examples/nxm_chain.py

Mirrors the reference workflow: Noise.calc_csd (detprocess/core/noise.py:374) fills the filter data, FilterData hands
the [n, n, N] csd and the [n, m, N] templates of the joint channel 'chanA|chanB' to qp.OFBase
(process/processing_data.py:294-381) and FeatureExtractors.ofnxm (core/algorithms.py:141) runs per event.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from detprocess_b200.core.filterdata import FilterData            # noqa: E402
from detprocess_b200.core.noise import NoiseCSD                   # noqa: E402
from detprocess_b200.process.features import FeatureProcessing    # noqa: E402
from detprocess_b200.synth import SynthNxM                        # noqa: E402

YAML = '''
global:
    trace_length_samples: 32768
    pretrigger_length_samples: 16384
chanA|chanB:
    feature_channel: pair
    ofnxm:
        run: True
        window_min_from_trig_usec: -400
        window_max_from_trig_usec: 400
        template_tag: shared
        amplitude_names: [phonon, glitch]
'''


def main():
    S = SynthNxM(32768, 2, 2)
    fs, n, pre = S.fs, S.nb_samples, S.nb_pretrigger
    # 1) cross-spectral density from 2000 noise-only events (streamed in batches; all-reduced when run under torchrun)
    est = NoiseCSD(n, fs, 2)
    for seed in range(4):
        est.update(torch.from_numpy(S.traces(500, np.random.default_rng(seed), pulse_fraction=0.0)).cuda())
    _, csd = est.finalize()
    print(f'csd from {est.count} events; |csd_AB / sqrt(csd_AA csd_BB)| at 1 kHz = '
          f'{abs(csd[0, 1, 26]) / np.sqrt(csd[0, 0, 26].real * csd[1, 1, 26].real):.2f}')
    # 2) filter data of the joint channel, 3) YAML-driven features
    fd = FilterData()
    fd.set_csd('chanA|chanB', csd, sample_rate=fs)
    fd.set_template('chanA|chanB', S.templates, sample_rate=fs, pretrigger_length_samples=pre, tag='shared')
    x, amps, delays = S.traces(4096, np.random.default_rng(99), pulse_fraction=1.0, return_truth=True)
    cfg = os.path.join('/tmp', 'nxm_example.yaml')
    with open(cfg, 'w') as f:
        f.write(YAML)
    fp = FeatureProcessing({'traces': torch.from_numpy(x), 'channels': ['chanA', 'chanB'], 'sample_rate': fs}, cfg,
                           filter_data=fd, verbose=False)
    df = fp.process()
    print(df[['phonon_ofnxm_constrained_pair', 'glitch_ofnxm_constrained_pair', 't0_ofnxm_constrained_pair',
              'chi2_ofnxm_constrained_pair']].head())
    print('injected:', amps[:5], delays[:5] / fs)
    res = df['phonon_ofnxm_constrained_pair'].values - amps[:, 0]
    print(f'phonon amplitude residual: mean {res.mean():.2e} A, rms {res.std():.2e} A; chi2 / dof = '
          f'{df["chi2_ofnxm_constrained_pair"].mean() / (2 * n):.3f}')


if __name__ == '__main__':
    main()
