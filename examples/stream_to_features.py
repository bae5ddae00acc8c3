"""
End-to-end example on synthetic data (needs a B200): continuous stream -> optimal-filter trigger ->
OF features of the triggered windows (no trace copy) -> the YAML-driven feature table.

    python examples/stream_to_features.py

Mirrors the reference workflow: OptimumFilterTrigger.update_trace / find_triggers (detprocess/core/oftrigger.py),
then FeatureProcessing.process with a trigger dataframe (detprocess/process/features.py).
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from detprocess_b200.core.filterdata import FilterData            # noqa: E402
from detprocess_b200.core.oftrigger import OptimumFilterTrigger   # noqa: E402
from detprocess_b200.core.plans import OFPlan                      # noqa: E402
from detprocess_b200.process.features import FeatureProcessing    # noqa: E402
from detprocess_b200.synth import SynthSetup, make_continuous     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    S = SynthSetup(32768)
    pre, fs, n = S.nb_pretrigger, S.fs, S.nb_samples
    L = 12_500_000                                   # 10 s at 1.25 MHz
    stream, t_true, a_true = make_continuous(L, S.template, S.psd, fs, np.random.default_rng(1), pulse_rate_hz=5.0,
                                             amp_range=(5e-8, 2e-7), return_truth=True)
    xs = torch.from_numpy(stream).cuda()

    # 1) trigger (same constructor arguments as the reference class)
    trig = OptimumFilterTrigger('chan1', fs, S.template, S.psd, pre, max_samples=L)
    trig.update_trace(xs)
    t0 = time.perf_counter()
    data = trig.find_triggers_once(thresh=10.0, pileup_window_msec=2.0)['chan1']
    torch.cuda.synchronize()
    idx = np.asarray(data['trigger_index'])
    print(f'{len(idx)} triggers in {L / fs:.0f} s of data ({1e3 * (time.perf_counter() - t0):.1f} ms incl. host); '
          f'{len(t_true)} pulses injected')

    # 2) OF features of the triggered windows, read straight from the stream
    plan = OFPlan(n, fs, 1, 'f64')
    plan.set_psd(0, S.psd, 'AC')
    f_con = plan.add_fit(0, plan.add_template(0, S.template, pre), pre - 500, pre + 500)
    plan.finalize()
    feats = plan.run_windows(xs, torch.from_numpy(idx - pre).cuda()).cpu().numpy()
    off = plan.fit_offset(0, f_con)
    ok = feats[:, 0] != -999999.0
    print('OF amplitude / trigger amplitude (first 5):', (feats[ok, off] / np.asarray(data['trigger_amplitude'])[ok])[:5])

    # 3) the YAML-driven feature table on the same windows (what FeatureProcessing does with a trigger dataframe)
    starts = idx[ok] - pre
    windows = torch.stack([xs[s:s + n] for s in starts])[:, None, :]
    fd = FilterData()
    fd.set_psd('chan1', S.psd, sample_rate=fs)
    fd.set_template('chan1', S.template, sample_rate=fs, pretrigger_length_samples=pre)
    fd.set_template('chan1', S.template_glitch, sample_rate=fs, pretrigger_length_samples=pre, tag='glitch')
    fp = FeatureProcessing({'traces': windows, 'channels': ['chan1'], 'sample_rate': fs,
                            'admin': {'trigger_index': idx[ok]}},
                           os.path.join(HERE, 'process_example.yaml'), filter_data=fd, verbose=False)
    df = fp.process()
    print(df[['trigger_index', 'amp_of1x1_constrained_chan1', 't0_of1x1_constrained_chan1', 'chi2_of1x1_constrained_chan1',
              'amp_of1x1_glitch_chan1', 'baseline_chan1', 'integral_chan1']].head())


if __name__ == '__main__':
    main()
