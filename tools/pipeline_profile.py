"""Where the time of FeatureProcessing.process goes (host profile + wall clock per batch size).
Usage: python tools/pipeline_profile.py [int16|float64] [batch_size]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from detprocess_b200.core.filterdata import FilterData
from detprocess_b200.io import ArrayReader
from detprocess_b200.process.features import FeatureProcessing
from detprocess_b200.synth import SynthSetup
from bench import make_device_traces

kind = sys.argv[1] if len(sys.argv) > 1 else 'int16'
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 512
N, FS, E = 32768, 1.25e6, int(os.environ.get("PP_EVENTS", "4096"))
dev = torch.device('cuda', 0)
S = SynthSetup(N, FS)
gain = 1.0e-11
xa = make_device_traces(S, E, dev, 779)
if kind == 'int16':
    host = torch.empty((E, 1, N), dtype=torch.int16).pin_memory()
    host[:, 0].copy_(torch.clamp(torch.round(xa / gain), -32768, 32767).to(torch.int16).cpu())
    kw = {'adc_gain': [gain], 'adc_offset': [0.0]}
else:
    host = torch.empty((E, 1, N), dtype=torch.float64).pin_memory()
    host[:, 0].copy_(xa.cpu())
    kw = {}
del xa
fd = FilterData()
fd.set_psd('chan1', S.psd, sample_rate=FS)
fd.set_template('chan1', S.template, sample_rate=FS, pretrigger_length_samples=S.nb_pretrigger)
fd.set_template('chan1', S.template_glitch, sample_rate=FS, pretrigger_length_samples=S.nb_pretrigger, tag='glitch')
with tempfile.TemporaryDirectory() as td:
    yml = os.path.join(td, 'c2.yaml')
    with open(yml, 'w') as f:
        f.write('global:\n    trace_length_samples: %d\n    pretrigger_length_samples: %d\n' % (N, S.nb_pretrigger)
                + 'chan1:\n'
                + '    of1x1_constrained:\n        run: True\n        template_tag: default\n'
                + '        window_min_from_trig_usec: -400\n        window_max_from_trig_usec: 400\n'
                + '    of1x1_glitch:\n        run: True\n        base_algorithm: of1x1_constrained\n        template_tag: glitch\n'
                + '        window_min_from_trig_usec: -400\n        window_max_from_trig_usec: 400\n'
                + '    baseline:\n        run: True\n        window_min_from_start_usec: 0\n        window_max_from_trig_usec: -1000\n'
                + '    integral:\n        run: True\n        window_min_from_trig_usec: -500\n        window_max_from_trig_usec: 500\n')
    fp = FeatureProcessing(ArrayReader(host, ['chan1'], FS, **kw), yml, filter_data=fd, verbose=False)
    fp.process(batch_size=bs, gather=False)
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        df = fp.process(batch_size=bs, gather=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'{kind} batch {bs}: {E / dt:.0f} events/s ({dt * 1e3:.1f} ms, {df.shape[1]} columns)')
    pr = cProfile.Profile()
    pr.enable()
    fp.process(batch_size=bs, gather=False)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
