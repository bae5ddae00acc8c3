"""Host-batch pipeline depth vs end-to-end rate (development aid)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from detprocess_b200.synth import SynthSetup

S = SynthSetup(32768)
E = 4096
x = bench.make_device_traces(S, E, torch.device('cuda', 0), 1)
gain = 1e-11
hosts = {'f64': torch.empty((E, 32768), dtype=torch.float64).pin_memory(), 'i16': torch.empty((E, 32768), dtype=torch.int16).pin_memory()}
hosts['f64'].copy_(x.cpu())
hosts['i16'].copy_(torch.clamp(torch.round(x / gain), -32768, 32767).to(torch.int16).cpu())
for name, host in hosts.items():
    plan = bench.build_plan(S, 'f64', adc=(gain, 0.0) if name == 'i16' else None)
    hout = np.empty((E, plan.n_out))
    row = []
    for stages in (2, 4, 8, 16, 32):
        os.environ['DP_HOST_STAGES'] = str(stages)
        plan.run_host(host, hout)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            plan.run_host(host, hout)
        torch.cuda.synchronize()
        row.append(f'{stages}: {5 * E / (time.perf_counter() - t0) / 1e3:.0f}k')
    print(name, '  '.join(row), '(stages: events/s)', flush=True)
