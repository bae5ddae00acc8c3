"""OF kernel rate by trace storage type (development aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthSetup, make_traces
from detprocess_b200.core.plans import OFPlan

N, B = 32768, 8192
S = SynthSetup(N)
pre = S.nb_pretrigger
base = torch.from_numpy(make_traces(256, S.template, S.psd, S.fs, np.random.default_rng(1))).cuda()
x64 = base.repeat(B // 256, 1).contiguous()
gain = 1e-11
for prec in ('f32', 'f64'):
    for two in (False, True):
        plan = OFPlan(N, S.fs, 1, prec)
        plan.set_psd(0, S.psd)
        plan.set_adc_conversion(0, gain, 0.0)
        t = plan.add_template(0, S.template, pre)
        plan.add_fit(0, t, pre - 500, pre + 500)
        if two:
            plan.add_fit(0, plan.add_template(0, S.template_glitch, pre), pre - 500, pre + 500)
        plan.finalize()
        row = []
        for name, x in (('f64', x64), ('f32', x64.float()), ('i16', torch.clamp(torch.round(x64 / gain), -32768, 32767).to(torch.int16))):
            out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
            for _ in range(3):
                plan.run(x, out)
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                plan.run(x, out)
                ms.append(plan.last_kernel_ms())
            row.append(f'{name} {B / np.median(ms) / 1e3:.2f} M/s')
        print(f'{prec} two_templates={two}: ' + ', '.join(row), flush=True)
