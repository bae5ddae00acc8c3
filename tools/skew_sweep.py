"""Sweep the start skew of the OF kernel (development aid): DP2_SKEW_NS is read at every launch."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthSetup, make_traces
from detprocess_b200.core.plans import OFPlan


def plan_for(N, prec, two):
    S = SynthSetup(N)
    pre = S.nb_pretrigger
    plan = OFPlan(N, S.fs, 1, prec)
    plan.set_psd(0, S.psd)
    t = plan.add_template(0, S.template, pre)
    plan.add_fit(0, t, pre - 500, pre + 500)
    if two:
        g = plan.add_template(0, S.template_glitch, pre)
        plan.add_fit(0, g, pre - 500, pre + 500)
    else:
        plan.add_fit(0, t, None, None)
    plan.finalize()
    return S, plan


for N, prec, two, B in ((32768, 'f64', True, 8192), (32768, 'f32', True, 8192), (32768, 'f32', False, 8192), (32768, 'f64', False, 8192),
                        (16384, 'f32', True, 16384), (65536, 'f64', True, 4096)):
    S, plan = plan_for(N, prec, two)
    base = torch.from_numpy(make_traces(256, S.template, S.psd, S.fs, np.random.default_rng(1))).cuda()
    x = base.repeat((B + 255) // 256, 1)[:B].contiguous()
    out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
    row = []
    for ns in (0, 200, 400, 600, 800, 1200, 2000):
        os.environ['DP2_SKEW_NS'] = str(ns)
        for _ in range(2):
            plan.run(x, out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            plan.run(x, out)
            ms.append(plan.last_kernel_ms())
        row.append(f'{ns}:{B / np.median(ms) / 1e3:.3f}')
    print(f'N={N} {prec} two_templ={two}:  ' + '  '.join(row) + '  (ns: M ev/s)', flush=True)
