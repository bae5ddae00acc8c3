"""A/B timing of the fused OF kernel (development aid): python tools/ab_of.py [tag]
Select the library with DP_LIB_NAME=<file in detprocess_b200/_C>."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthSetup, make_traces
from detprocess_b200.core.plans import OFPlan

tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get('DP_LIB_NAME', 'default')


def time_of(N, prec, B, mode):
    S = SynthSetup(N)
    pre = S.nb_pretrigger
    plan = OFPlan(N, S.fs, 1, prec)
    plan.set_psd(0, S.psd)
    t = plan.add_template(0, S.template, pre)
    if mode == 'c1':
        plan.add_fit_nodelay(0, t)
        plan.add_fit(0, t, None, None)
    elif mode == 'c2':
        plan.add_fit(0, t, pre - 500, pre + 500)
        g = plan.add_template(0, S.template_glitch, pre)
        plan.add_fit(0, g, pre - 500, pre + 500)
    elif mode == 'c2s':   # single template, constrained
        plan.add_fit(0, t, pre - 500, pre + 500)
    plan.finalize()
    base = torch.from_numpy(make_traces(256, S.template, S.psd, S.fs, np.random.default_rng(1))).cuda()
    x = base.repeat((B + 255) // 256, 1)[:B].contiguous()
    out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
    for _ in range(3):
        plan.run(x, out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        plan.run(x, out)
        ms.append(plan.last_kernel_ms())
    m = float(np.median(ms))
    print(f'[{tag}] OF N={N} {prec} B={B} {mode}: {m:.3f} ms  {B / (m * 1e-3) / 1e6:.3f} Mev/s  checksum {float(out[:, 1].sum()):.6e}', flush=True)


cfgs = [(32768, 'f64', 16384, 'c2'), (32768, 'f32', 16384, 'c2'), (32768, 'f64', 16384, 'c1'), (32768, 'f32', 16384, 'c1'),
        (32768, 'f64', 16384, 'c2s'), (16384, 'f64', 16384, 'c2'), (16384, 'f32', 16384, 'c2'), (65536, 'f64', 8192, 'c2'),
        (65536, 'f32', 8192, 'c2')]
if len(sys.argv) > 2:
    cfgs = cfgs[:int(sys.argv[2])]
for c in cfgs:
    time_of(*c)
