"""One config, few launches: target for ncu (development aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthSetup, make_traces
from detprocess_b200.core.plans import OFPlan

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
prec = sys.argv[2] if len(sys.argv) > 2 else 'f32'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
mode = sys.argv[4] if len(sys.argv) > 4 else 'c1'
S = SynthSetup(N)
pre = S.nb_pretrigger
plan = OFPlan(N, S.fs, 1, prec)
plan.set_psd(0, S.psd)
t = plan.add_template(0, S.template, pre)
if mode == 'c1':
    plan.add_fit_nodelay(0, t)
    plan.add_fit(0, t, None, None)
else:
    plan.add_fit(0, t, pre - 500, pre + 500)
    g = plan.add_template(0, S.template_glitch, pre)
    plan.add_fit(0, g, pre - 500, pre + 500)
plan.finalize()
base = torch.from_numpy(make_traces(128, S.template, S.psd, S.fs, np.random.default_rng(1))).cuda()
x = base.repeat((B + 127) // 128, 1)[:B].contiguous()
out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
for _ in range(3):
    plan.run(x, out)
torch.cuda.synchronize()
ms = plan.last_kernel_ms()
print(f'N={N} {prec} B={B} {mode}: {ms:.3f} ms {B/ms/1e3:.3f} Mev/s')
