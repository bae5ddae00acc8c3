"""One NxM configuration, few launches: target for ncu (development aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthNxM
from detprocess_b200.core.plans import NxMPlan

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
prec = sys.argv[2] if len(sys.argv) > 2 else 'f32'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
m = int(sys.argv[5]) if len(sys.argv) > 5 else 2
S = SynthNxM(N, n, m)
pre = S.nb_pretrigger
plan = NxMPlan(N, S.fs, n, m, prec)
plan.set_filter(S.templates, S.csd, pre, 'AC')
plan.set_window(pre - 500, pre + 500)
plan.finalize()
base = torch.from_numpy(S.traces(64, np.random.default_rng(1))).cuda()
x = base.repeat((B + 63) // 64, 1, 1)[:B].contiguous()
out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
for _ in range(3):
    plan.run(x, out)
torch.cuda.synchronize()
ms = plan.last_kernel_ms()
print(f'NxM N={N} {prec} B={B} n={n} m={m}: {ms:.3f} ms {B/ms/1e3:.3f} Mev/s')
