"""One CSD configuration, few launches: target for ncu (development aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.core.noise import NoiseCSD

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
prec = sys.argv[2] if len(sys.argv) > 2 else 'f64'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
x = torch.randn((B, n, N), dtype=torch.float64, device='cuda') * 1e-8
est = NoiseCSD(N, 1.25e6, n, precision=prec, typical_rms=1e-8)
for _ in range(3):
    est.update(x)
torch.cuda.synchronize()
ms = est.plan.last_kernel_ms()
print(f'CSD N={N} {prec} B={B} n={n}: {ms:.3f} ms {B/ms/1e3:.3f} Mev/s')
