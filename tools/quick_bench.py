"""Quick device-resident timing of the kernels (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from detprocess_b200.synth import SynthSetup, make_traces
from detprocess_b200.core.plans import OFPlan, ReducePlan


def time_of(N, prec, B, two_templ=False, windows='c1'):
    S = SynthSetup(N)
    pre = S.nb_pretrigger
    plan = OFPlan(N, S.fs, 1, prec)
    plan.set_psd(0, S.psd)
    t = plan.add_template(0, S.template, pre)
    if windows == 'c1':
        plan.add_fit_nodelay(0, t); plan.add_fit(0, t, None, None)
    else:
        plan.add_fit(0, t, pre - 500, pre + 500)
    if two_templ:
        g = plan.add_template(0, S.template_glitch, pre)
        plan.add_fit(0, g, pre - 500, pre + 500)
    plan.finalize()
    base = torch.from_numpy(make_traces(256, S.template, S.psd, S.fs, np.random.default_rng(1))).cuda()
    x = base.repeat((B + 255) // 256, 1)[:B].contiguous()
    out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
    for _ in range(3):
        plan.run(x, out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        plan.run(x, out)
        ms.append(plan.last_kernel_ms())
    m = float(np.median(ms))
    evs = B / (m * 1e-3)
    gbs = evs * N * 8 / 1e9
    print(f'OF N={N} {prec} B={B} two_templ={two_templ} win={windows}: {m:.3f} ms  {evs/1e6:.3f} Mev/s  {gbs:.0f} GB/s ({gbs/6551*100:.1f}% HBM)', flush=True)


def time_reduce(N, B):
    plan = ReducePlan(N, 1.25e6, 1)
    plan.add(0, 'baseline', 0, N // 2 - 1250)
    plan.add(0, 'integral', N // 2 - 625, N // 2 + 625)
    plan.finalize()
    x = torch.randn((B, N), dtype=torch.float64, device='cuda')
    out = torch.empty((B, plan.n_out), dtype=torch.float64, device='cuda')
    for _ in range(3):
        plan.run(x, out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        plan.run(x, out)
        ms.append(plan.last_kernel_ms())
    m = float(np.median(ms))
    byt = B * ((N // 2 - 1250) + 1250) * 8
    print(f'reduce N={N} B={B}: {m:.3f} ms {B/(m*1e-3)/1e6:.3f} Mev/s window-bytes {byt/(m*1e-3)/1e9:.0f} GB/s', flush=True)


def time_psd(N, prec, B):
    from detprocess_b200.core.plans import PSDPlan
    plan = PSDPlan(N, 1.25e6, precision=prec)
    plan.set_scale(1.0)
    x = torch.randn((B, N), dtype=torch.float64, device='cuda')
    for _ in range(2):
        plan.accumulate(x)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        plan.accumulate(x)
        ms.append(plan.last_kernel_ms())
    m = float(np.median(ms))
    print(f'PSD N={N} {prec} B={B}: {m:.3f} ms {B/(m*1e-3)/1e6:.3f} Mtraces/s {B*N*8/(m*1e-3)/1e9:.0f} GB/s ({B*N*8/(m*1e-3)/1e9/6551*100:.1f}% HBM)', flush=True)


def time_nxm(N, prec, B, n, m):
    from detprocess_b200.synth import SynthNxM
    from detprocess_b200.core.plans import NxMPlan
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
    ms = []
    for _ in range(5):
        plan.run(x, out)
        ms.append(plan.last_kernel_ms())
    mm = float(np.median(ms))
    gbs = B * n * N * 8 / (mm * 1e-3) / 1e9
    print(f'NxM N={N} {prec} B={B} n={n} m={m}: {mm:.3f} ms  {B/(mm*1e-3)/1e6:.3f} Mev/s  {gbs:.0f} GB/s ({gbs/6551*100:.1f}% HBM)', flush=True)


if __name__ == '__main__':
    import os
    print(torch.cuda.get_device_name(0), 'DP_OF_KERNEL=' + os.environ.get('DP_OF_KERNEL', 'v2'))
    only = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if only == 'nxm':
        time_nxm(32768, 'f32', 4096, 2, 2)
        time_nxm(32768, 'f64', 4096, 2, 2)
        time_nxm(16384, 'f32', 4096, 4, 1)
        time_nxm(16384, 'f64', 4096, 3, 3)
        time_nxm(32768, 'f32', 4096, 1, 1)
        time_nxm(32768, 'f64', 4096, 1, 1)
        sys.exit(0)
    time_of(32768, 'f32', 8192)
    time_of(32768, 'f32', 8192, windows='c2')
    time_of(32768, 'f32', 8192, two_templ=True, windows='c2')
    time_of(32768, 'f64', 8192)
    time_of(32768, 'f64', 8192, two_templ=True, windows='c2')
    if only == 'all':
        time_of(16384, 'f32', 16384)
        time_of(16384, 'f64', 8192)
        time_of(65536, 'f32', 4096)
        if os.environ.get('DP_OF_KERNEL', 'v2') != 'v1':
            time_of(65536, 'f64', 4096)
        time_of(8192, 'f64', 8192)
        time_reduce(32768, 8192)
        time_psd(65536, 'f64', 4096)
        time_psd(65536, 'f32', 4096)
        time_psd(32768, 'f64', 8192)
