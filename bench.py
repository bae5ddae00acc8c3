#!/usr/bin/env python
"""
bench.py -- events/s of the fused OF1x1 feature extraction on 32768-sample traces.

Workload (BASELINE.json configs[1], "C2"): of1x1_constrained (+-400 us window) with the
default template plus the glitch-template variant, single channel, 32768 samples @1.25 MHz.
A "step" is one pass of the hot path over one batch of synthetic events that is already
resident in HBM (float64 traces, `--events-per-gpu` per rank; 1 M events x 256 KiB do
not fit one GPU, so the per-GPU batch is fixed and N ranks scale weakly, each rank on its
own shard, no data-path collective).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference ...                     # CPU oracle, all host cores

`value` is measured in the float64 parity mode (same numbers as the reference's float64
path to 1e-9); the fp32 fast mode (north_star tolerances 1e-5 amp / 1e-4 chi2) is
reported next to it under "fast_mode".  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# BLAS / OpenMP pools pinned to one thread per process BEFORE numpy is imported, exactly as the reference does
# (detprocess/process/features.py:31-38): the CPU arm parallelises over processes, one per host core; a pool per
# worker process oversubscribed the host 16-32x in round 1 (63 events/s instead of ~1000 on the same 32 cores).
for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS', 'VECLIB_MAXIMUM_THREADS'):
    os.environ[_v] = '1'

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB_SAMPLES = 32768
FS = 1.25e6
WINDOW = 500            # +-400 us * 1.25 MHz
METRIC = 'events/sec OF1x1 amp+t0+chi2 (32768-sample)'
UNIT = 'events/s'
BYTES_PER_EVENT = NB_SAMPLES * 8    # algorithmic bytes: each float64 sample read once


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--events-per-gpu', type=int, default=16384)
    ap.add_argument('--e2e-events', type=int, default=4096)
    ap.add_argument('--cpu-sample', type=int, default=256, help='events per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fast-mode', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the C4 (trigger) / C5 (PSD) side measurements')
    ap.add_argument('--extras-at-scale', action='store_true',
                    help='run the per-GPU side measurements (other_rows) under torchrun as well; by default they are the N=1 line\'s')
    return ap.parse_args()


# ------------------------------------------------------------------ CPU (oracle) arm
def _cpu_worker(args):
    """Per-event loop shaped like FeatureProcessing._process (reference features.py:533-851):
    per event clear/update/filter (processing_data.py:731-772), then one qp.OF1x1 per algorithm."""
    traces, template, glitch, psd, pre = args
    from oracle.of1x1 import OFBaseOracle, OF1x1Oracle
    ofb = OFBaseOracle(FS)
    ofb.set_csd('ch', psd, coupling='AC')
    ofb.add_template('ch', template, 'default', pretrigger_samples=pre)
    ofb.add_template('ch', glitch, 'glitch', pretrigger_samples=pre)
    ofb.calc_phi('ch', 'default')
    ofb.calc_phi('ch', 'glitch')
    rows = []
    for x in traces:
        ofb.clear_signal()
        ofb.update_signal('ch', x, calc_fft=True)
        ofb.calc_signal_filt('ch')
        ofb.calc_signal_filt_td('ch')
        row = []
        for tag in ('default', 'glitch'):
            OF = OF1x1Oracle(ofb, 'ch', tag)
            OF.calc(window_min_index=pre - WINDOW, window_max_index=pre + WINDOW,
                    lowchi2_fcutoff=10000, lgc_fit_withdelay=True, lgc_fit_nodelay=False)
            row.extend(OF.get_result_withdelay())
            row.extend([OF.get_chisq_nopulse(), OF.get_energy_resolution(), OF.get_time_resolution()])
        rows.append(row)
    return rows


def cpu_arm(n_events, steps, warmup, cores=None):
    import multiprocessing as mp
    from detprocess_b200.synth import SynthSetup
    cores = cores or os.cpu_count() or 1
    S = SynthSetup(NB_SAMPLES, FS)
    traces = S.traces(n_events)
    chunks = [c for c in np.array_split(traces, cores) if len(c)]
    jobs = [(c, S.template, S.template_glitch, S.psd, S.nb_pretrigger) for c in chunks]
    ctx = mp.get_context('fork')
    times = []
    with ctx.Pool(len(jobs)) as pool:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    t = float(np.sum(times))
    return {'value': n_events * steps / t, 'unit': UNIT, 'cores': len(jobs), 'kind': 'port',
            'sample': f'{n_events} events/step x {steps} steps of the same C2 workload, '
                      f'CPU oracle restatement (not upstream detprocess+QETpy), per-event loop, '
                      f'multiprocessing.Pool({len(jobs)}), BLAS/OpenMP threads per process = '
                      f'{os.environ.get("OPENBLAS_NUM_THREADS")}/{os.environ.get("OMP_NUM_THREADS")} (set before numpy import)',
            'ms_per_step': 1e3 * t / steps}


def reference_main(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    steps, warmup = a.steps, a.warmup
    # bound the run to a few minutes: ~25 events/s/core
    cb = cpu_arm(a.cpu_sample, steps, min(warmup, 1) if warmup else 0)
    line = {'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'impl': 'reference', 'n_gpus': a.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C2: of1x1_constrained +-400us, default + glitch template, 1 ch x 32768 @1.25MHz',
                       'events_per_step': a.cpu_sample},
            'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------- GPU arm
# measured issue rates of the FP64 pipe on this part (tools/ubench/pipes.cu, profiles/r1_ubench_pipes.log), warp-instructions
# per clock and SM; the instruction counts of the headline kernel come from the committed ncu summary
PIPE_DFMA, PIPE_DADD, N_SM = 1.68, 1.97, 148


def ncu_summary(name):
    """profiles/<name>.json written by tools/ncu_summary.py from one `ncu --set full` capture of the kernel"""
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def fp_issue_floor_ms(summ, events, sm_hz):
    """time the FP64 instructions of `events` events need at the measured issue rates with every SM busy -- the floor
    of THIS instruction mix; kernel_ms / floor = how much of the launch is not FP64 issue"""
    per_event = (summ['dfma_warp_inst'] / PIPE_DFMA + (summ['dadd_warp_inst'] + summ['dmul_warp_inst']) / PIPE_DADD) / summ['events']
    return per_event * events / N_SM / sm_hz * 1e3


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(self.samples[0][1]),
                'reasons': reasons, 'samples': len(self.samples)}


def make_device_traces(S, n_events, device, seed):
    """Coloured Gaussian noise + template pulses, generated on the device with torch (data plumbing)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = S.nb_samples
    nh = n // 2 + 1
    amp = torch.from_numpy(np.sqrt(S.psd[:nh] * n * S.fs / 2.0)).to(device)
    tmpl = torch.from_numpy(S.template).to(device)
    out = torch.empty((n_events, n), dtype=torch.float64, device=device)
    chunk = 2048
    for i0 in range(0, n_events, chunk):
        m = min(chunk, n_events - i0)
        re = torch.randn((m, nh), generator=g, device=device, dtype=torch.float64)
        im = torch.randn((m, nh), generator=g, device=device, dtype=torch.float64)
        spec = torch.complex(re, im) * amp
        spec[:, 0] = spec[:, 0].real * (2.0 ** 0.5)
        spec[:, -1] = spec[:, -1].real * (2.0 ** 0.5)
        x = torch.fft.irfft(spec, n=n, dim=-1)
        a = torch.rand((m,), generator=g, device=device, dtype=torch.float64) * 2e-7
        has = torch.rand((m,), generator=g, device=device, dtype=torch.float64) < 0.9
        d = torch.randint(-300, 301, (m,), generator=g, device=device)
        idx = (torch.arange(n, device=device)[None, :] - d[:, None]) % n
        x += (a * has)[:, None] * tmpl[idx]
        out[i0:i0 + m] = x
    return out


def build_plan(S, precision, adc=None):
    from detprocess_b200.core.plans import OFPlan
    pre = S.nb_pretrigger
    plan = OFPlan(S.nb_samples, S.fs, 1, precision)
    plan.set_psd(0, S.psd, 'AC')
    if adc is not None:
        plan.set_adc_conversion(0, *adc)
    t0 = plan.add_template(0, S.template, pre)
    t1 = plan.add_template(0, S.template_glitch, pre)
    plan.add_fit(0, t0, pre - WINDOW, pre + WINDOW)
    plan.add_fit(0, t1, pre - WINDOW, pre + WINDOW)
    plan.finalize()
    return plan


def time_steps(plan, x, out, steps, warmup, dist, world):
    """K timed steps bracketed by barrier + synchronize; CUDA events on the launch stream;
    max over ranks.  Returns (ms_total_max, per-launch kernel ms list)."""
    import torch
    for _ in range(warmup):
        plan.run(x, out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.run(x, out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    # per-launch duration of the dominant kernel (CUDA events recorded inside the C ABI on the launch stream)
    kms = []
    for _ in range(3):
        plan.run(x, out)
        kms.append(plan.last_kernel_ms())
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=x.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, kms


def extras(device, dist, world, hbm_peak):
    """Side measurements of the other hot-path rows (not the headline metric): C5 noise PSD on
    65536-sample traces (+ the all-reduce of the per-GPU sums when N > 1) and C4 continuous-stream
    trigger on one 10 s stream.  Device-resident synthetic inputs, CUDA-event kernel times."""
    import torch
    from detprocess_b200.core.noise import NoisePSD
    from detprocess_b200.core.oftrigger import OptimumFilterTrigger
    from detprocess_b200.synth import make_template, make_psd, SynthSetup
    out = {}
    # ---- C1 window reductions (baseline_pre + integral, README.md:87-96 windows), bit-exact numpy arithmetic
    from detprocess_b200.core.plans import ReducePlan
    n, B = NB_SAMPLES, 8192
    red = ReducePlan(n, FS, 1)
    red.add(0, 'baseline', 0, n // 2 - 1250)
    red.add(0, 'integral', n // 2 - 625, n // 2 + 625)
    red.finalize(device)
    xr = torch.randn((B, n), dtype=torch.float64, device=device)
    ro = torch.empty((B, red.n_out), dtype=torch.float64, device=device)
    for _ in range(3):
        red.run(xr, ro)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        red.run(xr, ro)
        ms.append(red.last_kernel_ms())
    m = float(np.median(ms))
    wbytes = B * ((n // 2 - 1250) + 1250) * 8
    out['c1_window_reductions'] = {'events_per_s_per_gpu': B / (m * 1e-3), 'achieved_gbs': wbytes / (m * 1e-3) / 1e9,
                                   'roofline_frac': wbytes / (m * 1e-3) / 1e9 / hbm_peak, 'bound': 'hbm',
                                   'algorithmic_bytes_per_event': wbytes // B, 'kernel': 'dp_reduce_kernel<256>'}
    del xr, ro, red
    # ---- C5
    n, B = 65536, 4096
    x = torch.randn((B, n), dtype=torch.float64, device=device) * 1e-10
    for prec in ('f64', 'f32'):
        est = NoisePSD(n, FS, precision=prec, device=device, typical_rms=1e-10)
        for _ in range(2):
            est.update(x)
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            est.update(x)
            ms.append(est.plan.last_kernel_ms())
        m = float(np.mean(ms))
        t0 = time.perf_counter()
        est.finalize()                      # all-reduce of [N/2+1] sums + count (NCCL when world > 1)
        torch.cuda.synchronize()
        out[f'c5_psd_{prec}'] = {'traces_per_s_per_gpu': B / (m * 1e-3), 'achieved_gbs': B * n * 8 / (m * 1e-3) / 1e9,
                                 'roofline_frac': B * n * 8 / (m * 1e-3) / 1e9 / hbm_peak,
                                 'finalize_ms_incl_allreduce': 1e3 * (time.perf_counter() - t0), 'nb_samples': n}
        del est
    del x
    # ---- C4: one continuous 10 s stream @1.25 MHz, 32768-tap filter
    nt, L = 32768, 12_500_000
    g = torch.Generator(device=device)
    g.manual_seed(777)
    stream = torch.randn(L, generator=g, device=device, dtype=torch.float64) * 1.2e-8
    tmpl = make_template(nt, FS)
    psd = make_psd(nt, FS)
    shape = torch.from_numpy(tmpl[nt // 2:]).to(device)
    for t0 in range(300_000, L - 300_000, 250_000):             # 5 Hz pulses
        stream[t0:t0 + shape.shape[0]] += 1.5e-7 * shape
    for prec in ('f64', 'f32'):
        trig = OptimumFilterTrigger('ch', FS, tmpl, psd, nt // 2, precision=prec, max_samples=L, device=device)
        if prec == 'f32':
            trig._plan.set_scale(1.2e-8)
        trig.update_trace(stream)
        for _ in range(2):
            d = trig.find_triggers_once(5.0, pileup_window_msec=1.0)
        torch.cuda.synchronize()
        fms, gms = trig._plan.last_kernel_ms()
        out[f'c4_trigger_{prec}'] = {'samples_per_s_per_gpu': L / ((fms + gms) * 1e-3), 'filter_ms': fms, 'group_ms': gms,
                                     'stream_seconds': L / FS, 'n_triggers': len(d['ch']['trigger_index']),
                                     'achieved_gbs': L * 8 / (fms * 1e-3) / 1e9,
                                     'roofline_frac': L * 8 / (fms * 1e-3) / 1e9 / hbm_peak,
                                     'fft_size': trig._plan.fft_size, 'hop': trig._plan.hop}
        del trig
    # ---- C1 OF part (nodelay + unconstrained, one template) and C3 (8 channels x 16384, the full YAML feature set:
    # nodelay, constrained, unconstrained, constrained glitch + baseline, baseline_end, maximum, minimum, integral)
    from detprocess_b200.core.plans import OFPlan
    S1 = SynthSetup(NB_SAMPLES, FS)
    B = 8192
    x1 = make_device_traces(S1, B, device, 777)
    for prec in ('f64', 'f32'):
        pl = OFPlan(NB_SAMPLES, FS, 1, prec)
        pl.set_psd(0, S1.psd, 'AC')
        t0 = pl.add_template(0, S1.template, S1.nb_pretrigger)
        pl.add_fit_nodelay(0, t0)
        pl.add_fit(0, t0, None, None)
        pl.finalize(device)
        o1 = torch.empty((B, pl.n_out), dtype=torch.float64, device=device)
        for _ in range(3):
            pl.run(x1, o1)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            pl.run(x1, o1)
            ms.append(pl.last_kernel_ms())
        m = float(np.median(ms))
        gbs = B * BYTES_PER_EVENT / (m * 1e-3) / 1e9
        out[f'c1_of_nodelay_unconstrained_{prec}'] = {'events_per_s_per_gpu': B / (m * 1e-3), 'achieved_gbs': gbs,
                                                      'roofline_frac': gbs / hbm_peak, 'algorithmic_bytes_per_event': BYTES_PER_EVENT}
        del pl, o1
    del x1
    n3, c3, B = 16384, 8, 2048
    S3 = SynthSetup(n3, FS)
    x3 = make_device_traces(S3, B * c3, device, 778).reshape(B, c3, n3)
    pre3 = S3.nb_pretrigger
    red3 = ReducePlan(n3, FS, c3)
    for c in range(c3):
        red3.add(c, 'baseline', 0, pre3 - 1250)
        red3.add(c, 'baseline', n3 - 1250, n3)
        red3.add(c, 'maximum', None, None)
        red3.add(c, 'minimum', None, None)
        red3.add(c, 'integral', pre3 - 625, pre3 + 625)
    red3.finalize(device)
    r3 = torch.empty((B, red3.n_out), dtype=torch.float64, device=device)
    for prec in ('f64', 'f32'):
        pl = OFPlan(n3, FS, c3, prec)
        for c in range(c3):
            pl.set_psd(c, S3.psd, 'AC')
            t0 = pl.add_template(c, S3.template, pre3)
            t1 = pl.add_template(c, S3.template_glitch, pre3)
            pl.add_fit_nodelay(c, t0)
            pl.add_fit(c, t0, pre3 - WINDOW, pre3 + WINDOW)
            pl.add_fit(c, t0, None, None)
            pl.add_fit(c, t1, pre3 - WINDOW, pre3 + WINDOW)
        pl.finalize(device)
        o3 = torch.empty((B, pl.n_out), dtype=torch.float64, device=device)
        for _ in range(3):
            pl.run(x3, o3)
            red3.run(x3, r3)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            pl.run(x3, o3)
            red3.run(x3, r3)
            ms.append((pl.last_kernel_ms(), red3.last_kernel_ms()))
        mo, mr = (float(v) for v in np.median(np.array(ms), axis=0))
        byt = c3 * n3 * 8
        out[f'c3_8ch_16384_full_yaml_{prec}'] = {'events_per_s_per_gpu': B / ((mo + mr) * 1e-3), 'of_ms': mo, 'reduce_ms': mr,
                                                'achieved_gbs': B * byt / ((mo + mr) * 1e-3) / 1e9,
                                                'roofline_frac': B * byt / ((mo + mr) * 1e-3) / 1e9 / hbm_peak,
                                                'algorithmic_bytes_per_event': byt, 'features_per_event': pl.n_out + red3.n_out}
        del pl, o3
    del x3, red3, r3
    # ---- the YAML-driven pipeline (FeatureProcessing: reader -> H2D -> fused kernels -> feature table) on the C2 feature
    # set + baseline / integral, from pinned host events: int16 ADC counts (as stored on disk) and float64 amps
    import tempfile
    from detprocess_b200.core.filterdata import FilterData
    from detprocess_b200.io import ArrayReader
    from detprocess_b200.process.features import FeatureProcessing
    S2 = SynthSetup(NB_SAMPLES, FS)
    gain = 1.0e-11
    E16, E64 = 16384, 4096        # 1 GiB of pinned host memory each way
    host_f64 = torch.empty((E64, 1, NB_SAMPLES), dtype=torch.float64).pin_memory()
    host_i16 = torch.empty((E16, 1, NB_SAMPLES), dtype=torch.int16).pin_memory()
    for i0 in range(0, E16, 4096):
        xa = make_device_traces(S2, 4096, device, 779 + i0)
        if i0 == 0:
            host_f64[:, 0].copy_(xa.cpu())
        host_i16[i0:i0 + 4096, 0].copy_(torch.clamp(torch.round(xa / gain), -32768, 32767).to(torch.int16).cpu())
        del xa
    fd = FilterData()
    fd.set_psd('chan1', S2.psd, sample_rate=FS)
    fd.set_template('chan1', S2.template, sample_rate=FS, pretrigger_length_samples=S2.nb_pretrigger)
    fd.set_template('chan1', S2.template_glitch, sample_rate=FS, pretrigger_length_samples=S2.nb_pretrigger, tag='glitch')
    with tempfile.TemporaryDirectory() as td:
        yml = os.path.join(td, 'c2.yaml')
        with open(yml, 'w') as f:
            f.write('global:\n    trace_length_samples: %d\n    pretrigger_length_samples: %d\n' % (NB_SAMPLES, S2.nb_pretrigger)
                    + 'chan1:\n'
                    + '    of1x1_constrained:\n        run: True\n        template_tag: default\n'
                    + '        window_min_from_trig_usec: -400\n        window_max_from_trig_usec: 400\n'
                    + '    of1x1_glitch:\n        run: True\n        base_algorithm: of1x1_constrained\n        template_tag: glitch\n'
                    + '        window_min_from_trig_usec: -400\n        window_max_from_trig_usec: 400\n'
                    + '    baseline:\n        run: True\n        window_min_from_start_usec: 0\n        window_max_from_trig_usec: -1000\n'
                    + '    integral:\n        run: True\n        window_min_from_trig_usec: -500\n        window_max_from_trig_usec: 500\n')
        for name, host, bs, kw in (('int16', host_i16, 2048, {'adc_gain': [gain], 'adc_offset': [0.0]}), ('float64', host_f64, 512, {})):
            E = int(host.shape[0])
            fp = FeatureProcessing(ArrayReader(host, ['chan1'], FS, **kw), yml, filter_data=fd, verbose=False)
            fp.process(batch_size=bs, gather=False)           # warm-up (plans, staging)
            torch.cuda.synchronize()
            dts = []
            for _ in range(3):
                t0 = time.perf_counter()
                df = fp.process(batch_size=bs, gather=False)  # 8 batches: upload k+1 | kernels k | table columns k-1
                torch.cuda.synchronize()
                dts.append(time.perf_counter() - t0)
            dt = float(np.median(dts))
            # `process` shards the reader's events over the ranks: the job handles E events in dt, each GPU E / world
            out[f'pipeline_yaml_c2_{name}'] = {'events_per_s_job': E / dt, 'events_per_s_per_gpu': E / world / dt, 'events': E,
                                               'batch_size': bs, 'columns': int(df.shape[1]),
                                               'h2d_bytes': int(host.numel() * host.element_size()),
                                               'api': 'FeatureProcessing.process (YAML, FilterData, pinned host events)'}
            del fp, df
    del host_f64, host_i16
    # ---- (f)3 NxM optimal filter: 2 channels x 2 templates, 32768 samples, +-400 us window + no-delay fit
    from detprocess_b200.core.plans import NxMPlan
    from detprocess_b200.synth import SynthNxM
    SN = SynthNxM(NB_SAMPLES, 2, 2, FS)
    B = 4096
    base = torch.from_numpy(SN.traces(64, np.random.default_rng(12345))).to(device)
    xn = base.repeat(B // 64, 1, 1).contiguous()        # 2.0 GiB > L2
    for prec in ('f64', 'f32'):
        nx = NxMPlan(NB_SAMPLES, FS, 2, 2, prec)
        nx.set_filter(SN.templates, SN.csd, SN.nb_pretrigger, 'AC')
        nx.set_window(SN.nb_pretrigger - WINDOW, SN.nb_pretrigger + WINDOW)
        nx.finalize(device)
        no = torch.empty((B, nx.n_out), dtype=torch.float64, device=device)
        for _ in range(3):
            nx.run(xn, no)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            nx.run(xn, no)
            ms.append(nx.last_kernel_ms())
        m = float(np.median(ms))
        gbs = B * 2 * NB_SAMPLES * 8 / (m * 1e-3) / 1e9
        out[f'f3_ofnxm_2x2_{prec}'] = {'events_per_s_per_gpu': B / (m * 1e-3), 'achieved_gbs': gbs, 'roofline_frac': gbs / hbm_peak,
                                       'algorithmic_bytes_per_event': 2 * NB_SAMPLES * 8, 'kernel': 'dp_nxm_kernel<%s,4,2>' % ('double' if prec == 'f64' else 'f2')}
        del nx, no
    # ---- noise CSD of the same 2-channel shape (Noise.calc_csd)
    from detprocess_b200.core.noise import NoiseCSD
    for prec in ('f64', 'f32'):
        est = NoiseCSD(NB_SAMPLES, FS, 2, precision=prec, device=device, typical_rms=1e-8)
        for _ in range(2):
            est.update(xn)
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            est.update(xn)
            ms.append(est.plan.last_kernel_ms())
        m = float(np.median(ms))
        t0 = time.perf_counter()
        est.finalize()                   # reduce over CTAs + all-reduce over ranks + unfold to [2, 2, N]
        fin_ms = (time.perf_counter() - t0) * 1e3
        gbs = B * 2 * NB_SAMPLES * 8 / (m * 1e-3) / 1e9
        out[f'csd_2ch_{prec}'] = {'events_per_s_per_gpu': B / (m * 1e-3), 'achieved_gbs': gbs, 'roofline_frac': gbs / hbm_peak,
                                  'finalize_ms_incl_allreduce': fin_ms, 'nb_samples': NB_SAMPLES}
        del est
    del xn, base
    return out


def collective_rows(device, dist, world):
    """The two collectives the multi-GPU design rests on, timed warm on the device (CUDA events, max over ranks), NCCL
    over NVLink: (a) the gather of the per-event feature tables -- 1 M events x 17 float64 columns in total, through
    detprocess_b200.process.features.gather_frames' tensor path; (b) the all-reduce of the per-GPU periodogram sums
    [N/2+1] + count of the PSD estimator (N = 65536); (c) the same for the CSD sums [n^2][N/2+1], n = 2, N = 32768."""
    import torch
    out = {}
    if world == 1:
        return out

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows, ncol = 1_000_000 // world, 17
    block = torch.randn((rows, ncol), dtype=torch.float64, device=device)
    parts = [torch.empty_like(block) for _ in range(world)]
    ms = timed(lambda: dist.all_gather(parts, block))
    out['gather_feature_table_1M_x17'] = {'ms': ms, 'bytes_total': rows * world * ncol * 8,
                                          'gbs_per_rank_received': rows * (world - 1) * ncol * 8 / (ms * 1e-3) / 1e9,
                                          'collective': 'all_gather (NCCL), one [B/G, 17] float64 block per rank'}
    sums = torch.randn(65536 // 2 + 1, dtype=torch.float64, device=device)
    cnt = torch.ones(1, dtype=torch.int64, device=device)

    def psd_reduce():
        dist.all_reduce(sums)
        dist.all_reduce(cnt)
    out['allreduce_psd_sums_65536'] = {'ms': timed(psd_reduce), 'bytes': int(sums.numel() * 8 + 8),
                                       'collective': 'all_reduce(SUM) of [N/2+1] float64 + int64 count (NoisePSD.finalize)'}
    csd = torch.randn((4, 32768 // 2 + 1), dtype=torch.float64, device=device)

    def csd_reduce():
        dist.all_reduce(csd)
        dist.all_reduce(cnt)
    out['allreduce_csd_sums_2ch_32768'] = {'ms': timed(csd_reduce), 'bytes': int(csd.numel() * 8 + 8),
                                           'collective': 'all_reduce(SUM) of [n^2][N/2+1] float64 + count (NoiseCSD.finalize)'}
    return out


def gpu_main(a):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    from detprocess_b200.synth import SynthSetup
    from detprocess_b200 import build as _build
    _build.build(verbose=False)
    # CPU baseline first: its worker pool is forked before this process touches CUDA
    cb = None
    if not a.no_cpu_baseline and world == 1:
        cb = cpu_arm(a.cpu_sample, 3, 1)

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (detprocess_b200 has no CPU fallback)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    # one process per GPU: a disjoint slice of the GPU-local cores before any pinned allocation (detprocess_b200/utils/affinity.py)
    from detprocess_b200.utils.affinity import bind_to_gpu
    placement = bind_to_gpu(local, local, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    placement['bound_cpus'] = f"{len(placement['bound_cpus'])} cpus from {min(placement['bound_cpus'])}" if placement.get('bound_cpus') else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    S = SynthSetup(NB_SAMPLES, FS)
    B = a.events_per_gpu
    x = make_device_traces(S, B, device, 12345 + rank)       # per-shard seed
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'

    sampler = ClockSampler(local) if rank == 0 else None
    results = {}
    for prec in (['f64'] if a.no_fast_mode else ['f64', 'f32']):
        plan = build_plan(S, prec)
        out = torch.empty((B, plan.n_out), dtype=torch.float64, device=device)
        if prec == 'f64' and sampler:
            sampler.start()         # samples clocks / throttle reasons over every timed region below
        ms, kms = time_steps(plan, x, out, a.steps, a.warmup, dist, world)
        kms_avg = float(np.mean(kms))
        results[prec] = {'ms_total': ms, 'ms_per_step': ms / a.steps, 'value': world * B * a.steps / (ms * 1e-3),
                         'kernel_ms': kms_avg, 'achieved_gbs': B * BYTES_PER_EVENT / (kms_avg * 1e-3) / 1e9,
                         'launches': a.steps}
        # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region.  The
        # headline e2e ships the events the way the detector stores them, int16 ADC counts (converted in the kernel's
        # load with the channel's gain / offset); the float64-amps variant (4x the PCIe bytes) is reported beside it.
        E = min(a.e2e_events, B)
        reps = max(1, min(a.steps, 5))

        def time_host(plan_, host_):
            hout = np.empty((E, plan_.n_out), dtype=np.float64)
            plan_.run_host(host_, hout)      # warm-up (allocates staging once)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                plan_.run_host(host_, hout)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return world * E * reps / dt

        host = torch.empty((E, NB_SAMPLES), dtype=torch.float64).pin_memory()
        host.copy_(x[:E].cpu())
        results[prec]['e2e_f64'] = {'value': time_host(plan, host), 'unit': UNIT,
                                    'h2d_bytes_per_step': int(E * BYTES_PER_EVENT),
                                    'd2h_bytes_per_step': int(E * plan.n_out * 8), 'events_per_step': E,
                                    'input': 'pinned host float64 amps',
                                    'api': 'OFPlan.run_host -> dp_of1x1_batch_host'}
        results[prec]['n_out'] = plan.n_out
        del plan, out, host
        gain = 1.0e-11
        aplan = build_plan(S, prec, adc=(gain, 0.0))
        ahost = torch.empty((E, NB_SAMPLES), dtype=torch.int16).pin_memory()
        ahost.copy_(torch.clamp(torch.round(x[:E] / gain), -32768, 32767).to(torch.int16).cpu())
        results[prec]['e2e'] = {'value': time_host(aplan, ahost), 'unit': UNIT, 'h2d_bytes_per_step': int(E * NB_SAMPLES * 2),
                                'd2h_bytes_per_step': int(E * aplan.n_out * 8), 'events_per_step': E,
                                'input': 'pinned host int16 ADC counts as stored by the DAQ, adc->amps in the kernel load '
                                         '(dp_of_plan_set_adc_conversion)',
                                'api': 'OFPlan.run_host -> dp_of1x1_batch_host'}
        del aplan, ahost

    ex = None
    # the side rows are per-GPU measurements without a collective: the N = 1 line carries them; under torchrun every rank
    # would repeat them (minutes of plan building and pinned-memory set-up on a shared host) for no new information
    if not a.no_extras and (world == 1 or a.extras_at_scale):
        ex = extras(device, dist, world, hbm_peak)
    coll = collective_rows(device, dist, world)
    if sampler:
        sampler.stop_flag = True
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    r = results['f64']
    summ = ncu_summary('r2_ncu_of2_f64_c2.json')
    roofline = {'bound': 'hbm', 'achieved': r['achieved_gbs'], 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': r['achieved_gbs'] / hbm_peak, 'traffic': None, 'peak_source': peak_src,
                'kernel': 'dp_of2_kernel<double,4,0,true,true>', 'kernel_ms': r['kernel_ms'],
                'algorithmic_bytes_per_event': BYTES_PER_EVENT,
                'note': 'the FFT path is FP64-issue / latency bound, not HBM bound (10 FLOP/B): fp_issue_frac is the share '
                        'of the launch the FP64 instructions alone need at the measured pipe rates (DESIGN.md 4.1)'}
    if summ is not None:
        sm_hz = (sampler.summary().get('sm_mhz') or 1965.0) * 1e6 if sampler else 1965.0e6
        floor = fp_issue_floor_ms(summ, B, sm_hz)
        roofline.update({
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (ncu --set full), per event x events
            'traffic': summ['dram_bytes_per_event'] * B,
            'traffic_over_algorithmic': summ['dram_bytes_per_event'] / BYTES_PER_EVENT,
            'traffic_source': 'profiles/r2_ncu_of2_f64_c2.json (tools/ncu_summary.py of profiles/r2_prof_of2_f64_32k_c2.txt)',
            'fp_issue_floor_ms': floor, 'fp_issue_frac': floor / r['kernel_ms'],
            'fp_issue_model': f'(DFMA / {PIPE_DFMA} + (DADD + DMUL) / {PIPE_DADD}) warp-inst per clk and SM x {N_SM} SMs at the '
                              f'sampled SM clock; counts from the ncu summary: {summ["dfma_warp_inst"] / summ["events"]:.0f} DFMA + '
                              f'{(summ["dadd_warp_inst"] + summ["dmul_warp_inst"]) / summ["events"]:.0f} DADD/DMUL warp-inst per event'})
    line = {
        'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: of1x1_constrained +-400us, default + glitch template, 1 ch x 32768 @1.25MHz',
                   'events_per_gpu_per_step': B, 'input': 'float64 traces resident in HBM',
                   'l2': f'inputs ({B * BYTES_PER_EVENT / 2**30:.1f} GiB/GPU) larger than L2, no flush needed',
                   'sharding': 'events sharded by rank, no data-path collective'},
        'roofline': roofline,
        'e2e': {k: r['e2e'][k] for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'input', 'api')},
        'e2e_f64_host': r['e2e_f64'],
        'gpu_launches': r['launches'] * world,
        'clocks': sampler.summary() if sampler else None,
    }
    if 'f32' in results:
        f = results['f32']
        summ32 = ncu_summary('r2_ncu_of2_f32_c2.json')
        line['fast_mode'] = {'dtype': 'f32', 'value': f['value'], 'unit': UNIT, 'ms_per_step': f['ms_per_step'],
                             'roofline_frac': f['achieved_gbs'] / hbm_peak, 'achieved_gbs': f['achieved_gbs'],
                             'kernel': 'dp_of2_kernel<f2,4,0,true,true>', 'e2e': f['e2e']['value'],
                             'e2e_f64_host': f['e2e_f64']['value'],
                             'traffic': (summ32['dram_bytes_per_event'] * B) if summ32 else None,
                             'tolerance': 'amp 1e-5, chi2 1e-4 rel vs float64 oracle'}
    if ex is not None:
        line['other_rows'] = ex
    if coll:
        line['collectives'] = coll
    line['host_placement'] = placement
    if cb is not None:
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    args = parse()
    sys.exit(reference_main(args) if args.impl == 'reference' else gpu_main(args))
