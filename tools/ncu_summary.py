"""Summarise one kernel of an ncu report into the small JSON bench.py reads for `roofline.traffic` and
`roofline.fp_issue_frac` (development aid; run where ncu is installed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep <events in the profiled launch> profiles/r2_ncu_of2_f64_c2.json
"""
import csv, io, json, subprocess, sys

rep, events, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(key):
    v, u = m[key]
    x = float(v.replace(',', ''))
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3,
             'nsecond': 1e-9, 'second': 1.0, 'Ghz': 1e9, 'Mhz': 1e6, 'Kbyte/block': 1e3, 'byte/block': 1.0}
    return x * scale.get(u, 1.0)


def first(*keys):
    for k in keys:
        if k in m:
            return num(k)
    return None


d = {
    'report': rep, 'kernel': m['Kernel Name'][0], 'events': events,
    'duration_s': first('gpu__time_duration.sum'),
    'dram_bytes_read': first('dram__bytes_read.sum'), 'dram_bytes_write': first('dram__bytes_write.sum'),
    'inst_executed': first('smsp__inst_executed.sum', 'sm__inst_executed.sum'),
    # thread-level FP instruction rates (per elapsed SM cycle, summed over the chip); totals are formed below
    'rate_dfma': first('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed'),
    'rate_dadd': first('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed'),
    'rate_dmul': first('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed'),
    'rate_ffma': first('smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed'),
    'rate_fadd': first('smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed'),
    'rate_fmul': first('smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed'),
    'pipe_fp64_pct': first('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
    'issue_slots_busy_pct': first('sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct'),
    'l1tex_throughput_pct': first('l1tex__throughput.avg.pct_of_peak_sustained_active'),
    'smem_wavefronts_pct': first('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
    'registers_per_thread': first('launch__registers_per_thread'),
    'dyn_smem_bytes': first('launch__shared_mem_per_block_dynamic'),
    'sm_clock_hz': first('sm__cycles_elapsed.avg.per_second'),
}
d['duration_s'] *= 1.0          # ncu reports us: converted by num()
d['sm_clock_hz'] = d['sm_clock_hz'] * 1e9 if d['sm_clock_hz'] and d['sm_clock_hz'] < 1e6 else d['sm_clock_hz']
cycles = d['duration_s'] * d['sm_clock_hz']
for op in ('dfma', 'dadd', 'dmul', 'ffma', 'fadd', 'fmul'):
    r = d.pop('rate_' + op)
    d[op + '_warp_inst'] = None if r is None else r * cycles / 32.0      # thread instructions -> warp instructions
d['dyn_smem_bytes'] = d['dyn_smem_bytes'] * 1.0
d['dram_bytes_per_event'] = (d['dram_bytes_read'] + d['dram_bytes_write']) / events
json.dump(d, open(out, 'w'), indent=1)
print(json.dumps(d, indent=1))
