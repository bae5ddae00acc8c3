import sys, numpy as np, torch
sys.path.insert(0, '.')
from detprocess_b200.core.oftrigger import OptimumFilterTrigger
from detprocess_b200.synth import SynthSetup, make_continuous
S = SynthSetup(32768)
L = 12_500_000
for rate in (5.0, 50.0):
    x = torch.from_numpy(make_continuous(L, S.template, S.psd, S.fs, np.random.default_rng(1), pulse_rate_hz=rate)).cuda()
    for prec in ('f64', 'f32'):
        trig = OptimumFilterTrigger('ch', S.fs, S.template, S.psd, S.nb_pretrigger, precision=prec, max_samples=L)
        trig.update_trace(x)
        for _ in range(3):
            d = trig.find_triggers_once(5.0, pileup_window_msec=1.0)
        torch.cuda.synchronize()
        f, g = trig._plan.last_kernel_ms()
        print(f'rate {rate} {prec}: filter {f:.3f} ms group {g:.3f} ms triggers {len(d["ch"]["trigger_index"])}')
