"""Helper: drive the host emulator of the OF kernel and compare with the oracle."""
import os, struct, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, '_build', 'emu_of')


def build(tsan=False, v2=False, asan=False):
    os.makedirs(os.path.join(HERE, '_build'), exist_ok=True)
    exe = EXE + ('2' if v2 else '') + ('_tsan' if tsan else '') + ('_asan' if asan else '')
    src = os.path.join(HERE, 'emu_of2.cpp' if v2 else 'emu_of.cpp')
    deps = [src] + [os.path.join(HERE, '../../detprocess_b200/csrc', f) for f in
                    ('dp_of_kernel.cuh', 'dp_fft.cuh', 'dp_platform.cuh', 'dp_plan.hpp', 'dp_of2_kernel.cuh',
                     'dp_plan2.hpp', 'dp_f2.cuh')]
    if os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps):
        return exe
    cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-o', exe, src]
    if tsan:
        cmd[3:3] = ['-fsanitize=thread', '-g']
    if asan:
        cmd[3:3] = ['-fsanitize=address', '-fno-omit-frame-pointer', '-g']
    subprocess.check_call(cmd)
    return exe


def build_generic():
    """emulator of the mixed-radix kernel (dp_ofg_kernel.cuh)"""
    os.makedirs(os.path.join(HERE, '_build'), exist_ok=True)
    exe = os.path.join(HERE, '_build', 'emu_ofg')
    src = os.path.join(HERE, 'emu_ofg.cpp')
    deps = [src] + [os.path.join(HERE, '../../detprocess_b200/csrc', f) for f in
                    ('dp_of_kernel.cuh', 'dp_fft.cuh', 'dp_platform.cuh', 'dp_plan.hpp', 'dp_ofg_kernel.cuh', 'dp_ofg_plan.hpp')]
    if os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps):
        return exe
    subprocess.check_call(['g++', '-std=c++20', '-O1', '-pthread', '-o', exe, src])
    return exe


def run(traces, psd, templates, fits, fs, fcut=10000.0, precision='f64', ac=True,
        subtract_first=False, scale=1.0, tsan=False, force_p2=False, v2=False, asan=False, generic=False):
    """templates: list of (template, pretrigger, integralnorm); fits: list of (templ, lo, hi, outside)."""
    exe = build_generic() if generic else build(tsan, v2, asan)
    traces = np.ascontiguousarray(traces, dtype=np.float64)
    nev, n = traces.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('<6i', n, nev, len(templates), len(fits), int(ac), int(subtract_first)))
            f.write(struct.pack('<3d', fs, fcut, scale))
            f.write(np.asarray(psd, dtype=np.float64).tobytes())
            for tpl, pre, inorm in templates:
                f.write(struct.pack('<2i', pre, int(inorm)))
                f.write(np.asarray(tpl, dtype=np.float64).tobytes())
            for ft in fits:
                f.write(struct.pack('<4i', *ft))
            f.write(traces.tobytes())
        subprocess.check_call([exe, fin, fout, precision] + (['p2'] if force_p2 else []))
        out = np.fromfile(fout, dtype=np.float64).reshape(nev, 1 + 5 * len(fits))
    return out


def build_nxm(asan=False):
    os.makedirs(os.path.join(HERE, '_build'), exist_ok=True)
    exe = os.path.join(HERE, '_build', 'emu_nxm' + ('_asan' if asan else ''))
    src = os.path.join(HERE, 'emu_nxm.cpp')
    deps = [src] + [os.path.join(HERE, '../../detprocess_b200/csrc', f) for f in
                    ('dp_of_kernel.cuh', 'dp_fft.cuh', 'dp_platform.cuh', 'dp_plan.hpp', 'dp_of2_kernel.cuh',
                     'dp_plan2.hpp', 'dp_f2.cuh', 'dp_nxm_kernel.cuh', 'dp_nxm_plan.hpp')]
    if os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps):
        return exe
    cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-o', exe, src]
    if asan:
        cmd[3:3] = ['-fsanitize=address', '-fno-omit-frame-pointer', '-g']
    subprocess.check_call(cmd)
    return exe


def run_nxm(traces, templates, csd, fs, pretrigger, window=(None, None, False), precision='f64', ac=True, asan=False):
    """traces [B, n, N]; templates [n, m, N]; csd [n, n, N] complex; window (lo, hi, outside) in rolled indices."""
    exe = build_nxm(asan)
    traces = np.ascontiguousarray(traces, dtype=np.float64)
    nev, n, N = traces.shape
    m = templates.shape[1]
    lo, hi, outside = window
    lo = 0 if lo is None else lo
    hi = N if hi is None else hi
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('<9i', N, nev, n, m, pretrigger, int(ac), lo, hi, int(outside)))
            f.write(struct.pack('<d', fs))
            f.write(np.ascontiguousarray(templates, dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(csd, dtype=np.complex128).tobytes())
            f.write(traces.tobytes())
        subprocess.check_call([exe, fin, fout, precision])
        out = np.fromfile(fout, dtype=np.float64).reshape(nev, 4 + 2 * m)
    return out


def run_csd(traces, fs, mask=None, precision='f64', typical_rms=1e-8, asan=False):
    """traces [B, n, N] -> (component sums [n*n, N/2+1] in the CSDPlan.sums layout, accepted-event count)"""
    os.makedirs(os.path.join(HERE, '_build'), exist_ok=True)
    exe = os.path.join(HERE, '_build', 'emu_csd' + ('_asan' if asan else ''))
    src = os.path.join(HERE, 'emu_csd.cpp')
    deps = [src] + [os.path.join(HERE, '../../detprocess_b200/csrc', f) for f in
                    ('dp_of_kernel.cuh', 'dp_fft.cuh', 'dp_platform.cuh', 'dp_plan.hpp', 'dp_of2_kernel.cuh',
                     'dp_plan2.hpp', 'dp_f2.cuh', 'dp_csd_kernel.cuh')]
    if not (os.path.exists(exe) and all(os.path.getmtime(exe) > os.path.getmtime(d) for d in deps)):
        cmd = ['g++', '-std=c++20', '-O1', '-pthread', '-o', exe, src]
        if asan:
            cmd[3:3] = ['-fsanitize=address', '-fno-omit-frame-pointer', '-g']
        subprocess.check_call(cmd)
    traces = np.ascontiguousarray(traces, dtype=np.float64)
    nev, n, N = traces.shape
    mask = np.ones(nev, dtype=np.uint8) if mask is None else np.asarray(mask, dtype=np.uint8)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, 'in.bin'), os.path.join(td, 'out.bin')
        with open(fin, 'wb') as f:
            f.write(struct.pack('<4i', N, nev, n, int(precision == 'f32')))
            f.write(struct.pack('<2d', fs, typical_rms))
            f.write(mask.tobytes())
            f.write(traces.tobytes())
        subprocess.check_call([exe, fin, fout])
        out = np.fromfile(fout, dtype=np.float64)
    return out[:-1].reshape(n * n, N // 2 + 1), int(out[-1])
