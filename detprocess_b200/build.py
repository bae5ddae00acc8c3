"""
Build libdetprocess_b200.so in-tree with nvcc for sm_100a.

    python -m detprocess_b200.build [--force]

The library is a plain C-ABI shared object (include/detprocess_b200.h); it links the
CUDA runtime statically and has no torch / python dependency.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OUT_DIR = os.path.join(PKG, '_C')
# development A/B builds: DP_LIB_NAME=<name>.so DP_BUILD_DEFS='-DX=1 ...' python -m detprocess_b200.build --force
LIB = os.path.join(OUT_DIR, os.environ.get('DP_LIB_NAME', 'libdetprocess_b200.so'))
EXTRA_DEFS = os.environ.get('DP_BUILD_DEFS', '').split()

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O3', '-lineinfo',
              '--fmad=true', '-Xcompiler', '-fPIC,-O2', '-Xptxas', '-v']

# (object name, source, extra defines): the OF kernel is instantiated per
# (precision, input type) in separate translation units so they compile in parallel
UNITS = [('dp_capi', 'dp_capi.cu', []), ('dp_ofg_inst', 'dp_ofg_inst.cu', []), ('dp_band_inst', 'dp_band_inst.cu', [])] + [
    (f'dp_of2_inst_p{p}_{i}', 'dp_of2_inst.cu', [f'-DDP_INST_PREC={p}', f'-DDP_INST_IN={i}'])
    for p in (1, 0) for i in (0, 1, 2, 3, 4, 5)] + [
    (f'dp_trig_inst_p{p}', 'dp_trig_inst.cu', [f'-DDP_INST_PREC={p}']) for p in (1, 0)] + [
    (f'dp_nxm_inst_p{p}_{c}', 'dp_nxm_inst.cu', [f'-DDP_INST_PREC={p}', f'-DDP_INST_NCH={c}']) for p in (1, 0) for c in (1, 2, 3, 4)] + [
    (f'dp_csd_inst_p{p}_{c}', 'dp_csd_inst.cu', [f'-DDP_INST_PREC={p}', f'-DDP_INST_NCH={c}']) for p in (1, 0) for c in (2, 3, 4)] + [
    (f'dp_of_inst_p{p}_{i}', 'dp_of_inst.cu', [f'-DDP_INST_PREC={p}', f'-DDP_INST_IN={i}'])
    for p in (1, 0) for i in (0, 1, 2)]


def _deps():
    out = []
    for root in (CSRC, os.path.join(PKG, '..', 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cu', '.cuh', '.hpp', '.h')):
                out.append(os.path.join(root, f))
    return out


def _source_hash():
    """sha256 over the contents of every source / header and the compiler flags: the library is rebuilt when what it
    was built from changed, not when a copy of the tree shuffled the file times"""
    import hashlib
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS + EXTRA_DEFS + [repr(u) for u in UNITS]).encode())
    for path in sorted(_deps()):
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _stamp_path():
    return LIB + '.stamp'


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(_stamp_path()):
        return True
    with open(_stamp_path()) as f:
        return f.read().strip() != _source_hash()


def _unit_deps(src, seen=None):
    """the source and every local header it includes (transitively)"""
    import re
    seen = set() if seen is None else seen
    path = os.path.normpath(src)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', f.read(), flags=re.M):
            _unit_deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libdetprocess_b200.so')
    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, 'obj' + ('_' + os.path.basename(LIB) if EXTRA_DEFS else ''))
    os.makedirs(obj_dir, exist_ok=True)

    only = os.environ.get('DP_BUILD_UNITS', '').split()
    main_obj_dir = os.path.join(OUT_DIR, 'obj')

    def compile_unit(unit):
        name, src, defs = unit
        obj = os.path.join(obj_dir, name + '.o')
        if EXTRA_DEFS and only and not any(o in name for o in only):
            # development A/B build: units the experiment does not touch come from the main build
            return name, os.path.join(main_obj_dir, name + '.o'), '# reused from the main build', 0, ''
        cmd = [nvcc] + NVCC_FLAGS + EXTRA_DEFS + defs + ['-c', '-o', obj, os.path.join(CSRC, src)]
        # an object newer than its source, every header it includes and this script is reused
        deps = list(_unit_deps(os.path.join(CSRC, src))) + [os.path.abspath(__file__)]
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) > os.path.getmtime(d) for d in deps):
            return name, obj, ' '.join(cmd) + '   # up to date', 0, ''
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return name, obj, ' '.join(cmd), res.returncode, res.stdout

    if verbose:
        print(f'[detprocess_b200.build] nvcc sm_100a: {len(UNITS)} translation units', flush=True)
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    log = []
    ok = True
    for name, obj, cmd, rc, out in results:
        log.append(f'### {cmd}\n{out}')
        ok &= rc == 0
    link = [nvcc, '-shared', '-o', LIB] + [r[1] for r in results]
    if ok:
        res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(f'### {" ".join(link)}\n{res.stdout}')
        ok = res.returncode == 0
    with open(os.path.join(OUT_DIR, 'build.log' if not EXTRA_DEFS else 'build_' + os.path.basename(LIB) + '.log'), 'w') as f:
        f.write('\n'.join(log))
    if not ok:
        sys.stderr.write('\n'.join(log)[-8000:])
        raise RuntimeError('nvcc failed building libdetprocess_b200.so')
    with open(_stamp_path(), 'w') as f:
        f.write(_source_hash())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
