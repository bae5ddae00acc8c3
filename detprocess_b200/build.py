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
LIB = os.path.join(OUT_DIR, 'libdetprocess_b200.so')

SOURCES = ['dp_capi.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O3', '-lineinfo',
              '--fmad=true', '-Xcompiler', '-fPIC,-O2', '-shared']


def _deps():
    out = []
    for root in (CSRC, os.path.join(PKG, '..', 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cu', '.cuh', '.hpp', '.h')):
                out.append(os.path.join(root, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=True):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libdetprocess_b200.so')
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ['-Xptxas', '-v', '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print('[detprocess_b200.build]', ' '.join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(OUT_DIR, 'build.log'), 'w') as f:
        f.write(res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError('nvcc failed building libdetprocess_b200.so')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
