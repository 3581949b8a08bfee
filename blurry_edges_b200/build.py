"""Build libblurry_edges_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object loaded with ctypes, see _lib.py)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libblurry_edges_b200.so')
SOURCES = ['be_kernels.cu', 'be_run3.cu', 'be_train.cu', 'be_loss2.cu', 'be_ops.cu', 'be_capi.cu']
HEADERS = ['be_math.cuh', 'be_pack.cuh', 'be_internal.h', os.path.join('..', '..', 'include', 'blurry_edges_b200.h')]
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def _nvcc() -> str:
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libblurry_edges_b200.so')


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get('BE_NVCC_EXTRA', '').split()          # experiments: e.g. BE_NVCC_EXTRA=-DBE_LOSS2_RING=0
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, '-o', LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libblurry_edges_b200.so')
    with open(os.path.join(HERE, 'build_ptxas.log'), 'w') as f:
        f.write(r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
