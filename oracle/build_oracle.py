"""Compile the C++ side of the oracle (TEST INFRASTRUCTURE): oracle/be_hostmath.cpp -> oracle/_build/libbe_hostmath.so.
The reference is pure Python, so there is no oracle/_ref to compile."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build', 'libbe_hostmath.so')
SRC = os.path.join(HERE, 'be_hostmath.cpp')
DEPS = [os.path.join(HERE, '..', 'blurry_edges_b200', 'csrc', f) for f in ('be_math.cuh', 'be_pack.cuh')]


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(f) for f in [SRC] + DEPS):
        return OUT
    cmd = ['g++', '-O2', '-std=c++17', '-fopenmp', '-fPIC', '-shared', '-x', 'c++', SRC, '-o', OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('g++ failed building the oracle host-math library')
    return OUT


if __name__ == '__main__':
    print(build(force=True))
