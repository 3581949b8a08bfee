#!/usr/bin/env python
"""Stage a copy of the (read-only) reference under baseline/_ref/Blurry-Edges so that it travels to the GPU box with the gpurun
snapshot (baseline/_ref/ is git-ignored, never committed): tests/test_gpu_reference_scripts.py runs the UNMODIFIED scripts from
there, on the CPU, on the GPU with the reference's own classes, and on the GPU through this library's shim."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('BE_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref', 'Blurry-Edges')


def main():
    if not os.path.isfile(os.path.join(SRC, 'blurry_edges_test.py')):
        raise SystemExit(f'{SRC}: reference not found')
    if os.path.exists(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.pyc', 'pretrained_weights', 'assets', 'logs'))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f'staged {n} files of {SRC} under {DST} (git-ignored)')


if __name__ == '__main__':
    sys.exit(main())
