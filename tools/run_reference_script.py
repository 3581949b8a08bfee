#!/usr/bin/env python
"""Run an UNMODIFIED entry script of the reference on this library's kernels.

  python tools/run_reference_script.py [--fused] /path/to/Blurry-Edges/blurry_edges_test.py --cuda cuda:0 --data_path ... [script args]

Mechanism (blurry_edges_b200/shim.py): the reference's scripts do `from utils import DepthEtas,
PostProcessGlobalBase, ...` and, being run as files, always find their own `utils` package first.  This launcher imports that
package, copies its namespace into a shim module in which the four path classes are replaced by blurry_edges_b200's kernel-backed
mirrors, installs the shim as sys.modules['utils'], and runs the script with runpy.  models/, data/, args, metrics and
visualisation stay the reference's own.  `--fused` (before the script path) additionally makes the script's own composite class
(PostProcess / GlobalLoss / LocalLoss) resolve to the fused sibling at class-definition time, so the script's main loop drives the
fused kernels; without it every helper-class METHOD runs as one kernel of this library."""
from __future__ import annotations

import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def install_shim(ref_dir: str, fused: bool = False):
    from blurry_edges_b200 import shim
    return shim.install(ref_dir, fused=fused)


def main():
    argv = sys.argv[1:]
    fused = False
    if argv and argv[0] == '--fused':
        fused, argv = True, argv[1:]
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    ref_dir = os.path.dirname(script)
    install_shim(ref_dir, fused=fused)
    sys.argv = [script] + argv[1:]
    try:
        runpy.run_path(script, run_name='__main__')
    finally:
        from blurry_edges_b200 import _lib, shim
        print(f'[run_reference_script] mode={"fused" if fused else "methods"} substituted={shim.substituted} '
              f'kernel_launches={_lib.launch_count() if _lib._lib is not None else 0}', file=sys.stderr)


if __name__ == '__main__':
    main()
