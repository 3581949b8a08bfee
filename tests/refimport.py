"""Import the UNMODIFIED reference (read-only at /root/reference) as a test oracle.

Only available in the build container; on the GPU box /root/reference does not exist and
everything here reports `available() == False` (tests then rely on tests/golden/*.npz and on
oracle/be_oracle.py, which the golden vectors pin)."""
from __future__ import annotations

import importlib
import os
import sys
from contextlib import contextmanager

REF = os.environ.get('BE_REFERENCE', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_stubs')


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'utils', 'postprocessing_loss.py'))


@contextmanager
def _argv(argv):
    old = sys.argv
    sys.argv = ['x'] + [str(a) for a in argv]
    try:
        yield
    finally:
        sys.argv = old


def _ensure_path():
    try:
        import matplotlib  # noqa: F401
    except Exception:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def module(name: str):
    """Import a reference module (`utils`, `blurry_edges_test`, `global_training`, ...)."""
    _ensure_path()
    return importlib.import_module(name)


def get_args(mode: str, argv=(), big: bool = False):
    """utils/args.py:get_args with a patched sys.argv."""
    u = module('utils')
    with _argv(argv):
        return u.get_args(mode, big=big) if big else u.get_args(mode)


def to_dtype(obj, depth_cal, dtype):
    """Cast the constant buffers of a reference PostProcess*/DepthEtas pair."""
    for k in ('x', 'y', 'ridge', 'num_patches', 'sobel_x', 'sobel_y'):
        if hasattr(obj, k):
            setattr(obj, k, getattr(obj, k).to(dtype))
    if depth_cal is not None:
        for k in ('intercept', 'theta_mid', 'theta_wng'):
            setattr(depth_cal, k, getattr(depth_cal, k).to(dtype))
    return obj
