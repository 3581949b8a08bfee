"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header declares, derives the same
constants as the reference (no compute call is made: there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from blurry_edges_b200 import _lib, build
from oracle import be_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build_library()
    return _lib.load()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, 'include', 'blurry_edges_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|int64_t|const char\*)\s+(be_\w+)\s*\(', hdr, flags=re.M))
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version(lib):
    assert lib.be_abi_version() == 2


def test_derived_constants_match_reference_formulas(lib):
    for R, S in ((21, 147), (21, 29), (11, 31)):
        cfg = _lib.make_config(R=R, H=S, W=S)
        c = _lib.derive_constants(cfg)
        cam = O.Camera(R=R)
        g = O.Geometry(R=R, H=S, W=S)
        assert c[0] == pytest.approx(cam.numerator, rel=1e-15)
        assert c[1] == pytest.approx(cam.k_const, rel=1e-15)
        assert c[2] == pytest.approx(cam.k_root, rel=1e-15)
        assert c[3] == pytest.approx(cam.k_fac, rel=1e-15)
        assert abs(c[4] - cam.intercept) <= 1.2e-7 * cam.intercept      # fp32 chain, <= 1 ulp
        assert float(torch.tensor(c[5], dtype=torch.float32)) == g.lam
        assert (int(c[6]), int(c[7])) == (g.Hp, g.Wp)


def test_bad_config_is_rejected(lib):
    out = (ctypes.c_double * 8)()
    for kw in (dict(R=23), dict(stride=0), dict(H=10)):
        cfg = _lib.make_config(**kw)
        assert lib.be_derive_constants(ctypes.byref(cfg), out) != 0
        assert lib.be_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_silent_cpu_fallback(lib):
    with pytest.raises(_lib.BlurryEdgesError):
        _lib.Context(_lib.make_config(), 'cpu')
    h = ctypes.c_void_p()
    cfg = _lib.make_config()
    assert lib.be_ctx_create(ctypes.byref(h), ctypes.byref(cfg)) != 0
    assert b'no CUDA device' in lib.be_last_error()
