"""Import mechanics of tools/run_reference_script.py: with the shim installed, the unmodified reference
scripts bind THIS package's classes.  Runs only where the reference is mounted (the build container); there is no GPU
there, so the proof that the override is live is that constructing the script's own subclass reaches our constructor
and fails loudly with the no-CPU-fallback error."""
import importlib
import os
import sys

import pytest
import torch

import refimport

pytestmark = pytest.mark.skipif(not refimport.available(), reason='reference not mounted')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def shim():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    saved = {k: sys.modules.get(k) for k in ('utils', 'blurry_edges_test', 'global_training', 'local_training')}
    for k in saved:
        sys.modules.pop(k, None)
    import run_reference_script as rrs
    mod = rrs.install_shim(refimport.REF)
    yield mod
    for k, v in saved.items():
        sys.modules.pop(k, None)
        if v is not None:
            sys.modules[k] = v
    for k in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')]:
        sys.modules.pop(k, None)


def test_scripts_bind_our_classes(shim):
    import blurry_edges_b200 as be
    assert shim.PostProcessGlobalBase is be.PostProcessGlobalBase and shim.DepthEtas is be.DepthEtas
    assert callable(shim.get_args) and callable(shim.eval_depth)            # pass-throughs stay the reference's
    bt = importlib.import_module('blurry_edges_test')
    gtm = importlib.import_module('global_training')
    ltm = importlib.import_module('local_training')
    assert issubclass(bt.PostProcess, be.PostProcessGlobalBase)
    assert issubclass(gtm.GlobalLoss, be.PostProcessGlobalBase)
    assert issubclass(ltm.LocalLoss, be.PostProcessLocalBase)
    assert bt.DepthEtas is be.DepthEtas


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_script_subclass_reaches_our_constructor_and_fails_loudly(shim):
    import blurry_edges_b200 as be
    bt = importlib.import_module('blurry_edges_test')
    args = refimport.get_args('eval', ['--cuda', 'cpu'])
    with pytest.raises(be.BlurryEdgesError, match='CUDA'):
        bt.PostProcess(args, None, torch.device('cpu'))
    with pytest.raises(be.BlurryEdgesError, match='CUDA'):
        bt.DepthEtas(args, torch.device('cpu'))


def test_patch_reference_smish_swaps_the_class():
    """activations.patch_reference_smish replaces models.local_stage.Smish (models/local_stage.py:4-6); a LocalStage built
    afterwards holds SmishFused modules."""
    ls = refimport.module('models.local_stage')
    from blurry_edges_b200 import SmishFused, patch_reference_smish
    orig = patch_reference_smish(ls)
    try:
        net = ls.LocalStage()
        acts = [m for m in net.modules() if isinstance(m, SmishFused)]
        assert len(acts) > 0 and not any(isinstance(m, orig) for m in net.modules())
    finally:
        ls.Smish = orig
