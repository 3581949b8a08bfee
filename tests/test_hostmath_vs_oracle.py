"""The fp32 arithmetic the CUDA kernels use (csrc/be_math.cuh) compiled for the host and checked against the fp64
oracle - the closest thing to a kernel parity test that runs without a GPU.  Tolerances are the ones the GPU parity
tests use (see tests/test_gpu_inference.py for the derivation)."""
import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, MAPS, geom, inference_inputs, planar_pair, relmax
from oracle import be_oracle as O
from oracle import hostmath

CAM = O.Camera()


@pytest.mark.parametrize('gname', list(GEOMS))
def test_pass_a(gname):
    S = GEOMS[gname]
    g = geom(S)
    img = planar_pair(synth.image_pairs(1, S, S, seed=3))[0]
    est = synth.est_local(2, g.L, seed=5)
    ref = O.colors_only(est.to(F64), img.to(F64), g).numpy()
    got = hostmath.colors(est, img, g, CAM)
    assert relmax(got, ref) < 2e-5


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('densify', [None, 'w'])
@pytest.mark.parametrize('kind,tol', [('normal', 1e-5), ('stress', 5e-4)])
def test_pass_b(gname, densify, kind, tol):
    g, est, img = inference_inputs(gname, kind, F32)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    got = hostmath.render_fold(est, img, g, CAM, 10.39, densify)
    for name, r, o in zip(MAPS, ref, got):
        e = relmax(o, r.numpy())
        lim = tol if name not in ('sharp',) else max(tol, 2e-3)   # eta=1e-4 render: edge pixels amplify 1-ulp distance noise
        assert e < lim, (name, e)
