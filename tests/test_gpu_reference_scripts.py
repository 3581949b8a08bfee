"""The reference's UNMODIFIED entry scripts, run end to end on a B200 through this library (north_star: "local_training.py,
global_training.py and blurry_edges_test*.py run unchanged on it").

Needs a staged copy of the reference under baseline/_ref/Blurry-Edges (tools/stage_reference.py; git-ignored, travels with the
gpurun snapshot) - skipped otherwise.  Weights and datasets are synthetic (tools/make_assets.py: random-init state_dicts, scenes of
the reference's generator).  Every script is run as a subprocess
  cuda   : python <script> --cuda cuda:0                                  the reference's own classes, eager PyTorch on the GPU
  shim   : python tools/run_reference_script.py <script> --cuda cuda:0    every helper-class METHOD is one kernel of this library
  fused  : python tools/run_reference_script.py --fused <script> ...      the script's composite class resolves to the fused sibling
  cpu    : python <script> --cuda cpu                                     the reference's CPU path (evaluation script only: slow)
and what the scripts themselves print / save is compared: the depth metrics at the precision the script prints them (3 decimals;
tests/test_gpu_inference.py holds them to 5e-5 in-process), validation losses and saved weights after one epoch of training, the
pre-calculated global-stage inputs.  A summary goes to gpurun_out/scripts_on_b200.json."""
import json
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref', 'Blurry-Edges')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'blurry_edges_test.py')),
                                                  reason='reference not staged under baseline/_ref (tools/stage_reference.py)')]
SUMMARY = {}


def _run(mode, script, argv, timeout=900):
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(ROOT, 'tests', '_stubs'), env.get('PYTHONPATH', '')])     # matplotlib stand-in
    path = os.path.join(REF, script)
    dev = 'cpu' if mode == 'cpu' else 'cuda:0'
    if mode in ('cpu', 'cuda'):
        cmd = [sys.executable, path]
    else:
        cmd = [sys.executable, os.path.join(ROOT, 'tools', 'run_reference_script.py')] + (['--fused'] if mode == 'fused' else []) + [path]
    t0 = time.time()
    r = subprocess.run(cmd + ['--cuda', dev] + [str(a) for a in argv], capture_output=True, text=True, cwd=REF, env=env, timeout=timeout)
    assert r.returncode == 0, f'{mode} {script} failed:\n{r.stdout[-1500:]}\n{r.stderr[-3000:]}'
    launches = re.search(r'kernel_launches=(\d+)', r.stderr)
    if mode in ('shim', 'fused'):
        assert launches and int(launches.group(1)) > 0, 'the script did not reach this library\'s kernels'
    return r.stdout, r.stderr, time.time() - t0


@pytest.fixture(scope='module')
def assets(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('assets'))
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'make_assets.py'), REF, out], check=True, capture_output=True, timeout=600)
    yield out
    dst = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(dst) and SUMMARY:
        with open(os.path.join(dst, 'scripts_on_b200.json'), 'w') as f:
            json.dump(SUMMARY, f, indent=1)


def _metrics(stdout):
    per = re.findall(r'--- Error metrics: (.*)', stdout)
    avg = re.findall(r'Average metrics for whole dataset: (.*)', stdout)
    assert per and avg, stdout[-800:]
    return per, avg[0]


@pytest.mark.parametrize('densify', [None, 'w'])
def test_blurry_edges_test_py(assets, densify, tmp_path):
    """blurry_edges_test.py:174-202 (+ --densify w): the printed depth metrics of the shim and fused runs equal those of the
    reference's own classes on the same GPU, and (default rule) of the reference's CPU path."""
    base = ['--model_path', f'{assets}/weights', '--data_path', f'{assets}/eval'] + (['--densify', densify] if densify else [])
    got = {}
    modes = ['cuda', 'shim', 'fused'] + (['cpu'] if densify is None and os.environ.get('BE_SCRIPTS_CPU', '1') != '0' else [])
    for mode in modes:
        out, err, dt = _run(mode, 'blurry_edges_test.py', base + ['--log_path', str(tmp_path / mode)])
        got[mode] = _metrics(out)
        got[mode + '_s'] = round(dt, 1)
        assert os.path.isfile(tmp_path / mode / 'visualizations' / '1.png')
        if mode == 'fused':
            assert "('PostProcess', 'PostProcessFused')" in err
    SUMMARY[f'blurry_edges_test.py{" --densify w" if densify else ""}'] = {k: v for k, v in got.items()}
    for mode in modes[1:]:
        assert got[mode] == got['cuda'], (mode, got[mode], got['cuda'])


def _val_losses(log_dir, name):
    txt = open(os.path.join(log_dir, name)).read()
    rows = re.findall(r'^(\d+)\s+([0-9.eE+-]+)\s+\d+\s+[0-9.eE+-]+\s*$', txt, flags=re.M)
    assert rows, txt[-500:]
    return [float(v) for _, v in rows]


def test_global_training_py(assets, tmp_path):
    """global_training.py:173-225, two epochs on 4 + 2 scenes (batch 2): set_seed(1898, deterministic=True), xavier init, AdamW steps
    through the loss, validation with final gammas.  The shim / fused runs must give the validation losses and the saved weights of
    the reference's own classes on the same GPU (identical up to fp32 rounding amplified by two optimiser epochs)."""
    got = {}
    for mode in ('cuda', 'shim', 'fused'):
        d = tmp_path / mode
        out, err, dt = _run(mode, 'global_training.py', ['--data_path', f'{assets}/train', '--log_path', d, '--model_path', d,
                                                        '--epoch_num', 2, '--batch_size', 2])
        got[mode] = dict(val=_val_losses(d, 'exp_global_stage_training.txt'), s=round(dt, 1),
                         w=torch.load(d / 'best_run_exp_global_stage.pth', map_location='cpu'))
        if mode == 'fused':
            assert "('GlobalLoss', 'GlobalLossFused')" in err
    ref = got['cuda']
    SUMMARY['global_training.py'] = {m: dict(val_loss=got[m]['val'], seconds=got[m]['s']) for m in got}
    for mode in ('shim', 'fused'):
        np.testing.assert_allclose(got[mode]['val'], ref['val'], rtol=2e-4)
        worst = max(float((got[mode]['w'][k] - ref['w'][k]).abs().max()) for k in ref['w'])
        SUMMARY['global_training.py'][mode]['max_abs_weight_diff_vs_reference_classes'] = worst
        assert worst < 2e-4, (mode, worst)                     # AdamW steps of 1e-4 with sign-like updates: a flipped rounding moves a weight by <= 2 lr


def test_local_training_py(assets, tmp_path):
    """local_training.py:68-121, two epochs on 128 + 64 patches (batch 64)."""
    got = {}
    for mode in ('cuda', 'shim', 'fused'):
        d = tmp_path / mode
        out, err, dt = _run(mode, 'local_training.py', ['--data_path', f'{assets}/train/patches', '--log_path', d, '--model_path', d, '--epoch_num', 2])
        got[mode] = dict(val=_val_losses(d, 'exp_local_stage_training.txt'), s=round(dt, 1))
        if mode == 'fused':
            assert "('LocalLoss', 'LocalLossFused')" in err
    SUMMARY['local_training.py'] = {m: dict(val_loss=got[m]['val'], seconds=got[m]['s']) for m in got}
    for mode in ('shim', 'fused'):
        np.testing.assert_allclose(got[mode]['val'], got['cuda']['val'], rtol=5e-4)


def test_global_data_pre_cal_py(assets, tmp_path):
    """global_data_pre_cal.py:52-69: LocalStage + ridge colours -> params_src_{train,val}.npy (the script writes into --data_path)."""
    got = {}
    for mode in ('cuda', 'shim', 'fused'):
        d = tmp_path / mode
        shutil.copytree(f'{assets}/train', d, ignore=shutil.ignore_patterns('patches', 'params_src_*'))
        out, err, dt = _run(mode, 'global_data_pre_cal.py', ['--data_path', d, '--model_path', f'{assets}/weights'])
        got[mode] = (np.load(d / 'params_src_train.npy'), np.load(d / 'params_src_val.npy'), round(dt, 1))
    SUMMARY['global_data_pre_cal.py'] = {}
    for mode in ('shim', 'fused'):
        worst = max(float(np.abs(a - b).max()) for a, b in zip(got[mode][:2], got['cuda'][:2]))
        SUMMARY['global_data_pre_cal.py'][mode] = dict(max_abs_diff_vs_reference_classes=worst, seconds=got[mode][2])
        assert worst < 5e-3, (mode, worst)        # colours: the reference's fp32 trace-formula inverse is itself 2e-3 off fp64


def test_blurry_edges_test_big_py(assets, tmp_path):
    """blurry_edges_test_big.py:222-241 at 235x235 (2x2 blocks): method-granularity shim (the script folds with its own nn.Fold helpers)."""
    base = ['--model_path', f'{assets}/weights', '--data_path', f'{assets}/big', '--big_img_size', 235, 235]
    got = {}
    for mode in ('cuda', 'shim'):
        out, err, dt = _run(mode, 'blurry_edges_test_big.py', base + ['--log_path', str(tmp_path / mode)])
        got[mode] = _metrics(out)
        got[mode + '_s'] = round(dt, 1)
    SUMMARY['blurry_edges_test_big.py 235x235'] = got
    assert got['shim'] == got['cuda'], got
