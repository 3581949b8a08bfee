"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference, the one place outside tests/ and smoke()
that may execute oracle/) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run without a GPU
instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', ''))
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-pairs', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'patches/s' and d['higher_is_better'] is True
    for k in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['value'] > 0 and d['steps'] == 1 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    # both arms print the SAME config (the bounded sample the CPU arm times is named in cpu_baseline.sample)
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == bench.bench_config(32) and d['metric'] == bench.METRIC
    assert 'pair' in d['cpu_baseline']['sample']


def test_product_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return                                    # on a GPU box the product arm is what test_gpu_* and the driver run
    r = _run('--steps', '1', '--warmup', '0')
    assert r.returncode != 0
    assert 'no CUDA device' in (r.stderr + r.stdout) and 'no CPU fallback' in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith('{')]      # no bench line from a fallback
