"""The reference arm of bench.py runs on the host cores, so its JSON contract can be checked here without a GPU: one line,
`impl: reference`, the metric / unit / config of the GPU arm, `cpu_baseline` describing the run, zero-copy `e2e`."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, SSDK_REF_SAMPLE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'SSD300 images/sec (fwd+decode)' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['config']['workload'].startswith('SSD300 inference, batch 32')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: the product arm must not print a number on a machine without CUDA."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1', '--no-micro', '--no-cpu'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(l.strip().startswith('{') for l in r.stdout.splitlines())
