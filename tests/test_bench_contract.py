"""bench.py's output contract: one JSON line on stdout with the keys the driver reads.  The CPU
(reference) arm runs anywhere; the native arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
             'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches'}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout.decode()


def test_reference_arm_prints_one_json_line():
    out = _run(['--impl', 'reference', '--steps', '3', '--warmup', '3', '--envs', '64'])
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'env-steps/sec' and d['unit'] == 'env-steps/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_are_silent():
    out = _run(['--impl', 'reference', '--gpus', '2', '--steps', '3', '--warmup', '3', '--envs', '64'],
               env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''


@pytest.mark.gpu
def test_native_arm_prints_one_json_line():
    out = _run(['--steps', '20', '--warmup', '3', '--envs', '512', '--no-configs'])
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and 'impl' not in d
    assert d['n_gpus'] == 1 and d['steps'] == 20 and d['gpu_launches'] >= 20
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert d['cpu_baseline']['value'] > 0 and d['cpu_baseline']['kind'] == 'port'
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] == 512 * 8 and e['d2h_bytes_per_step'] == 512 * (519 * 4 + 5)
    assert d['clocks']['sm_mhz'] > 0 and d.get('short_run') is True
