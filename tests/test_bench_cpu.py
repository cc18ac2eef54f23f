"""CPU: the reference arm of bench.py (the one leg allowed to execute oracle/) prints exactly one
JSON line on stdout with the contract's keys; ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--cpu-scale', '6', '--cpu-queries', '500']
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'query_pts_per_sec' and d['unit'] == 'pts/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['dtype'] == 'f64'
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_are_silent():
    res = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert res.returncode == 0 and res.stdout.strip() == ''
