"""CPU: the reference arm of bench.py (the reference's own updateH/updateE -- the real modules from
oracle/_ref/ when present, else the oracle port -- timed on the host cores) prints one JSON line
with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, IES_BENCH_REF_MAX_RANKS='2', IES_BENCH_REF_FULL='0')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['unit'] == 'Mcell-updates/s' and j['higher_is_better'] is True
    assert j['value'] > 0 and j['cpu_baseline']['kind'] in ('reference', 'port') and j['cpu_baseline']['cores'] >= 1
    assert j['e2e'] == {'value': j['value'], 'unit': j['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_reference_arm_mie_config():
    env = dict(os.environ, IES_BENCH_REF_MAX_RANKS='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--config', 'mie'], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert j['impl'] == 'reference' and j['value'] > 0 and 'mie' in j['config']['workload']
