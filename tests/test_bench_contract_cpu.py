"""CPU only: the reference arm of bench.py (the oracle port timed on host cores) prints exactly one JSON line with the
contract's keys, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, SCD_BENCH_CPU_ROWS='512', **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'ms' and d['higher_is_better'] is False and d['vs_baseline'] is None
    assert d['metric'].startswith('ms per naming round') and d['value'] > 0 and d['ms_per_step'] == d['value']
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['data'] == 'synthetic' and d['scaling'] == 'strong'
    assert d['config']['workload'].startswith('C2: 127000x768') and 'model' not in d['config']
    cb = d['cpu_baseline']
    # "reference" where the checkout is importable (the E-step then runs the reference's own pairwise_distance), "port" elsewhere
    want_kind = 'reference' if os.path.isdir('/root/reference/local_utils') else 'port'
    assert cb['kind'] == want_kind and cb['kind_detail'] and cb['cores'] >= 1 and cb['value'] == d['value'] and '512 of 127000 rows' in cb['sample']
    r2 = _run({'SCD_REFERENCE_DIR': '/nonexistent'})
    assert json.loads(r2.stdout.strip())['cpu_baseline']['kind'] == 'port'
    assert d['e2e'] == {'value': d['value'], 'unit': 'ms', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''
