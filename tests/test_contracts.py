"""CPU: repository contracts the driver and the judge rely on."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under d3fields_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, 'd3fields_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn), encoding='utf-8', errors='replace').read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{fn} imports oracle'
                assert 'libd3f_oracle' not in src and 'd3f_oracle_eval' not in src, f'{fn} references the C oracle'


def test_reference_sources_are_not_vendored():
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in ('.git', 'gpurun_out', '__pycache__', '_lib', '_build', '_bin')]
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn), encoding='utf-8', errors='replace').read()
                assert 'grounded_instance_sam_new_ver' not in src and 'align_instance_mask_v3' not in src or fn == 'test_contracts.py'


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the reference's CPU operator sequence on a bounded sample and prints one JSON
    line with the keys the driver reads (kept tiny here: 1 step, 1 warm-up)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['unit'] == 'Mpts/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config']


def test_graft_entry_build_runs_on_cpu():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(os.path.join(ROOT, 'd3fields_b200', '_lib', 'libd3f.so'))
    assert os.path.exists(os.path.join(ROOT, 'oracle', '_build', 'libd3f_oracle.so'))


def test_algorithmic_bytes_match_the_survey():
    """bench.py's roofline numerator is SURVEY.md §8d's formula: 4.168 GB for cfg2a, 9.15 GB for cfg2b, 63.7 MB for cfg3 (u8)."""
    sys.path.insert(0, ROOT)
    import bench
    n, V, H, W = 1_000_000, 4, 480, 640
    assert abs(bench.algorithmic_bytes(n, V, H, W, [(48, 64, 1024, 4)]) - 4.168e9) < 2e6
    assert abs(bench.algorithmic_bytes(n, V, H, W, [(480, 640, 1024, 4)]) - 9.15e9) < 1e7
    assert abs(bench.algorithmic_bytes(n, V, H, W, [(480, 640, 8, 1)]) - 63.7e6) < 1e5
    assert bench.algorithmic_bytes(n, V, H, W, []) == 12 * n + V * H * W * 4 + 5 * n
