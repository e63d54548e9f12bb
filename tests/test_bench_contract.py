"""The JSON lines of bench.py follow the driver's contract: the committed B200 line (profiles/) and a live run of the
reference arm, which needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'cpu_baseline'}


def test_committed_b200_line_is_well_formed():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_mnist_bf16_1gpu.json')))
    assert BASE_KEYS | {'roofline', 'clocks'} <= set(d)
    assert d['metric'] == 'train_sequences_per_sec' and d['unit'] == 'sequences/s' and d['higher_is_better'] is True
    assert d['scaling'] == 'weak' and d['vs_baseline'] is None and d['dtype'] == 'bf16' and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['warmup'] >= 3 and d['gpu_launches'] > 0
    assert abs(d['value'] - d['config']['global_batch'] / (d['ms_per_step'] * 1e-3)) < 1e-6 * d['value']
    r = d['roofline']
    assert r['bound'] in ('tensor', 'hbm') and r['unit'] in ('TFLOP/s', 'GB/s') and r['traffic'] is not None
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['peak'] > 0 and 'peak_source' in r          # MEASURED_PEAKS.json is rewritten per pod: not compared here
    e = d['e2e']
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] <= d['value'] * 1.02
    c = d['cpu_baseline']
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(d['clocks'])
    assert not {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(d['clocks']['reasons'])


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert BASE_KEYS | {'impl'} <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'train_sequences_per_sec' and d['gpu_launches'] == 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['cores'] >= 1
