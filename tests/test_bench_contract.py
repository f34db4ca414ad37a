"""The bench.py output contract, checked on the recorded round-1 lines (profiles/) and on the host-side helpers (CPU only)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
            'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline']


def _load(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_recorded_bench_line_has_the_contract_keys():
    d = _load('r01_bench_final.json')
    for k in REQUIRED:
        assert k in d, k
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert set(['value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step']) <= set(d['e2e'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert d['gpu_launches'] > 0
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-3
    c = d['cpu_baseline']
    assert set(['value', 'unit', 'cores', 'kind', 'sample']) <= set(c) and c['kind'] in ('port', 'reference')
    assert not (set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'})
    # throughput bookkeeping: value = images per step / seconds per step
    assert abs(d['value'] - d['config']['global_batch'] / (d['ms_per_step'] / 1e3)) < 1e-2


def test_recorded_reference_arm_line():
    d = _load('r01_bench_reference_arm.json')
    assert d['impl'] == 'reference' and d['gpu_launches'] == 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['cores'] >= 1


@pytest.mark.parametrize('name,n', [('r01_bench_2gpu.json', 2), ('r01_bench_4gpu.json', 4)])
def test_recorded_multi_gpu_lines_scale_weakly(name, n):
    d, one = _load(name), _load('r01_bench_final.json')
    assert d['n_gpus'] == n and d['config']['global_batch'] == n * one['config']['global_batch']
    assert d['value'] > 0.9 * n * one['value']          # independent shards, one weight broadcast: near-linear


def test_host_thread_budget_is_sane():
    import bench
    n = bench._host_threads()
    assert 1 <= n <= 64 and n <= (os.cpu_count() or 1)
