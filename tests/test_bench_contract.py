"""bench.py contract checks that need no GPU: the reference arm (the reference graph over the oracle port on the host CPU) prints one
JSON line with the agreed keys; ranks other than 0 print nothing and exit 0; our arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line():
    r = _run(['--impl', 'reference', '--cpu-batch', '1', '--steps', '1', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train_meshes_per_sec' and d['unit'] == 'meshes/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'meshes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_reference_arm_uses_all_cores_and_never_maps_the_product_library():
    """Under torchrun OMP_NUM_THREADS=1 is exported: the arm must still use every host core.  It runs oracle/ only, so neither the
    product package nor libgeniconet_b200.so may appear in the process."""
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; "
            "t, threads = bench.cpu_reference_step_time('ico2ico', 5, 1, 1, 0); "
            "maps = open('/proc/self/maps').read(); "
            "print(json.dumps({'threads': threads, 'so': 'libgeniconet' in maps, "
            "'mods': [m for m in sys.modules if m.startswith('geniconet_b200')]}))" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['threads'] == (os.cpu_count() or 1) and not d['so'] and not d['mods'], d


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_our_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a CUDA device is present')
    r = _run(['--steps', '1', '--warmup', '1', '--no-cpu-baseline'])
    assert r.returncode != 0 and 'no CUDA device' in (r.stderr + r.stdout)
