"""Data-parallel parity on real GPUs: skipped on boxes with fewer than two devices."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world', [2, 4, 8])
def test_graphed_dp_step_averages_rank_gradients(world):
    """Post-allreduce gradients of the CAPTURED step == mean of the per-rank gradients, all-reduce launched from inside the
    fused chain's backward; also: every rank ends with the same gradients."""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d CUDA devices' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'diag', 'dp_parity_worker.py'), 'ico2ico', '6']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    d = json.loads(line)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'dp_parity_n%d.json' % world), 'w') as fh:
        fh.write(line + '\n')
    assert d['finite'] and d['params_handed_over_early'] > 0
    # the only difference between the two sides is the order in which NCCL adds `world` fp32 numbers
    assert d['max_abs_err_over_max_grad'] <= 1e-5 and d['worst_param_rel_l2'] <= 1e-4, d
