"""Row-sharded path on 2+ GPUs of one box (skipped on a single-GPU box):
torchrun launches tests/dist_worker.py, one rank per GPU over NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('nranks', [2, 4, 8])
def test_sharded_assembly_and_cg(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip('needs {0} GPUs'.format(nranks))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node', str(nranks), '--master-addr', '127.0.0.1',
                          '--master-port', str(29600 + nranks),
                          os.path.join(ROOT, 'tests', 'dist_worker.py')],
                         capture_output=True, text=True, env=env, timeout=420)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(' ok') == nranks
