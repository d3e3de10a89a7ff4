"""Multi-GPU pieces that need real devices (skipped on a one-GPU box): the peer-memory gradient all-reduce of
csrc/allreduce.cu against NCCL, run under torchrun on two GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_allreduce_matches_nccl(lib_built):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29577', os.path.join(ROOT, 'profiles', 'check_peer_allreduce.py')],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'PEER ALLREDUCE OK' in r.stdout
