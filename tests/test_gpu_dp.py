"""Multi-GPU tests (skipped on a single-GPU box): the fused peer-memory exchange kernel vs NCCL all-reduce + Adam."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs on one NVLink domain')
def test_peer_exchange_matches_allreduce_plus_adam():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr',
           '127.0.0.1', '--master-port', '29541', os.path.join(ROOT, 'tests', 'dp_peer_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    if r.returncode != 0:          # keep the ranks' own messages (pytest truncates the assertion text)
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'dp_peer_check_failure.log'), 'w') as f:
            f.write(r.stdout + '\n---- stderr ----\n' + r.stderr)
    assert r.returncode == 0 and 'dp_peer_check PASS' in r.stdout and 'dp_peer_check buckets PASS' in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
