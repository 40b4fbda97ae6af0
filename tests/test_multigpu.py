"""Multi-GPU parity: the fused gather + all-gather over NVLink peer memory must leave, on every
GPU, the same bits as a single-GPU precompute of the whole link list.  Needs >= 2 GPUs (skipped otherwise);
run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_exchange_equals_single_gpu_bit_for_bit():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tools', 'mgpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and 'MGPU_OK' in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
