"""Data parallelism on hardware: N = 2 ranks over NCCL (skipped with fewer than 2 GPUs; run it with
`gpurun --gpus 2`).  tools/dp_check.py asserts that the all-reduced gradient equals the average of
the per-shard oracle gradients, that all ranks hold bit-identical parameters after 3 optimizer
steps, and that losses / parameters follow the oracle's averaged-gradient trajectory."""
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
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_rank_nccl_step_matches_averaged_gradient_oracle():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tools', 'dp_check.py')]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res['ok'] and res['params_bit_identical_across_ranks'], res
