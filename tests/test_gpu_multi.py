"""Multi-GPU parity under pytest: runs tests/mgpu_check.py (every merge strategy of query_b200.dist against the oracle:
fused peer mailbox, peer arena with owner-sharded finalisation, NCCL all-gather / all-reduce, owner-bucketed all-to-all of
records and DISTINCT entries) under torch.distributed.run on min(device_count, 8) ranks.  Skipped on a box with one GPU -
the in-process peer-arena test of tests/test_gpu_parity.py and the world_size-2 gloo tests cover the logic there."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_every_merge_strategy_on_all_gpus_of_the_box():
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs (this box has %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "mgpu_check ok" in p.stdout, p.stdout[-4000:]
