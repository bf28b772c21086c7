"""Two-GPU parity (skipped on a single-GPU box): tests/mgpu_check.py under torchrun -- the row-sharded
reconstruction, gathered by NCCL and by the C++ row-shard group (six distinct scans pipelined over two slots, no
barrier between them), is byte-identical to the single-GPU reconstruction of the same frames.  On a one-GPU box the
single-rank form of the same schedule runs instead."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_row_shard_two_gpus_identical_to_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL gather" in r.stdout and "every cloud identical to 1 GPU: True" in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_row_shard_group_single_rank_schedule():
    """world = 1 through the same entry points and the same overlapped schedule (runs on a one-GPU box too)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "every cloud identical to 1 GPU: True" in r.stdout, r.stdout[-2000:]
