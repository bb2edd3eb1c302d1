"""BASELINE config 4 on more than one GPU: the batch-sharded training step (odcp_b200.train_step: convolutions on
cuDNN, fused head, DDP gradient all-reduce over NCCL, fused SGD) must leave the same parameters as a single-process
step over the whole batch.  Needs two GPUs on the box (`gpurun --gpus 2`, log kept under profiles/); skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_fused_step_with_peer_reduction_equals_the_whole_batch(cuda_device):
    """Two ranks: the loss terms summed inside the kernel over NVLink peer memory equal the whole batch's, the
    shards' gradients and kept boxes are the whole batch's slices (tests/multi_peer_check.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "multi_peer_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-3000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["world"] == 2


def test_cfg4_sharded_training_step_equals_the_whole_batch(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "cfg4_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["world"] == 2


def test_cfg4_step_single_process_check(cuda_device):
    """The same script with one rank: the 'sharded' and the whole-batch step are the same computation."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "cfg4_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    assert json.loads(lines[-1])["ok"]
