"""Utterance sharding with NCCL: over 2 GPUs when the box has them, and — on any GPU box — the same torchrun entry with ONE
rank (single-rank NCCL communicator: packed scatter, per-rank synthesis through the engine, waveform gather), so the NCCL wire path is
exercised on a one-GPU box too; the world-size-2 gloo tests cover the dealing and packing logic on the CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_synthesis_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "run_sharded.py")], capture_output=True, text=True, timeout=600)
    assert "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_synthesis_nccl_single_rank():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "scripts", "run_sharded.py")], capture_output=True, text=True, timeout=300)
    assert "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
