"""Data-parallel exchange over NVLink peer memory (grapes_allreduce_adam_peer).  Needs >= 2 GPUs on the box; the
1-GPU parity run skips it (the host-side sharding logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_allreduce_adam_matches_nccl_and_is_identical_across_ranks(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={w}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "check_peer_exchange.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "peer exchange ok" in res.stdout


def test_engine_data_parallel_step_equals_mean_of_oracle_gradients(cuda_device):
    """scripts/check_dp_engine.py under torchrun: the engine's own exchange (peer-memory kernel and NCCL) against the
    mean of the per-rank float64 oracle gradients, bit-identical parameters across ranks (eager and graph replay), and
    the flagged / no-op / sticky behaviour when a peer never arrives."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "scripts", "check_dp_engine.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-6000:] + res.stderr[-2000:]
    assert "dp engine ok" in res.stdout
