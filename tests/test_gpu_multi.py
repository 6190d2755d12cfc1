"""Multi-GPU checks (need >= 2 GPUs on the box; skipped otherwise): the NVLink peer-memory joint exchange of
csrc/xchg.cu against NCCL's all_reduce, one process per GPU under torch.distributed.run."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_peer_memory_exchange_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "xchg_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "XCHG PASS" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_product_losses_under_data_parallel_match_full_batch_oracle():
    """The PRODUCT path sharded over the GPUs of the box (tools/dp_parity_check.py): loss == full-batch fp64 oracle and
    bit-identical on all ranks, shard gradient == slice of the full-batch oracle gradient, both transports, one call at
    a time and as one batched exchange, eagerly and from a CUDA graph."""
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "dp_parity_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP PARITY PASS" in r.stdout, r.stdout[-3000:]
