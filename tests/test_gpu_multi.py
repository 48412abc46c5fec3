"""Multi-GPU parity as a driver-visible test: runs scripts/check_multigpu.py under torch.distributed.run on two GPUs
when the box has them (skipped on a one-GPU box).  That script asserts: row-block kNN and database-ring kNN bit-exact
against the single-GPU search; the edge-sharded optimiser with the fused peer-memory epoch tail (mmu_epoch_tail_peer)
equal to the single-GPU optimiser from the same state within the run-to-run noise of fp32 atomics, identical kept-edge
counts, replicas bit-identical across ranks.  (the host-side partitioning logic is covered on CPU by
tests/test_dist_cpu.py with gloo.)"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("tail", ["push", "fused", "legacy"])
def test_two_gpu_parity(tail):
    env = dict(os.environ, MMUMAP_PEER_TAIL=tail)
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(ROOT, "scripts", "check_multigpu.py")], env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    assert "multi-GPU parity OK" in run.stdout
