"""The shipped multi-GPU API (`coivo_b200.dist.sharded_loss` over `coivo_b200.photometric_loss`) on real devices:
world size 2, NCCL with one GPU per rank when the box has two, and -- so that a one-GPU box covers the same code --
two ranks sharing cuda:0 with gloo carrying the scalar all-reduce.  The check itself is tests/tools/dist_cuda_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(backend, reduce):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "dist_cuda_worker.py"), "--backend", backend,
           "--reduce", reduce]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "DIST_OK" in p.stdout, p.stdout[-2000:]
    from conftest import record_parity
    record_parity("dist", line=[l for l in p.stdout.splitlines() if "DIST_OK" in l][0])


@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_sharded_loss_two_ranks_one_gpu_gloo(reduce):
    _run("gloo", reduce)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_loss_two_ranks_nccl():
    _run("nccl", "sum")
