"""Key-partitioned 3N index (BASELINE config 5): routed lookups over NCCL must reproduce the replicated index
bit for bit.  World size 1 runs in-process on one GPU (self send/recv through NCCL); with >= 2 GPUs visible the
same worker runs under torchrun, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "partition_worker.py")


def test_partitioned_world_1(cuda):
    r = subprocess.run([sys.executable, WORKER, "8000", "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0/1 OK" in r.stdout


def test_partitioned_multi_gpu(cuda):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n = min(n, 8)
    port = str(23000 + os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", port, WORKER, "40000", "600000"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(n):
        assert "rank %d/%d OK" % (k, n) in r.stdout
