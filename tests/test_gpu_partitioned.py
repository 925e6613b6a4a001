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


STAGING_THREAD_SCRIPT = r"""
import sys, threading
import numpy as np, torch
sys.path.insert(0, %r)
import hashreadmapper_b200.api as api
from hashreadmapper_b200 import synth
torch.cuda.set_device(1)
genome, off = synth.make_genome([90000], seed=5)
reads, lens, _ = synth.make_reads(genome, off, 2000, 150, error_rate=0.02, seed=6)
reads = torch.from_numpy(reads).pin_memory().numpy()
mp = api.Mapper(api.directional_config())
mp.setGenome(genome, off, ["chrA"])
sq, rc, _, rec, cig = mp.mapReadsSam(reads, lens, first_read_id=0, cigar_pitch=128, want_records=True)
err = []
def stage():  # a fresh host thread: CUDA's current device is 0 here, the mapper lives on device 1
    try:
        mp.stageReads(0, reads, lens)
    except Exception as e:
        err.append(e)
t = threading.Thread(target=stage); t.start(); t.join()
assert not err, err
o_sq, o_tx = np.zeros(2000 * 40 + 16, np.uint8), np.zeros(2000 * 600, np.uint8)
mp.mapStaged(0, None, None, 128, 0, o_sq, o_tx)
sqw, recw = mp.finish(0)
assert o_sq[:sqw].tobytes() == sq.tobytes() and o_tx[:recw].tobytes() == rc.tobytes()
print("staging thread OK")
"""


def test_staging_thread_on_second_device(cuda):
    """a mapper created on device 1 is staged from a fresh host thread (whose current device is 0): every
    hrm_mapper_* entry point binds the calling thread to the mapper's device"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-c", STAGING_THREAD_SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "staging thread OK" in r.stdout
