"""Multi-GPU plumbing: one process per GPU, reads sharded contiguously, index replicated.

Every stage of the path is per read, so there is NO data-path collective: rank r maps reads
[lo_r, hi_r) against its own replica of the 3N index.  The only communication is the gather of the
fixed-size result records (and of timings) on rank 0, done with torch.distributed (NCCL on GPUs,
gloo in the CPU tests).  Concatenating the shards in rank order reproduces the single-GPU output
order, because the reference's per-read arg-min (main_gpu.cu:777-821) does not depend on other reads.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """read i -> rank floor(i * world / n): contiguous blocks, sizes differ by at most one"""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def gather_records(local: np.ndarray, n_total: int, group=None):
    """Gathers per-rank structured record arrays on rank 0 in rank order.  Returns the full array on
    rank 0 and None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    itemsize = local.dtype.itemsize
    counts = [shard_range(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in counts)
    buf = torch.zeros((maxn * itemsize,), dtype=torch.uint8, device=dev)
    raw = torch.from_numpy(np.frombuffer(local.tobytes(), dtype=np.uint8).copy())
    buf[:raw.numel()] = raw.to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    if backend == "nccl":
        # NCCL has no gather for uneven sizes: all_gather of the padded buffers
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
    else:
        dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    parts = []
    for r, (lo, hi) in enumerate(counts):
        b = outs[r][:(hi - lo) * itemsize].cpu().numpy().tobytes()
        parts.append(np.frombuffer(b, dtype=local.dtype))
    return np.concatenate(parts)


def max_over_ranks(value: float, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
