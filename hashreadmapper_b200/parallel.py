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


# ---- key-partitioned index: host model of the routed query (csrc/partition.cu) -------------------------
def _alltoallv(send: np.ndarray, scount, rcount, group=None):
    """variable all-to-all of a 1-D array with torch.distributed (gloo on CPU, NCCL on GPU tensors)"""
    out = torch.empty(int(sum(rcount)), dtype=torch.from_numpy(send[:0].copy()).dtype)
    dist.all_to_all_single(out, torch.from_numpy(np.ascontiguousarray(send)), [int(c) for c in rcount],
                           [int(c) for c in scount], group=group)
    return out.numpy()


def routed_query_model(sigs: np.ndarray, owner_of, lookup, group=None):
    """The exchange protocol of partition.cu restated with numpy + torch.distributed, for the CPU tests
    (world_size 2, gloo) and as executable documentation of the message layout.

    sigs [n, H] uint64: this rank's signatures.  owner_of(keys) -> int array: rank owning each key.
    lookup(keys, tables) -> list of int arrays: this rank's SHARD answering routed lookups.
    Returns (num_per_seq [n], offsets [n + 1], values): values of read i = buckets of tables 0..H-1
    concatenated, i.e. what the replicated index's retrieve writes."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n, H = sigs.shape
    flat = sigs.reshape(-1)
    dest = np.where(flat == np.uint64(0xFFFFFFFFFFFFFFFF), rank, owner_of(flat)).astype(np.int64)
    perm = np.argsort(dest, kind="stable")                       # 1. stable partition by destination
    scount = np.bincount(dest, minlength=world).astype(np.int64)
    mat = torch.zeros((world, world), dtype=torch.int64)          # 2. G x G count matrix
    dist.all_gather_into_tensor(mat.view(-1), torch.from_numpy(scount), group=group)
    rcount = mat[:, rank].numpy().copy()
    rkeys = _alltoallv(flat[perm].view(np.int64), scount, rcount, group).view(np.uint64)
    rtabs = _alltoallv((perm % H).astype(np.uint8), scount, rcount, group)
    lists = lookup(rkeys, rtabs)                                   # 3. owner: probe the shard
    rcnt = np.array([len(v) for v in lists], dtype=np.int32)
    cnt_back = _alltoallv(rcnt, rcount, scount, group)
    roff = np.concatenate([[0], np.cumsum(rcount)])
    vsend = np.array([rcnt[roff[p]:roff[p + 1]].sum() for p in range(world)], dtype=np.int64)
    vmat = torch.zeros((world, world), dtype=torch.int64)          # 4. value totals, then the values
    dist.all_gather_into_tensor(vmat.view(-1), torch.from_numpy(vsend), group=group)
    vrecv = vmat[:, rank].numpy().copy()
    svals = np.concatenate([np.asarray(v, dtype=np.uint32) for v in lists] + [np.zeros(0, np.uint32)])
    rvals = _alltoallv(svals.view(np.int32), vsend, vrecv, group).view(np.uint32)
    src_off = np.concatenate([[0], np.cumsum(cnt_back)])[:-1]     # 5. back to (read, table) order
    cnt_e = np.zeros(n * H, dtype=np.int64)
    src_e = np.zeros(n * H, dtype=np.int64)
    cnt_e[perm] = cnt_back
    src_e[perm] = src_off
    num = cnt_e.reshape(n, H).sum(axis=1)
    offsets = np.concatenate([[0], np.cumsum(num)])
    values = np.concatenate([rvals[src_e[e]:src_e[e] + cnt_e[e]] for e in range(n * H)] + [np.zeros(0, np.uint32)])
    return num.astype(np.int32), offsets.astype(np.int32), values
