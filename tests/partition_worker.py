"""Worker of the key-partitioned-index tests (TEST-ONLY).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/partition_worker.py [n_reads] [genome_bp]

Every rank maps its shard of the reads twice -- against its replica of the 3N index and through the
key-partitioned index (NCCL all-to-all) -- and requires bit-identical mapped reads, records and CIGARs.
Rank world-1 additionally runs one EMPTY batch while the others run a full one (collective with n = 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    genome_bp = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hashreadmapper_b200.api as api
    from hashreadmapper_b200 import synth, parallel
    genome, off = synth.make_genome([genome_bp * 2 // 3, genome_bp - genome_bp * 2 // 3], seed=21)
    reads, lens, _ = synth.make_reads(genome, off, n_reads, 150, error_rate=0.02, seed=22)
    lens[3] = 10  # shorter than k: invalid signatures
    lo, hi = parallel.shard_range(n_reads, rank, world)
    d_reads = torch.from_numpy(reads[lo:hi]).cuda()
    d_lens = torch.from_numpy(lens[lo:hi]).cuda()

    rep = api.Mapper(api.directional_config())
    rep.setGenome(genome, off)
    comm = api.Comm()
    os.environ["HRM_PART_CHUNK"] = "3000"  # several collective chunks per batch, uneven over the ranks
    par = api.Mapper(api.directional_config())
    del os.environ["HRM_PART_CHUNK"]
    par.setPartition(comm)
    par.setGenome(genome, off)
    ri, pi = rep.info(), par.info()
    assert pi.num_windows == ri.num_windows
    if world > 1:
        assert pi.table_slots_total < 0.75 * ri.table_slots_total, (pi.table_slots_total, ri.table_slots_total)

    m0, s0 = rep.mapBatch(d_reads, d_lens)
    m1, s1 = par.mapBatch(d_reads, d_lens)
    assert torch.equal(m0, m1), "mapped reads differ between the replicated and the partitioned index"
    assert s0.num_values == s1.num_values and s0.num_candidates == s1.num_candidates and s0.num_mapped == s1.num_mapped
    r0, c0, _ = rep.verifyBatch(d_reads, d_lens, m0)
    r1, c1, _ = par.verifyBatch(d_reads, d_lens, m1)
    assert torch.equal(r0, r1)
    # uneven round: the last rank passes an empty batch
    k = 0 if (rank == world - 1 and world > 1) else min(1000, hi - lo)
    m2, _ = par.mapBatch(d_reads[:k].contiguous(), d_lens[:k].contiguous())
    assert torch.equal(m2, m0[:k])
    info = comm.info()
    assert info.exchanges > 0
    if world > 1:
        assert info.bytes_sent > 0 and info.bytes_received > 0
    mapped = int((m0[:, 0] != 3).sum())
    print("rank %d/%d OK reads %d mapped %d slots %d -> %d sent %d B" % (rank, world, hi - lo, mapped,
                                                                       ri.table_slots_total, pi.table_slots_total,
                                                                       info.bytes_sent), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
