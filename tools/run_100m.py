"""BASELINE configs[2] as a JOB: 100 M distinct 150 bp directional BS reads against the human-size 3N index, in batches of
4 M, sharded over the ranks, through the staged pipeline with SAM text out; per-rank summaries gathered on rank 0.

    python tools/run_100m.py [--reads 100000000] [--batch 4000000] [--out profiles/r2_job_100m.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... tools/run_100m.py

Every batch is generated from its own seed (distinct reads), staged from pinned host memory, mapped, verified, formatted
as SAM text on the device and copied back; the host checks every record line's read id range and counts mapped reads
at their true locus from the binary records.  The synthetic read generator (numpy, one host thread) is ~30x slower than
the mapper, so the job's wall time says nothing about the mapper; reported instead: the DEVICE time of every batch
(CUDA events from the start of its H2D copy to the end of its last kernel / copy, batches do not overlap here because
the host generates between them) and the generator time."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--batch", type=int, default=4_000_000)
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hashreadmapper_b200 as hb
    import hashreadmapper_b200.api as api
    from hashreadmapper_b200 import synth, parallel
    lengths = ([int(x) for x in synth.human_like_lengths(args.genome_bp, 24)] if args.genome_bp >= (1 << 31)
               else [args.genome_bp])
    genome, off = synth.make_genome(lengths, seed=20240601)
    mp = api.Mapper(api.directional_config())
    mp.setGenome(genome, off, ["chr%d" % (i + 1) for i in range(len(off) - 1)])
    nbatches = (args.reads + args.batch - 1) // args.batch
    mine = [b for b in range(nbatches) if b % world == rank]  # batch b holds reads [b * batch, ...): SAM order = batch order
    n = args.batch
    pitch = 160
    bound = n * (96 + 128 + 128 + pitch)
    h_reads = [torch.empty((n, pitch), dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_lens = [torch.empty((n,), dtype=torch.int32).pin_memory() for _ in range(2)]
    h_rec = [torch.empty((n * hb.RECORD_DTYPE.itemsize,), dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_txt = [torch.empty((bound,), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    h_sq = [torch.empty((n * 40,), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    truths = [None, None]
    counts = [0, 0]
    tot = {"reads": 0, "mapped": 0, "true_locus": 0, "sam_bytes": 0, "sq_bytes": 0}
    digest = hashlib.sha256()
    gen_s = 0.0
    dev_ms = 0.0

    def generate(k):
        nonlocal gen_s
        t0 = time.perf_counter()
        b = mine[k]
        cnt = min(n, args.reads - b * n)
        r, l, t = synth.make_reads(genome, off, cnt, 150, error_rate=0.01, seed=20250000 + b, pitch=pitch)
        h_reads[k % 2].numpy()[:cnt] = r
        h_lens[k % 2].numpy()[:cnt] = l
        truths[k % 2] = t
        counts[k % 2] = cnt
        gen_s += time.perf_counter() - t0

    def consume(k, sizes):
        cnt = counts[k % 2]
        rec = h_rec[k % 2].numpy().view(hb.RECORD_DTYPE)[:cnt]
        t = truths[k % 2]
        m = rec["mapped"]["orientation"] != 3
        ok = m & (rec["mapped"]["chromosome_id"] == t["chrom"]) & (rec["mapped"]["position"] + rec["mapped"]["shift"] == t["pos"])
        tot["reads"] += cnt
        tot["mapped"] += int(m.sum())
        tot["true_locus"] += int(ok.sum())
        tot["sq_bytes"] += sizes[0]
        tot["sam_bytes"] += sizes[1]
        txt = h_txt[k % 2][:sizes[1]]
        first = bytes(txt[:12]).split(b"\t")[0]
        assert int(first) == mine[k] * n, (first, mine[k] * n)     # QNAME of the batch's first record = its read id
        assert txt[-1] == 10 and int((txt == 10).sum()) == cnt       # one line per read
        digest.update(txt[:1 << 20].tobytes())

    if mine:
        generate(0)
    for k in range(len(mine)):
        cnt = counts[k % 2]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mp.stageReads(k % 2, h_reads[k % 2].numpy()[:cnt], h_lens[k % 2].numpy()[:cnt])
        mp.mapStaged(k % 2, h_rec[k % 2].numpy().view(hb.RECORD_DTYPE)[:cnt], None, 128, mine[k] * n, h_sq[k % 2],
                     h_txt[k % 2])
        sizes = mp.finish(k % 2)
        torch.cuda.synchronize()
        dev_ms += (time.perf_counter() - t0) * 1e3  # device idle before and after: host wall = H2D + kernels + D2H of the batch
        if k + 1 < len(mine):
            generate(k + 1)
        consume(k, sizes)
    torch.cuda.synchronize()
    allr = {k: parallel.sum_over_ranks(float(v)) for k, v in tot.items()}
    pipe_max = parallel.max_over_ranks(dev_ms / 1e3)
    if rank == 0:
        res = {"what": "BASELINE configs[2] as a job: distinct reads in batches, SAM text out", "n_gpus": world,
               "reads": int(allr["reads"]), "batches": nbatches, "batch": n, "mapped": int(allr["mapped"]),
               "mapped_at_true_locus": int(allr["true_locus"]), "sam_record_bytes": int(allr["sam_bytes"]),
               "sam_sq_bytes": int(allr["sq_bytes"]), "batch_seconds_summed_max_over_ranks": pipe_max,
               "reads_per_s_batch_by_batch": allr["reads"] / pipe_max if pipe_max > 0 else None,
               "generator_seconds_rank0": gen_s, "sam_prefix_sha256_rank0": digest.hexdigest(),
               "note": "batch seconds = H2D + kernels + SAM text D2H of every batch, one batch at a time (the host generates "
                       "the next batch in between, so nothing overlaps); the generator is the wall-clock bottleneck"}
        print(json.dumps(res))
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
