"""BASELINE configs[4]: key-partitioned 3N index over N GPUs, k / window-size sweep.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sweep_kw.py [--genome-bp 3100000000] [--reads 500000] [--out profiles/r2_sweep_kw_nN.json]

For every (k, w) with w >= readLength / 2 (ref: include/referencewindows.hpp:25-26) the index is rebuilt with its tables
split by key over the ranks; every rank maps its own reads (seeding + routed lookups + collection + best window, K1..K5);
reported per setting: reads/s over all ranks, ms per batch, route (all-to-all) ms, bytes sent by rank 0, all-to-all GB/s,
index bytes per GPU, mapped fraction, identical-to-replicated on a 20 k-read sample (replicated index built per setting)."""
import argparse
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
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--reads", type=int, default=500_000)
    ap.add_argument("--ks", default="12,16,20,24,32")
    ap.add_argument("--ws", default="64,128,256")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--check-replicated", type=int, default=20_000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hashreadmapper_b200.api as api
    from hashreadmapper_b200 import synth, parallel
    lengths = ([int(x) for x in synth.human_like_lengths(args.genome_bp, 24)] if args.genome_bp >= (1 << 31)
               else [args.genome_bp])
    genome, off = synth.make_genome(lengths, seed=20240601)
    reads, lens, truth = synth.make_reads(genome, off, args.reads, args.read_len, error_rate=0.01, seed=20240700 + rank)
    d_reads, d_lens = torch.from_numpy(reads).cuda(), torch.from_numpy(lens).cuda()
    comm = api.Comm()
    rows = []
    for k in [int(x) for x in args.ks.split(",")]:
        for w in [int(x) for x in args.ws.split(",")]:
            row = {"k": k, "w": w}
            if w < args.read_len // 2 or w < k:
                row["skipped"] = "w < readLength / 2 (ref: referencewindows.hpp:25-26)" if w >= k else "w < k"
                rows.append(row)
                continue
            cfg = api.directional_config(k=k, window_size=w)
            # reads per batch: small k saturates the 3-letter k-mer space (buckets at the 65535 cap): bound the exchange
            nb = args.reads if k >= 16 else max(args.reads // 25, 2000)
            try:
                t0 = time.perf_counter()
                mp = api.Mapper(cfg)
                mp.setPartition(comm)
                mp.setGenome(genome, off)
                torch.cuda.synchronize()
                row["index_build_s"] = time.perf_counter() - t0
                info = mp.info()
                row["index_device_bytes_per_gpu"] = int(info.index_device_bytes)
                row["windows"] = int(info.num_windows)
                dr, dl = d_reads[:nb].contiguous(), d_lens[:nb].contiguous()
                out, st = mp.mapBatch(dr, dl, want_stats=True)  # warm-up + counters
                ci0 = comm.info()
                mp.setProfiling(True)
                mp.stageTimes()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out, _ = mp.mapBatch(dr, dl, want_stats=False)
                e1.record()
                torch.cuda.synchronize()
                ms = parallel.max_over_ranks(e0.elapsed_time(e1))
                stg = mp.stageTimes()
                ci1 = comm.info()
                o = out.cpu().numpy()
                mapped = float((o[:, 0] != 3).mean())
                row.update({"reads_per_gpu_per_batch": int(nb), "reads_per_s": world * nb / (ms / 1e3), "ms_per_batch": ms,
                            "route_ms": stg["route"][0], "probe_ms": stg["probe"][0], "collect_ms": stg["filter"][0],
                            "shd_ms": stg["shd"][0], "bytes_sent_rank0": int(ci1.bytes_sent - ci0.bytes_sent),
                            "alltoall_GBps_rank0": (ci1.bytes_sent - ci0.bytes_sent) / 1e9 / (max(stg["route"][0], 1e-6) / 1e3),
                            "exchanges": int(ci1.exchanges - ci0.exchanges), "values_per_read": st.num_values / max(nb, 1),
                            "candidates_per_read": st.num_candidates / max(nb, 1), "mapped_fraction": mapped})
                if args.check_replicated > 0:
                    nc = min(args.check_replicated, nb)
                    rep = api.Mapper(cfg)
                    rep.setGenome(genome, off)
                    a, _ = rep.mapBatch(d_reads[:nc].contiguous(), d_lens[:nc].contiguous(), want_stats=False)
                    b, _ = mp.mapBatch(d_reads[:nc].contiguous(), d_lens[:nc].contiguous(), want_stats=False)
                    same = 1.0 if torch.equal(a, b) else 0.0
                    row["identical_to_replicated"] = bool(-parallel.max_over_ranks(-same) == 1.0)
                    row["replicated_index_device_bytes"] = int(rep.info().index_device_bytes)
                    del rep
                del mp
            except Exception as e:  # a setting that does not fit must not end the sweep (all ranks fail alike)
                row["error"] = repr(e)[:300]
            torch.cuda.empty_cache()
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0:
        res = {"what": "BASELINE configs[4]: key-partitioned 3N index, k / w sweep", "n_gpus": world,
               "genome_bp": args.genome_bp, "read_len": args.read_len, "hashmaps": 16, "min_table_hits": 4, "rows": rows}
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
