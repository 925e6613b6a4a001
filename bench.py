#!/usr/bin/env python
"""bench.py -- reads mapped / second of the hashreadmapper hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the one its metric is quoted on -- "human-size 3N index"): 4 M synthetic
150 bp directional bisulfite reads (1 % substitution errors) per GPU per step against a 3.1 Gbp synthetic
reference with 24 GRCh38-like chromosome lengths, k=16, 16 hash tables, w=128, minTableHits=4,
maxHammingPercent=0.05, SW verification with CIGAR; both 3N indexes (C->T and G->A, 27.4 M windows each)
resident in HBM.  A step = one pass of the whole hot path (K1 pack, K2 minhash, K3 probe + retrieve, K4
collect, K5 SHD best window, K7 SW+CIGAR) over one 4 M-read batch per GPU (`--reads`; measured on B200: 8.6 M
reads/s with 1 M-read batches, 9.0 M with 2 M, 9.2 M with 4 M -- fixed per-batch costs).  `--genome-bp 46000000`
selects configs[1] (chr21-size reference).  N > 1: weak scaling, every rank maps its own 4 M-read shard against its
replica of the index; no data-path collective; `--index partitioned` = configs[4], the
key-partitioned index with NCCL all-to-all routing.

One JSON line on stdout (rank 0).  `value` = reads/s with the reads already in HBM; `e2e` = the same
through hrm_mapper_map_reads with pinned HOST buffers (H2D reads, D2H records + CIGARs inside the
timed region); `roofline` = the hash-probe kernel against the measured HBM peak; `cpu_baseline` /
`--impl reference` = the reference's own CPU functions (oracle/_ref, built from /root/reference)
timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_, W_, H_, T_ = 16, 128, 16, 4
READ_LEN = 150
ERR = 0.01


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    Q = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8 or (self.uuid and self.uuid not in f[0]):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = samples in the upper half of what was seen
        srt = sorted(sm)
        return {"sm_mhz": float(np.median(srt[len(srt) // 2:])), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


def workload(args, rank):
    from hashreadmapper_b200 import synth
    if args.genome_bp >= (1 << 31) or args.chromosomes > 1:  # BASELINE configs[2]: GRCh38-like chromosome lengths
        lengths = [int(x) for x in synth.human_like_lengths(args.genome_bp, max(args.chromosomes, 24))]
    else:
        lengths = [args.genome_bp]
    genome, off = synth.make_genome(lengths, seed=20240601)
    parts, at, piece = [], 0, 0
    while at < args.reads:  # pieces of 1 M reads bound the generator's temporaries
        cnt = min(1_000_000, args.reads - at)
        parts.append(synth.make_reads(genome, off, cnt, args.read_len, error_rate=args.error_rate,
                                      indel_frac=args.indel_frac, nondirectional=args.nondirectional,
                                      seed=20240602 + rank + 1000 * piece))
        at += cnt
        piece += 1
    reads = np.concatenate([p[0] for p in parts])
    lens = np.concatenate([p[1] for p in parts])
    truth = {k: np.concatenate([p[2][k] for p in parts]) for k in parts[0][2]}
    return genome, off, reads, lens, truth


def workload_name(reads, genome_bp, nchrom, shape="150bp directional BS reads"):
    rd = "%gM" % (reads / 1e6) if reads % 100_000 == 0 else str(reads)
    if genome_bp == 3_100_000_000:
        return ("%s x %s per GPU per step vs 3.1 Gbp human-size synthetic 3N index, 24 chromosomes "
                "(BASELINE configs[2]: batches of the 100M-read job)" % (rd, shape))
    if genome_bp == 46_000_000:
        tag = "BASELINE configs[1]" if shape == "150bp directional BS reads" else "read shape of BASELINE configs[3]"
        return "%s x %s per GPU per step vs 46 Mbp synthetic reference (%s)" % (rd, shape, tag)
    return "%s x %s per step vs %.4g Mbp synthetic reference, %d chromosome(s)" % (rd, shape, genome_bp / 1e6, nchrom)


def host_cores():
    """CPU cores this process may run on (torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample(args, n_reads, sample_reads=None, sample_bp=None):
    """bounded sample of the workload for the CPU arm: S reads drawn from a reference of sample_bp bases (one chromosome,
    same generator and seeds as the workload)"""
    from hashreadmapper_b200 import synth
    sub_bp = int(min(args.genome_bp, sample_bp or args.cpu_genome_bp))
    genome, off = synth.make_genome([sub_bp], seed=20240601)
    S = int(min(sample_reads or args.cpu_sample, n_reads))
    reads, lens, _ = synth.make_reads(genome, off, S, args.read_len, error_rate=args.error_rate,
                                      indel_frac=args.indel_frac, seed=20240602)
    return genome, off, reads, lens, S, args.genome_bp / float(sub_bp)


def reference_step(ref, port, genome_ct, genome_ga, off, reads, lens):
    """one step of the reference arm: the reference's CPU functions on a read sample, both 3N passes"""
    from oracle.pyoracle import ref_cpu_pipeline
    tot = np.zeros(4)
    mapped = np.zeros(len(lens), dtype=bool)
    for g in (genome_ct, genome_ga):
        out, _, _, times = ref_cpu_pipeline(ref, g, off, reads, lens, k=K_, w=W_, H=H_, min_hits=T_,
                                            want_alignments=False)
        tot += times
        mapped |= out["orientation"] != 3
    return tot, int(mapped.sum())


def read_shape(args):
    shape = "%dbp %s BS reads" % (args.read_len, "non-directional" if args.nondirectional else "directional")
    if args.error_rate != ERR or args.indel_frac:
        shape += " (%.3g %% errors, %.3g %% of them indels)" % (100 * args.error_rate, 100 * args.indel_frac)
    return shape


def base_config(args, world, partitioned=False):
    """the keys both arms share (the reference arm runs a bounded sample of exactly this workload)"""
    nchrom = 24 if (args.genome_bp >= (1 << 31) or args.chromosomes > 1) else 1
    return {"workload": workload_name(args.reads, args.genome_bp, nchrom, read_shape(args)), "reads_per_gpu": args.reads,
            "genome_bp": args.genome_bp, "k": K_, "hashmaps": H_, "window": W_, "min_table_hits": T_,
            "passes": "C->T index + G->A index" if not args.nondirectional else
                      "C->T and G->A reads x C->T and G->A index (4 passes)", "verification": "SW+CIGAR",
            "parallelism": ("reads sharded x%d, index replicated" % world) if not partitioned else
                           ("reads sharded x%d, index key-partitioned x%d, NCCL all-to-all" % (world, world))}


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone times the reference's CPU implementation of the path (all host cores).
    A step = the reference's whole CPU pipeline, both 3N passes, on a bounded sample (S reads against G_s bases) of the
    workload; the sample is sized from a calibration step so that K + W steps take about --ref-budget-s seconds, up to
    1 M reads against 500 Mbp.  value = reads / (window streaming scaled to the full genome + per-read
    stages scaled to the full batch): the reference streams every window once per batch whatever the batch holds."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle, have_ref
    cfg = base_config(args, world, args.index == "partitioned")
    if not have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhrm_ref.so was not built "
                          "(needs /root/reference at build time)"}))
        return
    ref, port = Oracle("ref"), Oracle("port")
    ref.lib.ref_set_num_threads(host_cores())
    cores = int(ref.lib.ref_num_threads())

    def prepare(S_want, bp_want):
        genome, off, reads, lens, S, wscale = cpu_sample(args, args.reads, S_want, bp_want)
        reads_ct = np.frombuffer(port.convert_ascii(reads.tobytes(), 1), dtype=np.uint8).reshape(reads.shape)
        return port.convert_ascii(genome, 1), port.convert_ascii(genome, 2), off, reads_ct, lens, S, wscale

    # calibration (untimed): cost per read and per streamed base on this box
    g_ct, g_ga, off, r_ct, lens, S0, _ = prepare(20_000, 23_000_000)
    reference_step(ref, port, g_ct, g_ga, off, r_ct[:2000], lens[:2000])
    t, _ = reference_step(ref, port, g_ct, g_ga, off, r_ct, lens)
    per_read = (t[0] + t[2] + t[3]) / S0
    per_bp = t[1] / float(off[-1])
    nsteps = args.steps + 0.25 * args.warmup
    step_budget = max(args.ref_budget_s / max(nsteps, 1.0), 1.0)
    # split the step evenly between the two terms, within [50 k, 1 M] reads and [46, 500] Mbp
    S = int(min(max(0.5 * step_budget / per_read, 50_000), 1_000_000, args.reads))
    bp = int(min(max(0.5 * step_budget / per_bp, 46_000_000), 500_000_000, args.genome_bp))
    if args.cpu_sample_fixed:
        S, bp = min(args.cpu_sample, args.reads), min(args.cpu_genome_bp, args.genome_bp)
    g_ct, g_ga, off, r_ct, lens, S, wscale = prepare(S, bp)
    wS = max(S // 4, 1000)
    for _ in range(args.warmup):
        reference_step(ref, port, g_ct, g_ga, off, r_ct[:wS], lens[:wS])
    t0 = time.perf_counter()
    acc = np.zeros(4)
    for _ in range(args.steps):
        t, nm = reference_step(ref, port, g_ct, g_ga, off, r_ct, lens)
        acc += t
    wall = time.perf_counter() - t0
    ms = wall / args.steps * 1e3
    per = acc / args.steps
    # fixed part: streaming + sketching all windows (times[1]); the rest scales with the number of reads
    scale = args.reads / S
    full_s = wscale * per[1] + scale * (per[0] + per[2] + per[3])
    value = args.reads / full_s
    line = {"metric": "reads mapped/sec", "value": value, "unit": "reads/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference",
                             "sample_reads": S, "sample_genome_bp": int(off[-1]),
                             "sample": ("each step = the reference's whole CPU pipeline (both 3N passes) on %d of %d reads "
                                        "against %.0f of %.0f Mbp, %d OpenMP threads; window streaming (%.2f s per step, "
                                        "measured) scaled x%.1f to the full genome and counted once per batch, per-read "
                                        "stages (%.2f s, measured) scaled x%.1f to the full batch; measured on the sample "
                                        "alone: %.0f reads/s"
                                        % (S, args.reads, off[-1] / 1e6, args.genome_bp / 1e6, cores, per[1], wscale,
                                           per[0] + per[2] + per[3], scale, S / (ms / 1e3))),
                             "stage_seconds_per_step": {"reads_build": per[0], "windows_query_filter": per[1],
                                                        "shd": per[2], "verify_ssw": per[3]}},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def parity_check(args, mp, cfg, genome, off, names, reads, lens, K):
    """After the timed region: K reads of the timed batch through the end-to-end SAM entry point, compared BYTE FOR BYTE
    with the oracle on the full-size index -- seeding by the reference's own CPU functions (oracle/_ref, OpenMP over the
    whole genome, both passes; the scalar port when the genome is small), stage V + SAM by the C restatement that is
    pinned to the reference's Mappinghandler.  Test infrastructure used as the checker, outside every timed region."""
    import hashreadmapper_b200._lib as L
    from oracle import pyoracle as po
    t0 = time.perf_counter()
    port = po.Oracle("port")
    ref = None
    if po.have_ref():
        ref = po.Oracle("ref")
        ref.lib.ref_set_num_threads(host_cores())
    elif off[-1] > 200_000_000:
        return {"parity_checked": 0, "parity_ok": None, "parity_note": "oracle/_ref not built and the genome is too large "
                "for the scalar port"}
    idx = np.linspace(0, len(lens) - 1, K).astype(np.int64)
    r, l = np.ascontiguousarray(reads[idx]), np.ascontiguousarray(lens[idx])
    genomes, rows, passes = [], [], []
    for p in range(cfg.num_passes):
        g = port.convert_ascii(genome, cfg.genome_conversion[p])
        rr = np.frombuffer(port.convert_ascii(r.tobytes(), cfg.read_conversion[p]), dtype=np.uint8).reshape(r.shape)
        genomes.append(g)
        rows.append(rr)
        if ref is not None:
            passes.append(po.ref_cpu_pipeline(ref, g, off, rr, l, k=K_, w=W_, H=H_, min_hits=T_, mapper_type=0,
                                              want_alignments=False)[0])
        else:
            passes.append(port.map_pass_refdir(g, off, rr, l, k=K_, w=W_, H=H_, min_hits=T_)[0])
    best = passes[0].copy()
    which = np.where(best["orientation"] != 3, 0, -1).astype(np.int32)
    for p, cur in enumerate(passes[1:], start=1):
        better = (cur["orientation"] != 3) & ((best["orientation"] == 3) | (cur["hammingDistance"] < best["hammingDistance"]))
        best[better] = cur[better]
        which[better] = p
    exp, _ = po.port_sam_format(port, genomes, off, names, rows, l, best, np.maximum(which, 0),
                                [cfg.verify_conversion[p] for p in range(cfg.num_passes)], w=W_)
    sq, rc, _, rec, cig = mp.mapReadsSam(r, l, cigar_pitch=256, want_records=True)
    got = L.SAM_HD + sq.tobytes() + L.SAM_PG_CO + rc.tobytes()
    m = best["orientation"] != 3
    same_mapped = bool((rec["mapped"]["orientation"] == best["orientation"]).all() and
                       (rec["mapped"]["position"][m] == best["position"][m]).all())
    nbad = 0
    if got != exp:
        a, b = got.split(b"\n"), exp.split(b"\n")
        nbad = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
    return {"parity_checked": int(K), "parity_ok": bool(got == exp), "parity_mapped_reads_identical": same_mapped,
            "parity_sam_bytes": len(exp), "parity_sam_lines_differing": int(nbad),
            "parity_mapped_in_sample": int(m.sum()),
            "parity_oracle": ("seeding: the reference's own CPU functions (oracle/_ref, %d threads) over the full %.0f Mbp, "
                              "every pass; stage V + SAM: the C restatement pinned to the reference's Mappinghandler"
                              % (host_cores(), off[-1] / 1e6)) if ref is not None else "scalar C restatement",
            "parity_seconds": time.perf_counter() - t0}


class StdoutToStderr:
    """routes file descriptor 1 to stderr while C libraries initialise (NCCL prints its version line on stdout): the
    bench prints exactly one line on stdout, the JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=4_000_000, help="reads per GPU per step (one batch)")
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--cpu-genome-bp", type=int, default=100_000_000,
                    help="reference bases the in-line CPU baseline streams (window streaming is scaled to --genome-bp)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="reads in the in-line CPU-baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=200.0,
                    help="--impl reference: seconds the K timed steps should take in total (sizes the sample)")
    ap.add_argument("--cpu-sample-fixed", action="store_true",
                    help="--impl reference: use --cpu-sample / --cpu-genome-bp instead of the calibrated sample")
    ap.add_argument("--check", type=int, default=2000,
                    help="reads of the timed batch compared with the oracle after the timed region (0: off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text", action="store_true", help="skip the SAM-text end-to-end figure")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ-text-in figure")
    ap.add_argument("--no-partitioned-segment", action="store_true",
                    help="N > 1: skip the short key-partitioned-index segment appended to the replicated run")
    ap.add_argument("--partitioned-sample", type=int, default=500_000, help="reads per GPU of that segment")
    ap.add_argument("--load-factor", type=float, default=None, help="hash-table load factor (default: the library's)")
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--error-rate", type=float, default=ERR)
    ap.add_argument("--indel-frac", type=float, default=0.0, help="fraction of the errors that are 1-bp indels")
    ap.add_argument("--nondirectional", action="store_true",
                    help="non-directional library: G->A reads too, four 3N passes (BASELINE configs[3] with --read-len 250 "
                         "--error-rate 0.03 --indel-frac 0.1)")
    ap.add_argument("--chromosomes", type=int, default=1, help="> 1: GRCh38-like chromosome lengths (configs[2])")
    ap.add_argument("--index", default="replicated", choices=["replicated", "partitioned"],
                    help="partitioned: key-partitioned tables, lookups routed with NCCL all-to-all (configs[4])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with StdoutToStderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    import hashreadmapper_b200 as hb
    import hashreadmapper_b200.api as api
    from hashreadmapper_b200 import parallel

    genome, off, reads, lens, truth = workload(args, rank)
    n = len(lens)
    cfg = api.nondirectional_config() if args.nondirectional else api.directional_config()
    if args.load_factor:
        cfg.load_factor = args.load_factor
    mp = api.Mapper(cfg)
    comm = None
    if args.index == "partitioned":
        with StdoutToStderr():
            comm = api.Comm()
        mp.setPartition(comm)
    t0 = time.perf_counter()
    names = ["chr%d" % (i + 1) for i in range(len(off) - 1)]
    mp.setGenome(genome, off, names)
    torch.cuda.synchronize()
    index_s = time.perf_counter() - t0
    info = mp.info()

    d_reads = torch.from_numpy(reads).cuda()
    d_lens = torch.from_numpy(lens).cuda()
    CIG = 64

    def step():
        mapped, _ = mp.mapBatch(d_reads, d_lens, want_stats=False)
        rec, cig, _ = mp.verifyBatch(d_reads, d_lens, mapped, cigar_pitch=CIG)
        return mapped, rec, cig

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # one accounted step (untimed) for the roofline bookkeeping
    torch.cuda.synchronize()
    i0 = mp.info()
    _, st = mp.mapBatch(d_reads, d_lens, want_stats=True)
    i1 = mp.info()
    st_collect = {"enumerated": i1.collect_ids_counted - i0.collect_ids_counted,
                  "skipped": i1.collect_ids_skipped - i0.collect_ids_skipped,
                  "block_reads": i1.collect_reads_block_kernel - i0.collect_reads_block_kernel}
    launches_map = st.num_kernel_launches
    mapped, _ = mp.mapBatch(d_reads, d_lens, want_stats=False)
    from hashreadmapper_b200 import _lib as L
    import ctypes as C
    st2 = L.BatchStats()
    rec = torch.empty((n, C.sizeof(L.ReadRecord) // 4), dtype=torch.int32, device=d_reads.device)
    cigs = torch.empty((2 * n, CIG), dtype=torch.uint8, device=d_reads.device)
    L.check(mp.lib.hrm_verify_batch(mp.h, C.c_void_p(d_reads.data_ptr()), reads.shape[1], C.c_void_p(d_lens.data_ptr()),
                                    n, C.c_void_p(mapped.data_ptr()), C.c_void_p(rec.data_ptr()),
                                    C.c_void_p(cigs.data_ptr()), CIG, C.byref(st2),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    launches_step = int(launches_map - 1 + st2.num_kernel_launches)  # minus the stats-only counting kernel
    del rec, cigs

    # ---- timed region: device-resident reads ---------------------------------------------------------
    uuid = ""
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid)
    # (1) one batch at a time on one stream: per-stage CUDA-event times and the launch durations of the roofline
    mp.setProfiling(True)
    mp.stageTimes()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    serial_ms = parallel.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    stages = mp.stageTimes()
    mp.setProfiling(False)

    # (2) the same K steps through the staged pipeline with device-resident reads (no host copies), for comparison
    def pipeline_run(steps):
        if comm is not None:
            for _ in range(steps):
                step()
            return
        mp.stageDevice(0, d_reads, d_lens, args.read_len)
        for i in range(steps):
            mp.mapStaged(i % 2, None, None, CIG)
            if i >= 1:
                mp.finish((i - 1) % 2)
            if i + 1 < steps:
                mp.stageDevice((i + 1) % 2, d_reads, d_lens, args.read_len)
        mp.finish((steps - 1) % 2)

    pipeline_run(3)
    barrier()
    p0 = torch.cuda.Event(enable_timing=True)
    p1 = torch.cuda.Event(enable_timing=True)
    p0.record()
    pipeline_run(args.steps)
    p1.record()
    barrier()
    pipe_ms = parallel.max_over_ranks(p0.elapsed_time(p1)) / args.steps

    # (3) the headline: K timed steps, one batch at a time, device-resident reads
    barrier()
    if rank == 0:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    total_ms = parallel.max_over_ranks(dev_ms)
    ms_per_step = total_ms / args.steps
    value = world * n * args.steps / (total_ms / 1e3)

    # ---- end to end: pinned host buffers through the double-buffered pipeline --------------------------------
    # every step: H2D of that step's reads (pinned), the whole hot path, D2H of its records + CIGARs; the copies of steps
    # i+1 and i-1 run under the kernels of step i (hrm_mapper_stage_reads / map_staged / finish)
    h_reads = torch.from_numpy(reads).pin_memory()
    h_lens = torch.from_numpy(lens).pin_memory()
    h_rec = [torch.empty((n * hb.RECORD_DTYPE.itemsize,), dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_cig = [torch.empty((2 * n, CIG), dtype=torch.uint8).pin_memory() for _ in range(2)]
    rec_np = [h.numpy().view(hb.RECORD_DTYPE) for h in h_rec]

    def e2e_run(steps, text=None):
        """steps batches through the pipeline; text = None: records + CIGARs out; else (sq buffers, record text buffers)"""
        if comm is not None:  # key-partitioned index: the routed queries are collective, batches run one at a time
            for i in range(steps):
                mp.mapReads(h_reads.numpy(), h_lens.numpy(), CIG, rec_np[i % 2], h_cig[i % 2].numpy())
            return (0, 0)
        trace = os.environ.get("BENCH_TRACE") is not None  # host time spent inside each call, per step, to stderr
        tt = []
        tp = time.perf_counter()
        mp.stageReads(0, h_reads.numpy(), h_lens.numpy())
        sizes = (0, 0)
        for i in range(steps):
            t_a = time.perf_counter()
            if text is None:
                mp.mapStaged(i % 2, rec_np[i % 2], h_cig[i % 2].numpy(), CIG)
            else:
                mp.mapStaged(i % 2, None, None, 128, i * n, text[0][i % 2], text[1][i % 2])
            t_b = time.perf_counter()
            if i >= 1:
                sizes = mp.finish((i - 1) % 2)
            t_c = time.perf_counter()
            if i + 1 < steps:
                mp.stageReads((i + 1) % 2, h_reads.numpy(), h_lens.numpy())
            tt.append((t_a - tp, t_b - t_a, t_c - t_b, time.perf_counter() - t_c))
        t_a = time.perf_counter()
        sizes = mp.finish((steps - 1) % 2)
        if trace and rank == 0:
            sys.stderr.write("trace %s: ms from start / in mapStaged / in finish / in stageReads per step: %s; last finish %.1f ms\n"
                             % ("text" if text is not None else "records",
                                " ".join("%.0f/%.0f/%.0f/%.0f" % tuple(1e3 * x for x in t) for t in tt),
                                1e3 * (time.perf_counter() - t_a)))
        return sizes

    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * args.steps / e2e_s
    last = rec_np[(args.steps - 1) % 2]
    n_mapped = int((last["mapped"]["orientation"] != 3).sum())
    ok = ((last["mapped"]["orientation"] != 3) & (last["mapped"]["chromosome_id"] == truth["chrom"]) &
          (last["mapped"]["position"] + last["mapped"]["shift"] == truth["pos"]))
    h2d = n * reads.shape[1] + n * 4
    d2h = n * hb.RECORD_DTYPE.itemsize + 2 * n * CIG
    # the one-shot call (serial copy -> compute -> copy), for comparison
    mp.mapReads(h_reads.numpy(), h_lens.numpy(), CIG, rec_np[0], h_cig[0].numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        mp.mapReads(h_reads.numpy(), h_lens.numpy(), CIG, rec_np[0], h_cig[0].numpy())
    torch.cuda.synchronize()
    oneshot_value = world * n * max(1, args.steps // 2) / parallel.max_over_ranks(time.perf_counter() - t0)
    # ---- end to end with TEXT out: reads in -> SAM text (V4 + O1 on the device) out, same pipeline ----------------
    e2e_text = None
    text_ok = 0.0
    if comm is None and not args.no_text:
        del h_rec, h_cig, rec_np
        line_bound = 96 + 128 + W_ + reads.shape[1]
        try:  # ~4.4 GB of pinned host memory per rank: every rank must get it, or all skip the figure together
            tx_rec = [torch.empty((n * line_bound,), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
            tx_sq = [torch.empty((n * 40,), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
            text_ok = 1.0
        except Exception:
            text_ok = 0.0
        text_ok = -parallel.max_over_ranks(-text_ok)
    if text_ok == 1.0:
        e2e_run(2, (tx_sq, tx_rec))
        barrier()
        t0 = time.perf_counter()
        sqw, recw = e2e_run(args.steps, (tx_sq, tx_rec))
        torch.cuda.synchronize()
        text_s = parallel.max_over_ranks(time.perf_counter() - t0)
        e2e_text = {"value": world * n * args.steps / text_s, "unit": "reads/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(sqw + recw),
                    "sam_bytes_per_read": float(sqw + recw) / n,
                    "api": "hrm_mapper_stage_reads / hrm_mapper_map_staged (V4 + SAM text on the device) / hrm_mapper_finish"}
        # ---- FASTQ text in -> SAM text out: the text goes H2D and is parsed by the device-side reader (hrm_ingest_reads)
        # while the other slot is verified ("FASTQ-in-memory -> SAM-records-in-memory")
        e2e_fastq = None
        if not args.no_fastq:
            L_ = args.read_len
            hdr = 11  # "@%09d\n"
            rec_len = hdr + L_ + 3 + L_ + 1
            fq = torch.empty((n, rec_len), dtype=torch.uint8).pin_memory()
            fqn = fq.numpy()
            ids = np.arange(n, dtype=np.int64)
            fqn[:, 0] = ord("@")
            for d_ in range(9):
                fqn[:, 9 - d_] = (ids // (10 ** d_)) % 10 + ord("0")
            fqn[:, 10] = 10
            fqn[:, hdr:hdr + L_] = reads[:, :L_]
            fqn[:, hdr + L_:hdr + L_ + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
            fqn[:, hdr + L_ + 3:hdr + 2 * L_ + 3] = ord("I")
            fqn[:, rec_len - 1] = 10
            fq_flat = fqn.reshape(-1)
            pitch_ = reads.shape[1]

            def fastq_run(steps):
                # same order as the row pipeline: queue batch i, fetch the text of batch i - 1 (its slot is free then),
                # stage batch i + 1 into that slot -- the H2D copy and the reader's kernels run under batch i's verification
                mp.stageFastq(0, fq_flat, pitch_, n, 0, 0)
                for i in range(steps):
                    mp.mapStaged(i % 2, None, None, 128, i * n, tx_sq[i % 2], tx_rec[i % 2])
                    if i >= 1:
                        mp.finish((i - 1) % 2)
                    if i + 1 < steps:
                        mp.stageFastq((i + 1) % 2, fq_flat, pitch_, n, (i + 1) * n, 0)
                return mp.finish((steps - 1) % 2)

            fastq_run(2)
            barrier()
            t0 = time.perf_counter()
            sqw2, recw2 = fastq_run(args.steps)
            torch.cuda.synchronize()
            fq_s = parallel.max_over_ranks(time.perf_counter() - t0)
            e2e_fastq = {"value": world * n * args.steps / fq_s, "unit": "reads/s", "h2d_bytes_per_step": int(fq_flat.size),
                         "d2h_bytes_per_step": int(sqw2 + recw2),
                         "api": "hrm_mapper_stage_fastq (H2D of the FASTQ text + device-side reader) / "
                                "hrm_mapper_map_staged (SAM text on the device) / hrm_mapper_finish"}
            del fq, fqn, fq_flat
        e2e_text["from_fastq_text"] = e2e_fastq
        del tx_rec, tx_sq
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: a short segment on the KEY-PARTITIONED index (BASELINE configs[4]) beside the replicated run ----------
    part_seg = None
    if world > 1 and comm is None and not args.no_partitioned_segment:
        with StdoutToStderr():
            comm2 = api.Comm()
        mp2 = api.Mapper(cfg)
        mp2.setPartition(comm2)
        t0 = time.perf_counter()
        mp2.setGenome(genome, off, names)
        torch.cuda.synchronize()
        pbuild = time.perf_counter() - t0
        ns = min(n, args.partitioned_sample)
        dr, dl = d_reads[:ns].contiguous(), d_lens[:ns].contiguous()
        m_rep, _ = mp.mapBatch(dr, dl, want_stats=False)
        m_par, _ = mp2.mapBatch(dr, dl, want_stats=False)
        same = 1.0 if torch.equal(m_rep, m_par) else 0.0
        same = -parallel.max_over_ranks(-same)  # min over the ranks
        ci0 = comm2.info()
        b0, x0 = ci0.bytes_sent, ci0.exchanges
        mp2.setProfiling(True)
        mp2.stageTimes()
        barrier()
        p0 = torch.cuda.Event(enable_timing=True)
        p1 = torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(2):
            mp2.mapBatch(dr, dl, want_stats=False)
        p1.record()
        barrier()
        pms = parallel.max_over_ranks(p0.elapsed_time(p1)) / 2
        pst = mp2.stageTimes()
        ci1 = comm2.info()
        part_seg = {"reads_per_gpu": int(ns), "seeding_reads_per_s": world * ns / (pms / 1e3), "ms_per_batch": pms,
                    "route_ms": pst["route"][0] / 2, "probe_ms": pst["probe"][0] / 2, "collect_ms": pst["filter"][0] / 2,
                    "shd_ms": pst["shd"][0] / 2, "bytes_sent_rank0_per_batch": int((ci1.bytes_sent - b0) // 2),
                    "alltoall_GBps_rank0": ((ci1.bytes_sent - b0) / 2) / 1e9 / (max(pst["route"][0] / 2, 1e-6) / 1e3),
                    "exchanges_per_batch": int((ci1.exchanges - x0) // 2), "identical_to_replicated": bool(same == 1.0),
                    "index_device_bytes_per_gpu": int(mp2.info().index_device_bytes), "index_build_s": pbuild,
                    "what": "K1..K5 (seeding, routed lookups over NCCL all-to-all, collection, best window) of the first "
                            "reads of every rank's batch on an index whose tables are split by key over the ranks; the "
                            "MappedReads are compared with the replicated index on every rank"}
        del mp2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    shape = "%dbp %s BS reads" % (args.read_len, "non-directional" if args.nondirectional else "directional")
    if args.error_rate != ERR or args.indel_frac:
        shape += " (%.3g %% errors, %.3g %% of them indels)" % (100 * args.error_rate, 100 * args.indel_frac)
    wl_name = workload_name(n, args.genome_bp, len(off) - 1, shape)
    # ---- roofline of the hash-probe kernel (K3b) -----------------------------------------------------
    peak, peak_src = peaks()
    probe_ms, probe_spans = stages["probe"]
    probe_launch_ms = probe_ms / max(probe_spans, 1)
    launches_per_step_probe = probe_spans / args.steps
    launches = max(launches_per_step_probe, 1.0)
    P = st.num_slot_touches / launches                # slots examined per probe launch
    Q = n * cfg.num_passes / launches                 # queries per probe launch
    alg_bytes = 16.0 * P + (8.0 * H_ + 4.0) * Q       # 16 B per slot touch + signatures in + count out
    achieved = alg_bytes / (probe_launch_ms / 1e3) / 1e9 if probe_launch_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "probe_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    stage_ms = {k: v[0] / args.steps for k, v in stages.items()}
    stage_sum = sum(stage_ms.values())
    dram_frac = (traffic / (probe_launch_ms / 1e3) / 1e9 / peak) if (traffic and probe_launch_ms > 0) else None
    min_bytes = 16.0 * Q * H_ + (8.0 * H_ + 4.0) * Q   # one 16-B slot per lookup: the least a probe must read
    probe_roofline = {"kernel": "hrm::probe_tm_kernel (K3b hash probe, table-major)" if comm is None else
                                "hrm::probe_keys_kernel (K3b hash probe of routed keys, owner side)", "bound": "hbm",
                      "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                      "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": probe_launch_ms,
                      "launches_per_step": launches_per_step_probe, "slot_touches_per_launch": P,
                      "lookups_per_launch": Q * H_,
                      "share_of_step": probe_ms / args.steps / serial_ms if serial_ms > 0 else None,
                      "frac_dram_side": dram_frac,
                      "frac_one_slot_per_lookup": (min_bytes / (probe_launch_ms / 1e3) / 1e9 / peak) if probe_launch_ms > 0 else None,
                      "note": "three accountings of the same launch: `frac` counts all four 16-B slots of every 64-B bucket "
                              "visited (16 B x slots examined + (8H+4) B x reads, SURVEY 8d); `frac_dram_side` = ncu DRAM "
                              "bytes (profiles/probe_traffic.json) / time, below `frac` because one 78 MB table at a time "
                              "stays in the 126 MB L2; `frac_one_slot_per_lookup` counts one 16-B slot per lookup"}
    # the kernel with the largest share of the step: the fused collection (K3b value retrieval + K4)
    dom_stage = max(stage_ms, key=lambda k_: stage_ms[k_])
    ids_step = float(st_collect["enumerated"])          # ids inserted into the duplicate filter, per step
    ids_skipped_step = float(st_collect["skipped"])     # ids of the two largest buckets: read and tested only
    coll_launches = max(stages["filter"][1] / args.steps, 1.0)
    coll_ms = stage_ms["filter"]
    coll_bytes = 4.0 * (ids_step + ids_skipped_step) + 8.0 * H_ * n * cfg.num_passes + 8.0 * n * cfg.num_passes
    coll_ach = coll_bytes / (coll_ms / 1e3) / 1e9 if coll_ms > 0 else 0.0
    ctraffic = None
    tp2 = os.path.join(ROOT, "profiles", "collect_traffic.json")
    if os.path.exists(tp2):
        try:
            ctraffic = json.load(open(tp2)).get("dram_bytes_per_launch")
        except Exception:
            ctraffic = None
    collect_roofline = {"kernel": "hrm::collect_dup_kernel<false> (warp per read; + <true>, block per read, for the ~2.5 % of "
                                  "skewed reads): K3b value retrieval + K4 candidate collection, fused", "bound": "hbm",
                        "achieved": coll_ach,
                        "peak": peak, "unit": "GB/s", "frac": coll_ach / peak, "traffic": ctraffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": coll_bytes / coll_launches, "launch_ms": coll_ms / coll_launches,
                        "launches_per_step": coll_launches, "share_of_step": coll_ms / serial_ms if serial_ms > 0 else None,
                        "ids_inserted_per_step": ids_step, "ids_tested_per_step": ids_skipped_step,
                        "frac_dram_side": (ctraffic / (coll_ms / coll_launches / 1e3) / 1e9 / peak)
                        if (ctraffic and coll_ms > 0) else None,
                        "reads_to_block_kernel_per_step": float(st_collect["block_reads"]),
                        "note": "algorithmic bytes = 4 B x every id of every bucket a read touches (each is read exactly once: "
                                "inserted into the duplicate filter or tested against it) + the 8-B (offset, count) range of "
                                "every (read, table) + the 8-B list header per read (DESIGN 3); launch_ms spans the kernels of "
                                "one pass' collection (CUDA events on the launching stream); traffic = ncu DRAM bytes of those "
                                "kernels (profiles/collect_traffic.json)"}
    # `roofline` = the HBM-bound kernel with the largest share of the step (the fused collection on the replicated index);
    # verification (K7) is integer-ALU bound: its figures are reported beside it, not against the HBM peak
    roofline = collect_roofline if comm is None else probe_roofline
    other = probe_roofline if roofline is collect_roofline else collect_roofline
    qp, rp = ((args.read_len + 15) // 16) * 16, W_
    cells = 2.0 * 2.0 * qp * rp * n   # 2 alignments x (forward + reverse pass) x padded read x window, upper bound
    verify_info = {"kernel": "hrm::sw_pair_passes_kernel (K7a: both SW passes, s16x2 DPX wavefront) + band ladder (K7b: trace "
                             "back + CIGAR)", "bound": "integer ALU (ncu: ALU pipe 80 % in K7a, profiles/README.md)",
                   "stage_ms": stage_ms["verify"], "share_of_step": stage_ms["verify"] / serial_ms if serial_ms > 0 else None,
                   "cell_updates_per_step_upper_bound": cells,
                   "gcups_over_the_whole_stage": cells / (stage_ms["verify"] / 1e3) / 1e9 if stage_ms["verify"] > 0 else None}
    dominant_stage = dom_stage

    line = {"metric": "reads mapped/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64 integer", "data": "synthetic",
            "config": dict(base_config(args, world, comm is not None),
                           load_factor=float(cfg.load_factor),
                           l2="inputs larger than L2 (reads %.0f MB + index %.0f MB per GPU)"
                              % (reads.nbytes / 1e6, info.index_device_bytes / 1e6),
                           index_build_s=index_s, windows=int(info.num_windows),
                           index_device_bytes=int(info.index_device_bytes), table_slots=int(info.table_slots_total),
                           table_keys=int(info.num_keys_total)),
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": ("hrm_mapper_stage_reads / hrm_mapper_map_staged / hrm_mapper_finish (pinned host buffers, two "
                            "batches in flight)") if comm is None else "hrm_mapper_map_reads (pinned host buffers)",
                    "one_shot_value": oneshot_value, "one_shot_api": "hrm_mapper_map_reads (serial copy, compute, copy)"},
            "e2e_text": e2e_text, "partitioned_segment": part_seg,
            "gpu_launches": int(launches_step * args.steps),
            "roofline": roofline,
            "roofline_probe" if roofline is collect_roofline else "roofline_collect": other,
            "verify_alu": verify_info, "dominant_stage": dominant_stage,
            "stages_ms_per_step": stage_ms, "stages_unaccounted_ms": serial_ms - stage_sum,
            "ms_per_step_with_stage_timers": serial_ms,
            "value_staged_pipeline_device_resident": world * n / (pipe_ms / 1e3),
            "mapped_fraction": n_mapped / n, "mapped_at_true_locus_fraction": float(ok.sum()) / max(n_mapped, 1),
            "candidates_per_read": st.num_candidates / n, "values_per_read": st.num_values / n,
            "collect_ids_skipped_fraction": (float(mp.info().collect_ids_skipped) /
                                             max(1, mp.info().collect_ids_skipped + mp.info().collect_ids_counted)),
            "clocks": clocks}
    if comm is not None:
        ci = comm.info()
        line["exchange"] = {"backend": "NCCL ncclSend/ncclRecv groups", "bytes_sent_rank0": int(ci.bytes_sent),
                            "exchanges_rank0": int(ci.exchanges)}

    # ---- parity at the quoted size: K reads of the timed batch against the oracle (outside every timed region) -----
    if world == 1 and args.check > 0:
        try:
            line.update(parity_check(args, mp, cfg, genome, off, names, reads, lens, min(args.check, n)))
        except Exception as e:
            line.update({"parity_checked": 0, "parity_ok": None, "parity_note": "failed: %r" % (e,)})

    # ---- CPU baseline: the reference's own functions on this box's host cores (rank 0, N = 1) -------
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.pyoracle import Oracle, have_ref
            port = Oracle("port")
            g_s, off_s, r_s, l_s, S, wscale = cpu_sample(args, n)
            r_ct = np.frombuffer(port.convert_ascii(r_s.tobytes(), 1), dtype=np.uint8).reshape(S, -1)
            g_ct, g_ga = port.convert_ascii(g_s, 1), port.convert_ascii(g_s, 2)
            if have_ref():
                ref = Oracle("ref")
                ref.lib.ref_set_num_threads(host_cores())
                t0 = time.perf_counter()
                per, nm = reference_step(ref, port, g_ct, g_ga, off_s, r_ct, l_s)
                wall = time.perf_counter() - t0
                scale = n / S
                full_s = wscale * per[1] + scale * (per[0] + per[2] + per[3])
                line["cpu_baseline"] = {
                    "value": n / full_s, "unit": "reads/s", "cores": int(ref.lib.ref_num_threads()), "kind": "reference",
                    "sample": ("%d of %d reads against %.0f Mbp of reference, both 3N passes, %.1f s wall; window "
                               "streaming (%.2f s) scaled x%.1f to %.0f Mbp, per-read stages (%.2f s) scaled x%.1f; "
                               "measured on the sample alone: %.0f reads/s"
                               % (S, n, min(args.genome_bp, args.cpu_genome_bp) / 1e6, wall, per[1], wscale,
                                  args.genome_bp / 1e6, per[0] + per[2] + per[3], scale, S / wall)),
                    "stage_seconds": {"reads_build": per[0], "windows_query_filter": per[1], "shd": per[2],
                                      "verify_ssw": per[3]}}
            else:
                S2 = min(S, 5000)
                t0 = time.perf_counter()
                for g in (g_ct, g_ga):
                    port.map_pass_refdir(g, off_s, r_ct[:S2], l_s[:S2])
                wall = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": S2 / wall, "unit": "reads/s", "cores": 1, "kind": "port",
                                        "sample": "%d reads, seeding+SHD only (scalar oracle port, no SSW)" % S2}
        except Exception as e:  # the baseline must never break the measurement
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference",
                                    "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
