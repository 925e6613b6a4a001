"""GPU parity of the fused path (hrm_map_batch / hrm_verify_batch / hrm_mapper_map_reads) against the
oracle's restatement of the REFERENCE'S pipeline in the reference's own direction (tables over reads,
windows streamed in batches of 2048): identical MappedRead per read, identical raw SSW fields and
CIGARs, identical edit distances -- on pre-converted input per pass, plus size-independent
properties at larger sizes.
"""
import numpy as np
import pytest

from hashreadmapper_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_pass(port, genome, off, reads, lens, rconv, gconv, **kw):
    g = port.convert_ascii(genome, gconv)
    r = np.frombuffer(port.convert_ascii(reads.tobytes(), rconv), dtype=np.uint8).reshape(reads.shape)
    return port.map_pass_refdir(g, off, r, lens, **kw)


def merge(passes):
    best = passes[0].copy()
    which = np.where(best["orientation"] != 3, 0, -1).astype(np.int32)
    for p, cur in enumerate(passes[1:], start=1):
        better = (cur["orientation"] != 3) & ((best["orientation"] == 3) | (cur["hammingDistance"] < best["hammingDistance"]))
        best[better] = cur[better]
        which[better] = p
    return best, which


def check_mapped(got, exp, which):
    m = exp["orientation"] != 3
    assert (got["orientation"] == exp["orientation"]).all()
    for a, b in (("hamming_distance", "hammingDistance"), ("shift", "shift"), ("chromosome_id", "chromosomeId"),
                 ("position", "position")):
        assert (got[a][m] == exp[b][m]).all(), a
    assert (got["pass"] == which).all()


@pytest.mark.parametrize("conf", ["single_none", "single_ct", "directional", "nondirectional", "config4"])
def test_map_and_verify_small(cuda, port, conf):
    import torch
    genome, off = synth.make_genome([60000, 25013], seed=11)
    nondir = conf in ("nondirectional", "config4")
    reads, lens, truth = synth.make_reads(genome, off, 3000, 150, error_rate=0.02 if conf != "single_none" else 0.0,
                                          nondirectional=nondir, seed=12)
    if conf == "config4":  # BASELINE configs[3]: non-directional library, 250 bp reads, 3 % errors of which 10 % indels
        reads, lens, truth = synth.make_reads(genome, off, 2000, 250, error_rate=0.03, indel_frac=0.1,
                                              nondirectional=True, seed=14)
    if conf == "single_none":  # plain 4-letter mapping of unconverted reads (the reference as shipped)
        reads, lens, truth = synth.make_reads(genome, off, 3000, 150, error_rate=0.01, conversion_rate=0.0, seed=12)
    # a few junk / short reads
    reads[5, :150] = np.frombuffer(b"ACGT" * 37 + b"AC", dtype=np.uint8)
    lens[7] = 12
    lens[9] = 40
    cfg = {"single_none": cuda.default_config(), "single_ct": cuda.default_config(),
           "directional": cuda.directional_config(), "nondirectional": cuda.nondirectional_config(),
           "config4": cuda.nondirectional_config()}[conf]
    if conf == "single_ct":
        cfg.read_conversion[0] = cfg.genome_conversion[0] = cfg.verify_conversion[0] = 1
    mp = cuda.Mapper(cfg)
    mp.setGenome(genome, off, ["chrA", "chrB"])
    info = mp.info()
    assert info.num_windows == sum((int(l) + 112) // 113 for l in np.diff(off))
    # oracle, pass by pass
    passes = []
    for p in range(cfg.num_passes):
        res, stats = oracle_pass(port, genome, off, reads, lens, cfg.read_conversion[p], cfg.genome_conversion[p])
        passes.append(res)
    exp, which = merge(passes)
    nmapped_exp = int((exp["orientation"] != 3).sum())
    # one strand only maps on the C->T index alone; 3 % errors over 250 bp exceed the 5 % Hamming threshold more often
    assert nmapped_exp > {"single_ct": 1200, "config4": 600}.get(conf, 2400)
    d_reads = torch.from_numpy(reads).cuda()
    d_lens = torch.from_numpy(lens).cuda()
    out, st = mp.mapBatch(d_reads, d_lens)
    import hashreadmapper_b200 as hb
    got = out.cpu().numpy().view(hb.MAPPED_DTYPE).reshape(-1)
    check_mapped(got, exp, which)
    assert st.num_mapped == (exp["orientation"] != 3).sum()
    assert st.num_kernel_launches > 0 and st.num_probes == len(lens) * 16 * cfg.num_passes
    # verification records through the host-buffer entry point
    rec, cig, st2 = mp.mapReads(reads, lens, cigar_pitch=96)
    check_mapped(rec["mapped"], exp, which)
    nchk = 0
    for i in range(len(lens)):
        if exp["orientation"][i] == 3:
            assert (rec["alignments"][i]["sw_score"] == 0).all()
            continue
        p = which[i]
        gconv, rconv, vconv = cfg.genome_conversion[p], cfg.read_conversion[p], cfg.verify_conversion[p]
        c = int(exp["chromosomeId"][i])
        chrom = port.convert_ascii(genome[off[c]:off[c + 1]], gconv)
        rd = port.convert_ascii(reads[i, :lens[i]].tobytes(), rconv)
        q, qrc, ref = port.verify_inputs(rd, int(exp["orientation"][i]), chrom, int(exp["position"][i]), 128, vconv)
        ml = max(15, int(lens[i]) // 2)
        assert rec["window_length"][i] == len(ref) and rec["mask_len"][i] == ml
        for a, qq in enumerate((q, qrc)):
            ea, ecig = port.ssw_align(qq, ref, ml)
            ga = tuple(int(rec["alignments"][i][a][n]) for n in hb.ALIGN_DTYPE.names[:9])
            clen = int(rec["alignments"][i][a]["cigar_len"])
            gc = bytes(cig[2 * i + a, :min(clen, cig.shape[1])]).decode()  # long cigars are cut at cigar_pitch
            assert ga == ea and clen == len(ecig) and gc == ecig[:cig.shape[1]], (i, a, ea, ecig, ga, gc)
        nchk += 1
    assert nchk == nmapped_exp
    sam = mp.samFormat(rec, cig, reads, lens)
    assert sam.startswith(b"@HD\tVN:1.4\n@SQ\tSN:0\tLN:")
    assert sam.count(b"\n") == 1 + len(lens) + 1 + len(lens)


def test_edlib_mode(cuda, port):
    genome, off = synth.make_genome([40000], seed=3)
    reads, lens, _ = synth.make_reads(genome, off, 800, 150, error_rate=0.01, seed=4)
    cfg = cuda.directional_config(mapper_type=1)
    mp = cuda.Mapper(cfg)
    mp.setGenome(genome, off)
    rec, cig, st = mp.mapReads(reads, lens)
    n = 0
    for i in range(len(lens)):
        m = rec["mapped"][i]
        if m["orientation"] == 3:
            continue
        p = int(m["pass"])
        chrom = port.convert_ascii(genome, cfg.genome_conversion[p])
        rd = port.convert_ascii(reads[i, :lens[i]].tobytes(), cfg.read_conversion[p])
        q, qrc, ref = port.verify_inputs(rd, int(m["orientation"]), chrom, int(m["position"]), 128,
                                         cfg.verify_conversion[p])
        assert int(rec["edit_distance"][i][0]) == port.edit_distance_nw(q, ref)
        assert int(rec["edit_distance"][i][1]) == port.edit_distance_nw(qrc, ref)
        n += 1
    assert n > 500


def test_properties_at_scale(cuda, port):
    """size-independent properties on a larger instance: truth recovery, idempotence, batch independence"""
    import torch
    import hashreadmapper_b200 as hb
    genome, off = synth.make_genome([3_000_000, 1_500_000], seed=21)
    reads, lens, truth = synth.make_reads(genome, off, 200_000, 150, error_rate=0.01, seed=22)
    mp = cuda.Mapper(cuda.directional_config())
    mp.setGenome(genome, off)
    d_reads = torch.from_numpy(reads).cuda()
    d_lens = torch.from_numpy(lens).cuda()
    out, st = mp.mapBatch(d_reads, d_lens)
    a = out.cpu().numpy().view(hb.MAPPED_DTYPE).reshape(-1).copy()
    mapped = a["orientation"] != 3
    assert mapped.mean() > 0.9
    # mapped reads land on their true locus: window start + shift == true start (substitution-only reads)
    ok = mapped & (a["chromosome_id"] == truth["chrom"]) & (a["position"] + a["shift"] == truth["pos"])
    assert ok.sum() / mapped.sum() > 0.99
    # + strand reads come from the C->T index in forward orientation, - strand reads from the G->A index as RC
    assert ((a["pass"][ok] == 0) == truth["strand"][ok]).all()
    assert ((a["orientation"][ok] == 1) == truth["strand"][ok]).all()
    # idempotence and independence of the batch composition
    out2, _ = mp.mapBatch(d_reads, d_lens)
    assert (out2.cpu().numpy() == out.cpu().numpy()).all()
    half = len(lens) // 2
    o1, _ = mp.mapBatch(d_reads[:half].contiguous(), d_lens[:half].contiguous())
    o2, _ = mp.mapBatch(d_reads[half:].contiguous(), d_lens[half:].contiguous())
    cat = np.concatenate([o1.cpu().numpy(), o2.cpu().numpy()])
    assert (cat == out.cpu().numpy()).all()
    # the general path (retrieve + K4 over value lists; HRM_COLLECT=0) against the default fused collection, and
    # candidate values beyond the value budget: the batch is halved into ranges of reads (human-genome scale
    # retrieves thousands of values per read); results and counters must not depend on path or split
    import os
    os.environ["HRM_COLLECT"] = "0"
    try:
        mp_gen = cuda.Mapper(cuda.directional_config())
        os.environ["HRM_VALUE_BUDGET"] = str(int(st.num_values // 2 // 7))
        mp_split = cuda.Mapper(cuda.directional_config())
    finally:
        os.environ.pop("HRM_VALUE_BUDGET", None)
        del os.environ["HRM_COLLECT"]
    mp_gen.setGenome(genome, off)
    mp_split.setGenome(genome, off)
    outg, stg = mp_gen.mapBatch(d_reads, d_lens)
    out3, st3 = mp_split.mapBatch(d_reads, d_lens)
    assert (outg.cpu().numpy() == out.cpu().numpy()).all()
    assert (out3.cpu().numpy() == out.cpu().numpy()).all()
    for s_ in (stg, st3):
        assert (s_.num_values, s_.num_candidates, s_.num_mapped) == (st.num_values, st.num_candidates, st.num_mapped)
    assert st3.num_kernel_launches > stg.num_kernel_launches  # it did split
    del mp_split, mp_gen
    # a sample against the oracle's reference-direction pipeline restricted to that sample
    idx = np.arange(0, len(lens), 400)
    passes = []
    for p in range(2):
        g = port.convert_ascii(genome, 1 + p)
        r = np.frombuffer(port.convert_ascii(reads[idx].tobytes(), 1), dtype=np.uint8).reshape(len(idx), -1)
        passes.append(port.map_pass_refdir(g, off, r, lens[idx])[0])
    exp, which = merge(passes)
    check_mapped(a[idx], exp, which)


@pytest.mark.parametrize("min_hits", [4, 2, 3])
def test_fused_collection_paths(cuda, tmp_path, min_hits):
    """K3b retrieval + K4 fused (k4_fused.cu): warp path, block path, many id ranges per read, tiny tables, skipped
    largest buckets -- all against the general retrieve + filter path, on a low-complexity genome with planted
    repeats (large buckets, many survivors)"""
    import os
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "collect_worker.py")
    variants = {"general": {"HRM_COLLECT": "0"},
                "default": {},  # duplicate detection: Bloom filter + exact event table, warp per read
                # small filters / small event tables: reads overflow to the block-wide variant and on to the table kernel
                "dup_tiny_filter": {"HRM_COLLECT_BLOOM_WORDS": "64"},
                "dup_tiny_tables": {"HRM_COLLECT_BLOOM_WORDS": "64", "HRM_COLLECT_XSLOTS": "256",
                                    "HRM_COLLECT_BLOOM_BLOCK_WORDS": "64", "HRM_COLLECT_BLOCK_XSLOTS": "2048"},
                "dup_small_events": {"HRM_COLLECT_XSLOTS": "256"},
                "dup_block_small": {"HRM_COLLECT_BLOOM_WORDS": "64", "HRM_COLLECT_BLOOM_BLOCK_WORDS": "512"},
                "dup_big": {"HRM_COLLECT_BLOOM_WORDS": "8192", "HRM_COLLECT_XSLOTS": "4096"},
                "dup_all_block": {"HRM_COLLECT_WARP_CAP": "8"},
                # ... through to the counting-table block kernel with many id ranges / tiny tables
                "dup_to_table": {"HRM_COLLECT_WARP_CAP": "8", "HRM_COLLECT_BLOOM_BLOCK_WORDS": "64", "HRM_COLLECT_SLOTS": "256",
                                 "HRM_COLLECT_FILL": "40"},
                "dup_to_table_tiny": {"HRM_COLLECT_WARP_CAP": "4", "HRM_COLLECT_BLOOM_BLOCK_WORDS": "64",
                                      "HRM_COLLECT_SLOTS": "64", "HRM_COLLECT_FILL": "12"},
                # the counting-table warp kernel with id ranges (HRM_COLLECT_RANGES=1)
                "warp_table": {"HRM_COLLECT_RANGES": "1"},
                "warp_ranges": {"HRM_COLLECT_RANGES": "1", "HRM_COLLECT_WARP_SLOTS": "64"},
                "warp_ranges_unpacked": {"HRM_COLLECT_RANGES": "1", "HRM_COLLECT_WARP_SLOTS": "128",
                                         "HRM_COLLECT_UNPACKED": "1"},
                "block": {"HRM_COLLECT_RANGES": "1", "HRM_COLLECT_WARP_CAP": "8"},
                "probe_qm": {"HRM_PROBE_TM": "0"},  # query-major probe ([n][H] signatures and ranges) + fused collection
                "probe_qm_general": {"HRM_PROBE_TM": "0", "HRM_COLLECT": "0"},
                "unpacked": {"HRM_COLLECT_UNPACKED": "1"},
                "unpacked_ranges": {"HRM_COLLECT_UNPACKED": "1", "HRM_COLLECT_WARP_CAP": "8", "HRM_COLLECT_SLOTS": "128",
                                    "HRM_COLLECT_FILL": "30"}}
    res, logs = {}, {}
    for name, env in variants.items():
        path = str(tmp_path / (name + ".npy"))
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, worker, path, str(min_hits)], capture_output=True, text=True, env=e,
                           timeout=600)
        assert r.returncode == 0, name + ": " + r.stdout + r.stderr
        res[name] = np.load(path)
        logs[name] = r.stdout.strip().splitlines()[-1]
    assert (res["general"][:, 0] != 3).mean() > 0.5
    for name in variants:
        assert (res[name] == res["general"]).all(), name
        assert logs[name] == logs["general"], (name, logs[name], logs["general"])


@pytest.mark.parametrize("conf", ["configs0_full", "configs1_sample"])
def test_baseline_configs_against_the_reference_cpu_path(cuda, port, ref, conf):
    """BASELINE configs at their stated index sizes against the REFERENCE'S OWN CPU functions (oracle/_ref:
    cpuhashtable.hpp tables over the reads, windows streamed in batches of 2048, hammingdistance templates, ssw.c):
    configs[0] in full (10 k reads x 5 Mbp); configs[1]'s 46 Mbp index with a 20 k-read sample of its 1 M reads
    (the table-major probe, the duplicate-detection collection on real bucket skew).  MappedRead + both raw SSW
    alignments per read, both 3N passes merged."""
    import torch
    from oracle.pyoracle import ref_cpu_pipeline
    if conf == "configs0_full":
        genome, off = synth.make_genome([5_000_000], seed=20240601)
        reads, lens, _ = synth.make_reads(genome, off, 10_000, 150, error_rate=0.0, seed=20240602)
    else:
        genome, off = synth.make_genome([46_000_000], seed=20240601)
        reads, lens, _ = synth.make_reads(genome, off, 20_000, 150, error_rate=0.01, seed=20240602)
    cfg = cuda.directional_config()
    mp = cuda.Mapper(cfg)
    mp.setGenome(genome, off, ["chrS"])
    rec, cig, st = mp.mapReads(reads, lens, cigar_pitch=128)
    passes, sws = [], []
    for p in range(2):
        g = port.convert_ascii(genome, cfg.genome_conversion[p])
        r = np.frombuffer(port.convert_ascii(reads.tobytes(), cfg.read_conversion[p]), dtype=np.uint8).reshape(reads.shape)
        # stage V of pass 1 verifies with G->A: the reference's C->T stage on the complemented text
        if cfg.verify_conversion[p] == 2:
            from oracle.pyoracle import complement_ascii
            out, _, _, _ = ref_cpu_pipeline(ref, g, off, r, lens, want_alignments=False)
            gc_ = complement_ascii(g)
            rc_ = np.frombuffer(complement_ascii(r.tobytes()), dtype=np.uint8).reshape(r.shape)
            sw = verify_with_reference(ref, gc_, off, rc_, lens, out)
        else:
            out, sw, _, _ = ref_cpu_pipeline(ref, g, off, r, lens, want_alignments=True)
        passes.append(out)
        sws.append(sw)
    exp, which = merge(passes)
    check_mapped(rec["mapped"], exp, which)
    m = exp["orientation"] != 3
    assert m.mean() > 0.9
    names = [n for n in rec["alignments"].dtype.names if n != "cigar_len"]
    for p in range(2):
        sel = m & (which == p)
        for a in range(2):
            for n in names:
                assert (rec["alignments"][n][sel, a] == sws[p][n][sel, a]).all(), (p, a, n)


def verify_with_reference(ref, genome, off, reads, lens, mapped):
    """2 x Aligner::Align per mapped read by the reference's own ssw (ref_ssw_align_batch), C->T inputs built as
    mappinghandler.cu:397-553 builds them"""
    import ctypes as C
    from oracle.pyoracle import ALIGN_DTYPE, _p
    n = len(lens)
    sw = np.zeros((n, 2), dtype=ALIGN_DTYPE)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in np.nonzero(mapped["orientation"] != 3)[0]:
        L = int(lens[i])
        rd = bytes(reads[i, :L])
        if mapped["orientation"][i] == 2:
            rd = rd.translate(comp)[::-1]
        rc = rd.translate(comp)[::-1]
        pos = int(mapped["position"][i])
        c = int(mapped["chromosomeId"][i])
        chrom = genome[off[c]:off[c + 1]]
        win = chrom[pos:pos + 128]
        q0, q1, w3 = rd.replace(b"C", b"T"), rc.replace(b"C", b"T"), win.replace(b"C", b"T")
        for a, q in enumerate((q0, q1)):
            al, _ = ref.ssw_align(q, w3, max(15, L // 2))
            for k, nme in enumerate(ALIGN_DTYPE.names[:9]):
                sw[nme][i, a] = al[k]
    return sw
