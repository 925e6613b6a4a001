"""GPU parity of V4 + O1 (score recalculation, conversion count, alignment choice, MAPQ, POS, SAM text) against the
oracle's restatement (oracle/hrm_oracle.c: orc_sam_format), which is pinned byte for byte to the reference's own
Mappinghandler (tests/test_oracle_pin.py, tests/golden/golden_sam_v1.json made by oracle/ref_shim_sam.cpp).
Everything goes through the C ABI: hrm_verify_batch -> hrm_sam_fields_batch / hrm_sam_format_device / hrm_sam_format,
and end to end through hrm_mapper_map_reads_sam."""
import hashlib
import json
import os

import numpy as np
import pytest

import samcase
from hashreadmapper_b200 import synth

pytestmark = pytest.mark.gpu
GOLD_SAM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sam_v1.json")


def first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return "line %d:\n got %r\n exp %r" % (i, x[:400], y[:400])
    return "lengths %d vs %d lines" % (len(la), len(lb))


def to_device_mapped(cuda, torch, mapped, passes):
    import hashreadmapper_b200._lib as L
    m = np.zeros(len(mapped), dtype=L.MAPPED_DTYPE)
    m["orientation"] = mapped["orientation"]
    m["hamming_distance"] = mapped["hammingDistance"]
    m["shift"] = mapped["shift"]
    m["chromosome_id"] = mapped["chromosomeId"]
    m["position"] = mapped["position"]
    m["pass"] = np.where(mapped["orientation"] != 3, passes, -1)
    return torch.from_numpy(m.view(np.int32).reshape(len(mapped), 8)).cuda()


@pytest.mark.parametrize("name", ["ct150", "ga150", "ct250", "none100"])
def test_sam_cases(cuda, port, name):
    """perturbed MappedReads (every branch of printtoSAM) -> fields and text identical to the oracle and the golden"""
    import torch
    import hashreadmapper_b200._lib as L
    from oracle import pyoracle as po
    case = samcase.make(port, name)
    n = len(case["lens"])
    cfg = cuda.default_config()
    cfg.read_conversion[0], cfg.genome_conversion[0], cfg.verify_conversion[0] = case["rconv"], case["gconv"], case["conv"]
    mp = cuda.Mapper(cfg)
    mp.setGenome(case["raw_genome"], case["off"], case["names"])
    d_reads = torch.from_numpy(case["raw_reads"]).cuda()
    d_lens = torch.from_numpy(case["lens"]).cuda()
    d_mapped = to_device_mapped(cuda, torch, case["mapped"], np.zeros(n, np.int32))
    rec, cig, _ = mp.verifyBatch(d_reads, d_lens, d_mapped, cigar_pitch=256)
    exp_sam, F = samcase.port_sam(po, port, case)
    got = mp.samFields(d_reads, d_lens, rec, cig)
    for f in ("sw_score", "sw_score_next_best", "num_conversions", "chosen", "flag", "mapq", "window_length", "pos"):
        assert (got[f] == F[f]).all(), (f, np.nonzero((got[f] != F[f]).reshape(n, -1).any(1))[0][:5])
    sq = mp.samFormatDevice(d_reads, d_lens, rec, cig, L.SAM_SQ_LINES).cpu().numpy().tobytes()
    rc = mp.samFormatDevice(d_reads, d_lens, rec, cig, L.SAM_RECORDS).cpu().numpy().tobytes()
    sam = L.SAM_HD + sq + L.SAM_PG_CO + rc
    assert sam == exp_sam, first_diff(sam, exp_sam)
    with open(GOLD_SAM) as f:
        g = json.load(f)["cases"][name]
    assert len(sam) == g["bytes"] and hashlib.sha256(sam).hexdigest() == g["sha256"]  # the reference's own bytes
    # host-buffer entry point: same text, with and without the header block, and a non-zero first read id
    h_rec = rec.cpu().numpy().view(L.RECORD_DTYPE).reshape(n)
    h_cig = cig.cpu().numpy()
    assert mp.samFormat(h_rec, h_cig, case["raw_reads"], case["lens"]) == exp_sam
    assert mp.samFormat(h_rec, h_cig, case["raw_reads"], case["lens"], with_header=False) == rc
    exp7, _ = po.port_sam_format(port, [case["genome"]], case["off"], case["names"], [case["reads"]], case["lens"],
                                 case["mapped"], np.zeros(n, np.int32), [case["conv"]], first_read_id=4000000000 - 5)
    assert mp.samFormat(h_rec, h_cig, case["raw_reads"], case["lens"], first_read_id=4000000000 - 5) == exp7


def oracle_run(port, po, cfg, genome, off, names, reads, lens, **kw):
    """the reference pipeline per pass on pre-converted input, merged as the mapper merges -> SAM text + fields"""
    genomes, rows, passes = [], [], []
    for p in range(cfg.num_passes):
        g = port.convert_ascii(genome, cfg.genome_conversion[p])
        r = np.frombuffer(port.convert_ascii(reads.tobytes(), cfg.read_conversion[p]), dtype=np.uint8).reshape(reads.shape)
        genomes.append(g)
        rows.append(r)
        passes.append(port.map_pass_refdir(g, off, r, lens, **kw)[0])
    best = passes[0].copy()
    which = np.where(best["orientation"] != 3, 0, -1).astype(np.int32)
    for p, cur in enumerate(passes[1:], start=1):
        better = (cur["orientation"] != 3) & ((best["orientation"] == 3) | (cur["hammingDistance"] < best["hammingDistance"]))
        best[better] = cur[better]
        which[better] = p
    vc = [cfg.verify_conversion[p] for p in range(cfg.num_passes)]
    sam, F = po.port_sam_format(port, genomes, off, names, rows, lens, best, np.maximum(which, 0), vc,
                                w=cfg.window_size)
    return sam, F, best, which


@pytest.mark.parametrize("conf", ["configs0", "configs3_shape"])
def test_sam_end_to_end(cuda, port, conf):
    """whole runs, FASTQ-less: host reads -> hrm_mapper_map_reads_sam -> the complete SAM text, byte for byte against
    the oracle.  configs0 = BASELINE configs[0] in full (10 k C->T-converted 150 bp reads vs 5 Mbp, directional
    passes); configs3_shape = non-directional library, 250 bp, 3 % errors of which 10 % indels (four passes)."""
    import hashreadmapper_b200._lib as L
    from oracle import pyoracle as po
    if conf == "configs0":
        genome, off = synth.make_genome([5_000_000], seed=20240601)
        reads, lens, _ = synth.make_reads(genome, off, 10_000, 150, error_rate=0.0, seed=20240602)
        cfg = cuda.directional_config()
        names = ["chrS"]
    else:
        genome, off = synth.make_genome([150_000, 70_001], seed=41)
        reads, lens, _ = synth.make_reads(genome, off, 4000, 250, error_rate=0.03, indel_frac=0.1, nondirectional=True,
                                          seed=42)
        cfg = cuda.nondirectional_config()
        names = ["chr1", "chr2"]
    lens[17] = 30
    mp = cuda.Mapper(cfg)
    mp.setGenome(genome, off, names)
    sq, rc, st, rec, cig = mp.mapReadsSam(reads, lens, cigar_pitch=256, want_records=True)
    sam = L.SAM_HD + sq.tobytes() + L.SAM_PG_CO + rc.tobytes()
    exp, F, best, which = oracle_run(port, po, cfg, genome, off, names, reads, lens)
    assert (rec["mapped"]["orientation"] == best["orientation"]).all() and (rec["mapped"]["pass"] == which).all()
    assert int(rec["alignments"]["cigar_len"].max()) <= 256
    assert sam == exp, first_diff(sam, exp)
    m = best["orientation"] != 3
    assert m.sum() > 0.3 * len(lens) and (F["num_conversions"][m].sum() > 0)
    # chunks of a run concatenate: records of [0, h) + records of [h, n) with first_read_id = h
    h = len(lens) // 3
    _, rc1, _ = mp.mapReadsSam(reads[:h], lens[:h], cigar_pitch=256)
    _, rc2, _ = mp.mapReadsSam(reads[h:], lens[h:], first_read_id=h, cigar_pitch=256)
    assert rc1.tobytes() + rc2.tobytes() == rc.tobytes()


def test_staged_pipeline(cuda, port):
    """hrm_mapper_stage_reads / map_staged / finish: batches in flight in two slots give the bytes the one-shot calls give"""
    import torch
    import hashreadmapper_b200._lib as L
    genome, off = synth.make_genome([90_000, 40_001], seed=51)
    cfg = cuda.directional_config()
    mp = cuda.Mapper(cfg)
    mp.setGenome(genome, off, ["chrA", "chrB"])
    sizes = [3000, 1, 2500, 0, 700]
    batches, first = [], 0
    for i, n in enumerate(sizes):
        reads, lens, _ = synth.make_reads(genome, off, max(n, 1), 150, error_rate=0.02, seed=60 + i)
        reads, lens = reads[:n], lens[:n]
        batches.append((torch.from_numpy(reads).pin_memory().numpy() if n else reads, lens, first))
        first += n
    exp = []
    for reads, lens, fid in batches:
        if len(lens) == 0:
            exp.append((b"", b"", None, None))
            continue
        sq, rc, _, rec, cig = mp.mapReadsSam(reads, lens, first_read_id=fid, cigar_pitch=128, want_records=True)
        exp.append((sq.tobytes(), rc.tobytes(), rec.copy(), cig.copy()))
    outs = []
    for reads, lens, fid in batches:
        n = len(lens)
        outs.append({"rec": np.zeros(max(n, 1), dtype=L.RECORD_DTYPE), "cig": np.zeros((2 * max(n, 1), 128), np.uint8),
                     "sq": np.zeros(max(n, 1) * 40, np.uint8), "txt": np.zeros(max(n, 1) * 600, np.uint8)})
    mp.stageReads(0, batches[0][0], batches[0][1])
    sizes_out = [None] * len(batches)
    for i, (reads, lens, fid) in enumerate(batches):
        o = outs[i]
        mp.mapStaged(i % 2, o["rec"], o["cig"], 128, fid, o["sq"], o["txt"])
        if i >= 1:
            sizes_out[i - 1] = mp.finish((i - 1) % 2)
        if i + 1 < len(batches):
            mp.stageReads((i + 1) % 2, batches[i + 1][0], batches[i + 1][1])
    sizes_out[-1] = mp.finish((len(batches) - 1) % 2)
    # staging into a slot whose text was never fetched is refused
    import hashreadmapper_b200 as hb
    mp.stageReads(0, batches[0][0], batches[0][1])
    mp.mapStaged(0, None, None, 128, 0, outs[0]["sq"], outs[0]["txt"])
    with pytest.raises(hb.HrmError):
        mp.stageReads(0, batches[0][0], batches[0][1])
    assert mp.finish(0) == sizes_out[0]
    for i, (reads, lens, fid) in enumerate(batches):
        n = len(lens)
        sqw, recw = sizes_out[i]
        assert outs[i]["sq"][:sqw].tobytes() == exp[i][0] and outs[i]["txt"][:recw].tobytes() == exp[i][1], i
        if n:
            for f in ("orientation", "hamming_distance", "shift", "chromosome_id", "position", "pass"):
                assert (outs[i]["rec"]["mapped"][f][:n] == exp[i][2]["mapped"][f]).all(), f
            assert (outs[i]["rec"]["alignments"][:n] == exp[i][2]["alignments"]).all()


def test_fastq_text_to_sam_text(cuda, port):
    """FASTQ text in host memory -> hrm_mapper_stage_fastq (H2D + device-side reader) -> map -> SAM text, staged from a
    second host thread while the other slot computes; same bytes as the row-based entry point on the parsed reads"""
    import threading
    import torch
    genome, off = synth.make_genome([120_000], seed=71)
    mp = cuda.Mapper(cuda.directional_config())
    mp.setGenome(genome, off, ["chrF"])
    texts, rows_, first = [], [], 0
    for b, n in enumerate([2000, 1500, 2500]):
        reads, lens, _ = synth.make_reads(genome, off, n, 150, error_rate=0.02, seed=80 + b)
        reads[3, 10] = ord("N")  # replaced by the reader's ACGT cycle
        recs = []
        for i in range(n):
            sq = bytes(reads[i, :lens[i]]).decode()
            recs.append("@r%d\n%s\n+\n%s\n" % (first + i, sq, "I" * len(sq)))
        t = np.frombuffer("".join(recs).encode(), dtype=np.uint8).copy()
        texts.append(torch.from_numpy(t).pin_memory().numpy())
        first += n
    # expected: parse each text on the device, then the one-shot row entry point
    exp, fid, carry = [], 0, 0
    for t in texts:
        r, l, a, c2 = cuda.ingest_reads(torch.from_numpy(t).cuda(), 160, 4000, fid, carry)
        sq, rc, _ = mp.mapReadsSam(r.cpu().numpy(), l.cpu().numpy(), first_read_id=fid, cigar_pitch=128)
        exp.append((sq.tobytes(), rc.tobytes()))
        fid += r.shape[0]
        carry = c2
    outs = [{"sq": np.zeros(4000 * 40, np.uint8), "txt": np.zeros(4000 * 600, np.uint8)} for _ in texts]
    state = {"fid": 0, "carry": 0, "n": [0] * len(texts), "first": [0] * len(texts)}

    def stage(i):
        state["first"][i] = state["fid"]
        n, c2 = mp.stageFastq(i % 2, texts[i], 160, 4000, state["fid"], state["carry"])
        state["n"][i] = n
        state["fid"] += n
        state["carry"] = c2
    stage(0)
    sizes = [None] * len(texts)
    for i in range(len(texts)):
        if i >= 1:
            sizes[i - 1] = mp.finish((i - 1) % 2)
        th = None
        if i + 1 < len(texts):  # parse the next batch (other slot, fetched above) while this one maps
            th = threading.Thread(target=stage, args=(i + 1,))
            th.start()
        mp.mapStaged(i % 2, None, None, 128, state["first"][i], outs[i]["sq"], outs[i]["txt"])
        if th is not None:
            th.join()
    sizes[-1] = mp.finish((len(texts) - 1) % 2)
    for i in range(len(texts)):
        assert outs[i]["sq"][:sizes[i][0]].tobytes() == exp[i][0], i
        assert outs[i]["txt"][:sizes[i][1]].tobytes() == exp[i][1], i
