"""CPU tests: the oracle (oracle/hrm_oracle.c) against (a) the golden fixture generated from the
reference's own code, (b) the known answers of SURVEY.md Appendix C, and (c) -- where
oracle/_ref/libhrm_ref.so exists -- the reference's own code on fresh random inputs.
"""
import json
import os
import random

import numpy as np
import pytest

from util import rs, mutate, ssw_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_known_answers(port):
    s = b"ACGTACGTTTGACCAGTAGGCATTACGGATCAGGCATCAGGACTTTACG"
    e = port.encode_2bit(s)
    sig, val = port.minhash_batch(e[None, :], np.array([len(s)]), 16, 4)
    assert sig[0].tolist() == [3092148114, 1517188485, 663417955, 2793855263]
    assert port.decode_2bit(e, len(s)) == s
    al, cig = port.ssw_align(b"CTGAGCCGGTAAATC", b"CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA", 15)
    assert al == (21, 8, 8, 21, 0, 14, 4, 2, 0) and cig == "4=1X4=1I5="
    assert port.edit_distance_nw(b"ACTTTGTTTGATTAG", b"ATTTTGTTGATTTAG") == 3
    assert port.convert_ascii(b"ACGTNCcG", 1) == b"ATGTNTcG"
    assert port.convert_ascii(b"ACGTNGgG", 2) == b"ACATNAgA"


def test_golden_encode_minhash(port, gold):
    for g in gold["encode"]:
        s = g["seq"].encode()
        e = port.encode_2bit(s)
        assert e.tolist() == g["words"]
        assert port.revcomp_2bit(e, len(s)).tolist() == g["rc_words"]
        assert port.revcomp_ascii(s).hex() == g["rc_ascii"]
    for g in gold["minhash"]:
        s = g["seq"].encode()
        sig, val = port.minhash_batch(port.encode_2bit(s)[None, :], np.array([len(s)]), g["k"], 16)
        assert sig[0].tolist() == g["sig"] and val[0].tolist() == g["valid"]
    for x, h in gold["murmur64"]:
        assert port.murmur64(x) == h


def test_golden_tables(port, gold):
    t = gold["tables"]
    sig = np.random.RandomState(t["sig_seed"]).randint(0, 60, size=(t["n"], t["H"])).astype(np.uint64)
    val = (np.random.RandomState(t["valid_seed"]).rand(t["n"], t["H"]) > 0.05).astype(np.uint8)
    q = np.random.RandomState(t["query_seed"]).randint(0, 70, size=(50, t["H"])).astype(np.uint64)
    qv = np.ones((50, t["H"]), np.uint8)
    qv[::7] = 0
    for cap, res in t["results"].items():
        T = port.tables_build(sig, val, None, int(cap))
        num, off, vals = port.tables_query(T, q, qv)
        assert num.tolist() == res["num"] and vals.tolist() == res["values"]
        port.tables_free(T)


def test_golden_window_shd(port, gold):
    for args, res in gold["window_location"]:
        assert list(port.window_location(*args)) == res
    for g in gold["shd"]:
        A, c = g["anchor"].encode(), g["cand"].encode()
        got = port.shd(port.encode_2bit(A), len(A), port.encode_2bit(c), len(c), g["rate"])
        assert list(got) == g["result"]


def test_golden_ssw_edit(port, gold):
    assert len(gold["ssw"]) > 300
    for g in gold["ssw"]:
        al, cig = port.ssw_align(g["q"].encode(), g["r"].encode(), g["mask"])
        assert list(al) == g["al"] and cig == g["cigar"], g
    for a, b, d in gold["edit"]:
        assert port.edit_distance_nw(a.encode(), b.encode()) == d


def test_filter_semantics(port):
    """C1 has no CPU twin in the reference; its stated semantics (cuda_unique_by_count.cuh:79-172)"""
    vals = np.array([5, 1, 5, 5, 9, 1, 5, 7, 7, 7, 7, 3], np.uint32)
    off = np.array([0, 7, 7, 12], np.int64)
    v, o = port.filter_by_frequency(vals, off, 4)
    assert v.tolist() == [5, 7] and o.tolist() == [0, 1, 1, 2]
    v, o = port.filter_by_frequency(vals, off, 1)
    assert v.tolist() == [1, 5, 9, 3, 7] and o.tolist() == [0, 3, 3, 5]
    v, o = port.filter_by_frequency(vals, off, 2)
    assert v.tolist() == [1, 5, 7]


# ---- against the reference's own code on fresh inputs (only where it was built) -------------------
def test_ref_encode_kmers_minhash(port, ref):
    rng = random.Random(1)
    for it in range(150):
        L = rng.randint(1, 300)
        s = rs(rng, L, "ACGTN" if it % 5 == 0 else "ACGT")
        a, b = port.encode_2bit(s), ref.encode_2bit(s)
        assert (a == b).all()
        assert (port.revcomp_2bit(a, L) == ref.revcomp_2bit(a, L)).all()
        assert port.revcomp_ascii(s) == ref.revcomp_ascii(s)
        for k in (4, 16, 21, 32):
            if L >= k:
                assert (port.canonical_kmers(a, L, k) == ref.canonical_kmers(a, L, k)).all()
            sa, va = port.minhash_batch(a[None, :], np.array([L]), k, 16)
            sb, vb = ref.minhash_batch(a[None, :], np.array([L]), k, 16)
            assert (sa == sb).all() and (va == vb).all()


def test_ref_tables(port, ref):
    rng = random.Random(2)
    for it in range(12):
        n = rng.randint(1, 2000)
        H = rng.choice([1, 4, 16])
        sig = np.random.RandomState(it).randint(0, rng.choice([5, 50, 5000]), size=(n, H)).astype(np.uint64)
        val = (np.random.RandomState(it + 99).rand(n, H) > 0.05).astype(np.uint8)
        cap = rng.choice([65535, 3, 10])
        hp, hr = port.tables_build(sig, val, None, cap), ref.tables_build(sig, val, None, cap)
        q = np.random.RandomState(it + 5).randint(0, 60, size=(100, H)).astype(np.uint64)
        qv = np.ones((100, H), np.uint8)
        qv[::7] = 0
        for x, y in zip(port.tables_query(hp, q, qv), ref.tables_query(hr, q, qv)):
            assert (x == y).all()
        port.tables_free(hp)
        ref.tables_free(hr)


def test_ref_shd_window(port, ref):
    rng = random.Random(3)
    for it in range(600):
        Lc = rng.randint(20, 260)
        La = max(Lc + rng.randint(-3, 140), 1)
        A = rs(rng, La, "AGT" if it % 2 else "ACGT")
        if it % 3 and La >= Lc:
            st = rng.randint(0, La - Lc)
            c = bytearray(A[st:st + Lc])
            for _ in range(rng.randint(0, 12)):
                c[rng.randrange(Lc)] = rng.choice(b"ACGT")
            c = bytes(c)
            if it % 6 == 1:
                c = port.revcomp_ascii(c)
        else:
            c = rs(rng, Lc)
        rate = rng.choice([0.05, 0.5, 0.03])
        ea, ec = port.encode_2bit(A), port.encode_2bit(c)
        assert port.shd(ea, La, ec, Lc, rate) == ref.shd(ea, La, ec, Lc, rate)
    for it in range(3000):
        args = (rng.randint(0, 500), rng.randint(500, 2000), rng.randint(500, 1900), rng.choice([64, 128, 256]),
                rng.randint(0, 130))
        assert port.window_location(*args) == ref.window_location(*args)


def test_ref_ssw_edit(port, ref):
    bad = 0
    for q, r, ml in ssw_cases(99, 2500):
        a, b = port.ssw_align(q, r, ml), ref.ssw_align(q, r, ml)
        if a != b and b[0][0] != 0:
            bad += 1
    assert bad == 0
    rng = random.Random(5)
    for it in range(500):
        q = rs(rng, rng.randint(1, 200), "AGT")
        t = (mutate(rng, q, 0.05, 0.05) if it % 2 else rs(rng, rng.randint(1, 200), "AGT")) or b"G"
        assert port.edit_distance_nw(q, t) == ref.edit_distance_nw(q, t)


def test_ref_whole_pipeline(port, ref):
    """the oracle's pipeline restatement == the reference's own functions driven in the reference's order"""
    from hashreadmapper_b200 import synth
    from oracle.pyoracle import ref_cpu_pipeline
    genome, off = synth.make_genome([50000, 20011], seed=5)
    reads, lens, _ = synth.make_reads(genome, off, 1500, 150, error_rate=0.02, seed=6)
    lens[3] = 10
    for conv_g in (1, 2):
        g = port.convert_ascii(genome, conv_g)
        exp, stats = port.map_pass_refdir(g, off, reads, lens)
        got, sw, ed, times = ref_cpu_pipeline(ref, g, off, reads, lens)
        m = exp["orientation"] != 3
        assert m.sum() > 500
        assert (got["orientation"] == exp["orientation"]).all()
        for f in ("hammingDistance", "shift", "chromosomeId", "position"):
            assert (got[f][m] == exp[f][m]).all(), f
        # verification inputs + SSW through the oracle == the reference's
        for i in np.nonzero(m)[0][:200]:
            chrom = g[off[exp["chromosomeId"][i]]:off[exp["chromosomeId"][i] + 1]]
            q, qrc, r = port.verify_inputs(reads[i, :lens[i]].tobytes(), int(exp["orientation"][i]), chrom,
                                           int(exp["position"][i]), 128, 1)
            for a, qq in enumerate((q, qrc)):
                ea, _ = port.ssw_align(qq, r, max(15, int(lens[i]) // 2))
                assert ea == tuple(int(sw[i][a][n]) for n in sw.dtype.names[:9])


def test_read_ingestion_model_against_reference_parser(ref, tmp_path):
    """8f-1: the record parsing model against the reference's own file reader (kseqpp through forEachReadInFile);
    the character normalisation is a restatement of a lambda that cannot be linked (cited in pyoracle)"""
    import random
    from oracle.pyoracle import parse_records_model, preprocess_reads_model
    rng = random.Random(12)
    for fmt in ("fastq", "fasta", "fasta_multiline"):
        seqs = ["".join(rng.choice("ACGTNacgtnRY-") for _ in range(rng.randint(1, 120))) for _ in range(700)]
        recs = []
        for i, sq in enumerate(seqs):
            if fmt == "fastq":
                recs.append("@r%d extra words\n%s\n+\n%s\n" % (i, sq, "I" * len(sq)))
            elif fmt == "fasta":
                recs.append(">r%d\n%s\n" % (i, sq))
            else:  # 60-column FASTA: the sequence continues until the next header
                w = rng.choice([60, 7, 200])
                recs.append(">r%d\n%s\n" % (i, "\n".join(sq[j:j + w] for j in range(0, len(sq), w))))
        text = "".join(recs).encode()
        path = tmp_path / ("reads." + fmt)
        path.write_bytes(text)
        rows, lens = ref.read_file(path, 128, 1000)
        got = [bytes(rows[i, :lens[i]]) for i in range(len(lens))]
        assert got == parse_records_model(text) == [s.encode() for s in seqs]
        norm, amb, _ = preprocess_reads_model(got)
        assert all(set(x) <= set(b"ACGT") for x in norm) and len(norm) == 700
        # the replacement cycle: k-th replaced character of the batch becomes "ACGT"[k % 4]
        k = 0
        for raw, fixed, a in zip(got, norm, amb):
            for c0, c1 in zip(raw, fixed):
                if c0 not in b"ACGTacgt":
                    assert c1 == b"ACGT"[k % 4]
                    k += 1
            assert a == any(c not in b"ACGTacgt" for c in raw)


# ---- V4 + O1: recalculation, alignment choice, MAPQ and the SAM text -------------------------------------------
GOLD_SAM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sam_v1.json")


@pytest.fixture(scope="module")
def gold_sam():
    with open(GOLD_SAM) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["ct150", "ga150", "ct250", "none100"])
def test_golden_sam(port, gold_sam, name):
    """the port's SAM text against the golden made by the reference's own Mappinghandler (byte for byte, by hash),
    its per-read recalculated scores / conversion counts / flags, and a few literal lines"""
    import hashlib
    import samcase
    from oracle import pyoracle as po
    g = gold_sam["cases"][name]
    case = samcase.make(port, name)
    sam, F = samcase.port_sam(po, port, case)
    per = np.array(g["per_read"], dtype=np.int64)
    m = case["mapped"]["orientation"] != 3
    assert (F["sw_score"][:, 0] == per[:, 0]).all() and (F["sw_score_next_best"][:, 0] == per[:, 1]).all()
    assert (F["sw_score"][:, 1] == per[:, 2]).all() and (F["sw_score_next_best"][:, 1] == per[:, 3]).all()
    assert (F["num_conversions"] == per[:, 4:6]).all()
    assert (F["chosen"] == (per[:, 0] < per[:, 2])).all()
    assert (F["flag"][m] == np.where(F["chosen"] == 0, per[:, 6], per[:, 7])[m]).all()
    n = len(case["lens"])
    rec = sam.split(b"\n")[n + 2:]
    for i, ln in g["lines"].items():
        assert rec[int(i)].decode() == ln
    assert len(sam) == g["bytes"] and hashlib.sha256(sam).hexdigest() == g["sha256"]
    # every branch is present in the fixture
    assert F["chosen"].sum() > 0 and (~m).sum() > 0
    if name != "none100":
        assert F["num_conversions"].sum() > 0
        assert (F["sw_score_next_best"] > 60000).any()  # the uint16 wrap of the reference's score fields


@pytest.mark.parametrize("name", ["ct150", "ga150"])
def test_ref_sam_live(port, name):
    """the same comparison against the reference's Mappinghandler run now (fresh seeds through case edits)"""
    import samcase
    from oracle import pyoracle as po
    if not po.have_ref_sam():
        pytest.skip("oracle/_ref/libhrm_ref_sam.so not built (needs /root/reference)")
    case = samcase.make(port, name, w=128)
    # move the hits around once more so that the live run is not the golden run
    rng = np.random.RandomState(5)
    mp = case["mapped"]
    idx = np.nonzero(mp["orientation"] != 3)[0]
    flip = idx[rng.rand(len(idx)) < 0.2]
    mp["orientation"][flip] = 3 - mp["orientation"][flip]
    sam_p, F = samcase.port_sam(po, port, case)
    sam_r, per = samcase.reference_sam(po, case)
    assert sam_p == sam_r
    assert (F["num_conversions"] == per[:, 4:6]).all()


def test_mapq_u16(port):
    """MAPQ (mapqfkt mappinghandler.cu:184-193): finite cases against the closed form, the out-of-range conversions
    against what the compiled reference prints (4)"""
    import math
    f = port.lib.orc_mapq_u16
    f.restype = np.ctypeslib.ctypes.c_uint32
    assert f(0, 0) == 4 and f(300, 0) == 4 and f(220, 65456) == 4 and f(100, 250) == 4
    for s1, s2 in ((300, 150), (300, 299), (300, 1), (65300, 20), (250, 249), (40, 38)):
        v = -4.343 * math.log(1 - abs(s1 - s2) / s1)
        assert f(s1, s2) == min(254, int(int(v) + 4.99))
