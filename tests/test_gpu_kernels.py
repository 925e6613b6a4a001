"""GPU parity tests, kernel by kernel, through the C ABI (hashreadmapper_b200.api -> libhrm_b200.so),
against the oracle (plain-C restatement of the reference, itself pinned against the reference's own
code in test_oracle_pin.py).  Bit-exact: all integer / byte / index work.
"""
import random

import numpy as np
import pytest

from util import rs, mutate, rows, pack_rows, ssw_cases

pytestmark = pytest.mark.gpu


def tt(api, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def as_u32(t):
    return t.cpu().numpy().view(np.uint32)


def as_u64(t):
    return t.cpu().numpy().view(np.uint64)


# ---- K1 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("conv", [0, 1, 2])
def test_k1_encode_rows(cuda, port, conv):
    rng = random.Random(10 + conv)
    seqs = [rs(rng, rng.randint(0, 260), "ACGTNacgt") for _ in range(300)] + [b"", b"A", b"C" * 16, b"G" * 17, b"T" * 256]
    a, lens = rows(seqs, fill=ord("T"))  # padding bytes are garbage on purpose
    out = as_u32(cuda.encode_2bit(tt(cuda, a), tt(cuda, lens), conv))
    for i, s in enumerate(seqs):
        exp = port.encode_2bit(port.convert_ascii(s, conv)) if s else np.zeros(0, np.uint32)
        assert (out[i, :len(exp)] == exp).all(), (i, s)
        assert (out[i, len(exp):] == 0).all()


def test_k1_encode_contiguous_unaligned(cuda, port):
    import torch
    rng = random.Random(3)
    g = rs(rng, 100003, "ACGTN")
    buf = torch.from_numpy(np.frombuffer(b"xxxxx" + g, dtype=np.uint8).copy()).cuda()
    for conv in (0, 1, 2):
        exp = port.encode_2bit(port.convert_ascii(g, conv))
        assert (as_u32(cuda.encode_2bit_contiguous(buf[5:], conv)) == exp).all()
        al = torch.from_numpy(np.frombuffer(g, dtype=np.uint8).copy()).cuda()
        assert (as_u32(cuda.encode_2bit_contiguous(al, conv)) == exp).all()


# ---- K2 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,H", [(16, 16), (4, 1), (12, 5), (20, 48), (31, 8), (32, 16)])
def test_k2_minhash_rows(cuda, port, k, H):
    rng = random.Random(k * 100 + H)
    seqs = [rs(rng, rng.choice([150, 150, 250, 15, 16, 17, 31, 32, 33, 1, 0, 100]), "ACGT") for _ in range(200)]
    seqs += [rs(rng, 150, "AGT") for _ in range(50)]
    enc = pack_rows(port, seqs, 17)
    lens = np.array([len(s) for s in seqs], np.int32)
    es, ev = port.minhash_batch(enc, lens, k, H)
    sigs, valid = cuda.minhash(tt(cuda, enc.view(np.int32)), tt(cuda, lens), k, H)
    assert (as_u64(sigs) == es).all()
    assert (valid.cpu().numpy() == ev).all()


@pytest.mark.parametrize("k,w", [(16, 128), (12, 64), (24, 256), (32, 128)])
def test_k2_minhash_windows(cuda, port, k, w):
    rng = random.Random(k + w)
    g = rs(rng, 20011, "ACGT")
    genome = cuda.Genome(g, [0, len(g)])
    stride = w - k + 1
    nw = genome.getNumWindowsInChromosome(0, k, w)
    assert nw == (len(g) + stride - 1) // stride
    sigs, valid = cuda.minhash_windows(genome.chromosome2BitPtr(0), len(g), k, w, 16, 0, nw)
    wins = [g[i * stride:i * stride + w] for i in range(nw)]
    enc = pack_rows(port, wins, (w + 15) // 16)
    es, ev = port.minhash_batch(enc, np.array([len(x) for x in wins], np.int32), k, 16)
    assert (as_u64(sigs) == es).all()
    assert (valid.cpu().numpy() == ev).all()
    # a sub-range
    s2, _ = cuda.minhash_windows(genome.chromosome2BitPtr(0), len(g), k, w, 16, 7, 20)
    assert (as_u64(s2) == es[7:27]).all()


# ---- K3 ------------------------------------------------------------------------------------------
def _query_both(cuda, port, mh, handle, orc_tables, qenc, qlens, k, H):
    num, total = mh.determineNumValues(handle, tt(cuda, qenc.view(np.int32)), tt(cuda, qlens))
    vals, offs = mh.retrieveValues(handle, len(qlens), total, num)
    qs, qv = port.minhash_batch(qenc, qlens, k, H)
    en, eo, evals = port.tables_query(orc_tables, qs, qv)
    assert total == len(evals)
    assert (num.cpu().numpy() == en).all()
    assert (offs.cpu().numpy().astype(np.int64) == eo).all()
    assert (as_u32(vals) == evals).all()  # same order: tables 0..H-1, insertion order inside a bucket


@pytest.mark.parametrize("cap", [65535, 3])
def test_k3_minhasher_reference_direction(cuda, port, cap):
    """reads inserted, windows query -- the reference's direction (gpuminhasherconstruction.cu, main_gpu.cu:534)"""
    rng = random.Random(77 + cap)
    k, H, w = 16, 16, 128
    g = rs(rng, 30000, "AGT")
    reads = []
    for _ in range(1500):
        p = rng.randint(0, len(g) - 150)
        reads.append(mutate(rng, g[p:p + 150], 0.01, 0)[:150])
    reads += [b"ACGT", b""] + [reads[0]] * 8  # too-short reads are skipped; duplicates share buckets
    enc = pack_rows(port, reads, 10)
    lens = np.array([len(r) for r in reads], np.int32)
    mh = cuda.Minhasher(len(reads), cap, k, 0.8)
    assert mh.addHashTables(H, list(range(H))) == H
    half = len(reads) // 2
    # two insert batches, the second split over two groups of tables
    mh.insert(tt(cuda, enc[:half].view(np.int32)), tt(cuda, lens[:half]), None, 0)
    mh.insert(tt(cuda, enc[half:].view(np.int32)), tt(cuda, lens[half:]), None, half, 0, 8)
    mh.insert(tt(cuda, enc[half:].view(np.int32)), tt(cuda, lens[half:]), None, half, 8, 8)
    mh.compact()
    mh.constructionIsFinished()
    info = mh.getInfo()
    assert info.is_compacted == 1 and info.num_inserted == len(reads)
    rsig, rval = port.minhash_batch(enc, lens, k, H)
    T = port.tables_build(rsig, rval, None, cap)
    stride = w - k + 1
    wins = [g[i * stride:i * stride + w] for i in range((len(g) + stride - 1) // stride)] + [b"ACG", rs(rng, 128)]
    qenc = pack_rows(port, wins, 8)
    qlens = np.array([len(x) for x in wins], np.int32)
    h = mh.makeMinhasherHandle()
    _query_both(cuda, port, mh, h, T, qenc, qlens, k, H)
    # stage order is enforced (ref: fakegpuminhasher.cuh:328)
    import hashreadmapper_b200 as hb
    with pytest.raises(hb.HrmError):
        mh.retrieveValues(h, len(qlens), 0, tt(cuda, np.zeros(len(qlens), np.int32)))
    # serialisation round trip
    mh2 = cuda.Minhasher.loadFromBytes(mh.writeToBytes())
    h2 = mh2.makeMinhasherHandle()
    _query_both(cuda, port, mh2, h2, T, qenc, qlens, k, H)
    mh.destroyHandle(h)
    port.tables_free(T)


@pytest.mark.parametrize("cap", [65535, 3])
def test_k3_reference_hashtable_files(cuda, port, ref, tmp_path, cap):
    """the reference's own hash-table file format (--save-hashtables-to / --load-hashtables-from), both ways:
    (a) tables built on the GPU, written by hrm_minhasher_write_reference_format, loaded by the REFERENCE's
    loadFromStream and queried by the reference; (b) tables built and written by the reference's code, read by
    hrm_minhasher_read_reference_format and queried on the GPU"""
    rng = random.Random(501 + cap)
    k, H, w = 16, 8, 128
    g = rs(rng, 20000, "AGT")
    reads = []
    for _ in range(900):
        p = rng.randint(0, len(g) - 150)
        reads.append(mutate(rng, g[p:p + 150], 0.01, 0)[:150])
    reads += [b"ACGT"] + [reads[0]] * 6
    enc = pack_rows(port, reads, 10)
    lens = np.array([len(r) for r in reads], np.int32)
    rsig, rval = port.minhash_batch(enc, lens, k, H)
    stride = w - k + 1
    wins = [g[i * stride:i * stride + w] for i in range((len(g) + stride - 1) // stride)] + [rs(rng, 128)]
    qenc = pack_rows(port, wins, 8)
    qlens = np.array([len(x) for x in wins], np.int32)
    qsig, qval = port.minhash_batch(qenc, qlens, k, H)
    # (a) GPU tables -> file -> the reference loads and answers
    mh = cuda.Minhasher(len(reads), cap, k, 0.8)
    assert mh.addHashTables(H, list(range(H))) == H
    mh.insert(tt(cuda, enc.view(np.int32)), tt(cuda, lens), None, 0)
    mh.compact()
    path_a = tmp_path / "ours.tables"
    assert mh.writeToStream(str(path_a)) == path_a.stat().st_size
    Ta, k_a, lf_a = ref.tables_load(path_a)
    assert k_a == k and abs(lf_a - 0.8) < 1e-6
    Tref = ref.tables_build(rsig, rval, None, cap, 0.8, 1)
    na, oa, va = ref.tables_query(Ta, qsig, qval)
    nr, orf, vr = ref.tables_query(Tref, qsig, qval)
    assert (na == nr).all() and (oa == orf).all() and (va == vr).all() and len(vr) > 100
    # (b) the reference's tables -> its own writeToStream -> read here -> queried on the GPU
    path_b = tmp_path / "ref.tables"
    ref.tables_save(Tref, path_b, k, 0.8)
    mh2 = cuda.Minhasher.loadFromStream(str(path_b))
    info = mh2.getInfo()
    assert info.k == k and info.num_tables == H and info.max_results_per_map == cap
    h2 = mh2.makeMinhasherHandle()
    _query_both(cuda, port, mh2, h2, port.tables_build(rsig, rval, None, cap), qenc, qlens, k, H)
    # and a file of ours read back by ourselves
    mh3 = cuda.Minhasher.loadFromStream(str(path_a), 5)  # numMapsUpperLimit
    assert mh3.getInfo().num_tables == 5
    # byte-identical where the format leaves no freedom: same size, same header, same values and key tables
    a, b = np.fromfile(path_a, dtype=np.uint8), np.fromfile(path_b, dtype=np.uint8)
    assert a.size == b.size and (a[:16] == b[:16]).all()
    ref.tables_free(Ta)
    ref.tables_free(Tref)


def test_k3_empty_and_all_miss(cuda, port):
    rng = random.Random(5)
    k, H = 16, 4
    reads = [rs(rng, 100) for _ in range(64)]
    enc = pack_rows(port, reads, 7)
    lens = np.array([len(r) for r in reads], np.int32)
    mh = cuda.Minhasher(64, 65535, k, 0.8)
    assert mh.addHashTables(H) == H
    mh.insert(tt(cuda, enc.view(np.int32)), tt(cuda, lens))
    mh.compact()
    h = mh.makeMinhasherHandle()
    q = [rs(rng, 100) for _ in range(33)]
    qenc = pack_rows(port, q, 7)
    qlens = np.array([len(x) for x in q], np.int32)
    num, total = mh.determineNumValues(h, tt(cuda, qenc.view(np.int32)), tt(cuda, qlens))
    assert total == 0 and int(num.sum()) == 0
    vals, offs = mh.retrieveValues(h, len(q), total, num)
    assert (offs.cpu().numpy() == 0).all()


# ---- K4 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("min_hits", [4, 1, 2, 16])
def test_k4_filter_by_frequency(cuda, port, min_hits):
    rng = np.random.RandomState(min_hits)
    sizes = list(rng.randint(0, 40, size=500)) + [0, 1, 255, 256, 257, 1000, 5000, 9000, 20000, 0]
    rng.shuffle(sizes)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    vals = np.zeros(int(offs[-1]), np.uint32)
    for i, n in enumerate(sizes):
        hi = max(2, n // 3 + 1)
        vals[offs[i]:offs[i + 1]] = rng.randint(0, hi, size=n).astype(np.uint32) * 65537 + (1 << 31) * (i % 2)
    ev, eo = port.filter_by_frequency(vals, offs, min_hits)
    dv = tt(cuda, vals.view(np.int32))
    dnum = tt(cuda, np.array(sizes, np.int32))
    doff = tt(cuda, offs.astype(np.int32))
    total = cuda.filter_by_frequency(dv, dnum, doff, min_hits)
    assert total == len(ev)
    assert (doff.cpu().numpy().astype(np.int64) == eo).all()
    assert (dnum.cpu().numpy().astype(np.int64) == np.diff(eo)).all()
    assert (as_u32(dv)[:total] == ev).all()
    seg = cuda.segment_ids(doff, total).cpu().numpy()
    exp_seg = np.repeat(np.arange(len(sizes)), np.diff(eo))
    assert (seg == exp_seg).all()


def test_k4_filter_counting_edge_cases(cuda, port):
    """counting path of K4: the id 0xFFFFFFFF (the table's empty marker), one id repeated, all distinct,
    segments around the warp/block/multi-pass boundaries, human-scale segments with 1-3 survivors"""
    rng = np.random.RandomState(99)
    segs = []
    for n in (1, 3, 4, 5, 255, 256, 257, 600, 5119, 5120, 5121, 12000, 30000):
        segs.append(np.full(n, 0xFFFFFFFF, np.uint32))                       # only the marker value
        segs.append(np.full(n, 12345, np.uint32))                            # one id
        segs.append(rng.permutation(n).astype(np.uint32) * 7919)             # all distinct
        mixed = rng.randint(0, 1 << 31, size=n).astype(np.uint32) * 2        # random + planted survivors
        for v, reps in ((0xFFFFFFFF, 4), (0xFFFFFFFE, 5), (0, 4), (77, 3), (2 ** 31 + 5, 16)):
            if n >= 40:
                mixed[rng.choice(n, reps, replace=False)] = v
        segs.append(mixed)
    rng.shuffle(segs)
    sizes = [len(x) for x in segs]
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    vals = np.concatenate(segs)
    for min_hits in (4, 2):
        ev, eo = port.filter_by_frequency(vals, offs, min_hits)
        dv = tt(cuda, vals.view(np.int32))
        dnum = tt(cuda, np.array(sizes, np.int32))
        doff = tt(cuda, offs.astype(np.int32))
        total = cuda.filter_by_frequency(dv, dnum, doff, min_hits)
        assert total == len(ev)
        assert (doff.cpu().numpy().astype(np.int64) == eo).all()
        assert (as_u32(dv)[:total] == ev).all()


# ---- S2 / S3 -------------------------------------------------------------------------------------
def test_s2_extended_windows(cuda, port):
    rng = random.Random(8)
    g = rs(rng, 5000, "ACGT")
    genome = cuda.Genome(g, [0, len(g)])
    w, k = 128, 16
    stride = w - k + 1
    nw = (len(g) + stride - 1) // stride
    pos = np.array([i * stride for i in range(nw)] * 3, np.int32)
    rl = np.array([rng.choice([150, 100, 250, 36]) for _ in pos], np.int32)
    pw = (w + 250 + 15) // 16
    out, left, right, ln = genome.extendedWindows(0, w, tt(cuda, pos), tt(cuda, rl), pw)
    out = as_u32(out)
    for i in range(len(pos)):
        M = 250
        secB = max(0, 0 - M // 2)
        secE = min(len(g), int(pos.max()) + w + M // 2)
        l, r, length, sp = port.window_location(secB, secE, int(pos[i]), w, int(rl[i]) // 2)
        assert (int(left[i]), int(right[i]), int(ln[i])) == (l, r, length)
        exp = port.encode_2bit(g[secB + sp:secB + sp + length])
        assert (out[i, :len(exp)] == exp).all()
        assert (out[i, len(exp):] == 0).all()


def test_s3_shifted_hamming(cuda, port):
    rng = random.Random(9)
    anchors, cands = [], []
    for it in range(1500):
        Lc = rng.choice([150, 150, 250, 100, 36, 33, 32, 31])
        La = Lc + rng.randint(-2, 160)
        La = max(La, 1)
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
        anchors.append(A)
        cands.append(c)
    for rate in (0.05, 0.1):
        ea = pack_rows(port, anchors, 26)
        ec = pack_rows(port, cands, 16)
        la = np.array([len(x) for x in anchors], np.int32)
        lc = np.array([len(x) for x in cands], np.int32)
        sh, sc, ori = cuda.shifted_hamming(tt(cuda, ea.view(np.int32)), tt(cuda, la), tt(cuda, ec.view(np.int32)),
                                           tt(cuda, lc), rate)
        sh, sc, ori = sh.cpu().numpy(), sc.cpu().numpy(), ori.cpu().numpy()
        nacc = 0
        for i in range(len(anchors)):
            exp = port.shd(ea[i], int(la[i]), ec[i], int(lc[i]), rate)
            assert int(ori[i]) == exp[2], (i, exp)
            if exp[2] != 3:  # shift/score of rejected candidates are unspecified
                nacc += 1
                assert (int(sh[i]), int(sc[i])) == exp[:2], (i, exp)
            elif lc[i] > la[i]:
                assert (int(sh[i]), int(sc[i])) == exp[:2]
        assert nacc > 300


# ---- V2 / V3 -------------------------------------------------------------------------------------
def test_v2_sw_align(cuda, port):
    cases = ssw_cases(21, 3000)
    qa, ql = rows([c[0] for c in cases], 272)
    ra, rl = rows([c[1] for c in cases], 272)
    ml = np.array([c[2] for c in cases], np.int32)
    al, cigs = cuda.sw_align(tt(cuda, qa), tt(cuda, ql), tt(cuda, ra), tt(cuda, rl), tt(cuda, ml), cigar_pitch=512)
    bad = 0
    for i, (q, r, m) in enumerate(cases):
        exp, ecig = port.ssw_align(q, r, m)
        got = tuple(int(al[n][i]) for n in al.dtype.names[:9])
        if got != exp or cigs[i] != ecig:
            bad += 1
            if bad < 5:
                print("MISMATCH", i, exp, ecig, got, cigs[i])
    assert bad == 0


def test_v3_edit_distance(cuda, port):
    rng = random.Random(4)
    qs, ts = [], []
    for it in range(1500):
        q = rs(rng, rng.randint(1, 300), "AGTN")
        t = mutate(rng, q, 0.05, 0.05) if it % 2 else rs(rng, rng.randint(1, 200), "AGT")
        qs.append(q)
        ts.append(t or b"G")
    qa, ql = rows(qs, 304)
    ta, tl = rows(ts, 400)
    d = cuda.edit_distance(tt(cuda, qa), tt(cuda, ql), tt(cuda, ta), tt(cuda, tl)).cpu().numpy()
    for i in range(len(qs)):
        assert int(d[i]) == port.edit_distance_nw(qs[i], ts[i]), i


# ---- 8f-1 read ingestion ---------------------------------------------------------------------------
def test_ingest_reads(cuda, tmp_path):
    """FASTQ / FASTA text -> normalised ASCII rows on the device, against the restatement of the reference's
    reader + preprocessSequence (65536-read replacement cycle, ambiguous flags, chunk carry, error paths)"""
    import torch
    import hashreadmapper_b200 as hb
    from oracle.pyoracle import parse_records_model, preprocess_reads_model
    rng = random.Random(21)

    def make(fmt, n, maxlen, trailing_newline=True):
        recs = []
        for i in range(n):
            sq = "".join(rng.choice("ACGTACGTACGTNacgtnR") for _ in range(rng.randint(1, maxlen)))
            if fmt == "fastq":
                recs.append("@r%d x\n%s\n+\n%s\n" % (i, sq, "I" * len(sq)))
            elif fmt == "fasta":
                recs.append(">r%d\n%s\n" % (i, sq))
            else:  # sequences over several lines (60-column files), CRLF line ends in some records
                w = rng.choice([60, 7, 200])
                eol = "\r\n" if i % 5 == 0 else "\n"
                recs.append(">r%d%s%s%s" % (i, eol, eol.join(sq[j:j + w] for j in range(0, len(sq), w)), eol))
        text = "".join(recs).encode()
        return text if trailing_newline else text[:-1]

    def run(text, pitch, first=0, carry=0, cap=None):
        t = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).cuda()
        return cuda.ingest_reads(t, pitch, cap or text.count(b"\n") + 1, first, carry)

    for fmt, n, maxlen, nl in (("fastq", 3000, 150, True), ("fasta", 2000, 100, False), ("fastq", 70000, 24, True),
                               ("fastq", 1, 5, False), ("fasta_ml", 3000, 150, True), ("fasta_ml", 66000, 30, False)):
        text = make(fmt, n, maxlen, nl)
        seqs = parse_records_model(text)
        exp, amb, carry_exp = preprocess_reads_model(seqs)
        rows, lens, a, carry = run(text, 160)
        assert rows.shape[0] == n and (lens.cpu().numpy() == [len(x) for x in exp]).all()
        r = rows.cpu().numpy()
        for i in (list(range(min(n, 400))) + list(range(max(0, n - 400), n)) + list(range(65400, min(n, 65700)))):
            assert bytes(r[i, :len(exp[i])]) == exp[i], i
            assert not r[i, len(exp[i]):].any()
        assert (a.cpu().numpy().astype(bool) == np.array(amb)).all()
        assert carry == (carry_exp if n % 65536 else 0)
    # two chunks of one file: the second starts inside a batch and continues its replacement cycle
    text = make("fastq", 5000, 60)
    seqs = parse_records_model(text)
    exp, amb, _ = preprocess_reads_model(seqs)
    cut = len(b"\n".join(text.split(b"\n")[:4 * 1234])) + 1
    r1, l1, a1, c1 = run(text[:cut], 64)
    r2, l2, a2, c2 = run(text[cut:], 64, first=1234, carry=c1)
    both = np.concatenate([r1.cpu().numpy(), r2.cpu().numpy()])
    assert both.shape[0] == 5000
    for i in range(5000):
        assert bytes(both[i, :len(exp[i])]) == exp[i], i
    # rows feed K1 directly
    enc = cuda.encode_2bit(r1, l1)
    assert enc.shape[0] == 1234
    # error paths: incomplete FASTQ record, row too short (both formats), garbage
    for bad in (b"@a\nACGT\n+\n", b"@a\n" + b"A" * 100 + b"\n+\n" + b"I" * 100 + b"\n",
                b">a\n" + b"A" * 40 + b"\n" + b"C" * 40 + b"\n", b"hello\n"):
        with pytest.raises(hb.HrmError):
            run(bad, 64)


@pytest.mark.parametrize("min_hits", [4, 2, 7])
def test_k4_against_reference_unique_by_count(cuda, min_hits):
    """C1 pinned to the REFERENCE'S OWN CUDA implementation: GpuSegmentedUniqueByCount::unique
    (include/gpu/cuda_unique_by_count.cuh:33-215, compiled from the header where it lies into oracle/_ref/libhrm_ref_c1.so)
    runs on this GPU on the same segments as hrm_filter_by_frequency"""
    from oracle import pyoracle as po
    if not po.have_ref_c1():
        pytest.skip("oracle/_ref/libhrm_ref_c1.so not built (needs /root/reference at build time)")
    rng = np.random.RandomState(100 + min_hits)
    sizes = list(rng.randint(0, 60, size=800)) + [0, 1, 255, 256, 257, 1000, 5000, 9000, 20000, 0, 4200, 4300]
    rng.shuffle(sizes)
    offs = np.zeros(len(sizes) + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    vals = np.zeros(int(offs[-1]), np.uint32)
    for i, n in enumerate(sizes):
        hi = max(2, n // 4 + 1)
        vals[offs[i]:offs[i + 1]] = rng.randint(0, hi, size=n).astype(np.uint32) * 65537 + (1 << 31) * (i % 2)
    ev, el = po.ref_unique_by_count(vals, offs.astype(np.int32), min_hits)
    dv = tt(cuda, vals.view(np.int32))
    dnum = tt(cuda, np.array(sizes, np.int32))
    doff = tt(cuda, offs.astype(np.int32))
    total = cuda.filter_by_frequency(dv, dnum, doff, min_hits)
    assert total == len(ev) and total > 0
    assert (dnum.cpu().numpy() == el).all()
    assert (as_u32(dv)[:total] == ev).all()
