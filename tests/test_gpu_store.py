"""GPU tests of S1 (read storage: hrm_readstore_*, ref GpuReadStorage include/gpu/gpureadstorage.cuh:22-119), of
ReferenceWindows (hrm_genome_window_info, ref include/referencewindows.hpp:31-62, genome.hpp:176-209) and of the two
adaptor classes driven through the reference's VIRTUAL interfaces by a C++ program compiled against the reference's
headers (tests/adaptor_prog/adaptor_check.cu, built by oracle/Makefile where /root/reference exists)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from util import rs as rand_seq, rows, pack_rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTOR = os.path.join(ROOT, "oracle", "_ref", "adaptor_check")


def make_reads(seed, n, with_n=True):
    import random
    rng = random.Random(seed)
    seqs = []
    for i in range(n):
        L = rng.choice([150, 150, 150, 100, 36, 151, 1, 16, 17])
        alpha = "ACGTN" if (with_n and i % 7 == 3) else "ACGT"
        seqs.append(rand_seq(rng, L, alpha))
    return seqs


def test_s1_readstore(cuda, port):
    import torch
    seqs = make_reads(1, 2000)
    seqs[5] = b"ACGT" * 30 + b"NNACGTNACG"
    a, lens = rows(seqs)
    exp = pack_rows(port, seqs)  # non-ACGT packs as A (sequencehelpers.hpp:195-211)
    amb = np.array([any(c not in b"ACGT" for c in s) for s in seqs])
    st = cuda.ReadStorage(a, lens)
    info = st.getInfo()
    assert info.num_reads == len(seqs) and info.pitch_words == exp.shape[1]
    assert (info.length_lower_bound, info.length_upper_bound) == (int(lens.min()), int(lens.max()))
    assert info.num_reads_with_n == int(amb.sum()) and info.is_paired_end == 0
    assert info.device_bytes >= exp.nbytes
    h = st.makeHandle()
    h2 = st.makeHandle()
    assert h != h2
    ids = np.random.RandomState(2).randint(0, len(seqs), size=5000).astype(np.uint32)
    d_ids = torch.from_numpy(ids.view(np.int32)).cuda()
    got = st.gatherSequences(h, d_ids).cpu().numpy().view(np.uint32)
    assert (got == exp[ids]).all()
    # a wider output pitch is zero filled, a narrower one truncates (callers size it from the length upper bound)
    wide = st.gatherSequences(h2, d_ids, out_pitch_words=exp.shape[1] + 3).cpu().numpy().view(np.uint32)
    assert (wide[:, :exp.shape[1]] == exp[ids]).all() and (wide[:, exp.shape[1]:] == 0).all()
    assert (st.gatherSequenceLengths(h, d_ids).cpu().numpy() == lens[ids]).all()
    c = st.gatherContiguousSequences(h, 700, 900).cpu().numpy().view(np.uint32)
    assert (c == exp[700:1600]).all()
    assert (st.areSequencesAmbiguous(h, d_ids).cpu().numpy().astype(bool) == amb[ids]).all()
    assert (st.getIdsOfAmbiguousReads() == np.nonzero(amb)[0]).all()
    assert st.gatherSequences(h, d_ids[:0]).shape[0] == 0
    st.destroyHandle(h2)
    with pytest.raises(Exception):
        st.gatherSequences(h2, d_ids)  # destroyed handle
    with pytest.raises(Exception):
        st.destroyHandle(h2)
    # 3N store: conversion applied while packing
    st_ct = cuda.ReadStorage(a, lens, conversion=1)
    exp_ct = pack_rows(port, [port.convert_ascii(s, 1) for s in seqs])
    assert (st_ct.gatherContiguousSequences(st_ct.makeHandle(), 0, len(seqs)).cpu().numpy().view(np.uint32) == exp_ct).all()
    # from packed rows + carried ambiguity flags (what hrm_ingest_reads hands over)
    d_rows = torch.from_numpy(exp.view(np.int32)).cuda()
    st2 = cuda.ReadStorage.from2Bit(d_rows, torch.from_numpy(lens).cuda(), torch.from_numpy(amb.astype(np.uint8)).cuda())
    hh = st2.makeHandle()
    assert (st2.gatherSequences(hh, d_ids).cpu().numpy().view(np.uint32) == exp[ids]).all()
    assert st2.getNumberOfReadsWithN() == int(amb.sum())
    assert (st2.areSequencesAmbiguous(hh, d_ids).cpu().numpy().astype(bool) == amb[ids]).all()
    # empty store
    e = cuda.ReadStorage(np.zeros((0, 16), np.uint8), np.zeros(0, np.int32))
    assert e.getNumberOfReads() == 0


@pytest.mark.parametrize("k,w", [(16, 128), (12, 64), (32, 256), (20, 20)])
def test_reference_windows(cuda, k, w):
    """global window id -> (chromosome, window id, position, length): Genome's enumeration (genome.hpp:176-209: window i
    of a chromosome starts at i * (w - k + 1), the last ones are cut at the chromosome end) restated with numpy"""
    lens = [1000, 113 * 7, 113 * 7 + 1, 5, w, w + 1, 40000]
    g = cuda.Genome(b"".join(b"A" * n for n in lens), np.concatenate([[0], np.cumsum(lens)]))
    stride = w - k + 1
    nwin = [(n + stride - 1) // stride for n in lens]
    assert g.getTotalNumWindows(k, w) == sum(nwin)
    for c, n in enumerate(lens):
        assert g.getNumWindowsInChromosome(c, k, w) == nwin[c]
    base = np.concatenate([[0], np.cumsum(nwin)])
    for gw in list(range(0, 30)) + [int(x) for x in base[1:] - 1] + [int(x) for x in base[:-1]] + [sum(nwin) - 1]:
        c = int(np.searchsorted(base, gw, side="right") - 1)
        wid = gw - int(base[c])
        pos = wid * stride
        assert g.windowInfo(k, w, gw) == (c, wid, pos, min(w, lens[c] - pos)), gw
    with pytest.raises(Exception):
        g.windowInfo(k, w, sum(nwin))


def test_adaptors_through_the_virtual_interface(cuda, port, tmp_path):
    """B200Minhasher / B200ReadStorage called through care::gpu::GpuMinhasher* / GpuReadStorage* by a program compiled
    against the reference's own headers: addHashTables -> insert x2 -> compact -> determineNumValues -> retrieveValues,
    gatherSequences / gatherContiguousSequences / gatherSequenceLengths / areSequencesAmbiguous -- same results as
    the ctypes path and as the oracle's tables"""
    import torch
    if not os.path.exists(ADAPTOR):
        pytest.skip("oracle/_ref/adaptor_check not built (needs /root/reference at build time)")
    import random
    rng = random.Random(7)
    G = rand_seq(rng, 30000, "AGT")
    n, nq, k, H, cap = 1500, 400, 16, 8, 5
    seqs = [G[p:p + 150] for p in (rng.randrange(0, len(G) - 150) for _ in range(n))]
    seqs[3] = seqs[3][:60] + b"N" + seqs[3][61:]
    seqs[9] = b"ACGTACGTAC"  # shorter than k: never inserted
    qs = [G[p:p + 128] for p in (rng.randrange(0, len(G) - 128) for _ in range(nq))]
    qs[1] = b"ACG"
    a, lens = rows(seqs, pitch=160)
    qa, qlens = rows(qs, pitch=160)
    gids = np.random.RandomState(3).randint(0, n, size=777).astype(np.uint32)
    fin = tmp_path / "in.bin"
    fout = tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<7q", n, 160, nq, len(gids), k, H, cap))
        f.write(lens.tobytes() + a.tobytes() + qlens.tobytes() + qa.tobytes() + gids.tobytes())
    r = subprocess.run([ADAPTOR, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(fout, "rb").read()
    pw, total, namb = struct.unpack_from("<3q", buf, 0)
    at = 24

    def take(dtype, count):
        nonlocal at
        v = np.frombuffer(buf, dtype=dtype, count=count, offset=at)
        at += v.nbytes
        return v
    g = take(np.uint32, len(gids) * pw).reshape(len(gids), pw)
    glen = take(np.int32, len(gids))
    gamb = take(np.uint8, len(gids))
    ambig = take(np.uint32, namb)
    allrows = take(np.uint32, n * pw).reshape(n, pw)
    num = take(np.int32, nq)
    off = take(np.int32, nq + 1)
    vals = take(np.uint32, total)
    info = take(np.int64, 6)
    assert at == len(buf)
    exp = pack_rows(port, seqs)
    assert pw == exp.shape[1] and (allrows == exp).all() and (g == exp[gids]).all() and (glen == lens[gids]).all()
    assert ambig.tolist() == [3] and (gamb.astype(bool) == (gids == 3)).all()
    assert info.tolist() == [n, int(lens.min()), int(lens.max()), H, k, cap]
    # the oracle's tables (pinned to cpuhashtable.hpp / groupbykey.hpp) over the same reads
    sig, val = port.minhash_batch(exp, lens, k, H)
    qsig, qval = port.minhash_batch(pack_rows(port, qs), qlens, k, H)
    T = port.tables_build(sig, val, None, cap)
    enum, eoff, evals = port.tables_query(T, qsig, qval)
    port.tables_free(T)
    assert total == len(evals) and (num == enum).all() and (off == eoff).all() and (vals == evals).all()
    # and the ctypes path
    mh = cuda.Minhasher(n, cap, k, 0.8)
    mh.addHashTables(H)
    d_rows = torch.from_numpy(exp.view(np.int32)).cuda()
    mh.insert(d_rows, torch.from_numpy(lens).cuda(), None, 0, 0, H)
    mh.compact()
    h = mh.makeMinhasherHandle()
    d_q = torch.from_numpy(pack_rows(port, qs).view(np.int32)).cuda()
    cnum, ctotal = mh.determineNumValues(h, d_q, torch.from_numpy(qlens).cuda())
    cvals, coff = mh.retrieveValues(h, nq, ctotal, cnum)
    assert ctotal == total and (cnum.cpu().numpy() == num).all() and (cvals.cpu().numpy().view(np.uint32) == vals).all()


@pytest.mark.parametrize("same_length", [False, True])
def test_preprocessed_reads_dump(cuda, port, tmp_path, same_length):
    """the reference's preprocessed-reads dump (ChunkedReadStorage::saveToFile / loadFromFile): a dump written here is
    loaded by the reference's own loadFromFile, a dump the reference wrote is loaded here, and for the same reads the
    two files are byte-identical (up to the order of the ambiguous ids)"""
    import torch
    from oracle import pyoracle as po
    seqs = make_reads(5, 3000)
    if same_length:
        seqs = [s[:36].ljust(36, b"A") for s in seqs]
    a, lens = rows(seqs)
    norm = [bytes(c if c in b"ACGT" else ord("A") for c in s) for s in seqs]  # what the packed rows hold
    exp = pack_rows(port, seqs)
    amb = np.array([any(c not in b"ACGT" for c in s) for s in seqs])
    st = cuda.ReadStorage(a, lens)
    dump = st.saveToBytes()
    # round trip through our own loader
    st2 = cuda.ReadStorage.loadFromBytes(dump)
    h = st2.makeHandle()
    assert (st2.gatherContiguousSequences(h, 0, len(seqs)).cpu().numpy().view(np.uint32) == exp).all()
    ids = torch.arange(len(seqs), dtype=torch.int32).cuda()
    assert (st2.gatherSequenceLengths(h, ids).cpu().numpy() == lens).all()
    assert (st2.getIdsOfAmbiguousReads() == np.nonzero(amb)[0]).all()
    i1, i2 = st.getInfo(), st2.getInfo()
    assert (i1.num_reads, i1.length_lower_bound, i1.length_upper_bound, i1.num_reads_with_n, i1.pitch_words) == \
           (i2.num_reads, i2.length_lower_bound, i2.length_upper_bound, i2.num_reads_with_n, i2.pitch_words)
    with pytest.raises(Exception):
        cuda.ReadStorage.loadFromBytes(dump[:len(dump) // 2])
    if not po.have_ref_sam():
        pytest.skip("oracle/_ref/libhrm_ref_sam.so not built: the reference side of the check needs /root/reference")
    # our file -> the reference's loadFromFile
    p1 = tmp_path / "ours.bin"
    p1.write_bytes(dump)
    r_rows, r_lens, r_amb = po.ref_readstorage_load(p1, exp.shape[1], len(seqs) + 10)
    assert (r_rows == exp).all() and (r_lens == lens).all() and (r_amb == np.nonzero(amb)[0]).all()
    # the reference's saveToFile -> our loader; and the two files byte for byte
    p2 = tmp_path / "ref.bin"
    po.ref_readstorage_save(p2, np.frombuffer(b"".join(s.ljust(a.shape[1], b"\0") for s in norm), dtype=np.uint8)
                            .reshape(len(seqs), a.shape[1]), lens, np.nonzero(amb)[0])
    ref_dump = p2.read_bytes()
    st3 = cuda.ReadStorage.loadFromBytes(ref_dump)
    assert (st3.gatherContiguousSequences(st3.makeHandle(), 0, len(seqs)).cpu().numpy().view(np.uint32) == exp).all()
    assert (st3.getIdsOfAmbiguousReads() == np.nonzero(amb)[0]).all()
    # byte for byte up to the ambiguous ids, which the reference writes in the iteration order of a hash set
    na = int(amb.sum())
    assert len(ref_dump) == len(dump) and ref_dump[:len(dump) - 4 * na] == dump[:len(dump) - 4 * na]
    assert sorted(np.frombuffer(ref_dump[len(dump) - 4 * na:], dtype=np.uint32).tolist()) == np.nonzero(amb)[0].tolist()
