"""ctypes bindings for the test-only checkers.  TEST INFRASTRUCTURE ONLY.

`Oracle("port")` wraps oracle/liboracle.so (hrm_oracle.c, the plain-C restatement) and
`Oracle("ref")` wraps oracle/_ref/libhrm_ref.so (the reference's own sources compiled from
/root/reference by oracle/Makefile).  Both expose the same Python methods so that tests can
run one against the other.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module; the product (hashreadmapper_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhrm_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(want_ref=True):
    """Compile the checkers (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if want_ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def have_ref():
    return os.path.exists(REF_SO)


class Alignment(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "sw_score", "sw_score_next_best", "ref_begin", "ref_end", "query_begin", "query_end",
        "ref_end_next_best", "mismatches", "flag", "cigar_len")]

    def astuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_[:9])


class MappedRead(C.Structure):
    _fields_ = [("orientation", C.c_int32), ("hammingDistance", C.c_int32), ("shift", C.c_int32),
                ("chromosomeId", C.c_int32), ("position", C.c_int64)]


MAPPED_DTYPE = np.dtype([("orientation", "<i4"), ("hammingDistance", "<i4"), ("shift", "<i4"),
                         ("chromosomeId", "<i4"), ("position", "<i8")])
ALIGN_DTYPE = np.dtype([(n, "<i4") for n, _ in Alignment._fields_])


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO):
                build(want_ref=False)
            self.lib = C.CDLL(PORT_SO)
            self.pfx = "orc_"
        elif kind == "ref":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            self.lib = C.CDLL(REF_SO)
            self.pfx = "ref_"
        else:
            raise ValueError(kind)
        f = self._f
        f("murmur64").restype = C.c_uint64
        f("murmur64").argtypes = [C.c_uint64]
        f("edit_distance_nw").restype = C.c_int
        f("edit_distance_nw").argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        f("tables_query").restype = C.c_int64
        if kind == "port":
            f("tables_build").restype = C.c_void_p
            f("filter_by_frequency").restype = C.c_int64
            f("mapq").restype = C.c_uint32
        else:
            f("tables_build").restype = C.c_void_p
            f("num_threads").restype = C.c_int

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # ---- P1 ----
    def convert_ascii(self, seq: bytes, mode: int) -> bytes:
        if self.kind != "port":
            raise NotImplementedError
        out = C.create_string_buffer(len(seq))
        self.lib.orc_convert_ascii(out, seq, C.c_int64(len(seq)), mode)
        return out.raw

    def encode_2bit(self, seq: bytes) -> np.ndarray:
        n = (len(seq) + 15) // 16
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._f("encode_2bit")(_p(out, C.c_uint32), seq, len(seq))
        return out[:n]

    def decode_2bit(self, enc: np.ndarray, length: int) -> bytes:
        out = C.create_string_buffer(max(length, 1))
        enc = np.ascontiguousarray(enc, dtype=np.uint32)
        self._f("decode_2bit")(out, _p(enc, C.c_uint32), length)
        return out.raw[:length]

    def revcomp_2bit(self, enc: np.ndarray, length: int) -> np.ndarray:
        enc = np.ascontiguousarray(enc, dtype=np.uint32)
        out = np.zeros_like(enc)
        self._f("revcomp_2bit")(_p(out, C.c_uint32), _p(enc, C.c_uint32), length)
        return out

    def revcomp_ascii(self, seq: bytes) -> bytes:
        out = C.create_string_buffer(max(len(seq), 1))
        self._f("revcomp_ascii")(out, seq, len(seq))
        return out.raw[:len(seq)]

    # ---- H1/H2 ----
    def murmur64(self, x: int) -> int:
        return self._f("murmur64")(C.c_uint64(x & 0xFFFFFFFFFFFFFFFF))

    def canonical_kmers(self, enc: np.ndarray, length: int, k: int) -> np.ndarray:
        enc = np.ascontiguousarray(enc, dtype=np.uint32)
        out = np.zeros(max(length - k + 1, 1), dtype=np.uint64)
        n = self._f("canonical_kmers")(_p(enc, C.c_uint32), length, k, _p(out, C.c_uint64))
        return out[:n]

    def minhash_batch(self, enc: np.ndarray, lens: np.ndarray, k: int, H: int):
        """enc: [n, pitch_words] u32, lens: [n] i32 -> (sigs [n,H] u64, valid [n,H] u8)"""
        enc = np.ascontiguousarray(enc, dtype=np.uint32)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = lens.shape[0]
        pitch = enc.shape[1] if enc.ndim == 2 else 0
        sigs = np.zeros((n, H), dtype=np.uint64)
        valid = np.zeros((n, H), dtype=np.uint8)
        if n:
            self._f("minhash_batch")(_p(enc, C.c_uint32), C.c_int64(pitch), _p(lens, C.c_int32), n, k, H,
                                     _p(sigs, C.c_uint64), _p(valid, C.c_uint8))
        return sigs, valid

    # ---- H3 ----
    def tables_build(self, sigs, valid, ids=None, max_results_per_map=65535, loadfactor=0.8, stable=1):
        sigs = np.ascontiguousarray(sigs, dtype=np.uint64)
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        n, H = sigs.shape
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            idp = _p(ids, C.c_uint32)
        if self.kind == "port":
            h = self.lib.orc_tables_build(_p(sigs, C.c_uint64), _p(valid, C.c_uint8), idp, C.c_int64(n), H,
                                          max_results_per_map)
        else:
            h = self.lib.ref_tables_build(_p(sigs, C.c_uint64), _p(valid, C.c_uint8), idp, C.c_int64(n), H,
                                          max_results_per_map, C.c_float(loadfactor), stable)
        return C.c_void_p(h), H

    def tables_free(self, handle):
        self._f("tables_free")(handle[0])

    def tables_query(self, handle, qsigs, qvalid):
        """-> (num_per_seq [nq] i32, offsets [nq+1] i64, values [total] u32)"""
        h, H = handle
        qsigs = np.ascontiguousarray(qsigs, dtype=np.uint64)
        qvalid = np.ascontiguousarray(qvalid, dtype=np.uint8)
        nq = qsigs.shape[0]
        num = np.zeros(nq, dtype=np.int32)
        off = np.zeros(nq + 1, dtype=np.int64)
        total = self._f("tables_query")(h, _p(qsigs, C.c_uint64), _p(qvalid, C.c_uint8), C.c_int64(nq),
                                        _p(num, C.c_int32), _p(off, C.c_int64), None)
        vals = np.zeros(max(total, 1), dtype=np.uint32)
        self._f("tables_query")(h, _p(qsigs, C.c_uint64), _p(qvalid, C.c_uint8), C.c_int64(nq),
                                _p(num, C.c_int32), _p(off, C.c_int64), _p(vals, C.c_uint32))
        return num, off, vals[:total]

    def tables_save(self, handle, path, k, loadfactor=0.8):
        """reference only: the reference's own writeToStream of every table (hash-table file format)"""
        assert self.kind == "ref"
        self.lib.ref_tables_save.restype = C.c_int
        rc = self.lib.ref_tables_save(handle[0], str(path).encode(), int(k), C.c_float(loadfactor))
        assert rc == 0, rc

    def tables_load(self, path):
        """reference only: the reference's own loadFromStream -> (handle, k, loadfactor)"""
        assert self.kind == "ref"
        self.lib.ref_tables_load.restype = C.c_void_p
        k, lf = C.c_int(0), C.c_float(0)
        h = self.lib.ref_tables_load(str(path).encode(), C.byref(k), C.byref(lf))
        assert h, "the reference could not load " + str(path)
        hh = C.c_void_p(h)
        # number of tables: stored in the handle; query functions take it from there, python needs it for shapes
        return (hh, None), k.value, lf.value

    def read_file(self, path, pitch, cap):
        """reference only: sequences of a FASTA/FASTQ file as the reference's own parser yields them
        (forEachReadInFile, readlibraryio.hpp:288-326) -> (rows [n, pitch] u8 unmodified, lengths)"""
        assert self.kind == "ref"
        rows = np.zeros((cap, pitch), dtype=np.uint8)
        lens = np.zeros(cap, dtype=np.int32)
        self.lib.ref_read_file.restype = C.c_int64
        n = self.lib.ref_read_file(str(path).encode(), _p(rows, C.c_char), C.c_int64(pitch), _p(lens, C.c_int32),
                                   C.c_int64(cap))
        assert 0 <= n <= cap, n
        return rows[:n], lens[:n]

    # ---- C1 (port only; the reference implementation is CUDA-only) ----
    def filter_by_frequency(self, values, offsets, min_hits):
        values = np.ascontiguousarray(values, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        nseg = offsets.shape[0] - 1
        outv = np.zeros(max(values.shape[0], 1), dtype=np.uint32)
        outo = np.zeros(nseg + 1, dtype=np.int64)
        total = self.lib.orc_filter_by_frequency(_p(values, C.c_uint32), _p(offsets, C.c_int64), C.c_int64(nseg),
                                                 min_hits, _p(outv, C.c_uint32), _p(outo, C.c_int64))
        return outv[:total], outo

    # ---- S2 / S3 ----
    def window_location(self, secB, secE, pos, w, ext):
        l, r, ln, sp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._f("window_location")(secB, secE, pos, w, ext, C.byref(l), C.byref(r), C.byref(ln), C.byref(sp))
        return l.value, r.value, ln.value, sp.value

    def shd(self, anchor2bit, La, cand2bit, Lc, rate=0.05):
        a = np.ascontiguousarray(anchor2bit, dtype=np.uint32)
        c = np.ascontiguousarray(cand2bit, dtype=np.uint32)
        s, sc, o = C.c_int(), C.c_int(), C.c_int()
        self._f("shd")(_p(a, C.c_uint32), La, _p(c, C.c_uint32), Lc, C.c_float(rate), C.byref(s), C.byref(sc),
                       C.byref(o))
        return s.value, sc.value, o.value

    # ---- whole seeding pass, reference direction (port only) ----
    def map_pass_refdir(self, genome: bytes, chrom_off, reads: np.ndarray, read_len, k=16, w=128, H=16,
                        min_hits=4, max_results_per_map=65535, rate=0.05, batchsize=2048):
        chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        read_len = np.ascontiguousarray(read_len, dtype=np.int32)
        n = read_len.shape[0]
        out = np.zeros(n, dtype=MAPPED_DTYPE)
        stats = np.zeros(4, dtype=np.int64)
        self.lib.orc_map_pass_refdir(genome, _p(chrom_off, C.c_int64), chrom_off.shape[0] - 1,
                                     _p(reads, C.c_char), reads.shape[1], _p(read_len, C.c_int32), C.c_int64(n),
                                     k, w, H, min_hits, max_results_per_map, C.c_float(rate), batchsize,
                                     out.ctypes.data_as(C.c_void_p), _p(stats, C.c_int64))
        return out, stats

    # ---- V2 / V3 ----
    def ssw_align(self, query: bytes, ref: bytes, mask_len: int):
        al = Alignment()
        cig = C.create_string_buffer(4096)
        self._f("ssw_align")(query, len(query), ref, len(ref), mask_len, C.byref(al), cig, 4096)
        return al.astuple(), cig.value.decode()

    def edit_distance_nw(self, q: bytes, t: bytes) -> int:
        return self._f("edit_distance_nw")(q, len(q), t, len(t))

    def verify_inputs(self, read: bytes, orientation: int, chrom: bytes, pos: int, w: int, conv: int):
        rl = len(read)
        q = C.create_string_buffer(rl + 1)
        qrc = C.create_string_buffer(rl + 1)
        ref = C.create_string_buffer(w + 1)
        wl = C.c_int()
        self.lib.orc_verify_inputs(read, rl, orientation, chrom, C.c_int64(len(chrom)), C.c_int64(pos), w, conv,
                                   q, qrc, ref, C.byref(wl))
        return q.raw[:rl], qrc.raw[:rl], ref.raw[:wl.value]

    def mapq(self, s1, s2):
        return self.lib.orc_mapq(s1, s2)


def ref_cpu_pipeline(ref: "Oracle", genome: bytes, chrom_off, reads: np.ndarray, read_len, k=16, w=128, H=16,
                     min_hits=4, max_results_per_map=65535, rate=0.05, batchsize=2048, mapper_type=0,
                     want_alignments=True):
    """The reference's CPU path assembled from its own functions (oracle/ref_shim.cpp: ref_cpu_pipeline).
    -> (mapped MAPPED_DTYPE [n], alignments ALIGN_DTYPE [n,2] or None, edit [n,2], times[4])"""
    assert ref.kind == "ref"
    chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    n = read_len.shape[0]
    out = np.zeros(n, dtype=MAPPED_DTYPE)
    sw = np.zeros((n, 2), dtype=ALIGN_DTYPE) if want_alignments else None
    ed = np.zeros((n, 2), dtype=np.int32)
    times = np.zeros(4, dtype=np.float64)
    ref.lib.ref_cpu_pipeline(genome, _p(chrom_off, C.c_int64), chrom_off.shape[0] - 1, _p(reads, C.c_char),
                             reads.shape[1], _p(read_len, C.c_int32), C.c_int64(n), k, w, H, min_hits,
                             max_results_per_map, C.c_float(rate), batchsize, mapper_type,
                             out.ctypes.data_as(C.c_void_p), sw.ctypes.data_as(C.c_void_p) if sw is not None else None,
                             ed.ctypes.data_as(C.c_void_p), _p(times, C.c_double))
    return out, sw, ed, times


# ---- 8f-1 read ingestion: restatement (TEST INFRASTRUCTURE) ------------------------------------------------
def preprocess_reads_model(seqs, first_read_id=0, carry=0):
    """ref: preprocessSequence include/chunkedreadstorageconstruction.hpp:70-95 as driven by the encoder thread
    (:273-314): a c g t -> upper case; any other character that is not A C G T -> "ACGT"[Ncount], Ncount =
    (Ncount + 1) % 4; Ncount = 0 at the start of every batch of 65536 reads (fileParserMaxBatchsize :107).
    seqs: list of bytes.  Returns (list of bytes, ambiguous flags, Ncount after the last read)."""
    out, amb = [], []
    ncount = carry & 3
    for i, s in enumerate(seqs):
        if (first_read_id + i) % 65536 == 0:
            ncount = 0
        b = bytearray(s)
        bad = False
        for j, c in enumerate(b):
            if c in b"ACGT":
                continue
            if c in b"acgt":
                b[j] = c - 32
            else:
                b[j] = b"ACGT"[ncount]
                ncount = (ncount + 1) % 4
                bad = True
        out.append(bytes(b))
        amb.append(bad)
    return out, amb, ncount


def parse_records_model(text):
    """4-line FASTQ records / FASTA records whose sequence spans any number of lines up to the next header ->
    list of sequences (what kseqpp yields for such files)"""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if not lines:
        return []
    if lines[0][:1] == b"@":
        assert len(lines) % 4 == 0
        return [lines[i + 1].rstrip(b"\r") for i in range(0, len(lines), 4)]
    out = []
    for ln in lines:
        if ln[:1] == b">":
            out.append(b"")
        else:
            out[-1] += ln.rstrip(b"\r")
    return out


# ---- V4 + O1: recalculation, alignment choice, MAPQ, SAM text (TEST INFRASTRUCTURE) --------------------------
REF_SAM_SO = os.path.join(HERE, "_ref", "libhrm_ref_sam.so")
SAM_FIELDS_DTYPE = np.dtype([("sw_score", "<i4", 2), ("sw_score_next_best", "<i4", 2), ("num_conversions", "<i4", 2),
                             ("chosen", "<i4"), ("flag", "<i4"), ("mapq", "<i4"), ("window_length", "<i4"),
                             ("pos", "<i8")])
_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def have_ref_sam():
    return os.path.exists(REF_SAM_SO)


def _names_array(names):
    return (C.c_char_p * len(names))(*[n if isinstance(n, bytes) else n.encode() for n in names])


def port_sam_format(port, genomes, chrom_off, names, reads, read_len, mapped, passes, verify_conv, w=128,
                    first_read_id=0, with_header=True):
    """oracle/hrm_oracle.c: orc_sam_format.  genomes[p]: converted genome bytes of pass p; reads[p]: converted read
    rows (uint8 [n, pitch]) of pass p; mapped: MAPPED_DTYPE [n]; passes: int32 [n] (pass of each hit).
    -> (sam bytes, SAM_FIELDS_DTYPE [n])"""
    assert port.kind == "port"
    chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    mapped = np.ascontiguousarray(mapped)
    passes = np.ascontiguousarray(passes, dtype=np.int32)
    vc = np.ascontiguousarray(verify_conv, dtype=np.int32)
    n = read_len.shape[0]
    reads = [np.ascontiguousarray(r, dtype=np.uint8) for r in reads]
    gp = (C.c_char_p * len(genomes))(*genomes)
    rp = (C.c_void_p * len(reads))(*[r.ctypes.data for r in reads])
    fields = np.zeros(n, dtype=SAM_FIELDS_DTYPE)
    fn = port.lib.orc_sam_format
    fn.restype = C.c_int64
    args = [gp, _p(chrom_off, C.c_int64), chrom_off.shape[0] - 1, _names_array(names), rp, reads[0].shape[1],
            _p(read_len, C.c_int32), C.c_int64(n), mapped.ctypes.data_as(C.c_void_p), _p(passes, C.c_int32),
            _p(vc, C.c_int32), w, C.c_uint32(first_read_id), int(with_header)]
    size = fn(*args, None, C.c_int64(0), None)
    out = C.create_string_buffer(size)
    fn(*args, out, C.c_int64(size), fields.ctypes.data_as(C.c_void_p))
    return out.raw[:size], fields


def ref_mapping_sam(genome, chrom_off, names, reads, read_len, mapped, w=128, threads=4, tmp_prefix=None):
    """The reference's own Mappinghandler::go (SW mode) on these inputs (oracle/ref_shim_sam.cpp), single pass:
    genome / reads as the reference would be given them.  -> (sam bytes, per-read int32 [n, 8]:
    sw_score0, next_best0, sw_score1, next_best1, num_conversions0, num_conversions1, flag, flag_rc)"""
    import tempfile
    lib = C.CDLL(REF_SAM_SO)
    chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    mapped = np.ascontiguousarray(mapped)
    n = read_len.shape[0]
    per = np.zeros((n, 8), dtype=np.int32)
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "out") if tmp_prefix is None else tmp_prefix
        rc = lib.ref_mapping_sam(genome, _p(chrom_off, C.c_int64), chrom_off.shape[0] - 1, _names_array(names),
                                 _p(reads, C.c_char), reads.shape[1], _p(read_len, C.c_int32), C.c_int64(n),
                                 mapped.ctypes.data_as(C.c_void_p), w, threads, prefix.encode(),
                                 per.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("ref_mapping_sam failed: %d" % rc)
        with open(prefix + ".SAM", "rb") as f:
            sam = f.read()
    return sam, per


def complement_ascii(b):
    return bytes(b).translate(_COMP)


def sam_complement_text_columns(sam):
    """complements the RNEXT (window) and SEQ columns of the record lines of a SAM text (the G->A mirror of
    oracle/hrm_oracle.c: orc_sam_record)"""
    out = []
    for ln in sam.split(b"\n"):
        if ln and not ln.startswith(b"@"):
            c = ln.split(b"\t")
            c[6] = c[6].translate(_COMP)
            c[9] = c[9].translate(_COMP)
            ln = b"\t".join(c)
        out.append(ln)
    return b"\n".join(out)


# ---- C1: the reference's own CUDA candidate filter (needs a GPU; TEST INFRASTRUCTURE) ------------------------------
REF_C1_SO = os.path.join(HERE, "_ref", "libhrm_ref_c1.so")


def have_ref_c1():
    return os.path.exists(REF_C1_SO)


def ref_unique_by_count(values, offsets, min_count):
    """GpuSegmentedUniqueByCount::unique (include/gpu/cuda_unique_by_count.cuh:33-215) on the current CUDA device.
    values uint32 [total], offsets int32 [nseg + 1] -> (survivors of all segments back to back, lengths [nseg])"""
    lib = C.CDLL(REF_C1_SO)
    values = np.ascontiguousarray(values, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    nseg = offsets.shape[0] - 1
    out = np.zeros(max(values.shape[0], 1), dtype=np.uint32)
    lens = np.zeros(max(nseg, 1), dtype=np.int32)
    rc = lib.ref_unique_by_count(_p(values, C.c_uint32), int(values.shape[0]), int(nseg), _p(offsets, C.c_int32),
                                 int(min_count), _p(out, C.c_uint32), _p(lens, C.c_int32))
    if rc != 0:
        raise RuntimeError("ref_unique_by_count failed: %d" % rc)
    lens = lens[:nseg]
    return out[:int(lens.sum())], lens


# ---- the reference's preprocessed-reads dump (TEST INFRASTRUCTURE) -------------------------------------------------
def ref_readstorage_save(path, reads, read_len, ambig_ids):
    """ChunkedReadStorage::saveToFile (include/chunkedreadstorage.hpp:246-400) for these ASCII rows"""
    lib = C.CDLL(REF_SAM_SO)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    ambig_ids = np.ascontiguousarray(ambig_ids, dtype=np.uint32)
    rc = lib.ref_readstorage_save(str(path).encode(), _p(reads, C.c_char), reads.shape[1], _p(read_len, C.c_int32),
                                  C.c_int64(read_len.shape[0]), _p(ambig_ids, C.c_uint32), C.c_int64(ambig_ids.shape[0]))
    if rc != 0:
        raise RuntimeError("ref_readstorage_save failed")


def ref_readstorage_load(path, pitch_ints, cap):
    """ChunkedReadStorage::loadFromFile (:160-243) -> (2-bit rows uint32 [n, pitch], lengths, ambiguous ids)"""
    lib = C.CDLL(REF_SAM_SO)
    lib.ref_readstorage_load.restype = C.c_int64
    rows = np.zeros((cap, pitch_ints), dtype=np.uint32)
    lens = np.zeros(cap, dtype=np.int32)
    amb = np.zeros(cap, dtype=np.uint32)
    na = C.c_int64(0)
    n = lib.ref_readstorage_load(str(path).encode(), _p(rows, C.c_uint32), C.c_int64(pitch_ints), _p(lens, C.c_int32),
                                 C.c_int64(cap), _p(amb, C.c_uint32), C.byref(na))
    if n < 0:
        raise RuntimeError("ref_readstorage_load failed: %d" % n)
    return rows[:n], lens[:n], np.sort(amb[:na.value])
