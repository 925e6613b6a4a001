"""Host-side mirror of the reference's handle API over the C ABI (plumbing only: torch supplies
device memory and streams; every computation happens in libhrm_b200.so).

Names follow the reference:
  GpuMinhasher      include/gpu/gpuminhasher.cuh:20-110       -> class Minhasher
  MinhasherHandle   include/minhasherhandle.hpp:13-30         -> int handle ids
  GpuReadStorage    include/gpu/gpureadstorage.cuh:22-119     -> class ReadStorage
  Genome            include/genome.hpp:84-446                 -> class Genome
  WindowBatchProcessor + Mappinghandler (main_gpu.cu:431-856, mappinghandler.cu) -> class Mapper
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._lib import check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _dev():
    if not torch.cuda.is_available():
        raise L.HrmError(L.HRM_ERR_CUDA, "no CUDA device: the hashreadmapper hot path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ---- K1 ----------------------------------------------------------------------------------------
def encode_2bit(ascii_rows: torch.Tensor, lengths: torch.Tensor, conversion=L.CONV_NONE, pitch_words=None):
    """ascii_rows: [n, pitch] uint8 (pitch % 16 == 0) on the device; returns [n, pitch_words] int32 bit patterns"""
    lib = L.load()
    n, pitch = ascii_rows.shape
    pw = pitch_words or (pitch + 15) // 16
    out = torch.empty((n, pw), dtype=torch.int32, device=ascii_rows.device)
    check(lib.hrm_encode_2bit(_ptr(ascii_rows), pitch, _ptr(lengths), n, conversion, _ptr(out), pw, _stream()))
    return out


def encode_2bit_contiguous(ascii: torch.Tensor, conversion=L.CONV_NONE):
    lib = L.load()
    n = ascii.numel()
    out = torch.empty(((n + 15) // 16,), dtype=torch.int32, device=ascii.device)
    check(lib.hrm_encode_2bit_contiguous(_ptr(ascii), n, conversion, _ptr(out), _stream()))
    return out


# ---- K2 ----------------------------------------------------------------------------------------
def minhash(seq2bit: torch.Tensor, lengths: torch.Tensor, k=16, H=16):
    lib = L.load()
    n, pw = seq2bit.shape
    sigs = torch.empty((n, H), dtype=torch.int64, device=seq2bit.device)
    valid = torch.empty((n, H), dtype=torch.uint8, device=seq2bit.device)
    check(lib.hrm_minhash(_ptr(seq2bit), pw, _ptr(lengths), n, k, H, _ptr(sigs), _ptr(valid), _stream()))
    return sigs, valid


def minhash_windows(chrom2bit_ptr, chrom_len, k, w, H, first_window, n_windows, device=None):
    lib = L.load()
    device = device or _dev()
    sigs = torch.empty((n_windows, H), dtype=torch.int64, device=device)
    valid = torch.empty((n_windows, H), dtype=torch.uint8, device=device)
    p = chrom2bit_ptr if isinstance(chrom2bit_ptr, (int, C.c_void_p)) else _ptr(chrom2bit_ptr)
    check(lib.hrm_minhash_windows(p, chrom_len, k, w, H, first_window, n_windows, _ptr(sigs), _ptr(valid), _stream()))
    return sigs, valid


# ---- K3 ----------------------------------------------------------------------------------------
class Minhasher:
    """Mirror of GpuMinhasher (include/gpu/gpuminhasher.cuh:20-110)."""

    def __init__(self, max_sequences, max_results_per_map=65535, k=16, load_factor=0.8):
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        check(self.lib.hrm_minhasher_create(C.byref(self.h), max_sequences, max_results_per_map, k, load_factor))

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.hrm_minhasher_destroy(self.h)
            self.h = None  # (module globals may be gone at interpreter exit)

    def addHashTables(self, n, hash_function_ids=None):
        ids = None
        if hash_function_ids is not None:
            ids = np.ascontiguousarray(hash_function_ids, dtype=np.int32)
        return self.lib.hrm_minhasher_add_tables(self.h, n, _ptr(ids), _stream())

    def insert(self, seq2bit, lengths, ids=None, first_id=0, first_hash_func=0, num_hash_funcs=None):
        info = self.getInfo()
        nf = num_hash_funcs if num_hash_funcs is not None else info.num_tables - first_hash_func
        n, pw = seq2bit.shape
        check(self.lib.hrm_minhasher_insert(self.h, _ptr(seq2bit), pw, _ptr(lengths), n, _ptr(ids), first_id,
                                            first_hash_func, nf, _stream()))

    def insertSignatures(self, sigs, valid, ids=None, first_id=0):
        check(self.lib.hrm_minhasher_insert_signatures(self.h, _ptr(sigs), _ptr(valid), sigs.shape[0], _ptr(ids),
                                                       first_id, _stream()))

    def checkInsertionErrors(self, first=0, num=0):
        return self.lib.hrm_minhasher_check_insertion_errors(self.h, first, num, _stream())

    def compact(self):
        check(self.lib.hrm_minhasher_compact(self.h, _stream()))

    def constructionIsFinished(self):
        check(self.lib.hrm_minhasher_finish(self.h, _stream()))

    def makeMinhasherHandle(self):
        h = self.lib.hrm_minhasher_handle_create(self.h)
        if h < 0:
            check(h)
        return h

    def destroyHandle(self, handle):
        check(self.lib.hrm_minhasher_handle_destroy(self.h, handle))

    def determineNumValues(self, handle, seq2bit, lengths):
        n, pw = seq2bit.shape
        num = torch.empty((max(n, 1),), dtype=torch.int32, device=seq2bit.device)
        total = C.c_int64(0)
        check(self.lib.hrm_minhasher_count(self.h, handle, _ptr(seq2bit), pw, _ptr(lengths), n, _ptr(num),
                                           C.byref(total), _stream()))
        return num[:n], total.value

    def determineNumValuesFromSignatures(self, handle, sigs):
        n = sigs.shape[0]
        num = torch.empty((max(n, 1),), dtype=torch.int32, device=sigs.device)
        total = C.c_int64(0)
        check(self.lib.hrm_minhasher_count_signatures(self.h, handle, _ptr(sigs), None, n, _ptr(num), C.byref(total),
                                                      _stream()))
        return num[:n], total.value

    def retrieveValues(self, handle, n, total, num_per_seq):
        dev = num_per_seq.device
        values = torch.empty((max(total, 1),), dtype=torch.int32, device=dev)
        offsets = torch.empty((n + 1,), dtype=torch.int32, device=dev)
        check(self.lib.hrm_minhasher_retrieve(self.h, handle, n, total, _ptr(values), _ptr(num_per_seq), _ptr(offsets),
                                              _stream()))
        return values[:total], offsets

    def getInfo(self):
        info = L.MinhasherInfo()
        check(self.lib.hrm_minhasher_info(self.h, C.byref(info)))
        return info

    def getNumberOfMaps(self):
        return self.getInfo().num_tables

    def getKmerSize(self):
        return self.getInfo().k

    def getNumResultsPerMapThreshold(self):
        return self.getInfo().max_results_per_map

    def hasGpuTables(self):
        return True

    def writeToBytes(self):
        size = C.c_int64(0)
        check(self.lib.hrm_minhasher_serialize(self.h, None, C.byref(size)))
        buf = np.empty(size.value, dtype=np.uint8)
        check(self.lib.hrm_minhasher_serialize(self.h, _ptr(buf), C.byref(size)))
        return buf

    @classmethod
    def loadFromBytes(cls, buf):
        self = cls.__new__(cls)
        self.lib = L.load()
        self.h = C.c_void_p()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        check(self.lib.hrm_minhasher_deserialize(C.byref(self.h), _ptr(buf), buf.size))
        return self


    def writeToStream(self, path):
        """the reference's own file format (ref: FakeGpuMinhasher::writeToStream fakegpuminhasher.cuh:498-510)"""
        size = C.c_int64(0)
        check(self.lib.hrm_minhasher_write_reference_format(self.h, None, C.byref(size)))
        buf = np.empty(size.value, dtype=np.uint8)
        check(self.lib.hrm_minhasher_write_reference_format(self.h, _ptr(buf), C.byref(size)))
        buf[:size.value].tofile(path)
        return int(size.value)

    @classmethod
    def loadFromStream(cls, path, num_maps_upper_limit=-1):
        """ref: FakeGpuMinhasher::loadFromStream fakegpuminhasher.cuh:512-532; returns the minhasher"""
        self = cls.__new__(cls)
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        buf = np.fromfile(path, dtype=np.uint8)
        check(self.lib.hrm_minhasher_read_reference_format(C.byref(self.h), _ptr(buf), buf.size, num_maps_upper_limit))
        return self


# ---- read ingestion --------------------------------------------------------------------------------
def ingest_reads(text: torch.Tensor, pitch, max_reads, first_read_id=0, carry_replaced=0):
    """FASTQ / FASTA text (uint8 tensor on the device, whole records) -> (rows [n, pitch] u8, lengths [n] i32,
    ambiguous [n] u8, carry for the next chunk).  ref: readlibraryio.hpp:288-326 +
    chunkedreadstorageconstruction.hpp:70-95"""
    lib = L.load()
    rows = torch.zeros((max_reads, pitch), dtype=torch.uint8, device=text.device)
    lens = torch.zeros((max_reads,), dtype=torch.int32, device=text.device)
    amb = torch.zeros((max_reads,), dtype=torch.uint8, device=text.device)
    n, carry = C.c_int64(0), C.c_int32(0)
    check(lib.hrm_ingest_reads(_ptr(text), text.numel(), first_read_id, carry_replaced, _ptr(rows), pitch, _ptr(lens),
                               _ptr(amb), max_reads, C.byref(n), C.byref(carry), _stream()))
    return rows[:n.value], lens[:n.value], amb[:n.value], carry.value


# ---- K4 ----------------------------------------------------------------------------------------
def filter_by_frequency(values, num_per_seq, offsets, min_hits):
    """in place; returns new total (values[:total] valid)"""
    lib = L.load()
    n = offsets.numel() - 1
    total = C.c_int64(0)
    check(lib.hrm_filter_by_frequency(_ptr(values), _ptr(num_per_seq), _ptr(offsets), n, min_hits, C.byref(total),
                                      _stream()))
    return total.value


def segment_ids(offsets, total):
    lib = L.load()
    n = offsets.numel() - 1
    out = torch.empty((max(total, 1),), dtype=torch.int32, device=offsets.device)
    check(lib.hrm_segment_ids(_ptr(offsets), n, total, _ptr(out), _stream()))
    return out[:total]


# ---- S1 ----------------------------------------------------------------------------------------
class ReadStorage:
    """Mirror of GpuReadStorage (include/gpu/gpureadstorage.cuh:22-119)."""

    @classmethod
    def from2Bit(cls, seq2bit: torch.Tensor, lengths: torch.Tensor, ambiguous: torch.Tensor = None):
        """adopts packed device rows (+ the ambiguity flags hrm_ingest_reads produced)"""
        self = cls.__new__(cls)
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        n, pw = seq2bit.shape
        check(self.lib.hrm_readstore_create_from_2bit(C.byref(self.h), _ptr(seq2bit), pw, _ptr(lengths), n, _stream()))
        if ambiguous is not None:
            check(self.lib.hrm_readstore_set_ambiguous(self.h, _ptr(ambiguous), _stream()))
        return self

    def __init__(self, ascii_rows: np.ndarray, lengths: np.ndarray, conversion=L.CONV_NONE):
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        ascii_rows = np.ascontiguousarray(ascii_rows, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        check(self.lib.hrm_readstore_create_from_ascii(C.byref(self.h), _ptr(ascii_rows), ascii_rows.shape[1],
                                                       _ptr(lengths), ascii_rows.shape[0], conversion, _stream()))

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.hrm_readstore_destroy(self.h)
            self.h = None  # (module globals may be gone at interpreter exit)

    def makeHandle(self):
        return self.lib.hrm_readstore_handle_create(self.h)

    def destroyHandle(self, handle):
        check(self.lib.hrm_readstore_handle_destroy(self.h, handle))

    def getInfo(self):
        info = L.ReadstoreInfo()
        check(self.lib.hrm_readstore_info(self.h, C.byref(info)))
        return info

    def getNumberOfReads(self):
        return self.getInfo().num_reads

    def getSequenceLengthUpperBound(self):
        return self.getInfo().length_upper_bound

    def getSequenceLengthLowerBound(self):
        return self.getInfo().length_lower_bound

    def gatherSequences(self, handle, ids: torch.Tensor, out_pitch_words=None):
        pw = out_pitch_words or self.getInfo().pitch_words
        n = ids.numel()
        out = torch.empty((n, pw), dtype=torch.int32, device=ids.device)
        check(self.lib.hrm_readstore_gather(self.h, handle, _ptr(out), pw, _ptr(ids), n, _stream()))
        return out

    def gatherContiguousSequences(self, handle, first_id, n, out_pitch_words=None):
        pw = out_pitch_words or self.getInfo().pitch_words
        out = torch.empty((n, pw), dtype=torch.int32, device=_dev())
        check(self.lib.hrm_readstore_gather_contiguous(self.h, handle, _ptr(out), pw, first_id, n, _stream()))
        return out

    def gatherSequenceLengths(self, handle, ids: torch.Tensor):
        n = ids.numel()
        out = torch.empty((n,), dtype=torch.int32, device=ids.device)
        check(self.lib.hrm_readstore_gather_lengths(self.h, handle, _ptr(out), _ptr(ids), n, _stream()))
        return out

    def saveToBytes(self) -> bytes:
        """the reference's preprocessed-reads dump (ChunkedReadStorage::saveToFile)"""
        size = C.c_int64(0)
        check(self.lib.hrm_readstore_write_reference_format(self.h, None, C.byref(size)))
        buf = np.empty(size.value, dtype=np.uint8)
        check(self.lib.hrm_readstore_write_reference_format(self.h, _ptr(buf), C.byref(size)))
        return buf[:size.value].tobytes()

    @classmethod
    def loadFromBytes(cls, data: bytes):
        """from a dump the reference (or saveToBytes) wrote"""
        self = cls.__new__(cls)
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        buf = np.frombuffer(data, dtype=np.uint8)
        check(self.lib.hrm_readstore_read_reference_format(C.byref(self.h), _ptr(buf), buf.size, _stream()))
        return self

    def areSequencesAmbiguous(self, handle, ids: torch.Tensor):
        n = ids.numel()
        out = torch.empty((n,), dtype=torch.uint8, device=ids.device)
        check(self.lib.hrm_readstore_are_ambiguous(self.h, handle, _ptr(out), _ptr(ids), n, _stream()))
        return out

    def getNumberOfReadsWithN(self):
        return self.getInfo().num_reads_with_n

    def getIdsOfAmbiguousReads(self):
        out = np.zeros(max(int(self.getNumberOfReadsWithN()), 1), dtype=np.uint32)
        check(self.lib.hrm_readstore_ambiguous_ids(self.h, _ptr(out)))
        return out[:int(self.getNumberOfReadsWithN())]


# ---- S2 ----------------------------------------------------------------------------------------
class Genome:
    """Mirror of Genome (include/genome.hpp:84-446), device resident and 2-bit packed."""

    def __init__(self, ascii: bytes, chrom_offsets, conversion=L.CONV_NONE):
        self.lib = L.load()
        _dev()
        self.h = C.c_void_p()
        self._off = np.ascontiguousarray(chrom_offsets, dtype=np.int64)
        self._buf = np.frombuffer(ascii, dtype=np.uint8)
        check(self.lib.hrm_genome_create_from_ascii(C.byref(self.h), _ptr(self._buf), _ptr(self._off),
                                                    len(self._off) - 1, conversion, _stream()))

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.hrm_genome_destroy(self.h)
            self.h = None  # (module globals may be gone at interpreter exit)

    def numChromosomes(self):
        return self.lib.hrm_genome_num_chromosomes(self.h)

    def chromosomeLength(self, c):
        return self.lib.hrm_genome_chromosome_length(self.h, c)

    def getNumWindowsInChromosome(self, c, k, w):
        return self.lib.hrm_genome_num_windows_in_chromosome(self.h, c, k, w)

    def getTotalNumWindows(self, k, w):
        return self.lib.hrm_genome_num_windows(self.h, k, w)

    def chromosome2BitPtr(self, c):
        return C.c_void_p(self.lib.hrm_genome_chromosome_2bit(self.h, c))

    def windowInfo(self, k, w, gw):
        c, wid, pos, ln = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32()
        check(self.lib.hrm_genome_window_info(self.h, k, w, gw, C.byref(c), C.byref(wid), C.byref(pos), C.byref(ln)))
        return c.value, wid.value, pos.value, ln.value

    def extendedWindows(self, chrom, w, window_pos: torch.Tensor, read_len: torch.Tensor, out_pitch_words):
        n = window_pos.numel()
        dev = window_pos.device
        out = torch.empty((n, out_pitch_words), dtype=torch.int32, device=dev)
        left = torch.empty((n,), dtype=torch.int32, device=dev)
        right = torch.empty((n,), dtype=torch.int32, device=dev)
        ln = torch.empty((n,), dtype=torch.int32, device=dev)
        check(self.lib.hrm_extended_windows(self.h, chrom, w, _ptr(window_pos), _ptr(read_len), n, _ptr(out),
                                            out_pitch_words, _ptr(left), _ptr(right), _ptr(ln), _stream()))
        return out, left, right, ln


# ---- S3 ----------------------------------------------------------------------------------------
def shifted_hamming(anchors, anchor_len, cands, cand_len, max_error_rate=0.05):
    lib = L.load()
    n = anchors.shape[0]
    dev = anchors.device
    shift = torch.empty((n,), dtype=torch.int32, device=dev)
    score = torch.empty((n,), dtype=torch.int32, device=dev)
    orient = torch.empty((n,), dtype=torch.int8, device=dev)
    check(lib.hrm_shifted_hamming(_ptr(anchors), anchors.shape[1], _ptr(anchor_len), _ptr(cands), cands.shape[1],
                                  _ptr(cand_len), n, max_error_rate, _ptr(shift), _ptr(score), _ptr(orient), _stream()))
    return shift, score, orient


# ---- V2 / V3 -----------------------------------------------------------------------------------
def sw_align(queries, query_len, refs, ref_len, mask_len, cigar_pitch=128):
    """queries [n, qp] uint8, refs [n, rp] uint8 on the device -> (alignments np structured, cigars list[str])"""
    lib = L.load()
    n = queries.shape[0]
    dev = queries.device
    out = torch.zeros((n, 10), dtype=torch.int32, device=dev)
    cig = torch.zeros((n, cigar_pitch), dtype=torch.uint8, device=dev)
    check(lib.hrm_sw_align(_ptr(queries), queries.shape[1], _ptr(query_len), _ptr(refs), refs.shape[1], _ptr(ref_len),
                           _ptr(mask_len), n, _ptr(out), _ptr(cig), cigar_pitch, _stream()))
    torch.cuda.synchronize()
    al = out.cpu().numpy().view(L.ALIGN_DTYPE).reshape(n)
    cg = cig.cpu().numpy()
    cigs = [bytes(cg[i, :min(al["cigar_len"][i], cigar_pitch)]).decode() for i in range(n)]
    return al, cigs


def edit_distance(queries, query_len, targets, target_len):
    lib = L.load()
    n = queries.shape[0]
    out = torch.empty((n,), dtype=torch.int32, device=queries.device)
    check(lib.hrm_edit_distance(_ptr(queries), queries.shape[1], _ptr(query_len), _ptr(targets), targets.shape[1],
                                _ptr(target_len), n, _ptr(out), _stream()))
    return out


def inflate_gzip(data: bytes, cap=None) -> bytes:
    """gzip'd FASTQ / FASTA bytes -> text (host side, zlib)"""
    lib = L.load()
    src = np.frombuffer(data, dtype=np.uint8)
    cap = cap or max(1 << 16, 8 * src.size)
    while True:
        out = np.empty(cap, dtype=np.uint8)
        written = C.c_int64(0)
        st = lib.hrm_inflate_gzip(_ptr(src), src.size, _ptr(out), cap, C.byref(written))
        if st == L.HRM_ERR_OVERFLOW:
            cap *= 4
            continue
        check(st)
        return out[:written.value].tobytes()


# ---- the fused mapper --------------------------------------------------------------------------
def directional_config(**kw):
    """Directional bisulfite library: C->T reads against the C->T index (forward strand hits) and
    against the G->A index (reverse strand hits)."""
    cfg = default_config(**kw)
    cfg.num_passes = 2
    cfg.read_conversion[0], cfg.genome_conversion[0], cfg.verify_conversion[0] = L.CONV_CT, L.CONV_CT, L.CONV_CT
    cfg.read_conversion[1], cfg.genome_conversion[1], cfg.verify_conversion[1] = L.CONV_CT, L.CONV_GA, L.CONV_GA
    return cfg


def nondirectional_config(**kw):
    cfg = directional_config(**kw)
    cfg.num_passes = 4
    cfg.read_conversion[2], cfg.genome_conversion[2], cfg.verify_conversion[2] = L.CONV_GA, L.CONV_GA, L.CONV_GA
    cfg.read_conversion[3], cfg.genome_conversion[3], cfg.verify_conversion[3] = L.CONV_GA, L.CONV_CT, L.CONV_CT
    return cfg


def default_config(**kw):
    cfg = L.MapperConfig()
    L.load().hrm_mapper_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


class Comm:
    """NCCL communicator of the key-partitioned index (hrm_comm_*).  The unique id is created on rank 0 and
    broadcast with torch.distributed (any backend) when a process group exists."""

    def __init__(self, rank=None, world=None, group=None):
        import torch.distributed as dist
        self.lib = L.load()
        _dev()
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        ident = np.zeros(L.COMM_ID_BYTES, dtype=np.uint8)
        if rank == 0:
            check(self.lib.hrm_comm_unique_id(_ptr(ident), ident.size))
        if world > 1:
            t = torch.from_numpy(ident)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = t.cpu().numpy()
        self.h = C.c_void_p()
        check(self.lib.hrm_comm_create(C.byref(self.h), rank, world, _ptr(ident)))
        self.rank, self.world = rank, world

    def info(self):
        info = L.CommInfo()
        check(self.lib.hrm_comm_info(self.h, C.byref(info)))
        return info

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.hrm_comm_destroy(self.h)
            self.h = None  # (module globals may be gone at interpreter exit)


def key_owner(key, world):
    return L.load().hrm_key_owner(C.c_uint64(int(key)), int(world))


class Mapper:
    def __init__(self, cfg=None):
        self.lib = L.load()
        _dev()
        self.cfg = cfg or default_config()
        self.h = C.c_void_p()
        check(self.lib.hrm_mapper_create(C.byref(self.h), C.byref(self.cfg)))
        self.chrom_names = None

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.hrm_mapper_destroy(self.h)
            self.h = None  # (module globals may be gone at interpreter exit)

    def setGenome(self, ascii: bytes, chrom_offsets, chrom_names=None):
        off = np.ascontiguousarray(chrom_offsets, dtype=np.int64)
        buf = np.frombuffer(ascii, dtype=np.uint8)
        check(self.lib.hrm_mapper_set_genome(self.h, _ptr(buf), _ptr(off), len(off) - 1, _stream()))
        self.chrom_names = chrom_names or ["chr%d" % i for i in range(len(off) - 1)]

    def info(self):
        info = L.MapperInfo()
        check(self.lib.hrm_mapper_info(self.h, C.byref(info)))
        return info

    STAGES = ["pack", "minhash", "probe", "scan", "retrieve", "filter", "shd", "merge", "verify", "route"]

    def setPartition(self, comm):
        """key-partitioned index over the ranks of `comm` (before setGenome); mapBatch becomes collective"""
        self.comm = comm
        check(self.lib.hrm_mapper_set_partition(self.h, comm.h if comm is not None else None))

    def setProfiling(self, enable=True):
        check(self.lib.hrm_mapper_set_profiling(self.h, int(enable)))

    def stageTimes(self):
        """-> {stage: (ms summed since the last call, number of spans)}"""
        ms = (C.c_float * len(self.STAGES))()
        sp = (C.c_int32 * len(self.STAGES))()
        check(self.lib.hrm_mapper_stage_times(self.h, ms, sp))
        return {n: (float(ms[i]), int(sp[i])) for i, n in enumerate(self.STAGES)}

    def mapBatch(self, reads_ascii: torch.Tensor, lengths: torch.Tensor, want_stats=True):
        """device-resident reads -> device tensor of hrm_mapped_read (as [n, 8] int32) + stats"""
        n, pitch = reads_ascii.shape
        out = torch.empty((n, 8), dtype=torch.int32, device=reads_ascii.device)
        st = L.BatchStats()
        check(self.lib.hrm_map_batch(self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, _ptr(out),
                                     C.byref(st) if want_stats else None, _stream()))
        return out, st

    def verifyBatch(self, reads_ascii, lengths, mapped, cigar_pitch=64, want_stats=False):
        n, pitch = reads_ascii.shape
        rec = torch.empty((n, C.sizeof(L.ReadRecord) // 4), dtype=torch.int32, device=reads_ascii.device)
        cig = torch.empty((2 * n, cigar_pitch), dtype=torch.uint8, device=reads_ascii.device)
        st = L.BatchStats()
        check(self.lib.hrm_verify_batch(self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, _ptr(mapped), _ptr(rec),
                                        _ptr(cig), cigar_pitch, C.byref(st) if want_stats else None, _stream()))
        return rec, cig, st

    def mapReads(self, reads_ascii: np.ndarray, lengths: np.ndarray, cigar_pitch=64, records=None, cigars=None):
        """host buffers in, host records out (the end-to-end call)"""
        n, pitch = reads_ascii.shape
        if records is None:
            records = np.empty(n, dtype=L.RECORD_DTYPE)
        if cigars is None:
            cigars = np.empty((2 * n, cigar_pitch), dtype=np.uint8)
        st = L.BatchStats()
        check(self.lib.hrm_mapper_map_reads(self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, _ptr(records),
                                            _ptr(cigars), cigar_pitch, C.byref(st), _stream()))
        return records, cigars, st

    def _names(self):
        return (C.c_char_p * len(self.chrom_names))(*[s.encode() for s in self.chrom_names])

    def samFields(self, reads_ascii, lengths, records, cigars):
        """V4 on the device (recalculated scores, conversion counts, chosen alignment, FLAG, MAPQ, POS):
        device tensors in -> numpy SAM_FIELDS_DTYPE [n]"""
        n, pitch = reads_ascii.shape
        out = torch.empty((n, L.SAM_FIELDS_DTYPE.itemsize // 4), dtype=torch.int32, device=reads_ascii.device)
        check(self.lib.hrm_sam_fields_batch(self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, _ptr(records),
                                            _ptr(cigars), cigars.shape[1], _ptr(out), _stream()))
        return out.cpu().numpy().view(L.SAM_FIELDS_DTYPE).reshape(n)

    def samFormatDevice(self, reads_ascii, lengths, records, cigars, part, first_read_id=0):
        """text of one part (L.SAM_SQ_LINES / L.SAM_RECORDS) as a device uint8 tensor"""
        n, pitch = reads_ascii.shape
        written = C.c_int64(0)
        args = (self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, _ptr(records), _ptr(cigars), cigars.shape[1],
                first_read_id, self._names(), part)
        check(self.lib.hrm_sam_format_device(*args, None, 0, C.byref(written), _stream()))
        out = torch.empty((max(written.value, 1),), dtype=torch.uint8, device=reads_ascii.device)
        check(self.lib.hrm_sam_format_device(*args, _ptr(out), written.value, C.byref(written), _stream()))
        return out[:written.value]

    def mapReadsSam(self, reads_ascii: np.ndarray, lengths: np.ndarray, first_read_id=0, cigar_pitch=128,
                    rec_out=None, sq_out=None, want_records=False):
        """host reads in, SAM text out (the @SQ lines and the record lines of these reads), end to end on the device.
        -> (sq bytes view, record bytes view, stats[, records, cigars])"""
        n, pitch = reads_ascii.shape
        bound = n * (96 + cigar_pitch + self.cfg.window_size + pitch)
        if rec_out is None:
            rec_out = np.empty(bound, dtype=np.uint8)
        if sq_out is None:
            sq_out = np.empty(n * 40 + 16, dtype=np.uint8)
        records = np.empty(n, dtype=L.RECORD_DTYPE) if want_records else None
        cigars = np.empty((2 * n, cigar_pitch), dtype=np.uint8) if want_records else None
        st = L.BatchStats()
        sqw, recw = C.c_int64(0), C.c_int64(0)
        check(self.lib.hrm_mapper_map_reads_sam(self.h, _ptr(reads_ascii), pitch, _ptr(lengths), n, first_read_id,
                                                self._names(), _ptr(sq_out), sq_out.size, C.byref(sqw), _ptr(rec_out),
                                                rec_out.size, C.byref(recw), _ptr(records) if want_records else None,
                                                _ptr(cigars) if want_records else None, cigar_pitch, C.byref(st),
                                                _stream()))
        if want_records:
            return sq_out[:sqw.value], rec_out[:recw.value], st, records, cigars
        return sq_out[:sqw.value], rec_out[:recw.value], st

    # ---- double-buffered pipeline (host buffers should be pinned) ----
    def stageReads(self, slot, reads_ascii: np.ndarray, lengths: np.ndarray):
        n, pitch = reads_ascii.shape
        check(self.lib.hrm_mapper_stage_reads(self.h, slot, _ptr(reads_ascii), pitch, _ptr(lengths), n))

    def stageDevice(self, slot, reads_ascii: torch.Tensor, lengths: torch.Tensor, max_length=None):
        """reads already on the device (kept alive by the caller until finish(slot))"""
        n, pitch = reads_ascii.shape
        check(self.lib.hrm_mapper_stage_device(self.h, slot, _ptr(reads_ascii), pitch, _ptr(lengths), n,
                                               int(max_length if max_length is not None else pitch), _stream()))

    def stageFastq(self, slot, text: np.ndarray, pitch, max_reads, first_read_id=0, carry=0):
        """FASTQ / FASTA text (uint8, pinned) -> staged batch; returns (number of reads, carry for the next chunk)"""
        nr, co = C.c_int64(0), C.c_int32(0)
        check(self.lib.hrm_mapper_stage_fastq(self.h, slot, _ptr(text), text.size, first_read_id, carry, pitch, max_reads,
                                              C.byref(nr), C.byref(co)))
        return nr.value, co.value

    def mapStaged(self, slot, records=None, cigars=None, cigar_pitch=64, first_read_id=0, sq_out=None, rec_out=None,
                  want_stats=False):
        st = L.BatchStats()
        check(self.lib.hrm_mapper_map_staged(self.h, slot, _ptr(records), _ptr(cigars), cigar_pitch, first_read_id,
                                             self._names(), _ptr(sq_out), sq_out.size if sq_out is not None else 0,
                                             _ptr(rec_out), rec_out.size if rec_out is not None else 0,
                                             C.byref(st) if want_stats else None, _stream()))
        return st

    def finish(self, slot):
        """-> (@SQ text bytes, record text bytes) now in the host buffers given to mapStaged"""
        sqw, recw = C.c_int64(0), C.c_int64(0)
        check(self.lib.hrm_mapper_finish(self.h, slot, C.byref(sqw), C.byref(recw)))
        return sqw.value, recw.value

    def samFormat(self, records, cigars, reads_ascii, lengths, first_read_id=0, with_header=True):
        n = len(records)
        names = self._names()
        written = C.c_int64(0)
        check(self.lib.hrm_sam_format(self.h, _ptr(records), _ptr(cigars), cigars.shape[1], _ptr(reads_ascii),
                                      reads_ascii.shape[1], _ptr(lengths), n, first_read_id, names, int(with_header),
                                      None, 0, C.byref(written)))
        buf = np.empty(written.value, dtype=np.uint8)
        check(self.lib.hrm_sam_format(self.h, _ptr(records), _ptr(cigars), cigars.shape[1], _ptr(reads_ascii),
                                      reads_ascii.shape[1], _ptr(lengths), n, first_read_id, names, int(with_header),
                                      _ptr(buf), buf.size, C.byref(written)))
        return buf.tobytes()
