"""ctypes binding of libhrm_b200.so (the C ABI declared in include/hrm_b200.h).

There is no fallback of any kind: if the CUDA library cannot be loaded the import raises, and if
no CUDA device is usable every compute entry point returns HRM_ERR_CUDA, which `check` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhrm_b200.so")

HRM_OK = 0
HRM_ERR_CUDA, HRM_ERR_INVALID, HRM_ERR_NOMEM, HRM_ERR_STATE, HRM_ERR_OVERFLOW = -1, -2, -3, -4, -5
CONV_NONE, CONV_CT, CONV_GA = 0, 1, 2
ORIENT_FORWARD, ORIENT_REVCOMP, ORIENT_NONE = 1, 2, 3
MAPPER_SW, MAPPER_EDLIB = 0, 1
MAX_PASSES = 4


class HrmError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("libhrm_b200: status %d: %s" % (status, text))
        self.status = status


class MinhasherInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("num_tables", C.c_int32), ("max_results_per_map", C.c_int32),
                ("load_factor", C.c_float), ("num_inserted", C.c_int64), ("num_keys_total", C.c_int64),
                ("num_values_total", C.c_int64), ("device_bytes", C.c_int64), ("is_compacted", C.c_int32),
                ("has_gpu_tables", C.c_int32)]


class ReadstoreInfo(C.Structure):
    _fields_ = [("num_reads", C.c_int64), ("length_lower_bound", C.c_int32), ("length_upper_bound", C.c_int32),
                ("num_reads_with_n", C.c_int64), ("pitch_words", C.c_int32), ("is_paired_end", C.c_int32),
                ("device_bytes", C.c_int64)]


class MappedRead(C.Structure):
    _fields_ = [("orientation", C.c_int32), ("hamming_distance", C.c_int32), ("shift", C.c_int32),
                ("chromosome_id", C.c_int32), ("position", C.c_int64), ("pass_", C.c_int32), ("reserved", C.c_int32)]


class Alignment(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("sw_score", "sw_score_next_best", "ref_begin", "ref_end", "query_begin",
                                         "query_end", "ref_end_next_best", "mismatches", "flag", "cigar_len")]


class ReadRecord(C.Structure):
    _fields_ = [("mapped", MappedRead), ("alignments", Alignment * 2), ("edit_distance", C.c_int32 * 2),
                ("window_length", C.c_int32), ("mask_len", C.c_int32)]


class MapperConfig(C.Structure):
    _fields_ = [("k", C.c_int32), ("window_size", C.c_int32), ("num_tables", C.c_int32), ("min_table_hits", C.c_int32),
                ("max_results_per_map", C.c_int32), ("load_factor", C.c_float), ("max_hamming_percent", C.c_float),
                ("mapper_type", C.c_int32), ("num_passes", C.c_int32), ("read_conversion", C.c_int32 * MAX_PASSES),
                ("genome_conversion", C.c_int32 * MAX_PASSES), ("verify_conversion", C.c_int32 * MAX_PASSES)]


class MapperInfo(C.Structure):
    _fields_ = [("num_windows", C.c_int64), ("index_device_bytes", C.c_int64), ("genome_device_bytes", C.c_int64),
                ("num_keys_total", C.c_int64), ("table_slots_total", C.c_int64), ("num_passes", C.c_int32),
                ("reserved", C.c_int32), ("collect_ids_counted", C.c_int64), ("collect_ids_skipped", C.c_int64),
                ("collect_reads_block_kernel", C.c_int64)]


class CommInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("bytes_sent", C.c_int64), ("bytes_received", C.c_int64),
                ("exchanges", C.c_int64)]


COMM_ID_BYTES = 128


class BatchStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("num_reads", "num_probes", "num_slot_touches", "num_values",
                                         "num_candidates", "num_mapped", "num_kernel_launches")]


import numpy as np  # noqa: E402

MAPPED_DTYPE = np.dtype([("orientation", "<i4"), ("hamming_distance", "<i4"), ("shift", "<i4"),
                         ("chromosome_id", "<i4"), ("position", "<i8"), ("pass", "<i4"), ("reserved", "<i4")])
ALIGN_DTYPE = np.dtype([(n, "<i4") for n, _ in Alignment._fields_])
RECORD_DTYPE = np.dtype([("mapped", MAPPED_DTYPE), ("alignments", ALIGN_DTYPE, (2,)), ("edit_distance", "<i4", (2,)),
                         ("window_length", "<i4"), ("mask_len", "<i4")])
SAM_FIELDS_DTYPE = np.dtype([("sw_score", "<i4", (2,)), ("sw_score_next_best", "<i4", (2,)),
                             ("num_conversions", "<i4", (2,)), ("chosen", "<i4"), ("flag", "<i4"), ("mapq", "<i4"),
                             ("window_length", "<i4"), ("pos", "<i8")])
SAM_HD = b"@HD\tVN:1.4\n"
SAM_PG_CO = b"@PG\tHashreadmapper\tID:1.0@CO: QNAME\tFLAG\tRNAME\tPOS\tMAPQ\tCIGAR\tRNEXT\tPNEXT\tTLEN\tSEQ\tQUAL\tTAG\n"
SAM_SQ_LINES, SAM_RECORDS = 0, 1
assert SAM_FIELDS_DTYPE.itemsize == 48
assert MAPPED_DTYPE.itemsize == C.sizeof(MappedRead) == 32
assert RECORD_DTYPE.itemsize == C.sizeof(ReadRecord)

# every symbol include/hrm_b200.h declares: name -> (restype, argtypes)
P, I32, I64, U32, F32, VP = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_void_p
SIGNATURES = {
    "hrm_last_error": (C.c_char_p, []),
    "hrm_abi_version": (C.c_int, []),
    "hrm_device_count": (C.c_int, []),
    "hrm_encode_2bit": (I32, [P, I64, P, I64, C.c_int, P, I64, VP]),
    "hrm_encode_2bit_contiguous": (I32, [P, I64, C.c_int, P, VP]),
    "hrm_minhash": (I32, [P, I64, P, I64, C.c_int, C.c_int, P, P, VP]),
    "hrm_minhash_windows": (I32, [P, I64, C.c_int, C.c_int, C.c_int, I64, I64, P, P, VP]),
    "hrm_minhasher_create": (I32, [C.POINTER(P), I64, C.c_int, C.c_int, F32]),
    "hrm_minhasher_destroy": (None, [P]),
    "hrm_minhasher_add_tables": (C.c_int, [P, C.c_int, P, VP]),
    "hrm_minhasher_insert": (I32, [P, P, I64, P, I64, P, U32, C.c_int, C.c_int, VP]),
    "hrm_minhasher_insert_signatures": (I32, [P, P, P, I64, P, U32, VP]),
    "hrm_minhasher_check_insertion_errors": (C.c_int, [P, C.c_int, C.c_int, VP]),
    "hrm_minhasher_compact": (I32, [P, VP]),
    "hrm_minhasher_finish": (I32, [P, VP]),
    "hrm_minhasher_handle_create": (C.c_int, [P]),
    "hrm_minhasher_handle_destroy": (I32, [P, C.c_int]),
    "hrm_minhasher_count": (I32, [P, C.c_int, P, I64, P, C.c_int, P, C.POINTER(I64), VP]),
    "hrm_minhasher_count_signatures": (I32, [P, C.c_int, P, P, C.c_int, P, C.POINTER(I64), VP]),
    "hrm_minhasher_retrieve": (I32, [P, C.c_int, C.c_int, I64, P, P, P, VP]),
    "hrm_minhasher_info": (I32, [P, C.POINTER(MinhasherInfo)]),
    "hrm_minhasher_serialize": (I32, [P, P, C.POINTER(I64)]),
    "hrm_minhasher_deserialize": (I32, [C.POINTER(P), P, I64]),
    "hrm_minhasher_write_reference_format": (I32, [P, P, C.POINTER(I64)]),
    "hrm_minhasher_read_reference_format": (I32, [C.POINTER(P), P, I64, C.c_int]),
    "hrm_filter_by_frequency": (I32, [P, P, P, C.c_int, C.c_int, C.POINTER(I64), VP]),
    "hrm_segment_ids": (I32, [P, C.c_int, I64, P, VP]),
    "hrm_readstore_create_from_ascii": (I32, [C.POINTER(P), P, I64, P, I64, C.c_int, VP]),
    "hrm_readstore_create_from_2bit": (I32, [C.POINTER(P), P, I64, P, I64, VP]),
    "hrm_readstore_destroy": (None, [P]),
    "hrm_readstore_handle_create": (C.c_int, [P]),
    "hrm_readstore_handle_destroy": (I32, [P, C.c_int]),
    "hrm_readstore_gather": (I32, [P, C.c_int, P, I64, P, I64, VP]),
    "hrm_readstore_gather_contiguous": (I32, [P, C.c_int, P, I64, U32, I64, VP]),
    "hrm_readstore_gather_lengths": (I32, [P, C.c_int, P, P, I64, VP]),
    "hrm_readstore_are_ambiguous": (I32, [P, C.c_int, P, P, I64, VP]),
    "hrm_readstore_ambiguous_ids": (I32, [P, P]),
    "hrm_readstore_set_ambiguous": (I32, [P, P, VP]),
    "hrm_readstore_info": (I32, [P, C.POINTER(ReadstoreInfo)]),
    "hrm_readstore_write_reference_format": (I32, [P, P, C.POINTER(I64)]),
    "hrm_readstore_read_reference_format": (I32, [C.POINTER(P), P, I64, VP]),
    "hrm_genome_create_from_ascii": (I32, [C.POINTER(P), P, P, C.c_int, C.c_int, VP]),
    "hrm_genome_destroy": (None, [P]),
    "hrm_genome_num_chromosomes": (C.c_int, [P]),
    "hrm_genome_chromosome_length": (I64, [P, C.c_int]),
    "hrm_genome_num_windows_in_chromosome": (I64, [P, C.c_int, C.c_int, C.c_int]),
    "hrm_genome_num_windows": (I64, [P, C.c_int, C.c_int]),
    "hrm_genome_chromosome_2bit": (P, [P, C.c_int]),
    "hrm_genome_window_info": (I32, [P, C.c_int, C.c_int, I64, C.POINTER(I32), C.POINTER(I64), C.POINTER(I64),
                                     C.POINTER(I32)]),
    "hrm_extended_windows": (I32, [P, C.c_int, C.c_int, P, P, I64, P, I64, P, P, P, VP]),
    "hrm_shifted_hamming": (I32, [P, I64, P, P, I64, P, I64, F32, P, P, P, VP]),
    "hrm_sw_align": (I32, [P, I64, P, P, I64, P, P, I64, P, P, I64, VP]),
    "hrm_edit_distance": (I32, [P, I64, P, P, I64, P, I64, P, VP]),
    "hrm_mapper_default_config": (None, [C.POINTER(MapperConfig)]),
    "hrm_mapper_create": (I32, [C.POINTER(P), C.POINTER(MapperConfig)]),
    "hrm_mapper_destroy": (None, [P]),
    "hrm_mapper_set_genome": (I32, [P, P, P, C.c_int, VP]),
    "hrm_mapper_info": (I32, [P, C.POINTER(MapperInfo)]),
    "hrm_map_batch": (I32, [P, P, I64, P, I64, P, C.POINTER(BatchStats), VP]),
    "hrm_verify_batch": (I32, [P, P, I64, P, I64, P, P, P, I64, C.POINTER(BatchStats), VP]),
    "hrm_mapper_map_reads": (I32, [P, P, I64, P, I64, P, P, I64, C.POINTER(BatchStats), VP]),
    "hrm_mapper_set_profiling": (I32, [P, C.c_int]),
    "hrm_mapper_stage_times": (I32, [P, C.POINTER(C.c_float), C.POINTER(I32)]),
    "hrm_ingest_reads": (I32, [P, I64, I64, I32, P, I64, P, P, I64, C.POINTER(I64), C.POINTER(I32), VP]),
    "hrm_inflate_gzip": (I32, [P, I64, P, I64, C.POINTER(I64)]),
    "hrm_comm_unique_id": (I32, [P, I64]),
    "hrm_comm_create": (I32, [C.POINTER(P), C.c_int, C.c_int, P]),
    "hrm_comm_destroy": (None, [P]),
    "hrm_comm_info": (I32, [P, C.POINTER(CommInfo)]),
    "hrm_key_owner": (C.c_int, [C.c_uint64, C.c_int]),
    "hrm_minhasher_set_partition": (I32, [P, C.c_int, C.c_int]),
    "hrm_mapper_set_partition": (I32, [P, P]),
    "hrm_sam_format": (I32, [P, P, P, I64, P, I64, P, I64, U32, P, C.c_int, P, I64, C.POINTER(I64)]),
    "hrm_sam_fields_batch": (I32, [P, P, I64, P, I64, P, P, I64, P, VP]),
    "hrm_sam_format_device": (I32, [P, P, I64, P, I64, P, P, I64, U32, P, C.c_int, P, I64, C.POINTER(I64), VP]),
    "hrm_mapper_stage_reads": (I32, [P, C.c_int, P, I64, P, I64]),
    "hrm_mapper_stage_device": (I32, [P, C.c_int, P, I64, P, I64, C.c_int, VP]),
    "hrm_mapper_stage_fastq": (I32, [P, C.c_int, P, I64, I64, I32, I64, I64, C.POINTER(I64), C.POINTER(I32)]),
    "hrm_mapper_map_staged": (I32, [P, C.c_int, P, P, I64, U32, P, P, I64, P, I64, C.POINTER(BatchStats), VP]),
    "hrm_mapper_finish": (I32, [P, C.c_int, C.POINTER(I64), C.POINTER(I64)]),
    "hrm_mapper_map_reads_sam": (I32, [P, P, I64, P, I64, U32, P, P, I64, C.POINTER(I64), P, I64, C.POINTER(I64), P, P,
                                       I64, C.POINTER(BatchStats), VP]),
}

_lib = None


def load():
    """Loads the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libhrm_b200.so is not built (%s). Run `python -m hashreadmapper_b200.build` -- there is no "
            "CPU or PyTorch fallback for the hashreadmapper hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != HRM_OK:
        raise HrmError(status, load().hrm_last_error().decode("utf-8", "replace"))
