/*
 * hrm_b200.h -- C ABI of libhrm_b200.so: the B200-native (sm_100a) read-mapping hot path of
 * hashreadmapper, behind the reference's own handle API.
 *
 * Every entry point takes plain pointers and sizes (no C++/torch types).  `d_` pointers are
 * caller-allocated DEVICE memory, `h_` pointers are HOST memory, `hrm_stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  All functions return an
 * hrm_status; the C++ adaptor (hashreadmapper_b200/csrc/hrm_adaptor.hpp) re-throws it the way
 * the reference's CUDACHECK does (include/gpu/cudaerrorcheck.cuh:42-58).  Calls are
 * asynchronous with respect to the given stream unless stated otherwise.  There is NO CPU
 * fallback anywhere: without a CUDA device every compute call returns HRM_ERR_CUDA.
 *
 * "ref:" citations are file:line in the reference repository (clubby93421234/hashreadmapper)
 * and name the interface each entry point replaces.  See INTEGRATION.md for the adaptor.
 */
#ifndef HRM_B200_H
#define HRM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t hrm_status;
#define HRM_OK 0
#define HRM_ERR_CUDA (-1)     /* a CUDA runtime call failed (ref: CUDACHECK -> std::runtime_error) */
#define HRM_ERR_INVALID (-2)  /* bad argument (ref: assert in the callee) */
#define HRM_ERR_NOMEM (-3)    /* device allocation failed (ref: gpuminhasherconstruction.cu:71,226) */
#define HRM_ERR_STATE (-4)    /* call order violated (ref: stage assert fakegpuminhasher.cuh:328) */
#define HRM_ERR_OVERFLOW (-5) /* a 32-bit count of the reference's API would overflow */

typedef void* hrm_stream;

/* 3N conversion applied while packing (north_star item 1).  HRM_CONV_CT is the reference's
 * NucleoideConverer (ref: src/gpu/mappinghandler.cu:163-179); HRM_CONV_GA is its mirror for the
 * reverse-strand index. */
#define HRM_CONV_NONE 0
#define HRM_CONV_CT 1
#define HRM_CONV_GA 2

/* ref: include/alignmentorientation.hpp:4 */
#define HRM_ORIENT_FORWARD 1
#define HRM_ORIENT_REVCOMP 2
#define HRM_ORIENT_NONE 3

const char* hrm_last_error(void);   /* thread-local text of the last failure */
int hrm_abi_version(void);          /* HRM_ABI_VERSION this library was built with */
#define HRM_ABI_VERSION 2
int hrm_device_count(void);         /* number of visible CUDA devices (0 => nothing can run) */

/* ------------------------------------------------------------------------------------------
 * K1 -- 3N conversion + 2-bit packing.
 * ref: SequenceHelpers::encodeSequence2Bit include/sequencehelpers.hpp:185-218 and
 *      callEncodeSequencesTo2BitKernel src/gpu/sequenceconversionkernels.cu:448-517 (call site
 *      src/gpu/main_gpu.cu:522).  A=0 C=1 G=2 T=3, anything else 0; 16 bases per word, MSB first,
 *      last word left-aligned.  Conversion (none / C->T / G->A) happens on the ASCII byte first.
 * d_ascii: n rows of ascii_pitch bytes (ascii_pitch % 16 == 0, rows 16-byte aligned);
 * d_out:   n rows of out_pitch_words words; words beyond ceil(len/16) are written as 0.
 * ---------------------------------------------------------------------------------------- */
hrm_status hrm_encode_2bit(const char* d_ascii, int64_t ascii_pitch, const int32_t* d_lengths,
                           int64_t n, int conversion, uint32_t* d_out, int64_t out_pitch_words,
                           hrm_stream stream);

/* One long sequence (a chromosome).  d_ascii need not be aligned.  d_out gets ceil(len/16) words. */
hrm_status hrm_encode_2bit_contiguous(const char* d_ascii, int64_t len, int conversion,
                                      uint32_t* d_out, hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * K2 -- minhash signatures.
 * ref: GPUSequenceHasher::hash include/gpu/gpusequencehasher.cuh:720-776,
 *      minhashSignatures3264Kernel :116-169, MurmurHash<u64> include/helpers/hashers.cuh:129-137.
 * sig[i][j] = (min over canonical k-mers c of murmur64(c + j)) & (2^(2k)-1); valid[i][j] = len>=k,
 * otherwise sig = ~0.  Hash-function ids are 0..H-1 (ref: gpuminhasherconstruction.cu:128-133).
 * d_sigs: [n][H] u64; d_valid: [n][H] u8 (may be NULL).  1 <= k <= 32, 1 <= H <= 64.
 * ---------------------------------------------------------------------------------------- */
hrm_status hrm_minhash(const uint32_t* d_seq2bit, int64_t pitch_words, const int32_t* d_lengths,
                       int64_t n, int k, int H, uint64_t* d_sigs, uint8_t* d_valid,
                       hrm_stream stream);

/* Same, for the windows of one packed chromosome (ref: Genome::forEachWindowInChromosome
 * include/genome.hpp:176-209: window i starts at i*(w-k+1), length min(w, len - start)).
 * Sketches windows [first_window, first_window + n_windows). */
hrm_status hrm_minhash_windows(const uint32_t* d_chrom2bit, int64_t chrom_len, int k, int w, int H,
                               int64_t first_window, int64_t n_windows, uint64_t* d_sigs,
                               uint8_t* d_valid, hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * K3 -- the minhasher: H multi-value hash tables, signature -> ids.
 * ref: class GpuMinhasher include/gpu/gpuminhasher.cuh:20-110 (abstract interface),
 *      MinhasherHandle include/minhasherhandle.hpp:13-30,
 *      FakeGpuMinhasher include/gpu/fakegpuminhasher.cuh (the default, CPU-table backend whose
 *      results this reproduces), constructGpuMinhasherFromReadStorage
 *      src/gpu/gpuminhasherconstruction.cu:36-252 (the construction loop that calls these).
 * Direction-agnostic: the reference inserts reads and queries windows; hrm_mapper_* inserts
 * windows and queries reads.
 * ---------------------------------------------------------------------------------------- */
typedef struct hrm_minhasher hrm_minhasher;

typedef struct {
    int32_t k;
    int32_t num_tables;
    int32_t max_results_per_map;
    float load_factor;
    int64_t num_inserted;      /* sequences inserted so far */
    int64_t num_keys_total;    /* distinct keys over all tables (after compact) */
    int64_t num_values_total;  /* stored values over all tables (after compact) */
    int64_t device_bytes;      /* ref: GpuMinhasher::getMemoryInfo() */
    int32_t is_compacted;
    int32_t has_gpu_tables;    /* always 1 (ref: hasGpuTables()) */
} hrm_minhasher_info_t;

/* ref: FakeGpuMinhasher(maxNumKeys, maxValuesPerKey, k, loadfactor) fakegpuminhasher.cuh:150-153.
 * max_sequences bounds the number of sequences that will be inserted. */
hrm_status hrm_minhasher_create(hrm_minhasher** out, int64_t max_sequences, int max_results_per_map,
                                int k, float load_factor);
void hrm_minhasher_destroy(hrm_minhasher* mh);

/* ref: GpuMinhasher::addHashTables(int n, const int* hashFunctionIds, cudaStream_t) -> number added
 * (gpuminhasher.cuh:43; fewer than n signals memory shortage).  Ids must continue 0,1,2,... */
int hrm_minhasher_add_tables(hrm_minhasher* mh, int n, const int32_t* h_hash_function_ids,
                             hrm_stream stream);

/* ref: GpuMinhasher::insert (gpuminhasher.cuh:45-57; impl fakegpuminhasher.cuh:568-728).
 * Hashes n packed sequences with hash functions [first_hash_func, first_hash_func+num_hash_funcs)
 * and appends (signature, id) to those tables; sequences shorter than k are skipped.
 * d_ids may be NULL (ids = first_id .. first_id+n-1). */
hrm_status hrm_minhasher_insert(hrm_minhasher* mh, const uint32_t* d_seq2bit, int64_t pitch_words,
                                const int32_t* d_lengths, int64_t n, const uint32_t* d_ids,
                                uint32_t first_id, int first_hash_func, int num_hash_funcs,
                                hrm_stream stream);

/* Same, with signatures already computed by hrm_minhash / hrm_minhash_windows (all tables). */
hrm_status hrm_minhasher_insert_signatures(hrm_minhasher* mh, const uint64_t* d_sigs,
                                           const uint8_t* d_valid, int64_t n, const uint32_t* d_ids,
                                           uint32_t first_id, hrm_stream stream);

/* ref: GpuMinhasher::checkInsertionErrors (gpuminhasher.cuh:59) -> number of failed insertions (0) */
int hrm_minhasher_check_insertion_errors(hrm_minhasher* mh, int first_hash_func, int num_hash_funcs,
                                         hrm_stream stream);

/* ref: GpuMinhasher::compact (gpuminhasher.cuh:65; fakegpuminhasher.cuh:394-442 + groupbykey.hpp):
 * per table sort pairs by key (stable), unique keys, keep the first
 * min(max_results_per_map, 65535) values of each key, build the open-addressing key table.
 * Synchronises the stream (it sizes the tables from the distinct-key counts). */
hrm_status hrm_minhasher_compact(hrm_minhasher* mh, hrm_stream stream);
/* ref: GpuMinhasher::constructionIsFinished (gpuminhasher.cuh:67) -- frees build scratch */
hrm_status hrm_minhasher_finish(hrm_minhasher* mh, hrm_stream stream);

/* ref: makeMinhasherHandle / destroyHandle (gpuminhasher.cuh:35-37).  A handle owns per-caller
 * query scratch; one per host thread.  Returns handle id >= 0, or a negative hrm_status. */
int hrm_minhasher_handle_create(hrm_minhasher* mh);
hrm_status hrm_minhasher_handle_destroy(hrm_minhasher* mh, int handle);

/* ref: GpuMinhasher::determineNumValues (gpuminhasher.cuh:69-79; fakegpuminhasher.cuh:199-310).
 * d_num_per_seq[i] = sum over tables of the bucket size of sig_j(seq i), buckets larger than
 * max_results_per_map dropped.  *h_total receives the sum (the call synchronises the stream, as
 * the reference does at fakegpuminhasher.cuh:260). */
hrm_status hrm_minhasher_count(hrm_minhasher* mh, int handle, const uint32_t* d_seq2bit,
                               int64_t pitch_words, const int32_t* d_lengths, int n,
                               int32_t* d_num_per_seq, int64_t* h_total, hrm_stream stream);
/* Same, from precomputed signatures. */
hrm_status hrm_minhasher_count_signatures(hrm_minhasher* mh, int handle, const uint64_t* d_sigs,
                                          const uint8_t* d_valid, int n, int32_t* d_num_per_seq,
                                          int64_t* h_total, hrm_stream stream);

/* ref: GpuMinhasher::retrieveValues (gpuminhasher.cuh:81-90; fakegpuminhasher.cuh:312-392).
 * Must follow hrm_minhasher_count* on the same handle (else HRM_ERR_STATE).  d_offsets gets n+1
 * entries (exclusive prefix sum of d_num_per_seq); d_values gets, per sequence, the buckets of
 * tables 0..H-1 concatenated, each bucket ascending by insertion order. */
hrm_status hrm_minhasher_retrieve(hrm_minhasher* mh, int handle, int n, int64_t total,
                                  uint32_t* d_values, const int32_t* d_num_per_seq,
                                  int32_t* d_offsets, hrm_stream stream);

hrm_status hrm_minhasher_info(const hrm_minhasher* mh, hrm_minhasher_info_t* out);

/* ref: GpuMinhasher::writeToStream / loadFromStream (gpuminhasher.cuh:98-104).  The byte format is
 * this library's own; *h_size receives the size; pass h_buf = NULL to query it. */
hrm_status hrm_minhasher_serialize(const hrm_minhasher* mh, void* h_buf, int64_t* h_size);
hrm_status hrm_minhasher_deserialize(hrm_minhasher** out, const void* h_buf, int64_t size);

/* The reference's own on-disk format (`--save-hashtables-to` / `--load-hashtables-from`):
 * ref: FakeGpuMinhasher::writeToStream / loadFromStream include/gpu/fakegpuminhasher.cuh:498-532,
 *      CpuReadOnlyMultiValueHashTable::writeToStream / loadFromStream include/cpuhashtable.hpp:624-646,
 *      AoSCpuSingleValueHashTable::writeToStream / loadFromStream :217-244.
 * write: the byte stream the reference would save for tables with these contents (every key re-inserted into
 * the reference's linear-probing layout on the host); h_buf == NULL returns the size in *h_size.
 * read: builds device tables from such a stream; max_tables >= 0 limits the number of tables loaded
 * (ref: numMapsUpperLimit). */
hrm_status hrm_minhasher_write_reference_format(const hrm_minhasher* mh, void* h_buf, int64_t* h_size);
hrm_status hrm_minhasher_read_reference_format(hrm_minhasher** out, const void* h_buf, int64_t size,
                                               int max_tables);

/* ------------------------------------------------------------------------------------------
 * K4 -- candidate collection: per segment sort ascending, run-length count, keep ids whose
 * multiplicity >= min_hits (min_hits <= 1: plain distinct).
 * ref: GpuMinhashQueryFilter::keepDistinctByFrequency include/gpu/minhashqueryfilter.cuh:239-278
 *      -> GpuSegmentedUniqueByCount::unique include/gpu/cuda_unique_by_count.cuh:33-215;
 *      keepDistinct minhashqueryfilter.cuh:217-236 (call site src/gpu/main_gpu.cu:233-254).
 * In place: d_values is compacted, d_num_per_seq / d_offsets (n+1) rewritten, *h_total = new total
 * (synchronises the stream like the reference, main_gpu.cu:266-275).
 * ---------------------------------------------------------------------------------------- */
hrm_status hrm_filter_by_frequency(uint32_t* d_values, int32_t* d_num_per_seq, int32_t* d_offsets,
                                   int n, int min_hits, int64_t* h_total, hrm_stream stream);

/* ref: getSegmentIdsPerElement src/gpu/main_gpu.cu:289-324: d_segment_ids[e] = segment of element e */
hrm_status hrm_segment_ids(const int32_t* d_offsets, int n, int64_t total, int32_t* d_segment_ids,
                           hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * S1 -- read storage.
 * ref: class GpuReadStorage include/gpu/gpureadstorage.cuh:22-119, ReadStorageHandle
 *      include/readstoragehandle.hpp:13-30 (impl include/gpu/multigpureadstorage.cuh:655-905).
 * ---------------------------------------------------------------------------------------- */
typedef struct hrm_readstore hrm_readstore;

typedef struct {
    int64_t num_reads;
    int32_t length_lower_bound;
    int32_t length_upper_bound;
    int64_t num_reads_with_n;   /* reads in which a non-ACGT character was replaced (always by A) */
    int32_t pitch_words;
    int32_t is_paired_end;
    int64_t device_bytes;
} hrm_readstore_info_t;

/* Packs n ASCII reads (host rows of ascii_pitch bytes) into device 2-bit rows with `conversion`
 * applied (K1 runs on the device; the host only copies). */
hrm_status hrm_readstore_create_from_ascii(hrm_readstore** out, const char* h_ascii,
                                           int64_t ascii_pitch, const int32_t* h_lengths, int64_t n,
                                           int conversion, hrm_stream stream);
/* Adopts already packed device rows (copied). */
hrm_status hrm_readstore_create_from_2bit(hrm_readstore** out, const uint32_t* d_seq2bit,
                                          int64_t pitch_words, const int32_t* d_lengths, int64_t n,
                                          hrm_stream stream);
void hrm_readstore_destroy(hrm_readstore* rs);
int hrm_readstore_handle_create(hrm_readstore* rs);
hrm_status hrm_readstore_handle_destroy(hrm_readstore* rs, int handle);
/* ref: gatherSequences (gpureadstorage.cuh:49-58) / gatherContiguousSequences (:60-68) /
 *      gatherSequenceLengths (:94-100) */
hrm_status hrm_readstore_gather(const hrm_readstore* rs, int handle, uint32_t* d_out,
                                int64_t out_pitch_words, const uint32_t* d_ids, int64_t n,
                                hrm_stream stream);
hrm_status hrm_readstore_gather_contiguous(const hrm_readstore* rs, int handle, uint32_t* d_out,
                                           int64_t out_pitch_words, uint32_t first_id, int64_t n,
                                           hrm_stream stream);
hrm_status hrm_readstore_gather_lengths(const hrm_readstore* rs, int handle, int32_t* d_lengths,
                                        const uint32_t* d_ids, int64_t n, hrm_stream stream);
/* ref: areSequencesAmbiguous (gpureadstorage.cuh:31-37): d_result[e] = 1 if read d_ids[e] held a non-ACGT character */
hrm_status hrm_readstore_are_ambiguous(const hrm_readstore* rs, int handle, uint8_t* d_result,
                                       const uint32_t* d_ids, int64_t n, hrm_stream stream);
/* ref: getIdsOfAmbiguousReads (gpureadstorage.cuh:89-91): ascending ids into h_ids (num_reads_with_n entries) */
hrm_status hrm_readstore_ambiguous_ids(const hrm_readstore* rs, uint32_t* h_ids);
/* carries hrm_ingest_reads' d_ambiguous flags (one byte per read, device) into a store created from 2-bit rows */
hrm_status hrm_readstore_set_ambiguous(hrm_readstore* rs, const uint8_t* d_flags, hrm_stream stream);
hrm_status hrm_readstore_info(const hrm_readstore* rs, hrm_readstore_info_t* out);
/* The reference's preprocessed-reads dump (`--save-preprocessedreads-to` / `--load-preprocessedreads-from`):
 * ref: ChunkedReadStorage::saveToFile / loadFromFile include/chunkedreadstorage.hpp:160-400, LengthStore
 *      include/lengthstorage.hpp:164-204.  write: the bytes the reference would save for a storage with these
 * reads (no quality scores); h_buf == NULL returns the size in *h_size.  read: builds a device read store from
 * such a dump.  Dumps are interchangeable with the reference's. */
hrm_status hrm_readstore_write_reference_format(const hrm_readstore* rs, void* h_buf, int64_t* h_size);
hrm_status hrm_readstore_read_reference_format(hrm_readstore** out, const void* h_buf, int64_t size,
                                               hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * S2 -- genome and reference windows.
 * ref: struct Genome include/genome.hpp:84-446 (window enumeration :176-209, :304-354),
 *      ReferenceWindows include/referencewindows.hpp:13-91.
 * The genome lives on the device 2-bit packed, one word-aligned run per chromosome.
 * ---------------------------------------------------------------------------------------- */
typedef struct hrm_genome hrm_genome;

/* h_ascii: chromosomes concatenated (already upper-case, ref: genome.hpp str_toupper);
 * h_chrom_offsets: n_chrom+1 byte offsets.  Each chromosome must be shorter than 2^31
 * (ref: positions are int, genome.hpp:92). */
hrm_status hrm_genome_create_from_ascii(hrm_genome** out, const char* h_ascii,
                                        const int64_t* h_chrom_offsets, int n_chrom, int conversion,
                                        hrm_stream stream);
void hrm_genome_destroy(hrm_genome* g);
int hrm_genome_num_chromosomes(const hrm_genome* g);
int64_t hrm_genome_chromosome_length(const hrm_genome* g, int chrom);
/* ref: Genome::getNumWindowsInChromosome / getTotalNumWindows genome.hpp:176-196 */
int64_t hrm_genome_num_windows_in_chromosome(const hrm_genome* g, int chrom, int k, int w);
int64_t hrm_genome_num_windows(const hrm_genome* g, int k, int w);
/* device pointer to the packed words of one chromosome (for tests / custom kernels) */
const uint32_t* hrm_genome_chromosome_2bit(const hrm_genome* g, int chrom);
/* ref: ReferenceWindows (referencewindows.hpp:31-62): global window id -> (chromosome, window id,
 * start position, length) on the host */
hrm_status hrm_genome_window_info(const hrm_genome* g, int k, int w, int64_t global_window_id,
                                  int32_t* chrom, int64_t* window_id, int64_t* position,
                                  int32_t* length);

/* ref: detail::computeWindowLocation include/gpu/windowgenerationkernels.cuh:17-48 +
 *      generateExtendedWindows2BitKernel :162-262 (call site main_gpu.cu:691).  For each candidate e
 * (window d_window_pos[e] of chromosome `chrom`, read length d_read_len[e]) writes the extended
 * window [pos-left, pos-left+len) packed to d_out[e] and left/right/len. */
hrm_status hrm_extended_windows(const hrm_genome* g, int chrom, int w, const int32_t* d_window_pos,
                                const int32_t* d_read_len, int64_t n, uint32_t* d_out,
                                int64_t out_pitch_words, int32_t* d_ext_left, int32_t* d_ext_right,
                                int32_t* d_ext_len, hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * S3 -- shifted Hamming distance with full overlap.
 * ref: computeShiftedHammingDistancesFullOverlap -> callShiftedHammingDistanceWithFullOverlapKernelSmem1
 *      src/gpu/hammingdistancekernels.cu:266-353 (kernel :132-263; call site main_gpu.cu:714).
 * Candidate e: anchor (extended window) row e vs candidate (read) row e; all shifts
 * 0..La-Lc, read forward then reverse-complemented; result = first minimum.  orientation None when
 * min > int(float(Lc)*max_error_rate) or Lc > La; shift/score are then unspecified
 * except Lc > La which yields (0, Lc).
 * ---------------------------------------------------------------------------------------- */
hrm_status hrm_shifted_hamming(const uint32_t* d_anchor2bit, int64_t anchor_pitch_words,
                               const int32_t* d_anchor_len, const uint32_t* d_cand2bit,
                               int64_t cand_pitch_words, const int32_t* d_cand_len, int64_t n,
                               float max_error_rate, int32_t* d_best_shift, int32_t* d_best_score,
                               int8_t* d_best_orientation, hrm_stream stream);

/* ref: struct MappedRead include/gpu/mappedread.cuh:6-12 (+ the pass that produced it) */
typedef struct {
    int32_t orientation;      /* HRM_ORIENT_* */
    int32_t hamming_distance;
    int32_t shift;            /* shift - extensionLeft (main_gpu.cu:787-790) */
    int32_t chromosome_id;
    int64_t position;         /* window start in the chromosome */
    int32_t pass;             /* index of the pass (3N index) that produced the hit, -1 if none */
    int32_t reserved;
} hrm_mapped_read;

/* ------------------------------------------------------------------------------------------
 * V2 -- local alignment with CIGAR (replaces SSW).
 * ref: StripedSmithWaterman::Aligner::Align(query, ref, ref_len, filter, alignment, maskLen)
 *      src/ssw_cpp.cpp:361-400 with the default Aligner (+2/-2, gap 3/1) and default Filter;
 *      ssw_align src/ssw.c:818-922 and everything below it; call site mappinghandler.cu:556-595.
 * Bit-exact fields incl. second-best rules, byte->word switch and banded trace back.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t sw_score;
    int32_t sw_score_next_best;
    int32_t ref_begin;
    int32_t ref_end;
    int32_t query_begin;
    int32_t query_end;
    int32_t ref_end_next_best;
    int32_t mismatches;
    int32_t flag;        /* return value of Align(): 0 ok, 1 trace back failed, 2 partial */
    int32_t cigar_len;   /* bytes in the cigar string (not NUL-terminated when == cigar_pitch) */
} hrm_alignment;

#define HRM_SW_MAX_QUERY 512
#define HRM_SW_MAX_REF 512

/* n alignments; query e = d_queries + e*query_pitch (ASCII, d_query_len[e] bytes), likewise refs.
 * d_cigars: n rows of cigar_pitch bytes (>= 64 recommended; longer strings are truncated but
 * cigar_len keeps the full length). */
hrm_status hrm_sw_align(const char* d_queries, int64_t query_pitch, const int32_t* d_query_len,
                        const char* d_refs, int64_t ref_pitch, const int32_t* d_ref_len,
                        const int32_t* d_mask_len, int64_t n, hrm_alignment* d_out, char* d_cigars,
                        int64_t cigar_pitch, hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * V3 -- global edit distance (replaces edlib as the reference calls it).
 * ref: edlibAlign(q, qlen, t, tlen, edlibDefaultAlignConfig()) src/edlib.cpp:1474-1476, call site
 *      src/gpu/mappinghandler.cu:968-987: mode NW, task DISTANCE, k = -1.
 * ---------------------------------------------------------------------------------------- */
hrm_status hrm_edit_distance(const char* d_queries, int64_t query_pitch, const int32_t* d_query_len,
                             const char* d_targets, int64_t target_pitch, const int32_t* d_target_len,
                             int64_t n, int32_t* d_distance, hrm_stream stream);

/* ------------------------------------------------------------------------------------------
 * The fused mapper (north_star direction: 3N index over reference windows, reads probe it).
 * ref: WindowBatchProcessor::operator() src/gpu/main_gpu.cu:471-854 (seeding + filter + SHD + best
 *      window) and Mappinghandler::go src/gpu/mappinghandler.cu:67, CSSW :383-766, edlibAligner
 *      :841-1010 (verification), printtoSAM :196-293 (output).
 * ---------------------------------------------------------------------------------------- */
typedef struct hrm_mapper hrm_mapper;

#define HRM_MAPPER_SW 0     /* ref: --mappertype SW (default, options.hpp) */
#define HRM_MAPPER_EDLIB 1  /* ref: --mappertype edlib */
#define HRM_MAX_PASSES 4

typedef struct {
    int32_t k;                    /* ref: --kmerlength, default 16 */
    int32_t window_size;          /* ref: --windowSize, default 128 */
    int32_t num_tables;           /* ref: --hashmaps, default 16 */
    int32_t min_table_hits;       /* ref: --minTableHits, default 4 */
    int32_t max_results_per_map;  /* ref: --maxResultsPerMap, default 65535 */
    float load_factor;            /* ref: --hashtableLoadfactor, default 0.8 */
    float max_hamming_percent;    /* ref: --maxHammingPercent, default 0.05 */
    int32_t mapper_type;          /* HRM_MAPPER_SW / HRM_MAPPER_EDLIB */
    /* passes: pass p maps reads converted with read_conversion[p] against the genome converted
     * with genome_conversion[p]; verification converts with verify_conversion[p].  The reference
     * on pre-converted input is one pass.  Directional bisulfite = {CT/CT/CT, CT/GA/GA}. */
    int32_t num_passes;
    int32_t read_conversion[HRM_MAX_PASSES];
    int32_t genome_conversion[HRM_MAX_PASSES];
    int32_t verify_conversion[HRM_MAX_PASSES];
} hrm_mapper_config;

void hrm_mapper_default_config(hrm_mapper_config* cfg); /* reference defaults, one NONE pass */

/* A mapper belongs to the CUDA device that is current at hrm_mapper_create; every hrm_mapper_* entry point makes that
 * device current on the calling thread (the staging calls below are meant for a second host thread, which starts with
 * device 0 current).  The other handles follow the CUDA convention: the caller selects the device. */
hrm_status hrm_mapper_create(hrm_mapper** out, const hrm_mapper_config* cfg);
void hrm_mapper_destroy(hrm_mapper* m);

/* Packs the genome once per distinct genome_conversion, sketches every window and builds the
 * window index of each pass on the current device.  Synchronous. */
hrm_status hrm_mapper_set_genome(hrm_mapper* m, const char* h_ascii, const int64_t* h_chrom_offsets,
                                 int n_chrom, hrm_stream stream);

typedef struct {
    int64_t num_windows;
    int64_t index_device_bytes;
    int64_t genome_device_bytes;
    int64_t num_keys_total;
    int64_t table_slots_total;
    int32_t num_passes;
    int32_t reserved;
    int64_t collect_ids_counted;  /* fused collection, since creation: ids counted in shared memory */
    int64_t collect_ids_skipped;  /* ... ids of the largest buckets that were only looked up */
    int64_t collect_reads_block_kernel; /* ... (read, pass) pairs the warp kernel handed to the block kernel */
} hrm_mapper_info_t;
hrm_status hrm_mapper_info(const hrm_mapper* m, hrm_mapper_info_t* out);

/* Per-batch record of the verification stage: both alignments of the reference
 * (alignments[0] = 3N(read) vs 3N(window), alignments[1] = 3N(RC(read)) vs 3N(window)),
 * cigar strings in d_cigars rows 2e and 2e+1. */
typedef struct {
    hrm_mapped_read mapped;
    hrm_alignment alignments[2];
    int32_t edit_distance[2];   /* HRM_MAPPER_EDLIB: NW distances; else -1 */
    int32_t window_length;
    int32_t mask_len;
} hrm_read_record;

/* Counters of the last hrm_map_batch / hrm_mapper_map_reads call (for the roofline bookkeeping) */
typedef struct {
    int64_t num_reads;
    int64_t num_probes;          /* (read, table) lookups issued, over all passes */
    int64_t num_slot_touches;    /* slots examined, over all passes */
    int64_t num_values;          /* candidate values retrieved before filtering */
    int64_t num_candidates;      /* (read, window) pairs after the frequency filter */
    int64_t num_mapped;
    int64_t num_kernel_launches; /* kernels of this library launched by the call */
} hrm_batch_stats;

/* hrm_map_batch: seeding + filter + SHD + best window for n reads resident on the device as ASCII
 * rows (K1..K5 + the per-read arg-min of main_gpu.cu:777-821).  d_out: n hrm_mapped_read.
 * hrm_verify_batch: verification of those reads (K6/K7); d_records: n hrm_read_record,
 * d_cigars: 2n rows of cigar_pitch bytes.  Both synchronise internally where sizes are needed. */
hrm_status hrm_map_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                         const int32_t* d_lengths, int64_t n, hrm_mapped_read* d_out,
                         hrm_batch_stats* h_stats, hrm_stream stream);
hrm_status hrm_verify_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                            const int32_t* d_lengths, int64_t n, const hrm_mapped_read* d_mapped,
                            hrm_read_record* d_records, char* d_cigars, int64_t cigar_pitch,
                            hrm_batch_stats* h_stats, hrm_stream stream);

/* End to end with HOST buffers (what a user of the reference binary gets from STEP 1 + STEP 2):
 * copies reads H2D, runs hrm_map_batch + hrm_verify_batch, copies records and cigars D2H.
 * h_reads_ascii should be pinned for full PCIe speed.  Synchronous. */
hrm_status hrm_mapper_map_reads(hrm_mapper* m, const char* h_reads_ascii, int64_t ascii_pitch,
                                const int32_t* h_lengths, int64_t n, hrm_read_record* h_records,
                                char* h_cigars, int64_t cigar_pitch, hrm_batch_stats* h_stats,
                                hrm_stream stream);

/* Per-stage device timing of the fused path (replaces the reference's nvtx ranges + CpuTimer lines,
 * ref: include/helpers/nvtx_markers.cuh:15-58, main_gpu.cu:484-775, include/helpers/timers.cuh).
 * When enabled, CUDA events are recorded around every stage on the caller's stream (no extra
 * synchronisation); hrm_mapper_stage_times sums them per stage since the last call (it waits for
 * the recorded events) and resets the accumulation. */
#define HRM_STAGE_PACK 0      /* K1 */
#define HRM_STAGE_MINHASH 1   /* K2 */
#define HRM_STAGE_PROBE 2     /* K3b: the probe kernel alone */
#define HRM_STAGE_SCAN 3      /* offsets + the pass' one host sync */
#define HRM_STAGE_RETRIEVE 4  /* K3b: value gather */
#define HRM_STAGE_FILTER 5    /* K4 */
#define HRM_STAGE_SHD 6       /* K5 + per-read arg-min */
#define HRM_STAGE_MERGE 7
#define HRM_STAGE_VERIFY 8    /* K6/K7 */
#define HRM_STAGE_ROUTE 9     /* key-partitioned index: routing + NCCL all-to-all exchanges */
#define HRM_NUM_STAGES 10
hrm_status hrm_mapper_set_profiling(hrm_mapper* m, int enable);
hrm_status hrm_mapper_stage_times(hrm_mapper* m, float* h_ms /* [HRM_NUM_STAGES] */,
                                  int32_t* h_spans /* [HRM_NUM_STAGES], may be NULL */);

/* ---- read ingestion on the device -------------------------------------------------------
 * ref: forEachReadInFile include/readlibraryio.hpp:288-326 (kseqpp, one host parser thread) and the encoder
 * threads of constructChunkedReadStorageFromFiles include/chunkedreadstorageconstruction.hpp:70-95,:273-314:
 * a c g t -> upper case, every other non-ACGT character -> "ACGT"[Ncount++ % 4] with Ncount = 0 at the start of
 * every batch of 65536 reads of the file, reads with such characters recorded as ambiguous.
 * d_text: FASTQ (4-line records) or FASTA (2-line records) text of whole records in device memory, < 2 GiB per
 * call.  Output: ASCII rows (d_rows, `pitch` bytes each, zero padded) + lengths, exactly what the reference hands to
 * encodeSequence2Bit and what hrm_encode_2bit / hrm_map_batch / hrm_readstore_create_from_ascii take.
 * first_read_id: id of the first read of the text within its file; carry_replaced: replaced-character count
 * carried from the previous chunk of the same batch (0 at a multiple of 65536 reads); *h_carry_replaced_out: the
 * value to pass with the next chunk.  d_ambiguous (may be NULL): 1 per read with replaced characters. */
hrm_status hrm_ingest_reads(const char* d_text, int64_t nbytes, int64_t first_read_id, int32_t carry_replaced,
                            char* d_rows, int64_t pitch, int32_t* d_lengths, uint8_t* d_ambiguous,
                            int64_t max_reads, int64_t* h_num_reads, int32_t* h_carry_replaced_out,
                            hrm_stream stream);

/* gzip'd read files (ref: the reference's reader goes through zlib, include/kseqpp/, readlibraryio.hpp:288-326):
 * host-side inflate of a whole gzip stream (all members) into a host buffer; the text then goes to
 * hrm_ingest_reads / hrm_mapper_stage_fastq.  *h_written = bytes produced; HRM_ERR_OVERFLOW when cap is too small. */
hrm_status hrm_inflate_gzip(const void* h_in, int64_t nbytes, void* h_out, int64_t cap, int64_t* h_written);

/* ---- key-partitioned index over the GPUs of one box (BASELINE config 5) -------------
 * ref: the reference's multi-GPU minhasher distributes whole tables (hash function j on GPU j mod G),
 * broadcasts every query batch to all GPUs and gathers the results with cudaMemcpyPeerAsync
 * (include/gpu/multigpuminhasher.cuh:257-333, :659-675, :724, :835-870).  Here every table is split by
 * key, owner(key) = hrm_key_owner(key, G); one process per GPU; the (read, table) lookups of a batch are
 * routed to their owners and the value lists come back with NCCL all-to-all (ncclSend/ncclRecv groups).
 * Results are identical to the replicated index.
 *
 * Bootstrap: rank 0 calls hrm_comm_unique_id and ships the HRM_COMM_ID_BYTES bytes to the other ranks by
 * whatever channel the application has (torch.distributed broadcast in the Python harness, a file, MPI);
 * every rank then calls hrm_comm_create (collective).  hrm_mapper_set_partition must precede
 * hrm_mapper_set_genome; afterwards hrm_map_batch is collective: all ranks call it once per batch
 * (a rank without reads passes n = 0). */
#define HRM_COMM_ID_BYTES 128
typedef struct hrm_comm hrm_comm;
typedef struct {
    int32_t rank, world;
    int64_t bytes_sent;      /* data-path bytes sent to other ranks since creation */
    int64_t bytes_received;
    int64_t exchanges;       /* all-to-all rounds */
} hrm_comm_info_t;
hrm_status hrm_comm_unique_id(void* out_id, int64_t capacity);
hrm_status hrm_comm_create(hrm_comm** out, int rank, int world, const void* id_bytes);
void hrm_comm_destroy(hrm_comm* c);
hrm_status hrm_comm_info(const hrm_comm* c, hrm_comm_info_t* out);
int hrm_key_owner(uint64_t key, int world);
/* restricts a minhasher under construction to the keys owned by `rank` (call before the first insert) */
hrm_status hrm_minhasher_set_partition(hrm_minhasher* mh, int rank, int world);
hrm_status hrm_mapper_set_partition(hrm_mapper* m, hrm_comm* comm);

/* ------------------------------------------------------------------------------------------
 * V4 + O1 -- score recalculation, conversion count, choice of the alignment, MAPQ, POS and the SAM text,
 * all on the device.
 * ref: recalculateAlignmentScorefk / comparefk src/gpu/mappinghandler.cu:601-766, mapqfkt :184-193,
 *      printtoSAM :196-293 (SW mode).
 * Semantics: the reference with the two patches of SURVEY's parity contract (the two query strings owned;
 * rc_ref reads NUL beyond the chromosome) and every other quirk kept -- alignment 0 is walked with the
 * reverse-complement query, only the first 82 bases are examined, scores are uint16_t and wrap, MAPQ is 4
 * whenever the reference's double -> uint32_t conversion is out of range (x86-64).  A pass sees its own
 * converted reads and genome (= the reference run on pre-converted input); unmapped reads print from pass 0.
 * The parity tests check this against the reference's own Mappinghandler compiled unmodified.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t sw_score[2];            /* after the recalculation, as the reference's uint16_t holds them */
    int32_t sw_score_next_best[2];
    int32_t num_conversions[2];     /* Yf:i:<n> of alignment 0 / 1 */
    int32_t chosen;                 /* 0: alignment 0 printed (YZ:A:<+>), 1: alignment 1 (YZ:A:<->); ref :222 */
    int32_t flag;                   /* FLAG column */
    int32_t mapq;                   /* MAPQ column */
    int32_t window_length;          /* LN of the read's @SQ line */
    int64_t pos;                    /* POS column = window start + query_begin (ref :236) */
} hrm_sam_fields;

/* V4 alone: d_fields[n] from the records and cigars of hrm_verify_batch (same reads, same order). */
hrm_status hrm_sam_fields_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                                const int32_t* d_lengths, int64_t n, const hrm_read_record* d_records,
                                const char* d_cigars, int64_t cigar_pitch, hrm_sam_fields* d_fields,
                                hrm_stream stream);

/* The SAM file of printtoSAM is  HRM_SAM_HD | the @SQ line of every read | HRM_SAM_PG_CO | the record line of
 * every read.  The two per-read parts are produced separately so that a run of many batches can be assembled. */
#define HRM_SAM_HD "@HD\tVN:1.4\n"
#define HRM_SAM_PG_CO "@PG\tHashreadmapper\tID:1.0@CO: QNAME\tFLAG\tRNAME\tPOS\tMAPQ\tCIGAR\tRNEXT\tPNEXT\tTLEN\tSEQ\tQUAL\tTAG\n"
#define HRM_SAM_SQ_LINES 0
#define HRM_SAM_RECORDS 1

/* Text of one part for n reads into DEVICE memory.  *h_written = bytes of the part (pass d_out = NULL to size
 * it); lines that would cross `cap` are not written.  cigar_pitch must cover the longest cigar (longer strings
 * print truncated, hrm_alignment.cigar_len tells).  h_chrom_names: n_chrom NUL-terminated names (NULL: the
 * chromosome index).  first_read_id = id of read 0 (QNAME = read id, ref :252).  d_reads_ascii = NULL: the
 * batch is still packed inside the mapper (hrm_verify_batch just ran on these reads).  Synchronises the stream. */
hrm_status hrm_sam_format_device(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                                 const int32_t* d_lengths, int64_t n, const hrm_read_record* d_records,
                                 const char* d_cigars, int64_t cigar_pitch, uint32_t first_read_id,
                                 const char* const* h_chrom_names, int part, char* d_out, int64_t cap,
                                 int64_t* h_written, hrm_stream stream);

/* Host buffers in, host text out (records / cigars as hrm_mapper_map_reads returned them): copies them to the
 * device, formats there and copies the text back.  `with_header` emits HRM_SAM_HD, the @SQ lines and
 * HRM_SAM_PG_CO in front of the records.  *h_written = bytes needed (h_out = NULL to size). */
hrm_status hrm_sam_format(hrm_mapper* m, const hrm_read_record* h_records,
                          const char* h_cigars, int64_t cigar_pitch, const char* h_reads_ascii,
                          int64_t ascii_pitch, const int32_t* h_lengths, int64_t n,
                          uint32_t first_read_id, const char* const* h_chrom_names, int with_header,
                          char* h_out, int64_t cap, int64_t* h_written);

/* End to end, text out: reads H2D, seeding + filter + SHD + verification + V4, SAM text D2H.
 * ref: STEP 1 + STEP 2 of performMappingGpu (main_gpu.cu:1123-1160) incl. the SAM writer.
 * h_sq_out / h_rec_out receive the @SQ lines / the record lines of these reads (pinned for full PCIe speed);
 * h_records / h_cigars (may be NULL) additionally receive the binary records.  Synchronous. */
hrm_status hrm_mapper_map_reads_sam(hrm_mapper* m, const char* h_reads_ascii, int64_t ascii_pitch,
                                    const int32_t* h_lengths, int64_t n, uint32_t first_read_id,
                                    const char* const* h_chrom_names, char* h_sq_out, int64_t sq_cap,
                                    int64_t* h_sq_written, char* h_rec_out, int64_t rec_cap,
                                    int64_t* h_rec_written, hrm_read_record* h_records, char* h_cigars,
                                    int64_t cigar_pitch, hrm_batch_stats* h_stats, hrm_stream stream);

/* ---- double-buffered end-to-end pipeline ----------------------------------------------------------------
 * ref: the batch loop of performMappingGpu (src/gpu/main_gpu.cu:1123-1160), which overlaps nothing.  A batch lives in
 * one of HRM_PIPE_SLOTS slots:
 *   hrm_mapper_stage_reads  enqueues the H2D copy of a batch into a slot (returns at once; waits first if the slot's
 *                           previous results have not left the device yet);
 *   hrm_mapper_map_staged   runs seeding + filter + SHD + best window (K1..K5) of the staged batch on `stream` and queues
 *                           its verification (K6/K7), V4 and -- when h_rec_out != NULL -- the SAM text behind it, and
 *                           the D2H copies of records / CIGARs on the slot's copy-out stream; returns when the seeding has
 *                           run, not when the rest has finished (HRM_PIPE_OVERLAP=1: verification on a second stream,
 *                           under the seeding of the next batch);
 *   hrm_mapper_finish       waits until the slot's results are in the host buffers (copies the SAM text out, whose
 *                           size is known only then); returns the text sizes.
 * Steady state:  map_staged(i) ; finish(i-1) ; stage(i+1)  -- staging into a slot waits for the slot's previous batch,
 * so the next batch is staged after the previous one was fetched.  Host buffers should be pinned.  h_records / h_cigars /
 * h_sq_out / h_rec_out may each be NULL; with text, rec_cap >= n * (64 + longest chromosome name + cigar_pitch +
 * window + ascii_pitch) and sq_cap >= n * 40 (the text is produced into device buffers of that bound). */
#define HRM_PIPE_SLOTS 2
hrm_status hrm_mapper_stage_reads(hrm_mapper* m, int slot, const char* h_reads_ascii, int64_t ascii_pitch,
                                  const int32_t* h_lengths, int64_t n);
/* Same for reads that are ALREADY in device memory (no copy; they must stay valid until hrm_mapper_finish of the slot).
 * max_length = an upper bound of the read lengths (<= ascii_pitch); ready_on = the stream that produced the buffers. */
hrm_status hrm_mapper_stage_device(hrm_mapper* m, int slot, const char* d_reads_ascii, int64_t ascii_pitch,
                                   const int32_t* d_lengths, int64_t n, int max_length, hrm_stream ready_on);
/* Same as hrm_mapper_stage_reads from FASTQ / FASTA TEXT in host memory (whole records, < 2 GiB): H2D on the copy-in
 * stream, parsed on the device by hrm_ingest_reads (same arguments and semantics).  Blocks the calling thread until
 * the batch is parsed; to overlap it with hrm_mapper_map_staged of the other slot call it from a second host thread
 * (the two calls share no state). */
hrm_status hrm_mapper_stage_fastq(hrm_mapper* m, int slot, const char* h_text, int64_t nbytes, int64_t first_read_id,
                                  int32_t carry_replaced, int64_t ascii_pitch, int64_t max_reads,
                                  int64_t* h_num_reads, int32_t* h_carry_replaced_out);
hrm_status hrm_mapper_map_staged(hrm_mapper* m, int slot, hrm_read_record* h_records, char* h_cigars,
                                 int64_t cigar_pitch, uint32_t first_read_id, const char* const* h_chrom_names,
                                 char* h_sq_out, int64_t sq_cap, char* h_rec_out, int64_t rec_cap,
                                 hrm_batch_stats* h_stats, hrm_stream stream);
hrm_status hrm_mapper_finish(hrm_mapper* m, int slot, int64_t* h_sq_written, int64_t* h_rec_written);

#ifdef __cplusplus
}
#endif
#endif /* HRM_B200_H */
