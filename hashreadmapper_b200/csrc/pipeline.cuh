// pipeline.cuh -- internal declarations shared by mapper.cu and the kernel files.
#pragma once
#include "runtime.cuh"
#include "store.cuh"

struct hrm_minhasher;

namespace hrm {

// d_scratch: as many entries as d_values (partition space of the counting kernel; contents undefined afterwards)
hrm_status filter_segments(uint32_t* d_values, uint32_t* d_scratch, const int32_t* d_offsets, int n, int min_hits,
                           int32_t* d_new_counts, int32_t* d_new_offsets, int64_t* d_total64, cudaStream_t s);
hrm_status compact_segments(const uint32_t* d_values, const int32_t* d_old_offsets, const int32_t* d_new_offsets, int n,
                            uint32_t* d_out, cudaStream_t s);
hrm_status best_windows(const uint32_t* d_reads, int64_t read_pitch, const int32_t* d_read_len, int64_t n,
                        const uint32_t* d_cand_windows, const int32_t* d_cand_offsets, const hrm_genome* g,
                        const int64_t* d_win_prefix, int k, int w, float rate, int pass, hrm_mapped_read* d_out,
                        cudaStream_t s, const int2* d_cand_lists = nullptr);

// K3b retrieval + K4 fused (k4_fused.cu)
hrm_status collect_candidates(const hrm_minhasher* mh, const struct QueryHandle* qh, int n, int min_hits, uint32_t id_space,
                              uint32_t* d_out, int64_t out_cap, int2* d_lists, int64_t* h_total, int* h_overflow,
                              int64_t* h_stats3, cudaStream_t s);

hrm_status collect_candidates_from(const uint2* d_ranges, int64_t rq, int64_t rt, const uint32_t* d_table_values, int H, int n,
                                   int min_hits, uint32_t id_space, uint32_t* d_out, int64_t out_cap, int2* d_lists,
                                   int64_t* h_total, int* h_overflow, int64_t* h_stats3, cudaStream_t s);

// verification inputs of one pass: reads packed with the pass' read conversion, genome packed with
// its genome conversion, stage-V conversion applied on the fly
struct VerifyPass {
    const uint32_t* reads;
    int64_t read_pitch;
    GenomeDev G;
    int32_t verify_conv;
};
struct VerifyParams {
    VerifyPass pass[HRM_MAX_PASSES];
    int num_passes;
    int w;
    int mapper_type;
};
hrm_status verify_reads(const VerifyParams& VP, const int32_t* d_read_len, int64_t n, int max_read_len,
                        const hrm_mapped_read* d_mapped, hrm_read_record* d_records, char* d_cigars, int64_t cigar_pitch,
                        cudaStream_t s);

} // namespace hrm
