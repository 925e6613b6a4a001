// pipeline.cuh -- internal declarations shared by mapper.cu and the kernel files.
#pragma once
#include "runtime.cuh"
#include "store.cuh"

namespace hrm {

hrm_status filter_segments(uint32_t* d_values, const int32_t* d_offsets, int n, int min_hits, int32_t* d_new_counts,
                           int32_t* d_new_offsets, int64_t* d_total64, cudaStream_t s);
hrm_status compact_segments(const uint32_t* d_values, const int32_t* d_old_offsets, const int32_t* d_new_offsets, int n,
                            uint32_t* d_out, cudaStream_t s);
hrm_status best_windows(const uint32_t* d_reads, int64_t read_pitch, const int32_t* d_read_len, int64_t n,
                        const uint32_t* d_cand_windows, const int32_t* d_cand_offsets, const hrm_genome* g,
                        const int64_t* d_win_prefix, int k, int w, float rate, int pass, hrm_mapped_read* d_out,
                        cudaStream_t s);

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
