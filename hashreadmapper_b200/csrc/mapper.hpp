// mapper.hpp -- the mapper object (shared by mapper.cu and sam.cu).
#pragma once
#include "runtime.cuh"
#include "store.cuh"
#include <string>
#include <vector>

struct hrm_minhasher;
struct hrm_comm;

namespace hrm {
// CUDA-event spans around the stages of the fused path (enabled by hrm_mapper_set_profiling)
struct StageTimer {
    bool enabled = false;
    struct Span {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get()
    {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int stage, cudaStream_t s)
    {
        if (!enabled) return;
        Span sp{stage, get(), get()};
        cudaEventRecord(sp.a, s);
        spans.push_back(sp);
    }
    void end(cudaStream_t s)
    {
        if (!enabled || spans.empty()) return;
        cudaEventRecord(spans.back().b, s);
    }
    ~StageTimer()
    {
        for (auto& sp : spans) {
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        for (auto e : pool) cudaEventDestroy(e);
    }
};
} // namespace hrm

namespace hrm {
// per-batch device buffers of the fused path (they only grow)
struct BatchCtx {
    GrowBuf packed[3];  // reads packed per read conversion
    GrowBuf sigs, num, off, newoff, passres, misc;
    int64_t packed_pitch = 0;
};

// one batch of the double-buffered end-to-end pipeline (mapper.cu: hrm_mapper_stage_reads / map_staged / finish)
struct PipeSlot {
    GrowBuf ascii, len, mapped, rec, cig, text, sq, fastq, fields, lens2, offs;
    const char* d_ascii = nullptr;  // reads staged from device memory (hrm_mapper_stage_device): not owned
    const int32_t* d_len = nullptr;
    cudaEvent_t staged = nullptr, seeded = nullptr, computed = nullptr, drained = nullptr;
    int64_t n = 0, pitch = 0, sq_written = 0, rec_written = 0;
    int maxlen = 0;
    char* h_rec = nullptr;          // host text buffers of the batch in flight (copied out by hrm_mapper_finish)
    char* h_sq = nullptr;
    bool is_staged = false, busy = false, want_text = false;
};
} // namespace hrm

struct hrm_mapper {
    hrm_mapper_config cfg;
    int device = 0; // the device that was current at hrm_mapper_create: every entry point makes it current again
    // one genome + index per distinct genome conversion
    hrm_genome* genome[3] = {nullptr, nullptr, nullptr};
    hrm_minhasher* index[3] = {nullptr, nullptr, nullptr};
    int index_handle[3] = {-1, -1, -1};
    int64_t* d_win_prefix = nullptr; // n_chrom + 1
    int64_t num_windows = 0;
    int n_chrom = 0;
    std::vector<int64_t> chrom_len;
    std::vector<int64_t> chrom_off;
    // chromosome names on the device for the SAM writer (sam.cu), uploaded when they change
    std::string names_flat;
    hrm::GrowBuf d_names, d_name_off;
    // per-batch device state: one context per pipeline slot + one for the plain (one batch at a time) entry points;
    // `bc` is the context the next hrm_map_batch / hrm_verify_batch / SAM call works on
    hrm::BatchCtx ctx[HRM_PIPE_SLOTS + 1];
    hrm::BatchCtx* bc = &ctx[HRM_PIPE_SLOTS];
    unsigned long long touches_seen[3] = {0, 0, 0};
    hrm::StageTimer timer;
    int64_t value_budget = 1LL << 30; // candidate values retrieved per range of reads (int offsets, 8 B scratch each)
    bool use_fused = true;            // K3b retrieval + K4 fused (k4_fused.cu) on the replicated index
    int64_t collect_enumerated = 0, collect_skipped = 0; // ids counted / skipped (largest buckets) by the fused path
    int64_t collect_block_reads = 0;                     // reads the warp kernel handed to the block kernel
    int64_t part_chunk = 1 << 17;     // reads per routed query of the key-partitioned index
    // reads per pipelined chunk of hrm_mapper_map_reads (copies of neighbouring chunks under compute).  Default: one
    // chunk -- measured on B200 at 1 M reads: 8.05 M reads/s as one chunk, 6.44 M in chunks of 262 144, 5.08 M in
    // chunks of 131 072 (per-chunk launch and sync overheads cost more than the 8 ms of hidden PCIe copies)
    int64_t e2e_chunk = 1LL << 40;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    hrm_comm* comm = nullptr; // key-partitioned index (partition.cu); not owned
    // double-buffered end-to-end pipeline
    hrm::PipeSlot slot[HRM_PIPE_SLOTS];
    // one copy-out stream per slot: the text copy of batch i - 1 is queued (in hrm_mapper_finish, once its size is known)
    // AFTER batch i has made its own stream wait for its verification -- on a shared stream it would wait for that too
    cudaStream_t pipe_in = nullptr, pipe_out[HRM_PIPE_SLOTS] = {}, pipe_verify = nullptr;
    int64_t* pipe_host = nullptr; // pinned: text sizes of the batches in flight
    bool pipe_ready = false;
};

namespace hrm {
// packs the batch once per distinct read conversion used by the passes (mapper.cu)
hrm_status mapper_pack_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch, const int32_t* d_lengths,
                             int64_t n, hrm_stream stream);
// V4 + O1 on the device (sam.cu)
hrm_status sam_fields(hrm_mapper* m, const int32_t* d_lengths, int64_t n, const hrm_read_record* d_records,
                      const char* d_cigars, int64_t cigar_pitch, uint32_t first_read_id, hrm_sam_fields* d_fields,
                      int32_t* d_line_len, int32_t* d_sq_len, cudaStream_t s);
hrm_status sam_upload_names(hrm_mapper* m, const char* const* h_chrom_names, cudaStream_t s);
// same as sam_text without the host round trip: offsets into d_off (n + 2 entries), the byte total into *h_total_pinned
// when the stream gets there
hrm_status sam_text_async(hrm_mapper* m, const int32_t* d_lengths, int64_t n, const hrm_read_record* d_records,
                          const char* d_cigars, int64_t cigar_pitch, const hrm_sam_fields* d_fields, const int32_t* d_len,
                          uint32_t first_read_id, int part, char* d_out, int64_t cap, int64_t* d_off,
                          int64_t* h_total_pinned, cudaStream_t s);
hrm_status sam_text(hrm_mapper* m, const int32_t* d_lengths, int64_t n, const hrm_read_record* d_records,
                    const char* d_cigars, int64_t cigar_pitch, const hrm_sam_fields* d_fields, const int32_t* d_len,
                    uint32_t first_read_id, int part, char* d_out, int64_t cap, int64_t* h_written, cudaStream_t s);
} // namespace hrm
