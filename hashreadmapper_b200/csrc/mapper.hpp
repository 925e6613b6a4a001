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

struct hrm_mapper {
    hrm_mapper_config cfg;
    // one genome + index per distinct genome conversion
    hrm_genome* genome[3] = {nullptr, nullptr, nullptr};
    hrm_minhasher* index[3] = {nullptr, nullptr, nullptr};
    int index_handle[3] = {-1, -1, -1};
    int64_t* d_win_prefix = nullptr; // n_chrom + 1
    int64_t num_windows = 0;
    int n_chrom = 0;
    std::vector<int64_t> chrom_len;
    std::string host_genome; // kept for SAM output (RNEXT column prints the window, ref: mappinghandler.cu:257)
    std::vector<int64_t> chrom_off;
    // per-batch buffers that only grow
    hrm::GrowBuf packed[3];  // reads packed per read conversion
    hrm::GrowBuf sigs, num, off, newoff, passres, misc;
    int64_t packed_pitch = 0;
    unsigned long long touches_seen[3] = {0, 0, 0};
    hrm::StageTimer timer;
    int64_t value_budget = 1LL << 30; // candidate values retrieved per range of reads (int offsets, 8 B scratch each)
    bool use_fused = true;            // K3b retrieval + K4 fused (k4_fused.cu) on the replicated index
    int64_t collect_enumerated = 0, collect_skipped = 0; // ids counted / skipped (largest buckets) by the fused path
    int64_t part_chunk = 1 << 17;     // reads per routed query of the key-partitioned index
    // reads per pipelined chunk of hrm_mapper_map_reads (copies of neighbouring chunks under compute).  Default: one
    // chunk -- measured on B200 at 1 M reads: 8.05 M reads/s as one chunk, 6.44 M in chunks of 262 144, 5.08 M in
    // chunks of 131 072 (per-chunk launch and sync overheads cost more than the 8 ms of hidden PCIe copies)
    int64_t e2e_chunk = 1LL << 40;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    hrm_comm* comm = nullptr; // key-partitioned index (partition.cu); not owned
};
