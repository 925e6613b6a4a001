// k3_table.cuh -- internal interface of the minhasher object (K3).
#pragma once
#include "runtime.cuh"
#include <vector>
#include <memory>
#include <mutex>

namespace hrm {

// 16-byte slot; four slots form one 64-byte bucket.  64 B is what one HBM access moves on B200 (ncu: a
// random 32-byte sector read costs 64 DRAM bytes), so a lookup that ends in its home bucket costs exactly
// one DRAM access and examines everything that access brought in.
struct alignas(16) Slot {
    uint64_t key;
    uint32_t off;   // index into the minhasher's value array
    uint32_t cnt;   // bucket size after truncation
};
static_assert(sizeof(Slot) == 16, "slot must be 16 bytes (SURVEY 8d: 16 B per probe)");

constexpr uint64_t SLOT_EMPTY = ~0ULL;   // ref: emptySlot cpuhashtable.hpp:277-278
constexpr int MAX_TABLES = 64;           // ref: assert(numTables <= 64) fakegpuminhasher.cuh:541

constexpr int BUCKET_SLOTS = 4;
constexpr int BUCKET_BYTES = BUCKET_SLOTS * (int)sizeof(Slot);

struct TableRef {
    const Slot* slots;
    uint32_t nbuckets; // any count >= 1: home bucket = mulhi(hash32, nbuckets), then linear over buckets
    uint32_t pad;
};
struct TablesParam {
    TableRef t[MAX_TABLES];
};

struct QueryHandle {
    GrowBuf sigs;    // [n][H] u64 (when hashing inside count)
    GrowBuf ranges;  // uint2 (off, cnt): [n][H], or [H][n] after a table-major probe (tm)
    GrowBuf sigs_tm; // [H][n] u64: signatures transposed for the table-major probe
    bool tm = false; // layout of `ranges`
    int64_t rq = 0, rt = 0; // index of (query q, table t) in `ranges` = q * rq + t * rt
    GrowBuf misc;    // totals
    int stage = 0;   // 0 none, 1 counted
    int n = 0;
    bool in_use = false;
};

} // namespace hrm

struct hrm_minhasher {
    int k = 16;
    int max_results = 65535;
    float load = 0.8f;
    int64_t max_sequences = 0;
    int H = 0;
    int64_t inserted = 0;
    bool compacted = false;
    bool finished = false;
    int device = 0;
    int part_rank = 0, part_world = 1; // key partition: only keys owned by part_rank are inserted
    // build staging, one pair of device arrays per table
    std::vector<uint64_t*> stage_keys;
    std::vector<uint32_t*> stage_vals;
    std::vector<int64_t> table_count; // sequences staged per table
    // compacted form
    uint32_t* values = nullptr;      // H * inserted entries; table j at [j*inserted, ...)
    int64_t values_count = 0;
    std::vector<hrm::Slot*> slots;   // per table
    std::vector<int64_t> nbuckets;
    std::vector<int64_t> nkeys;
    hrm::TablesParam param;
    hrm::TablesParam* d_param = nullptr; // device copy read by the probe kernel
    // handles
    std::mutex mtx;
    std::vector<std::unique_ptr<hrm::QueryHandle>> handles;
    // counters of the last count call (slot touches), device side
    unsigned long long* d_touches = nullptr;
};

namespace hrm {
// device-side pieces used by the fused mapper
hrm_status minhasher_count_sigs(hrm_minhasher* mh, QueryHandle* qh, const uint64_t* d_sigs, int n,
                                int32_t* d_num_per_seq, cudaStream_t s);
// table-major probe for indexes far larger than L2 / the TLB reach (k3_table.cu): all lookups of table 0, then table
// 1, ... so that the lookups in flight share one table.  Three steps so that the caller can time the probe alone.
bool minhasher_wants_table_major(const hrm_minhasher* mh);
hrm_status minhasher_tm_prepare(hrm_minhasher* mh, QueryHandle* qh, const uint64_t* d_sigs, int n, cudaStream_t s);
hrm_status minhasher_tm_probe(hrm_minhasher* mh, QueryHandle* qh, int n, cudaStream_t s);
hrm_status minhasher_tm_totals(hrm_minhasher* mh, QueryHandle* qh, int n, int32_t* d_num_per_seq, cudaStream_t s);
// values of queries [first, first + n) of the last count; d_offsets: n + 1 offsets relative to d_values
hrm_status minhasher_retrieve(hrm_minhasher* mh, QueryHandle* qh, int64_t first, int n, uint32_t* d_values,
                              const int32_t* d_offsets, cudaStream_t s);
QueryHandle* minhasher_handle(hrm_minhasher* mh, int id);
hrm_status minhash_rows(const uint32_t* d_seq2bit, int64_t pitch_words, const int32_t* d_lengths, int64_t n, int k,
                        int H, uint64_t* d_sigs, uint8_t* d_valid, cudaStream_t s);
hrm_status minhash_windows(const uint32_t* d_chrom2bit, int64_t chrom_len, int k, int w, int H, int64_t first_window,
                           int64_t n_windows, uint64_t* d_sigs, uint8_t* d_valid, cudaStream_t s);
} // namespace hrm
