// mapper.cu -- the fused read-mapping path in the north_star direction: a 3N index over the
// reference windows lives in HBM and batches of reads probe it.
// ref: performMappingGpu src/gpu/main_gpu.cu:859-1160 (driver), WindowBatchProcessor::operator()
//      :471-854 (per 2048-window batch: H2D ASCII, encode, minhash, CPU table query, H2D ids, filter,
//      gather reads, H2D genome slice, extended windows, HiLo, SHD, D2H, host arg-min; >= 6 stream
//      syncs per batch), Mappinghandler::go src/gpu/mappinghandler.cu:67 (verification on the host).
// Candidate (read, window) pairs are symmetric in the direction of the index, so
// indexing windows and probing with reads yields the same pairs as long as no bucket exceeds
// min(maxResultsPerMap, 65535) entries (then the truncation applies to windows instead of reads --
// DESIGN.md "bucket caps").  One pass = one (read conversion, genome conversion) pair = exactly the
// reference pipeline on pre-converted input; passes are merged per read by smaller Hamming distance,
// earlier pass first.  One stream sync per pass (to size the candidate buffer), none per window.
#include "pipeline.cuh"
#include "k3_table.cuh"
#include "mapper.hpp"
#include "partition.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace hrm {

__global__ void __launch_bounds__(256) merge_pass_kernel(hrm_mapped_read* __restrict__ best,
                                                         const hrm_mapped_read* __restrict__ cur, int64_t n, int first)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const hrm_mapped_read c = cur[i];
        if (first) {
            best[i] = c;
            continue;
        }
        const hrm_mapped_read b = best[i];
        if (c.orientation != HRM_ORIENT_NONE &&
            (b.orientation == HRM_ORIENT_NONE || c.hamming_distance < b.hamming_distance))
            best[i] = c;
    }
}

__global__ void __launch_bounds__(256) count_mapped_kernel(const hrm_mapped_read* __restrict__ m, int64_t n,
                                                           unsigned long long* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        local += m[i].orientation != HRM_ORIENT_NONE ? 1 : 0;
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

__global__ void max_len_kernel(const int32_t* __restrict__ len, int64_t n, int* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) local = max(local, len[i]);
    local = __reduce_max_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0) atomicMax(out, local);
}

static unsigned mgrid(int64_t items)
{
    int64_t g = HRM_SDIV(items, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}
} // namespace hrm

using namespace hrm;

// A mapper belongs to one device.  Staging threads (hrm_mapper_stage_* are meant to be called from a second host thread)
// start with device 0 current, so every entry point binds the calling thread to the mapper's device first.
static hrm_status bind_device(const hrm_mapper* m)
{
    int cur = -1;
    HRM_CUDA(cudaGetDevice(&cur));
    if (cur != m->device) HRM_CUDA(cudaSetDevice(m->device));
    return HRM_OK;
}

extern "C" void hrm_mapper_default_config(hrm_mapper_config* cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->k = 16;                      // ref: options.hpp kmerlength
    cfg->window_size = 128;           // ref: options.hpp windowSize
    cfg->num_tables = 16;             // ref: options.hpp hashmaps
    cfg->min_table_hits = 4;          // ref: options.hpp minTableHits
    cfg->max_results_per_map = 65535; // ref: options.hpp maxResultsPerMap
    // ref: options.hpp hashtableLoadfactor is 0.8 for the reference's 16-byte-slot linear-probing table.  The
    // bucketized table here is sized at 0.5: 1.06 HBM accesses per lookup instead of 2.0 at 0.8 (measured)
    cfg->load_factor = 0.5f;
    cfg->max_hamming_percent = 0.05f; // ref: options.hpp maxHammingPercent
    cfg->mapper_type = HRM_MAPPER_SW;
    cfg->num_passes = 1;
    cfg->read_conversion[0] = HRM_CONV_NONE;
    cfg->genome_conversion[0] = HRM_CONV_NONE;
    cfg->verify_conversion[0] = HRM_CONV_CT; // ref: stage V always converts C->T (mappinghandler.cu:456-473)
}

extern "C" hrm_status hrm_mapper_create(hrm_mapper** out, const hrm_mapper_config* cfg)
{
    HRM_REQUIRE(out != nullptr && cfg != nullptr, "args");
    *out = nullptr;
    HRM_TRY(ensure_device());
    HRM_REQUIRE(cfg->k >= 1 && cfg->k <= 32, "1 <= k <= 32 (ref: main_gpu.cu:1305)");
    HRM_REQUIRE(cfg->num_tables >= 1 && cfg->num_tables <= 48, "1 <= hashmaps <= 48 (ref: main_gpu.cu:1306)");
    HRM_REQUIRE(cfg->num_tables >= cfg->min_table_hits, "hashmaps >= minTableHits (ref: main_gpu.cu:1309)");
    HRM_REQUIRE(cfg->window_size >= cfg->k && cfg->window_size <= HRM_SW_MAX_REF, "k <= windowSize <= 512");
    HRM_REQUIRE(cfg->num_passes >= 1 && cfg->num_passes <= HRM_MAX_PASSES, "1 <= num_passes <= 4");
    HRM_REQUIRE(cfg->load_factor > 0.f && cfg->load_factor <= 1.f, "load factor");
    for (int p = 0; p < cfg->num_passes; p++) {
        HRM_REQUIRE(cfg->read_conversion[p] >= 0 && cfg->read_conversion[p] <= 2, "read_conversion");
        HRM_REQUIRE(cfg->genome_conversion[p] >= 0 && cfg->genome_conversion[p] <= 2, "genome_conversion");
        HRM_REQUIRE(cfg->verify_conversion[p] >= 0 && cfg->verify_conversion[p] <= 2, "verify_conversion");
    }
    int dev = 0;
    HRM_CUDA(cudaGetDevice(&dev));
    auto* m = new hrm_mapper;
    m->cfg = *cfg;
    m->device = dev;
    if (const char* pc = getenv("HRM_PART_CHUNK")) { // test hook: chunked collective query at small sizes
        const long long v = atoll(pc);
        if (v > 0) m->part_chunk = v;
    }
    if (const char* ec = getenv("HRM_E2E_CHUNK")) { // reads per pipelined chunk of hrm_mapper_map_reads
        const long long v = atoll(ec);
        if (v > 0) m->e2e_chunk = v;
    }
    if (const char* cf = getenv("HRM_COLLECT")) m->use_fused = atoi(cf) != 0; // 0: general retrieve + filter path only
    if (const char* vb = getenv("HRM_VALUE_BUDGET")) { // test hook: forces the range splitting at small sizes
        const long long v = atoll(vb);
        if (v > 0 && v < m->value_budget) m->value_budget = v;
    }
    *out = m;
    return HRM_OK;
}

extern "C" void hrm_mapper_destroy(hrm_mapper* m)
{
    if (!m) return;
    bind_device(m);
    for (int c = 0; c < 3; c++) {
        if (m->index[c]) hrm_minhasher_destroy(m->index[c]);
        if (m->genome[c]) hrm_genome_destroy(m->genome[c]);
    }
    if (m->d_win_prefix) cudaFree(m->d_win_prefix);
    for (auto e : m->copy_events) cudaEventDestroy(e);
    if (m->pipe_ready) {
        for (int i = 0; i < HRM_PIPE_SLOTS; i++) {
            cudaEventDestroy(m->slot[i].staged);
            cudaEventDestroy(m->slot[i].seeded);
            cudaEventDestroy(m->slot[i].computed);
            cudaEventDestroy(m->slot[i].drained);
        }
        cudaStreamDestroy(m->pipe_in);
        for (int i = 0; i < HRM_PIPE_SLOTS; i++) cudaStreamDestroy(m->pipe_out[i]);
        cudaStreamDestroy(m->pipe_verify);
        cudaFreeHost(m->pipe_host);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    delete m;
}

static hrm_status set_genome_impl(hrm_mapper* m, const char* h_ascii, const int64_t* h_chrom_offsets, int n_chrom,
                                  hrm_stream stream);

extern "C" hrm_status hrm_mapper_set_genome(hrm_mapper* m, const char* h_ascii, const int64_t* h_chrom_offsets,
                                            int n_chrom, hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr && h_ascii != nullptr && h_chrom_offsets != nullptr && n_chrom >= 1, "args");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(m->d_win_prefix == nullptr && m->genome[0] == nullptr && m->genome[1] == nullptr && m->genome[2] == nullptr,
                "genome already set");
    const hrm_status st = set_genome_impl(m, h_ascii, h_chrom_offsets, n_chrom, stream);
    if (st != HRM_OK) { // nothing half-built stays behind: the mapper is as it was before the call and may be retried
        cudaStreamSynchronize(as_stream(stream));
        for (int c = 0; c < 3; c++) {
            if (m->index[c]) hrm_minhasher_destroy(m->index[c]);
            if (m->genome[c]) hrm_genome_destroy(m->genome[c]);
            m->index[c] = nullptr;
            m->genome[c] = nullptr;
            m->index_handle[c] = -1;
        }
        if (m->d_win_prefix) cudaFree(m->d_win_prefix);
        m->d_win_prefix = nullptr;
        m->num_windows = 0;
        m->n_chrom = 0;
        m->chrom_len.clear();
        m->chrom_off.clear();
    }
    return st;
}

static hrm_status set_genome_impl(hrm_mapper* m, const char* h_ascii, const int64_t* h_chrom_offsets, int n_chrom,
                                  hrm_stream stream)
{
    cudaStream_t s = as_stream(stream);
    const hrm_mapper_config& cfg = m->cfg;
    m->n_chrom = n_chrom;
    m->chrom_off.assign(h_chrom_offsets, h_chrom_offsets + n_chrom + 1);
    const int64_t stride = cfg.window_size - cfg.k + 1;
    std::vector<int64_t> prefix(n_chrom + 1, 0);
    for (int c = 0; c < n_chrom; c++) {
        const int64_t len = h_chrom_offsets[c + 1] - h_chrom_offsets[c];
        m->chrom_len.push_back(len);
        prefix[c + 1] = prefix[c] + (len + stride - 1) / stride;
    }
    m->num_windows = prefix[n_chrom];
    HRM_REQUIRE(m->num_windows < (1LL << 32), "window ids are 32 bit");
    HRM_CUDA(cudaMalloc(&m->d_win_prefix, sizeof(int64_t) * (size_t)(n_chrom + 1)));
    HRM_CUDA(cudaMemcpy(m->d_win_prefix, prefix.data(), sizeof(int64_t) * (size_t)(n_chrom + 1), cudaMemcpyHostToDevice));

    for (int p = 0; p < cfg.num_passes; p++) {
        const int gc = cfg.genome_conversion[p];
        if (m->genome[gc]) continue;
        HRM_TRY(hrm_genome_create_from_ascii(&m->genome[gc], h_ascii, h_chrom_offsets, n_chrom, gc, stream));
        hrm_minhasher* mh = nullptr;
        HRM_TRY(hrm_minhasher_create(&mh, m->num_windows, cfg.max_results_per_map, cfg.k, cfg.load_factor));
        m->index[gc] = mh;
        if (m->comm) HRM_TRY(hrm_minhasher_set_partition(mh, comm_rank(m->comm), comm_world(m->comm)));
        const int added = hrm_minhasher_add_tables(mh, cfg.num_tables, nullptr, stream);
        if (added != cfg.num_tables) {
            set_error("not enough device memory for %d hash tables (got %d)", cfg.num_tables, added);
            return HRM_ERR_NOMEM; // ref: gpuminhasherconstruction.cu:71
        }
        // sketch the windows chromosome by chromosome in bounded chunks and stage them
        const int64_t CHUNK = 4LL << 20;
        Scratch sg, vd;
        const int64_t maxchunk = m->num_windows < CHUNK ? m->num_windows : CHUNK;
        HRM_TRY(sg.alloc(sizeof(uint64_t) * (size_t)maxchunk * cfg.num_tables, s));
        HRM_TRY(vd.alloc((size_t)maxchunk * cfg.num_tables, s));
        for (int c = 0; c < n_chrom; c++) {
            const int64_t nw = prefix[c + 1] - prefix[c];
            for (int64_t at = 0; at < nw; at += CHUNK) {
                const int64_t cnt = (nw - at) < CHUNK ? (nw - at) : CHUNK;
                HRM_TRY(minhash_windows(m->genome[gc]->chrom_words[c], m->chrom_len[c], cfg.k, cfg.window_size,
                                        cfg.num_tables, at, cnt, sg.as<uint64_t>(), vd.as<uint8_t>(), s));
                HRM_TRY(hrm_minhasher_insert_signatures(mh, sg.as<uint64_t>(), vd.as<uint8_t>(), cnt, nullptr,
                                                        (uint32_t)(prefix[c] + at), stream));
            }
        }
        HRM_TRY(hrm_minhasher_compact(mh, stream));
        HRM_TRY(hrm_minhasher_finish(mh, stream));
        m->index_handle[gc] = hrm_minhasher_handle_create(mh);
    }
    HRM_CUDA(cudaStreamSynchronize(s));
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_info(const hrm_mapper* m, hrm_mapper_info_t* out)
{
    HRM_REQUIRE(m != nullptr && out != nullptr, "args");
    memset(out, 0, sizeof *out);
    out->num_windows = m->num_windows;
    out->num_passes = m->cfg.num_passes;
    out->collect_ids_counted = m->collect_enumerated;
    out->collect_ids_skipped = m->collect_skipped;
    out->collect_reads_block_kernel = m->collect_block_reads;
    for (int c = 0; c < 3; c++) {
        if (m->genome[c]) out->genome_device_bytes += m->genome[c]->device_bytes();
        if (m->index[c]) {
            hrm_minhasher_info_t mi;
            hrm_minhasher_info(m->index[c], &mi);
            out->index_device_bytes += mi.device_bytes;
            out->num_keys_total += mi.num_keys_total;
            for (int j = 0; j < m->index[c]->H; j++) out->table_slots_total += m->index[c]->nbuckets[j] * hrm::BUCKET_SLOTS;
        }
    }
    return HRM_OK;
}

// packs the batch once per distinct read conversion used by the passes
hrm_status hrm::mapper_pack_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch, const int32_t* d_lengths,
                                  int64_t n, hrm_stream stream)
{
    const hrm_mapper_config& cfg = m->cfg;
    m->bc->packed_pitch = ascii_pitch / 16;
    bool done[3] = {false, false, false};
    for (int p = 0; p < cfg.num_passes; p++) {
        const int rc = cfg.read_conversion[p];
        if (done[rc]) continue;
        done[rc] = true;
        HRM_TRY(m->bc->packed[rc].reserve(sizeof(uint32_t) * (size_t)n * m->bc->packed_pitch));
        HRM_TRY(hrm_encode_2bit(d_reads_ascii, ascii_pitch, d_lengths, n, rc, m->bc->packed[rc].as<uint32_t>(),
                                m->bc->packed_pitch, stream));
    }
    return HRM_OK;
}

extern "C" hrm_status hrm_map_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                                    const int32_t* d_lengths, int64_t n, hrm_mapped_read* d_out,
                                    hrm_batch_stats* h_stats, hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr, "mapper");
    HRM_REQUIRE(m->d_win_prefix != nullptr, "hrm_mapper_set_genome has not been called");
    HRM_REQUIRE(n >= 0 && n < (1LL << 31) / (m->cfg.num_tables > 0 ? m->cfg.num_tables : 1),
                "batch too large: n * hashmaps must fit int");
    HRM_REQUIRE(ascii_pitch > 0 && ascii_pitch % 16 == 0, "ascii_pitch must be a positive multiple of 16");
    cudaStream_t s = as_stream(stream);
    const hrm_mapper_config& cfg = m->cfg;
    const int64_t launches0 = g_launches.load();
    hrm_batch_stats st;
    memset(&st, 0, sizeof st);
    st.num_reads = n;
    // key-partitioned index: the routed query is collective and bounded per call, so every rank runs the same number
    // of chunks of at most part_chunk reads (a rank that has run out of reads serves the others' lookups with n = 0)
    int64_t nmax_all = n;
    if (m->comm) HRM_TRY(comm_max_i64(m->comm, &nmax_all, s));
    if (n == 0) {
        if (m->comm) {
            Scratch numoff, values;
            HRM_TRY(numoff.alloc(sizeof(int32_t) * 4, s));
            for (int p = 0; p < cfg.num_passes; p++) {
                int64_t chunk = m->part_chunk; // the same sequence of collective queries as the ranks that hold reads
                for (int64_t base = 0; base < nmax_all || base == 0;) {
                    int64_t total = 0;
                    const hrm_status qs = partitioned_query(m->comm, m->index[cfg.genome_conversion[p]], nullptr, 0,
                                                            numoff.as<int32_t>(), numoff.as<int32_t>() + 1, &total, values,
                                                            m->timer, s);
                    if (qs == HRM_ERR_OVERFLOW && chunk > 1024) {
                        chunk /= 2;
                        continue;
                    }
                    HRM_TRY(qs);
                    base += chunk;
                }
            }
            HRM_CUDA(cudaStreamSynchronize(s));
        }
        if (h_stats) *h_stats = st;
        return HRM_OK;
    }
    const int H = cfg.num_tables;
    if (h_stats) // slot-touch counters are reported per call
        for (int c = 0; c < 3; c++)
            if (m->index[c]) {
                HRM_CUDA(cudaMemsetAsync(m->index[c]->d_touches, 0, sizeof(unsigned long long), s));
                m->touches_seen[c] = 0;
            }
    StageTimer& T = m->timer;
    T.begin(HRM_STAGE_PACK, s);
    HRM_TRY(mapper_pack_batch(m, d_reads_ascii, ascii_pitch, d_lengths, n, stream));
    T.end(s);
    HRM_TRY(m->bc->sigs.reserve(sizeof(uint64_t) * (size_t)n * H));
    HRM_TRY(m->bc->num.reserve(sizeof(int32_t) * ((size_t)n + 1)));
    HRM_TRY(m->bc->off.reserve(sizeof(int32_t) * ((size_t)n + 1)));
    HRM_TRY(m->bc->newoff.reserve(sizeof(int32_t) * ((size_t)n + 1)));
    HRM_TRY(m->bc->passres.reserve(sizeof(hrm_mapped_read) * (size_t)n));
    HRM_TRY(m->bc->misc.reserve(64));
    int64_t* d_tot = m->bc->misc.as<int64_t>();           // [0] values total, [1] filtered total
    unsigned long long* d_cnt = m->bc->misc.as<unsigned long long>() + 2;
    int last_rc = -1;
    for (int p = 0; p < cfg.num_passes; p++) {
        const int rc = cfg.read_conversion[p], gc = cfg.genome_conversion[p];
        hrm_minhasher* mh = m->index[gc];
        QueryHandle* qh = minhasher_handle(mh, m->index_handle[gc]);
        HRM_REQUIRE(qh != nullptr, "index handle");
        const uint32_t* reads = m->bc->packed[rc].as<uint32_t>();
        // K2 (skipped when the previous pass sketched the same converted reads)
        if (rc != last_rc) {
            T.begin(HRM_STAGE_MINHASH, s);
            HRM_TRY(minhash_rows(reads, m->bc->packed_pitch, d_lengths, n, cfg.k, H, m->bc->sigs.as<uint64_t>(), nullptr, s));
            T.end(s);
        }
        last_rc = rc;
        hrm_mapped_read* passout = m->bc->passres.as<hrm_mapped_read>();
        // K4 + K5 of reads [lo, lo + cnt): values in table order at `values`, offsets in m->off[0 .. cnt]
        auto filter_and_select = [&](int64_t lo, int64_t cnt, int64_t total, Scratch& values) -> hrm_status {
            Scratch cands;
            HRM_TRY(cands.alloc(sizeof(uint32_t) * (size_t)(total > 0 ? total : 1), s));
            T.begin(HRM_STAGE_FILTER, s);
            // K4 count / sort + threshold, then dense candidate lists (cands doubles as K4's partition space first)
            HRM_TRY(filter_segments(values.as<uint32_t>(), cands.as<uint32_t>(), m->bc->off.as<int32_t>(), (int)cnt,
                                    cfg.min_table_hits, m->bc->num.as<int32_t>() + lo, m->bc->newoff.as<int32_t>(), d_tot + 1, s));
            HRM_TRY(compact_segments(values.as<uint32_t>(), m->bc->off.as<int32_t>(), m->bc->newoff.as<int32_t>(), (int)cnt,
                                     cands.as<uint32_t>(), s));
            T.end(s);
            // K5 + per-read arg-min
            T.begin(HRM_STAGE_SHD, s);
            HRM_TRY(best_windows(reads + lo * m->bc->packed_pitch, m->bc->packed_pitch, d_lengths + lo, cnt, cands.as<uint32_t>(),
                                 m->bc->newoff.as<int32_t>(), m->genome[gc], m->d_win_prefix, cfg.k, cfg.window_size,
                                 cfg.max_hamming_percent, p, passout + lo, s));
            T.end(s);
            st.num_values += total;
            if (h_stats) {
                int64_t ftotal = 0;
                HRM_CUDA(cudaMemcpyAsync(&ftotal, d_tot + 1, sizeof ftotal, cudaMemcpyDeviceToHost, s));
                HRM_CUDA(cudaStreamSynchronize(s));
                st.num_candidates += ftotal;
            }
            return HRM_OK;
        };
        if (m->comm) {
            // key-partitioned index: route the lookups to their owners, values come back in table order.  Every rank
            // runs the same sequence of collective queries: chunk boundaries depend on (nmax, part_chunk) only, and a
            // collective overflow (decided identically on all ranks) halves the chunk size for everyone.
            int64_t chunk = m->part_chunk;
            for (int64_t base = 0; base < nmax_all || base == 0;) {
                const int64_t lo = base < n ? base : n;
                const int64_t cnt = (n - lo) < chunk ? (n - lo) : chunk;
                Scratch values, rng;
                int64_t total = 0;
                const bool fuse = m->use_fused && cfg.min_table_hits >= 2 && m->num_windows < 0xFFFFFFFFLL;
                if (fuse) HRM_TRY(rng.alloc(sizeof(uint2) * (size_t)(cnt > 0 ? cnt : 1) * H, s));
                const hrm_status qs = partitioned_query(m->comm, mh, m->bc->sigs.as<uint64_t>() + lo * H, (int)cnt,
                                                        m->bc->num.as<int32_t>() + lo, m->bc->off.as<int32_t>(), &total, values, T, s,
                                                        fuse ? rng.as<uint2>() : nullptr);
                if (qs == HRM_ERR_OVERFLOW && chunk > 1024) {
                    chunk /= 2; // collective decision: every rank retries this base with half the chunk
                    continue;
                }
                HRM_TRY(qs);
                base += chunk;
                if (cnt == 0) continue;
                bool collected = false;
                if (fuse) { // the fused collection over the routed value lists (k4_fused.cu)
                    Scratch cands, lists;
                    const int64_t cap = cnt * 32 > (1 << 16) ? cnt * 32 : (1 << 16);
                    HRM_TRY(cands.alloc(sizeof(uint32_t) * (size_t)cap, s));
                    HRM_TRY(lists.alloc(sizeof(int2) * (size_t)cnt, s));
                    int64_t ctotal = 0, cst[3] = {0, 0, 0};
                    int overflow = 0;
                    T.begin(HRM_STAGE_FILTER, s);
                    HRM_TRY(collect_candidates_from(rng.as<uint2>(), H, 1, values.as<uint32_t>(), H, (int)cnt, cfg.min_table_hits,
                                                    (uint32_t)m->num_windows, cands.as<uint32_t>(), cap, lists.as<int2>(),
                                                    &ctotal, &overflow, cst, s));
                    T.end(s);
                    if (!overflow) {
                        T.begin(HRM_STAGE_SHD, s);
                        HRM_TRY(best_windows(reads + lo * m->bc->packed_pitch, m->bc->packed_pitch, d_lengths + lo, cnt,
                                             cands.as<uint32_t>(), nullptr, m->genome[gc], m->d_win_prefix, cfg.k,
                                             cfg.window_size, cfg.max_hamming_percent, p, passout + lo, s, lists.as<int2>()));
                        T.end(s);
                        st.num_values += cst[0] + cst[1];
                        st.num_candidates += ctotal;
                        m->collect_enumerated += cst[0];
                        m->collect_skipped += cst[1];
                        m->collect_block_reads += cst[2];
                        collected = true;
                    }
                }
                if (!collected) HRM_TRY(filter_and_select(lo, cnt, total, values));
            }
        } else {
            // K3b probe of the whole batch -> per-read counts and bucket ranges
            if (minhasher_wants_table_major(mh)) {
                // index far larger than L2: probe table by table (transpose + totals are booked under "scan")
                T.begin(HRM_STAGE_SCAN, s);
                HRM_TRY(minhasher_tm_prepare(mh, qh, m->bc->sigs.as<uint64_t>(), (int)n, s));
                T.end(s);
                T.begin(HRM_STAGE_PROBE, s);
                HRM_TRY(minhasher_tm_probe(mh, qh, (int)n, s));
                T.end(s);
                T.begin(HRM_STAGE_SCAN, s);
                HRM_TRY(minhasher_tm_totals(mh, qh, (int)n, m->bc->num.as<int32_t>(), s));
                T.end(s);
            } else {
                T.begin(HRM_STAGE_PROBE, s);
                HRM_TRY(minhasher_count_sigs(mh, qh, m->bc->sigs.as<uint64_t>(), (int)n, m->bc->num.as<int32_t>(), s));
                T.end(s);
            }
            // fused retrieval + collection straight from the index (k4_fused.cu); falls through to the general
            // path when a read or the output does not fit its bounds
            bool collected = false;
            if (m->use_fused && cfg.min_table_hits >= 2 && m->num_windows < 0xFFFFFFFFLL) {
                Scratch cands, lists;
                const int64_t cap = n * 32 > (1 << 16) ? n * 32 : (1 << 16);
                HRM_REQUIRE(cap < (1LL << 31), "batch too large for int candidate offsets");
                HRM_TRY(cands.alloc(sizeof(uint32_t) * (size_t)cap, s));
                HRM_TRY(lists.alloc(sizeof(int2) * (size_t)n, s));
                int64_t ctotal = 0, cst[3] = {0, 0, 0};
                int overflow = 0;
                T.begin(HRM_STAGE_FILTER, s);
                HRM_TRY(collect_candidates(mh, qh, (int)n, cfg.min_table_hits, (uint32_t)m->num_windows, cands.as<uint32_t>(),
                                           cap, lists.as<int2>(), &ctotal, &overflow, cst, s));
                T.end(s);
                if (!overflow) {
                    T.begin(HRM_STAGE_SHD, s);
                    HRM_TRY(best_windows(reads, m->bc->packed_pitch, d_lengths, n, cands.as<uint32_t>(), nullptr, m->genome[gc],
                                         m->d_win_prefix, cfg.k, cfg.window_size, cfg.max_hamming_percent, p, passout, s,
                                         lists.as<int2>()));
                    T.end(s);
                    st.num_values += cst[0] + cst[1];
                    st.num_candidates += ctotal;
                    m->collect_enumerated += cst[0];
                    m->collect_skipped += cst[1];
                    m->collect_block_reads += cst[2];
                    collected = true;
                }
            }
            // ranges of reads whose candidate values fit the budget (int offsets, bounded scratch): scan -> offsets +
            // total (one host sync per range: sizes the buffers), retrieve, K4, K5.  A batch normally is one range;
            // at human-genome scale (thousands of values per read) it is halved until the ranges fit.
            struct Range {
                int64_t lo, cnt;
            };
            std::vector<Range> todo;
            if (!collected) todo.push_back({0, n});
            while (!todo.empty()) {
                const Range r = todo.back();
                todo.pop_back();
                int64_t total = 0;
                T.begin(HRM_STAGE_SCAN, s);
                HRM_TRY(exclusive_scan_i32(m->bc->num.as<int32_t>() + r.lo, m->bc->off.as<int32_t>(), r.cnt, d_tot, s));
                HRM_CUDA(cudaMemcpyAsync(&total, d_tot, sizeof total, cudaMemcpyDeviceToHost, s));
                T.end(s);
                HRM_CUDA(cudaStreamSynchronize(s));
                if (total > m->value_budget) {
                    if (r.cnt == 1) {
                        set_error("one read retrieves more candidate values than the value budget");
                        return HRM_ERR_OVERFLOW;
                    }
                    const int64_t half = r.cnt / 2;
                    todo.push_back({r.lo + half, r.cnt - half}); // processed second
                    todo.push_back({r.lo, half});
                    continue;
                }
                Scratch values;
                HRM_TRY(values.alloc(sizeof(uint32_t) * (size_t)(total > 0 ? total : 1), s));
                T.begin(HRM_STAGE_RETRIEVE, s);
                if (total > 0)
                    HRM_TRY(minhasher_retrieve(mh, qh, r.lo, (int)r.cnt, values.as<uint32_t>(), m->bc->off.as<int32_t>(), s));
                T.end(s);
                HRM_TRY(filter_and_select(r.lo, r.cnt, total, values));
            }
        }
        st.num_probes += n * H;
        qh->stage = 0;
        T.begin(HRM_STAGE_MERGE, s);
        HRM_LAUNCH(merge_pass_kernel, mgrid(n), 256, 0, s, d_out, passout, n, p == 0 ? 1 : 0);
        T.end(s);
    }
    if (h_stats) {
        HRM_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
        HRM_LAUNCH(count_mapped_kernel, mgrid(n), 256, 0, s, d_out, n, d_cnt);
        unsigned long long mapped = 0;
        HRM_CUDA(cudaMemcpyAsync(&mapped, d_cnt, sizeof mapped, cudaMemcpyDeviceToHost, s));
        for (int c = 0; c < 3; c++) {
            if (!m->index[c]) continue;
            unsigned long long t = 0;
            HRM_CUDA(cudaMemcpyAsync(&t, m->index[c]->d_touches, sizeof t, cudaMemcpyDeviceToHost, s));
            HRM_CUDA(cudaStreamSynchronize(s));
            st.num_slot_touches += (int64_t)(t - m->touches_seen[c]);
            m->touches_seen[c] = t;
        }
        HRM_CUDA(cudaStreamSynchronize(s));
        st.num_mapped = (int64_t)mapped;
        st.num_kernel_launches = g_launches.load() - launches0;
        *h_stats = st;
    }
    return HRM_OK;
}

// verification of the batch whose reads are packed in m->bc.  maxlen < 0: taken from the device (one host sync);
// d_reads_ascii != NULL: pack first.
static hrm_status verify_batch_impl(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch, const int32_t* d_lengths,
                                    int64_t n, int maxlen, const hrm_mapped_read* d_mapped, hrm_read_record* d_records,
                                    char* d_cigars, int64_t cigar_pitch, hrm_batch_stats* h_stats, hrm_stream stream)
{
    cudaStream_t s = as_stream(stream);
    const int64_t launches0 = g_launches.load();
    if (n == 0) return HRM_OK;
    const hrm_mapper_config& cfg = m->cfg;
    m->timer.begin(HRM_STAGE_VERIFY, s);
    if (d_reads_ascii) HRM_TRY(mapper_pack_batch(m, d_reads_ascii, ascii_pitch, d_lengths, n, stream));
    if (maxlen < 0) {
        HRM_TRY(m->bc->misc.reserve(64));
        int* d_maxlen = m->bc->misc.as<int>() + 8;
        HRM_CUDA(cudaMemsetAsync(d_maxlen, 0, sizeof(int), s));
        HRM_LAUNCH(max_len_kernel, mgrid(n), 256, 0, s, d_lengths, n, d_maxlen);
        HRM_CUDA(cudaMemcpyAsync(&maxlen, d_maxlen, sizeof maxlen, cudaMemcpyDeviceToHost, s));
        HRM_CUDA(cudaStreamSynchronize(s));
    }
    HRM_REQUIRE(maxlen <= HRM_SW_MAX_QUERY, "read longer than HRM_SW_MAX_QUERY");
    VerifyParams VP;
    memset(&VP, 0, sizeof VP);
    VP.num_passes = cfg.num_passes;
    VP.w = cfg.window_size;
    VP.mapper_type = cfg.mapper_type;
    for (int p = 0; p < cfg.num_passes; p++) {
        VP.pass[p].reads = m->bc->packed[cfg.read_conversion[p]].as<uint32_t>();
        VP.pass[p].read_pitch = m->bc->packed_pitch;
        VP.pass[p].G = m->genome[cfg.genome_conversion[p]]->dev();
        VP.pass[p].verify_conv = cfg.verify_conversion[p];
    }
    HRM_TRY(verify_reads(VP, d_lengths, n, maxlen, d_mapped, d_records, d_cigars, cigar_pitch, s));
    m->timer.end(s);
    if (h_stats) h_stats->num_kernel_launches += g_launches.load() - launches0;
    return HRM_OK;
}

extern "C" hrm_status hrm_verify_batch(hrm_mapper* m, const char* d_reads_ascii, int64_t ascii_pitch,
                                       const int32_t* d_lengths, int64_t n, const hrm_mapped_read* d_mapped,
                                       hrm_read_record* d_records, char* d_cigars, int64_t cigar_pitch,
                                       hrm_batch_stats* h_stats, hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr, "mapper");
    HRM_REQUIRE(m->d_win_prefix != nullptr, "hrm_mapper_set_genome has not been called");
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0 && cigar_pitch > 0, "sizes");
    return verify_batch_impl(m, d_reads_ascii, ascii_pitch, d_lengths, n, -1, d_mapped, d_records, d_cigars, cigar_pitch,
                             h_stats, stream);
}

extern "C" hrm_status hrm_mapper_map_reads(hrm_mapper* m, const char* h_reads_ascii, int64_t ascii_pitch,
                                           const int32_t* h_lengths, int64_t n, hrm_read_record* h_records,
                                           char* h_cigars, int64_t cigar_pitch, hrm_batch_stats* h_stats,
                                           hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr && h_reads_ascii != nullptr && h_lengths != nullptr && h_records != nullptr, "args");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0 && cigar_pitch > 0, "sizes");
    cudaStream_t s = as_stream(stream);
    if (n == 0) {
        if (h_stats) memset(h_stats, 0, sizeof *h_stats);
        // key-partitioned index: the call is collective, a rank without reads still serves the others' lookups
        if (m->comm) return hrm_map_batch(m, nullptr, ascii_pitch, nullptr, 0, nullptr, nullptr, stream);
        return HRM_OK;
    }
    Scratch d_ascii, d_len, d_mapped, d_rec, d_cig;
    HRM_TRY(d_ascii.alloc((size_t)(n * ascii_pitch), s));
    HRM_TRY(d_len.alloc(sizeof(int32_t) * (size_t)n, s));
    HRM_TRY(d_mapped.alloc(sizeof(hrm_mapped_read) * (size_t)n, s));
    HRM_TRY(d_rec.alloc(sizeof(hrm_read_record) * (size_t)n, s));
    HRM_TRY(d_cig.alloc((size_t)(2 * n * cigar_pitch), s));
    hrm_batch_stats st;
    memset(&st, 0, sizeof st);
    // Chunks of reads pipelined over a second stream: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 run
    // under the kernels of chunk c (reads are independent, so chunking does not change any result).  The
    // key-partitioned index keeps one chunk: its routed queries are collective and every rank must issue the same
    // sequence of them.
    const int64_t chunk = m->comm ? n : m->e2e_chunk;
    const int nchunks = (int)HRM_SDIV(n, chunk);
    if (!m->copy_stream) HRM_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = m->copy_stream;
    while ((int)m->copy_events.size() < 2 * nchunks + 1) {
        cudaEvent_t e;
        HRM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->copy_events.push_back(e);
    }
    cudaEvent_t ev_ready = m->copy_events[2 * nchunks]; // the device buffers exist (allocated in order on s)
    HRM_CUDA(cudaEventRecord(ev_ready, s));
    HRM_CUDA(cudaStreamWaitEvent(cs, ev_ready, 0));
    for (int c = 0; c < nchunks; c++) {
        const int64_t lo = (int64_t)c * chunk, cnt = (n - lo) < chunk ? (n - lo) : chunk;
        HRM_CUDA(cudaMemcpyAsync(d_ascii.as<char>() + lo * ascii_pitch, h_reads_ascii + lo * ascii_pitch,
                                 (size_t)(cnt * ascii_pitch), cudaMemcpyHostToDevice, cs));
        HRM_CUDA(cudaMemcpyAsync(d_len.as<int32_t>() + lo, h_lengths + lo, sizeof(int32_t) * (size_t)cnt,
                                 cudaMemcpyHostToDevice, cs));
        HRM_CUDA(cudaEventRecord(m->copy_events[2 * c], cs));
    }
    // on an error nothing may stay in flight on the copy stream when the scratch buffers die (they are freed on `s`)
    struct Drain {
        cudaStream_t a, b;
        bool armed = true;
        ~Drain()
        {
            if (armed) {
                cudaStreamSynchronize(a);
                cudaStreamSynchronize(b);
            }
        }
    } drain{cs, s};
    for (int c = 0; c < nchunks; c++) {
        const int64_t lo = (int64_t)c * chunk, cnt = (n - lo) < chunk ? (n - lo) : chunk;
        HRM_CUDA(cudaStreamWaitEvent(s, m->copy_events[2 * c], 0));
        hrm_batch_stats cst;
        memset(&cst, 0, sizeof cst);
        HRM_TRY(hrm_map_batch(m, d_ascii.as<char>() + lo * ascii_pitch, ascii_pitch, d_len.as<int32_t>() + lo, cnt,
                              d_mapped.as<hrm_mapped_read>() + lo, h_stats ? &cst : nullptr, stream));
        HRM_TRY(hrm_verify_batch(m, d_ascii.as<char>() + lo * ascii_pitch, ascii_pitch, d_len.as<int32_t>() + lo, cnt,
                                 d_mapped.as<hrm_mapped_read>() + lo, d_rec.as<hrm_read_record>() + lo,
                                 d_cig.as<char>() + 2 * lo * cigar_pitch, cigar_pitch, h_stats ? &cst : nullptr, stream));
        st.num_reads += cst.num_reads;
        st.num_probes += cst.num_probes;
        st.num_slot_touches += cst.num_slot_touches;
        st.num_values += cst.num_values;
        st.num_candidates += cst.num_candidates;
        st.num_mapped += cst.num_mapped;
        st.num_kernel_launches += cst.num_kernel_launches;
        HRM_CUDA(cudaEventRecord(m->copy_events[2 * c + 1], s));
        HRM_CUDA(cudaStreamWaitEvent(cs, m->copy_events[2 * c + 1], 0));
        HRM_CUDA(cudaMemcpyAsync(h_records + lo, d_rec.as<hrm_read_record>() + lo, sizeof(hrm_read_record) * (size_t)cnt,
                                 cudaMemcpyDeviceToHost, cs));
        if (h_cigars)
            HRM_CUDA(cudaMemcpyAsync(h_cigars + 2 * lo * cigar_pitch, d_cig.as<char>() + 2 * lo * cigar_pitch,
                                     (size_t)(2 * cnt * cigar_pitch), cudaMemcpyDeviceToHost, cs));
    }
    HRM_CUDA(cudaStreamSynchronize(cs));
    HRM_CUDA(cudaStreamSynchronize(s));
    drain.armed = false;
    if (h_stats) *h_stats = st;
    return HRM_OK;
}

// End to end with text out: reads H2D -> K1..K5 -> K7 -> V4 -> SAM text -> D2H (ref: performMappingGpu STEP 1 + STEP 2
// incl. printtoSAM, main_gpu.cu:1123-1160).  The text buffers are sized by an upper bound per read so that the only
// host round trip of the output side is the one that returns the byte counts.
extern "C" hrm_status hrm_mapper_map_reads_sam(hrm_mapper* m, const char* h_reads_ascii, int64_t ascii_pitch,
                                               const int32_t* h_lengths, int64_t n, uint32_t first_read_id,
                                               const char* const* h_chrom_names, char* h_sq_out, int64_t sq_cap,
                                               int64_t* h_sq_written, char* h_rec_out, int64_t rec_cap,
                                               int64_t* h_rec_written, hrm_read_record* h_records, char* h_cigars,
                                               int64_t cigar_pitch, hrm_batch_stats* h_stats, hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr && h_reads_ascii != nullptr && h_lengths != nullptr && h_rec_written != nullptr, "args");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0 && cigar_pitch > 0, "sizes");
    HRM_REQUIRE(m->comm == nullptr, "hrm_mapper_map_reads_sam runs on the replicated index");
    cudaStream_t s = as_stream(stream);
    *h_rec_written = 0;
    if (h_sq_written) *h_sq_written = 0;
    if (h_stats) memset(h_stats, 0, sizeof *h_stats);
    if (n == 0) return HRM_OK;
    size_t maxname = 12;
    for (int c = 0; c < m->n_chrom && h_chrom_names; c++)
        if (h_chrom_names[c] && strlen(h_chrom_names[c]) > maxname) maxname = strlen(h_chrom_names[c]);
    const int64_t line_bound = 64 + (int64_t)maxname + cigar_pitch + m->cfg.window_size + ascii_pitch;
    Scratch d_ascii, d_len, d_mapped, d_rec, d_cig, d_fields, d_ll, d_text, d_sq;
    HRM_TRY(d_ascii.alloc((size_t)(n * ascii_pitch), s));
    HRM_TRY(d_len.alloc(sizeof(int32_t) * (size_t)n, s));
    HRM_TRY(d_mapped.alloc(sizeof(hrm_mapped_read) * (size_t)n, s));
    HRM_TRY(d_rec.alloc(sizeof(hrm_read_record) * (size_t)n, s));
    HRM_TRY(d_cig.alloc((size_t)(2 * n * cigar_pitch), s));
    HRM_TRY(d_fields.alloc(sizeof(hrm_sam_fields) * (size_t)n, s));
    HRM_TRY(d_ll.alloc(sizeof(int32_t) * (size_t)(2 * n), s));
    HRM_TRY(d_text.alloc((size_t)(n * line_bound), s));
    if (h_sq_out) HRM_TRY(d_sq.alloc((size_t)(n * 40), s));
    HRM_CUDA(cudaMemcpyAsync(d_ascii.p, h_reads_ascii, (size_t)(n * ascii_pitch), cudaMemcpyHostToDevice, s));
    HRM_CUDA(cudaMemcpyAsync(d_len.p, h_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
    hrm_batch_stats st;
    memset(&st, 0, sizeof st);
    HRM_TRY(hrm_map_batch(m, d_ascii.as<char>(), ascii_pitch, d_len.as<int32_t>(), n, d_mapped.as<hrm_mapped_read>(),
                          h_stats ? &st : nullptr, stream));
    HRM_TRY(hrm_verify_batch(m, d_ascii.as<char>(), ascii_pitch, d_len.as<int32_t>(), n, d_mapped.as<hrm_mapped_read>(),
                             d_rec.as<hrm_read_record>(), d_cig.as<char>(), cigar_pitch, h_stats ? &st : nullptr, stream));
    // the reads of the batch are still packed in the mapper (hrm_verify_batch packed them)
    const int64_t launches0 = g_launches.load();
    HRM_TRY(hrm_sam_format_device(m, nullptr, ascii_pitch, d_len.as<int32_t>(), n, d_rec.as<hrm_read_record>(),
                                  d_cig.as<char>(), cigar_pitch, first_read_id, h_chrom_names, HRM_SAM_RECORDS,
                                  d_text.as<char>(), n * line_bound, h_rec_written, stream));
    if (h_rec_out) {
        HRM_REQUIRE(*h_rec_written <= rec_cap, "record text does not fit rec_cap");
        HRM_CUDA(cudaMemcpyAsync(h_rec_out, d_text.p, (size_t)*h_rec_written, cudaMemcpyDeviceToHost, s));
    }
    if (h_sq_out && h_sq_written) {
        HRM_TRY(hrm_sam_format_device(m, nullptr, ascii_pitch, d_len.as<int32_t>(), n, d_rec.as<hrm_read_record>(),
                                      d_cig.as<char>(), cigar_pitch, first_read_id, h_chrom_names, HRM_SAM_SQ_LINES,
                                      d_sq.as<char>(), n * 40, h_sq_written, stream));
        HRM_REQUIRE(*h_sq_written <= sq_cap, "@SQ text does not fit sq_cap");
        HRM_CUDA(cudaMemcpyAsync(h_sq_out, d_sq.p, (size_t)*h_sq_written, cudaMemcpyDeviceToHost, s));
    }
    if (h_records)
        HRM_CUDA(cudaMemcpyAsync(h_records, d_rec.p, sizeof(hrm_read_record) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (h_cigars) HRM_CUDA(cudaMemcpyAsync(h_cigars, d_cig.p, (size_t)(2 * n * cigar_pitch), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    if (h_stats) {
        st.num_kernel_launches += g_launches.load() - launches0;
        *h_stats = st;
    }
    return HRM_OK;
}


// ---- double-buffered end-to-end pipeline ---------------------------------------------------------------------
// ref: the reference's driver overlaps nothing: every window batch is H2D -> kernels -> D2H with >= 6 stream syncs
// (main_gpu.cu:471-854) and STEP 2 starts when STEP 1 has ended (main_gpu.cu:1123-1160).  Here a batch lives in one of
// HRM_PIPE_SLOTS slots: its reads are staged (H2D on a copy-in stream) while the previous batch computes, its results
// leave (D2H on the copy-out stream of its slot) while the next batch computes; the kernels of all batches run in order on the
// caller's stream.  Host syncs inside the compute (candidate totals, text size) only ever wait for the batch at hand.
static hrm_status pipe_init(hrm_mapper* m)
{
    static std::mutex mtx; // staging may be called from a second host thread
    std::lock_guard<std::mutex> lk(mtx);
    if (m->pipe_ready) return HRM_OK;
    HRM_CUDA(cudaStreamCreateWithFlags(&m->pipe_in, cudaStreamNonBlocking));
    for (int i = 0; i < HRM_PIPE_SLOTS; i++) HRM_CUDA(cudaStreamCreateWithFlags(&m->pipe_out[i], cudaStreamNonBlocking));
    HRM_CUDA(cudaStreamCreateWithFlags(&m->pipe_verify, cudaStreamNonBlocking));
    HRM_CUDA(cudaMallocHost(&m->pipe_host, sizeof(int64_t) * 4 * HRM_PIPE_SLOTS));
    for (int i = 0; i < HRM_PIPE_SLOTS; i++) {
        HRM_CUDA(cudaEventCreateWithFlags(&m->slot[i].staged, cudaEventDisableTiming));
        HRM_CUDA(cudaEventCreateWithFlags(&m->slot[i].seeded, cudaEventDisableTiming));
        HRM_CUDA(cudaEventCreateWithFlags(&m->slot[i].computed, cudaEventDisableTiming));
        HRM_CUDA(cudaEventCreateWithFlags(&m->slot[i].drained, cudaEventDisableTiming));
    }
    m->pipe_ready = true;
    return HRM_OK;
}

// the slot is free again when everything its previous batch queued has finished
static hrm_status slot_release(hrm_mapper* m, hrm::PipeSlot& S)
{
    if (S.busy && S.want_text) {
        set_error("the slot holds a batch whose SAM text has not been fetched: call hrm_mapper_finish first");
        return HRM_ERR_STATE;
    }
    if (S.busy) {
        HRM_CUDA(cudaEventSynchronize(S.computed));
        HRM_CUDA(cudaEventSynchronize(S.drained));
    }
    S.busy = false;
    (void)m;
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_stage_reads(hrm_mapper* m, int slot, const char* h_reads_ascii, int64_t ascii_pitch,
                                             const int32_t* h_lengths, int64_t n)
{
    HRM_REQUIRE(m != nullptr && slot >= 0 && slot < HRM_PIPE_SLOTS, "mapper / slot");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0, "sizes");
    HRM_REQUIRE(n == 0 || (h_reads_ascii != nullptr && h_lengths != nullptr), "buffers");
    HRM_REQUIRE(m->comm == nullptr, "the staged pipeline runs on the replicated index");
    HRM_TRY(pipe_init(m));
    PipeSlot& S = m->slot[slot];
    HRM_TRY(slot_release(m, S)); // the slot's previous batch must have left the device before its buffers are overwritten
    S.n = n;
    S.pitch = ascii_pitch;
    S.d_ascii = nullptr;
    int mx = 0;
    for (int64_t i = 0; i < n; i++) mx = h_lengths[i] > mx ? h_lengths[i] : mx; // sizes the SW frames without a device sync
    S.maxlen = mx;
    HRM_TRY(S.ascii.reserve((size_t)(n * ascii_pitch)));
    HRM_TRY(S.len.reserve(sizeof(int32_t) * (size_t)n));
    if (n > 0) {
        HRM_CUDA(cudaMemcpyAsync(S.ascii.p, h_reads_ascii, (size_t)(n * ascii_pitch), cudaMemcpyHostToDevice, m->pipe_in));
        HRM_CUDA(cudaMemcpyAsync(S.len.p, h_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, m->pipe_in));
    }
    HRM_CUDA(cudaEventRecord(S.staged, m->pipe_in));
    S.is_staged = true;
    return HRM_OK;
}

// reads that are already in device memory (they must stay valid until hrm_mapper_finish of the slot)
extern "C" hrm_status hrm_mapper_stage_device(hrm_mapper* m, int slot, const char* d_reads_ascii, int64_t ascii_pitch,
                                              const int32_t* d_lengths, int64_t n, int max_length, hrm_stream ready_on)
{
    HRM_REQUIRE(m != nullptr && slot >= 0 && slot < HRM_PIPE_SLOTS, "mapper / slot");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0 && max_length >= 0 && max_length <= ascii_pitch, "sizes");
    HRM_REQUIRE(n == 0 || (d_reads_ascii != nullptr && d_lengths != nullptr), "buffers");
    HRM_REQUIRE(m->comm == nullptr, "the staged pipeline runs on the replicated index");
    HRM_TRY(pipe_init(m));
    PipeSlot& S = m->slot[slot];
    HRM_TRY(slot_release(m, S));
    S.n = n;
    S.pitch = ascii_pitch;
    S.d_ascii = d_reads_ascii;
    S.d_len = d_lengths;
    S.maxlen = max_length;
    HRM_CUDA(cudaEventRecord(S.staged, as_stream(ready_on)));
    S.is_staged = true;
    return HRM_OK;
}

// FASTQ / FASTA text instead of ASCII rows: the text goes H2D on the copy-in stream and is parsed there by
// hrm_ingest_reads (device-side reader, ingest.cu).  Blocks the calling thread until the batch is parsed (the reader
// sizes its buffers from the line count); call it from a second host thread to overlap it with hrm_mapper_map_staged
// of the other slot -- the two calls share no state.
extern "C" hrm_status hrm_mapper_stage_fastq(hrm_mapper* m, int slot, const char* h_text, int64_t nbytes,
                                             int64_t first_read_id, int32_t carry_replaced, int64_t ascii_pitch,
                                             int64_t max_reads, int64_t* h_num_reads, int32_t* h_carry_replaced_out)
{
    HRM_REQUIRE(m != nullptr && slot >= 0 && slot < HRM_PIPE_SLOTS, "mapper / slot");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(nbytes >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0 && max_reads >= 0 && h_num_reads != nullptr, "sizes");
    HRM_REQUIRE(nbytes == 0 || h_text != nullptr, "text");
    HRM_REQUIRE(m->comm == nullptr, "the staged pipeline runs on the replicated index");
    HRM_TRY(pipe_init(m));
    PipeSlot& S = m->slot[slot];
    HRM_TRY(slot_release(m, S));
    S.pitch = ascii_pitch;
    S.n = 0;
    S.d_ascii = nullptr;
    S.maxlen = (int)(ascii_pitch < HRM_SW_MAX_QUERY ? ascii_pitch : HRM_SW_MAX_QUERY);
    HRM_TRY(S.fastq.reserve((size_t)nbytes + 16));
    HRM_TRY(S.ascii.reserve((size_t)(max_reads * ascii_pitch)));
    HRM_TRY(S.len.reserve(sizeof(int32_t) * (size_t)max_reads));
    if (nbytes > 0) HRM_CUDA(cudaMemcpyAsync(S.fastq.p, h_text, (size_t)nbytes, cudaMemcpyHostToDevice, m->pipe_in));
    HRM_TRY(hrm_ingest_reads(S.fastq.as<char>(), nbytes, first_read_id, carry_replaced, S.ascii.as<char>(), ascii_pitch,
                             S.len.as<int32_t>(), nullptr, max_reads, h_num_reads, h_carry_replaced_out,
                             (hrm_stream)m->pipe_in));
    S.n = *h_num_reads;
    HRM_CUDA(cudaEventRecord(S.staged, m->pipe_in));
    S.is_staged = true;
    return HRM_OK;
}

// Seeding (K1..K5) of the staged batch, then its verification (K7 / K6), V4 and the SAM text, all queued without a host
// round trip after the seeding; the D2H copies on the slot's copy-out stream.  Returns when the seeding kernels have run (their
// candidate totals come back to the host) and everything else is queued.
extern "C" hrm_status hrm_mapper_map_staged(hrm_mapper* m, int slot, hrm_read_record* h_records, char* h_cigars,
                                            int64_t cigar_pitch, uint32_t first_read_id,
                                            const char* const* h_chrom_names, char* h_sq_out, int64_t sq_cap,
                                            char* h_rec_out, int64_t rec_cap, hrm_batch_stats* h_stats, hrm_stream stream)
{
    HRM_REQUIRE(m != nullptr && slot >= 0 && slot < HRM_PIPE_SLOTS, "mapper / slot");
    HRM_TRY(bind_device(m));
    HRM_REQUIRE(cigar_pitch > 0, "cigar_pitch");
    PipeSlot& S = m->slot[slot];
    HRM_REQUIRE(m->pipe_ready && S.is_staged, "nothing is staged in this slot");
    // verification on the caller's stream by default.  HRM_PIPE_OVERLAP=1 puts it on a second stream so that it runs under
    // the seeding of the next batch; measured on B200 (profiles/README.md) the two do not speed each other up -- each kernel
    // fills the SMs with persistent blocks and the pair is bound by instruction issue -- so it is off by default.
    static const bool overlap = getenv("HRM_PIPE_OVERLAP") != nullptr && atoi(getenv("HRM_PIPE_OVERLAP")) != 0;
    cudaStream_t s = as_stream(stream), vs = overlap ? m->pipe_verify : s, co = m->pipe_out[slot];
    const int64_t n = S.n;
    S.is_staged = false;
    S.sq_written = S.rec_written = 0;
    S.h_sq = S.h_rec = nullptr;
    S.want_text = false;
    if (h_stats) memset(h_stats, 0, sizeof *h_stats);
    if (n == 0) return HRM_OK;
    const bool want_text = h_rec_out != nullptr;
    size_t maxname = 12;
    for (int c = 0; c < m->n_chrom && h_chrom_names; c++)
        if (h_chrom_names[c] && strlen(h_chrom_names[c]) > maxname) maxname = strlen(h_chrom_names[c]);
    const int64_t line_bound = 64 + (int64_t)maxname + cigar_pitch + m->cfg.window_size + S.pitch;
    HRM_REQUIRE(!want_text || (n * line_bound <= rec_cap && (!h_sq_out || n * 40 <= sq_cap)),
                "text buffers must hold n * (64 + longest chromosome name + cigar_pitch + window + pitch) / n * 40 bytes");
    HRM_TRY(S.mapped.reserve(sizeof(hrm_mapped_read) * (size_t)n));
    HRM_TRY(S.rec.reserve(sizeof(hrm_read_record) * (size_t)n));
    HRM_TRY(S.cig.reserve((size_t)(2 * n * cigar_pitch)));
    if (want_text) {
        HRM_TRY(S.text.reserve((size_t)(n * line_bound)));
        HRM_TRY(S.fields.reserve(sizeof(hrm_sam_fields) * (size_t)n));
        HRM_TRY(S.lens2.reserve(sizeof(int32_t) * (size_t)(2 * n)));
        HRM_TRY(S.offs.reserve(sizeof(int64_t) * (size_t)(2 * n + 4)));
        if (h_sq_out) HRM_TRY(S.sq.reserve((size_t)(n * 40)));
        HRM_TRY(sam_upload_names(m, h_chrom_names, s));
    }
    hrm_batch_stats st;
    memset(&st, 0, sizeof st);
    const char* d_ascii = S.d_ascii ? S.d_ascii : S.ascii.as<char>();
    const int32_t* d_len = S.d_ascii ? S.d_len : S.len.as<int32_t>();
    // seeding on the caller's stream, in this slot's context (its packed reads stay valid for the verification)
    m->bc = &m->ctx[slot];
    HRM_CUDA(cudaStreamWaitEvent(s, S.staged, 0));
    hrm_status rc = hrm_map_batch(m, d_ascii, S.pitch, d_len, n, S.mapped.as<hrm_mapped_read>(), h_stats ? &st : nullptr, stream);
    if (rc == HRM_OK) rc = cudaEventRecord(S.seeded, s) == cudaSuccess ? HRM_OK : HRM_ERR_CUDA;
    // verification + V4 + SAM text on the second stream
    const int64_t launches0 = g_launches.load();
    if (rc == HRM_OK) rc = cudaStreamWaitEvent(vs, S.seeded, 0) == cudaSuccess ? HRM_OK : HRM_ERR_CUDA;
    if (rc == HRM_OK)
        rc = verify_batch_impl(m, nullptr, S.pitch, d_len, n, S.maxlen, S.mapped.as<hrm_mapped_read>(),
                               S.rec.as<hrm_read_record>(), S.cig.as<char>(), cigar_pitch, nullptr, (hrm_stream)vs);
    int64_t* hp = m->pipe_host + 4 * slot; // pinned: [0] record text bytes, [1] @SQ text bytes
    if (rc == HRM_OK && want_text) {
        int32_t* ll = S.lens2.as<int32_t>();
        int64_t* off = S.offs.as<int64_t>();
        rc = sam_fields(m, d_len, n, S.rec.as<hrm_read_record>(), S.cig.as<char>(), cigar_pitch, first_read_id,
                        S.fields.as<hrm_sam_fields>(), ll, ll + n, vs);
        if (rc == HRM_OK)
            rc = sam_text_async(m, d_len, n, S.rec.as<hrm_read_record>(), S.cig.as<char>(), cigar_pitch,
                                S.fields.as<hrm_sam_fields>(), ll, first_read_id, HRM_SAM_RECORDS, S.text.as<char>(),
                                n * line_bound, off, hp, vs);
        if (rc == HRM_OK && h_sq_out)
            rc = sam_text_async(m, d_len, n, S.rec.as<hrm_read_record>(), S.cig.as<char>(), cigar_pitch,
                                S.fields.as<hrm_sam_fields>(), ll + n, first_read_id, HRM_SAM_SQ_LINES, S.sq.as<char>(),
                                n * 40, off + n + 2, hp + 1, vs);
        S.want_text = true;
        S.h_rec = h_rec_out;
        S.h_sq = h_sq_out;
    }
    st.num_kernel_launches += g_launches.load() - launches0;
    m->bc = &m->ctx[HRM_PIPE_SLOTS];
    if (rc == HRM_OK) rc = cudaEventRecord(S.computed, vs) == cudaSuccess ? HRM_OK : HRM_ERR_CUDA;
    S.busy = true; // something is queued on the slot's buffers
    // the next batch's seeding reuses the other context only; this context is reused two batches on, after `computed`
    if (rc == HRM_OK) rc = cudaStreamWaitEvent(co, S.computed, 0) == cudaSuccess ? HRM_OK : HRM_ERR_CUDA;
    if (rc == HRM_OK && h_records)
        rc = cudaMemcpyAsync(h_records, S.rec.p, sizeof(hrm_read_record) * (size_t)n, cudaMemcpyDeviceToHost, co) == cudaSuccess
                 ? HRM_OK : HRM_ERR_CUDA;
    if (rc == HRM_OK && h_cigars)
        rc = cudaMemcpyAsync(h_cigars, S.cig.p, (size_t)(2 * n * cigar_pitch), cudaMemcpyDeviceToHost, co) == cudaSuccess
                 ? HRM_OK : HRM_ERR_CUDA;
    cudaEventRecord(S.drained, co);
    // the other context must not be overwritten by the NEXT seeding before ITS verification has finished: make the
    // caller's stream wait for the other slot's `computed` (queued two calls ago)
    PipeSlot& O = m->slot[(slot + 1) % HRM_PIPE_SLOTS];
    if (O.busy) cudaStreamWaitEvent(s, O.computed, 0);
    if (rc != HRM_OK) {
        if (rc == HRM_ERR_CUDA) set_error("CUDA call failed while queueing the staged batch: %s", cudaGetErrorString(cudaGetLastError()));
        cudaStreamSynchronize(vs); // leave nothing in flight on buffers the caller may now free
        cudaStreamSynchronize(co);
        return rc;
    }
    if (h_stats) *h_stats = st;
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_finish(hrm_mapper* m, int slot, int64_t* h_sq_written, int64_t* h_rec_written)
{
    HRM_REQUIRE(m != nullptr && slot >= 0 && slot < HRM_PIPE_SLOTS, "mapper / slot");
    HRM_TRY(bind_device(m));
    PipeSlot& S = m->slot[slot];
    if (S.busy) {
        HRM_CUDA(cudaEventSynchronize(S.computed));
        if (S.want_text) { // the text sizes are known now: copy exactly that much
            const int64_t* hp = m->pipe_host + 4 * slot;
            S.rec_written = hp[0];
            S.sq_written = S.h_sq ? hp[1] : 0;
            HRM_CUDA(cudaMemcpyAsync(S.h_rec, S.text.p, (size_t)S.rec_written, cudaMemcpyDeviceToHost, m->pipe_out[slot]));
            if (S.h_sq) HRM_CUDA(cudaMemcpyAsync(S.h_sq, S.sq.p, (size_t)S.sq_written, cudaMemcpyDeviceToHost, m->pipe_out[slot]));
            HRM_CUDA(cudaEventRecord(S.drained, m->pipe_out[slot]));
            S.want_text = false;
        }
        HRM_CUDA(cudaEventSynchronize(S.drained));
    }
    S.busy = false;
    if (h_sq_written) *h_sq_written = S.sq_written;
    if (h_rec_written) *h_rec_written = S.rec_written;
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_set_partition(hrm_mapper* m, hrm_comm* comm)
{
    HRM_REQUIRE(m != nullptr, "mapper");
    HRM_TRY(bind_device(m));
    if (m->d_win_prefix != nullptr) {
        set_error("hrm_mapper_set_partition must precede hrm_mapper_set_genome");
        return HRM_ERR_STATE;
    }
    m->comm = comm;
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_set_profiling(hrm_mapper* m, int enable)
{
    HRM_REQUIRE(m != nullptr, "mapper");
    m->timer.enabled = enable != 0;
    return HRM_OK;
}

extern "C" hrm_status hrm_mapper_stage_times(hrm_mapper* m, float* h_ms, int32_t* h_spans)
{
    HRM_REQUIRE(m != nullptr && h_ms != nullptr, "args");
    HRM_TRY(bind_device(m));
    for (int i = 0; i < HRM_NUM_STAGES; i++) {
        h_ms[i] = 0.f;
        if (h_spans) h_spans[i] = 0;
    }
    for (auto& sp : m->timer.spans) {
        HRM_CUDA(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        HRM_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        if (sp.stage >= 0 && sp.stage < HRM_NUM_STAGES) {
            h_ms[sp.stage] += ms;
            if (h_spans) h_spans[sp.stage]++;
        }
        m->timer.pool.push_back(sp.a);
        m->timer.pool.push_back(sp.b);
    }
    m->timer.spans.clear();
    return HRM_OK;
}
