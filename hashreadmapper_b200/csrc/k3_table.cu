// k3_table.cu -- K3: GPU-resident multi-value open-addressing hash tables (one per hash function).
// ref: GpuMinhasher include/gpu/gpuminhasher.cuh:20-110; the results reproduced are those of the
//      default CPU backend FakeGpuMinhasher include/gpu/fakegpuminhasher.cuh:199-442,568-728 with
//      CpuReadOnlyMultiValueHashTable include/cpuhashtable.hpp:465-679 finalised by GroupByKey
//      include/groupbykey.hpp:97-228 / 312-530.  Design precedent for a device table: warpcore
//      (include/gpu/gpuhashtable.cuh:303-1111) -- not used as an oracle.
//
// Layout in HBM (DESIGN.md "K3"): table j = array of 64-byte buckets, each bucket four 16-byte slots
// {u64 key, u32 value offset, u32 count}; values of all tables in one u32 array.  A bucket is exactly
// one HBM access (64 B), so a lookup that ends in its home bucket moves one DRAM atom and uses all of it.
// Build (one-off): stable radix sort of (key, id) pairs per table (CUB, build path only), run
// detection, truncation to the first min(maxResultsPerMap, 65535) ids, CAS insertion of the distinct
// keys.  Probe (hot path): one thread per (query, table) lookup reads the whole bucket with two 256-bit
// loads (LDG.E.256), linear probing over buckets; query keys staged into shared memory with TMA
// (cp.async.bulk + mbarrier, double buffered); ranges written coalesced.  Two tile orders: table-major
// (probe_tm_kernel, the mapper's default: signatures transposed to [H][n], all lookups in flight hit ONE table, so
// they share its pages and part of it stays in L2 -- 27 % -> 78-92 % of the HBM copy peak on a human-size index) and
// query-major (probe_count_kernel, the GpuMinhasher API path: per-query totals by warp shuffle).
#include "k3_table.cuh"
#include "core_minhash.cuh"
#include "k3_probe.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace hrm {

// ------------------------------------------------------------------------------------------
// build kernels
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t invalid_key(int k) { return k < 32 ? (1ULL << (2 * k)) : ~0ULL; }

// sigs[n][H] -> per-table staging (key, id); invalid signatures get a key that sorts last
__global__ void __launch_bounds__(256) stage_pairs_kernel(const uint64_t* __restrict__ sigs,
                                                          const uint8_t* __restrict__ valid, int64_t n, int H,
                                                          int Hsig, int first_func, int k, const uint32_t* __restrict__ ids,
                                                          uint32_t first_id, int64_t at,
                                                          uint64_t* const* __restrict__ stage_keys,
                                                          uint32_t* const* __restrict__ stage_vals, int part_rank,
                                                          int part_world)
{
    // thread per (table, sequence), sequence fastest => coalesced staging writes
    const int64_t total = n * H;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int j = (int)(t / n);
        const int64_t i = t - (int64_t)j * n;
        uint64_t key = sigs[i * Hsig + j];
        const bool ok = valid ? valid[i * Hsig + j] != 0 : key != ~0ULL;
        if (!ok || (part_world > 1 && (int)key_owner(key, (uint32_t)part_world) != part_rank)) key = invalid_key(k);
        stage_keys[first_func + j][at + i] = key;
        stage_vals[first_func + j][at + i] = ids ? ids[i] : first_id + (uint32_t)i;
    }
}

// flags[i] = 1 where a new valid key run starts; counts valid pairs
__global__ void __launch_bounds__(256) mark_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t inv,
                                                         int32_t* __restrict__ flags,
                                                         unsigned long long* __restrict__ valid_count)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        const bool v = key != inv && key != SLOT_EMPTY;
        flags[i] = (v && (i == 0 || keys[i - 1] != key)) ? 1 : 0;
        local += v ? 1 : 0;
    }
    // warp aggregate then one atomic per warp
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(valid_count, local);
}

// number of pairs with a valid key
__global__ void __launch_bounds__(256) count_valid_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t inv,
                                                          unsigned long long* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        local += (key != inv && key != SLOT_EMPTY) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// head_pos[u] = index of the first pair of distinct key u
__global__ void __launch_bounds__(256) head_positions_kernel(const int32_t* __restrict__ flags,
                                                             const int32_t* __restrict__ excl, int64_t n,
                                                             int32_t* __restrict__ head_pos)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (flags[i]) head_pos[excl[i]] = (int32_t)i;
}

// one thread per distinct key: claim a slot with CAS, then publish (offset, count)
__global__ void __launch_bounds__(256) insert_keys_kernel(const uint64_t* __restrict__ keys,
                                                          const int32_t* __restrict__ head_pos, int64_t nkeys,
                                                          int64_t nvalid, uint32_t value_base, uint32_t upper,
                                                          Slot* __restrict__ slots, uint32_t nbuckets,
                                                          unsigned long long* __restrict__ errors)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nkeys; u += stride) {
        const int64_t pos = head_pos[u];
        const int64_t end = u + 1 < nkeys ? (int64_t)head_pos[u + 1] : nvalid;
        const uint64_t key = keys[pos];
        int64_t cnt = end - pos;
        if (cnt > upper) cnt = upper; // ref: groupbykey.hpp:177-191 keeps the FIRST `upper` values
        uint32_t b = home_bucket(key, nbuckets);
        bool done = false;
        for (uint32_t probe = 0; probe < nbuckets && !done; probe++) {
            for (int sub = 0; sub < BUCKET_SLOTS && !done; sub++) {
                Slot* s = slots + ((size_t)b * BUCKET_SLOTS + sub);
                const unsigned long long prev =
                    atomicCAS(reinterpret_cast<unsigned long long*>(&s->key), (unsigned long long)SLOT_EMPTY,
                              (unsigned long long)key);
                if (prev == SLOT_EMPTY) {
                    s->off = value_base + (uint32_t)pos;
                    s->cnt = (uint32_t)cnt;
                    done = true;
                }
            }
            b = next_bucket(b, nbuckets);
        }
        if (!done) atomicAdd(errors, 1ULL);
    }
}

__global__ void fill_u64_kernel(uint64_t* p, int64_t n, uint64_t v)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// probe kernel (hot path)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

constexpr int PROBE_THREADS = 256;
constexpr int PROBE_LOOKUPS = 1024; // lookups (query x table) per tile

__global__ void __launch_bounds__(PROBE_THREADS) probe_count_kernel(const uint64_t* __restrict__ sigs, int n, int H,
                                                                    int TQ, const TablesParam* __restrict__ tabs_g,
                                                                    uint32_t max_results, uint2* __restrict__ ranges,
                                                                    int32_t* __restrict__ num_per_seq,
                                                                    unsigned long long* __restrict__ touches_out,
                                                                    int sigs_aligned)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(smem_raw);                       // [2][PROBE_LOOKUPS] keys
    uint32_t* cbuf = reinterpret_cast<uint32_t*>(sbuf + 2 * PROBE_LOOKUPS);       // [PROBE_LOOKUPS] counts (general H)
    uint64_t* bars = reinterpret_cast<uint64_t*>(cbuf + PROBE_LOOKUPS);           // [2]
    TableRef* tabs = reinterpret_cast<TableRef*>(bars + 2);                       // [H]

    const int tid = threadIdx.x;
    const int numTiles = (n + TQ - 1) / TQ;
    const int tileElems = TQ * H;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    for (int t = tid; t < H; t += PROBE_THREADS) tabs[t] = tabs_g->t[t];
    __syncthreads();

    auto tile_elems = [&](int tile) {
        const int q0 = tile * TQ;
        const int nq = (n - q0) < TQ ? (n - q0) : TQ;
        return nq * H;
    };
    auto tma_ok = [&](int tile) {
        return sigs_aligned && ((tile_elems(tile) * 8) % 16 == 0) && ((((size_t)tile * tileElems) * 8) % 16 == 0);
    };
    auto issue = [&](int tile, int st) {
        const uint32_t bytes = (uint32_t)tile_elems(tile) * 8u;
        mbar_expect_tx(&bars[st], bytes);
        tma_load_1d(sbuf + (size_t)st * PROBE_LOOKUPS, sigs + (size_t)tile * tileElems, bytes, &bars[st]);
    };

    uint32_t phases = 0u; // bit st = parity of the next completion of barrier st
    uint32_t visited = 0;
    // H a power of two <= 32: the H lookups of a query sit in H consecutive lanes -> totals by shuffle
    const bool seg = H <= 32 && (H & (H - 1)) == 0;
    const int lane = tid & 31;

    int tile = blockIdx.x;
    int it = 0;
    if (tid == 0 && tile < numTiles && tma_ok(tile)) issue(tile, 0);
    for (; tile < numTiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const int next = tile + gridDim.x;
        if (tid == 0 && next < numTiles && tma_ok(next)) issue(next, st ^ 1);
        const int elems = tile_elems(tile);
        uint64_t* keys = sbuf + (size_t)st * PROBE_LOOKUPS;
        if (tma_ok(tile)) {
            mbar_wait(&bars[st], (phases >> st) & 1u);
            phases ^= 1u << st;
        } else {
            for (int e = tid; e < elems; e += PROBE_THREADS) keys[e] = sigs[(size_t)tile * tileElems + e];
            __syncthreads();
        }
        // lookups: element e = local query * H + table; thread e, e + 256, ...
        uint2* gout = ranges + (size_t)tile * tileElems;
        const int q0 = tile * TQ;
        for (int e0 = 0; e0 < elems; e0 += PROBE_THREADS) {
            const int e = e0 + tid;
            uint2 r = make_uint2(0u, 0u);
            if (e < elems) {
                const int t = seg ? (e & (H - 1)) : (e % H);
                const TableRef T = tabs[t];
                r = probe_bucket_sequence(T.slots, T.nbuckets, keys[e], max_results, visited);
                gout[e] = r;
            }
            if (seg) {
                uint32_t sum = r.y;
                for (int d = 1; d < H; d <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
                if (e < elems && (lane & (H - 1)) == 0) num_per_seq[q0 + e / H] = (int)sum;
            } else {
                if (e < elems) cbuf[e] = r.y;
            }
        }
        if (!seg) {
            __syncthreads();
            for (int q = tid; q * H < elems; q += PROBE_THREADS) {
                int sum = 0;
                for (int t = 0; t < H; t++) sum += (int)cbuf[q * H + (t + q) % H]; // rotated: spreads the banks
                num_per_seq[q0 + q] = sum;
            }
        }
        __syncthreads(); // sbuf[st] (and cbuf) are free again
    }
    // slot touches: every bucket visit examines all of its slots
    visited *= BUCKET_SLOTS;
    for (int d = 16; d > 0; d >>= 1) visited += __shfl_xor_sync(0xffffffffu, visited, d);
    if (lane == 0 && visited && touches_out) atomicAdd(touches_out, (unsigned long long)visited);
}

// values of query q: buckets of tables 0..H-1 concatenated at d_values[offsets[q] ...]; one warp per query
__global__ void __launch_bounds__(256) retrieve_kernel(const uint2* __restrict__ ranges, int64_t rq, int64_t rt, int n,
                                                       int H, const uint32_t* __restrict__ table_values,
                                                       const int32_t* __restrict__ offsets,
                                                       uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp0; q < n; q += nwarps) {
        int64_t w = offsets[q];
        for (int t0 = 0; t0 < H; t0 += 32) {
            const int t = t0 + lane;
            const uint2 r = t < H ? ranges[q * rq + t * rt] : make_uint2(0u, 0u);
            // exclusive prefix of counts over the lanes
            int incl = (int)r.y;
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - (int)r.y;
            if (total <= 64) { // common case: each lane copies its own short bucket
                for (uint32_t v = 0; v < r.y; v++) out[w + excl + v] = table_values[r.x + v];
            } else { // long buckets: the whole warp copies bucket after bucket
                for (int tt = 0; tt < 32; tt++) {
                    const uint32_t off = __shfl_sync(0xffffffffu, r.x, tt);
                    const uint32_t cnt = __shfl_sync(0xffffffffu, r.y, tt);
                    const int ex = __shfl_sync(0xffffffffu, excl, tt);
                    for (uint32_t v = lane; v < cnt; v += 32) out[w + ex + v] = table_values[off + v];
                }
            }
            w += total;
        }
    }
}

static unsigned capped_grid(int64_t items, int block, int waves)
{
    int64_t g = HRM_SDIV(items, (int64_t)block);
    const int64_t cap = (int64_t)num_sms() * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

hrm_status minhasher_count_sigs(hrm_minhasher* mh, QueryHandle* qh, const uint64_t* d_sigs, int n,
                                int32_t* d_num_per_seq, cudaStream_t s)
{
    const int H = mh->H;
    HRM_TRY(qh->ranges.reserve(sizeof(uint2) * (size_t)n * H));
    const int TQ = PROBE_LOOKUPS / H > 0 ? PROBE_LOOKUPS / H : 1;
    const int numTiles = (n + TQ - 1) / TQ;
    const size_t smem = sizeof(uint64_t) * 2 * PROBE_LOOKUPS + sizeof(uint32_t) * PROBE_LOOKUPS + 2 * sizeof(uint64_t) +
                        sizeof(TableRef) * (size_t)H;
    int grid = numTiles;
    const int cap = num_sms() * 8; // 8 resident CTAs of 256 threads per SM
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const int aligned = (reinterpret_cast<uintptr_t>(d_sigs) & 15) == 0 ? 1 : 0;
    HRM_LAUNCH(probe_count_kernel, grid, PROBE_THREADS, smem, s, d_sigs, n, H, TQ, mh->d_param,
               (uint32_t)mh->max_results, qh->ranges.as<uint2>(), d_num_per_seq, mh->d_touches, aligned);
    qh->stage = 1;
    qh->n = n;
    qh->tm = false;
    qh->rq = H;
    qh->rt = 1;
    return HRM_OK;
}

// ---- table-major probe -----------------------------------------------------------------------------
// [n][H] -> [H][n]; block = 64 queries, staged through shared memory so that both sides are coalesced
__global__ void __launch_bounds__(256) transpose_sigs_kernel(const uint64_t* __restrict__ sigs, int n, int H,
                                                             uint64_t* __restrict__ sigs_tm)
{
    extern __shared__ uint64_t tsm[]; // [64][H + 1]
    const int nblk = (n + 63) / 64;
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int q0 = b * 64;
        const int nq = (n - q0) < 64 ? (n - q0) : 64;
        for (int e = threadIdx.x; e < nq * H; e += blockDim.x) tsm[(e / H) * (H + 1) + (e % H)] = sigs[(size_t)q0 * H + e];
        __syncthreads();
        for (int e = threadIdx.x; e < nq * H; e += blockDim.x) {
            const int t = e / nq, q = e - t * nq;
            sigs_tm[(size_t)t * n + q0 + q] = tsm[q * (H + 1) + t];
        }
        __syncthreads();
    }
}

// Same lookup as probe_count_kernel, tiles in table-major order: tile = (table t, 1024 consecutive queries).
// Keys of a tile are contiguous in sigs_tm -> one TMA bulk copy per tile, double buffered; ranges written
// coalesced at ranges_tm[t * n + q].
__global__ void __launch_bounds__(PROBE_THREADS) probe_tm_kernel(const uint64_t* __restrict__ sigs_tm, int n, int H,
                                                                 const TablesParam* __restrict__ tabs_g,
                                                                 uint32_t max_results, uint2* __restrict__ ranges_tm,
                                                                 unsigned long long* __restrict__ touches_out)
{
    __shared__ __align__(16) uint64_t sbuf[2][PROBE_LOOKUPS];
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x;
    const int tilesPerTable = (n + PROBE_LOOKUPS - 1) / PROBE_LOOKUPS;
    const int64_t numTiles = (int64_t)tilesPerTable * H;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto q0_of = [&](int64_t tile) { return (int)(tile % tilesPerTable) * PROBE_LOOKUPS; };
    auto t_of = [&](int64_t tile) { return (int)(tile / tilesPerTable); };
    auto elems_of = [&](int64_t tile) {
        const int q0 = q0_of(tile);
        return (n - q0) < PROBE_LOOKUPS ? (n - q0) : PROBE_LOOKUPS;
    };
    auto src_of = [&](int64_t tile) { return sigs_tm + (size_t)t_of(tile) * n + q0_of(tile); };
    auto tma_ok = [&](int64_t tile) {
        return ((reinterpret_cast<uintptr_t>(src_of(tile)) & 15) == 0) && ((elems_of(tile) * 8) % 16 == 0);
    };
    auto issue = [&](int64_t tile, int st) {
        const uint32_t bytes = (uint32_t)elems_of(tile) * 8u;
        mbar_expect_tx(&bars[st], bytes);
        tma_load_1d(sbuf[st], src_of(tile), bytes, &bars[st]);
    };
    uint32_t phases = 0u, visited = 0;
    int64_t tile = blockIdx.x;
    int it = 0;
    if (tid == 0 && tile < numTiles && tma_ok(tile)) issue(tile, 0);
    for (; tile < numTiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const int64_t next = tile + gridDim.x;
        if (tid == 0 && next < numTiles && tma_ok(next)) issue(next, st ^ 1);
        const int elems = elems_of(tile);
        uint64_t* keys = sbuf[st];
        if (tma_ok(tile)) {
            mbar_wait(&bars[st], (phases >> st) & 1u);
            phases ^= 1u << st;
        } else {
            const uint64_t* src = src_of(tile);
            for (int e = tid; e < elems; e += PROBE_THREADS) keys[e] = src[e];
            __syncthreads();
        }
        const TableRef T = tabs_g->t[t_of(tile)];
        uint2* gout = ranges_tm + (size_t)t_of(tile) * n + q0_of(tile);
        for (int e = tid; e < elems; e += PROBE_THREADS)
            gout[e] = probe_bucket_sequence(T.slots, T.nbuckets, keys[e], max_results, visited);
        __syncthreads(); // sbuf[st] is free again
    }
    visited *= BUCKET_SLOTS;
    for (int d = 16; d > 0; d >>= 1) visited += __shfl_xor_sync(0xffffffffu, visited, d);
    if ((tid & 31) == 0 && visited && touches_out) atomicAdd(touches_out, (unsigned long long)visited);
}

__global__ void __launch_bounds__(256) totals_tm_kernel(const uint2* __restrict__ ranges_tm, int n, int H,
                                                        int32_t* __restrict__ num_per_seq)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        int sum = 0;
        for (int t = 0; t < H; t++) sum += (int)ranges_tm[(size_t)t * n + q].y;
        num_per_seq[q] = sum;
    }
}

bool minhasher_wants_table_major(const hrm_minhasher* mh)
{
    // Measured on B200 (profiles/README.md): with all H tables in flight at once a human-size index (2.5 GB of
    // buckets per conversion) is probed at 27 % of the HBM copy peak, table by table at 78 %; a chr21-size
    // index (143 MB) goes from 65 % to 85 %.  HRM_PROBE_TM=0 selects the query-major kernel.
    static const int forced = getenv("HRM_PROBE_TM") ? atoi(getenv("HRM_PROBE_TM")) : -1;
    (void)mh;
    return forced != 0;
}

hrm_status minhasher_tm_prepare(hrm_minhasher* mh, QueryHandle* qh, const uint64_t* d_sigs, int n, cudaStream_t s)
{
    const int H = mh->H;
    HRM_TRY(qh->sigs_tm.reserve(sizeof(uint64_t) * (size_t)n * H));
    HRM_TRY(qh->ranges.reserve(sizeof(uint2) * (size_t)n * H));
    if (n == 0) return HRM_OK;
    const size_t smem = sizeof(uint64_t) * 64 * (size_t)(H + 1);
    HRM_LAUNCH(transpose_sigs_kernel, capped_grid((int64_t)((n + 63) / 64) * 256, 256, 8), 256, smem, s, d_sigs, n, H,
               qh->sigs_tm.as<uint64_t>());
    return HRM_OK;
}

hrm_status minhasher_tm_probe(hrm_minhasher* mh, QueryHandle* qh, int n, cudaStream_t s)
{
    const int H = mh->H;
    if (n > 0) {
        const int64_t numTiles = (int64_t)((n + PROBE_LOOKUPS - 1) / PROBE_LOOKUPS) * H;
        int64_t grid = numTiles;
        const int64_t cap = (int64_t)num_sms() * 8;
        if (grid > cap) grid = cap;
        HRM_LAUNCH(probe_tm_kernel, (unsigned)grid, PROBE_THREADS, 0, s, qh->sigs_tm.as<uint64_t>(), n, H, mh->d_param,
                   (uint32_t)mh->max_results, qh->ranges.as<uint2>(), mh->d_touches);
    }
    qh->stage = 1;
    qh->n = n;
    qh->tm = true;
    qh->rq = 1;
    qh->rt = n;
    return HRM_OK;
}

hrm_status minhasher_tm_totals(hrm_minhasher* mh, QueryHandle* qh, int n, int32_t* d_num_per_seq, cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(totals_tm_kernel, capped_grid(n, 256, 16), 256, 0, s, qh->ranges.as<uint2>(), n, mh->H, d_num_per_seq);
    return HRM_OK;
}

hrm_status minhasher_retrieve(hrm_minhasher* mh, QueryHandle* qh, int64_t first, int n, uint32_t* d_values,
                              const int32_t* d_offsets, cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(retrieve_kernel, capped_grid((int64_t)n * 32, 256, 16), 256, 0, s, qh->ranges.as<uint2>() + first * qh->rq,
               qh->rq, qh->rt, n, mh->H,
               mh->values, d_offsets, d_values);
    return HRM_OK;
}

QueryHandle* minhasher_handle(hrm_minhasher* mh, int id)
{
    std::lock_guard<std::mutex> lk(mh->mtx);
    if (id < 0 || id >= (int)mh->handles.size() || !mh->handles[id]) return nullptr;
    return mh->handles[id].get();
}

static void free_staging(hrm_minhasher* mh)
{
    for (auto p : mh->stage_keys)
        if (p) cudaFree(p);
    for (auto p : mh->stage_vals)
        if (p) cudaFree(p);
    mh->stage_keys.clear();
    mh->stage_vals.clear();
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_minhasher_create(hrm_minhasher** out, int64_t max_sequences, int max_results_per_map, int k,
                                           float load_factor)
{
    HRM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HRM_TRY(ensure_device());
    HRM_REQUIRE(k >= 1 && k <= 32, "1 <= k <= 32");
    HRM_REQUIRE(max_sequences >= 0 && max_sequences < (1LL << 32), "max_sequences must fit read_number (u32)");
    HRM_REQUIRE(max_results_per_map >= 0, "max_results_per_map");
    HRM_REQUIRE(load_factor > 0.f && load_factor <= 1.f, "0 < load_factor <= 1");
    auto* mh = new hrm_minhasher;
    mh->k = k;
    mh->max_results = max_results_per_map;
    mh->load = load_factor;
    mh->max_sequences = max_sequences;
    cudaGetDevice(&mh->device);
    memset(&mh->param, 0, sizeof mh->param);
    if (cudaMalloc(&mh->d_touches, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&mh->d_param, sizeof(TablesParam)) != cudaSuccess) {
        set_error("cudaMalloc failed");
        delete mh;
        return HRM_ERR_NOMEM;
    }
    cudaMemset(mh->d_touches, 0, sizeof(unsigned long long));
    *out = mh;
    return HRM_OK;
}

extern "C" void hrm_minhasher_destroy(hrm_minhasher* mh)
{
    if (!mh) return;
    free_staging(mh);
    if (mh->values) cudaFree(mh->values);
    for (auto p : mh->slots)
        if (p) cudaFree(p);
    if (mh->d_touches) cudaFree(mh->d_touches);
    if (mh->d_param) cudaFree(mh->d_param);
    delete mh;
}

extern "C" int hrm_minhasher_add_tables(hrm_minhasher* mh, int n, const int32_t* h_hash_function_ids, hrm_stream)
{
    if (!mh || n < 0 || mh->compacted) return 0;
    int added = 0;
    for (int t = 0; t < n; t++) {
        if (mh->H >= MAX_TABLES) break;
        if (h_hash_function_ids && h_hash_function_ids[t] != mh->H) break; // ids must be 0,1,2,...
        uint64_t* kp = nullptr;
        uint32_t* vp = nullptr;
        const size_t cap = (size_t)(mh->max_sequences > 0 ? mh->max_sequences : 1);
        if (cudaMalloc(&kp, sizeof(uint64_t) * cap) != cudaSuccess) {
            cudaGetLastError();
            break; // ref: fewer tables than requested signals memory shortage
        }
        if (cudaMalloc(&vp, sizeof(uint32_t) * cap) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(kp);
            break;
        }
        mh->stage_keys.push_back(kp);
        mh->stage_vals.push_back(vp);
        mh->table_count.push_back(0);
        mh->H++;
        added++;
    }
    return added;
}

static hrm_status stage_signatures(hrm_minhasher* mh, const uint64_t* d_sigs, const uint8_t* d_valid, int64_t n, int Hsig,
                                   int first_func, int num_funcs, const uint32_t* d_ids, uint32_t first_id,
                                   cudaStream_t s)
{
    // device copies of the staging pointer tables
    Scratch kp, vp;
    HRM_TRY(kp.alloc(sizeof(uint64_t*) * MAX_TABLES, s));
    HRM_TRY(vp.alloc(sizeof(uint32_t*) * MAX_TABLES, s));
    HRM_CUDA(cudaMemcpyAsync(kp.p, mh->stage_keys.data(), sizeof(uint64_t*) * mh->H, cudaMemcpyHostToDevice, s));
    HRM_CUDA(cudaMemcpyAsync(vp.p, mh->stage_vals.data(), sizeof(uint32_t*) * mh->H, cudaMemcpyHostToDevice, s));
    HRM_LAUNCH(stage_pairs_kernel, capped_grid(n * num_funcs, 256, 16), 256, 0, s, d_sigs, d_valid, n, num_funcs, Hsig,
               first_func, mh->k, d_ids, first_id, mh->table_count[first_func], kp.as<uint64_t*>(), vp.as<uint32_t*>(),
               mh->part_rank, mh->part_world);
    // the pointer tables are pageable host memory: make sure the copies are done before returning
    HRM_CUDA(cudaStreamSynchronize(s));
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_insert(hrm_minhasher* mh, const uint32_t* d_seq2bit, int64_t pitch_words,
                                           const int32_t* d_lengths, int64_t n, const uint32_t* d_ids, uint32_t first_id,
                                           int first_hash_func, int num_hash_funcs, hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    if (mh->compacted) {
        set_error("insert after compact");
        return HRM_ERR_STATE;
    }
    HRM_REQUIRE(first_hash_func >= 0 && num_hash_funcs >= 1 && first_hash_func + num_hash_funcs <= mh->H,
                "hash function range");
    for (int j = first_hash_func; j < first_hash_func + num_hash_funcs; j++)
        HRM_REQUIRE(mh->table_count[j] == mh->table_count[first_hash_func], "tables of one insert call must be equally filled");
    HRM_REQUIRE(n >= 0 && mh->table_count[first_hash_func] + n <= mh->max_sequences, "more sequences than max_sequences");
    if (n == 0) return HRM_OK;
    cudaStream_t s = as_stream(stream);
    // hash with functions 0..first+num-1 and keep the requested columns (ids are 0..H-1)
    const int Hsig = first_hash_func + num_hash_funcs;
    Scratch sigs, valid;
    HRM_TRY(sigs.alloc(sizeof(uint64_t) * (size_t)n * Hsig, s));
    HRM_TRY(valid.alloc((size_t)n * Hsig, s));
    HRM_TRY(minhash_rows(d_seq2bit, pitch_words, d_lengths, n, mh->k, Hsig, sigs.as<uint64_t>(), valid.as<uint8_t>(), s));
    // stage columns [first, first+num): pass a pointer offset by first_hash_func columns
    HRM_TRY(stage_signatures(mh, sigs.as<uint64_t>() + first_hash_func, valid.as<uint8_t>() + first_hash_func, n, Hsig,
                             first_hash_func, num_hash_funcs, d_ids, first_id, s));
    for (int j = first_hash_func; j < first_hash_func + num_hash_funcs; j++) mh->table_count[j] += n;
    mh->inserted = mh->table_count[0];
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_insert_signatures(hrm_minhasher* mh, const uint64_t* d_sigs, const uint8_t* d_valid,
                                                      int64_t n, const uint32_t* d_ids, uint32_t first_id,
                                                      hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    if (mh->compacted) {
        set_error("insert after compact");
        return HRM_ERR_STATE;
    }
    HRM_REQUIRE(mh->H > 0, "no tables");
    for (int j = 0; j < mh->H; j++)
        HRM_REQUIRE(mh->table_count[j] == mh->table_count[0], "tables must be equally filled");
    HRM_REQUIRE(n >= 0 && mh->table_count[0] + n <= mh->max_sequences, "more sequences than max_sequences");
    if (n == 0) return HRM_OK;
    HRM_TRY(stage_signatures(mh, d_sigs, d_valid, n, mh->H, 0, mh->H, d_ids, first_id, as_stream(stream)));
    for (int j = 0; j < mh->H; j++) mh->table_count[j] += n;
    mh->inserted = mh->table_count[0];
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_set_partition(hrm_minhasher* mh, int rank, int world)
{
    HRM_REQUIRE(mh != nullptr && world >= 1 && rank >= 0 && rank < world, "args");
    for (int j = 0; j < mh->H; j++)
        if (mh->table_count[j] != 0 || mh->compacted) {
            set_error("set_partition after insert");
            return HRM_ERR_STATE;
        }
    mh->part_rank = rank;
    mh->part_world = world;
    return HRM_OK;
}

extern "C" int hrm_minhasher_check_insertion_errors(hrm_minhasher*, int, int, hrm_stream) { return 0; }

extern "C" hrm_status hrm_minhasher_compact(hrm_minhasher* mh, hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    if (mh->compacted) return HRM_OK;
    cudaStream_t s = as_stream(stream);
    const int H = mh->H;
    for (int j = 0; j < H; j++)
        HRM_REQUIRE(mh->table_count[j] == mh->table_count[0], "compact: tables are not equally filled");
    const int64_t n = H > 0 ? mh->table_count[0] : 0;
    mh->inserted = n;
    HRM_REQUIRE((int64_t)H * n < (1LL << 32), "value offsets must fit 32 bits");
    mh->slots.assign(H, nullptr);
    mh->nbuckets.assign(H, 0);
    mh->nkeys.assign(H, 0);
    // values: only the pairs with a valid key are kept (a key-partitioned table stages the keys of other ranks
    // as invalid, so its value array shrinks with the number of ranks like its slots do)
    std::vector<int64_t> vbase(H + 1, 0);
    {
        Scratch vc;
        HRM_TRY(vc.alloc(sizeof(unsigned long long) * (size_t)(H > 0 ? H : 1), s));
        HRM_CUDA(cudaMemsetAsync(vc.p, 0, sizeof(unsigned long long) * (size_t)(H > 0 ? H : 1), s));
        const uint64_t inv0 = mh->k < 32 ? (1ULL << (2 * mh->k)) : ~0ULL;
        for (int j = 0; j < H && n > 0; j++)
            HRM_LAUNCH(count_valid_kernel, capped_grid(n, 256, 16), 256, 0, s, mh->stage_keys[j], n, inv0,
                       vc.as<unsigned long long>() + j);
        std::vector<unsigned long long> hv((size_t)(H > 0 ? H : 1), 0);
        HRM_CUDA(cudaMemcpyAsync(hv.data(), vc.p, sizeof(unsigned long long) * hv.size(), cudaMemcpyDeviceToHost, s));
        HRM_CUDA(cudaStreamSynchronize(s));
        for (int j = 0; j < H; j++) vbase[j + 1] = vbase[j] + (int64_t)hv[j];
    }
    mh->values_count = vbase[H];
    HRM_CUDA(cudaMalloc(&mh->values, sizeof(uint32_t) * (size_t)(mh->values_count > 0 ? mh->values_count : 1)));
    const uint16_t bs = (uint16_t)mh->max_results; // ref: BucketSize(maxValuesPerKey) groupbykey.hpp:178
    const uint32_t upper = bs < 65535 ? bs : 65535;
    const int end_bit = mh->k < 32 ? 2 * mh->k + 1 : 64;
    const uint64_t inv = mh->k < 32 ? (1ULL << (2 * mh->k)) : ~0ULL;

    Scratch keys_sorted, vals_sorted, flags, excl, head_pos, cub_tmp, counters;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    HRM_TRY(keys_sorted.alloc(sizeof(uint64_t) * nn, s));
    HRM_TRY(vals_sorted.alloc(sizeof(uint32_t) * nn, s));
    HRM_TRY(flags.alloc(sizeof(int32_t) * nn, s));
    HRM_TRY(excl.alloc(sizeof(int32_t) * (nn + 1), s));
    HRM_TRY(head_pos.alloc(sizeof(int32_t) * (nn + 1), s));
    HRM_TRY(counters.alloc(sizeof(unsigned long long) * 4, s));
    size_t tmp_bytes = 0;
    if (n > 0) {
        HRM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                 (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, end_bit, s));
        HRM_TRY(cub_tmp.alloc(tmp_bytes, s));
    }
    for (int j = 0; j < H; j++) {
        int64_t nkeys = 0, nvalid = 0;
        uint32_t* vals_out = vals_sorted.as<uint32_t>();
        if (n > 0) {
            HRM_REQUIRE(n < (1LL << 31), "at most 2^31-1 sequences per table");
            HRM_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, mh->stage_keys[j], keys_sorted.as<uint64_t>(),
                                                     mh->stage_vals[j], vals_out, (int)n, 0, end_bit, s));
            g_launches.fetch_add(1);
            HRM_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * 4, s));
            HRM_LAUNCH(mark_heads_kernel, capped_grid(n, 256, 16), 256, 0, s, keys_sorted.as<uint64_t>(), n, inv,
                       flags.as<int32_t>(), counters.as<unsigned long long>());
            HRM_TRY(exclusive_scan_i32(flags.as<int32_t>(), excl.as<int32_t>(), n,
                                       reinterpret_cast<int64_t*>(counters.as<unsigned long long>() + 1), s));
            HRM_LAUNCH(head_positions_kernel, capped_grid(n, 256, 16), 256, 0, s, flags.as<int32_t>(), excl.as<int32_t>(),
                       n, head_pos.as<int32_t>());
            unsigned long long h_cnt[2];
            HRM_CUDA(cudaMemcpyAsync(h_cnt, counters.p, sizeof h_cnt, cudaMemcpyDeviceToHost, s));
            HRM_CUDA(cudaStreamSynchronize(s));
            nvalid = (int64_t)h_cnt[0];
            nkeys = (int64_t)h_cnt[1];
            HRM_REQUIRE(nvalid == vbase[j + 1] - vbase[j], "compact: valid pair count changed");
            if (nvalid > 0) // valid keys sort first: their values are the prefix
                HRM_CUDA(cudaMemcpyAsync(mh->values + vbase[j], vals_out, sizeof(uint32_t) * (size_t)nvalid,
                                         cudaMemcpyDeviceToDevice, s));
        }
        // smallest bucket count with nkeys / (BUCKET_SLOTS * nbuckets) <= load factor
        const int64_t nb = (int64_t)((double)nkeys / (double)mh->load / (double)BUCKET_SLOTS) + 1;
        HRM_REQUIRE(nb < (1LL << 32), "table too large");
        Slot* sl = nullptr;
        HRM_CUDA(cudaMalloc(&sl, (size_t)BUCKET_BYTES * (size_t)nb));
        mh->slots[j] = sl;
        mh->nbuckets[j] = nb;
        mh->nkeys[j] = nkeys;
        // empty pattern: every 64-bit word = ~0 (key == SLOT_EMPTY; payload irrelevant)
        HRM_CUDA(cudaMemsetAsync(sl, 0xFF, (size_t)BUCKET_BYTES * (size_t)nb, s));
        if (nkeys > 0) {
            HRM_LAUNCH(insert_keys_kernel, capped_grid(nkeys, 256, 16), 256, 0, s, keys_sorted.as<uint64_t>(),
                       head_pos.as<int32_t>(), nkeys, nvalid, (uint32_t)vbase[j], upper, sl, (uint32_t)nb,
                       counters.as<unsigned long long>() + 2);
        }
        mh->param.t[j].slots = sl;
        mh->param.t[j].nbuckets = (uint32_t)nb;
        // staging of this table is no longer needed
        cudaFree(mh->stage_keys[j]);
        cudaFree(mh->stage_vals[j]);
        mh->stage_keys[j] = nullptr;
        mh->stage_vals[j] = nullptr;
    }
    unsigned long long h_err = 0;
    HRM_CUDA(cudaMemcpyAsync(&h_err, counters.as<unsigned long long>() + 2, sizeof h_err, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    if (n > 0 && h_err != 0) {
        set_error("hash table insertion failed for %llu keys", h_err);
        return HRM_ERR_CUDA;
    }
    HRM_CUDA(cudaMemcpy(mh->d_param, &mh->param, sizeof mh->param, cudaMemcpyHostToDevice));
    mh->compacted = true;
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_finish(hrm_minhasher* mh, hrm_stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    free_staging(mh);
    mh->finished = true;
    return HRM_OK;
}

extern "C" int hrm_minhasher_handle_create(hrm_minhasher* mh)
{
    if (!mh) return HRM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(mh->mtx);
    mh->handles.emplace_back(new QueryHandle);
    mh->handles.back()->in_use = true;
    return (int)mh->handles.size() - 1;
}

extern "C" hrm_status hrm_minhasher_handle_destroy(hrm_minhasher* mh, int handle)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    std::lock_guard<std::mutex> lk(mh->mtx);
    HRM_REQUIRE(handle >= 0 && handle < (int)mh->handles.size() && mh->handles[handle], "handle");
    mh->handles[handle].reset(); // ref: destroyHandle nulls the slot (fakegpuminhasher.cuh:181-190)
    return HRM_OK;
}

static hrm_status finish_count(hrm_minhasher* mh, QueryHandle* qh, int n, const int32_t* d_num_per_seq, int64_t* h_total,
                               cudaStream_t s)
{
    // total = sum of counts (one scan over a scratch copy; the reference reduces on the host)
    Scratch tmp;
    HRM_TRY(tmp.alloc(sizeof(int32_t) * ((size_t)n + 1) + sizeof(int64_t), s));
    HRM_TRY(qh->misc.reserve(sizeof(int64_t)));
    HRM_TRY(exclusive_scan_i32(d_num_per_seq, tmp.as<int32_t>(), n, qh->misc.as<int64_t>(), s));
    int64_t total = 0;
    HRM_CUDA(cudaMemcpyAsync(&total, qh->misc.p, sizeof total, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s)); // ref: fakegpuminhasher.cuh:260 synchronises as well
    if (total > 0x7fffffffLL) {
        set_error("total number of values %lld exceeds int (ref: main_gpu.cu:87)", (long long)total);
        return HRM_ERR_OVERFLOW;
    }
    if (h_total) *h_total = total;
    (void)mh;
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_count_signatures(hrm_minhasher* mh, int handle, const uint64_t* d_sigs,
                                                     const uint8_t* /*d_valid: invalid <=> sig == ~0*/, int n,
                                                     int32_t* d_num_per_seq, int64_t* h_total, hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    if (!mh->compacted) {
        set_error("query before compact");
        return HRM_ERR_STATE;
    }
    QueryHandle* qh = minhasher_handle(mh, handle);
    HRM_REQUIRE(qh != nullptr, "handle");
    HRM_REQUIRE(n >= 0, "n");
    if (n == 0) { // ref: returns immediately (fakegpuminhasher.cuh:216)
        if (h_total) *h_total = 0;
        qh->stage = 1;
        qh->n = 0;
        return HRM_OK;
    }
    cudaStream_t s = as_stream(stream);
    HRM_TRY(minhasher_count_sigs(mh, qh, d_sigs, n, d_num_per_seq, s));
    return finish_count(mh, qh, n, d_num_per_seq, h_total, s);
}

extern "C" hrm_status hrm_minhasher_count(hrm_minhasher* mh, int handle, const uint32_t* d_seq2bit, int64_t pitch_words,
                                          const int32_t* d_lengths, int n, int32_t* d_num_per_seq, int64_t* h_total,
                                          hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    if (!mh->compacted) {
        set_error("query before compact");
        return HRM_ERR_STATE;
    }
    QueryHandle* qh = minhasher_handle(mh, handle);
    HRM_REQUIRE(qh != nullptr, "handle");
    HRM_REQUIRE(n >= 0, "n");
    if (n == 0) {
        if (h_total) *h_total = 0;
        qh->stage = 1;
        qh->n = 0;
        return HRM_OK;
    }
    cudaStream_t s = as_stream(stream);
    HRM_TRY(qh->sigs.reserve(sizeof(uint64_t) * (size_t)n * mh->H));
    HRM_TRY(minhash_rows(d_seq2bit, pitch_words, d_lengths, n, mh->k, mh->H, qh->sigs.as<uint64_t>(), nullptr, s));
    HRM_TRY(minhasher_count_sigs(mh, qh, qh->sigs.as<uint64_t>(), n, d_num_per_seq, s));
    return finish_count(mh, qh, n, d_num_per_seq, h_total, s);
}

extern "C" hrm_status hrm_minhasher_retrieve(hrm_minhasher* mh, int handle, int n, int64_t total, uint32_t* d_values,
                                             const int32_t* d_num_per_seq, int32_t* d_offsets, hrm_stream stream)
{
    HRM_REQUIRE(mh != nullptr, "minhasher");
    QueryHandle* qh = minhasher_handle(mh, handle);
    HRM_REQUIRE(qh != nullptr, "handle");
    if (qh->stage != 1 || qh->n != n) { // ref: assert(previousStage == NumValues) fakegpuminhasher.cuh:328
        set_error("retrieve must follow count on the same handle with the same n");
        return HRM_ERR_STATE;
    }
    qh->stage = 0;
    if (n == 0) return HRM_OK;
    cudaStream_t s = as_stream(stream);
    if (total == 0) { // ref: fakegpuminhasher.cuh:332-335
        HRM_CUDA(cudaMemsetAsync(d_offsets, 0, sizeof(int32_t) * ((size_t)n + 1), s));
        return HRM_OK;
    }
    HRM_TRY(exclusive_scan_i32(d_num_per_seq, d_offsets, n, nullptr, s));
    return minhasher_retrieve(mh, qh, 0, n, d_values, d_offsets, s);
}

extern "C" hrm_status hrm_minhasher_info(const hrm_minhasher* mh, hrm_minhasher_info_t* out)
{
    HRM_REQUIRE(mh != nullptr && out != nullptr, "args");
    memset(out, 0, sizeof *out);
    out->k = mh->k;
    out->num_tables = mh->H;
    out->max_results_per_map = mh->max_results;
    out->load_factor = mh->load;
    out->num_inserted = mh->inserted;
    out->is_compacted = mh->compacted ? 1 : 0;
    out->has_gpu_tables = 1;
    int64_t bytes = 0;
    if (mh->compacted) {
        for (int j = 0; j < mh->H; j++) {
            out->num_keys_total += mh->nkeys[j];
            bytes += mh->nbuckets[j] * BUCKET_BYTES;
        }
        out->num_values_total = mh->values_count;
        bytes += mh->values_count * 4;
    }
    for (auto p : mh->stage_keys)
        if (p) bytes += mh->max_sequences * 12;
    out->device_bytes = bytes;
    return HRM_OK;
}

// ---- serialisation (own format; ref: writeToStream/loadFromStream fakegpuminhasher.cuh:498-532) ----
namespace {
struct SerHeader {
    char magic[8]; // "HRMB200\0"
    int32_t version, k, max_results, H;
    float load;
    int32_t pad;
    int64_t inserted, values_count;
};
} // namespace

extern "C" hrm_status hrm_minhasher_serialize(const hrm_minhasher* mh, void* h_buf, int64_t* h_size)
{
    HRM_REQUIRE(mh != nullptr && h_size != nullptr, "args");
    if (!mh->compacted) {
        set_error("serialize before compact");
        return HRM_ERR_STATE;
    }
    int64_t need = sizeof(SerHeader) + sizeof(int64_t) * 2 * mh->H + mh->values_count * 4;
    for (int j = 0; j < mh->H; j++) need += mh->nbuckets[j] * BUCKET_BYTES;
    if (!h_buf) {
        *h_size = need;
        return HRM_OK;
    }
    HRM_REQUIRE(*h_size >= need, "buffer too small");
    char* p = (char*)h_buf;
    SerHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "HRMB200", 8);
    hd.version = 2; // 2: 64-byte buckets, multiply-shift home bucket
    hd.k = mh->k;
    hd.max_results = mh->max_results;
    hd.H = mh->H;
    hd.load = mh->load;
    hd.inserted = mh->inserted;
    hd.values_count = mh->values_count;
    memcpy(p, &hd, sizeof hd);
    p += sizeof hd;
    for (int j = 0; j < mh->H; j++) {
        memcpy(p, &mh->nbuckets[j], 8);
        p += 8;
        memcpy(p, &mh->nkeys[j], 8);
        p += 8;
    }
    HRM_CUDA(cudaMemcpy(p, mh->values, (size_t)mh->values_count * 4, cudaMemcpyDeviceToHost));
    p += mh->values_count * 4;
    for (int j = 0; j < mh->H; j++) {
        HRM_CUDA(cudaMemcpy(p, mh->slots[j], (size_t)mh->nbuckets[j] * BUCKET_BYTES, cudaMemcpyDeviceToHost));
        p += mh->nbuckets[j] * BUCKET_BYTES;
    }
    *h_size = need;
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_deserialize(hrm_minhasher** out, const void* h_buf, int64_t size)
{
    HRM_REQUIRE(out != nullptr && h_buf != nullptr, "args");
    *out = nullptr;
    HRM_REQUIRE(size >= (int64_t)sizeof(SerHeader), "truncated");
    const char* p = (const char*)h_buf;
    SerHeader hd;
    memcpy(&hd, p, sizeof hd);
    p += sizeof hd;
    HRM_REQUIRE(memcmp(hd.magic, "HRMB200", 8) == 0 && hd.version == 2, "bad magic/version");
    HRM_REQUIRE(hd.H >= 0 && hd.H <= MAX_TABLES, "bad table count");
    // counts from the image are bounded by the image before anything is sized from them
    HRM_REQUIRE(hd.values_count >= 0 && hd.values_count <= (size - (int64_t)sizeof(SerHeader)) / 4 && hd.inserted >= 0 &&
                    size >= (int64_t)sizeof(SerHeader) + 16LL * hd.H,
                "bad counts");
    hrm_minhasher* mh = nullptr;
    HRM_TRY(hrm_minhasher_create(&mh, hd.inserted, hd.max_results, hd.k, hd.load));
    mh->H = hd.H;
    mh->inserted = hd.inserted;
    mh->values_count = hd.values_count;
    mh->slots.assign(hd.H, nullptr);
    mh->nbuckets.assign(hd.H, 0);
    mh->nkeys.assign(hd.H, 0);
    int64_t need = sizeof(SerHeader) + 16LL * hd.H + hd.values_count * 4;
    for (int j = 0; j < hd.H; j++) {
        memcpy(&mh->nbuckets[j], p, 8);
        p += 8;
        memcpy(&mh->nkeys[j], p, 8);
        p += 8;
        if (mh->nbuckets[j] < 0 || mh->nbuckets[j] > (size - need) / BUCKET_BYTES || mh->nkeys[j] < 0) {
            need = INT64_MAX;
            break;
        }
        need += mh->nbuckets[j] * BUCKET_BYTES;
    }
    // every slot's value range must lie inside the value array: a corrupt image must not turn into out-of-bounds
    // device reads at probe time
    if (need != INT64_MAX && size >= need) {
        const char* sp = p + hd.values_count * 4;
        for (int j = 0; j < hd.H && need != INT64_MAX; j++) {
            const int64_t nslots = mh->nbuckets[j] * BUCKET_SLOTS;
            for (int64_t i = 0; i < nslots; i++) {
                Slot sl;
                memcpy(&sl, sp + i * (int64_t)sizeof(Slot), sizeof sl);
                if (sl.key != SLOT_EMPTY && (uint64_t)sl.off + sl.cnt > (uint64_t)hd.values_count) {
                    need = INT64_MAX;
                    break;
                }
            }
            sp += mh->nbuckets[j] * BUCKET_BYTES;
        }
    }
    if (need == INT64_MAX || size < need) {
        hrm_minhasher_destroy(mh);
        set_error("truncated or corrupt minhasher image");
        return HRM_ERR_INVALID;
    }
    cudaError_t e = cudaMalloc(&mh->values, (size_t)(hd.values_count > 0 ? hd.values_count : 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(mh->values, p, (size_t)hd.values_count * 4, cudaMemcpyHostToDevice);
    p += hd.values_count * 4;
    for (int j = 0; j < hd.H && e == cudaSuccess; j++) {
        Slot* sl = nullptr;
        e = cudaMalloc(&sl, (size_t)mh->nbuckets[j] * BUCKET_BYTES);
        if (e != cudaSuccess) break;
        mh->slots[j] = sl;
        e = cudaMemcpy(sl, p, (size_t)mh->nbuckets[j] * BUCKET_BYTES, cudaMemcpyHostToDevice);
        p += mh->nbuckets[j] * BUCKET_BYTES;
        mh->param.t[j].slots = sl;
        mh->param.t[j].nbuckets = (uint32_t)mh->nbuckets[j];
    }
    if (e == cudaSuccess) e = cudaMemcpy(mh->d_param, &mh->param, sizeof mh->param, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("deserialize: %s", cudaGetErrorString(e));
        hrm_minhasher_destroy(mh);
        return HRM_ERR_CUDA;
    }
    mh->compacted = true;
    mh->finished = true;
    *out = mh;
    return HRM_OK;
}

// ---- the reference's own file format (--save-hashtables-to / --load-hashtables-from) ----------------------
// ref: FakeGpuMinhasher::writeToStream / loadFromStream include/gpu/fakegpuminhasher.cuh:498-532:
//        int kmerSize, int resultsPerMapThreshold, float loadfactor, int numTables, then per table
//      CpuReadOnlyMultiValueHashTable::writeToStream include/cpuhashtable.hpp:624-646:
//        size_t nValues, u32 values[nValues] (grouped by ascending key, truncated buckets already cut), then
//      AoSCpuSingleValueHashTable::writeToStream :217-229:
//        float load, size_t numKeys, size_t maxProbes, size_t size, size_t capacity, size_t elements,
//        Data storage[capacity], Data = pair<u64 key, pair<u32 offset, u16 count>> (16 bytes, 2 padding bytes),
//        empty slot = {~0, {0, 0}} (:277-278), capacity = size_t(float(size) / load) (:60-64), keys inserted in
//        ascending order at murmur64(key) % capacity with linear probing (:96-124, built at :524-544).
// The tables of this library are bucketised differently, so writing re-inserts every key on the host exactly
// as the reference does and reading re-inserts every stored key into 64-byte buckets on the device.
namespace {
struct RefSlot {
    uint64_t key;
    uint32_t off;
    uint16_t cnt;
    uint16_t pad;
};
static_assert(sizeof(RefSlot) == 16, "reference Data is 16 bytes");

struct KeyEntry {
    uint64_t key;
    uint32_t off, cnt;
};

template <class T>
void put(std::vector<char>& out, const T& v)
{
    const char* p = reinterpret_cast<const char*>(&v);
    out.insert(out.end(), p, p + sizeof(T));
}
template <class T>
bool get(const char*& p, const char* end, T& v)
{
    if (end - p < (ptrdiff_t)sizeof(T)) return false;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return true;
}

// explicit (key, offset, count) entries into the bucketised device table
__global__ void __launch_bounds__(256) insert_entries_kernel(const KeyEntry* __restrict__ entries, int64_t n,
                                                             uint32_t value_base, Slot* __restrict__ slots,
                                                             uint32_t nbuckets, unsigned long long* __restrict__ errors)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride) {
        const KeyEntry en = entries[u];
        uint32_t b = home_bucket(en.key, nbuckets);
        bool done = false;
        for (uint32_t probe = 0; probe < nbuckets && !done; probe++) {
            for (int sub = 0; sub < BUCKET_SLOTS && !done; sub++) {
                Slot* s = slots + ((size_t)b * BUCKET_SLOTS + sub);
                const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&s->key),
                                                          (unsigned long long)SLOT_EMPTY, (unsigned long long)en.key);
                if (prev == SLOT_EMPTY) {
                    s->off = value_base + en.off;
                    s->cnt = en.cnt;
                    done = true;
                }
            }
            b = next_bucket(b, nbuckets);
        }
        if (!done) atomicAdd(errors, 1ULL);
    }
}
} // namespace

extern "C" hrm_status hrm_minhasher_write_reference_format(const hrm_minhasher* mh, void* h_buf, int64_t* h_size)
{
    HRM_REQUIRE(mh != nullptr && h_size != nullptr, "args");
    if (!mh->compacted) {
        set_error("write_reference_format before compact");
        return HRM_ERR_STATE;
    }
    std::vector<char> out;
    put(out, (int32_t)mh->k);
    put(out, (int32_t)mh->max_results);
    put(out, (float)mh->load);
    put(out, (int32_t)mh->H);
    std::vector<uint32_t> hvalues((size_t)(mh->values_count > 0 ? mh->values_count : 1));
    if (mh->values_count > 0)
        HRM_CUDA(cudaMemcpy(hvalues.data(), mh->values, (size_t)mh->values_count * 4, cudaMemcpyDeviceToHost));
    for (int j = 0; j < mh->H; j++) {
        std::vector<Slot> hs((size_t)mh->nbuckets[j] * BUCKET_SLOTS);
        HRM_CUDA(cudaMemcpy(hs.data(), mh->slots[j], hs.size() * sizeof(Slot), cudaMemcpyDeviceToHost));
        std::vector<KeyEntry> keys;
        keys.reserve((size_t)mh->nkeys[j]);
        for (const Slot& s : hs)
            if (s.key != SLOT_EMPTY) keys.push_back({s.key, s.off, s.cnt});
        std::sort(keys.begin(), keys.end(), [](const KeyEntry& a, const KeyEntry& b) { return a.key < b.key; });
        // values grouped by ascending key, each bucket cut to its stored count (ref: groupbykey.hpp:177-191)
        std::vector<uint32_t> vals;
        std::vector<uint32_t> newoff(keys.size());
        for (size_t u = 0; u < keys.size(); u++) {
            newoff[u] = (uint32_t)vals.size();
            vals.insert(vals.end(), hvalues.begin() + keys[u].off, hvalues.begin() + keys[u].off + keys[u].cnt);
        }
        put(out, (uint64_t)vals.size());
        if (!vals.empty()) out.insert(out.end(), (const char*)vals.data(), (const char*)(vals.data() + vals.size()));
        // the key -> (offset, count) table exactly as AoSCpuSingleValueHashTable builds it
        const uint64_t size = keys.size();
        const uint64_t capacity = (uint64_t)((float)size / mh->load); // float arithmetic, as `capacity(size/load)`
        std::vector<RefSlot> storage((size_t)capacity, RefSlot{~0ULL, 0u, 0, 0});
        uint64_t maxProbes = 0;
        if (!keys.empty()) HRM_REQUIRE(capacity >= size, "load factor above 1");
        for (size_t u = 0; u < keys.size(); u++) {
            HRM_REQUIRE(keys[u].cnt <= 65535u, "bucket count exceeds the reference's 16-bit BucketSize");
            uint64_t pos = murmur64(keys[u].key) % capacity, probes = 0;
            while (storage[pos].key != ~0ULL) {
                pos = pos + 1 == capacity ? 0 : pos + 1;
                probes++;
            }
            storage[pos] = RefSlot{keys[u].key, newoff[u], (uint16_t)keys[u].cnt, 0};
            maxProbes = probes > maxProbes ? probes : maxProbes;
        }
        put(out, (float)mh->load);
        put(out, (uint64_t)size);      // numKeys
        put(out, (uint64_t)maxProbes);
        put(out, (uint64_t)size);      // size
        put(out, (uint64_t)capacity);
        put(out, (uint64_t)storage.size());
        if (!storage.empty())
            out.insert(out.end(), (const char*)storage.data(), (const char*)(storage.data() + storage.size()));
    }
    if (!h_buf) {
        *h_size = (int64_t)out.size();
        return HRM_OK;
    }
    HRM_REQUIRE(*h_size >= (int64_t)out.size(), "buffer too small");
    memcpy(h_buf, out.data(), out.size());
    *h_size = (int64_t)out.size();
    return HRM_OK;
}

extern "C" hrm_status hrm_minhasher_read_reference_format(hrm_minhasher** out, const void* h_buf, int64_t size,
                                                          int max_tables)
{
    HRM_REQUIRE(out != nullptr && h_buf != nullptr && size >= 16, "args");
    *out = nullptr;
    HRM_TRY(ensure_device());
    const char* p = (const char*)h_buf;
    const char* end = p + size;
    int32_t k = 0, thr = 0, H = 0;
    float load = 0.f;
    HRM_REQUIRE(get(p, end, k) && get(p, end, thr) && get(p, end, load) && get(p, end, H), "truncated header");
    HRM_REQUIRE(k >= 1 && k <= 32 && H >= 0 && H <= MAX_TABLES && load > 0.f && load <= 1.f, "bad header");
    if (max_tables >= 0 && H > max_tables) H = max_tables; // ref: loadFromStream(is, numMapsUpperLimit)
    // first pass over the image: table extents
    struct Ext {
        const uint32_t* vals;
        uint64_t nvals;
        const RefSlot* storage;
        uint64_t elements;
    };
    std::vector<Ext> ext((size_t)H);
    int64_t total_vals = 0, max_inserted = 0;
    for (int j = 0; j < H; j++) {
        uint64_t nvals = 0, numKeys = 0, maxProbes = 0, sz = 0, cap = 0, elements = 0;
        float tl = 0.f;
        // sizes from the file are bounded by what is left of it BEFORE they are multiplied (no wrap-around)
        HRM_REQUIRE(get(p, end, nvals) && nvals <= (uint64_t)(end - p) / 4, "truncated values");
        ext[j].vals = (const uint32_t*)p;
        ext[j].nvals = nvals;
        p += nvals * 4;
        HRM_REQUIRE(get(p, end, tl) && get(p, end, numKeys) && get(p, end, maxProbes) && get(p, end, sz) &&
                        get(p, end, cap) && get(p, end, elements),
                    "truncated table header");
        HRM_REQUIRE(elements <= (uint64_t)(end - p) / sizeof(RefSlot), "truncated table storage");
        ext[j].storage = (const RefSlot*)p;
        ext[j].elements = elements;
        p += elements * sizeof(RefSlot);
        total_vals += (int64_t)nvals;
        max_inserted = (int64_t)nvals > max_inserted ? (int64_t)nvals : max_inserted;
    }
    HRM_REQUIRE(total_vals < (1LL << 32), "value offsets must fit 32 bits");
    hrm_minhasher* mh = nullptr;
    HRM_TRY(hrm_minhasher_create(&mh, max_inserted, thr, k, load));
    mh->H = H;
    mh->inserted = max_inserted;
    mh->values_count = total_vals;
    mh->slots.assign(H, nullptr);
    mh->nbuckets.assign(H, 0);
    mh->nkeys.assign(H, 0);
    mh->table_count.assign(H, max_inserted);
    auto fail = [&](const char* what, cudaError_t e) {
        set_error("read_reference_format: %s: %s", what, cudaGetErrorString(e));
        hrm_minhasher_destroy(mh);
        return e == cudaErrorMemoryAllocation ? HRM_ERR_NOMEM : HRM_ERR_CUDA;
    };
    cudaError_t e = cudaMalloc(&mh->values, (size_t)(total_vals > 0 ? total_vals : 1) * 4);
    if (e != cudaSuccess) return fail("values", e);
    unsigned long long* d_err = nullptr;
    e = cudaMalloc(&d_err, sizeof(unsigned long long));
    if (e != cudaSuccess) return fail("counter", e);
    cudaMemset(d_err, 0, sizeof(unsigned long long));
    int64_t vbase = 0;
    for (int j = 0; j < H; j++) {
        std::vector<KeyEntry> keys;
        for (uint64_t i = 0; i < ext[j].elements; i++) {
            RefSlot s;
            memcpy(&s, ext[j].storage + i, sizeof s);
            if (s.key == ~0ULL && s.off == 0u && s.cnt == 0) continue; // ref: emptySlot cpuhashtable.hpp:277-278
            if (s.key == SLOT_EMPTY || (uint64_t)s.off + s.cnt > ext[j].nvals) { // ~0 is the device table's empty marker
                cudaFree(d_err);
                hrm_minhasher_destroy(mh);
                set_error("read_reference_format: value range of a key exceeds the value array");
                return HRM_ERR_INVALID;
            }
            keys.push_back({s.key, s.off, s.cnt});
        }
        if (ext[j].nvals > 0) {
            std::vector<uint32_t> tmp(ext[j].nvals); // the image need not be 4-byte aligned
            memcpy(tmp.data(), ext[j].vals, ext[j].nvals * 4);
            e = cudaMemcpy(mh->values + vbase, tmp.data(), ext[j].nvals * 4, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) break;
        }
        const int64_t nb = (int64_t)((double)keys.size() / (double)mh->load / (double)BUCKET_SLOTS) + 1;
        Slot* sl = nullptr;
        e = cudaMalloc(&sl, (size_t)BUCKET_BYTES * (size_t)nb);
        if (e != cudaSuccess) break;
        mh->slots[j] = sl;
        mh->nbuckets[j] = nb;
        mh->nkeys[j] = (int64_t)keys.size();
        mh->param.t[j].slots = sl;
        mh->param.t[j].nbuckets = (uint32_t)nb;
        e = cudaMemset(sl, 0xFF, (size_t)BUCKET_BYTES * (size_t)nb);
        if (e != cudaSuccess) break;
        if (!keys.empty()) {
            KeyEntry* d_keys = nullptr;
            e = cudaMalloc(&d_keys, keys.size() * sizeof(KeyEntry));
            if (e != cudaSuccess) break;
            e = cudaMemcpy(d_keys, keys.data(), keys.size() * sizeof(KeyEntry), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) {
                insert_entries_kernel<<<capped_grid((int64_t)keys.size(), 256, 16), 256>>>(d_keys, (int64_t)keys.size(),
                                                                                          (uint32_t)vbase, sl, (uint32_t)nb,
                                                                                          d_err);
                g_launches.fetch_add(1);
                e = cudaDeviceSynchronize();
            }
            cudaFree(d_keys);
            if (e != cudaSuccess) break;
        }
        vbase += (int64_t)ext[j].nvals;
    }
    unsigned long long h_err = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    if (e == cudaSuccess) e = cudaMemcpy(mh->d_param, &mh->param, sizeof mh->param, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail("tables", e);
    if (h_err != 0) {
        hrm_minhasher_destroy(mh);
        set_error("read_reference_format: %llu keys could not be inserted", h_err);
        return HRM_ERR_CUDA;
    }
    mh->compacted = true;
    mh->finished = true;
    *out = mh;
    return HRM_OK;
}
