// partition.cu -- key-partitioned 3N index over the GPUs of one box, queries routed with NCCL all-to-all
// over NVLink (BASELINE config 5).
// ref: what the reference does for several GPUs is different -- MultiGpuMinhasher puts whole tables
//      (hash function j on GPU j mod G) on the devices, BROADCASTS every query batch to all of them with
//      cudaMemcpyPeerAsync and gathers counts/values back by peer copies
//      (include/gpu/multigpuminhasher.cuh:257-333 layout, :659-675 broadcast, :724,:835-870 gather, merge
//      kernels :48-200).  Here every table is split by key: owner(key) = mulhi(high word of
//      murmur64(key + seed), G), independent of the home-bucket hash (low word).  Rank r keeps the slots of
//      its keys only, so the slot arrays -- 8/9 of the index -- shrink by G.
//
// One query batch of a rank (its own shard of the reads), all device-side except two count exchanges:
//   1. route: dest = owner(sig) per (read, table) key, counting partition by dest (histogram, block-wise reservation
//      of output ranges, scatter), keys + table ids gathered into send order;
//   2. counts all-gather (G x G matrix) -> host; all-to-all-v of keys (8 B) and table ids (1 B);
//   3. owner: probe its shard (one 64-B bucket access per lookup, as K3b), counts back (4 B per key);
//   4. owner: gather the value lists contiguously per origin; value totals all-gather -> host;
//      all-to-all-v of values (4 B each);
//   5. origin: per-read totals, offsets and the values scattered into table order -- byte-identical to what
//      the replicated index's retrieve produces, so K4/K5 run unchanged.
// NCCL is resolved at run time (dlopen libnccl.so.2: in a torch process that is the library torch uses).
#include "pipeline.cuh"
#include "k3_table.cuh"
#include "mapper.hpp"
#include "partition.cuh"
#include <dlfcn.h>
#include <string.h>
#include <vector>

namespace hrm {

// ---- NCCL, resolved lazily -------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("NCCL is not available: %s", dlerror());
        return nullptr;
    }
#define HRM_NCCL_SYM(name)                                              \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); \
    if (!api.name) {                                                    \
        set_error("NCCL symbol nccl" #name " not found");               \
        return nullptr;                                                 \
    }
    HRM_NCCL_SYM(GetUniqueId)
    HRM_NCCL_SYM(CommInitRank)
    HRM_NCCL_SYM(CommDestroy)
    HRM_NCCL_SYM(GroupStart)
    HRM_NCCL_SYM(GroupEnd)
    HRM_NCCL_SYM(Send)
    HRM_NCCL_SYM(Recv)
    HRM_NCCL_SYM(AllGather)
    HRM_NCCL_SYM(GetErrorString)
#undef HRM_NCCL_SYM
    api.handle = h;
    return &api;
}

#define HRM_NCCL(api, call)                                                                             \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            ::hrm::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
            return HRM_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

} // namespace hrm

struct hrm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    hrm::NcclApi* api = nullptr;
    int64_t bytes_sent = 0, bytes_received = 0, exchanges = 0; // data-path traffic since creation
};

namespace hrm {

// max over the ranks of one host value (collective; synchronises the stream)
hrm_status comm_max_i64(hrm_comm* c, int64_t* v, cudaStream_t s)
{
    Scratch buf;
    HRM_TRY(buf.alloc(sizeof(int64_t) * (size_t)(c->world + 1), s));
    int64_t* d = buf.as<int64_t>();
    HRM_CUDA(cudaMemcpyAsync(d, v, sizeof(int64_t), cudaMemcpyHostToDevice, s));
    HRM_NCCL(c->api, c->api->AllGather(d, d + 1, 1, ncclInt64, c->comm, s));
    std::vector<int64_t> h((size_t)c->world);
    HRM_CUDA(cudaMemcpyAsync(h.data(), d + 1, sizeof(int64_t) * h.size(), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    for (int64_t x : h) *v = x > *v ? x : *v;
    return HRM_OK;
}

// collective agreement on a status: every rank returns the first non-OK status any rank holds (synchronises the
// stream).  A rank that failed locally (allocation, argument) never leaves its peers inside a collective.
hrm_status comm_agree(hrm_comm* c, hrm_status local, cudaStream_t s)
{
    int64_t v = -(int64_t)local; // statuses are <= 0
    hrm_status st = comm_max_i64(c, &v, s);
    if (st != HRM_OK) return st;
    if (v != 0 && local == HRM_OK) set_error("another rank failed with status %lld", (long long)-v);
    return (hrm_status)(-v);
}

int comm_rank(const hrm_comm* c) { return c ? c->rank : 0; }
int comm_world(const hrm_comm* c) { return c ? c->world : 1; }

// byte-wise all-to-all-v: counts / offsets in elements of `elem` bytes
static hrm_status alltoallv(hrm_comm* c, const void* send, const int64_t* scount, const int64_t* soff, void* recv,
                            const int64_t* rcount, const int64_t* roff, size_t elem, cudaStream_t s)
{
    NcclApi* api = c->api;
    HRM_NCCL(api, api->GroupStart());
    for (int p = 0; p < c->world; p++) {
        if (scount[p] > 0)
            HRM_NCCL(api, api->Send((const char*)send + (size_t)soff[p] * elem, (size_t)scount[p] * elem, ncclUint8, p,
                                    c->comm, s));
        if (rcount[p] > 0)
            HRM_NCCL(api, api->Recv((char*)recv + (size_t)roff[p] * elem, (size_t)rcount[p] * elem, ncclUint8, p,
                                    c->comm, s));
        if (p != c->rank) {
            c->bytes_sent += scount[p] * (int64_t)elem;
            c->bytes_received += rcount[p] * (int64_t)elem;
        }
    }
    HRM_NCCL(api, api->GroupEnd());
    c->exchanges++;
    return HRM_OK;
}

// ---- kernels ---------------------------------------------------------------------------------------
// dest[e] = owner of key e (invalid signatures stay at home: they count 0 anywhere); hist[p] += 1
__global__ void __launch_bounds__(256) route_dest_kernel(const uint64_t* __restrict__ sigs, int64_t total, int rank,
                                                         int world, uint8_t* __restrict__ dest,
                                                         unsigned long long* __restrict__ hist)
{
    __shared__ unsigned int sh[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) sh[t] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t key = sigs[e];
        const int d = key == SLOT_EMPTY ? rank : (int)key_owner(key, (uint32_t)world);
        dest[e] = (uint8_t)d;
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < world; t += blockDim.x)
        if (sh[t]) atomicAdd(&hist[t], (unsigned long long)sh[t]);
}

// counting partition by destination: perm[q] = lookup e, lookups of the same destination contiguous (any order inside
// a destination: answers come back in the order they were sent).  A block counts its items per destination in shared
// memory, reserves its ranges with one atomic per destination and scatters.
__global__ void __launch_bounds__(256) route_partition_kernel(const uint8_t* __restrict__ dest, int64_t total, int world,
                                                              const unsigned long long* __restrict__ hist,
                                                              unsigned long long* __restrict__ cursor,
                                                              uint32_t* __restrict__ perm)
{
    __shared__ unsigned int cnt[256], base_lo[256], soff_s[256];
    __shared__ unsigned long long start[256];
    constexpr int ITEMS = 16;
    const int64_t tile = (int64_t)blockDim.x * ITEMS;
    for (int64_t t0 = (int64_t)blockIdx.x * tile; t0 < total; t0 += (int64_t)gridDim.x * tile) {
        for (int t = threadIdx.x; t < 256; t += blockDim.x) cnt[t] = 0u;
        __syncthreads();
        unsigned int mypos[ITEMS];
        uint8_t myd[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int64_t e = t0 + (int64_t)k * blockDim.x + threadIdx.x;
            myd[k] = 0;
            mypos[k] = 0u;
            if (e < total) {
                myd[k] = dest[e];
                mypos[k] = atomicAdd(&cnt[myd[k]], 1u);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < world) {
            unsigned long long off = 0; // exclusive prefix of the global histogram = start of the destination's range
            for (int p = 0; p < (int)threadIdx.x; p++) off += hist[p];
            start[threadIdx.x] = off + atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const int64_t e = t0 + (int64_t)k * blockDim.x + threadIdx.x;
            if (e < total) perm[start[myd[k]] + mypos[k]] = (uint32_t)e;
        }
        __syncthreads();
    }
    (void)base_lo;
    (void)soff_s;
}

// send order: keys and table ids of the routed lookups
__global__ void __launch_bounds__(256) route_gather_kernel(const uint64_t* __restrict__ sigs,
                                                           const uint32_t* __restrict__ perm, int64_t total, int H,
                                                           uint64_t* __restrict__ keys, uint8_t* __restrict__ tabs)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const uint32_t e = perm[q];
        keys[q] = sigs[e];
        tabs[q] = (uint8_t)(e % (uint32_t)H);
    }
}

// owner side: one lookup per thread against the local shard (the bucket access of K3b)
__global__ void __launch_bounds__(256) probe_keys_kernel(const uint64_t* __restrict__ keys,
                                                         const uint8_t* __restrict__ tabs, int64_t m,
                                                         const TablesParam* __restrict__ tabs_g, int H,
                                                         uint32_t max_results, uint2* __restrict__ ranges,
                                                         int32_t* __restrict__ counts,
                                                         unsigned long long* __restrict__ touches_out)
{
    __shared__ TableRef tr[MAX_TABLES];
    for (int t = threadIdx.x; t < H; t += blockDim.x) tr[t] = tabs_g->t[t];
    __syncthreads();
    uint32_t visited = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += stride) {
        const TableRef T = tr[tabs[q]];
        const uint2 r = probe_bucket_sequence(T.slots, T.nbuckets, keys[q], max_results, visited);
        ranges[q] = r;
        counts[q] = (int32_t)r.y;
    }
    visited *= BUCKET_SLOTS;
    for (int d = 16; d > 0; d >>= 1) visited += __shfl_xor_sync(0xffffffffu, visited, d);
    if ((threadIdx.x & 31) == 0 && visited && touches_out) atomicAdd(touches_out, (unsigned long long)visited);
}

// owner side: value lists in received order (contiguous per origin); a warp copies one list, coalesced
__global__ void __launch_bounds__(256) gather_values_kernel(const uint2* __restrict__ ranges,
                                                            const int32_t* __restrict__ voff, int64_t m,
                                                            const uint32_t* __restrict__ table_values,
                                                            uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp0; q < m; q += nwarps) {
        const uint2 r = ranges[q];
        const int64_t w = voff[q];
        for (uint32_t v = lane; v < r.y; v += 32) out[w + v] = __ldg(table_values + r.x + v);
    }
}

// picks src[idx[t]] for a handful of positions (segment boundaries of a scan)
__global__ void pick_kernel(const int32_t* __restrict__ src, const int64_t* __restrict__ idx, int n,
                            int64_t* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int64_t)src[idx[t]];
}

// origin side: counts / source offsets back in (read, table) order
__global__ void __launch_bounds__(256) unsort_kernel(const uint32_t* __restrict__ perm,
                                                     const int32_t* __restrict__ cnt_back,
                                                     const int32_t* __restrict__ src_off, int64_t total,
                                                     int32_t* __restrict__ cnt_e, int32_t* __restrict__ src_e)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const uint32_t e = perm[q];
        cnt_e[e] = cnt_back[q];
        src_e[e] = src_off[q];
    }
}

__global__ void __launch_bounds__(256) read_totals_kernel(const int32_t* __restrict__ cnt_e, int n, int H,
                                                          int32_t* __restrict__ num_per_seq)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int sum = 0;
        for (int j = 0; j < H; j++) sum += cnt_e[i * H + j];
        num_per_seq[i] = sum;
    }
}

// values of read i: buckets of tables 0..H-1 concatenated at out[offsets[i] ...] (the layout of retrieve_kernel)
__global__ void __launch_bounds__(256) scatter_values_kernel(const int32_t* __restrict__ cnt_e,
                                                             const int32_t* __restrict__ src_e,
                                                             const int32_t* __restrict__ offsets, int n, int H,
                                                             const uint32_t* __restrict__ recv_values,
                                                             uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarps) {
        int64_t w = offsets[i];
        for (int t0 = 0; t0 < H; t0 += 32) {
            const int t = t0 + lane;
            const int cnt = t < H ? cnt_e[i * H + t] : 0;
            const int src = t < H ? src_e[i * H + t] : 0;
            int incl = cnt;
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - cnt;
            if (total <= 64) {
                for (int v = 0; v < cnt; v++) out[w + excl + v] = recv_values[src + v];
            } else {
                for (int tt = 0; tt < 32; tt++) {
                    const int c2 = __shfl_sync(0xffffffffu, cnt, tt);
                    const int s2 = __shfl_sync(0xffffffffu, src, tt);
                    const int ex = __shfl_sync(0xffffffffu, excl, tt);
                    for (int v = lane; v < c2; v += 32) out[w + ex + v] = recv_values[s2 + v];
                }
            }
            w += total;
        }
    }
}

// (offset into the scattered value array, count) of every (read, table) bucket: what the fused collection takes
__global__ void __launch_bounds__(256) value_ranges_kernel(const int32_t* __restrict__ cnt_e,
                                                           const int32_t* __restrict__ offsets, int n, int H,
                                                           uint2* __restrict__ ranges)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarps) {
        int w = offsets[i];
        for (int t0 = 0; t0 < H; t0 += 32) {
            const int t = t0 + lane;
            const int cnt = t < H ? cnt_e[i * H + t] : 0;
            int incl = cnt;
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            if (t < H) ranges[i * H + t] = make_uint2((uint32_t)(w + incl - cnt), (uint32_t)cnt);
            w += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

static unsigned pgrid(int64_t items)
{
    int64_t g = HRM_SDIV(items, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ---- the routed query ------------------------------------------------------------------------------
hrm_status partitioned_query(hrm_comm* c, hrm_minhasher* mh, const uint64_t* d_sigs, int n, int32_t* d_num_per_seq,
                             int32_t* d_offsets, int64_t* h_total, Scratch& values, StageTimer& T, cudaStream_t s,
                             uint2* d_ranges)
{
    const int H = mh->H, G = c->world;
    NcclApi* api = c->api;
    const int64_t total = (int64_t)n * H;
    HRM_REQUIRE(total < (1LL << 31), "batch too large: n * hashmaps must fit int");
    HRM_REQUIRE(G <= 255, "at most 255 ranks");
    // 1. route
    T.begin(HRM_STAGE_ROUTE, s);
    Scratch dest, perm, small, skeys, stabs;
    const size_t tt = (size_t)(total > 0 ? total : 1);
    hrm_status ast = HRM_OK; // allocation status: agreed on collectively before the first exchange
    auto A = [&](hrm_status st) {
        if (ast == HRM_OK && st != HRM_OK) ast = st;
    };
    A(dest.alloc(tt, s));
    A(perm.alloc(sizeof(uint32_t) * tt, s));
    A(skeys.alloc(sizeof(uint64_t) * tt, s));
    A(stabs.alloc(tt, s));
    // small: [0..G) hist, [G..G+G*G) count matrix, then value-total row + matrix, then pick indices/outputs, cursors
    const size_t small_words = (size_t)(3 * G + 2 * G * G + 4 * (G + 1) + 16);
    A(small.alloc(sizeof(int64_t) * small_words, s));
    HRM_TRY(comm_agree(c, ast, s)); // every rank returns the same status: nobody is left inside a collective
    int64_t* d_hist = small.as<int64_t>();
    int64_t* d_cmat = d_hist + G;
    int64_t* d_vrow = d_cmat + (size_t)G * G;
    int64_t* d_vmat = d_vrow + G;
    int64_t* d_idx = d_vmat + (size_t)G * G;
    int64_t* d_pick = d_idx + 2 * (G + 1);
    int64_t* d_cursor = d_pick + 2 * (G + 1);
    HRM_CUDA(cudaMemsetAsync(small.p, 0, sizeof(int64_t) * small_words, s));
    if (total > 0) {
        HRM_LAUNCH(route_dest_kernel, pgrid(total), 256, 0, s, d_sigs, total, c->rank, G, dest.as<uint8_t>(),
                   reinterpret_cast<unsigned long long*>(d_hist));
        HRM_LAUNCH(route_partition_kernel, pgrid(HRM_SDIV(total, (int64_t)16)), 256, 0, s, dest.as<uint8_t>(), total, G,
                   reinterpret_cast<const unsigned long long*>(d_hist), reinterpret_cast<unsigned long long*>(d_cursor),
                   perm.as<uint32_t>());
        HRM_LAUNCH(route_gather_kernel, pgrid(total), 256, 0, s, d_sigs, perm.as<uint32_t>(), total, H,
                   skeys.as<uint64_t>(), stabs.as<uint8_t>());
    }
    // 2. counts: every rank learns the whole G x G matrix (row r = what rank r sends to each peer)
    HRM_NCCL(api, api->AllGather(d_hist, d_cmat, (size_t)G, ncclInt64, c->comm, s));
    std::vector<int64_t> cmat((size_t)G * G);
    HRM_CUDA(cudaMemcpyAsync(cmat.data(), d_cmat, sizeof(int64_t) * cmat.size(), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    std::vector<int64_t> scount(G), soff(G + 1, 0), rcount(G), roff(G + 1, 0);
    for (int p = 0; p < G; p++) {
        scount[p] = cmat[(size_t)c->rank * G + p];
        rcount[p] = cmat[(size_t)p * G + c->rank];
        soff[p + 1] = soff[p] + scount[p];
        roff[p + 1] = roff[p] + rcount[p];
    }
    const int64_t m = roff[G];
    // every rank holds the whole matrix, so every rank takes the same decision for every rank's totals
    for (int p = 0; p < G; p++) {
        int64_t mp = 0;
        for (int q = 0; q < G; q++) mp += cmat[(size_t)q * G + p];
        if (mp >= (1LL << 31)) {
            set_error("too many routed lookups for rank %d (%lld): use smaller batches", p, (long long)mp);
            return HRM_ERR_OVERFLOW;
        }
    }
    Scratch rkeys, rtabs, ranges, rcnt, voff, cnt_back;
    const size_t mm = (size_t)(m > 0 ? m : 1);
    A(soff[G] == total ? HRM_OK : HRM_ERR_INVALID);
    A(rkeys.alloc(sizeof(uint64_t) * mm, s));
    A(rtabs.alloc(mm, s));
    A(ranges.alloc(sizeof(uint2) * mm, s));
    A(rcnt.alloc(sizeof(int32_t) * mm, s));
    A(voff.alloc(sizeof(int32_t) * (mm + 1), s));
    A(cnt_back.alloc(sizeof(int32_t) * tt, s));
    HRM_TRY(comm_agree(c, ast, s));
    HRM_TRY(alltoallv(c, skeys.p, scount.data(), soff.data(), rkeys.p, rcount.data(), roff.data(), sizeof(uint64_t), s));
    HRM_TRY(alltoallv(c, stabs.p, scount.data(), soff.data(), rtabs.p, rcount.data(), roff.data(), 1, s));
    T.end(s);
    // 3. owner: probe, counts back
    T.begin(HRM_STAGE_PROBE, s);
    if (m > 0)
        HRM_LAUNCH(probe_keys_kernel, pgrid(m), 256, 0, s, rkeys.as<uint64_t>(), rtabs.as<uint8_t>(), m, mh->d_param, H,
                   (uint32_t)mh->max_results, ranges.as<uint2>(), rcnt.as<int32_t>(), mh->d_touches);
    T.end(s);
    T.begin(HRM_STAGE_ROUTE, s);
    HRM_TRY(alltoallv(c, rcnt.p, rcount.data(), roff.data(), cnt_back.p, scount.data(), soff.data(), sizeof(int32_t), s));
    // 4. owner: value lists contiguous per origin; totals per origin -> matrix -> host
    HRM_TRY(exclusive_scan_i32(rcnt.as<int32_t>(), voff.as<int32_t>(), m, nullptr, s));
    HRM_CUDA(cudaMemcpyAsync(d_idx, roff.data(), sizeof(int64_t) * (size_t)(G + 1), cudaMemcpyHostToDevice, s));
    HRM_LAUNCH(pick_kernel, 1, 256, 0, s, voff.as<int32_t>(), d_idx, G + 1, d_pick);
    std::vector<int64_t> vb(G + 1);
    HRM_CUDA(cudaMemcpyAsync(vb.data(), d_pick, sizeof(int64_t) * (size_t)(G + 1), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    std::vector<int64_t> vsend(G), vsoff(G + 1, 0), vrecv(G), vroff(G + 1, 0), vrow(G);
    for (int p = 0; p < G; p++) {
        vsend[p] = vb[p + 1] - vb[p];
        vsoff[p + 1] = vsoff[p] + vsend[p];
        vrow[p] = vsend[p];
    }
    HRM_CUDA(cudaMemcpyAsync(d_vrow, vrow.data(), sizeof(int64_t) * (size_t)G, cudaMemcpyHostToDevice, s));
    HRM_NCCL(api, api->AllGather(d_vrow, d_vmat, (size_t)G, ncclInt64, c->comm, s));
    std::vector<int64_t> vmat((size_t)G * G);
    HRM_CUDA(cudaMemcpyAsync(vmat.data(), d_vmat, sizeof(int64_t) * vmat.size(), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    // collective decision: values any rank sends (row sums) or receives (column sums) must fit int offsets
    for (int p = 0; p < G; p++) {
        int64_t sent = 0, recv = 0;
        for (int q = 0; q < G; q++) {
            sent += vmat[(size_t)p * G + q];
            recv += vmat[(size_t)q * G + p];
        }
        if (sent > 0x7fffffffLL || recv > 0x7fffffffLL) {
            set_error("candidate values routed through rank %d exceed int (%lld sent, %lld received): smaller batches",
                      p, (long long)sent, (long long)recv);
            return HRM_ERR_OVERFLOW;
        }
    }
    for (int p = 0; p < G; p++) {
        vrecv[p] = vmat[(size_t)p * G + c->rank];
        vroff[p + 1] = vroff[p] + vrecv[p];
    }
    const int64_t vtotal = vroff[G];
    Scratch svals, rvals, src_off, cnt_e, src_e;
    A(svals.alloc(sizeof(uint32_t) * (size_t)(vsoff[G] > 0 ? vsoff[G] : 1), s));
    A(rvals.alloc(sizeof(uint32_t) * (size_t)(vtotal > 0 ? vtotal : 1), s));
    A(src_off.alloc(sizeof(int32_t) * (tt + 1), s));
    A(cnt_e.alloc(sizeof(int32_t) * tt, s));
    A(src_e.alloc(sizeof(int32_t) * tt, s));
    A(values.alloc(sizeof(uint32_t) * (size_t)(vtotal > 0 ? vtotal : 1), s));
    HRM_TRY(comm_agree(c, ast, s));
    if (m > 0)
        HRM_LAUNCH(gather_values_kernel, pgrid(m * 32), 256, 0, s, ranges.as<uint2>(), voff.as<int32_t>(), m, mh->values,
                   svals.as<uint32_t>());
    HRM_TRY(alltoallv(c, svals.p, vsend.data(), vsoff.data(), rvals.p, vrecv.data(), vroff.data(), sizeof(uint32_t), s));
    T.end(s);
    // 5. origin: per-read totals, offsets, values in table order
    T.begin(HRM_STAGE_RETRIEVE, s);
    HRM_TRY(exclusive_scan_i32(cnt_back.as<int32_t>(), src_off.as<int32_t>(), total, nullptr, s));
    if (total > 0) {
        HRM_LAUNCH(unsort_kernel, pgrid(total), 256, 0, s, perm.as<uint32_t>(), cnt_back.as<int32_t>(),
                   src_off.as<int32_t>(), total, cnt_e.as<int32_t>(), src_e.as<int32_t>());
        HRM_LAUNCH(read_totals_kernel, pgrid(n), 256, 0, s, cnt_e.as<int32_t>(), n, H, d_num_per_seq);
    }
    HRM_TRY(exclusive_scan_i32(d_num_per_seq, d_offsets, n, nullptr, s));
    if (vtotal > 0)
        HRM_LAUNCH(scatter_values_kernel, pgrid((int64_t)n * 32), 256, 0, s, cnt_e.as<int32_t>(), src_e.as<int32_t>(),
                   d_offsets, n, H, rvals.as<uint32_t>(), values.as<uint32_t>());
    if (d_ranges && total > 0)
        HRM_LAUNCH(value_ranges_kernel, pgrid((int64_t)n * 32), 256, 0, s, cnt_e.as<int32_t>(), d_offsets, n, H, d_ranges);
    T.end(s);
    *h_total = vtotal;
    return HRM_OK;
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_comm_unique_id(void* out_id, int64_t capacity)
{
    HRM_REQUIRE(out_id != nullptr && capacity >= (int64_t)HRM_COMM_ID_BYTES, "unique id buffer needs HRM_COMM_ID_BYTES");
    NcclApi* api = nccl_api();
    if (!api) return HRM_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == HRM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    HRM_NCCL(api, api->GetUniqueId(&id));
    memcpy(out_id, &id, sizeof id);
    return HRM_OK;
}

extern "C" hrm_status hrm_comm_create(hrm_comm** out, int rank, int world, const void* id_bytes)
{
    HRM_REQUIRE(out != nullptr && id_bytes != nullptr && world >= 1 && rank >= 0 && rank < world, "args");
    *out = nullptr;
    HRM_TRY(ensure_device());
    NcclApi* api = nccl_api();
    if (!api) return HRM_ERR_CUDA;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    auto* c = new hrm_comm;
    c->api = api;
    c->rank = rank;
    c->world = world;
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return HRM_ERR_CUDA;
    }
    *out = c;
    return HRM_OK;
}

extern "C" void hrm_comm_destroy(hrm_comm* c)
{
    if (!c) return;
    if (c->comm && c->api) c->api->CommDestroy(c->comm);
    delete c;
}

extern "C" hrm_status hrm_comm_info(const hrm_comm* c, hrm_comm_info_t* out)
{
    HRM_REQUIRE(c != nullptr && out != nullptr, "args");
    out->rank = c->rank;
    out->world = c->world;
    out->bytes_sent = c->bytes_sent;
    out->bytes_received = c->bytes_received;
    out->exchanges = c->exchanges;
    return HRM_OK;
}

extern "C" int hrm_key_owner(uint64_t key, int world)
{
    return world > 0 ? (int)key_owner(key, (uint32_t)world) : 0;
}
