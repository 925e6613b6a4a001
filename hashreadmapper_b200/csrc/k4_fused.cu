// k4_fused.cu -- K3b value retrieval + K4 candidate collection in one step, for the replicated index.
// ref: FakeGpuMinhasher::retrieveValues include/gpu/fakegpuminhasher.cuh:312-392 (copy every bucket of every
//      query into one list) followed by GpuMinhashQueryFilter::keepDistinctByFrequency
//      include/gpu/minhashqueryfilter.cuh:239-278 -> GpuSegmentedUniqueByCount::unique
//      include/gpu/cuda_unique_by_count.cuh:33-215 (segmented radix sort + run-length count + threshold).
// Result per read: the ascending list of ids that occur in at least minTableHits of its H buckets -- the same
// list the reference builds.  How: the H buckets of a read are never copied.  Every bucket is an ascending id
// list inside the index (stable sort at build time), so
//   * multiplicities are counted in a shared-memory table straight from the index (atomicCAS on the id,
//     atomicAdd on its count);
//   * a read with more ids than one table holds is cut into id RANGES: a range is a contiguous piece of each
//     bucket (binary search), ranges are visited in ascending order, so their sorted survivors concatenate
//     into the sorted result;
//   * the L = min(2, minTableHits - 2) LARGEST buckets are not enumerated at all: an id with multiplicity >=
//     minTableHits still reaches minTableHits - L in the other buckets, and only ids that do are looked up in
//     the big buckets (binary search).  Bucket sizes are heavy-tailed on a 3-letter genome (T-rich k-mers), so
//     this removes a large share of the ids.
// At human-genome scale a read retrieves ~4000 ids per pass and keeps 1-3; this path moves each id once
// (index -> shared memory) instead of index -> value list -> partition scratch -> table.
// Anything that does not fit (more than COLLECT_FINAL_CAP survivors for one read, output space exhausted,
// minTableHits < 2) is reported through the overflow flag and the caller reruns the batch on the general
// path (retrieve + filter_segments).
#include "pipeline.cuh"
#include "k3_table.cuh"
#include "k4_sort.cuh"
#include <stdlib.h>

namespace hrm {

constexpr int COLLECT_THREADS = 256;     // warp-per-read kernel
constexpr int COLLECT_BIG_THREADS = 256; // block-per-read kernel (128 and 512 threads measured slower: 247 and 148 ms vs 96 ms per 1M reads)
constexpr int COLLECT_FINAL_CAP = 1024; // survivors per read the block kernel can hold
constexpr int COLLECT_MLP = 4;         // independent id loads a thread keeps in flight
constexpr int COLLECT_CHUNK = 16;       // id ranges whose bucket boundaries are searched together
constexpr uint32_t COLLECT_EMPTY = 0xFFFFFFFFu;

struct CollectParams {
    const uint2* ranges;        // (value offset, count) of the last probe; (q, t) at q * rq + t * rt
    int64_t rq, rt;
    const uint32_t* table_values;
    int n, H, min_hits;
    uint32_t id_space;          // ids are < id_space
    uint32_t* out;              // candidate lists, allocated from `cursor`
    unsigned long long out_cap;
    unsigned long long* cursor;
    int2* lists;                // per read: (start in out, count)
    int* overflow;
    int32_t* big_list;
    int32_t* big_count;
    int32_t* big2_list;         // reads the block-wide filter kernel hands on to the counting-table kernel
    int32_t* big2_count;
    int xslots_warp, xslots_block; // exact-table slots of the duplicate-detection kernels (powers of two)
    int* work;                  // work counter of the block variant
    int warp_cap;               // reads with more ids go straight to the block kernel (test hook)
    int warp_slots;             // table slots of a warp (power of two)
    int slots;                  // table slots of the block kernel (power of two)
    int fill;                   // target ids per range
    unsigned long long* stats;  // [0] ids enumerated, [1] ids skipped (largest buckets), [2] ranges, [3] reads -> block kernel
};

__device__ __forceinline__ uint32_t collect_hash(uint32_t v) { return v * 0x9E3779B1u; }

// Counting table.  PACKED (ids < 2^26): one word per slot, id << 6 | count (count <= 63 >= any number of hash
// tables), so an id costs ONE atomic -- the CAS that claims the slot or the add that bumps it.  Otherwise two
// words per slot (id, count).  A slot is empty iff its (first) word is all ones.
constexpr int COLLECT_PACK_BITS = 6;
template <bool PACKED>
struct CountTable {
    uint32_t* w;  // PACKED: [slots]; else ids [cap] followed by counts [cap]
    int cap;      // allocated slots (power of two)
    __device__ __forceinline__ void clear_all(int tid, int nthr) const
    {
        for (int i = tid; i < cap; i += nthr) {
            w[i] = COLLECT_EMPTY;
            if (!PACKED) w[cap + i] = 0u;
        }
    }
    // counts one id and returns its new count
    __device__ __forceinline__ uint32_t count(uint32_t mask, int shift, uint32_t v) const
    {
        uint32_t h = collect_hash(v) >> shift;
        if (PACKED) {
            while (true) {
                uint32_t cur = w[h];
                if (cur == COLLECT_EMPTY) {
                    cur = atomicCAS(&w[h], COLLECT_EMPTY, (v << COLLECT_PACK_BITS) | 1u);
                    if (cur == COLLECT_EMPTY) return 1u;
                }
                if ((cur >> COLLECT_PACK_BITS) == v)
                    return (atomicAdd(&w[h], 1u) & ((1u << COLLECT_PACK_BITS) - 1u)) + 1u;
                h = (h + 1) & mask;
            }
        } else {
            while (true) {
                const uint32_t prev = atomicCAS(&w[h], COLLECT_EMPTY, v);
                if (prev == COLLECT_EMPTY || prev == v) return atomicAdd(&w[cap + h], 1u) + 1u;
                h = (h + 1) & mask;
            }
        }
    }
    // count of an id that is in the table
    __device__ __forceinline__ uint32_t lookup(uint32_t mask, int shift, uint32_t v) const
    {
        uint32_t h = collect_hash(v) >> shift;
        while (true) {
            const uint32_t cur = w[h];
            if (cur == COLLECT_EMPTY) return 0u;
            if (PACKED) {
                if ((cur >> COLLECT_PACK_BITS) == v) return cur & ((1u << COLLECT_PACK_BITS) - 1u);
            } else if (cur == v) {
                return w[cap + h];
            }
            h = (h + 1) & mask;
        }
    }
    // empties slots [0, n), n a multiple of 128: 128-bit stores
    __device__ __forceinline__ void clear_prefix(int n, int tid, int nthr) const
    {
        uint4* p = reinterpret_cast<uint4*>(w);
        for (int i = tid; i < n / 4; i += nthr) p[i] = make_uint4(COLLECT_EMPTY, COLLECT_EMPTY, COLLECT_EMPTY, COLLECT_EMPTY);
        if (!PACKED) {
            uint4* q = reinterpret_cast<uint4*>(w + cap);
            for (int i = tid; i < n / 4; i += nthr) q[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // reads slot i and leaves it empty; returns false for an empty slot
    __device__ __forceinline__ bool take(int i, uint32_t& id, uint32_t& cnt) const
    {
        const uint32_t x = w[i];
        if (x == COLLECT_EMPTY) return false;
        w[i] = COLLECT_EMPTY;
        if (PACKED) {
            id = x >> COLLECT_PACK_BITS;
            cnt = x & ((1u << COLLECT_PACK_BITS) - 1u);
        } else {
            id = x;
            cnt = w[cap + i];
            w[cap + i] = 0u;
        }
        return true;
    }
};

// first position in [lo, hi) of the ascending list p whose id is >= key
__device__ __forceinline__ int collect_lower_bound(const uint32_t* __restrict__ p, int lo, int hi, uint32_t key)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// The same for a bucket of cnt ids spread over [0, id_space): window ids of a bucket are close to uniform, so
// the answer lies within a few sqrt(cnt) of cnt * key / id_space.  One probe at the estimate, a gallop away from
// it, then the bisection of a bracket that sits in one or two sectors: 3-4 dependent DRAM accesses instead of
// log2(cnt).  Any distribution is handled (the gallop doubles), uniformity only makes it fast.
__device__ __forceinline__ int collect_lower_bound_interp(const uint32_t* __restrict__ p, int cnt, uint32_t key,
                                                          uint32_t id_space)
{
    if (cnt <= 0 || key == 0u) return 0;
    int lo = 0, hi = cnt; // the answer is in [lo, hi]
    // the estimate only steers the search: single precision is enough and avoids a 64-bit division
    int pos = (int)(((float)key / (float)id_space) * (float)cnt);
    pos = pos >= cnt ? cnt - 1 : (pos < 0 ? 0 : pos);
    int step = 8;
    while (step * step < cnt) step <<= 1; // ~ sqrt(cnt), at least one sector
    if (__ldg(p + pos) < key) {
        lo = pos + 1;
        while (true) {
            const int nx = lo + step - 1;
            if (nx >= cnt) break;
            if (__ldg(p + nx) < key) {
                lo = nx + 1;
                step <<= 1;
            } else {
                hi = nx;
                break;
            }
        }
    } else {
        hi = pos;
        while (true) {
            const int nx = hi - step;
            if (nx < 0) break;
            if (__ldg(p + nx) < key) {
                lo = nx + 1;
                break;
            }
            hi = nx;
            step <<= 1;
        }
    }
    return collect_lower_bound(p, lo, hi, key);
}

__device__ __forceinline__ void collect_sort_warp(uint32_t* s, int cnt, int lane)
{
    // odd-even transposition is enough for the handful of survivors a warp sees
    for (int round = 0; round < cnt; round++) {
        for (int i = 2 * lane + (round & 1); i + 1 < cnt; i += 64) {
            const uint32_t a = s[i], b = s[i + 1];
            if (a > b) {
                s[i] = b;
                s[i + 1] = a;
            }
        }
        __syncwarp();
    }
}

// ---- warp per read ---------------------------------------------------------------------------------
// The whole scheme (skip the L largest buckets, id ranges, count, look the survivors up in the big buckets) run by
// ONE WARP per read on its own slice of shared memory: no block barriers, and an SM keeps as many reads in flight
// as it has resident warps, which is what hides the DRAM latency of a path that is a chain of dependent accesses
// (bucket ranges -> range boundaries -> ids -> big-bucket lookups).  Reads whose ids are far from uniform or that
// keep more survivors than the slice holds go to the block kernel through big_list.
constexpr int COLLECT_WCHUNK = 4;   // id ranges whose boundaries a warp searches together
constexpr int COLLECT_WCAND = 128;  // ids reaching the reduced threshold per range a warp can hold
constexpr int COLLECT_WFIN = 128;   // survivors per read a warp can hold
constexpr int COLLECT_WMLP = 4;     // independent id loads per lane

__host__ __device__ inline size_t collect_warp_slice_words(int slots, int H, bool packed)
{
    const size_t w = (size_t)(packed ? 1 : 2) * slots + 2 * COLLECT_WCAND + COLLECT_WFIN + 3 * (size_t)H +
                     (size_t)(COLLECT_WCHUNK + 1) * H + (size_t)H + 2;
    return (w + 3) & ~(size_t)3; // slices stay 16-byte aligned (128-bit table clears)
}

template <bool PACKED>
__global__ void __launch_bounds__(COLLECT_THREADS) collect_warp_kernel(CollectParams P)
{
    extern __shared__ __align__(16) uint32_t wdyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int H = P.H, T = P.min_hits, SW = P.warp_slots;
    uint32_t* base = wdyn + (size_t)wid * collect_warp_slice_words(SW, H, PACKED);
    const CountTable<PACKED> tab{base, SW};
    uint32_t* cand = base + (PACKED ? 1 : 2) * SW; // [WCAND]
    uint32_t* ccnt = cand + COLLECT_WCAND;          // [WCAND]
    uint32_t* fin = ccnt + COLLECT_WCAND;           // [WFIN]
    uint32_t* offv = fin + COLLECT_WFIN;            // [H]
    int* cntv = reinterpret_cast<int*>(offv + H);   // [H]
    int* skipf = cntv + H;                          // [H]
    int* bnd = skipf + H;                           // [(WCHUNK + 1) * H]
    int* gpre = bnd + (COLLECT_WCHUNK + 1) * H;     // [H + 1]
    int* wcount = gpre + H + 1;                     // candidates of the current range
    const int L = T - 2 < 2 ? (T - 2 < 0 ? 0 : T - 2) : 2; // buckets not enumerated
    const int thr = T - L;
    const int warp0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    unsigned long long st_enum = 0, st_skip = 0, st_ranges = 0;
    tab.clear_all(lane, 32); // a range leaves the slots it used empty again
    if (lane == 0) *wcount = 0;
    __syncwarp();
    for (int rd = warp0; rd < P.n; rd += nwarps) {
        // bucket ranges of the read
        int total = 0;
        for (int t0 = 0; t0 < H; t0 += 32) {
            const int t = t0 + lane;
            const uint2 r = t < H ? P.ranges[(int64_t)rd * P.rq + (int64_t)t * P.rt] : make_uint2(0u, 0u);
            if (t < H) {
                offv[t] = r.x;
                cntv[t] = (int)r.y;
                skipf[t] = 0;
            }
            total += __reduce_add_sync(0xffffffffu, (int)r.y);
        }
        __syncwarp();
        if (total < T) {
            if (lane == 0) {
                P.lists[rd] = make_int2(0, 0);
                st_enum += (unsigned long long)total;
            }
            continue;
        }
        if (total > P.warp_cap) { // test hook / safety valve: straight to the block kernel
            if (lane == 0) P.big_list[atomicAdd(P.big_count, 1)] = rd;
            continue;
        }
        // the L largest buckets are looked up, not enumerated
        int skipped = 0;
        for (int l = 0; l < L; l++) {
            unsigned best = 0u; // (count << 8 | table) of the lane's best candidate
            for (int t = lane; t < H; t += 32)
                if (!skipf[t] && cntv[t] > 0) {
                    const unsigned key = ((unsigned)cntv[t] << 8) | (unsigned)(255 - t); // ties: lowest table
                    best = key > best ? key : best;
                }
            best = __reduce_max_sync(0xffffffffu, best);
            if (best == 0u) break;
            const int t = 255 - (int)(best & 255u);
            if (lane == 0) skipf[t] = 1;
            skipped += (int)(best >> 8);
            __syncwarp();
        }
        const int E = total - skipped;
        const int fill = (3 * SW) / 8;
        const int nranges = E > 0 ? (E + fill - 1) / fill : 1;
        const uint64_t width = ((uint64_t)P.id_space + (uint64_t)nranges - 1) / (uint64_t)nranges;
        if (lane == 0) {
            st_enum += (unsigned long long)E;
            st_skip += (unsigned long long)skipped;
            st_ranges += (unsigned long long)nranges;
        }
        int nfin = 0;
        bool bad = false;
        for (int g0 = 0; g0 < nranges && !bad; g0 += COLLECT_WCHUNK) {
            const int ng = (nranges - g0) < COLLECT_WCHUNK ? (nranges - g0) : COLLECT_WCHUNK;
            for (int x = lane; x < (ng + 1) * H; x += 32) { // range boundaries, all searches of the chunk in flight
                const int g = x / H, t = x - g * H;
                int pos = 0;
                if (!skipf[t]) {
                    const uint64_t key = (uint64_t)(g0 + g) * width;
                    pos = key >= (uint64_t)P.id_space
                              ? cntv[t]
                              : collect_lower_bound_interp(P.table_values + offv[t], cntv[t], (uint32_t)key, P.id_space);
                }
                bnd[g * H + t] = pos;
            }
            __syncwarp();
            for (int g = 0; g < ng && !bad; g++) {
                int gtotal = 0; // flat prefix of the range's bucket pieces
                for (int t0 = 0; t0 < H; t0 += 32) {
                    const int t = t0 + lane;
                    const int len = t < H ? bnd[(g + 1) * H + t] - bnd[g * H + t] : 0;
                    int incl = len;
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    if (t < H) gpre[t] = gtotal + incl - len;
                    gtotal += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) gpre[H] = gtotal;
                __syncwarp();
                if (gtotal < thr) continue;
                if (8 * gtotal > 7 * SW) { // ids far from uniform: this range would clog the table
                    bad = true;
                    break;
                }
                int slots = 128, shift = 32 - 7;
                while (slots < SW && 5 * slots < 8 * gtotal) {
                    slots <<= 1;
                    shift--;
                }
                // flat index e -> bucket piece: pbase + e is the id's position while e < pnext (e only grows per lane)
                int tcur = 0, pnext = 0;
                int64_t pbase = 0;
                for (int e0 = 0; e0 < gtotal; e0 += 32 * COLLECT_WMLP) {
                    uint32_t v[COLLECT_WMLP];
#pragma unroll
                    for (int u = 0; u < COLLECT_WMLP; u++) {
                        const int e = e0 + u * 32 + lane;
                        v[u] = COLLECT_EMPTY;
                        if (e < gtotal) {
                            if (e >= pnext) {
                                while (gpre[tcur + 1] <= e) tcur++;
                                pnext = gpre[tcur + 1];
                                pbase = (int64_t)offv[tcur] + bnd[g * H + tcur] - gpre[tcur];
                            }
                            v[u] = __ldg(P.table_values + pbase + e);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < COLLECT_WMLP; u++)
                        if (v[u] != COLLECT_EMPTY && tab.count((uint32_t)slots - 1u, shift, v[u]) == (uint32_t)thr) {
                            // the id just reached the reduced threshold: a candidate (once per id)
                            const int at = atomicAdd(wcount, 1);
                            if (at < COLLECT_WCAND) cand[at] = v[u];
                        }
                }
                __syncwarp();
                const int ncand = *wcount;
                __syncwarp();
                if (ncand > 0 && ncand <= COLLECT_WCAND)
                    for (int c = lane; c < ncand; c += 32) ccnt[c] = tab.lookup((uint32_t)slots - 1u, shift, cand[c]);
                if (lane == 0) *wcount = 0;
                __syncwarp();
                tab.clear_prefix(slots, lane, 32); // leave the table empty (128-bit stores; slots >= 128)
                __syncwarp();
                if (ncand == 0) continue; // the common case: nothing in this id range reaches the threshold
                if (ncand > COLLECT_WCAND) {
                    bad = true;
                    break;
                }
                int gf = 0;
                for (int c0 = 0; c0 < ncand; c0 += 32) {
                    const int c = c0 + lane;
                    bool keep = false;
                    uint32_t id = 0;
                    if (c < ncand) {
                        id = cand[c];
                        uint32_t m = ccnt[c];
                        for (int t = 0; t < H && m < (uint32_t)T; t++) {
                            if (!skipf[t]) continue;
                            const uint32_t* p = P.table_values + offv[t];
                            const int pos = collect_lower_bound_interp(p, cntv[t], id, P.id_space);
                            if (pos < cntv[t] && __ldg(p + pos) == id) m++;
                        }
                        keep = m >= (uint32_t)T;
                    }
                    const unsigned m2 = __ballot_sync(0xffffffffu, keep);
                    const int at = nfin + gf + __popc(m2 & ((1u << lane) - 1u));
                    if (keep && at < COLLECT_WFIN) fin[at] = id;
                    gf += __popc(m2);
                }
                __syncwarp();
                if (nfin + gf > COLLECT_WFIN) {
                    bad = true;
                    break;
                }
                // ranges ascend, so sorting each range's survivors keeps the read's list sorted
                if (gf > 1) collect_sort_warp(fin + nfin, gf, lane);
                nfin += gf;
            }
        }
        if (bad) { // the aborted range may have left ids behind; the block kernel redoes the read
            __syncwarp();
            tab.clear_all(lane, 32);
            if (lane == 0) *wcount = 0;
            if (lane == 0) {
                P.big_list[atomicAdd(P.big_count, 1)] = rd;
                st_enum -= (unsigned long long)E; // counted again there
                st_skip -= (unsigned long long)skipped;
                st_ranges -= (unsigned long long)nranges;
            }
            __syncwarp();
            continue;
        }
        unsigned long long start = 0;
        if (lane == 0 && nfin > 0) start = atomicAdd(P.cursor, (unsigned long long)nfin);
        start = __shfl_sync(0xffffffffu, start, 0);
        if (start + (unsigned long long)nfin > P.out_cap) {
            if (lane == 0) *P.overflow = 1;
            nfin = 0;
        }
        for (int i = lane; i < nfin; i += 32) P.out[start + i] = fin[i];
        if (lane == 0) P.lists[rd] = make_int2((int)start, nfin);
        __syncwarp();
    }
    if (lane == 0 && P.stats) {
        if (st_enum) atomicAdd(P.stats, st_enum);
        if (st_skip) atomicAdd(P.stats + 1, st_skip);
        if (st_ranges) atomicAdd(P.stats + 2, st_ranges);
    }
}

// ---- duplicate detection: blocked Bloom filter in front of a small exact table ---------------------------------
// An id with multiplicity >= T over the H buckets of a read has multiplicity >= T - L >= 2 in the H - L smallest
// ("enumerated") buckets, i.e. it is a DUPLICATE there.  At human-genome scale a read-pass walks ~2000 enumerated ids
// of which ~100 are second or later occurrences (low-complexity windows that share many sketches) and 1-3 survive.
//   phase 1  the enumerated buckets are streamed once, front to back, with 128-bit loads.  Every id sets three bits
//            of ONE 32-bit word of a shared-memory bitmap with one atomicOr; an id whose three bits were all set
//            before is an EVENT: a second or later occurrence (always), or a false positive (~0.1 %).  Events -- 5 % of
//            the ids -- are counted in a small exact table (id << 6 | count, one atomic each).  Two lanes holding the same
//            id in the same instruction are serialised by the atomic on their common word: the later one sees the bits.
//   phase 2  the L largest buckets are streamed the same way but only TESTED: an id whose bits are set is looked up in
//            the exact table and, if present, counted.  No bucket is ever binary-searched for a candidate.
//   phase 3  table entry (id, c): c = events + hits in the largest buckets, and the id's multiplicity is c + 1 unless its
//            very first occurrence was itself a false positive (then c).  c >= T: survivor.  c == T - 1: the one ambiguous
//            case, settled exactly by searching all H buckets (lane per bucket).  c < T - 1: dropped.
// Everything a read touches is read once and coalesced: 4 B per id.  The same code runs warp-per-read (table in a
// slice of shared memory, most reads) and block-per-read (BLOCK = true: the reads of big_list, whose enumerated ids or
// events exceed the warp's tables).  What exceeds the block's tables goes on to big2_list (counting-table kernel).
constexpr int DUP_MLP = 4;         // 128-bit id loads a lane keeps in flight
constexpr int DUP_WFIN = 128;      // survivors / ambiguous ids per read (warp)
constexpr int DUP_BTHREADS = 256;  // threads of the block variant
// events are queued and handled by all lanes together; a queue holds one iteration's worth of ids (every id an event)
constexpr int DUP_WQ = 32 * 4 * DUP_MLP;           // warp: 512 entries
constexpr int DUP_BQ = DUP_BTHREADS * 4 * DUP_MLP; // block: 4096 entries
constexpr uint32_t DUP_PAD = 0xFFFFFF00u;          // padding ids (window ids are < 2^26 here: id << 6 | count)

__host__ __device__ inline size_t dup_slice_words(int words, int xslots, int qcap, int H)
{
    const size_t w = (size_t)words + xslots + (size_t)qcap + 5 * (size_t)H + 16;
    return (w + 3) & ~(size_t)3;
}

// word of the bitmap (top bits of a multiplicative hash) and three bits inside it (its low 15 bits: a permutation of
// the id's low 15 bits, so two ids share a mask only if they are a multiple of 32768 apart)
__device__ __forceinline__ void dup_hash(uint32_t id, int shift, uint32_t& word, uint32_t& mask)
{
    const uint32_t h = id * 0x9E3779B1u;
    word = h >> shift;
    mask = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u)) | (1u << ((h >> 10) & 31u));
}

// Measured and dropped (4 M reads, human-size index; this form: 96 ms per step): two bits per id instead of three
// (99 ms: more false events); events compacted with a warp ballot instead of one atomic on the queue tail (103 ms); one
// queue per thread, applied by the thread itself (136 ms: duplicates come in runs, a few lanes hold most events); one
// copy of the streaming loop per phase (+5 ms: the kernel is ~45 KB of instructions); two loads in flight (103 ms).
template <bool BLOCK>
__global__ void __launch_bounds__(BLOCK ? DUP_BTHREADS : COLLECT_THREADS) collect_dup_kernel(CollectParams P)
{
    extern __shared__ __align__(16) uint32_t ddyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tid = BLOCK ? (int)threadIdx.x : lane;
    const int nthr = BLOCK ? DUP_BTHREADS : 32;
    const int H = P.H, T = P.min_hits;
    const int WMAX = BLOCK ? P.slots : P.warp_slots;      // bitmap words (power of two)
    const int XS = BLOCK ? P.xslots_block : P.xslots_warp; // exact-table slots (power of two)
    const int QCAP = BLOCK ? DUP_BQ : DUP_WQ;
    const int FINCAP = BLOCK ? COLLECT_FINAL_CAP : DUP_WFIN;
    uint32_t* base = ddyn + (BLOCK ? (size_t)0 : (size_t)wid * dup_slice_words(WMAX, XS, QCAP, H));
    uint32_t* bm = base;                          // [WMAX]
    uint32_t* xt = bm + WMAX;                     // [XS] id << 6 | count
    uint32_t* q = xt + XS;                        // [QCAP] queued events of phases 1 / 2; afterwards:
    uint32_t* fin = q;                            //   [FINCAP] survivors
    uint32_t* amb = q + FINCAP;                   //   [FINCAP] ambiguous ids (their exact counts reuse the bitmap)
    // bucket records, one 16-byte load each: .x / .y = END of the bucket in the flat sequence of 16-byte chunks of
    // phase 1 (enumerated buckets) / phase 2 (largest buckets) -- a bucket that is not part of a phase ends where its
    // predecessor ends --, .z = offset of its first id, .w = number of ids
    uint4* rec = reinterpret_cast<uint4*>(q + QCAP); // [H]
    int* sc = reinterpret_cast<int*>(rec + H);       // 0 distinct ids in xt, 1 bad, 2 nfin, 3 namb, 4 E, 5 SK, 6 total, 7 start,
                                                     // 8 queue tail, 9 work item, 10 / 11 chunks of phase 1 / 2
    const uint32_t xmask = (uint32_t)XS - 1u;
    int xshift = 32;
    for (int x = XS; x > 1; x >>= 1) xshift--;
    const int xcap = (XS / 4) * 3;
    const int L = T - 2 < 2 ? (T - 2 < 0 ? 0 : T - 2) : 2; // buckets tested, not inserted
    auto gsync = [&]() {
        if (BLOCK) __syncthreads();
        else __syncwarp();
    };
    unsigned long long st_enum = 0, st_skip = 0, st_big = 0;
    const uint4* __restrict__ vals4 = reinterpret_cast<const uint4*>(P.table_values);
    const int first = BLOCK ? 0 : (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int step = BLOCK ? 1 : (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    const int nitems = BLOCK ? *P.big_count : P.n;
    for (int it = first; it < nitems; it += step) {
        if (BLOCK) { // reads of very different sizes: blocks take the next one from a counter
            __syncthreads();
            if (tid == 0) sc[9] = atomicAdd(P.work, 1);
            __syncthreads();
            it = sc[9];
            if (it >= nitems) break;
        }
        const int rd = BLOCK ? P.big_list[it] : it;
        gsync();
        if (!BLOCK || wid == 0) { // one warp sets the read up: lane t owns buckets t and t + 32
            const int ta = lane, tb = lane + 32;
            const uint2 ra = ta < H ? P.ranges[(int64_t)rd * P.rq + (int64_t)ta * P.rt] : make_uint2(0u, 0u);
            const uint2 rb = tb < H ? P.ranges[(int64_t)rd * P.rq + (int64_t)tb * P.rt] : make_uint2(0u, 0u);
            int sa = 0, sb = 0;
            for (int l = 0; l < L; l++) { // the L largest buckets (ties: lowest table)
                const unsigned ka = (!sa && ra.y > 0u) ? ((ra.y << 8) | (unsigned)(255 - ta)) : 0u;
                const unsigned kb = (!sb && rb.y > 0u) ? ((rb.y << 8) | (unsigned)(255 - tb)) : 0u;
                const unsigned best = __reduce_max_sync(0xffffffffu, ka > kb ? ka : kb);
                if (best == 0u) break;
                const int t = 255 - (int)(best & 255u);
                sa |= t == ta;
                sb |= t == tb;
            }
            const int cha = ra.y > 0u ? (int)(((ra.x & 3u) + ra.y + 3u) >> 2) : 0;
            const int chb = rb.y > 0u ? (int)(((rb.x & 3u) + rb.y + 3u) >> 2) : 0;
            // flat chunk prefixes in table order: tables 0..31 then 32..63
            int v1a = sa ? 0 : cha, v2a = sa ? cha : 0, v1b = sb ? 0 : chb, v2b = sb ? chb : 0;
            int i1a = v1a, i2a = v2a, i1b = v1b, i2b = v2b;
            for (int d = 1; d < 32; d <<= 1) {
                const int o1 = __shfl_up_sync(0xffffffffu, i1a, d), o2 = __shfl_up_sync(0xffffffffu, i2a, d);
                const int o3 = __shfl_up_sync(0xffffffffu, i1b, d), o4 = __shfl_up_sync(0xffffffffu, i2b, d);
                if (lane >= d) {
                    i1a += o1;
                    i2a += o2;
                    i1b += o3;
                    i2b += o4;
                }
            }
            const int t1a = __shfl_sync(0xffffffffu, i1a, 31), t2a = __shfl_sync(0xffffffffu, i2a, 31);
            const int t1b = __shfl_sync(0xffffffffu, i1b, 31), t2b = __shfl_sync(0xffffffffu, i2b, 31);
            if (ta < H) rec[ta] = make_uint4((uint32_t)i1a, (uint32_t)i2a, ra.x, ra.y);
            if (tb < H) rec[tb] = make_uint4((uint32_t)(t1a + i1b), (uint32_t)(t2a + i2b), rb.x, rb.y);
            const unsigned ea = sa ? 0u : ra.y, eb = sb ? 0u : rb.y, ka2 = sa ? ra.y : 0u, kb2 = sb ? rb.y : 0u;
            const unsigned E = __reduce_add_sync(0xffffffffu, ea + eb), SK = __reduce_add_sync(0xffffffffu, ka2 + kb2);
            if (lane == 0) {
                sc[10] = t1a + t1b;
                sc[11] = t2a + t2b;
                sc[0] = 0;
                sc[1] = 0;
                sc[2] = 0;
                sc[3] = 0;
                sc[4] = (int)E;   // <= 64 * 65535
                sc[5] = (int)SK;
                sc[6] = (int)(E + SK);
                sc[8] = 0;
            }
        }
        gsync();
        const int E = sc[4], SK = sc[5], total = sc[6];
        if (total < T) {
            if (tid == 0) {
                P.lists[rd] = make_int2(0, 0);
                st_enum += (unsigned long long)total;
            }
            continue;
        }
        if ((!BLOCK && total > P.warp_cap) || E > 6 * WMAX) { // more ids than the bitmap resolves
            if (tid == 0) {
                if (BLOCK) P.big2_list[atomicAdd(P.big2_count, 1)] = rd;
                else P.big_list[atomicAdd(P.big_count, 1)] = rd;
                st_big++;
            }
            continue;
        }
        int W = 64, shift = 32 - 6;
        while (W < WMAX && 2 * W < E) { // >= half a word per id: ~6 bits of 32 set in a word at the end
            W <<= 1;
            shift--;
        }
        {
            uint4* b4 = reinterpret_cast<uint4*>(bm);
            for (int i = tid; i < W / 4; i += nthr) b4[i] = make_uint4(0u, 0u, 0u, 0u);
            uint4* x4 = reinterpret_cast<uint4*>(xt);
            for (int i = tid; i < XS / 4; i += nthr) x4[i] = make_uint4(COLLECT_EMPTY, COLLECT_EMPTY, COLLECT_EMPTY, COLLECT_EMPTY);
        }
        gsync();
        // queued events -> exact table, all lanes busy.  phase 0: insert / count; phase 1: count where present
        auto apply = [&](uint32_t id, int phase) {
            {
                if (id >= DUP_PAD) return; // padding of a bucket's first / last 16-byte chunk
                uint32_t hh = collect_hash(id) >> xshift;
                if (phase == 0) {
                    if (sc[0] >= xcap) { // the exact table is full: a bigger one takes the read
                        sc[1] = 1;
                        return;
                    }
                    while (true) {
                        uint32_t cur = xt[hh];
                        if (cur == COLLECT_EMPTY) {
                            cur = atomicCAS(&xt[hh], COLLECT_EMPTY, (id << COLLECT_PACK_BITS) | 1u);
                            if (cur == COLLECT_EMPTY) {
                                atomicAdd(&sc[0], 1);
                                break;
                            }
                        }
                        if ((cur >> COLLECT_PACK_BITS) == id) {
                            atomicAdd(&xt[hh], 1u);
                            break;
                        }
                        hh = (hh + 1) & xmask;
                    }
                } else {
                    while (true) {
                        const uint32_t cur = xt[hh];
                        if (cur == COLLECT_EMPTY) break;
                        if ((cur >> COLLECT_PACK_BITS) == id) {
                            atomicAdd(&xt[hh], 1u);
                            break;
                        }
                        hh = (hh + 1) & xmask;
                    }
                }
            }
        };
        auto drain = [&](int phase) {
            gsync();
            const int nq = sc[8];
            for (int i = tid; i < nq; i += nthr) apply(q[i], phase);
            gsync();
            if (tid == 0) sc[8] = 0;
            gsync();
        };
        // phases 1 and 2: the same streaming loop over a flat sequence of 16-byte chunks (one copy of the code: the
        // compiler switches on the phase around the ids of an iteration)
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
            const int ctotal = sc[10 + phase];
            int tnext = 0, cfirst = 0, cnext = 0; // the lane's current bucket holds the chunks [cfirst, cnext)
            uint32_t lo = 0u, hi = 0u;            // its element range
            uint32_t cbase = 0u;                  // chunk c of it lies at vals4[cbase + c] (offsets are 32 bit)
            for (int c0 = 0; c0 < ctotal; c0 += nthr * DUP_MLP) {
                uint4 v[DUP_MLP];
                uint32_t inv = 0u; // nibble u: the elements of v[u] that are not ids of this read
#pragma unroll
                for (int u = 0; u < DUP_MLP; u++) {
                    const int c = c0 + u * nthr + tid;
                    if (c < ctotal) {
                        if (c >= cnext) {
                            uint4 r;
                            int end = cnext;
                            do {
                                cfirst = end;
                                r = rec[tnext++];
                                end = (int)(phase == 0 ? r.x : r.y);
                            } while (end <= c);
                            cnext = end;
                            lo = r.z;
                            hi = r.z + r.w;
                            cbase = (lo >> 2) - (uint32_t)cfirst;
                        }
                        v[u] = __ldg(vals4 + (cbase + (uint32_t)c));
                        if (c == cfirst || c + 1 == cnext) { // a bucket begins / ends inside its first / last chunk
                            uint32_t nb = c == cfirst ? (1u << (lo & 3u)) - 1u : 0u;
                            if (c + 1 == cnext && (hi & 3u) != 0u) nb |= (0xFu << (hi & 3u)) & 0xFu;
                            inv |= nb << (4 * u);
                        }
                    } else {
                        v[u] = make_uint4(0u, 0u, 0u, 0u);
                        inv |= 0xFu << (4 * u);
                    }
                }
#pragma unroll
                for (int u = 0; u < DUP_MLP; u++) {
                    uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    const uint32_t nb = (inv >> (4 * u)) & 0xFu;
                    if (nb != 0u) {
                        if (nb == 0xFu) continue; // past the end of the sequence
                        // foreign elements become padding ids (distinct per lane and element: distinct bitmap words);
                        // they go through the filter like ids and are dropped when the queue is drained
                        const uint32_t pad = DUP_PAD | ((uint32_t)lane << 2);
                        if (nb & 1u) x[0] = pad;
                        if (nb & 2u) x[1] = pad | 1u;
                        if (nb & 4u) x[2] = pad | 2u;
                        if (nb & 8u) x[3] = pad | 3u;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t word, mask;
                        dup_hash(x[j], shift, word, mask);
                        bool ev;
                        if (phase == 0) ev = (atomicOr(&bm[word], mask) & mask) == mask;
                        else ev = (bm[word] & mask) == mask;
                        if (ev) q[atomicAdd(&sc[8], 1)] = x[j];
                    }
                }
                // the queue holds one iteration's worth of ids: empty it before the next iteration adds to it
                if (c0 + nthr * DUP_MLP < ctotal) {
                    gsync();
                    if (sc[8] > 0) drain(phase);
                }
            }
            drain(phase);
            if (sc[1]) break;
        }
        if (sc[1]) {
            if (tid == 0) {
                if (BLOCK) P.big2_list[atomicAdd(P.big2_count, 1)] = rd;
                else P.big_list[atomicAdd(P.big_count, 1)] = rd;
                st_big++;
            }
            continue;
        }
        if (tid == 0) {
            st_enum += (unsigned long long)E;
            st_skip += (unsigned long long)SK;
        }
        // phase 3: c >= T survives, c == T - 1 is settled exactly
        if (sc[0] > 0) {
            for (int i = tid; i < XS; i += nthr) {
                const uint32_t x = xt[i];
                if (x == COLLECT_EMPTY) continue;
                const uint32_t c = x & ((1u << COLLECT_PACK_BITS) - 1u), id = x >> COLLECT_PACK_BITS;
                if (c >= (uint32_t)T) {
                    const int at = atomicAdd(&sc[2], 1);
                    if (at < FINCAP) fin[at] = id;
                } else if (c + 1u == (uint32_t)T) {
                    const int at = atomicAdd(&sc[3], 1);
                    if (at < FINCAP) amb[at] = id;
                }
            }
            gsync();
            const int namb = sc[3];
            if (namb > FINCAP || sc[2] > FINCAP) {
                if (tid == 0) sc[1] = 1;
            } else if (namb > 0) {
                uint32_t* acnt = bm; // the bitmap is free now
                for (int c = tid; c < namb; c += nthr) acnt[c] = 0u;
                gsync();
                for (int x = tid; x < namb * H; x += nthr) { // exact multiplicity over all H buckets
                    const int c = x / H, t = x - c * H;
                    const int cnt = (int)rec[t].w;
                    if (cnt > 0) {
                        const uint32_t id = amb[c];
                        const uint32_t* p = P.table_values + rec[t].z;
                        const int pos = collect_lower_bound_interp(p, cnt, id, P.id_space);
                        if (pos < cnt && __ldg(p + pos) == id) atomicAdd(&acnt[c], 1u);
                    }
                }
                gsync();
                for (int c = tid; c < namb; c += nthr)
                    if (acnt[c] >= (uint32_t)T) {
                        const int at = atomicAdd(&sc[2], 1);
                        if (at < FINCAP) fin[at] = amb[c];
                    }
            }
            gsync();
            if (sc[1] || sc[2] > FINCAP) {
                if (tid == 0) {
                    if (BLOCK) P.big2_list[atomicAdd(P.big2_count, 1)] = rd;
                    else P.big_list[atomicAdd(P.big_count, 1)] = rd;
                    st_big++;
                    st_enum -= (unsigned long long)E;
                    st_skip -= (unsigned long long)SK;
                }
                continue;
            }
        }
        const int nfin = sc[2];
        if (nfin > 1) {
            if (BLOCK) bitonic_sort<true>(fin, nfin, tid, nthr);
            else collect_sort_warp(fin, nfin, lane);
        }
        gsync();
        if (tid == 0) {
            unsigned long long start = 0;
            int nf = nfin;
            if (nf > 0) start = atomicAdd(P.cursor, (unsigned long long)nf);
            if (start + (unsigned long long)nf > P.out_cap) {
                *P.overflow = 1;
                nf = 0;
            }
            P.lists[rd] = make_int2((int)start, nf);
            sc[2] = nf;
            sc[7] = (int)(start & 0x7fffffffu); // out_cap < 2^31
        }
        gsync();
        {
            const int nf = sc[2];
            const unsigned long long start = (unsigned long long)(unsigned)sc[7];
            for (int i = tid; i < nf; i += nthr) P.out[start + i] = fin[i];
        }
    }
    if (tid == 0 && P.stats) {
        if (st_enum) atomicAdd(P.stats, st_enum);
        if (st_skip) atomicAdd(P.stats + 1, st_skip);
        if (st_big && !BLOCK) atomicAdd(P.stats + 3, st_big);
        if (st_big && BLOCK) atomicAdd(P.stats + 4, st_big);
    }
}

// ---- block per read ----------------------------------------------------------------------------------
template <bool PACKED>
__global__ void __launch_bounds__(COLLECT_BIG_THREADS) collect_big_kernel(CollectParams P)
{
    extern __shared__ uint32_t cdyn[];
    const int S = P.slots;
    const CountTable<PACKED> tab{cdyn, S};
    uint32_t* cand = cdyn + (PACKED ? 1 : 2) * S; // [S / 2] ids that reached the reduced threshold in the enumerated buckets
    uint32_t* ccnt = cand + S / 2;                 // [S / 2] their counts
    uint32_t* fin = ccnt + S / 2;                  // [COLLECT_FINAL_CAP] survivors of the read, ascending
    __shared__ uint32_t offv[MAX_TABLES];
    __shared__ int cntv[MAX_TABLES];
    __shared__ int skip[MAX_TABLES];
    __shared__ int bnd[(COLLECT_CHUNK + 1) * MAX_TABLES]; // bucket positions of the range boundaries of a chunk
    __shared__ int gtot[COLLECT_CHUNK];                   // ids per range of the chunk
    __shared__ int gpre[COLLECT_CHUNK * (MAX_TABLES + 1)]; // per range: exclusive prefix of its bucket pieces
    __shared__ int s_ncand, s_gfin, s_nfin, s_bad;
    __shared__ unsigned long long s_start;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, H = P.H, T = P.min_hits;
    const int L = T - 2 < 2 ? (T - 2 < 0 ? 0 : T - 2) : 2; // buckets not enumerated
    const int thr = T - L;                                  // >= 2 for T >= 2
    const int nbig = *P.big_count;
    unsigned long long st_enum = 0, st_skip = 0, st_ranges = 0;
    tab.clear_all(tid, COLLECT_BIG_THREADS); // every range leaves the slots it used empty again
    if (tid == 0) {
        s_ncand = 0;
        s_gfin = 0;
    }
    __syncthreads();
    for (int bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const int rd = P.big_list[bi];
        if (tid < H) {
            const uint2 r = P.ranges[(int64_t)rd * P.rq + (int64_t)tid * P.rt];
            offv[tid] = r.x;
            cntv[tid] = (int)r.y;
            skip[tid] = 0;
        }
        if (tid == 0) {
            s_nfin = 0;
            s_bad = 0;
        }
        __syncthreads();
        if (tid == 0) {
            for (int l = 0; l < L; l++) { // the L largest buckets
                int best = -1;
                for (int t = 0; t < H; t++)
                    if (!skip[t] && (best < 0 || cntv[t] > cntv[best])) best = t;
                if (best >= 0 && cntv[best] > 0) skip[best] = 1;
            }
        }
        __syncthreads();
        int64_t E = 0, SK = 0;
        for (int t = 0; t < H; t++) {
            if (skip[t]) SK += cntv[t];
            else E += cntv[t];
        }
        const int64_t nranges = E > 0 ? HRM_SDIV(E, (int64_t)P.fill) : 1;
        const uint64_t width = ((uint64_t)P.id_space + (uint64_t)nranges - 1) / (uint64_t)nranges;
        if (tid == 0) {
            st_enum += (unsigned long long)E;
            st_skip += (unsigned long long)SK;
            st_ranges += (unsigned long long)nranges;
        }
        bool bad = false; // block-uniform copy of s_bad (read only after a barrier)
        for (int64_t g0 = 0; g0 < nranges && !bad; g0 += COLLECT_CHUNK) {
            const int ng = (int)((nranges - g0) < COLLECT_CHUNK ? (nranges - g0) : COLLECT_CHUNK);
            // boundaries of ranges g0 .. g0 + ng in every enumerated bucket, all searches in flight together
            for (int x = tid; x < (ng + 1) * H; x += COLLECT_BIG_THREADS) {
                const int g = x / H, t = x - g * H;
                int pos = 0;
                if (!skip[t]) {
                    const uint64_t key = (uint64_t)(g0 + g) * width;
                    pos = key >= (uint64_t)P.id_space
                              ? cntv[t]
                              : collect_lower_bound_interp(P.table_values + offv[t], cntv[t], (uint32_t)key, P.id_space);
                }
                bnd[g * MAX_TABLES + t] = pos;
            }
            __syncthreads();
            if (tid < ng) {
                int tot = 0;
                for (int t = 0; t < H; t++) {
                    gpre[tid * (MAX_TABLES + 1) + t] = tot; // flat index of the first id of bucket t inside range tid
                    tot += bnd[(tid + 1) * MAX_TABLES + t] - bnd[tid * MAX_TABLES + t];
                }
                gpre[tid * (MAX_TABLES + 1) + H] = tot;
                gtot[tid] = tot;
            }
            __syncthreads();
            for (int g = 0; g < ng && !bad; g++) {
                const int gtotal = gtot[g];
                if (gtotal < thr) continue;
                if (8 * gtotal > 7 * S) { // ids far from uniform: this range would clog the table
                    bad = true;
                    break;
                }
                int slots = 64, shift = 32 - 6;
                while (slots < S && 5 * slots < 8 * gtotal) {
                    slots <<= 1;
                    shift--;
                }
                // the range's ids as one flat sequence (balanced over the block whatever the bucket sizes);
                // COLLECT_MLP independent loads per thread are issued before the first atomic, so a thread waits for
                // DRAM once per batch instead of once per id
                const int* gp = gpre + g * (MAX_TABLES + 1);
                int tcur = 0;
                for (int e0 = 0; e0 < gtotal; e0 += COLLECT_BIG_THREADS * COLLECT_MLP) {
                    uint32_t v[COLLECT_MLP];
#pragma unroll
                    for (int u = 0; u < COLLECT_MLP; u++) {
                        const int e = e0 + u * COLLECT_BIG_THREADS + tid;
                        v[u] = COLLECT_EMPTY;
                        if (e < gtotal) {
                            while (gp[tcur + 1] <= e) tcur++; // bucket with gp[t] <= e < gp[t + 1]; e only grows
                            v[u] = __ldg(P.table_values + offv[tcur] + bnd[g * MAX_TABLES + tcur] + (e - gp[tcur]));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < COLLECT_MLP; u++)
                        if (v[u] != COLLECT_EMPTY) tab.count((uint32_t)slots - 1u, shift, v[u]);
                }
                __syncthreads();
                for (int i = tid; i < slots; i += COLLECT_BIG_THREADS) { // collect + leave the table empty
                    uint32_t id = 0, c = 0;
                    if (tab.take(i, id, c) && c >= (uint32_t)thr) {
                        const int at = atomicAdd(&s_ncand, 1); // <= gtotal / thr <= S / 2
                        cand[at] = id;
                        ccnt[at] = c;
                    }
                }
                __syncthreads();
                const int ncand = s_ncand;
                if (ncand == 0) continue; // the common case: nothing in this id range reaches the threshold
                const int base = s_nfin;
                for (int c = tid; c < ncand; c += COLLECT_BIG_THREADS) {
                    const uint32_t id = cand[c];
                    uint32_t m = ccnt[c];
                    for (int t = 0; t < H && m < (uint32_t)T; t++) {
                        if (!skip[t]) continue;
                        const uint32_t* p = P.table_values + offv[t];
                        const int pos = collect_lower_bound_interp(p, cntv[t], id, P.id_space);
                        if (pos < cntv[t] && __ldg(p + pos) == id) m++;
                    }
                    if (m >= (uint32_t)T) {
                        const int at = base + atomicAdd(&s_gfin, 1);
                        if (at < COLLECT_FINAL_CAP) fin[at] = id;
                        else s_bad = 1;
                    }
                }
                __syncthreads();
                bad = s_bad != 0;
                const int gf = s_gfin;
                // sort this range's survivors; ranges ascend, so the read's list stays sorted
                if (!bad && gf > 1) bitonic_sort<true>(fin + base, gf, tid, COLLECT_BIG_THREADS);
                __syncthreads();
                if (tid == 0) {
                    s_nfin = base + (bad ? 0 : gf);
                    s_ncand = 0;
                    s_gfin = 0;
                }
                __syncthreads();
            }
        }
        if (bad) {
            __syncthreads();
            tab.clear_all(tid, COLLECT_BIG_THREADS); // the aborted range may have left ids behind
            if (tid == 0) {
                *P.overflow = 1;
                P.lists[rd] = make_int2(0, 0);
                s_ncand = 0;
                s_gfin = 0;
            }
            __syncthreads();
            continue;
        }
        __syncthreads();
        const int ns = s_nfin;
        if (tid == 0) {
            unsigned long long start = 0;
            if (ns > 0) start = atomicAdd(P.cursor, (unsigned long long)ns);
            if (start + (unsigned long long)ns > P.out_cap) {
                *P.overflow = 1;
                start = ~0ull;
            }
            s_start = start;
        }
        __syncthreads();
        const unsigned long long start = s_start;
        if (start != ~0ull) {
            for (int i = tid; i < ns; i += COLLECT_BIG_THREADS) P.out[start + i] = fin[i];
            if (tid == 0) P.lists[rd] = make_int2((int)start, ns);
        } else if (tid == 0) {
            P.lists[rd] = make_int2(0, 0);
        }
        __syncthreads();
    }
    if (tid == 0 && P.stats) {
        if (st_enum) atomicAdd(P.stats, st_enum);
        if (st_skip) atomicAdd(P.stats + 1, st_skip);
        if (st_ranges) atomicAdd(P.stats + 2, st_ranges);
    }
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    if (!v) return dflt;
    const int x = atoi(v);
    return x > 0 ? x : dflt;
}

// Candidate lists of n reads from the bucket ranges of the last probe.  d_lists: n (start, count) pairs into
// d_out (capacity out_cap ids).  *h_overflow != 0: rerun on the general path.  Synchronises once.
hrm_status collect_candidates(const hrm_minhasher* mh, const QueryHandle* qh, int n, int min_hits, uint32_t id_space,
                              uint32_t* d_out, int64_t out_cap, int2* d_lists, int64_t* h_total, int* h_overflow,
                              int64_t* h_stats3, cudaStream_t s)
{
    return collect_candidates_from(qh->ranges.as<uint2>(), qh->rq, qh->rt, mh->values, mh->H, n, min_hits, id_space, d_out,
                                   out_cap, d_lists, h_total, h_overflow, h_stats3, s);
}

// The same over any array of ascending id lists: (offset, count) of list (q, t) at ranges[q * rq + t * rt], offsets into
// table_values (the key-partitioned index hands the routed value lists over this way).
hrm_status collect_candidates_from(const uint2* d_ranges, int64_t rq, int64_t rt, const uint32_t* d_table_values, int H, int n,
                                   int min_hits, uint32_t id_space, uint32_t* d_out, int64_t out_cap, int2* d_lists,
                                   int64_t* h_total, int* h_overflow, int64_t* h_stats3, cudaStream_t s)
{
    *h_overflow = 0;
    *h_total = 0;
    if (n == 0) return HRM_OK;
    HRM_REQUIRE(min_hits >= 2 && id_space < 0xFFFFFFFFu, "collect_candidates needs minTableHits >= 2");
    const int warp_cap = env_int("HRM_COLLECT_WARP_CAP", 0x7fffffff);
    const int wslots_env = env_int("HRM_COLLECT_WARP_SLOTS", 1024);
    int wslots = 128;
    while (wslots < wslots_env && wslots < 8192) wslots <<= 1;
    const int slots_env = env_int("HRM_COLLECT_SLOTS", 2048);
    int slots = 64;
    while (slots < slots_env && slots < 16384) slots <<= 1;
    const int fill_env = env_int("HRM_COLLECT_FILL", 0);
    const int fill = fill_env > 0 ? fill_env : (slots * 3) / 8; // expected ids per range: 3/8 of the table, 7/8 tolerated
    Scratch ctl, big, big2;
    HRM_TRY(ctl.alloc(sizeof(unsigned long long) * 32, s));
    HRM_TRY(big.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    HRM_TRY(big2.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    HRM_CUDA(cudaMemsetAsync(ctl.p, 0, sizeof(unsigned long long) * 32, s));
    HRM_CUDA(cudaMemsetAsync(big.p, 0, sizeof(int32_t), s));
    HRM_CUDA(cudaMemsetAsync(big2.p, 0, sizeof(int32_t), s));
    CollectParams P;
    P.ranges = d_ranges;
    P.rq = rq;
    P.rt = rt;
    P.table_values = d_table_values;
    P.n = n;
    P.H = H;
    P.min_hits = min_hits;
    P.id_space = id_space;
    P.out = d_out;
    P.out_cap = (unsigned long long)out_cap;
    P.cursor = ctl.as<unsigned long long>();
    P.overflow = reinterpret_cast<int*>(ctl.as<unsigned long long>() + 1);
    P.stats = ctl.as<unsigned long long>() + 2;
    P.lists = d_lists;
    P.big_count = big.as<int32_t>();
    P.big_list = big.as<int32_t>() + 1;
    P.big2_count = big2.as<int32_t>();
    P.big2_list = big2.as<int32_t>() + 1;
    P.warp_cap = warp_cap;
    P.warp_slots = wslots;
    P.slots = slots;
    P.fill = fill;
    const int impl_ranges = env_int("HRM_COLLECT_RANGES", 0); // 1: the counting-table warp kernel (id ranges)
    const int bloom_words_env = env_int("HRM_COLLECT_BLOOM_WORDS", 1024);
    int bwords = 64;
    while (bwords < bloom_words_env && bwords < 8192) bwords <<= 1;
    const int bblock_env = env_int("HRM_COLLECT_BLOOM_BLOCK_WORDS", 8192);
    int bblock_words = 64;
    while (bblock_words < bblock_env && bblock_words < 32768) bblock_words <<= 1;
    const int xwarp_env = env_int("HRM_COLLECT_XSLOTS", 512), xblock_env = env_int("HRM_COLLECT_BLOCK_XSLOTS", 4096);
    int xwarp = 256, xblock = 2048; // the table keeps a quarter of its slots free: more than one slot per thread
    while (xwarp < xwarp_env && xwarp < 4096) xwarp <<= 1;
    while (xblock < xblock_env && xblock < 8192) xblock <<= 1;
    P.xslots_warp = xwarp;
    P.xslots_block = xblock;
    P.work = reinterpret_cast<int*>(ctl.as<unsigned long long>() + 24);
    const bool allow_packed = env_int("HRM_COLLECT_UNPACKED", 0) == 0;
    const bool packed = allow_packed && id_space < (1u << (32 - COLLECT_PACK_BITS)) - 1u && H < (1 << COLLECT_PACK_BITS);
    const size_t smem = sizeof(uint32_t) * ((size_t)((packed ? 1 : 2) + 1) * slots + COLLECT_FINAL_CAP);
    // warp kernel: 2 warps per block so that the slices of many blocks fill the SM's shared memory
    const int wthreads = 64;
    const size_t wsmem = sizeof(uint32_t) * collect_warp_slice_words(wslots, H, packed) * (wthreads / 32);
    int wres = 1;
    // one wave of resident blocks, each loops over the list of big reads
    int resident = 1;
    if (!impl_ranges && packed) {
        P.warp_slots = bwords;
        P.slots = bblock_words;
        auto wkern = collect_dup_kernel<false>;
        auto bkern = collect_dup_kernel<true>;
        const size_t bsmem = sizeof(uint32_t) * dup_slice_words(bwords, xwarp, DUP_WQ, H) * (wthreads / 32);
        cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wres, wkern, wthreads, bsmem));
        const int wcap = env_int("HRM_COLLECT_BLOCKS_PER_SM", 0);
        if (wcap > 0 && wres > wcap) wres = wcap;
        HRM_LAUNCH(wkern, (unsigned)(num_sms() * (wres > 0 ? wres : 1)), wthreads, bsmem, s, P);
        // skewed reads: the same scheme block-wide, then (what is left) the counting-table kernel on big2_list
        const size_t bbsmem = sizeof(uint32_t) * dup_slice_words(bblock_words, xblock, DUP_BQ, H);
        int bres = 1;
        cudaFuncSetAttribute(bkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bbsmem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bres, bkern, DUP_BTHREADS, bbsmem));
        HRM_LAUNCH(bkern, (unsigned)(num_sms() * (bres > 0 ? bres : 1)), DUP_BTHREADS, bbsmem, s, P);
        P.slots = slots;
        CollectParams PC = P;
        PC.big_count = P.big2_count;
        PC.big_list = P.big2_list;
        if (packed) {
            cudaFuncSetAttribute(collect_big_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, collect_big_kernel<true>, COLLECT_BIG_THREADS, smem));
            HRM_LAUNCH(collect_big_kernel<true>, (unsigned)(num_sms() * (resident > 0 ? resident : 1)), COLLECT_BIG_THREADS,
                       smem, s, PC);
        } else {
            cudaFuncSetAttribute(collect_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, collect_big_kernel<false>, COLLECT_BIG_THREADS, smem));
            HRM_LAUNCH(collect_big_kernel<false>, (unsigned)(num_sms() * (resident > 0 ? resident : 1)), COLLECT_BIG_THREADS,
                       smem, s, PC);
        }
    } else if (packed) {
        cudaFuncSetAttribute(collect_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wres, collect_warp_kernel<true>, wthreads, wsmem));
        HRM_LAUNCH(collect_warp_kernel<true>, (unsigned)(num_sms() * (wres > 0 ? wres : 1)), wthreads, wsmem, s, P);
        cudaFuncSetAttribute(collect_big_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, collect_big_kernel<true>, COLLECT_BIG_THREADS, smem));
        HRM_LAUNCH(collect_big_kernel<true>, (unsigned)(num_sms() * (resident > 0 ? resident : 1)), COLLECT_BIG_THREADS, smem,
                   s, P);
    } else {
        cudaFuncSetAttribute(collect_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wres, collect_warp_kernel<false>, wthreads, wsmem));
        HRM_LAUNCH(collect_warp_kernel<false>, (unsigned)(num_sms() * (wres > 0 ? wres : 1)), wthreads, wsmem, s, P);
        cudaFuncSetAttribute(collect_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        HRM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, collect_big_kernel<false>, COLLECT_BIG_THREADS, smem));
        HRM_LAUNCH(collect_big_kernel<false>, (unsigned)(num_sms() * (resident > 0 ? resident : 1)), COLLECT_BIG_THREADS,
                   smem, s, P);
    }
    unsigned long long h[32];
    HRM_CUDA(cudaMemcpyAsync(h, ctl.p, sizeof h, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s)); // the one sync of the pass: overflow flag + candidate total
    *h_total = (int64_t)h[0];
    *h_overflow = (int)(h[1] & 0xFFFFFFFFull);
    if (getenv("HRM_COLLECT_DEBUG"))
        fprintf(stderr, "collect: n=%d inserted=%llu tested=%llu warp->block=%llu block->table=%llu\n", n, h[2], h[3], h[5], h[6]);
    if (h_stats3) {
        h_stats3[0] = (int64_t)h[2];
        h_stats3[1] = (int64_t)h[3];
        h_stats3[2] = (int64_t)h[5]; // reads handed to the block kernel
    }
    return HRM_OK;
}

} // namespace hrm
