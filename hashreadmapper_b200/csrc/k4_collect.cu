// k4_collect.cu -- K4: candidate collection = per-segment sort + run-length reduction + threshold.
// ref: GpuSegmentedUniqueByCount::unique include/gpu/cuda_unique_by_count.cuh:33-215 (cub segmented
//      radix sort + ~9 thrust launches) via GpuMinhashQueryFilter::keepDistinctByFrequency
//      include/gpu/minhashqueryfilter.cuh:239-278; keepDistinct :217-236 (minTableHits <= 1).
// Here (minHits >= 2, the mapper's case): multiplicities are COUNTED, not sorted out -- every id of a
// segment goes into an open-addressing table in shared memory (atomicCAS on the id, atomicAdd on its
// count; one warp per segment up to 256 ids, one block per segment above, ids hashed into several
// passes when the segment exceeds the table), the few ids that reach minHits are collected and only
// those are sorted: O(ids) work instead of O(ids log^2 ids).  At human-genome scale a read retrieves
// ~4000 ids per pass of which 1-3 survive.  minHits <= 1 (keepDistinct), and any segment whose
// survivors overflow the shared list, take the sort path: normalised bitonic network (all
// compare-exchanges ascending, so virtual +inf padding needs no storage), run heads, threshold by
// looking minHits-1 places ahead.  Then one scan and one gather compact the lists.
// Traffic: 4 B read + <= 4 B written per candidate id, plus 8 B per segment.
#include "runtime.cuh"
#include "k4_sort.cuh"

namespace hrm {

constexpr int K4_WARP_CAP = 256;    // ids per segment handled by one warp
constexpr int K4_BLOCK_CAP = 8192;  // ids per segment sorted in shared memory by one block
constexpr int K4_THREADS = 256;
constexpr int K4_WARP_SLOTS = 512;   // counting-table slots of a warp (load <= 0.5)
constexpr int K4_BLOCK_SLOTS = 8192; // counting-table slots of a block
constexpr int K4_BLOCK_FILL = 5120;  // ids hashed into one pass of the block table (load <= 0.63)
constexpr int K4_SURV_CAP = 1024;    // survivors a block can hold before it falls back to the sort path
constexpr int K4_MAX_GROUPS = 512;   // hash groups of one segment (segments up to 2.6 M ids are counted)
constexpr uint32_t K4_EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t k4_hash(uint32_t v) { return v * 0x9E3779B1u; }

// counts one id; slots = power of two.  The id 0xFFFFFFFF (the empty marker) is counted by the caller.
__device__ __forceinline__ void k4_count(uint32_t* keys, uint32_t* cnts, uint32_t mask, int shift, uint32_t v)
{
    uint32_t h = k4_hash(v) >> shift;
    while (true) {
        const uint32_t prev = atomicCAS(&keys[h], K4_EMPTY, v);
        if (prev == K4_EMPTY || prev == v) {
            atomicAdd(&cnts[h], 1u);
            return;
        }
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ bool keep_head(const uint32_t* s, int i, int cnt, int min_hits)
{
    if (i >= cnt) return false;
    const uint32_t v = s[i];
    if (i > 0 && s[i - 1] == v) return false;
    if (min_hits <= 1) return true;
    const int j = i + min_hits - 1;
    return j < cnt && s[j] == v;
}

__global__ void __launch_bounds__(K4_THREADS) filter_small_kernel(uint32_t* __restrict__ values,
                                                                  const int32_t* __restrict__ offsets, int n,
                                                                  int min_hits, int32_t* __restrict__ new_counts,
                                                                  int32_t* __restrict__ big_list,
                                                                  int32_t* __restrict__ big_count)
{
    __shared__ uint32_t sm[K4_THREADS / 32][K4_WARP_CAP];
    __shared__ uint32_t hk[K4_THREADS / 32][K4_WARP_SLOTS];
    __shared__ uint32_t hc[K4_THREADS / 32][K4_WARP_SLOTS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* s = sm[wid];
    uint32_t* keys = hk[wid];
    uint32_t* cnts = hc[wid];
    const int warp0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    for (int seg = warp0; seg < n; seg += nwarps) {
        const int b = offsets[seg];
        const int cnt = offsets[seg + 1] - b;
        if (cnt <= 0) {
            if (lane == 0) new_counts[seg] = 0;
            continue;
        }
        if (cnt > K4_WARP_CAP) {
            if (lane == 0) {
                big_list[atomicAdd(big_count, 1)] = seg;
                new_counts[seg] = 0; // filled in by the block kernel
            }
            continue;
        }
        if (min_hits >= 2) {
            // count multiplicities in a table of >= 2 * cnt slots, collect the ids that reach min_hits, sort those
            int slots = 32, shift = 27;
            while (slots < 2 * cnt) {
                slots <<= 1;
                shift--;
            }
            for (int i = lane; i < slots; i += 32) {
                keys[i] = K4_EMPTY;
                cnts[i] = 0u;
            }
            __syncwarp();
            int nff = 0;
            for (int i = lane; i < cnt; i += 32) {
                const uint32_t v = values[b + i];
                if (v == K4_EMPTY) nff++;
                else k4_count(keys, cnts, (uint32_t)slots - 1u, shift, v);
            }
            nff = __reduce_add_sync(0xffffffffu, nff);
            __syncwarp();
            int ns = 0;
            for (int base = 0; base < slots; base += 32) {
                const int i = base + lane;
                const bool keep = cnts[i] >= (uint32_t)min_hits;
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) s[ns + __popc(m & ((1u << lane) - 1u))] = keys[i]; // ns <= cnt / min_hits < K4_WARP_CAP
                ns += __popc(m);
            }
            if (nff >= min_hits) { // the largest id sorts last
                if (lane == 0) s[ns] = K4_EMPTY;
                ns++;
            }
            __syncwarp();
            if (ns > 1) bitonic_sort<false>(s, ns, lane, 32);
            for (int i = lane; i < ns; i += 32) values[b + i] = s[i];
            if (lane == 0) new_counts[seg] = ns;
            __syncwarp();
            continue;
        }
        for (int i = lane; i < cnt; i += 32) s[i] = values[b + i];
        __syncwarp();
        bitonic_sort<false>(s, cnt, lane, 32);
        int outpos = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int i = base + lane;
            const bool keep = keep_head(s, i, cnt, min_hits);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) values[b + outpos + __popc(m & ((1u << lane) - 1u))] = s[i];
            outpos += __popc(m);
        }
        if (lane == 0) new_counts[seg] = outpos;
        __syncwarp();
    }
}

// block per large segment, counting version (min_hits >= 2).  A segment that exceeds one table fill is first
// partitioned by hash group into `scratch` (histogram, scan, scatter -- O(ids)), then every group is counted on
// its own; survivors are collected in shared memory, sorted and written to the front of the segment once all
// groups are done.  A segment with more survivors than the list holds, more groups than the histogram, or a
// wildly skewed group goes to the sort kernel through sort_list.
__global__ void __launch_bounds__(K4_THREADS) filter_count_kernel(uint32_t* __restrict__ values,
                                                                  uint32_t* __restrict__ scratch,
                                                                  const int32_t* __restrict__ offsets, int min_hits,
                                                                  int32_t* __restrict__ new_counts,
                                                                  const int32_t* __restrict__ big_list,
                                                                  const int32_t* __restrict__ big_count,
                                                                  int32_t* __restrict__ sort_list,
                                                                  int32_t* __restrict__ sort_count)
{
    extern __shared__ uint32_t k4_dyn[];
    uint32_t* keys = k4_dyn;                              // [K4_BLOCK_SLOTS]
    uint32_t* cnts = keys + K4_BLOCK_SLOTS;               // [K4_BLOCK_SLOTS]
    uint32_t* surv = cnts + K4_BLOCK_SLOTS;               // [K4_SURV_CAP]
    int* ghist = reinterpret_cast<int*>(surv + K4_SURV_CAP); // [K4_MAX_GROUPS] ids per group
    int* gpos = ghist + K4_MAX_GROUPS;                    // [K4_MAX_GROUPS] start, then running position
    __shared__ int s_ns, s_nff, s_bad;
    const int tid = threadIdx.x;
    const int nbig = *big_count;
    for (int bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const int seg = big_list[bi];
        const int b = offsets[seg];
        const int cnt = offsets[seg + 1] - b;
        uint32_t* g = values + b;
        const int groups = HRM_SDIV(cnt, K4_BLOCK_FILL);
        if (tid == 0) {
            s_ns = 0;
            s_nff = 0;
            s_bad = groups > K4_MAX_GROUPS ? 1 : 0;
        }
        const uint32_t* src = g;
        if (groups > 1 && groups <= K4_MAX_GROUPS) { // partition by hash group into scratch
            for (int i = tid; i < groups; i += K4_THREADS) ghist[i] = 0;
            __syncthreads();
            int nff = 0;
            for (int i = tid; i < cnt; i += K4_THREADS) {
                const uint32_t v = g[i];
                if (v == K4_EMPTY) nff++;
                else atomicAdd(&ghist[(int)((k4_hash(v) & 0xFFFFu) * (uint32_t)groups >> 16)], 1);
            }
            if (nff) atomicAdd(&s_nff, nff);
            __syncthreads();
            if (tid < 32) { // exclusive scan of the histogram by one warp
                int carry = 0;
                for (int base = 0; base < groups; base += 32) {
                    const int i = base + tid;
                    const int x = i < groups ? ghist[i] : 0;
                    int incl = x;
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (tid >= d) incl += o;
                    }
                    if (i < groups) {
                        gpos[i] = carry + incl - x;
                        if (x > K4_BLOCK_SLOTS - K4_BLOCK_SLOTS / 8) s_bad = 1; // skewed group: would clog the table
                    }
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
            uint32_t* dst = scratch + b;
            for (int i = tid; i < cnt; i += K4_THREADS) {
                const uint32_t v = g[i];
                if (v != K4_EMPTY) dst[atomicAdd(&gpos[(int)((k4_hash(v) & 0xFFFFu) * (uint32_t)groups >> 16)], 1)] = v;
            }
            __syncthreads(); // gpos[p] is now the END of group p; the scattered ids are visible to the block
            src = dst;
        } else {
            __syncthreads();
        }
        bool bad = s_bad != 0;
        for (int p = 0; p < groups && !bad; p++) {
            int lo = 0, hi = cnt;
            if (groups > 1) {
                hi = gpos[p];
                lo = hi - ghist[p];
            }
            const int np = hi - lo;
            int slots = 512, shift = 32 - 9; // table of >= 1.6 * np slots
            while (slots < K4_BLOCK_SLOTS && 5 * slots < 8 * np) {
                slots <<= 1;
                shift--;
            }
            for (int i = tid; i < slots; i += K4_THREADS) {
                keys[i] = K4_EMPTY;
                cnts[i] = 0u;
            }
            __syncthreads();
            int nff = 0;
            for (int i0 = lo; i0 < hi; i0 += K4_THREADS * 4) { // four independent loads in flight before the atomics
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * K4_THREADS + tid;
                    v[u] = i < hi ? src[i] : K4_EMPTY;
                    if (i < hi && v[u] == K4_EMPTY) nff++; // only when groups == 1
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (v[u] != K4_EMPTY) k4_count(keys, cnts, (uint32_t)slots - 1u, shift, v[u]);
            }
            if (nff) atomicAdd(&s_nff, nff);
            __syncthreads();
            for (int i = tid; i < slots; i += K4_THREADS)
                if (cnts[i] >= (uint32_t)min_hits) {
                    const int at = atomicAdd(&s_ns, 1);
                    if (at < K4_SURV_CAP) surv[at] = keys[i];
                }
            __syncthreads();
            bad = s_ns >= K4_SURV_CAP; // keeps one slot free for the id 0xFFFFFFFF
            __syncthreads();
        }
        if (bad) {
            if (tid == 0) sort_list[atomicAdd(sort_count, 1)] = seg;
            __syncthreads();
            continue;
        }
        if (tid == 0 && s_nff >= min_hits) surv[s_ns++] = K4_EMPTY;
        __syncthreads();
        const int ns = s_ns;
        if (ns > 1) bitonic_sort<true>(surv, ns, tid, K4_THREADS);
        for (int i = tid; i < ns; i += K4_THREADS) g[i] = surv[i];
        if (tid == 0) new_counts[seg] = ns;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K4_THREADS) filter_large_kernel(uint32_t* __restrict__ values,
                                                                  const int32_t* __restrict__ offsets, int min_hits,
                                                                  int32_t* __restrict__ new_counts,
                                                                  const int32_t* __restrict__ big_list,
                                                                  const int32_t* __restrict__ big_count)
{
    extern __shared__ uint32_t k4_dyn[];
    uint32_t* sdata = k4_dyn; // K4_BLOCK_CAP ids
    __shared__ int wsum[K4_THREADS / 32];
    __shared__ int s_out;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nbig = *big_count;
    for (int bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
        const int seg = big_list[bi];
        const int b = offsets[seg];
        const int cnt = offsets[seg + 1] - b;
        uint32_t* g = values + b;
        uint32_t* s;
        if (cnt <= K4_BLOCK_CAP) {
            for (int i = tid; i < cnt; i += K4_THREADS) sdata[i] = g[i];
            s = sdata;
        } else {
            s = g; // in place in global memory
        }
        __syncthreads();
        bitonic_sort<true>(s, cnt, tid, K4_THREADS);
        if (tid == 0) s_out = 0;
        __syncthreads();
        for (int base = 0; base < cnt; base += K4_THREADS) {
            const int i = base + tid;
            const bool keep = keep_head(s, i, cnt, min_hits);
            const uint32_t v = i < cnt ? s[i] : 0u;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) wsum[wid] = __popc(m);
            __syncthreads(); // all reads of this chunk are done before any write
            int before = s_out;
            for (int w = 0; w < wid; w++) before += wsum[w];
            if (keep) g[before + __popc(m & ((1u << lane) - 1u))] = v;
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < K4_THREADS / 32; w++) tot += wsum[w];
                s_out += tot;
            }
            __syncthreads();
        }
        if (tid == 0) new_counts[seg] = s_out;
        __syncthreads();
    }
}

// gather the kept heads of every segment to their final dense position; one warp per segment
__global__ void __launch_bounds__(256) compact_segments_kernel(const uint32_t* __restrict__ values,
                                                               const int32_t* __restrict__ old_offsets,
                                                               const int32_t* __restrict__ new_offsets, int n,
                                                               uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warp0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    for (int seg = warp0; seg < n; seg += nwarps) {
        const int ob = old_offsets[seg], nb = new_offsets[seg];
        const int cnt = new_offsets[seg + 1] - nb;
        for (int i = lane; i < cnt; i += 32) out[nb + i] = values[ob + i];
    }
}

__global__ void __launch_bounds__(256) segment_ids_kernel(const int32_t* __restrict__ offsets, int n, int64_t total,
                                                          int32_t* __restrict__ seg_ids)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int lo = 0, hi = n; // largest s with offsets[s] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)offsets[mid] <= e) lo = mid;
            else hi = mid;
        }
        seg_ids[e] = lo;
    }
}

static unsigned k4_grid(int64_t threads_wanted, int waves)
{
    int64_t g = HRM_SDIV(threads_wanted, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// Sort+filter every segment in place (kept heads at the front of each segment), write the new
// counts and their exclusive scan.  d_new_offsets: n+1.  d_total64: device int64.
hrm_status filter_segments(uint32_t* d_values, uint32_t* d_scratch, const int32_t* d_offsets, int n, int min_hits,
                           int32_t* d_new_counts, int32_t* d_new_offsets, int64_t* d_total64, cudaStream_t s)
{
    if (n == 0) return exclusive_scan_i32(d_new_counts, d_new_offsets, 0, d_total64, s);
    Scratch big;
    HRM_TRY(big.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    int32_t* big_count = big.as<int32_t>();
    int32_t* big_list = big.as<int32_t>() + 1;
    HRM_CUDA(cudaMemsetAsync(big_count, 0, sizeof(int32_t), s));
    HRM_LAUNCH(filter_small_kernel, k4_grid((int64_t)n * 32, 16), K4_THREADS, 0, s, d_values, d_offsets, n, min_hits,
               d_new_counts, big_list, big_count);
    // the attribute is per device: set before every launch (a process may drive several GPUs)
    HRM_CUDA(cudaFuncSetAttribute(filter_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(uint32_t) * K4_BLOCK_CAP)));
    if (min_hits >= 2) { // counting kernel first; what it cannot hold goes on to the sort kernel
        Scratch srt;
        HRM_TRY(srt.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
        int32_t* sort_count = srt.as<int32_t>();
        int32_t* sort_list = srt.as<int32_t>() + 1;
        HRM_CUDA(cudaMemsetAsync(sort_count, 0, sizeof(int32_t), s));
        const size_t smem_count = sizeof(uint32_t) * (2 * K4_BLOCK_SLOTS + K4_SURV_CAP + 2 * K4_MAX_GROUPS);
        HRM_CUDA(cudaFuncSetAttribute(filter_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        HRM_LAUNCH(filter_count_kernel, (unsigned)(num_sms() * 3), K4_THREADS, smem_count, s, d_values, d_scratch, d_offsets,
                   min_hits,
                   d_new_counts, big_list, big_count, sort_list, sort_count);
        HRM_LAUNCH(filter_large_kernel, (unsigned)(num_sms() * 2), K4_THREADS, sizeof(uint32_t) * K4_BLOCK_CAP, s, d_values,
                   d_offsets, min_hits, d_new_counts, sort_list, sort_count);
        return exclusive_scan_i32(d_new_counts, d_new_offsets, n, d_total64, s);
    }
    HRM_LAUNCH(filter_large_kernel, (unsigned)(num_sms() * 2), K4_THREADS, sizeof(uint32_t) * K4_BLOCK_CAP, s, d_values,
               d_offsets, min_hits, d_new_counts, big_list, big_count);
    return exclusive_scan_i32(d_new_counts, d_new_offsets, n, d_total64, s);
}

hrm_status compact_segments(const uint32_t* d_values, const int32_t* d_old_offsets, const int32_t* d_new_offsets, int n,
                            uint32_t* d_out, cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(compact_segments_kernel, k4_grid((int64_t)n * 32, 16), 256, 0, s, d_values, d_old_offsets, d_new_offsets,
               n, d_out);
    return HRM_OK;
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_filter_by_frequency(uint32_t* d_values, int32_t* d_num_per_seq, int32_t* d_offsets, int n,
                                              int min_hits, int64_t* h_total, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0, "n");
    cudaStream_t s = as_stream(stream);
    if (n == 0) {
        if (h_total) *h_total = 0;
        return HRM_OK;
    }
    Scratch newoff, tot, part;
    HRM_TRY(newoff.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    HRM_TRY(tot.alloc(sizeof(int64_t), s));
    int32_t total_in = 0; // ref: the reference sizes its temporaries from the host-known total too (main_gpu.cu:225)
    HRM_CUDA(cudaMemcpyAsync(&total_in, d_offsets + n, sizeof total_in, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    HRM_REQUIRE(total_in >= 0, "offsets[n] is negative");
    HRM_TRY(part.alloc(sizeof(uint32_t) * (size_t)(total_in > 0 ? total_in : 1), s));
    HRM_TRY(filter_segments(d_values, part.as<uint32_t>(), d_offsets, n, min_hits, d_num_per_seq, newoff.as<int32_t>(),
                            tot.as<int64_t>(), s));
    int64_t total = 0;
    HRM_CUDA(cudaMemcpyAsync(&total, tot.p, sizeof total, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s)); // ref: main_gpu.cu:266-275 synchronises for the total too
    if (total > 0) {
        Scratch tmp;
        HRM_TRY(tmp.alloc(sizeof(uint32_t) * (size_t)total, s));
        HRM_TRY(compact_segments(d_values, d_offsets, newoff.as<int32_t>(), n, tmp.as<uint32_t>(), s));
        HRM_CUDA(cudaMemcpyAsync(d_values, tmp.p, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToDevice, s));
    }
    HRM_CUDA(cudaMemcpyAsync(d_offsets, newoff.p, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, s));
    if (h_total) *h_total = total;
    return HRM_OK;
}

extern "C" hrm_status hrm_segment_ids(const int32_t* d_offsets, int n, int64_t total, int32_t* d_segment_ids,
                                      hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && total >= 0, "sizes");
    if (total == 0 || n == 0) return HRM_OK;
    HRM_LAUNCH(segment_ids_kernel, k4_grid(total, 16), 256, 0, as_stream(stream), d_offsets, n, total, d_segment_ids);
    return HRM_OK;
}
