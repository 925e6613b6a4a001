// k4_collect.cu -- K4: candidate collection = per-segment sort + run-length reduction + threshold.
// ref: GpuSegmentedUniqueByCount::unique include/gpu/cuda_unique_by_count.cuh:33-215 (cub segmented
//      radix sort + ~9 thrust launches) via GpuMinhashQueryFilter::keepDistinctByFrequency
//      include/gpu/minhashqueryfilter.cuh:239-278; keepDistinct :217-236 (minTableHits <= 1).
// Here: one warp per segment sorts in shared memory with a normalised bitonic network (all
// compare-exchanges ascending, so virtual +inf padding needs no storage), detects run heads and
// keeps a head iff the element minHits-1 places further is equal (sorted => multiplicity >= minHits).
// Segments above 256 ids go to a block-per-segment kernel (shared memory up to 8192 ids, in-place
// in global memory beyond that).  Then one scan and one gather compact the lists.
// Traffic: 4 B read + <= 4 B written per candidate id, plus 8 B per segment.
#include "runtime.cuh"

namespace hrm {

constexpr int K4_WARP_CAP = 256;    // ids per segment handled by one warp
constexpr int K4_BLOCK_CAP = 8192;  // ids per segment sorted in shared memory by one block
constexpr int K4_THREADS = 256;

__device__ __forceinline__ void cmpswap(uint32_t* s, int lo, int hi)
{
    const uint32_t a = s[lo], b = s[hi];
    if (a > b) {
        s[lo] = b;
        s[hi] = a;
    }
}

// normalised bitonic sort of cnt elements by `nthreads` cooperating threads (tid in [0,nthreads))
template <bool BLOCK>
__device__ __forceinline__ void bitonic_sort(uint32_t* s, int cnt, int tid, int nthreads)
{
    int npow = 1;
    while (npow < cnt) npow <<= 1;
    const int half = npow >> 1;
    for (int k = 2; k <= npow; k <<= 1) {
        const int hk = k >> 1;
        for (int i = tid; i < half; i += nthreads) {
            const int blk = i / hk, r = i - blk * hk;
            const int lo = blk * k + r, hi = blk * k + k - 1 - r;
            if (hi < cnt) cmpswap(s, lo, hi);
        }
        if (BLOCK) __syncthreads();
        else __syncwarp();
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int i = tid; i < half; i += nthreads) {
                const int lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                if (hi < cnt) cmpswap(s, lo, hi);
            }
            if (BLOCK) __syncthreads();
            else __syncwarp();
        }
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
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* s = sm[wid];
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

__global__ void __launch_bounds__(K4_THREADS) filter_large_kernel(uint32_t* __restrict__ values,
                                                                  const int32_t* __restrict__ offsets, int min_hits,
                                                                  int32_t* __restrict__ new_counts,
                                                                  const int32_t* __restrict__ big_list,
                                                                  const int32_t* __restrict__ big_count)
{
    extern __shared__ uint32_t sdata[]; // K4_BLOCK_CAP ids
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
hrm_status filter_segments(uint32_t* d_values, const int32_t* d_offsets, int n, int min_hits, int32_t* d_new_counts,
                           int32_t* d_new_offsets, int64_t* d_total64, cudaStream_t s)
{
    if (n == 0) return exclusive_scan_i32(d_new_counts, d_new_offsets, 0, d_total64, s);
    Scratch big;
    HRM_TRY(big.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    int32_t* big_count = big.as<int32_t>();
    int32_t* big_list = big.as<int32_t>() + 1;
    HRM_CUDA(cudaMemsetAsync(big_count, 0, sizeof(int32_t), s));
    HRM_LAUNCH(filter_small_kernel, k4_grid((int64_t)n * 32, 16), K4_THREADS, 0, s, d_values, d_offsets, n, min_hits,
               d_new_counts, big_list, big_count);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(filter_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(uint32_t) * K4_BLOCK_CAP));
        attr_set = true;
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
    Scratch newoff, tot;
    HRM_TRY(newoff.alloc(sizeof(int32_t) * ((size_t)n + 1), s));
    HRM_TRY(tot.alloc(sizeof(int64_t), s));
    HRM_TRY(filter_segments(d_values, d_offsets, n, min_hits, d_num_per_seq, newoff.as<int32_t>(), tot.as<int64_t>(), s));
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
