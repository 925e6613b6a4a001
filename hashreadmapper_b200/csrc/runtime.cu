// runtime.cu -- see runtime.cuh
#include "runtime.cuh"
#include <stdarg.h>
#include <mutex>

namespace hrm {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static std::once_flag g_pool_once[64];

hrm_status ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libhrm_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return HRM_ERR_CUDA;
    }
    int dev = 0;
    HRM_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) {
        std::call_once(g_pool_once[dev], [dev]() {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t thr = UINT64_MAX; // keep freed scratch cached in the pool
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
                // scratch freed on one stream must not be handed to another stream by making that stream WAIT for the
                // free (the default policy): it would serialise the verification stream and the seeding stream of
                // the staged pipeline.  Memory whose free has completed is still reused across streams.
                int off = 0;
                cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
            }
        });
    }
    return HRM_OK;
}

int num_sms()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

hrm_status scratch_alloc(void** p, size_t bytes, cudaStream_t s)
{
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes, s);
    if (e != cudaSuccess) {
        set_error("cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? HRM_ERR_NOMEM : HRM_ERR_CUDA;
    }
    return HRM_OK;
}

void scratch_free(void* p, cudaStream_t s)
{
    if (p) cudaFreeAsync(p, s);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan: reduce per tile -> scan of tile sums (one block) -> scan per tile + offset
// ---------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t warp_incl_scan64(int64_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int64 per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ int64_t block_excl_scan64(int64_t v, int64_t* total)
{
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    __shared__ int64_t tot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t incl = warp_incl_scan64(v);
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int64_t w = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        const int64_t wi = warp_incl_scan64(w);
        if (lane < SCAN_THREADS / 32) wsum[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) tot = wi;
    }
    __syncthreads();
    const int64_t r = incl - v + wsum[wid];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                      int64_t* __restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t v = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        const int64_t i = base + (int64_t)t * SCAN_THREADS + threadIdx.x;
        if (i < n) v += in[i];
    }
    int64_t total;
    block_excl_scan64(v, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(int64_t* __restrict__ tile_sums,
                                                                         int64_t ntiles,
                                                                         int64_t* __restrict__ d_total)
{
    int64_t carry = 0;
    for (int64_t base = 0; base < ntiles; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < ntiles ? tile_sums[i] : 0;
        int64_t total;
        const int64_t ex = block_excl_scan64(v, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && d_total) *d_total = carry;
}

template <class OutT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* __restrict__ in,
                                                                  OutT* __restrict__ out, int64_t n,
                                                                  const int64_t* __restrict__ tile_offsets)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t sum = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        const int64_t i = base + t;
        v[t] = i < n ? in[i] : 0;
        sum += v[t];
    }
    int64_t total;
    int64_t ex = block_excl_scan64(sum, &total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        const int64_t i = base + t;
        if (i < n) out[i] = (OutT)ex;
        ex += v[t];
        if (i == n - 1) out[n] = (OutT)ex;
    }
}

template <class OutT>
__global__ void scan_empty_kernel(OutT* out, int64_t* d_total)
{
    out[0] = 0;
    if (d_total) *d_total = 0;
}

size_t exclusive_scan_scratch_bytes(int64_t n) { return sizeof(int64_t) * (size_t)(HRM_SDIV(n, (int64_t)SCAN_TILE) + 1); }

template <class OutT>
static hrm_status exclusive_scan_impl(const int32_t* d_in, OutT* d_out, int64_t n, int64_t* d_total64, cudaStream_t s)
{
    if (n <= 0) {
        HRM_LAUNCH(scan_empty_kernel<OutT>, 1, 1, 0, s, d_out, d_total64);
        return HRM_OK;
    }
    const int64_t ntiles = HRM_SDIV(n, (int64_t)SCAN_TILE);
    Scratch tiles;
    HRM_TRY(tiles.alloc(sizeof(int64_t) * (size_t)ntiles, s));
    HRM_LAUNCH(scan_tile_sums_kernel, (unsigned)ntiles, SCAN_THREADS, 0, s, d_in, n, tiles.as<int64_t>());
    HRM_LAUNCH(scan_tile_offsets_kernel, 1, SCAN_THREADS, 0, s, tiles.as<int64_t>(), ntiles, d_total64);
    HRM_LAUNCH(scan_apply_kernel<OutT>, (unsigned)ntiles, SCAN_THREADS, 0, s, d_in, d_out, n, tiles.as<int64_t>());
    return HRM_OK;
}

hrm_status exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int64_t* d_total64, cudaStream_t s)
{
    return exclusive_scan_impl<int32_t>(d_in, d_out, n, d_total64, s);
}
// same input, 64-bit offsets (byte offsets of text lines: a batch of SAM text exceeds 2 GiB)
hrm_status exclusive_scan_i32_to_i64(const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total64, cudaStream_t s)
{
    return exclusive_scan_impl<int64_t>(d_in, d_out, n, d_total64, s);
}

} // namespace hrm

extern "C" {

const char* hrm_last_error(void) { return hrm::g_err; }
int hrm_abi_version(void) { return HRM_ABI_VERSION; }
int hrm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

} // extern "C"
