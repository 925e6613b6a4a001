// runtime.cuh -- thin host runtime of libhrm_b200: status codes, error text, stream-ordered
// scratch, launch accounting.  Replaces the reference's rmm pools / CUDACHECK glue
// (ref: include/gpu/cudaerrorcheck.cuh:42-58, include/gpu/rmm_utilities.cuh) -- no rmm, no thrust.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include "../../include/hrm_b200.h"

#ifndef HRM_SDIV
#define HRM_SDIV(a, b) (((a) + (b)-1) / (b))
#endif

namespace hrm {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches; // kernels launched by this library (hrm_batch_stats)

inline cudaStream_t as_stream(hrm_stream s) { return (cudaStream_t)s; }

#define HRM_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::hrm::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return e__ == cudaErrorMemoryAllocation ? HRM_ERR_NOMEM : HRM_ERR_CUDA;                 \
        }                                                                                           \
    } while (0)

#define HRM_REQUIRE(cond, msg)                                              \
    do {                                                                    \
        if (!(cond)) {                                                      \
            ::hrm::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return HRM_ERR_INVALID;                                         \
        }                                                                   \
    } while (0)

#define HRM_TRY(call)                     \
    do {                                  \
        hrm_status s__ = (call);          \
        if (s__ != HRM_OK) return s__;    \
    } while (0)

// counted kernel launch + launch-error check
#define HRM_LAUNCH(kernel, grid, block, smem, stream, ...)                  \
    do {                                                                    \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);         \
        ::hrm::g_launches.fetch_add(1, std::memory_order_relaxed);          \
        HRM_CUDA(cudaGetLastError());                                       \
    } while (0)

// stream-ordered scratch (cudaMallocAsync on the device's default pool, which this library
// configures to keep freed memory cached: no per-call allocation cost in steady state)
hrm_status scratch_alloc(void** p, size_t bytes, cudaStream_t s);
void scratch_free(void* p, cudaStream_t s);
hrm_status ensure_device();      // HRM_ERR_CUDA when no CUDA device is usable (no CPU fallback)
int num_sms();

// RAII scratch for use inside functions that return hrm_status
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = 0;
    Scratch() {}
    ~Scratch() { release(); }
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    hrm_status alloc(size_t bytes, cudaStream_t stream)
    {
        release();
        s = stream;
        return scratch_alloc(&p, bytes ? bytes : 16, stream);
    }
    void release()
    {
        if (p) scratch_free(p, s);
        p = nullptr;
    }
    template <class T> T* as() const { return (T*)p; }
};

// persistent device buffer that only grows (handle-owned scratch; ref: per-handle QueryData,
// fakegpuminhasher.cuh:60-142)
struct GrowBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~GrowBuf() { if (p) cudaFree(p); }
    hrm_status reserve(size_t bytes)
    {
        if (bytes <= cap) return HRM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            p = nullptr;
            return HRM_ERR_NOMEM;
        }
        cap = want;
        return HRM_OK;
    }
    template <class T> T* as() const { return (T*)p; }
};

// ---- device-wide exclusive scan of int32 (hand-written; used on the hot path) -----------------
// out[i] = sum_{t<i} in[t] for i in [0, n]; out has n+1 entries; *d_total64 (optional) = int64 sum.
hrm_status exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int64_t* d_total64,
                              cudaStream_t s);
hrm_status exclusive_scan_i32_to_i64(const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* d_total64,
                                     cudaStream_t s);
size_t exclusive_scan_scratch_bytes(int64_t n);

} // namespace hrm
