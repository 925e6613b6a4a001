// k4_sort.cuh -- normalised bitonic network in shared (or global) memory, shared by the K4 kernels.
#pragma once
#include <stdint.h>

namespace hrm {

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

} // namespace hrm
