// hrm_common.cuh -- shared helpers of libhrm_b200 (sm_100a).  The per-item arithmetic lives in
// HRM_HD functions so that tests/host_harness can compile the very same code with g++ and check
// it on the CPU box before GPU time is spent; the shipped library only ever runs it on the device.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define HRM_HD __host__ __device__ __forceinline__
#define HRM_D __device__ __forceinline__
#else
#define HRM_HD inline
#define HRM_D inline
#endif

#ifndef HRM_SDIV
#define HRM_SDIV(a, b) (((a) + (b)-1) / (b))
#endif

namespace hrm {

HRM_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

HRM_HD uint32_t brev32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

HRM_HD uint64_t brev64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    return ((uint64_t)brev32((uint32_t)x) << 32) | brev32((uint32_t)(x >> 32));
#endif
}

// (hi:lo) << s, upper 32 bits; 0 <= s <= 31
HRM_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, int s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s == 0 ? hi : (hi << s) | (lo >> (32 - s));
#endif
}

// murmur64 finaliser (ref: include/helpers/hashers.cuh:129-137)
HRM_HD uint64_t murmur64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// 2-bit code of base i of a packed sequence (ref: include/sequencehelpers.hpp:228-233)
HRM_HD uint32_t get_nuc(const uint32_t* enc, int64_t i)
{
    return (enc[i >> 4] >> (30 - 2 * (int)(i & 15))) & 3u;
}

// keep the even (low) bit of each 2-bit code, compacted to 16 bits
// (ref: SequenceHelpers::extractEvenBits include/sequencehelpers.hpp:644-652)
HRM_HD uint32_t extract_even_bits(uint32_t x)
{
    x = x & 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// 32 bits of a packed 2-bit stream starting at base `b` (16 bases), words beyond `nwords` read as 0
HRM_HD uint32_t stream16(const uint32_t* w, int64_t nwords, int64_t b)
{
    const int64_t i = b >> 4;
    const int s = (int)(b & 15) * 2;
    const uint32_t a = i < nwords ? w[i] : 0u;
    const uint32_t c = (s != 0 && i + 1 < nwords) ? w[i + 1] : 0u;
    return funnel_l(a, c, s);
}

// hi/lo bit planes of the 32 bases starting at base b (MSB = first base)
// (ref: HiLo layout include/sequencehelpers.hpp:415-457, conversion :654-690)
HRM_HD void planes32(const uint32_t* w, int64_t nwords, int64_t b, uint32_t& hi, uint32_t& lo)
{
    const uint32_t x0 = stream16(w, nwords, b);
    const uint32_t x1 = stream16(w, nwords, b + 16);
    hi = (extract_even_bits(x0 >> 1) << 16) | extract_even_bits(x1 >> 1);
    lo = (extract_even_bits(x0) << 16) | extract_even_bits(x1);
}

} // namespace hrm
