// core_pack.cuh -- K1 arithmetic: 3N conversion + 2-bit packing of 16 ASCII bases into one word.
// ref: SequenceHelpers::encodeSequence2Bit include/sequencehelpers.hpp:185-218 (codes, bit order,
//      left-aligned tail); Mappinghandler::NucleoideConverer src/gpu/mappinghandler.cu:163-179 (C->T).
#pragma once
#include "hrm_common.cuh"

namespace hrm {

// per-byte equality mask (0xff where equal) of the four bytes of x against byte c
HRM_HD uint32_t bytes_eq(uint32_t x, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vcmpeq4(x, c * 0x01010101u);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((x >> (8 * i)) & 0xffu) == c) r |= 0xffu << (8 * i);
    return r;
#endif
}

// four ASCII bytes (first base in the lowest byte, as loaded little-endian) -> 8 bits of codes,
// first base in bits 7:6.  conv: 0 none, 1 C->T, 2 G->A.  Only exact 'A','C','G','T' get a
// non-zero code; everything else (N, lower case, padding) packs as A = 0.
HRM_HD uint32_t pack4(uint32_t x, int conv)
{
    const uint32_t isC = bytes_eq(x, 'C');
    const uint32_t isG = bytes_eq(x, 'G');
    const uint32_t isT = bytes_eq(x, 'T');
    const uint32_t cC = conv == 1 ? 0x03030303u : 0x01010101u; // C->T: C packs as T
    const uint32_t cG = conv == 2 ? 0x00000000u : 0x02020202u; // G->A: G packs as A
    const uint32_t codes = (isC & cC) | (isG & cG) | (isT & 0x03030303u);
    return ((codes << 6) & 0xC0u) | ((codes >> 4) & 0x30u) | ((codes >> 14) & 0x0Cu) | ((codes >> 24) & 0x03u);
}

// zero the bytes at positions >= valid (0..4) of a little-endian 4-byte group
HRM_HD uint32_t keep_bytes(uint32_t x, int valid)
{
    if (valid >= 4) return x;
    if (valid <= 0) return 0u;
    return x & (0xFFFFFFFFu >> (8 * (4 - valid)));
}

// 16 ASCII bytes given as four little-endian words; only the first `valid` bytes (1..16) count
HRM_HD uint32_t pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d, int valid, int conv)
{
    if (valid < 16) {
        a = keep_bytes(a, valid);
        b = keep_bytes(b, valid - 4);
        c = keep_bytes(c, valid - 8);
        d = keep_bytes(d, valid - 12);
    }
    return (pack4(a, conv) << 24) | (pack4(b, conv) << 16) | (pack4(c, conv) << 8) | pack4(d, conv);
}

// byte-wise fallback for unaligned input
HRM_HD uint32_t pack16_bytes(const char* p, int valid, int conv)
{
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < valid; i++) w[i >> 2] |= (uint32_t)(unsigned char)p[i] << (8 * (i & 3));
    return pack16(w[0], w[1], w[2], w[3], valid, conv);
}

} // namespace hrm
