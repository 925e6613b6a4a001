// core_shd.cuh -- S3 arithmetic: shifted Hamming distance on hi/lo bit planes.
// ref: shiftedHammingDistanceWithFullOverlapKernelSmem1 src/gpu/hammingdistancekernels.cu:132-263,
//      hammingdistanceHiLo :78-118, reverseComplementSequenceInplace2BitHiLo
//      include/sequencehelpers.hpp:580-614.
//
// The reference walks shifts sequentially with an early exit that depends on the best score so far.
// For accepted results (orientation != None) that is exactly "first minimum of the exact Hamming
// distance over (orientation, shift) in the order forward 0..S, reverse-complement 0..S" -- an
// order-independent statement, so shifts can be evaluated in parallel and reduced with the key
// (hd, orientation, shift).  Rejected results carry unspecified shift/score in the reference
// (partial sums).
#pragma once
#include "hrm_common.cuh"

namespace hrm {

#define HRM_SHD_MAX_READ_WORDS 16   /* 512 bases */
#define HRM_SHD_MAX_ANCHOR_WORDS 40 /* 1280 bases */
#define HRM_SHD_INF 0x7FFFFFFFu

// threshold: int(float(Lc) * rate) in float32 (ref: hammingdistancekernels.cu:212,246)
HRM_HD int shd_threshold(int Lc, float rate) { return (int)((float)Lc * rate); }

// word j (of nw) of the reverse-complemented plane; plane = left-aligned Lc bits, zero padded.
HRM_HD uint32_t rc_plane_word(const uint32_t* plane, int nw, int Lc, int j)
{
    const int pad = nw * 32 - Lc; // 0..31
    const uint32_t t0 = brev32(plane[nw - 1 - j]);
    const uint32_t t1 = (j + 1 < nw) ? brev32(plane[nw - 2 - j]) : 0u;
    uint32_t r = ~funnel_l(t0, t1, pad);
    const int bits_before = 32 * j;
    const int remain = Lc - bits_before; // valid bits in this word
    if (remain < 32) r &= remain <= 0 ? 0u : (0xFFFFFFFFu << (32 - remain));
    return r;
}

// exact Hamming distance of read planes (Lc bits) against anchor planes shifted by s, stopping
// early (returning a value > limit) once the running sum exceeds `limit`.
HRM_HD int shd_at_shift(const uint32_t* ahi, const uint32_t* alo, int anw, const uint32_t* rhi,
                        const uint32_t* rlo, int Lc, int s, int limit)
{
    const int nw = HRM_SDIV(Lc, 32);
    const int wi = s >> 5, sh = s & 31;
    int acc = 0;
    for (int j = 0; j < nw; j++) {
        const int a = wi + j;
        const uint32_t h0 = a < anw ? ahi[a] : 0u, h1 = a + 1 < anw ? ahi[a + 1] : 0u;
        const uint32_t l0 = a < anw ? alo[a] : 0u, l1 = a + 1 < anw ? alo[a + 1] : 0u;
        uint32_t bits = (funnel_l(h0, h1, sh) ^ rhi[j]) | (funnel_l(l0, l1, sh) ^ rlo[j]);
        const int remain = Lc - 32 * j;
        if (remain < 32) bits &= 0xFFFFFFFFu << (32 - remain);
        acc += popc32(bits);
        if (acc > limit) return acc;
    }
    return acc;
}

// reduction key: smaller is better; ties -> forward before RC, smaller shift first
HRM_HD uint32_t shd_key(int hd, int orientation01, int shift)
{
    return ((uint32_t)hd << 20) | ((uint32_t)orientation01 << 16) | (uint32_t)shift;
}

// Sequential evaluation of one candidate (host harness + generic fallback).
// out: shift, score, orientation (1 fwd, 2 rc, 3 none)
HRM_HD void shd_sequential(const uint32_t* anchor, int64_t anchor_words, int64_t anchor_base, int La,
                           const uint32_t* read, int64_t read_words, int Lc, float rate, int* out_shift,
                           int* out_score, int* out_orientation)
{
    if (Lc > La) {
        *out_shift = 0;
        *out_score = Lc;
        *out_orientation = 3;
        return;
    }
    uint32_t ahi[HRM_SHD_MAX_ANCHOR_WORDS], alo[HRM_SHD_MAX_ANCHOR_WORDS];
    uint32_t rhi[2][HRM_SHD_MAX_READ_WORDS], rlo[2][HRM_SHD_MAX_READ_WORDS];
    const int anw = HRM_SDIV(La, 32), nw = HRM_SDIV(Lc, 32);
    for (int j = 0; j < anw; j++) {
        planes32(anchor, anchor_words, anchor_base + 32 * (int64_t)j, ahi[j], alo[j]);
        const int remain = La - 32 * j;
        if (remain < 32) {
            ahi[j] &= 0xFFFFFFFFu << (32 - remain);
            alo[j] &= 0xFFFFFFFFu << (32 - remain);
        }
    }
    for (int j = 0; j < nw; j++) {
        planes32(read, read_words, 32 * (int64_t)j, rhi[0][j], rlo[0][j]);
        const int remain = Lc - 32 * j;
        if (remain < 32) {
            rhi[0][j] &= 0xFFFFFFFFu << (32 - remain);
            rlo[0][j] &= 0xFFFFFFFFu << (32 - remain);
        }
    }
    for (int j = 0; j < nw; j++) {
        rhi[1][j] = rc_plane_word(rhi[0], nw, Lc, j);
        rlo[1][j] = rc_plane_word(rlo[0], nw, Lc, j);
    }
    const int thr = shd_threshold(Lc, rate);
    uint32_t best = HRM_SHD_INF;
    for (int o = 0; o < 2; o++)
        for (int s = 0; s <= La - Lc; s++) {
            const int hd = shd_at_shift(ahi, alo, anw, rhi[o], rlo[o], Lc, s, thr);
            if (hd <= thr) {
                const uint32_t key = shd_key(hd, o, s);
                if (key < best) best = key;
            }
        }
    if (best == HRM_SHD_INF) {
        *out_shift = 0;
        *out_score = thr + 1;
        *out_orientation = 3;
    } else {
        *out_shift = (int)(best & 0xFFFFu);
        *out_score = (int)(best >> 20);
        *out_orientation = ((best >> 16) & 1u) ? 2 : 1;
    }
}

// extended-window location (ref: detail::computeWindowLocation
// include/gpu/windowgenerationkernels.cuh:17-48) with the section end replaced by the chromosome
// end: the reference's per-batch section [pos0 - M/2, pos_last + w + M/2) only clips differently
// from the chromosome when extension > M/2, which cannot happen (extension = readLen/2 <= M/2).
HRM_HD void window_location(int64_t chromLen, int64_t windowPos, int windowSize, int extension, int* left,
                            int* right, int* length)
{
    int len = windowSize, l = 0, r = 0;
    if ((int64_t)extension < windowPos) {
        l = extension;
        len += extension;
    }
    if (windowPos + windowSize <= chromLen) {
        if (windowPos + windowSize + extension < chromLen) r = extension;
        else r = (int)(chromLen - (windowPos + windowSize));
        len += r;
    } else {
        len -= (int)((windowPos + windowSize) - chromLen);
    }
    *left = l;
    *right = r;
    *length = len;
}

} // namespace hrm
