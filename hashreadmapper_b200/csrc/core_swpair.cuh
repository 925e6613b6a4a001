// core_swpair.cuh -- V2 arithmetic, SIMD form: the two alignments of one mapped read (read and RC(read)
// against the same window; ref: mappinghandler.cu:556-595) run in the two 16-bit halves of every register,
// on the Blackwell DPX instructions (VIMNMX3.S16x2, VIADDMNMX.S16x2).  Same DP as core_sw.cuh: sw_pass
// (ref: sw_sse2_byte src/ssw.c:197-386, sw_sse2_word :412-588), restated per lane of a G-lane group:
//
//   lane l owns rows [l*R, (l+1)*R) of a G*R-row frame and sweeps the reference columns as a wavefront
//   (column t-l at step t); the bottom H / F of its strip and the running column maximum go to lane l+1.
//
// Frame layout, forward pass (read length L, padW = L rounded up to 8, padB = to 16, Rtot = G*R):
//   rows [top, top+L) with top = Rtot-8-padW hold the read; rows above are EMPTY (never match: they stay 0
//   and act as the matrix border); rows below are the striped kernels' PAD rows (score 0 against anything).
//   The word-mode pad boundary is therefore always after row Rtot-9 and the byte-mode boundary after row
//   Rtot-9 (padB == padW) or Rtot-1: the column maxima that feed the second-best score are tapped at fixed
//   rows.  Pad rows only ever occur in the last 16 rows of the last lane.
// Value domains (all halves non-negative, so plain 32-bit IMAD/IADD act on both halves without borrows):
//   S = h, E2 = E+2, F2 = F+2, t = (diag + s) + 2 = S_diag + K*M with M the profile's match bit, K = 4
//   (match +2 / mismatch -2) or, for pad rows, M = 1 and K = 2 (score 0).
#pragma once
#include "hrm_common.cuh"

namespace hrm {

// ---- packed signed 16x2 primitives --------------------------------------------------------------
#if defined(__CUDA_ARCH__)
HRM_HD uint32_t px_max(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
HRM_HD uint32_t px_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
HRM_HD uint32_t px_addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
#else
HRM_HD int16_t px_lo(uint32_t x) { return (int16_t)(x & 0xFFFFu); }
HRM_HD int16_t px_hi(uint32_t x) { return (int16_t)(x >> 16); }
HRM_HD uint32_t px_pack(int lo, int hi) { return ((uint32_t)(uint16_t)(int16_t)lo) | ((uint32_t)(uint16_t)(int16_t)hi << 16); }
HRM_HD uint32_t px_max(uint32_t a, uint32_t b)
{
    return px_pack(px_lo(a) > px_lo(b) ? px_lo(a) : px_lo(b), px_hi(a) > px_hi(b) ? px_hi(a) : px_hi(b));
}
HRM_HD uint32_t px_max3(uint32_t a, uint32_t b, uint32_t c) { return px_max(px_max(a, b), c); }
HRM_HD uint32_t px_addmax(uint32_t a, uint32_t b, uint32_t c)
{
    return px_max(px_pack((int16_t)(px_lo(a) + px_lo(b)), (int16_t)(px_hi(a) + px_hi(b))), c);
}
#endif

constexpr uint32_t PX_ONES = 0x00010001u;
constexpr uint32_t PX_TWOS = 0x00020002u;
constexpr int PAIR_CODE_PAD = 5;   // row below the read: scores 0 against every reference base
constexpr int PAIR_CODE_EMPTY = 6; // row outside the frame's read: never matches

// 5-bit profile field of one row: bit b = "matches reference base b" (b < 4), bit 4 = selected by a
// reference N (code 4): real rows never match it, pad rows score 0 there as everywhere
HRM_HD uint32_t pair_field(int code)
{
    if (code < 4) return 1u << code;
    if (code == PAIR_CODE_PAD) return 0x1Fu;
    return 0u; // N in the read, EMPTY
}

template <int R>
struct PairLane {
    static constexpr int NP = (R + 2) / 3;
    uint32_t S[R];    // h of the previous column
    uint32_t E[R];    // E + 2
    uint32_t prof[NP]; // 3 rows per register and half: row r at bits 5*(r%3) of register r/3
    uint32_t Kmid[8]; // multiplier of rows R-16 .. R-9
    uint32_t Klast;   // multiplier of rows R-8 .. R-1
};

template <int R>
HRM_HD void pair_reset(PairLane<R>& L)
{
#pragma unroll
    for (int r = 0; r < R; r++) {
        L.S[r] = 0u;
        L.E[r] = PX_TWOS;
    }
}

// codeA/codeB: code of frame row g for the two halves (0..3 base, 4 N, PAD, EMPTY)
template <int R, class CodeFn>
HRM_HD void pair_build(PairLane<R>& L, int row0, CodeFn code)
{
#pragma unroll
    for (int p = 0; p < PairLane<R>::NP; p++) L.prof[p] = 0u;
    L.Klast = 4u;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ca = code(0, row0 + r), cb = code(1, row0 + r);
        L.prof[r / 3] |= (pair_field(ca) | (pair_field(cb) << 16)) << (5 * (r % 3));
        const uint32_t K = (ca == PAIR_CODE_PAD) ? 2u : 4u; // both halves share the read length: pads coincide
        if (r >= R - 16 && r < R - 8) L.Kmid[r - (R - 16)] = K;
        if (r == R - 8) L.Klast = K; // rows R-8.. are all pad or all not (frame layout)
    }
}

// One column of the lane's strip.  rc: reference code 0..4; hmask: PX_ONES restricted to the active halves.
// diag_in: h of the row above the strip in the previous column; F_in: F+2 entering from above.
// kW: running key maximum after row R-9; kAll: after row R-1.  key = h*64 + (63 - r), both halves.
template <int R>
HRM_HD void pair_column(PairLane<R>& L, int rc, uint32_t hmask, uint32_t diag_in, uint32_t F_in, uint32_t& S_bot,
                        uint32_t& F_out, uint32_t& kW, uint32_t& kAll)
{
    const int sh0 = rc, sh1 = rc + 5, sh2 = rc + 10;
    uint32_t diag = diag_in, F = F_in, kc = 0u, kprev = 0u;
    kW = 0u;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int sh = (r % 3 == 0) ? sh0 : ((r % 3 == 1) ? sh1 : sh2);
        const uint32_t M = (L.prof[r / 3] >> sh) & hmask;
        const uint32_t K = (r < R - 16) ? 4u : ((r < R - 8) ? L.Kmid[(r - (R - 16)) & 7] : L.Klast);
        const uint32_t t = M * K + diag;                            // (diag + s) + 2
        const uint32_t H = px_max3(t, L.E[r], F);                   // h + 2 (E2 >= 2 is the zero floor)
        const uint32_t hob = px_addmax(H, 0xFFFDFFFDu, PX_TWOS);    // max(h - gapO, 0) + 2
        L.E[r] = px_addmax(L.E[r], 0xFFFFFFFFu, hob);               // max(E - gapE, h - gapO, 0) + 2
        F = px_addmax(F, 0xFFFFFFFFu, hob);
        const uint32_t Sn = H - PX_TWOS;
        diag = L.S[r];
        L.S[r] = Sn;
        const uint32_t key = Sn * 64u + (uint32_t)(63 - r) * PX_ONES;
        if (r & 1) kc = px_max3(kc, kprev, key);
        else kprev = key;
        if (r == R - 9) kW = (r & 1) ? kc : px_max(kc, key);
    }
    if (R & 1) kc = px_max(kc, kprev);
    S_bot = L.S[R - 1];
    F_out = F;
    kAll = kc;
}

// per-lane running best of one half: strictly larger value wins, so the first column (in processing order)
// and, through the key's low bits, the smallest row of that column are kept
struct PairBest {
    uint32_t key, cmp, col; // key = value*64 + (63 - r); cmp = key | 63
};
HRM_HD void pair_best_reset(PairBest& b)
{
    b.key = 0u;
    b.cmp = 63u;
    b.col = 0u;
}
HRM_HD void pair_best_update(PairBest& b, uint32_t k16, uint32_t col)
{
    const bool up = k16 > b.cmp;
    b.key = up ? k16 : b.key;
    b.cmp = up ? (k16 | 63u) : b.cmp;
    b.col = up ? col : b.col;
}
// (value, first column, smallest frame row) as one comparable word; 0 when nothing positive was seen
HRM_HD uint32_t pair_best_word(const PairBest& b, int row0)
{
    const uint32_t v = b.key >> 6;
    if (v == 0u) return 0u;
    const uint32_t g = (uint32_t)row0 + (63u - (b.key & 63u));
    return (v << 20) | ((1023u - b.col) << 10) | (1023u - g);
}


// ---- one lane of the wavefront ---------------------------------------------------------------------
template <int R>
struct PairWave {
    PairLane<R> L;
    PairBest bestA, bestB;
    uint32_t outS, outF, outCm; // handed to the next lane: bottom h, bottom F+2, running column maximum
    uint32_t prevRecvS;         // bottom h of the lane above, one column back (the strip's first diagonal)
};

template <int R>
HRM_HD void pair_wave_reset(PairWave<R>& w)
{
    pair_reset(w.L);
    pair_best_reset(w.bestA);
    pair_best_reset(w.bestB);
    w.outS = 0u;
    w.outF = PX_TWOS;
    w.outCm = 0u;
    w.prevRecvS = 0u;
}

// One step: the lane processes its column number c (0-based in processing order) if 0 <= c < ncols.
// recv*: the out* values of the lane above after ITS previous step (0 / PX_TWOS / 0 for the first lane).
// rc / hmask: reference code and active halves of that column; order: the column's rank in processing
// order.  cmW / cmB: column maximum (value only) including this lane's rows up to R-9 / R-1.
template <int R>
HRM_HD bool pair_wave_step(PairWave<R>& w, int c, int ncols, int rc, uint32_t hmask, uint32_t recvS, uint32_t recvF,
                           uint32_t recvCm, uint32_t order, uint32_t& cmW, uint32_t& cmB)
{
    if (c < 0 || c >= ncols) return false;
    const uint32_t diag_in = c == 0 ? 0u : w.prevRecvS;
    w.prevRecvS = recvS;
    uint32_t kW, kAll;
    pair_column(w.L, rc, hmask, diag_in, recvF, w.outS, w.outF, kW, kAll);
    pair_best_update(w.bestA, kAll & 0xFFFFu, order);
    pair_best_update(w.bestB, kAll >> 16, order);
    cmW = px_max(recvCm, (kW >> 6) & 0x03FF03FFu);
    cmB = px_max(recvCm, (kAll >> 6) & 0x03FF03FFu);
    w.outCm = cmB;
    return true;
}

// frame geometry of the forward pass
HRM_HD int pair_pad8(int L) { return (L + 7) & ~7; }
HRM_HD int pair_pad16(int L) { return (L + 15) & ~15; }
HRM_HD int pair_top(int rtot, int L) { return rtot - 8 - pair_pad8(L); }
// a read of length L and a window of length wl fit a G*R frame (keys are 16-bit: h*64+63 < 32768)
HRM_HD bool pair_fits(int rtot, int L, int wl)
{
    const int m = L < wl ? L : wl;
    return L >= 1 && pair_pad8(L) + 8 <= rtot && rtot <= 1023 && wl <= 1023 && 2 * m <= 500;
}

} // namespace hrm
