// core_swband.cuh -- V2 trace back: ONE band iteration of the reference's banded_sw restated in
// diagonal coordinates, one alignment per thread, plus the trace back and the =/X/I/D/S string.
// ref: banded_sw src/ssw.c:590-774 (band doubling :614-672, trace back :675-764), ConvertAlignment
//      src/ssw_cpp.cpp:54-90, CalculateNumberMismatch :126-210.
//
// The reference keeps a row of the band in h_b / e_b / h_c indexed by u = j - max(i - band, 0) + 1 and
// shifts its frame once i > band.  Here cell (i, j) lives at d = j - i + band (0 <= d <= 2*band), so
//   up (i-1, j) = slot d+1,   diagonal (i-1, j-1) = slot d,   left (i, j-1) = slot d-1
// for EVERY row, and one array of (H, E) pairs is updated in place, ascending d.  Slots that were never
// written hold 0, which is exactly what the reference reads outside the band: h_b[0] (left border),
// and the slot it zeroes before every row, h_b[edge] / e_b[edge] (ssw.c:635).  That zeroing uses
// edge = min(end + 1, width - 1) with `end` in ABSOLUTE coordinates; it coincides with "the slot right of
// the band" except for rows  refLen - band <= i <= band + 1  (band wider than half the reference and the
// row already clipped at the last column), where it wipes the valid cell (i-1, refLen-1) instead: the
// `kill` flag below reproduces that.  Cells outside the band are "not written" for the trace back, which
// then fails (flag 1) as in the reference (there: a zero byte from the freshly allocated direction matrix).
//
// Direction nibble per cell, first cell of a row (slot dlo = max(band - i, 0)) in the TOP nibble of the row's
// first word: bit 3 = H did not come from the diagonal, bit 2 = it came from E rather than F, bit 1 = E opened
// (code 3 instead of 2), bit 0 = F opened (code 5 instead of 4) -- the reference's direction_line bytes, 8
// cells per word.  Each bit is the sign of a difference, shifted in with one funnel shift.
#pragma once
#include "core_sw.cuh"

namespace hrm {

// match masks of the reference sub-sequence: bit (j & 31) of word (j >> 5) of mask c is set iff ref[j] == c
// (c = 0..3; N never matches).  Words are strided so that the lanes of a warp hit different banks.
struct BandMasks {
    uint32_t* p;
    int stride; // distance between consecutive words of one thread's slice
    int MW;     // words per mask
    HRM_HD void clear() const
    {
        for (int t = 0; t < 4 * MW; t++) p[(int64_t)t * stride] = 0u;
    }
    HRM_HD void set(int c, int j) const
    {
        if (c >= 0 && c < 4) p[(int64_t)(c * MW + (j >> 5)) * stride] |= 1u << (j & 31);
    }
    // 32 match bits of base code c for columns [s, s + 32); columns outside [0, 32 * MW) read 0
    HRM_HD uint32_t window(int c, int s) const
    {
        if (c < 0 || c >= 4) return 0u;
        const int w = s >> 5, sh = s & 31; // arithmetic shift: floor
        const uint32_t lo = (w >= 0 && w < MW) ? p[(int64_t)(c * MW + w) * stride] : 0u;
        if (sh == 0) return lo;
        const uint32_t hi = (w + 1 >= 0 && w + 1 < MW) ? p[(int64_t)(c * MW + w + 1) * stride] : 0u;
        return (lo >> sh) | (hi << (32 - sh));
    }
};

// (H, E) pairs of one band row, signed 16 bit each in one word; slot stride as above
struct BandState {
    uint32_t* p;
    int stride;
    HRM_HD void zero(int nslots) const
    {
        for (int t = 0; t < nslots; t++) p[(int64_t)t * stride] = 0u;
    }
    HRM_HD uint32_t get(int d) const { return p[(int64_t)d * stride]; }
    HRM_HD void set(int d, int h, int e) const
    {
        p[(int64_t)d * stride] = ((uint32_t)h & 0xFFFFu) | ((uint32_t)e << 16);
    }
};

// acc' = (acc << 1) | sign(x)
HRM_HD uint32_t sw_shift_in_sign(uint32_t acc, int x)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)x, acc, 1);
#else
    return (acc << 1) | ((uint32_t)x >> 31);
#endif
}

// direction nibbles: word (row * nw + (c >> 3)) of the thread's slice, c = cell index within the row
struct BandDirs {
    uint32_t* p;
    int stride;
    int nw; // words per row
    HRM_HD uint32_t nibble(int row, int c) const
    {
        return (p[((int64_t)row * nw + (c >> 3)) * stride] >> (4 * (7 - (c & 7)))) & 15u;
    }
};

// One band iteration (ref: ssw.c:614-670).  qcode(i): code of read base i of the sub-sequence.  Returns the
// maximum H of the band.  Needs 2 * band + 3 state slots and (2 * band + 8) / 8 direction words per row.
//
// The row's only serial dependency is F.  With X = max(E, 0, diagonal + s) (independent of F) the reference's
//   f' = max(h_c - gapO, f - gapE),  h_c = max(E+, F+, diagonal + s) = max(f, X)
// collapses to f' = max(f - gapE, X - gapO) (f - gapO < f - gapE), and its direction flag
// (h_c - gapO > f - gapE) to X > f + gapO - gapE: one dependent operation per cell.  The (H, E) pair two
// slots ahead is loaded before the current slot is stored, so no load waits for a store.
template <class QCode>
HRM_HD int sw_band_iteration(const BandMasks& M, QCode qcode, int refLen, int readLen, int band, const BandState& st,
                             const BandDirs& dirs)
{
    const int W = 2 * band + 1;
    st.zero(W + 2);
    int mx = 0;
    const int64_t sst = st.stride, dst = dirs.stride;
    for (int i = 0; i < readLen; i++) {
        const int j0 = i - band; // column of slot 0
        const int dlo = j0 < 0 ? -j0 : 0;
        const int dhi = (refLen - 1 - j0) < (W - 1) ? (refLen - 1 - j0) : (W - 1);
        if (dhi < dlo) continue; // row beyond the band's reach (readLen > refLen + band): nothing is written
        const int qc = qcode(i);
        // ssw.c:635 (see the header): the up-neighbour of column refLen-1 is wiped; that slot is never read again
        if (i >= 1 && i <= band + 1 && i >= refLen - band) st.set(refLen - j0, 0, 0);
        int f = -HRM_SW_GAPE; // max(h_c[0] - gapO, 0 - gapE) with h_c[0] = 0
        int xdf = 0;          // sign = "F opened" of the current cell
        uint32_t mbits = 0u, acc = 0u;
        uint32_t* ps = st.p + dlo * sst;
        uint32_t* pd = dirs.p + (int64_t)i * dirs.nw * dst;
        uint32_t cur = ps[0], up = ps[sst];
        auto cell = [&]() {
            const uint32_t up_next = ps[2 * sst];
            const int diag = (int)(int16_t)(cur & 0xFFFFu);
            const int uh = (int)(int16_t)(up & 0xFFFFu), ue = (int)up >> 16;
            const int t1 = uh - HRM_SW_GAPO, t2 = ue - HRM_SW_GAPE;
            const int e = t1 > t2 ? t1 : t2;
            const int temp2 = (int)(mbits & 1u) * 4 + (diag - 2);
            mbits >>= 1;
            const int e1 = e > 0 ? e : 0;
            const int X = e1 > temp2 ? e1 : temp2;
            const int h = f > X ? f : X;
            const int f1 = f > 0 ? f : 0;
            const int temp1 = e1 > f1 ? e1 : f1;
            acc = sw_shift_in_sign(acc, temp2 - temp1); // H not from the diagonal (temp1 > temp2)
            acc = sw_shift_in_sign(acc, f1 - e1);       // ... from E (e1 > f1)
            acc = sw_shift_in_sign(acc, t2 - t1);       // E opened (t1 > t2)
            acc = sw_shift_in_sign(acc, xdf);           // F opened
            ps[0] = ((uint32_t)h & 0xFFFFu) | ((uint32_t)e << 16);
            mx = h > mx ? h : mx;
            xdf = f + (HRM_SW_GAPO - HRM_SW_GAPE) - X;
            const int fa = f - HRM_SW_GAPE, fb = X - HRM_SW_GAPO;
            f = fa > fb ? fa : fb;
            cur = up;
            up = up_next;
            ps += sst;
        };
        const int ncell = dhi - dlo + 1;
        for (int c = 0; c < ncell; c += 8) {
            if ((c & 31) == 0) mbits = M.window(qc, j0 + dlo + c);
            const int lim = ncell - c;
            if (lim >= 8) {
                cell(); cell(); cell(); cell(); cell(); cell(); cell(); cell();
            } else {
                for (int t = 0; t < lim; t++) cell();
                acc <<= 4 * (8 - lim);
            }
            *pd = acc;
            pd += dst;
        }
    }
    return mx;
}

// Trace back (ref: ssw.c:675-764) into 2-bit steps (0 M, 1 I, 2 D) in TRACE order (end of the alignment
// first); the reference's closing "M" is appended as the last step.  Returns the number of steps or -1.
HRM_HD int sw_band_traceback(const BandDirs& dirs, int band, int refLen, int readLen, uint32_t* steps, int max_steps)
{
    int i = readLen - 1, j = refLen - 1, state = 2, n = 0;
    while (i >= 0 && j > 0) {
        const int d = j - i + band;
        if (d < 0 || d > 2 * band) return -1; // outside the band: never written
        const uint32_t nb = dirs.nibble(i, d - (band > i ? band - i : 0));
        const int ce = (nb & 2u) ? 3 : 2, cf = (nb & 1u) ? 5 : 4;
        int code;
        if (state == 2) code = !(nb & 8u) ? 1 : ((nb & 4u) ? ce : cf);
        else code = state == 0 ? ce : cf;
        uint32_t op;
        switch (code) {
        case 1: --i; --j; state = 2; op = 0u; break;
        case 2: --i; state = 0; op = 1u; break;
        case 3: --i; state = 2; op = 1u; break;
        case 4: --j; state = 1; op = 2u; break;
        default: --j; state = 2; op = 2u; break;
        }
        if (n >= max_steps - 1) return -1;
        if ((n & 15) == 0) steps[n >> 4] = 0u;
        steps[n >> 4] |= op << (2 * (n & 15));
        n++;
    }
    if ((n & 15) == 0) steps[n >> 4] = 0u;
    n++; // closing M (ssw.c:736-753)
    return n;
}

HRM_HD uint32_t sw_step_at(const uint32_t* steps, int t) { return (steps[t >> 4] >> (2 * (t & 15))) & 3u; }

// ConvertAlignment + CalculateNumberMismatch on the step list (same text as sw_emit_cigar on the merged op
// list: zero-length runs print nothing).  qcode / rcode: codes of the WHOLE read / window.
template <class QCode, class RCode>
HRM_HD void sw_emit_steps(QCode qcode, int qlen, RCode rcode, SwAlignment* al, const uint32_t* steps, int nsteps,
                          char* cigar, int cigar_cap)
{
    const int read_end1 = al->query_end, ref_begin1 = al->ref_begin, read_begin1 = al->query_begin;
    int pos = 0, mism = 0;
    if (read_begin1 > 0) pos = sw_append(cigar, pos, cigar_cap, read_begin1, 'S');
    int rp = ref_begin1, qp = read_begin1;
    bool in_M = false, in_X = false;
    int length_M = 0, length_X = 0;
    int t = nsteps - 1; // forward order = reverse trace order
    while (t >= 0) {
        const uint32_t op = sw_step_at(steps, t);
        if (op == 0u) {
            if (rcode(rp) != qcode(qp)) {
                ++mism;
                if (in_M) pos = sw_append(cigar, pos, cigar_cap, length_M, '=');
                length_M = 0;
                ++length_X;
                in_M = false;
                in_X = true;
            } else {
                if (in_X) pos = sw_append(cigar, pos, cigar_cap, length_X, 'X');
                ++length_M;
                length_X = 0;
                in_M = true;
                in_X = false;
            }
            ++rp;
            ++qp;
            --t;
        } else {
            int length = 0;
            while (t >= 0 && sw_step_at(steps, t) == op) {
                ++length;
                --t;
            }
            if (op == 1u) qp += length;
            else rp += length;
            mism += length;
            if (in_M) pos = sw_append(cigar, pos, cigar_cap, length_M, '=');
            else if (in_X) pos = sw_append(cigar, pos, cigar_cap, length_X, 'X');
            in_M = in_X = false;
            length_M = length_X = 0;
            pos = sw_append(cigar, pos, cigar_cap, length, op == 1u ? 'I' : 'D');
        }
    }
    if (in_M) pos = sw_append(cigar, pos, cigar_cap, length_M, '=');
    else if (in_X) pos = sw_append(cigar, pos, cigar_cap, length_X, 'X');
    const int endS = qlen - read_end1 - 1;
    if (endS > 0) pos = sw_append(cigar, pos, cigar_cap, endS, 'S');
    al->mismatches = mism;
    al->cigar_len = pos;
}

// band class: all bands in [2^c, 2^(c+1)) -- one iteration of the doubling sequence falls into each class
HRM_HD int sw_band_class(int band)
{
    int c = 0;
    while ((2 << c) <= band) c++;
    return c;
}

} // namespace hrm
