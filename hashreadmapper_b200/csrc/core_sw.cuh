// core_sw.cuh -- V2 arithmetic: the reference's striped Smith-Waterman semantics in scalar form,
// one alignment per caller (thread or host harness), all scratch caller-provided.
// ref: ssw_align src/ssw.c:818-922, sw_sse2_byte :197-386, sw_sse2_word :412-588, banded_sw :590-774,
//      Aligner::Align src/ssw_cpp.cpp:361-400, ConvertAlignment :54-90, CalculateNumberMismatch :126-210,
//      default scoring ssw_cpp.cpp:230-242 (+2 / -2, N never matches, gap open 3, extend 1).
//
// What the striped SSE2 kernels compute, stated without SIMD (DESIGN.md "K7"):
//  * H = Gotoh local DP (zero floor) over the read padded to a multiple of 16 rows (byte mode) or
//    8 rows (word mode); pad rows score 0 against every reference base and count in the per-column
//    maxima that feed the second-best score.
//  * score1 = max H; ref_end1 = first column where the running max reaches it; read_end1 = smallest
//    row holding it in that column.  Byte mode is abandoned for word mode when max + 2 >= 255.
//  * second best = largest column maximum outside [ref_end1 - maskLen, ref_end1 + maskLen (+1 in
//    byte mode)), first column attaining it.
//  * begin = same pass over reversed prefixes, stopped at the first column whose maximum == score1.
//  * CIGAR = banded DP (band doubling) with the reference's exact direction codes and band-edge
//    zeroing, traced back from the bottom-right cell.
#pragma once
#include "hrm_common.cuh"

namespace hrm {

#define HRM_SW_GAPO 3
#define HRM_SW_GAPE 1

struct SwAlignment {
    int32_t sw_score, sw_score_next_best, ref_begin, ref_end, query_begin, query_end, ref_end_next_best,
        mismatches, flag, cigar_len;
};

// ref: kBaseTranslation src/ssw_cpp.cpp:12-25 (A/a 0, C/c 1, G/g 2, T/t 3, U/u 0, else 4)
HRM_HD int8_t sw_translate(unsigned char c)
{
    switch (c & 127) {
    case 'A': case 'a': case 'U': case 'u': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

HRM_HD int sw_score(int a, int b) { return (a == b && a < 4) ? 2 : -2; }

struct SwEnds {
    int score, ref, read, score2, ref2;
};

// One pass.  q: read codes (index qoff + dir*j, j = 0..readLen-1 so that a reversed prefix needs no
// copy), r: ref codes.  ref_dir 0: columns 0..refLen-1, 1: refLen-1..0.  padTo 16/8.
// H, E: rows+1 int16 scratch; maxColumn: refLen int16 scratch.
HRM_HD SwEnds sw_pass(const int8_t* r, int ref_dir, int refLen, const int8_t* q, int qoff, int qdir,
                      int readLen, int padTo, int terminate, int maskLen, int16_t* H, int16_t* E,
                      int16_t* maxColumn)
{
    const int rows = HRM_SDIV(readLen, padTo) * padTo;
    const bool byteMode = padTo == 16;
    for (int j = 0; j < rows; j++) {
        H[j] = 0;
        E[j] = 0;
    }
    int max = 0, end_ref = byteMode ? -1 : 0, end_read = 0;
    bool overflow = false;
    int begin = 0, end = refLen, step = 1;
    if (ref_dir == 1) {
        begin = refLen - 1;
        end = -1;
        step = -1;
    }
    for (int i = begin; i != end; i += step) {
        int F = 0, colmax = 0, argrow = 0, diag = 0;
        const int rc = r[i];
        for (int j = 0; j < rows; j++) {
            const int hprev = H[j];
            const int s = j < readLen ? sw_score(rc, q[qoff + qdir * j]) : 0;
            int h = diag + s;
            const int e = E[j];
            h = h > 0 ? h : 0;
            h = h > e ? h : e;
            h = h > F ? h : F;
            diag = hprev;
            H[j] = (int16_t)h;
            if (h > colmax) {
                colmax = h;
                argrow = j;
            }
            int ho = h - HRM_SW_GAPO;
            ho = ho > 0 ? ho : 0;
            int e2 = e - HRM_SW_GAPE;
            e2 = e2 > 0 ? e2 : 0;
            E[j] = (int16_t)(e2 > ho ? e2 : ho);
            int f2 = F - HRM_SW_GAPE;
            f2 = f2 > 0 ? f2 : 0;
            F = f2 > ho ? f2 : ho;
        }
        if (colmax > max) {
            max = colmax;
            if (byteMode && max + 2 >= 255) {
                overflow = true;
                break;
            }
            end_ref = i;
            end_read = argrow;
        }
        maxColumn[i] = (int16_t)colmax;
        if (colmax == terminate) break;
    }
    if (end_read > readLen - 1) end_read = readLen - 1;
    SwEnds b;
    b.score = overflow ? 255 : max;
    b.ref = end_ref;
    b.read = end_read;
    b.score2 = 0;
    b.ref2 = 0;
    if (!overflow && ref_dir == 0) { // the reverse pass' second best is never consumed
        int edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
        for (int i = 0; i < edge; i++)
            if (maxColumn[i] > b.score2) {
                b.score2 = maxColumn[i];
                b.ref2 = i;
            }
        edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
        for (int i = edge + (byteMode ? 1 : 0); i < refLen; i++)
            if (maxColumn[i] > b.score2) {
                b.score2 = maxColumn[i];
                b.ref2 = i;
            }
    }
    return b;
}

// direction byte: bit7 written, bits0-2 dh (1..5), bit3 de==3, bit4 df==5
HRM_HD int sw_dir_code(uint8_t cell, int state)
{
    if (!(cell & 0x80)) return 0;
    if (state == 2) return cell & 7;
    if (state == 0) return (cell & 8) ? 3 : 2;
    return (cell & 16) ? 5 : 4;
}


// banded_sw.  ref/read point at the sub-sequences.  hb/eb/hc: 2*len+16 ints each where
// len = max(refLen, readLen); dir: dir_cap bytes.  ops/lens: cigar out (op chars M/I/D), capacity
// maxops.  Returns number of ops, -1 if the trace back fails (reference: flag 1), -2 if scratch
// is too small (caller sizes it so that this cannot happen).
// direction-byte storage: row-major cells (row = read position, x = band-relative column)
struct DirLinear { // private contiguous slice
    uint8_t* p;
    HRM_HD uint8_t& at(int64_t idx) const { return p[idx]; }
};
struct DirInterleaved { // 32 lanes of a warp interleaved byte-wise: lanes in lock step coalesce
    uint8_t* p;
    int lane;
    HRM_HD uint8_t& at(int64_t idx) const { return p[idx * 32 + lane]; }
};

template <class DirT>
HRM_HD int sw_traceback(DirT dir, int row_stride, int width_d, int band_width, int refLen, int readLen, char* ops,
                        int32_t* lens, int maxops);

// one band iteration of banded_sw (ref: ssw.c:614-670): fills the direction bytes (row i at
// row_stride * i, width_d = 2*band_width+1 cells per row, cells outside the band zero = "not written")
// and returns the maximum H seen.  hb/eb/hc need width + 8 = 2*band_width + 11 ints.
template <class DirT>
HRM_HD int sw_banded_once(const int8_t* ref, const int8_t* read, int refLen, int readLen, int band_width,
                          int32_t* h_b, int32_t* e_b, int32_t* h_c, DirT dir, int row_stride)
{
    int i, j, f, temp1, temp2, max = 0;
    const int width = band_width * 2 + 3;
    const int width_d = band_width * 2 + 1;
    for (j = 0; j < width + 8; j++) {
        h_b[j] = 0;
        e_b[j] = 0;
        h_c[j] = 0;
    }
    for (i = 0; i < readLen; i++) {
        int beg = 0, end = refLen - 1, u = 0, edge;
        j = i - band_width;
        beg = beg > j ? beg : j;
        j = i + band_width;
        end = end < j ? end : j;
        edge = end + 1 < width - 1 ? end + 1 : width - 1;
        f = h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
        const int64_t line = (int64_t)row_stride * i;
        const int xi = (i - band_width) > 0 ? (i - band_width) : 0;
        const int xp = (i - 1 - band_width) > 0 ? (i - 1 - band_width) : 0;
        for (j = beg; j <= end; j++) {
            u = j - xi + 1;
            const int ue = j - xp + 1;     // (i-1, j)
            const int ub = j - 1 - xi + 1; // (i, j-1)
            const int ud = j - 1 - xp + 1; // (i-1, j-1)
            temp1 = i == 0 ? -HRM_SW_GAPO : h_b[ue] - HRM_SW_GAPO;
            temp2 = i == 0 ? -HRM_SW_GAPE : e_b[ue] - HRM_SW_GAPE;
            e_b[u] = temp1 > temp2 ? temp1 : temp2;
            const int de = temp1 > temp2 ? 3 : 2;
            temp1 = h_c[ub] - HRM_SW_GAPO;
            temp2 = f - HRM_SW_GAPE;
            f = temp1 > temp2 ? temp1 : temp2;
            const int df = temp1 > temp2 ? 5 : 4;
            const int e1 = e_b[u] > 0 ? e_b[u] : 0;
            const int f1 = f > 0 ? f : 0;
            temp1 = e1 > f1 ? e1 : f1;
            temp2 = h_b[ud] + sw_score(ref[j], read[i]);
            h_c[u] = temp1 > temp2 ? temp1 : temp2;
            if (h_c[u] > max) max = h_c[u];
            int dh;
            if (temp1 <= temp2) dh = 1;
            else dh = e1 > f1 ? de : df;
            dir.at(line + (j - xi)) = (uint8_t)(0x80 | dh | (de == 3 ? 8 : 0) | (df == 5 ? 16 : 0));
        }
        for (j = (end - xi + 1) > 0 ? (end - xi + 1) : 0; j < width_d; j++) dir.at(line + j) = 0; // not written
        for (j = 1; j <= u; j++) h_b[j] = h_c[j];
    }
    return max;
}

HRM_HD int sw_banded(const int8_t* ref, const int8_t* read, int refLen, int readLen, int score,
                     int band_width, int32_t* h_b, int32_t* e_b, int32_t* h_c, uint8_t* dir, int64_t dir_cap,
                     char* ops, int32_t* lens, int maxops)
{
    int max = 0;
    const int len = refLen > readLen ? refLen : readLen;
    int width_d;
    do {
        width_d = band_width * 2 + 1;
        if ((int64_t)width_d * readLen > dir_cap) return -2;
        const int m = sw_banded_once(ref, read, refLen, readLen, band_width, h_b, e_b, h_c, DirLinear{dir}, width_d);
        if (m > max) max = m; // ref: `max` persists across band doublings (ssw.c:600)
        band_width *= 2;
    } while (max < score && band_width <= len);
    band_width /= 2;
    return sw_traceback(DirLinear{dir}, width_d, width_d, band_width, refLen, readLen, ops, lens, maxops);
}

// trace back through the direction bytes of the last band iteration (ref: ssw.c:675-764).
// Returns the number of cigar ops (op chars M/I/D, in alignment order) or -1 on failure.
template <class DirT>
HRM_HD int sw_traceback(DirT dir, int row_stride, int width_d, int band_width, int refLen, int readLen, char* ops,
                        int32_t* lens, int maxops)
{
    int i = readLen - 1;
    int j = refLen - 1;
    int e = 0;
    int l = 0;
    char op = 'M', prev_op = 'M';
    int state = 2;
    while (i >= 0 && j > 0) {
        const int xi = (i - band_width) > 0 ? (i - band_width) : 0;
        const int x = j - xi;
        int code = 0;
        if (x >= 0 && x < width_d) code = sw_dir_code(dir.at((int64_t)row_stride * i + x), state);
        switch (code) {
        case 1: --i; --j; state = 2; op = 'M'; break;
        case 2: --i; state = 0; op = 'I'; break;
        case 3: --i; state = 2; op = 'I'; break;
        case 4: --j; state = 1; op = 'D'; break;
        case 5: --j; state = 2; op = 'D'; break;
        default: return -1;
        }
        if (op == prev_op) ++e;
        else {
            ++l;
            if (l - 1 < maxops) {
                ops[l - 1] = prev_op;
                lens[l - 1] = e;
            }
            prev_op = op;
            e = 1;
        }
    }
    if (op == 'M') {
        ++l;
        if (l - 1 < maxops) {
            ops[l - 1] = op;
            lens[l - 1] = e + 1;
        }
    } else {
        l += 2;
        if (l - 1 < maxops) {
            ops[l - 2] = op;
            lens[l - 2] = e;
            ops[l - 1] = 'M';
            lens[l - 1] = 1;
        }
    }
    if (l > maxops) l = maxops;
    for (int s = 0, t = l - 1; s < t; s++, t--) {
        const char co = ops[s];
        ops[s] = ops[t];
        ops[t] = co;
        const int cl = lens[s];
        lens[s] = lens[t];
        lens[t] = cl;
    }
    return l;
}

// decimal + op appended to a bounded char buffer; returns the new (unbounded) position
HRM_HD int sw_append(char* s, int pos, int cap, int len, char op)
{
    char buf[12];
    int n = 0;
    unsigned v = (unsigned)len;
    do {
        buf[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    for (int t = n - 1; t >= 0; t--, pos++)
        if (pos < cap) s[pos] = buf[t];
    if (pos < cap) s[pos] = op;
    return pos + 1;
}

struct SwScratch {
    int16_t* H;         // maxRows + 16
    int16_t* E;         // maxRows + 16
    int16_t* maxColumn; // maxRef
    int32_t* hb;        // 2*maxLen + 16
    int32_t* eb;
    int32_t* hc;
    uint8_t* dir;       // dir_cap
    int64_t dir_cap;
    char* ops;          // maxops
    int32_t* lens;      // maxops
    int maxops;
};

HRM_HD void sw_finish(const int8_t* q, int qlen, const int8_t* r, const SwScratch& S, SwAlignment* al, char* cigar,
                      int cigar_cap);
HRM_HD void sw_emit_cigar(const int8_t* q, int qlen, const int8_t* r, SwAlignment* al, const char* ops,
                          const int32_t* lens, int nops, char* cigar, int cigar_cap);

// Whole Align().  q/r: translated codes.  cigar: cap bytes (no NUL needed).
HRM_HD void sw_align(const int8_t* q, int qlen, const int8_t* r, int rlen, int maskLen, const SwScratch& S,
                     SwAlignment* al, char* cigar, int cigar_cap)
{
    al->sw_score = al->sw_score_next_best = al->ref_begin = al->ref_end = al->query_begin = al->query_end = 0;
    al->ref_end_next_best = al->mismatches = al->flag = al->cigar_len = 0;
    if (qlen <= 0) return; // ref: Align() returns false for an empty query
    bool word = false;
    SwEnds b = sw_pass(r, 0, rlen, q, 0, 1, qlen, 16, 255, maskLen, S.H, S.E, S.maxColumn);
    if (b.score == 255) {
        b = sw_pass(r, 0, rlen, q, 0, 1, qlen, 8, 65535, maskLen, S.H, S.E, S.maxColumn);
        word = true;
    }
    const int score1 = b.score, ref_end1 = b.ref, read_end1 = b.read;
    al->sw_score = score1;
    al->sw_score_next_best = maskLen >= 15 ? b.score2 : 0;
    al->ref_end = ref_end1;
    al->query_end = read_end1;
    al->ref_end_next_best = maskLen >= 15 ? b.ref2 : -1;
    if (score1 == 0 || ref_end1 < 0) { // undefined in the reference (ssw.c:220); deterministic here
        al->ref_begin = -1;
        al->query_begin = -1;
        return;
    }
    // reverse pass: read[read_end1 .. 0] against ref[ref_end1 .. 0]
    const SwEnds br = sw_pass(r, 1, ref_end1 + 1, q, read_end1, -1, read_end1 + 1, word ? 8 : 16,
                              word ? score1 : (score1 & 255), maskLen, S.H, S.E, S.maxColumn);
    al->ref_begin = br.ref;
    al->query_begin = read_end1 - br.read;
    al->flag = score1 > br.score ? 2 : 0;
    sw_finish(q, qlen, r, S, al, cigar, cigar_cap);
}

// Second half of Align(): banded trace back + ConvertAlignment + CalculateNumberMismatch, given the
// ends/begins found by the two passes (al->sw_score, ref_begin/end, query_begin/end, flag 0|2).
HRM_HD void sw_finish(const int8_t* q, int qlen, const int8_t* r, const SwScratch& S, SwAlignment* al, char* cigar,
                      int cigar_cap)
{
    const int score1 = al->sw_score, ref_end1 = al->ref_end, read_end1 = al->query_end;
    const int ref_begin1 = al->ref_begin, read_begin1 = al->query_begin;
    int flag = al->flag;
    const int refLen = ref_end1 - ref_begin1 + 1;
    const int readLen = read_end1 - read_begin1 + 1;
    int nops = 0;
    if (ref_begin1 < 0 || refLen <= 0 || readLen <= 0) {
        flag = 1;
    } else {
        int band = refLen - readLen;
        band = (band < 0 ? -band : band) + 1;
        nops = sw_banded(r + ref_begin1, q + read_begin1, refLen, readLen, score1, band, S.hb, S.eb, S.hc,
                         S.dir, S.dir_cap, S.ops, S.lens, S.maxops);
        if (nops < 0) {
            flag = 1;
            nops = 0;
        }
    }
    al->flag = flag;
    sw_emit_cigar(q, qlen, r, al, S.ops, S.lens, nops, cigar, cigar_cap);
}

// ConvertAlignment + CalculateNumberMismatch (ref: ssw_cpp.cpp:54-90, :126-210): S / = / X / I / D string
HRM_HD void sw_emit_cigar(const int8_t* q, int qlen, const int8_t* r, SwAlignment* al, const char* ops,
                          const int32_t* lens, int nops, char* cigar, int cigar_cap)
{
    const int read_end1 = al->query_end, ref_begin1 = al->ref_begin, read_begin1 = al->query_begin;
    int pos = 0, mism = 0;
    if (read_begin1 > 0) pos = sw_append(cigar, pos, cigar_cap, read_begin1, 'S');
    const int8_t* rp = r + ref_begin1;
    const int8_t* qp = q + read_begin1;
    bool in_M = false, in_X = false;
    int length_M = 0, length_X = 0;
    for (int c = 0; c < nops; c++) {
        const char op = ops[c];
        const int length = lens[c];
        if (op == 'M') {
            for (int t = 0; t < length; t++) {
                if (*rp != *qp) {
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
            }
        } else {
            if (op == 'I') qp += length;
            else rp += length;
            mism += length;
            if (in_M) pos = sw_append(cigar, pos, cigar_cap, length_M, '=');
            else if (in_X) pos = sw_append(cigar, pos, cigar_cap, length_X, 'X');
            in_M = in_X = false;
            length_M = length_X = 0;
            pos = sw_append(cigar, pos, cigar_cap, length, op);
        }
    }
    if (in_M) pos = sw_append(cigar, pos, cigar_cap, length_M, '=');
    else if (in_X) pos = sw_append(cigar, pos, cigar_cap, length_X, 'X');
    const int endS = qlen - read_end1 - 1;
    if (endS > 0) pos = sw_append(cigar, pos, cigar_cap, endS, 'S');
    al->mismatches = mism;
    al->cigar_len = pos;
}

// ---- V3: Myers bit-vector global edit distance (replaces edlibAlign NW/DISTANCE) ----------------
// ref: edlibAlign src/edlib.cpp:1474-1476 as called at src/gpu/mappinghandler.cu:968-987.
// Query up to 512 symbols in 64-bit blocks; Peq built from codes 0..4 (5 symbols + wildcard-free).
#define HRM_MYERS_MAX_BLOCKS 8
HRM_HD int myers_nw(const unsigned char* q, int qlen, const unsigned char* t, int tlen)
{
    if (qlen == 0) return tlen;
    if (tlen == 0) return qlen;
    const int nb = HRM_SDIV(qlen, 64);
    uint64_t Pv[HRM_MYERS_MAX_BLOCKS], Mv[HRM_MYERS_MAX_BLOCKS];
    uint64_t Peq[4][HRM_MYERS_MAX_BLOCKS]; // match masks of the four letters; other bytes on the fly
    for (int b = 0; b < nb; b++) {
        Pv[b] = ~0ULL;
        Mv[b] = 0;
        Peq[0][b] = Peq[1][b] = Peq[2][b] = Peq[3][b] = 0;
    }
    for (int z = 0; z < qlen; z++) {
        const unsigned char c = q[z];
        const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
        if (code >= 0) Peq[code][z >> 6] |= 1ULL << (z & 63);
    }
    const int lastbits = qlen - 64 * (nb - 1);
    const uint64_t topbit = 1ULL << (lastbits - 1);
    int score = qlen;
    for (int c = 0; c < tlen; c++) {
        const unsigned char tc = t[c];
        const int tcode = tc == 'A' ? 0 : tc == 'C' ? 1 : tc == 'G' ? 2 : tc == 'T' ? 3 : -1;
        int hin = 1; // D[0][c] - D[0][c-1] = +1 (global)
        for (int b = 0; b < nb; b++) {
            uint64_t Eq = 0;
            if (tcode >= 0) {
                Eq = Peq[tcode][b];
            } else {
                const int base = 64 * b;
                const int lim = (qlen - base) < 64 ? (qlen - base) : 64;
                for (int z = 0; z < lim; z++) Eq |= (uint64_t)(q[base + z] == tc) << z;
            }
            const uint64_t pv = Pv[b], mv = Mv[b];
            const uint64_t hinNeg = hin < 0 ? 1ULL : 0ULL;
            const uint64_t hinPos = hin > 0 ? 1ULL : 0ULL;
            const uint64_t Xv = Eq | mv;
            const uint64_t EqH = Eq | hinNeg;
            const uint64_t Xh = (((EqH & pv) + pv) ^ pv) | EqH;
            uint64_t Ph = mv | ~(Xh | pv);
            uint64_t Mh = pv & Xh;
            int hout = 0;
            const uint64_t hb = (b == nb - 1) ? topbit : (1ULL << 63);
            if (Ph & hb) hout = 1;
            else if (Mh & hb) hout = -1;
            Ph = (Ph << 1) | hinPos;
            Mh = (Mh << 1) | hinNeg;
            Pv[b] = Mh | ~(Xv | Ph);
            Mv[b] = Ph & Xv;
            hin = hout;
        }
        score += hin;
    }
    return score;
}

} // namespace hrm
