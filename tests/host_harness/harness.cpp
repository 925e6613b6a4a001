// tests/host_harness/harness.cpp -- TEST-ONLY.  Compiles the product's HRM_HD per-item arithmetic
// (hashreadmapper_b200/csrc/core_*.cuh) with g++ so that the "not gpu" test-suite can check it
// against the oracle on the CPU box.  It is never part of libhrm_b200.so and never used by the
// product; the shipped library runs this arithmetic on the device only.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../hashreadmapper_b200/csrc/core_pack.cuh"
#include "../../hashreadmapper_b200/csrc/core_minhash.cuh"
#include "../../hashreadmapper_b200/csrc/core_shd.cuh"
#include "../../hashreadmapper_b200/csrc/core_sw.cuh"
#include "../../hashreadmapper_b200/csrc/core_swpair.cuh"
#include "../../hashreadmapper_b200/csrc/core_swband.cuh"

using namespace hrm;

extern "C" {

void hh_encode_2bit(const char* ascii, int len, int conv, uint32_t* out, int aligned)
{
    const int nw = HRM_SDIV(len, 16);
    for (int i = 0; i < nw; i++) {
        const int valid = (len - 16 * i) < 16 ? (len - 16 * i) : 16;
        if (aligned) {
            uint32_t w[4] = {0, 0, 0, 0};
            std::memcpy(w, ascii + 16 * i, (size_t)valid); // bytes beyond `valid` may be anything
            if (valid < 16) std::memset((char*)w + valid, 'T', (size_t)(16 - valid)); // garbage on purpose
            out[i] = pack16(w[0], w[1], w[2], w[3], valid, conv);
        } else {
            out[i] = pack16_bytes(ascii + 16 * i, valid, conv);
        }
    }
}

void hh_minhash(const uint32_t* enc, int64_t nwords, int64_t start, int len, int k, int H, uint64_t* sig,
                uint8_t* valid)
{
    if (len >= k) {
        uint64_t minv[64];
        minhash_sequential(enc, nwords, start, len, k, H, minv);
        for (int j = 0; j < H; j++) {
            sig[j] = minv[j] & kmer_mask(k);
            valid[j] = 1;
        }
    } else {
        for (int j = 0; j < H; j++) {
            sig[j] = ~0ULL;
            valid[j] = 0;
        }
    }
}

void hh_shd(const uint32_t* anchor, int64_t anchor_words, int64_t anchor_base, int La, const uint32_t* read,
            int64_t read_words, int Lc, float rate, int* shift, int* score, int* orientation)
{
    shd_sequential(anchor, anchor_words, anchor_base, La, read, read_words, Lc, rate, shift, score, orientation);
}

void hh_window_location(int64_t chromLen, int64_t pos, int w, int ext, int* l, int* r, int* len)
{
    window_location(chromLen, pos, w, ext, l, r, len);
}

void hh_sw_align(const char* query, int qlen, const char* ref, int rlen, int maskLen, SwAlignment* al,
                 char* cigar, int cigar_cap)
{
    std::vector<int8_t> q(qlen + 1), r(rlen + 1);
    for (int i = 0; i < qlen; i++) q[i] = sw_translate((unsigned char)query[i]);
    for (int i = 0; i < rlen; i++) r[i] = sw_translate((unsigned char)ref[i]);
    const int maxLen = (qlen > rlen ? qlen : rlen) + 16;
    std::vector<int16_t> H(maxLen + 32), E(maxLen + 32), mc(maxLen + 32);
    std::vector<int32_t> hb(2 * maxLen + 32), eb(2 * maxLen + 32), hc(2 * maxLen + 32);
    std::vector<uint8_t> dir((size_t)(2 * maxLen + 1) * maxLen);
    std::vector<char> ops(2 * maxLen + 8);
    std::vector<int32_t> lens(2 * maxLen + 8);
    SwScratch S{H.data(), E.data(), mc.data(), hb.data(), eb.data(), hc.data(), dir.data(),
                (int64_t)dir.size(), ops.data(), lens.data(), (int)ops.size()};
    sw_align(q.data(), qlen, r.data(), rlen, maskLen, S, al, cigar, cigar_cap);
    if (al->cigar_len < cigar_cap) cigar[al->cigar_len] = 0;
}

// band statistics of the trace-back stage (used by tools/band_stats.py to size the kernels)
void hh_sw_band_stats(const char* query, int qlen, const char* ref, int rlen, int maskLen, int* out)
{
    SwAlignment al;
    char cig[8];
    hh_sw_align(query, qlen, ref, rlen, maskLen, &al, cig, 0);
    out[0] = al.sw_score;
    out[1] = al.ref_end - al.ref_begin + 1;
    out[2] = al.query_end - al.query_begin + 1;
    out[3] = out[4] = 0;
    if (al.sw_score <= 0 || al.ref_begin < 0) return;
    std::vector<int8_t> q(qlen + 1), r(rlen + 1);
    for (int i = 0; i < qlen; i++) q[i] = sw_translate((unsigned char)query[i]);
    for (int i = 0; i < rlen; i++) r[i] = sw_translate((unsigned char)ref[i]);
    const int refLen = out[1], readLen = out[2], len = refLen > readLen ? refLen : readLen;
    int band = refLen - readLen;
    band = (band < 0 ? -band : band) + 1;
    std::vector<int32_t> hb(2 * len + 32), eb(2 * len + 32), hc(2 * len + 32);
    std::vector<uint8_t> dir((size_t)(2 * len + 3) * (readLen + 1));
    int mx = 0, iters = 0;
    do {
        const int m = sw_banded_once(r.data() + al.ref_begin, q.data() + al.query_begin, refLen, readLen, band, hb.data(),
                                     eb.data(), hc.data(), DirLinear{dir.data()}, 2 * band + 1);
        mx = m > mx ? m : mx;
        band *= 2;
        iters++;
    } while (mx < al.sw_score && band <= len);
    out[3] = band / 2;
    out[4] = iters;
}

// the thread-per-alignment band ladder of k7_verify.cu (core_swband.cuh) on top of the scalar passes:
// same begin/end search as hh_sw_align, then one sw_band_iteration per band of the doubling sequence
void hh_sw_align_band(const char* query, int qlen, const char* ref, int rlen, int maskLen, SwAlignment* al,
                      char* cigar, int cigar_cap, int stride)
{
    std::vector<int8_t> q(qlen + 1), r(rlen + 1);
    for (int i = 0; i < qlen; i++) q[i] = sw_translate((unsigned char)query[i]);
    for (int i = 0; i < rlen; i++) r[i] = sw_translate((unsigned char)ref[i]);
    const int maxLen = (qlen > rlen ? qlen : rlen) + 16;
    std::vector<int16_t> H(maxLen + 32), E(maxLen + 32), mc(maxLen + 32);
    al->sw_score = al->sw_score_next_best = al->ref_begin = al->ref_end = al->query_begin = al->query_end = 0;
    al->ref_end_next_best = al->mismatches = al->flag = al->cigar_len = 0;
    if (cigar_cap > 0) cigar[0] = 0;
    if (qlen <= 0) return;
    bool word = false;
    SwEnds b = sw_pass(r.data(), 0, rlen, q.data(), 0, 1, qlen, 16, 255, maskLen, H.data(), E.data(), mc.data());
    if (b.score == 255) {
        b = sw_pass(r.data(), 0, rlen, q.data(), 0, 1, qlen, 8, 65535, maskLen, H.data(), E.data(), mc.data());
        word = true;
    }
    al->sw_score = b.score;
    al->sw_score_next_best = maskLen >= 15 ? b.score2 : 0;
    al->ref_end = b.ref;
    al->query_end = b.read;
    al->ref_end_next_best = maskLen >= 15 ? b.ref2 : -1;
    if (b.score == 0 || b.ref < 0) {
        al->ref_begin = -1;
        al->query_begin = -1;
        return;
    }
    const SwEnds br = sw_pass(r.data(), 1, b.ref + 1, q.data(), b.read, -1, b.read + 1, word ? 8 : 16,
                              word ? b.score : (b.score & 255), maskLen, H.data(), E.data(), mc.data());
    al->ref_begin = br.ref;
    al->query_begin = b.read - br.read;
    al->flag = b.score > br.score ? 2 : 0;
    const int refLen = al->ref_end - al->ref_begin + 1, readLen = al->query_end - al->query_begin + 1;
    const int len = refLen > readLen ? refLen : readLen;
    int band = refLen - readLen;
    band = (band < 0 ? -band : band) + 1;
    const int MW = HRM_SDIV(refLen, 32);
    std::vector<uint32_t> mk((size_t)4 * MW * stride + 1, 0xdeadbeefu);
    BandMasks M{mk.data(), stride, MW};
    M.clear();
    for (int j = 0; j < refLen; j++) M.set(r[al->ref_begin + j], j);
    auto qc = [&](int i) -> int { return q[al->query_begin + i]; };
    std::vector<uint32_t> steps((qlen + rlen) / 16 + 4, 0xdeadbeefu);
    int nsteps = -1;
    int prev_class = -1;
    while (true) {
        const int cls = sw_band_class(band);
        if (cls <= prev_class) std::abort(); // one iteration per class
        prev_class = cls;
        std::vector<uint32_t> st((size_t)(2 * band + 3) * stride + 1, 0xdeadbeefu);
        const int nw = (2 * band + 8) / 8;
        std::vector<uint32_t> dd((size_t)readLen * nw * stride + 1, 0xdeadbeefu);
        BandState S{st.data(), stride};
        BandDirs D{dd.data(), stride, nw};
        const int m = sw_band_iteration(M, qc, refLen, readLen, band, S, D);
        if (m < al->sw_score && band * 2 <= len) {
            band *= 2;
            continue;
        }
        nsteps = sw_band_traceback(D, band, refLen, readLen, steps.data(), (int)steps.size() * 16);
        break;
    }
    if (nsteps < 0) {
        al->flag = 1;
        nsteps = 0;
    }
    sw_emit_steps([&](int i) -> int { return q[i]; }, qlen, [&](int j) -> int { return r[j]; }, al, steps.data(), nsteps,
                  cigar, cigar_cap);
    if (al->cigar_len < cigar_cap) cigar[al->cigar_len] = 0;
}

// cell-by-cell comparison of sw_band_iteration (diagonal coordinates) with the literal sw_banded_once for an
// ARBITRARY band: returns the number of differing direction cells, +1000000 if the band maxima differ
int hh_band_compare(const char* ref, int refLen, const char* read, int readLen, int band, int stride)
{
    std::vector<int8_t> q(readLen + 1), r(refLen + 1);
    for (int i = 0; i < readLen; i++) q[i] = sw_translate((unsigned char)read[i]);
    for (int i = 0; i < refLen; i++) r[i] = sw_translate((unsigned char)ref[i]);
    const int width_d = 2 * band + 1;
    std::vector<int32_t> hb(2 * band + 32), eb(2 * band + 32), hc(2 * band + 32);
    std::vector<uint8_t> dir((size_t)width_d * readLen + 8);
    const int m1 = sw_banded_once(r.data(), q.data(), refLen, readLen, band, hb.data(), eb.data(), hc.data(),
                                  DirLinear{dir.data()}, width_d);
    const int MW = HRM_SDIV(refLen, 32);
    std::vector<uint32_t> mk((size_t)4 * MW * stride + 1, 0xdeadbeefu);
    BandMasks M{mk.data(), stride, MW};
    M.clear();
    for (int j = 0; j < refLen; j++) M.set(r[j], j);
    std::vector<uint32_t> st((size_t)(2 * band + 3) * stride + 1, 0xdeadbeefu);
    const int nw = (2 * band + 8) / 8;
    std::vector<uint32_t> dd((size_t)readLen * nw * stride + 1, 0xdeadbeefu);
    BandState S{st.data(), stride};
    BandDirs D{dd.data(), stride, nw};
    const int m2 = sw_band_iteration(M, [&](int i) -> int { return q[i]; }, refLen, readLen, band, S, D);
    int bad = m1 != m2 ? 1000000 : 0;
    for (int i = 0; i < readLen; i++) {
        const int xi = (i - band) > 0 ? (i - band) : 0;
        const int beg = xi, end = (i + band) < (refLen - 1) ? (i + band) : (refLen - 1);
        for (int j = beg; j <= end; j++) {
            const uint8_t cell = dir[(size_t)width_d * i + (j - xi)];
            const uint32_t nb = D.nibble(i, j - xi); // cell index within the row
            const int dh = cell & 7;
            // bit 3: not diagonal, bit 2: from E (only meaningful when bit 3), bit 1: E opened, bit 0: F opened
            const uint32_t want_hi = dh == 1 ? 0u : ((dh == 2 || dh == 3) ? 12u : 8u);
            const uint32_t want = want_hi | ((cell & 8) ? 2u : 0u) | ((cell & 16) ? 1u : 0u);
            const uint32_t got = (nb & 8u) ? nb : (nb & 11u);
            if (!(cell & 0x80) || got != want) bad++;
        }
    }
    return bad;
}

int hh_myers(const char* q, int qlen, const char* t, int tlen)
{
    return myers_nw((const unsigned char*)q, qlen, (const unsigned char*)t, tlen);
}

} // extern "C"

// ---- the paired SIMD passes (core_swpair.cuh), G lanes emulated one after the other -----------------
namespace {
struct PairOut {
    int score, score2, ref_begin, ref_end, query_begin, query_end, ref2, flag;
};

template <int G, int R>
void pair_passes_host(const int8_t* A, const int8_t* B, int L, const int8_t* ref, int rl, int ml, PairOut out[2])
{
    const int rtot = G * R, top = pair_top(rtot, L);
    std::vector<PairWave<R>> w(G);
    auto fwd_code = [&](int half, int g) {
        const int j = g - top;
        if (j < 0) return PAIR_CODE_EMPTY;
        if (j >= L) return PAIR_CODE_PAD;
        return (int)(half ? B[j] : A[j]);
    };
    for (int l = 0; l < G; l++) {
        pair_wave_reset(w[l]);
        pair_build(w[l].L, l * R, fwd_code);
    }
    std::vector<uint32_t> cmW(rl + 1), cmB(rl + 1);
    auto run = [&](int ncols, auto colinfo, auto on_last) {
        for (int t = 0; t < ncols + G - 1; t++) {
            uint32_t oS[G], oF[G], oC[G];
            for (int l = 0; l < G; l++) {
                oS[l] = w[l].outS;
                oF[l] = w[l].outF;
                oC[l] = w[l].outCm;
            }
            bool stop = false;
            for (int l = 0; l < G; l++) {
                const int c = t - l;
                int rc = 0;
                uint32_t hmask = 0;
                if (c >= 0 && c < ncols) colinfo(c, rc, hmask);
                uint32_t a = 0, b = 0;
                const bool did = pair_wave_step(w[l], c, ncols, rc, hmask, l ? oS[l - 1] : 0u, l ? oF[l - 1] : PX_TWOS,
                                                l ? oC[l - 1] : 0u, (uint32_t)c, a, b);
                if (did && l == G - 1) stop = on_last(c, a, b);
            }
            if (stop) break;
        }
    };
    run(rl, [&](int c, int& rc, uint32_t& hm) { rc = ref[c]; hm = PX_ONES; },
        [&](int c, uint32_t a, uint32_t b) { cmW[c] = a; cmB[c] = b; return false; });
    int score1[2], ref_end1[2], read_end1[2];
    const bool samePad = pair_pad8(L) == pair_pad16(L);
    for (int h = 0; h < 2; h++) {
        uint32_t word = 0;
        for (int l = 0; l < G; l++) {
            const uint32_t x = pair_best_word(h ? w[l].bestB : w[l].bestA, l * R);
            word = x > word ? x : word;
        }
        score1[h] = (int)(word >> 20);
        ref_end1[h] = word ? 1023 - (int)((word >> 10) & 1023u) : -1;
        int g = 1023 - (int)(word & 1023u);
        read_end1[h] = g - top < L - 1 ? g - top : L - 1;
        PairOut& o = out[h];
        o = PairOut{score1[h], 0, -1, ref_end1[h], -1, word ? read_end1[h] : 0, 0, 0};
        o.ref2 = ml >= 15 ? 0 : -1;
        if (!word) continue; // score 0: ref_end -1 (byte kernel's initial end), query_end 0, begins -1
        const bool word_mode = score1[h] + 2 >= 255;
        const std::vector<uint32_t>& cm = (word_mode || samePad) ? cmW : cmB;
        auto val = [&](int c) { return (int)((cm[c] >> (16 * h)) & 0xFFFFu); };
        int s2 = 0, r2 = 0;
        const int e1 = ref_end1[h] - ml > 0 ? ref_end1[h] - ml : 0;
        const int e2 = ref_end1[h] + ml > rl ? rl : ref_end1[h] + ml;
        for (int c = 0; c < e1; c++)
            if (val(c) > s2) {
                s2 = val(c);
                r2 = c;
            }
        for (int c = e2 + (word_mode ? 0 : 1); c < rl; c++)
            if (val(c) > s2) {
                s2 = val(c);
                r2 = c;
            }
        o.score2 = ml >= 15 ? s2 : 0;
        o.ref2 = ml >= 15 ? r2 : -1;
    }
    // reverse pass: both halves over the shared columns cmax .. 0
    const bool vA = score1[0] > 0, vB = score1[1] > 0;
    if (!vA && !vB) return;
    const int cmax = std::max(vA ? ref_end1[0] : -1, vB ? ref_end1[1] : -1);
    auto rev_code = [&](int half, int g) {
        if (!(half ? vB : vA)) return PAIR_CODE_EMPTY;
        const int nr = read_end1[half] + 1;
        if (g >= nr) return PAIR_CODE_EMPTY;
        return (int)(half ? B[read_end1[1] - g] : A[read_end1[0] - g]);
    };
    for (int l = 0; l < G; l++) {
        pair_wave_reset(w[l]);
        pair_build(w[l].L, l * R, rev_code);
    }
    bool doneA = !vA, doneB = !vB;
    run(cmax + 1,
        [&](int j, int& rc, uint32_t& hm) {
            const int c = cmax - j;
            rc = ref[c];
            hm = ((vA && c <= ref_end1[0]) ? 1u : 0u) | ((vB && c <= ref_end1[1]) ? 0x10000u : 0u);
        },
        [&](int, uint32_t, uint32_t b) {
            if ((int)(b & 0xFFFFu) == score1[0]) doneA = true;
            if ((int)(b >> 16) == score1[1]) doneB = true;
            return doneA && doneB;
        });
    for (int h = 0; h < 2; h++) {
        if (!(h ? vB : vA)) continue;
        uint32_t word = 0;
        for (int l = 0; l < G; l++) {
            const uint32_t x = pair_best_word(h ? w[l].bestB : w[l].bestA, l * R);
            word = x > word ? x : word;
        }
        const int maxv = (int)(word >> 20);
        const int j = 1023 - (int)((word >> 10) & 1023u);
        int g = 1023 - (int)(word & 1023u);
        const int nr = read_end1[h] + 1;
        if (g > nr - 1) g = nr - 1;
        out[h].ref_begin = cmax - j;
        out[h].query_begin = read_end1[h] - g;
        out[h].flag = score1[h] > maxv ? 2 : 0;
    }
}
} // namespace

extern "C" int hh_sw_pair(const int8_t* A, const int8_t* B, int L, const int8_t* ref, int rl, int ml, int G, int R,
                          int* out16)
{
    PairOut o[2];
    if (!pair_fits(G * R, L, rl)) return -1;
    if (G == 4 && R == 40) pair_passes_host<4, 40>(A, B, L, ref, rl, ml, o);
    else if (G == 8 && R == 32) pair_passes_host<8, 32>(A, B, L, ref, rl, ml, o);
    else if (G == 8 && R == 34) pair_passes_host<8, 34>(A, B, L, ref, rl, ml, o);
    else if (G == 4 && R == 32) pair_passes_host<4, 32>(A, B, L, ref, rl, ml, o);
    else if (G == 2 && R == 33) pair_passes_host<2, 33>(A, B, L, ref, rl, ml, o);
    else return -2;
    for (int h = 0; h < 2; h++) {
        int* p = out16 + 8 * h;
        p[0] = o[h].score;
        p[1] = o[h].score2;
        p[2] = o[h].ref_begin;
        p[3] = o[h].ref_end;
        p[4] = o[h].query_begin;
        p[5] = o[h].query_end;
        p[6] = o[h].ref2;
        p[7] = o[h].flag;
    }
    return 0;
}
