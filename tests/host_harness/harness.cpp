// tests/host_harness/harness.cpp -- TEST-ONLY.  Compiles the product's HRM_HD per-item arithmetic
// (hashreadmapper_b200/csrc/core_*.cuh) with g++ so that the "not gpu" test-suite can check it
// against the oracle on the CPU box.  It is never part of libhrm_b200.so and never used by the
// product; the shipped library runs this arithmetic on the device only.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../hashreadmapper_b200/csrc/core_pack.cuh"
#include "../../hashreadmapper_b200/csrc/core_minhash.cuh"
#include "../../hashreadmapper_b200/csrc/core_shd.cuh"
#include "../../hashreadmapper_b200/csrc/core_sw.cuh"

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

int hh_myers(const char* q, int qlen, const char* t, int tlen)
{
    return myers_nw((const unsigned char*)q, qlen, (const unsigned char*)t, tlen);
}

} // extern "C"
