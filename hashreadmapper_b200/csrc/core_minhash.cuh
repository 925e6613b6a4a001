// core_minhash.cuh -- K2 arithmetic: canonical k-mer at a position of a packed sequence and the
// murmur64 min-hash update.
// ref: forEachEncodedCanonicalKmerFromEncodedSequence include/sequencehelpers.hpp:847-926,
//      getEncodedKmerFromEncodedSequence :704-726, EncodedReverseComplement2Bit :14-47,
//      minhashSignatures3264Kernel include/gpu/gpusequencehasher.cuh:116-169.
#pragma once
#include "hrm_common.cuh"

namespace hrm {

// top 64 bits (32 bases) of the packed stream starting at base b; words >= nwords read as 0
HRM_HD uint64_t stream32(const uint32_t* w, int64_t nwords, int64_t b)
{
    const int64_t i = b >> 4;
    const int s = (int)(b & 15) * 2;
    const uint32_t w0 = i < nwords ? w[i] : 0u;
    const uint32_t w1 = i + 1 < nwords ? w[i + 1] : 0u;
    const uint32_t w2 = (s != 0 && i + 2 < nwords) ? w[i + 2] : 0u;
    return ((uint64_t)funnel_l(w0, w1, s) << 32) | funnel_l(w1, w2, s);
}

// reverse complement of a k-mer held in the low 2k bits
HRM_HD uint64_t revcomp_kmer(uint64_t kmer, int k)
{
    uint64_t x = brev64(kmer); // bit reversal also swaps the two bits of each base: swap them back
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    return (~x) >> (64 - 2 * k);
}

// canonical k-mer starting at base b (caller guarantees b + k <= sequence end)
HRM_HD uint64_t canonical_kmer(const uint32_t* w, int64_t nwords, int64_t b, int k)
{
    const uint64_t kmer = stream32(w, nwords, b) >> (64 - 2 * k);
    const uint64_t rc = revcomp_kmer(kmer, k);
    return kmer < rc ? kmer : rc;
}

HRM_HD uint64_t kmer_mask(int k) { return 0xFFFFFFFFFFFFFFFFULL >> ((32 - k) * 2); }

// Sequential signature of one sequence of `len` bases starting at base `start` of the packed
// stream w: minv[j] = min over positions of murmur64(canon + j).  Used by the host harness and by
// the generic (any H) device path; the warp kernel splits the position loop over lanes.
HRM_HD void minhash_sequential(const uint32_t* w, int64_t nwords, int64_t start, int len, int k, int H,
                               uint64_t* minv)
{
    for (int j = 0; j < H; j++) minv[j] = 0xFFFFFFFFFFFFFFFFULL;
    for (int p = 0; p + k <= len; p++) {
        const uint64_t c = canonical_kmer(w, nwords, start + p, k);
        for (int j = 0; j < H; j++) {
            const uint64_t h = murmur64(c + (uint64_t)j);
            if (h < minv[j]) minv[j] = h;
        }
    }
}

} // namespace hrm
