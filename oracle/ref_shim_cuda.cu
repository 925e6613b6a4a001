// ref_shim_cuda.cu -- host-side calls into the reference's __host__ __device__ templates that
// live in .cu/.cuh files (so they need nvcc to parse): detail::computeWindowLocation
// (include/gpu/windowgenerationkernels.cuh:17-48) and the shifted-Hamming primitives
// shiftBitArrayLeftBy<1> / hammingdistanceHiLo (src/gpu/hammingdistancekernels.cu:44-118) plus
// the HiLo helpers of include/sequencehelpers.hpp.  Only host code is executed.
//
// TEST INFRASTRUCTURE ONLY (see ref_shim.cpp).  The reference sources are included from
// /root/reference where they lie; nothing is copied.  The loop in ref_shd() re-drives the
// reference primitives in the order of shiftedHammingDistanceWithFullOverlapKernelSmem1
// (hammingdistancekernels.cu:132-263) because a __global__ kernel cannot run on the host.
#include <cstdint>
#include <climits>
#include <vector>
#include <algorithm>

#include "../../reference/src/gpu/hammingdistancekernels.cu"
#include <cooperative_groups.h>
namespace cg = cooperative_groups; // the includer provides this alias in the reference (main_gpu.cu)
#include <gpu/windowgenerationkernels.cuh>

extern "C" {

void ref_window_location(int sectionBegin, int sectionEnd, int windowPos, int windowSize,
                         int extension, int* left, int* right, int* length, int* startpos)
{
    auto loc = detail::computeWindowLocation(sectionBegin, sectionEnd, windowPos, windowSize, extension);
    *left = loc.left;
    *right = loc.right;
    *length = loc.length;
    *startpos = loc.startpos;
}

void ref_shd(const uint32_t* anchor2bit, int anchorLength, const uint32_t* cand2bit,
             int candidateLength, float maxErrorRate, int* out_shift, int* out_score,
             int* out_orientation)
{
    auto identity = [](int i) { return i; };
    const int anchorints = SequenceHelpers::getEncodedNumInts2BitHiLo(anchorLength);
    const int candidateints = SequenceHelpers::getEncodedNumInts2BitHiLo(candidateLength);
    if (candidateLength <= anchorLength) {
        std::vector<unsigned int> anchorHiLo(std::max(anchorints, 2)), candHiLo(std::max(candidateints, 2));
        std::vector<unsigned int> mySharedAnchor(std::max(anchorints, 2));
        SequenceHelpers::convert2BitTo2BitHiLo(anchorHiLo.data(), anchor2bit, anchorLength);
        SequenceHelpers::convert2BitTo2BitHiLo(candHiLo.data(), cand2bit, candidateLength);
        unsigned int* const mySharedCandidate = candHiLo.data();
        unsigned int* const shiftptr_hi = mySharedAnchor.data();
        unsigned int* const shiftptr_lo = mySharedAnchor.data() + anchorints / 2;
        unsigned int* const otherptr_hi = mySharedCandidate;
        unsigned int* const otherptr_lo = mySharedCandidate + candidateints / 2;
        int bestScore = std::numeric_limits<int>::max();
        int bestShift = -1;
        int bestOrientation = -1;
        for (int orientation = 0; orientation < 2; orientation++) {
            const bool isReverseComplement = orientation == 1;
            if (isReverseComplement) {
                SequenceHelpers::reverseComplementSequenceInplace2BitHiLo(mySharedCandidate, candidateLength, identity);
            }
            for (int i = 0; i < anchorints; i++) mySharedAnchor[i] = anchorHiLo[i];
            for (int shift = 0; shift < anchorLength - candidateLength + 1; shift += 1) {
                const int max_errors = std::min(int(float(candidateLength) * maxErrorRate), std::max(0, bestScore - 1));
                if (shift != 0) {
                    shiftBitArrayLeftBy<1>(shiftptr_hi, anchorints / 2, identity);
                    shiftBitArrayLeftBy<1>(shiftptr_lo, anchorints / 2, identity);
                }
                const int score = hammingdistanceHiLo(shiftptr_hi, shiftptr_lo, otherptr_hi, otherptr_lo,
                                                      candidateLength, candidateLength, max_errors,
                                                      identity, identity,
                                                      [](auto i) { return __builtin_popcount(i); });
                if (score < bestScore) {
                    bestScore = score;
                    bestShift = shift;
                    bestOrientation = orientation;
                }
            }
        }
        *out_shift = bestShift;
        *out_score = bestScore;
        if (bestScore > int(float(candidateLength) * maxErrorRate)) *out_orientation = int(AlignmentOrientation::None);
        else *out_orientation = bestOrientation == 0 ? int(AlignmentOrientation::Forward) : int(AlignmentOrientation::ReverseComplement);
    } else {
        *out_shift = 0;
        *out_score = candidateLength;
        *out_orientation = int(AlignmentOrientation::None);
    }
}

} // extern "C"
