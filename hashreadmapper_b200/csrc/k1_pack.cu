// k1_pack.cu -- K1: fused 3N conversion + 2-bit packing (HBM-bound streaming kernel).
// ref: encodeSequencesTo2BitKernel<8> src/gpu/sequenceconversionkernels.cu:448-517 (call site
//      src/gpu/main_gpu.cu:522), SequenceHelpers::encodeSequence2Bit include/sequencehelpers.hpp:185-218.
// One thread produces one output word from one 128-bit load (16 ASCII bases): consecutive threads
// read consecutive 16-byte groups, so a warp moves 512 B in and 128 B out per step, fully coalesced.
// Algorithmic traffic: 1 B in + 0.25 B out per base.
#include "runtime.cuh"
#include "core_pack.cuh"

namespace hrm {

__global__ void __launch_bounds__(256) pack_rows_kernel(const char* __restrict__ ascii, int64_t ascii_pitch,
                                                        const int32_t* __restrict__ lengths, int64_t n, int conv,
                                                        uint32_t* __restrict__ out, int64_t pitch_words)
{
    const int64_t total = n * pitch_words;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t row = t / pitch_words;
        const int wi = (int)(t - row * pitch_words);
        const int len = lengths[row];
        const int valid = len - 16 * wi;
        uint32_t w = 0;
        if (valid > 0) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(ascii + row * ascii_pitch + 16 * (int64_t)wi));
            w = pack16(v.x, v.y, v.z, v.w, valid < 16 ? valid : 16, conv);
        }
        out[t] = w;
    }
}

// one long sequence; handles an unaligned base pointer by shifting to the aligned grid
__global__ void __launch_bounds__(256) pack_contiguous_kernel(const char* __restrict__ ascii, int64_t len, int conv,
                                                              uint32_t* __restrict__ out)
{
    const int64_t nwords = (len + 15) / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(ascii) & 15) == 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nwords; t += stride) {
        const int64_t rem = len - 16 * t;
        const int valid = rem < 16 ? (int)rem : 16;
        uint32_t w;
        if (aligned && valid == 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(ascii + 16 * t));
            w = pack16(v.x, v.y, v.z, v.w, 16, conv);
        } else {
            w = pack16_bytes(ascii + 16 * t, valid, conv);
        }
        out[t] = w;
    }
}

static inline unsigned grid_for(int64_t items, int block)
{
    int64_t g = HRM_SDIV(items, (int64_t)block);
    const int64_t cap = (int64_t)num_sms() * 16; // grid-stride beyond 16 resident-size waves
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_encode_2bit(const char* d_ascii, int64_t ascii_pitch, const int32_t* d_lengths, int64_t n,
                                      int conversion, uint32_t* d_out, int64_t out_pitch_words, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && out_pitch_words > 0, "sizes");
    HRM_REQUIRE(ascii_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(d_ascii) & 15) == 0,
                "ascii rows must be 16-byte aligned (pitch % 16 == 0)");
    HRM_REQUIRE(conversion >= 0 && conversion <= 2, "conversion");
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(pack_rows_kernel, grid_for(n * out_pitch_words, 256), 256, 0, as_stream(stream), d_ascii, ascii_pitch,
               d_lengths, n, conversion, d_out, out_pitch_words);
    return HRM_OK;
}

extern "C" hrm_status hrm_encode_2bit_contiguous(const char* d_ascii, int64_t len, int conversion, uint32_t* d_out,
                                                 hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(len >= 0, "len");
    HRM_REQUIRE(conversion >= 0 && conversion <= 2, "conversion");
    if (len == 0) return HRM_OK;
    HRM_LAUNCH(pack_contiguous_kernel, grid_for((len + 15) / 16, 256), 256, 0, as_stream(stream), d_ascii, len,
               conversion, d_out);
    return HRM_OK;
}
