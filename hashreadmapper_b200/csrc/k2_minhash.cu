// k2_minhash.cu -- K2: k-mer minhash sketching, one warp per sequence (integer-ALU bound).
// ref: minhashSignatures3264Kernel include/gpu/gpusequencehasher.cuh:116-169 (one thread per
//      (sequence, hash function), each re-walking the sequence), launcher :720-776.
// Here the 32 lanes of a warp split the k-mer positions of one sequence; every lane evaluates all
// hash functions of a chunk on its canonical k-mers (kept in registers) and the per-function minima
// are combined with warp shuffles, so a sequence is walked once per chunk instead of once per
// hash function and its packed words are read through L1 only.
// Work per sequence: (len-k+1) * H murmur64 evaluations; traffic: ceil(len/16)*4 B in, 8H(+H) B out.
#include "runtime.cuh"
#include "core_minhash.cuh"

namespace hrm {

struct SeqSource {
    const uint32_t* words;   // rows: first row; windows: the packed chromosome
    int64_t pitch_words;     // rows: words per row; windows: total words of the chromosome
    const int32_t* lengths;  // rows only
    int64_t chrom_len;       // windows only
    int64_t first_window;    // windows only
    int32_t stride;          // windows: w - k + 1 ; rows: 0
    int32_t window_size;     // windows only
};

__device__ __forceinline__ void seq_locate(const SeqSource& src, int64_t i, const uint32_t*& w, int64_t& nwords,
                                           int64_t& start, int& len)
{
    if (src.stride == 0) {
        w = src.words + i * src.pitch_words;
        nwords = src.pitch_words;
        start = 0;
        len = src.lengths[i];
    } else {
        w = src.words;
        nwords = src.pitch_words;
        start = (src.first_window + i) * (int64_t)src.stride;
        const int64_t rem = src.chrom_len - start;
        len = rem < src.window_size ? (int)(rem > 0 ? rem : 0) : src.window_size;
    }
}

constexpr int MH_CHUNK = 8;

__global__ void __launch_bounds__(256) minhash_warp_kernel(SeqSource src, int64_t n, int k, int H,
                                                           uint64_t* __restrict__ sigs, uint8_t* __restrict__ valid)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t mask = kmer_mask(k);
    for (int64_t i = warp0; i < n; i += nwarps) {
        const uint32_t* w;
        int64_t nwords, start;
        int len;
        seq_locate(src, i, w, nwords, start, len);
        if (len < k) {
            for (int j = lane; j < H; j += 32) {
                sigs[i * H + j] = ~0ULL;
                if (valid) valid[i * H + j] = 0;
            }
            continue;
        }
        for (int c = 0; c < H; c += MH_CHUNK) {
            uint64_t minv[MH_CHUNK];
#pragma unroll
            for (int j = 0; j < MH_CHUNK; j++) minv[j] = ~0ULL;
            for (int p = lane; p + k <= len; p += 32) {
                const uint64_t canon = canonical_kmer(w, nwords, start + p, k);
#pragma unroll
                for (int j = 0; j < MH_CHUNK; j++) {
                    const uint64_t h = murmur64(canon + (uint64_t)(c + j));
                    minv[j] = h < minv[j] ? h : minv[j];
                }
            }
#pragma unroll
            for (int j = 0; j < MH_CHUNK; j++) {
                uint64_t v = minv[j];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v, d);
                    v = o < v ? o : v;
                }
                minv[j] = v;
            }
            // lane j writes function c+j (all lanes hold every minimum after the xor reduction)
#pragma unroll
            for (int j = 0; j < MH_CHUNK; j++) {
                if (lane == j && c + j < H) {
                    sigs[i * H + c + j] = minv[j] & mask;
                    if (valid) valid[i * H + c + j] = 1;
                }
            }
        }
    }
}

static hrm_status launch_minhash(const SeqSource& src, int64_t n, int k, int H, uint64_t* d_sigs, uint8_t* d_valid,
                                 cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    int64_t blocks = HRM_SDIV(n, (int64_t)8);
    const int64_t cap = (int64_t)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    HRM_LAUNCH(minhash_warp_kernel, (unsigned)blocks, 256, 0, s, src, n, k, H, d_sigs, d_valid);
    return HRM_OK;
}

hrm_status minhash_rows(const uint32_t* d_seq2bit, int64_t pitch_words, const int32_t* d_lengths, int64_t n, int k,
                        int H, uint64_t* d_sigs, uint8_t* d_valid, cudaStream_t s)
{
    SeqSource src{d_seq2bit, pitch_words, d_lengths, 0, 0, 0, 0};
    return launch_minhash(src, n, k, H, d_sigs, d_valid, s);
}

hrm_status minhash_windows(const uint32_t* d_chrom2bit, int64_t chrom_len, int k, int w, int H, int64_t first_window,
                           int64_t n_windows, uint64_t* d_sigs, uint8_t* d_valid, cudaStream_t s)
{
    SeqSource src{d_chrom2bit, (chrom_len + 15) / 16, nullptr, chrom_len, first_window, w - k + 1, w};
    return launch_minhash(src, n_windows, k, H, d_sigs, d_valid, s);
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_minhash(const uint32_t* d_seq2bit, int64_t pitch_words, const int32_t* d_lengths, int64_t n,
                                  int k, int H, uint64_t* d_sigs, uint8_t* d_valid, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(k >= 1 && k <= 32, "1 <= k <= 32 (ref: main_gpu.cu:1305)");
    HRM_REQUIRE(H >= 1 && H <= 64, "1 <= H <= 64 (ref: gpusequencehasher.cuh:139)");
    HRM_REQUIRE(n >= 0 && pitch_words > 0, "sizes");
    return minhash_rows(d_seq2bit, pitch_words, d_lengths, n, k, H, d_sigs, d_valid, as_stream(stream));
}

extern "C" hrm_status hrm_minhash_windows(const uint32_t* d_chrom2bit, int64_t chrom_len, int k, int w, int H,
                                          int64_t first_window, int64_t n_windows, uint64_t* d_sigs, uint8_t* d_valid,
                                          hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(k >= 1 && k <= 32, "1 <= k <= 32");
    HRM_REQUIRE(H >= 1 && H <= 64, "1 <= H <= 64");
    HRM_REQUIRE(w >= k, "window size must be >= k");
    HRM_REQUIRE(n_windows >= 0 && first_window >= 0 && chrom_len >= 0, "sizes");
    return minhash_windows(d_chrom2bit, chrom_len, k, w, H, first_window, n_windows, d_sigs, d_valid,
                           as_stream(stream));
}
