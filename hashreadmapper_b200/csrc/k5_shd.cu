// k5_shd.cu -- K5: shifted-Hamming best-window selection (integer ALU / shared-memory bound).
// ref: shiftedHammingDistanceWithFullOverlapKernelSmem1 src/gpu/hammingdistancekernels.cu:132-263
//      (one THREAD per candidate walking <=129 shifts x 2 orientations sequentially, anchors staged
//      block-transposed in shared memory, preceded by two 2bit->HiLo conversion kernels :293-312 and
//      by generateExtendedWindows2BitKernel include/gpu/windowgenerationkernels.cuh:162-262 which
//      re-packs every extended window from an ASCII genome slice copied H2D per batch);
//      host arg-min src/gpu/main_gpu.cu:777-821.
// Here: one WARP per candidate; the 32 lanes evaluate 32 (orientation, shift) pairs at a time from
// hi/lo bit planes held in shared memory, and a single __reduce_min_sync on the key
// (hd, orientation, shift) reproduces the reference's "first minimum" rule.  The extended window is
// cut directly out of the resident packed genome (no ASCII slice, no re-pack, no HiLo pre-pass), and
// in the fused kernel the per-read arg-min over candidate windows happens in the same warp.
// Per candidate: 4*(ceil(Lc/16) + ceil(La/16)) B read, 16 B written (fused: 32 B per READ).
#include "runtime.cuh"
#include "store.cuh"
#include "core_shd.cuh"

namespace hrm {

struct WarpShdSmem {
    uint32_t ahi[HRM_SHD_MAX_ANCHOR_WORDS];
    uint32_t alo[HRM_SHD_MAX_ANCHOR_WORDS];
    uint32_t rhi[2][HRM_SHD_MAX_READ_WORDS];
    uint32_t rlo[2][HRM_SHD_MAX_READ_WORDS];
};

// read planes (forward + reverse complement) into shared memory; all lanes call
__device__ __forceinline__ void warp_load_read(WarpShdSmem& S, const uint32_t* read, int64_t read_words, int Lc,
                                               int lane)
{
    const int nw = HRM_SDIV(Lc, 32);
    for (int j = lane; j < nw; j += 32) {
        uint32_t hi, lo;
        planes32(read, read_words, 32 * (int64_t)j, hi, lo);
        const int remain = Lc - 32 * j;
        if (remain < 32) {
            hi &= 0xFFFFFFFFu << (32 - remain);
            lo &= 0xFFFFFFFFu << (32 - remain);
        }
        S.rhi[0][j] = hi;
        S.rlo[0][j] = lo;
    }
    __syncwarp();
    for (int j = lane; j < nw; j += 32) {
        S.rhi[1][j] = rc_plane_word(S.rhi[0], nw, Lc, j);
        S.rlo[1][j] = rc_plane_word(S.rlo[0], nw, Lc, j);
    }
    __syncwarp();
}

// best key of one candidate (all lanes return it); HRM_SHD_INF when nothing is within thr
__device__ __forceinline__ uint32_t warp_shd(WarpShdSmem& S, const uint32_t* anchor, int64_t anchor_words,
                                             int64_t anchor_base, int La, int Lc, int thr, int lane)
{
    const int anw = HRM_SDIV(La, 32);
    __syncwarp();
    for (int j = lane; j < anw; j += 32) {
        uint32_t hi, lo;
        planes32(anchor, anchor_words, anchor_base + 32 * (int64_t)j, hi, lo);
        const int remain = La - 32 * j;
        if (remain < 32) {
            hi &= 0xFFFFFFFFu << (32 - remain);
            lo &= 0xFFFFFFFFu << (32 - remain);
        }
        S.ahi[j] = hi;
        S.alo[j] = lo;
    }
    __syncwarp();
    const int nshift = La - Lc + 1;
    uint32_t best = HRM_SHD_INF;
    for (int c = lane; c < 2 * nshift; c += 32) {
        const int o = c >= nshift ? 1 : 0;
        const int sft = c - o * nshift;
        const int hd = shd_at_shift(S.ahi, S.alo, anw, S.rhi[o], S.rlo[o], Lc, sft, thr);
        if (hd <= thr) {
            const uint32_t key = shd_key(hd, o, sft);
            best = key < best ? key : best;
        }
    }
    return __reduce_min_sync(0xffffffffu, best);
}

// ---- function-level API kernel: explicit anchors (extended windows), one warp per candidate -------
__global__ void __launch_bounds__(256) shd_rows_kernel(const uint32_t* __restrict__ anchors, int64_t anchor_pitch,
                                                       const int32_t* __restrict__ anchor_len,
                                                       const uint32_t* __restrict__ cands, int64_t cand_pitch,
                                                       const int32_t* __restrict__ cand_len, int64_t n, float rate,
                                                       int32_t* __restrict__ best_shift,
                                                       int32_t* __restrict__ best_score,
                                                       int8_t* __restrict__ best_orient)
{
    __shared__ WarpShdSmem sm[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpShdSmem& S = sm[wid];
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp0; e < n; e += nwarps) {
        const int La = anchor_len[e], Lc = cand_len[e];
        int shift = 0, score = Lc, orient = HRM_ORIENT_NONE;
        if (Lc <= La && Lc > 0 && Lc <= 32 * HRM_SHD_MAX_READ_WORDS && La <= 32 * HRM_SHD_MAX_ANCHOR_WORDS) {
            const int thr = shd_threshold(Lc, rate);
            warp_load_read(S, cands + e * cand_pitch, cand_pitch, Lc, lane);
            const uint32_t key = warp_shd(S, anchors + e * anchor_pitch, anchor_pitch, 0, La, Lc, thr, lane);
            if (key != HRM_SHD_INF) {
                shift = (int)(key & 0xFFFFu);
                score = (int)(key >> 20);
                orient = ((key >> 16) & 1u) ? HRM_ORIENT_REVCOMP : HRM_ORIENT_FORWARD;
            } else {
                score = thr + 1;
            }
        }
        if (lane == 0) {
            best_shift[e] = shift;
            best_score[e] = score;
            best_orient[e] = (int8_t)orient;
        }
    }
}

// ---- extended windows cut out of the packed genome (function-level API, S2) ----------------------
__global__ void __launch_bounds__(256) extended_windows_kernel(const uint32_t* __restrict__ chrom, int64_t chrom_len,
                                                               int w, const int32_t* __restrict__ window_pos,
                                                               const int32_t* __restrict__ read_len, int64_t n,
                                                               uint32_t* __restrict__ out, int64_t out_pitch,
                                                               int32_t* __restrict__ ext_left,
                                                               int32_t* __restrict__ ext_right,
                                                               int32_t* __restrict__ ext_len)
{
    const int64_t chrom_words = (chrom_len + 15) / 16;
    const int64_t total = n * out_pitch;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t e = t / out_pitch;
        const int wi = (int)(t - e * out_pitch);
        int left, right, len;
        window_location(chrom_len, window_pos[e], w, read_len[e] / 2, &left, &right, &len);
        uint32_t word = 0;
        const int remain = len - 16 * wi;
        if (remain > 0) {
            word = stream16(chrom, chrom_words, (int64_t)window_pos[e] - left + 16 * (int64_t)wi);
            if (remain < 16) word &= 0xFFFFFFFFu << (2 * (16 - remain));
        }
        out[t] = word;
        if (wi == 0) {
            ext_left[e] = left;
            ext_right[e] = right;
            ext_len[e] = len;
        }
    }
}

// ---- fused: per read, best window over its candidate list (S2 + S3 + S4) --------------------------
// cand_windows: global window ids, ascending inside each read's segment (K4 output).
__global__ void __launch_bounds__(256) best_window_kernel(const uint32_t* __restrict__ reads, int64_t read_pitch,
                                                          const int32_t* __restrict__ read_len, int64_t n,
                                                          const uint32_t* __restrict__ cand_windows,
                                                          const int32_t* __restrict__ cand_offsets,
                                                          const int2* __restrict__ cand_lists, GenomeDev G,
                                                          const int64_t* __restrict__ win_prefix, int k, int w,
                                                          float rate, int pass, hrm_mapped_read* __restrict__ out)
{
    __shared__ WarpShdSmem sm[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpShdSmem& S = sm[wid];
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int stride_bases = w - k + 1;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const int Lc = read_len[r];
        int cb, ce; // the read's candidate list: dense offsets (K4) or (start, count) (fused collection)
        if (cand_lists) {
            const int2 l = cand_lists[r];
            cb = l.x;
            ce = l.x + l.y;
        } else {
            cb = cand_offsets[r];
            ce = cand_offsets[r + 1];
        }
        hrm_mapped_read best;
        best.orientation = HRM_ORIENT_NONE;
        best.hamming_distance = 0;
        best.shift = 0;
        best.chromosome_id = 0;
        best.position = 0;
        best.pass = -1;
        best.reserved = 0;
        if (ce > cb && Lc > 0 && Lc <= 32 * HRM_SHD_MAX_READ_WORDS) {
            const int thr = shd_threshold(Lc, rate);
            warp_load_read(S, reads + r * read_pitch, read_pitch, Lc, lane);
            for (int c = cb; c < ce; c++) {
                const int64_t gw = cand_windows[c];
                int lo = 0, hi = G.n_chrom; // chromosome of global window gw
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (win_prefix[mid] <= gw) lo = mid;
                    else hi = mid;
                }
                const int chrom = lo;
                const int64_t clen = G.chrom_len[chrom];
                const int64_t p = (gw - win_prefix[chrom]) * stride_bases;
                int left, right, len;
                window_location(clen, p, w, Lc / 2, &left, &right, &len);
                if (Lc > len || len > 32 * HRM_SHD_MAX_ANCHOR_WORDS) continue; // ref: candidate longer than anchor => None
                const uint32_t key = warp_shd(S, G.chrom_words[chrom], (clen + 15) / 16, p - left, len, Lc, thr, lane);
                if (key == HRM_SHD_INF) continue;
                const int hd = (int)(key >> 20);
                // ref: main_gpu.cu:800-812 -- strictly smaller wins; windows arrive in ascending order
                if (best.orientation == HRM_ORIENT_NONE || best.hamming_distance > hd) {
                    best.orientation = ((key >> 16) & 1u) ? HRM_ORIENT_REVCOMP : HRM_ORIENT_FORWARD;
                    best.hamming_distance = hd;
                    best.shift = (int)(key & 0xFFFFu) - left;
                    best.chromosome_id = chrom;
                    best.position = p;
                    best.pass = pass;
                }
            }
        }
        if (lane == 0) out[r] = best;
    }
}

static unsigned warp_grid(int64_t nwarps_wanted, int waves)
{
    int64_t g = HRM_SDIV(nwarps_wanted, (int64_t)8);
    const int64_t cap = (int64_t)num_sms() * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

hrm_status best_windows(const uint32_t* d_reads, int64_t read_pitch, const int32_t* d_read_len, int64_t n,
                        const uint32_t* d_cand_windows, const int32_t* d_cand_offsets, const hrm_genome* g,
                        const int64_t* d_win_prefix, int k, int w, float rate, int pass, hrm_mapped_read* d_out,
                        cudaStream_t s, const int2* d_cand_lists)
{
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(best_window_kernel, warp_grid(n, 32), 256, 0, s, d_reads, read_pitch, d_read_len, n, d_cand_windows,
               d_cand_offsets, d_cand_lists, g->dev(), d_win_prefix, k, w, rate, pass, d_out);
    return HRM_OK;
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_shifted_hamming(const uint32_t* d_anchor2bit, int64_t anchor_pitch_words,
                                          const int32_t* d_anchor_len, const uint32_t* d_cand2bit,
                                          int64_t cand_pitch_words, const int32_t* d_cand_len, int64_t n,
                                          float max_error_rate, int32_t* d_best_shift, int32_t* d_best_score,
                                          int8_t* d_best_orientation, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && anchor_pitch_words > 0 && cand_pitch_words > 0, "sizes");
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(shd_rows_kernel, warp_grid(n, 32), 256, 0, as_stream(stream), d_anchor2bit, anchor_pitch_words,
               d_anchor_len, d_cand2bit, cand_pitch_words, d_cand_len, n, max_error_rate, d_best_shift, d_best_score,
               d_best_orientation);
    return HRM_OK;
}

extern "C" hrm_status hrm_extended_windows(const hrm_genome* g, int chrom, int w, const int32_t* d_window_pos,
                                           const int32_t* d_read_len, int64_t n, uint32_t* d_out,
                                           int64_t out_pitch_words, int32_t* d_ext_left, int32_t* d_ext_right,
                                           int32_t* d_ext_len, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(g != nullptr && chrom >= 0 && chrom < g->n_chrom, "genome/chromosome");
    HRM_REQUIRE(n >= 0 && out_pitch_words > 0 && w > 0, "sizes");
    if (n == 0) return HRM_OK;
    int64_t blocks = HRM_SDIV(n * out_pitch_words, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    HRM_LAUNCH(extended_windows_kernel, (unsigned)blocks, 256, 0, as_stream(stream), g->chrom_words[chrom],
               g->chrom_len[chrom], w, d_window_pos, d_read_len, n, d_out, out_pitch_words, d_ext_left, d_ext_right,
               d_ext_len);
    return HRM_OK;
}
