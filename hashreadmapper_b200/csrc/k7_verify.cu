// k7_verify.cu -- K6/K7: verification of mapped reads on the device (integer-ALU bound).
// ref: Mappinghandler::CSSW src/gpu/mappinghandler.cu:383-600 (serial per-read prep on one host
//      thread, then 2 x Aligner::Align per read on a host ThreadPool), edlibAligner :841-1010,
//      StripedSmithWaterman::Aligner::Align src/ssw_cpp.cpp:361-400, ssw_align src/ssw.c:818-922
//      (sw_sse2_byte :197-386, sw_sse2_word :412-588, banded_sw :590-774), edlibAlign src/edlib.cpp:1474.
//
// Per batch of alignments:
//  A. passes -- score, ends, begins, second best.  Fused path: sw_pair_passes_kernel, 4 lanes per read, both
//     alignments of a read in the s16x2 halves of every register (core_swpair.cuh, DPX VIMNMX3 / VIADDMNMX),
//     anti-diagonal wavefront, forward + reverse pass.  Function-level API (ASCII rows, any size up to 512):
//     sw_passes_kernel, one warp per alignment, 32-bit cells.  Inputs are built on the fly from the packed read
//     and the packed genome (stage-V 3N conversion applied on the codes): no ASCII, no host prep loop.
//  B. trace back + CIGAR -- the reference's banded_sw doubles its band until the banded score reaches the
//     alignment score.  sw_classify_kernel puts every alignment on the list of its first band's class
//     [2^c, 2^(c+1)); sw_finish_band_kernel runs ONE band iteration per alignment and class, one thread per
//     alignment (core_swband.cuh), and either finishes the alignment (trace back, =/X/I/D/S string) or hands it to
//     the next class; one launch per class, ascending.  sw_finish_kernel (one warp per alignment, the literal
//     band arrays) takes what exceeds the ladder's limits.
// Work per alignment: ~(L+pad) x w cell updates forward, <= that backward, rows x (2 band + 1) per band iteration.
#include "pipeline.cuh"
#include "core_sw.cuh"
#include "core_swpair.cuh"
#include "core_swband.cuh"
#include <stdlib.h>

namespace hrm {

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// sources of (query codes, ref codes, maskLen)
// ------------------------------------------------------------------------------------------------
struct AsciiSrc { // function-level API: ASCII rows
    const char* queries;
    int64_t qpitch;
    const int32_t* qlen;
    const char* refs;
    int64_t rpitch;
    const int32_t* rlen;
    const int32_t* mask_len;
    int maxQ, maxR;
    static constexpr int SKIP_FLAG = 1; // sizes outside the limits: reported like a failed trace back
    struct Item { // random access to the codes of one alignment (band kernels)
        const char* q_;
        const char* r_;
        int ql, rl;
        __device__ __forceinline__ int q(int i) const { return sw_translate((unsigned char)q_[i]); }
        __device__ __forceinline__ int r(int j) const { return sw_translate((unsigned char)r_[j]); }
    };
    __device__ __forceinline__ bool open(int64_t e, Item& it) const
    {
        it.ql = qlen[e];
        it.rl = rlen[e];
        it.q_ = queries + e * qpitch;
        it.r_ = refs + e * rpitch;
        return !(it.ql < 0 || it.ql > maxQ || it.rl < 0 || it.rl > maxR);
    }
    // returns false when the item must be skipped
    __device__ __forceinline__ bool load(int64_t e, int8_t* sq, int8_t* sr, int tid, int nthr, int& ql, int& rl,
                                         int& ml) const
    {
        ql = qlen[e];
        rl = rlen[e];
        ml = mask_len[e];
        if (ql < 0 || ql > maxQ || rl < 0 || rl > maxR) return false;
        for (int i = tid; i < ql; i += nthr) sq[i] = sw_translate((unsigned char)queries[e * qpitch + i]);
        for (int i = tid; i < rl; i += nthr) sr[i] = sw_translate((unsigned char)refs[e * rpitch + i]);
        return true;
    }
};

__device__ __forceinline__ int conv_code(int c, int conv)
{
    if (conv == HRM_CONV_CT && c == 1) return 3;
    if (conv == HRM_CONV_GA && c == 2) return 0;
    return c;
}

struct PackedSrc { // fused path: item t = 2 * read + a
    VerifyParams VP;
    const int32_t* read_len;
    const hrm_mapped_read* mapped;
    int maxQ, maxR;
    static constexpr int SKIP_FLAG = 0; // unmapped read: zero-initialised alignments (ref: mappinghandler.cu:548)
    struct Item { // random access to the codes of one alignment (band kernels)
        const uint32_t* rw;
        const uint32_t* cw;
        int64_t pos;
        int ql, rl, conv;
        bool rc;
        __device__ __forceinline__ int q(int i) const
        {
            return conv_code(rc ? 3 - (int)get_nuc(rw, ql - 1 - i) : (int)get_nuc(rw, i), conv);
        }
        __device__ __forceinline__ int r(int j) const { return conv_code((int)get_nuc(cw, pos + j), conv); }
    };
    __device__ __forceinline__ bool open(int64_t t, Item& it) const
    {
        const int64_t rd = t >> 1;
        const hrm_mapped_read m = mapped[rd];
        const int L = read_len[rd];
        it.ql = L;
        it.rl = 0;
        if (m.orientation == HRM_ORIENT_NONE || m.pass < 0 || m.pass >= VP.num_passes || L <= 0 || L > maxQ)
            return false;
        const VerifyPass& P = VP.pass[m.pass];
        it.rw = P.reads + rd * P.read_pitch;
        it.cw = P.G.chrom_words[m.chromosome_id];
        const int64_t clen = P.G.chrom_len[m.chromosome_id];
        int wl = (int)((m.position + VP.w < clen) ? VP.w : clen - m.position); // ref: mappinghandler.cu:434-440
        if (wl > maxR) wl = maxR;
        it.rl = wl;
        it.pos = m.position;
        it.conv = P.verify_conv;
        it.rc = (m.orientation == HRM_ORIENT_REVCOMP) != ((t & 1) == 1);
        return true;
    }
    __device__ __forceinline__ bool load(int64_t t, int8_t* sq, int8_t* sr, int tid, int nthr, int& ql, int& rl,
                                         int& ml) const
    {
        const int64_t rd = t >> 1;
        const int a = (int)(t & 1);
        const hrm_mapped_read m = mapped[rd];
        const int L = read_len[rd];
        ql = L;
        rl = 0;
        ml = L / 2 < 15 ? 15 : L / 2; // ref: mappinghandler.cu:453-454
        if (m.orientation == HRM_ORIENT_NONE || m.pass < 0 || m.pass >= VP.num_passes || L <= 0 || L > maxQ)
            return false;
        const VerifyPass& P = VP.pass[m.pass];
        const uint32_t* rw = P.reads + rd * P.read_pitch;
        const int64_t clen = P.G.chrom_len[m.chromosome_id];
        const uint32_t* cw = P.G.chrom_words[m.chromosome_id];
        int wl = (int)((m.position + VP.w < clen) ? VP.w : clen - m.position); // ref: mappinghandler.cu:434-440
        if (wl > maxR) wl = maxR;
        rl = wl;
        // readsequence = RC(read) when SHD chose RC (:420-423); alignment 1 uses RC(readsequence) (:461-465)
        const bool rc = (m.orientation == HRM_ORIENT_REVCOMP) != (a == 1);
        for (int j = tid; j < L; j += nthr) {
            const int c = rc ? 3 - (int)get_nuc(rw, L - 1 - j) : (int)get_nuc(rw, j);
            sq[j] = (int8_t)conv_code(c, P.verify_conv);
        }
        for (int j = tid; j < wl; j += nthr) sr[j] = (int8_t)conv_code((int)get_nuc(cw, m.position + j), P.verify_conv);
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// kernel A: wavefront passes
// ------------------------------------------------------------------------------------------------
struct WfOut {
    int maxv, end_col, end_row;
};

// rows [lane*R, lane*R+R) of the (padded) read; columns col0, col0+cstep, ... (ncols of them)
template <int R>
__device__ __forceinline__ WfOut wf_pass(const int8_t* sq, int qoff, int qdir, int nreal, int padB, int padW,
                                         const int8_t* sr, int col0, int cstep, int ncols, int terminate,
                                         int init_end_col, int16_t* cmB, int16_t* cmW, int lane)
{
    const unsigned FULL = 0xffffffffu;
    int Hp[R], E[R], qc[R];
    const int row0 = lane * R;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = row0 + r;
        Hp[r] = 0;
        E[r] = 0;
        qc[r] = row < nreal ? (int)sq[qoff + qdir * row] : -1; // pad rows score 0 (ref: qP_byte ssw.c:177)
    }
    int outH = 0, outF = 0;
    unsigned outKB = 0, outKW = 0; // (column max << 16) | (0xFFFF - first row attaining it)
    int prevRecvH = 0;
    int runmax = 0, end_col = init_end_col, end_row = 0, done = 0;
    const int nsteps = ncols + 31;
    for (int t = 0; t < nsteps; ++t) {
        int recvH = __shfl_up_sync(FULL, outH, 1);
        int recvF = __shfl_up_sync(FULL, outF, 1);
        unsigned recvKB = __shfl_up_sync(FULL, outKB, 1);
        unsigned recvKW = __shfl_up_sync(FULL, outKW, 1);
        if (lane == 0) {
            recvH = 0;
            recvF = 0;
            recvKB = 0;
            recvKW = 0;
        }
        const int crel = t - lane;
        if (crel >= 0 && crel < ncols) {
            const int col = col0 + cstep * crel;
            const int rc = sr[col];
            int diag = crel == 0 ? 0 : prevRecvH;
            prevRecvH = recvH;
            int F = recvF;
            unsigned kB = recvKB, kW = recvKW;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int row = row0 + r;
                const int hprev = Hp[r];
                const int s = qc[r] < 0 ? 0 : ((rc == qc[r] && rc < 4) ? 2 : -2);
                int h = max(max(diag + s, 0), max(E[r], F));
                diag = hprev;
                Hp[r] = h;
                const unsigned key = ((unsigned)h << 16) | (unsigned)(0xFFFF - row);
                if (row < padB) kB = max(kB, key);
                if (row < padW) kW = max(kW, key);
                const int ho = max(h - HRM_SW_GAPO, 0);
                E[r] = max(E[r] - HRM_SW_GAPE, ho);
                F = max(F - HRM_SW_GAPE, ho);
            }
            outH = Hp[R - 1];
            outF = F;
            outKB = kB;
            outKW = kW;
            if (lane == 31) {
                cmB[col] = (int16_t)(kB >> 16);
                cmW[col] = (int16_t)(kW >> 16);
                if (!done) {
                    const int cm = (int)(kB >> 16);
                    if (cm > runmax) { // ref: ssw.c:321-335 strict increase; :343-351 smallest row
                        runmax = cm;
                        end_col = col;
                        end_row = 0xFFFF - (int)(kB & 0xFFFFu);
                    }
                    if (cm == terminate) done = 1; // ref: ssw.c:339
                }
            }
        }
        if (terminate >= 0 && __shfl_sync(FULL, done, 31)) break;
    }
    WfOut o;
    o.maxv = __shfl_sync(FULL, runmax, 31);
    o.end_col = __shfl_sync(FULL, end_col, 31);
    o.end_row = __shfl_sync(FULL, end_row, 31);
    return o;
}

__device__ __forceinline__ void zero_alignment(hrm_alignment& o)
{
    o.sw_score = o.sw_score_next_best = o.ref_begin = o.ref_end = o.query_begin = o.query_end = 0;
    o.ref_end_next_best = o.mismatches = o.flag = o.cigar_len = 0;
}

// best column maximum over [lo, hi) of cm[]: (value, first column) ; warp-wide
__device__ __forceinline__ unsigned scan_cols(const int16_t* cm, int lo, int hi, int lane)
{
    unsigned best = 0;
    for (int c = lo + lane; c < hi; c += 32) {
        const unsigned key = ((unsigned)(uint16_t)cm[c] << 16) | (unsigned)(0xFFFF - c);
        best = key > best ? key : best;
    }
    return best;
}

template <int R, class Src>
__global__ void __launch_bounds__(256) sw_passes_kernel(Src src, int64_t n, int QP, int RP,
                                                        hrm_alignment* __restrict__ out, int64_t out_stride_items,
                                                        int64_t out_offset_items)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = (size_t)QP + RP + 4 * (size_t)RP;
    int8_t* sq = (int8_t*)(smem + per_warp * wid);
    int8_t* sr = sq + QP;
    int16_t* cmB = (int16_t*)(sr + RP);
    int16_t* cmW = cmB + RP;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp0; e < n; e += nwarps) {
        int ql, rl, ml;
        __syncwarp();
        const bool ok = src.load(e, sq, sr, lane, 32, ql, rl, ml);
        __syncwarp();
        hrm_alignment o;
        zero_alignment(o);
        if (ok && ql > 0) {
            const int padB = (int)align_up(ql, 16), padW = (int)align_up(ql, 8);
            WfOut f = wf_pass<R>(sq, 0, 1, ql, padB, padW, sr, 0, 1, rl, -1, -1, cmB, cmW, lane);
            __syncwarp();
            const int score1 = f.maxv, ref_end1 = f.end_col, read_end1 = f.end_row < ql - 1 ? f.end_row : ql - 1;
            const bool word = score1 + 2 >= 255; // ref: ssw.c:329 byte overflow -> word pass (:846-849)
            o.sw_score = score1;
            o.ref_end = ref_end1;
            o.query_end = read_end1;
            // second best (ref: ssw.c:368-381 byte, :570-583 word)
            int score2 = 0, ref2 = 0;
            if (score1 > 0 || word) {
                const int16_t* cm = word ? cmW : cmB;
                const int e1 = (ref_end1 - ml) > 0 ? (ref_end1 - ml) : 0;
                int e2 = (ref_end1 + ml) > rl ? rl : (ref_end1 + ml);
                unsigned k1 = scan_cols(cm, 0, e1, lane);
                unsigned k2 = scan_cols(cm, e2 + (word ? 0 : 1), rl, lane);
                k1 = __reduce_max_sync(0xffffffffu, k1);
                k2 = __reduce_max_sync(0xffffffffu, k2);
                // the reference scans the left range first and replaces only on strictly greater
                const unsigned v1 = k1 >> 16, v2 = k2 >> 16;
                if (v1 > 0) {
                    score2 = (int)v1;
                    ref2 = 0xFFFF - (int)(k1 & 0xFFFFu);
                }
                if (v2 > (unsigned)score2) {
                    score2 = (int)v2;
                    ref2 = 0xFFFF - (int)(k2 & 0xFFFFu);
                }
            }
            o.sw_score_next_best = ml >= 15 ? score2 : 0;
            o.ref_end_next_best = ml >= 15 ? ref2 : -1;
            if (score1 == 0 || ref_end1 < 0) { // undefined in the reference (ssw.c:220): deterministic convention
                o.ref_begin = -1;
                o.query_begin = -1;
            } else {
                __syncwarp();
                const int nr = read_end1 + 1;
                const int pad = (int)align_up(nr, word ? 8 : 16);
                WfOut b = wf_pass<R>(sq, read_end1, -1, nr, pad, pad, sr, ref_end1, -1, ref_end1 + 1, score1,
                                     word ? 0 : -1, cmB, cmW, lane);
                const int brow = b.end_row < nr - 1 ? b.end_row : nr - 1;
                o.ref_begin = b.end_col;
                o.query_begin = read_end1 - brow;
                o.flag = score1 > b.maxv ? 2 : 0; // ref: ssw.c:890-893
            }
        }
        if (!ok) o.flag = Src::SKIP_FLAG;
        if (lane == 0) out[out_offset_items + e * out_stride_items] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// kernel A2 (fused path): both alignments of a read in the halves of one register (core_swpair.cuh)
// ------------------------------------------------------------------------------------------------
// A group of G lanes owns one read: its two alignments (read and RC(read) against the same window) share
// the reference column, so every DP instruction (VIMNMX3 / VIADDMNMX .S16x2) updates two cells.  Lane l of
// the group keeps rows [l*R, l*R+R) of the frame in registers; the wavefront needs 3 shuffles per column.
// The reverse pass runs both halves over the shared columns cmax..0, a half joining at its own end column.
template <int G>
__device__ __forceinline__ uint32_t group_max(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, v, d, G);
        v = o > v ? o : v;
    }
    return v;
}

template <int G, int R>
__global__ void __launch_bounds__(128) sw_pair_passes_kernel(PackedSrc src, int64_t nreads, int QP, int RP,
                                                             hrm_alignment* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int PPW = 32 / G; // reads per warp
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane % G, gp = lane / G;
    const size_t per_pair = (size_t)QP + RP + 8 * (size_t)RP;
    unsigned char* pbase = smem + per_pair * (size_t)(wid * PPW + gp);
    int8_t* sq = (int8_t*)pbase;               // oriented read, unconverted codes
    int8_t* sr = sq + QP;                      // window, stage-V converted codes
    uint32_t* cmW = (uint32_t*)(sr + RP);      // column maxima incl. word-mode pad rows (A lo, B hi)
    uint32_t* cmB = cmW + RP;                  // ... incl. byte-mode pad rows
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int rtot = G * R;
    PairWave<R> w;
    for (int64_t rd0 = warp0 * PPW; rd0 < nreads; rd0 += nwarps * PPW) {
        const int64_t rd = rd0 + gp;
        bool ok = rd < nreads;
        int L = 0, rl = 0, conv = 0;
        __syncwarp();
        if (ok) {
            const hrm_mapped_read m = src.mapped[rd];
            L = src.read_len[rd];
            if (m.orientation == HRM_ORIENT_NONE || m.pass < 0 || m.pass >= src.VP.num_passes || L <= 0 || L > src.maxQ) {
                ok = false;
            } else {
                const uint32_t* rw = src.VP.pass[m.pass].reads + rd * src.VP.pass[m.pass].read_pitch;
                const int64_t clen = src.VP.pass[m.pass].G.chrom_len[m.chromosome_id];
                const uint32_t* cw = src.VP.pass[m.pass].G.chrom_words[m.chromosome_id];
                conv = src.VP.pass[m.pass].verify_conv;
                int wl = (int)((m.position + src.VP.w < clen) ? src.VP.w : clen - m.position); // ref: mappinghandler.cu:434-440
                if (wl > src.maxR) wl = src.maxR;
                rl = wl;
                const bool rcq = m.orientation == HRM_ORIENT_REVCOMP; // ref: mappinghandler.cu:420-423
                for (int j = gl; j < L; j += G) sq[j] = (int8_t)(rcq ? 3 - (int)get_nuc(rw, L - 1 - j) : (int)get_nuc(rw, j));
                for (int j = gl; j < wl; j += G) sr[j] = (int8_t)conv_code((int)get_nuc(cw, m.position + j), conv);
            }
        }
        __syncwarp();
        const int ml = L / 2 < 15 ? 15 : L / 2; // ref: mappinghandler.cu:453-454
        const int top = pair_top(rtot, L);
        // alignment 0: 3N(readsequence), alignment 1: 3N(RC(readsequence)) (ref: mappinghandler.cu:456-465)
        auto qcode = [&](int half, int j) -> int { return conv_code(half ? 3 - (int)sq[L - 1 - j] : (int)sq[j], conv); };

        // ---- forward pass -----------------------------------------------------------------------------
        pair_wave_reset(w);
        pair_build(w.L, gl * R, [&](int half, int g) -> int {
            const int j = g - top;
            if (!ok || j < 0) return PAIR_CODE_EMPTY;
            if (j >= L) return PAIR_CODE_PAD;
            return qcode(half, j);
        });
        const int ncolsF = ok ? rl : 0;
        int nsteps = __reduce_max_sync(FULL, ncolsF > 0 ? ncolsF + G - 1 : 0);
        for (int t = 0; t < nsteps; ++t) {
            uint32_t recvS = __shfl_up_sync(FULL, w.outS, 1, G);
            uint32_t recvF = __shfl_up_sync(FULL, w.outF, 1, G);
            uint32_t recvC = __shfl_up_sync(FULL, w.outCm, 1, G);
            if (gl == 0) {
                recvS = 0u;
                recvF = PX_TWOS;
                recvC = 0u;
            }
            const int c = t - gl;
            const int rc = (c >= 0 && c < ncolsF) ? (int)sr[c] : 0;
            uint32_t a, b;
            if (pair_wave_step(w, c, ncolsF, rc, PX_ONES, recvS, recvF, recvC, (uint32_t)c, a, b) && gl == G - 1) {
                cmW[c] = a;
                cmB[c] = b;
            }
        }
        __syncwarp();
        const uint32_t wordF[2] = {group_max<G>(pair_best_word(w.bestA, gl * R)),
                                   group_max<G>(pair_best_word(w.bestB, gl * R))};
        hrm_alignment o[2];
        int score1[2], ref_end1[2], read_end1[2];
        bool valid[2];
        const bool samePad = pair_pad8(L) == pair_pad16(L);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            zero_alignment(o[h]);
            const uint32_t word = wordF[h];
            score1[h] = (int)(word >> 20);
            ref_end1[h] = word ? 1023 - (int)((word >> 10) & 1023u) : -1;
            const int g = 1023 - (int)(word & 1023u);
            read_end1[h] = word ? ((g - top) < L - 1 ? (g - top) : L - 1) : 0;
            valid[h] = ok && word != 0u;
            // second best (ref: ssw.c:368-381 byte, :570-583 word): first column of the largest column maximum
            // outside the mask around the end, left range first
            const bool wordMode = score1[h] + 2 >= 255; // ref: ssw.c:329 byte overflow -> word pass (:846-849)
            const uint32_t* cm = (wordMode || samePad) ? cmW : cmB;
            unsigned k1 = 0, k2 = 0;
            if (valid[h]) {
                const int e1 = (ref_end1[h] - ml) > 0 ? (ref_end1[h] - ml) : 0;
                const int e2 = (ref_end1[h] + ml) > rl ? rl : (ref_end1[h] + ml);
                for (int c = gl; c < e1; c += G) {
                    const unsigned key = (((cm[c] >> (16 * h)) & 0xFFFFu) << 16) | (unsigned)(0xFFFF - c);
                    k1 = key > k1 ? key : k1;
                }
                for (int c = e2 + (wordMode ? 0 : 1) + gl; c < rl; c += G) {
                    const unsigned key = (((cm[c] >> (16 * h)) & 0xFFFFu) << 16) | (unsigned)(0xFFFF - c);
                    k2 = key > k2 ? key : k2;
                }
            }
            k1 = group_max<G>(k1);
            k2 = group_max<G>(k2);
            int score2 = 0, ref2 = 0;
            if ((k1 >> 16) > 0) {
                score2 = (int)(k1 >> 16);
                ref2 = 0xFFFF - (int)(k1 & 0xFFFFu);
            }
            if ((int)(k2 >> 16) > score2) {
                score2 = (int)(k2 >> 16);
                ref2 = 0xFFFF - (int)(k2 & 0xFFFFu);
            }
            if (ok) {
                o[h].sw_score = score1[h];
                o[h].ref_end = ref_end1[h];
                o[h].query_end = read_end1[h];
                o[h].sw_score_next_best = ml >= 15 ? score2 : 0;
                o[h].ref_end_next_best = ml >= 15 ? ref2 : -1;
                o[h].ref_begin = -1; // score 0 is undefined in the reference (ssw.c:220): deterministic convention
                o[h].query_begin = -1;
            }
        }
        // ---- reverse pass: reversed prefixes, both halves over the shared columns cmax .. 0 -----------
        const int cmax = max(valid[0] ? ref_end1[0] : -1, valid[1] ? ref_end1[1] : -1);
        const int ncolsR = cmax + 1;
        __syncwarp();
        pair_wave_reset(w);
        pair_build(w.L, gl * R, [&](int half, int g) -> int {
            if (!valid[half] || g > read_end1[half]) return PAIR_CODE_EMPTY;
            return qcode(half, read_end1[half] - g);
        });
        bool done = !valid[0] && !valid[1];
        bool doneA = !valid[0], doneB = !valid[1];
        nsteps = __reduce_max_sync(FULL, ncolsR > 0 ? ncolsR + G - 1 : 0);
        for (int t = 0; t < nsteps; ++t) {
            uint32_t recvS = __shfl_up_sync(FULL, w.outS, 1, G);
            uint32_t recvF = __shfl_up_sync(FULL, w.outF, 1, G);
            uint32_t recvC = __shfl_up_sync(FULL, w.outCm, 1, G);
            if (gl == 0) {
                recvS = 0u;
                recvF = PX_TWOS;
                recvC = 0u;
            }
            const int j = t - gl;
            int rc = 0;
            uint32_t hmask = 0u;
            if (j >= 0 && j < ncolsR) {
                const int c = cmax - j;
                rc = (int)sr[c];
                hmask = ((valid[0] && c <= ref_end1[0]) ? 1u : 0u) | ((valid[1] && c <= ref_end1[1]) ? 0x10000u : 0u);
            }
            uint32_t a, b;
            if (pair_wave_step(w, j, ncolsR, rc, hmask, recvS, recvF, recvC, (uint32_t)j, a, b) && gl == G - 1) {
                if ((int)(b & 0xFFFFu) == score1[0]) doneA = true; // ref: ssw.c:339 stop at the first column reaching score1
                if ((int)(b >> 16) == score1[1]) doneB = true;
                done = doneA && doneB;
            }
            const int fin = __shfl_sync(FULL, (int)done, gp * G + G - 1);
            if (__all_sync(FULL, fin)) break;
        }
        const uint32_t wordR[2] = {group_max<G>(pair_best_word(w.bestA, gl * R)),
                                   group_max<G>(pair_best_word(w.bestB, gl * R))};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (valid[h]) {
                const uint32_t word = wordR[h];
                const int maxv = (int)(word >> 20);
                const int j = 1023 - (int)((word >> 10) & 1023u);
                int g = 1023 - (int)(word & 1023u);
                if (g > read_end1[h]) g = read_end1[h];
                o[h].ref_begin = cmax - j;
                o[h].query_begin = read_end1[h] - g;
                o[h].flag = score1[h] > maxv ? 2 : 0; // ref: ssw.c:890-893
            }
        }
        if (rd < nreads) {
            if (gl == 0) out[2 * rd] = o[0];
            if (gl == 1) out[2 * rd + 1] = o[1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel B: banded trace back + CIGAR, one WARP per alignment
// ------------------------------------------------------------------------------------------------
// The reference's banded_sw (ssw.c:590-774) walks the band cell by cell.  Its F recurrence
//   f(j) = max(h_c(j-1) - gapO, f(j-1) - gapE),  df(j) = (h_c(j-1) - gapO > f(j-1) - gapE) ? 5 : 4
// does not change -- value and direction code alike -- when h_c(j-1) = max(H'(j-1), max(f(j-1),0)) is
// replaced by H'(j-1) = max(E+, diagonal): where f dominates, opening from it (f - gapO) loses to
// extending it (f - gapE).  With g(j) = f(j) + j the recurrence becomes a prefix maximum of
// a(j) = H'(j-1) - gapO + j, so the lanes of a warp evaluate up to 32 band cells of a row at once and
// combine them with a shuffle scan; E and the diagonal come from the previous row kept in shared
// memory with the reference's own band-relative indexing (including its band-edge zeroing).
struct TracePool {
    unsigned char* base; // global: direction bytes, one slice per resident warp
    int64_t per_warp;
    int wmax;            // ints per h_b / e_b / h_c array
    int maxops;
};

__device__ __forceinline__ int warp_banded(const int8_t* ref, const int8_t* read, int refLen, int readLen, int score,
                                           int32_t* hb, int32_t* eb, int32_t* hc, uint8_t* dir, int64_t dir_cap,
                                           int lane, int& out_band, int& out_width_d)
{
    const unsigned FULL = 0xffffffffu;
    const int NEG = -(1 << 28);
    int band = refLen - readLen;
    band = (band < 0 ? -band : band) + 1;
    const int len = refLen > readLen ? refLen : readLen;
    int maxv = 0, width = 0, width_d = 0;
    do {
        width = band * 2 + 3;
        width_d = band * 2 + 1;
        if ((int64_t)width_d * readLen > dir_cap) return -2;
        for (int j = lane; j < width + 8; j += 32) {
            hb[j] = 0;
            eb[j] = 0;
            hc[j] = 0;
        }
        __syncwarp();
        int lmax = 0;
        for (int i = 0; i < readLen; i++) {
            const int beg = (i - band) > 0 ? (i - band) : 0;
            const int end = (i + band) < (refLen - 1) ? (i + band) : (refLen - 1);
            const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            if (lane == 0) { // ref: ssw.c:635
                hb[0] = 0;
                eb[0] = 0;
                hb[edge] = 0;
                eb[edge] = 0;
                hc[0] = 0;
            }
            __syncwarp();
            const int xi = beg; // max(i - band, 0)
            const int xp = (i - 1 - band) > 0 ? (i - 1 - band) : 0;
            uint8_t* line = dir + (int64_t)width_d * i;
            const int ri = read[i];
            int carry_g = beg - 1; // g(beg-1) = f_init + beg - 1 with f_init = 0
            int carry_H = 0;       // h_c[0]
            const int ncell = end - beg + 1;
            for (int c0 = 0; c0 < ncell; c0 += 32) {
                const int j = beg + c0 + lane;
                const bool act = j <= end;
                const int u = j - xi + 1;
                int e = 0, de = 2, e1 = 0, diagv = 0, Hq = 0;
                if (act) {
                    const int ue = j - xp + 1, ud = j - xp;
                    const int t1 = i == 0 ? -HRM_SW_GAPO : hb[ue] - HRM_SW_GAPO;
                    const int t2 = i == 0 ? -HRM_SW_GAPE : eb[ue] - HRM_SW_GAPE;
                    e = t1 > t2 ? t1 : t2;
                    de = t1 > t2 ? 3 : 2;
                    e1 = e > 0 ? e : 0;
                    diagv = hb[ud] + sw_score(ref[j], ri);
                    Hq = e1 > diagv ? e1 : diagv;
                }
                int Hleft = __shfl_up_sync(FULL, Hq, 1);
                if (lane == 0) Hleft = carry_H;
                const int a = act ? Hleft - HRM_SW_GAPO + j : NEG;
                int incl = a;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl = incl > o ? incl : o;
                }
                int prevIncl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) prevIncl = NEG;
                const int gprev = carry_g > prevIncl ? carry_g : prevIncl;
                const int g = carry_g > incl ? carry_g : incl;
                const int f = g - j;
                const int df = a > gprev ? 5 : 4;
                const int f1 = f > 0 ? f : 0;
                const int temp1 = e1 > f1 ? e1 : f1;
                const int h = temp1 > diagv ? temp1 : diagv;
                const int dh = temp1 <= diagv ? 1 : (e1 > f1 ? de : df);
                __syncwarp(); // every lane has read the previous row before anyone overwrites e_b
                if (act) {
                    eb[u] = e;
                    hc[u] = h;
                    line[j - xi] = (uint8_t)(0x80 | dh | (de == 3 ? 8 : 0) | (df == 5 ? 16 : 0));
                    lmax = h > lmax ? h : lmax;
                }
                const int last = (ncell - 1 - c0) < 31 ? (ncell - 1 - c0) : 31;
                carry_g = __shfl_sync(FULL, g, last);
                carry_H = __shfl_sync(FULL, Hq, last);
                __syncwarp();
            }
            for (int x = ncell + lane; x < width_d; x += 32) line[x] = 0; // cells outside the band: not written
            const int ulast = end - xi + 1;
            for (int j = 1 + lane; j <= ulast; j += 32) hb[j] = hc[j]; // ref: ssw.c:669
            __syncwarp();
        }
        lmax = __reduce_max_sync(FULL, lmax);
        maxv = lmax > maxv ? lmax : maxv;
        band *= 2;
    } while (maxv < score && band <= len);
    band /= 2;
    out_band = band;
    out_width_d = width_d;
    return 0;
}

// ---- work lists of the trace-back ladder ---------------------------------------------------------------
// list c (c <= max_cls): alignments whose next band iteration has a band in [2^c, 2^(c+1)); list max_cls + 1:
// everything the thread-per-alignment kernels cannot hold (kernel B3).  The band doubles after a failed
// iteration, so an alignment visits each class at most once, in ascending order: one launch per class.
struct BandLists {
    int32_t* items;  // (max_cls + 2) lists of `cap` entries
    int32_t* counts; // max_cls + 2
    int64_t cap;
    int max_cls;
    int dir_rows; // rows the direction pool holds per alignment
    __device__ __forceinline__ void push(int list, int32_t e) const
    {
        items[(int64_t)list * cap + atomicAdd(counts + list, 1)] = e;
    }
};

// classifier: every alignment with a trace back to do goes to the list of its first band; the rest get their
// terminating NUL (ref: ssw_align returns without a cigar when the score is 0 / banded_sw is never reached)
__global__ void __launch_bounds__(256) sw_classify_kernel(const hrm_alignment* __restrict__ out, int64_t n,
                                                          char* __restrict__ cigars, int64_t cigar_pitch, BandLists BL)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const hrm_alignment o = out[e];
        if (o.flag == 1 || o.sw_score <= 0 || o.ref_begin < 0) {
            if (o.cigar_len < cigar_pitch) cigars[e * cigar_pitch + o.cigar_len] = 0;
            continue;
        }
        const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.query_end - o.query_begin + 1;
        int band = refLen - readLen;
        band = (band < 0 ? -band : band) + 1;
        const int c = sw_band_class(band);
        BL.push(c <= BL.max_cls && readLen <= BL.dir_rows ? c : BL.max_cls + 1, (int32_t)e);
    }
}

// ---- B2: one band iteration per alignment and class, one THREAD per alignment (core_swband.cuh) -------------
// State of the row ((H, E) pairs, match masks of the window) in shared memory, word-interleaved over the
// threads of the block (conflict free); direction nibbles in a global pool, word-interleaved over the lanes
// of the warp (rows in lock step coalesce).  Success -> trace back + CIGAR in the same thread; a band that is
// still too narrow -> next class.
constexpr int BAND_TILE = 1024;
template <class Src>
__global__ void __launch_bounds__(128) sw_finish_band_kernel(Src src, BandLists BL, int cls, int band_max, int MW, int nw,
                                                             uint32_t* __restrict__ dir_pool, int64_t dir_words_per_warp,
                                                             hrm_alignment* __restrict__ out,
                                                             char* __restrict__ cigars, int64_t cigar_pitch)
{
    extern __shared__ __align__(16) uint32_t smw[];
    const int T = blockDim.x;
    const int nstate = 2 * band_max + 3;
    const BandState S{smw + threadIdx.x, T};
    const BandMasks M{smw + (size_t)nstate * T + threadIdx.x, T, MW};
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const BandDirs D{dir_pool + (gtid >> 5) * dir_words_per_warp + (threadIdx.x & 31), 32, nw};
    // Work of an item ~ rows x band cells and the lanes of a warp run rows and cells in lock step: every block
    // takes a contiguous share of the list and sorts it tile-wise by (band, rows) in shared memory, so that the
    // 32 items a warp holds at a time are nearly alike.
    __shared__ unsigned long long tile[BAND_TILE];
    __shared__ int next_group;
    const int64_t total = BL.counts[cls];
    const int32_t* list = BL.items + (int64_t)cls * BL.cap;
    const int64_t share = HRM_SDIV(total, (int64_t)gridDim.x);
    const int64_t share_lo = (int64_t)blockIdx.x * share;
    const int64_t share_hi = (share_lo + share) < total ? (share_lo + share) : total;
    for (int64_t tile0 = share_lo; tile0 < share_hi; tile0 += BAND_TILE) {
    __syncthreads();
    if (threadIdx.x == 0) next_group = 0;
    for (int t = threadIdx.x; t < BAND_TILE; t += T) {
        unsigned long long key = ~0ull;
        if (tile0 + t < share_hi) {
            const int32_t e = list[tile0 + t];
            const hrm_alignment o = out[e];
            const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.query_end - o.query_begin + 1;
            int band = refLen - readLen;
            band = (band < 0 ? -band : band) + 1;
            while (sw_band_class(band) < cls) band *= 2;
            key = ((unsigned long long)(unsigned)band << 48) | ((unsigned long long)(unsigned)readLen << 32) | (unsigned)e;
        }
        tile[t] = key;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= BAND_TILE; k2 <<= 1)
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            for (int t = threadIdx.x; t < BAND_TILE; t += T) {
                const int p = t ^ j2;
                if (p > t) {
                    const unsigned long long a = tile[t], b = tile[p];
                    if ((a > b) == ((t & k2) == 0)) {
                        tile[t] = b;
                        tile[p] = a;
                    }
                }
            }
            __syncthreads();
        }
    const int ntile = (share_hi - tile0) < BAND_TILE ? (int)(share_hi - tile0) : BAND_TILE;
    // warps pull groups of 32 neighbours, heaviest first (longest-processing-time order balances the warps)
    while (true) {
        int grp = 0;
        if ((threadIdx.x & 31) == 0) grp = atomicAdd(&next_group, 1);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        const int it = ntile - 1 - (grp * 32 + (int)(threadIdx.x & 31));
        if (grp * 32 >= ntile) break;
        if (it < 0) continue;
        const unsigned long long key = tile[it];
        const int64_t e = (int64_t)(uint32_t)key;
        hrm_alignment o = out[e];
        const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.query_end - o.query_begin + 1;
        const int len = refLen > readLen ? refLen : readLen;
        const int band = (int)(key >> 48); // the element of the doubling sequence in this class
        typename Src::Item item;
        if (band > band_max || readLen > BL.dir_rows || refLen > 32 * MW || !src.open(e, item)) {
            BL.push(BL.max_cls + 1, (int32_t)e);
            continue;
        }
        for (int c0 = 0; c0 < refLen; c0 += 32) { // match masks of the window sub-sequence
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
            const int lim = (refLen - c0) < 32 ? (refLen - c0) : 32;
            for (int t = 0; t < lim; t++) {
                const int c = item.r(o.ref_begin + c0 + t);
                m0 |= (uint32_t)(c == 0) << t;
                m1 |= (uint32_t)(c == 1) << t;
                m2 |= (uint32_t)(c == 2) << t;
                m3 |= (uint32_t)(c == 3) << t;
            }
            const int w = c0 >> 5;
            M.p[(int64_t)(0 * MW + w) * T] = m0;
            M.p[(int64_t)(1 * MW + w) * T] = m1;
            M.p[(int64_t)(2 * MW + w) * T] = m2;
            M.p[(int64_t)(3 * MW + w) * T] = m3;
        }
        for (int w = HRM_SDIV(refLen, 32); w < MW; w++)
            for (int c = 0; c < 4; c++) M.p[(int64_t)(c * MW + w) * T] = 0u;
        const int qb = o.query_begin;
        const int m = sw_band_iteration(M, [&](int i) -> int { return item.q(qb + i); }, refLen, readLen, band, S, D);
        if (m < o.sw_score && band * 2 <= len) { // ref: while (max < score && band_width <= len), ssw.c:672
            BL.push(sw_band_class(band * 2) <= BL.max_cls ? cls + 1 : BL.max_cls + 1, (int32_t)e);
            continue;
        }
        uint32_t steps[(HRM_SW_MAX_QUERY + HRM_SW_MAX_REF) / 16 + 2];
        int nsteps = sw_band_traceback(D, band, refLen, readLen, steps, (HRM_SW_MAX_QUERY + HRM_SW_MAX_REF) + 16);
        SwAlignment al;
        al.sw_score = o.sw_score;
        al.sw_score_next_best = o.sw_score_next_best;
        al.ref_begin = o.ref_begin;
        al.ref_end = o.ref_end;
        al.query_begin = o.query_begin;
        al.query_end = o.query_end;
        al.ref_end_next_best = o.ref_end_next_best;
        al.mismatches = 0;
        al.cigar_len = 0;
        al.flag = o.flag;
        if (nsteps < 0) { // ref: banded_sw failed -> flag 1, empty path (ssw.c:910)
            al.flag = 1;
            nsteps = 0;
        }
        char* cig = cigars + e * cigar_pitch;
        sw_emit_steps([&](int i) -> int { return item.q(i); }, item.ql, [&](int j) -> int { return item.r(j); }, &al, steps,
                      nsteps, cig, (int)cigar_pitch);
        o.mismatches = al.mismatches;
        o.flag = al.flag;
        o.cigar_len = al.cigar_len;
        out[e] = o;
        if (o.cigar_len < cigar_pitch) cig[o.cigar_len] = 0;
    }
    }
}

// ---- B1 / B2: one THREAD per alignment -------------------------------------------------------------
// The literal banded DP (core_sw.cuh: sw_banded_once) is most instruction-efficient with one alignment per
// thread; what limits it is where its state lives.  B1 (BWMAX = 1): true alignments of substitution-only
// reads need band 1 and one band iteration -- h_b/e_b/h_c in local arrays, direction bytes and codes in a
// private shared-memory slice.  B2 (BWMAX = 16): the gapped "other strand" alignments -- same code, direction
// bytes in a global scratch interleaved across the 32 lanes of the warp (lanes in lock step coalesce).
// Anything that needs a wider band or a longer op list than the kernel holds is appended to the next work
// list; the last stage (B3) is the warp-per-alignment kernel, which has no limits.
template <class Src, int BWMAX, int MAXOPS, bool FROM_LIST, bool SMEM_DIR>
__global__ void __launch_bounds__(128) sw_finish_thread_kernel(Src src, int64_t n, const int32_t* __restrict__ in_list,
                                                               const int32_t* __restrict__ in_count,
                                                               int slice_bytes, int QP, int RP, int dir_rows,
                                                               uint8_t* __restrict__ dir_pool,
                                                               hrm_alignment* __restrict__ out,
                                                               char* __restrict__ cigars, int64_t cigar_pitch,
                                                               BandLists BL)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int8_t* q = (int8_t*)(smem + (size_t)slice_bytes * threadIdx.x);
    int8_t* r = q + QP;
    constexpr int WD = 2 * BWMAX + 1;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = FROM_LIST ? (int64_t)*in_count : n;
    // direction bytes: shared slice (linear) or global, interleaved over the lanes of the warp
    DirLinear dlin{(uint8_t*)(r + RP)};
    DirInterleaved dint{dir_pool + (gtid >> 5) * ((int64_t)WD * dir_rows * 32), (int)(threadIdx.x & 31)};
    for (int64_t it = gtid; it < total; it += stride) {
        const int64_t e = FROM_LIST ? (int64_t)in_list[it] : it;
        hrm_alignment o = out[e];
        char* cig = cigars + e * cigar_pitch;
        if (o.flag == 1 || o.sw_score <= 0 || o.ref_begin < 0) {
            if (o.cigar_len < cigar_pitch) cig[o.cigar_len] = 0;
            continue;
        }
        const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.query_end - o.query_begin + 1;
        const int len = refLen > readLen ? refLen : readLen;
        int band = refLen - readLen;
        band = (band < 0 ? -band : band) + 1;
        bool defer = band > BWMAX || readLen > dir_rows;
        int defer_band = band; // the band the next stage has to run
        int ql = 0, rl = 0, ml = 0;
        if (!defer && !src.load(e, q, r, 0, 1, ql, rl, ml)) {
            defer = true;
            defer_band = -1; // not loadable here: the general kernel decides
        }
        if (!defer) {
            int32_t hb[2 * BWMAX + 3 + 8], eb[2 * BWMAX + 3 + 8], hc[2 * BWMAX + 3 + 8];
            int mx = 0;
            while (true) { // ref: do { ... band_width *= 2; } while (max < score && band_width <= len)
                int m;
                if (SMEM_DIR) m = sw_banded_once(r + o.ref_begin, q + o.query_begin, refLen, readLen, band, hb, eb, hc, dlin, WD);
                else m = sw_banded_once(r + o.ref_begin, q + o.query_begin, refLen, readLen, band, hb, eb, hc, dint, WD);
                mx = m > mx ? m : mx;
                if (mx < o.sw_score && band * 2 <= len) {
                    band *= 2;
                    if (band > BWMAX) {
                        defer = true;
                        defer_band = band;
                        break;
                    }
                    continue;
                }
                break;
            }
            if (!defer) {
                char ops[MAXOPS];
                int32_t lens[MAXOPS];
                int nops;
                if (SMEM_DIR) nops = sw_traceback(dlin, WD, 2 * band + 1, band, refLen, readLen, ops, lens, MAXOPS);
                else nops = sw_traceback(dint, WD, 2 * band + 1, band, refLen, readLen, ops, lens, MAXOPS);
                if (nops >= MAXOPS) { // may have been truncated: the band kernel redoes this band without an op limit
                    defer = true;
                    defer_band = band;
                }
                if (!defer) {
                    SwAlignment al;
                    al.sw_score = o.sw_score;
                    al.sw_score_next_best = o.sw_score_next_best;
                    al.ref_begin = o.ref_begin;
                    al.ref_end = o.ref_end;
                    al.query_begin = o.query_begin;
                    al.query_end = o.query_end;
                    al.ref_end_next_best = o.ref_end_next_best;
                    al.mismatches = 0;
                    al.cigar_len = 0;
                    al.flag = o.flag;
                    if (nops < 0) {
                        al.flag = 1;
                        nops = 0;
                    }
                    sw_emit_cigar(q, ql, r, &al, ops, lens, nops, cig, (int)cigar_pitch);
                    o.mismatches = al.mismatches;
                    o.flag = al.flag;
                    o.cigar_len = al.cigar_len;
                    out[e] = o;
                    if (o.cigar_len < cigar_pitch) cig[o.cigar_len] = 0;
                }
            }
        }
        if (defer) BL.push(defer_band > 0 && readLen <= BL.dir_rows ? sw_band_class(defer_band) : BL.max_cls + 1, (int32_t)e);
    }
}

// ---- B3: everything else, one WARP per alignment (work list written by B2) ---------------------------
template <class Src>
__global__ void __launch_bounds__(256) sw_finish_kernel(Src src, const int32_t* __restrict__ worklist,
                                                        const int32_t* __restrict__ work_count, int QP, int RP,
                                                        TracePool P, hrm_alignment* __restrict__ out,
                                                        char* __restrict__ cigars, int64_t cigar_pitch)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = (size_t)QP + RP + 3 * sizeof(int32_t) * (size_t)P.wmax +
                            align_up(P.maxops, 16) + sizeof(int32_t) * (size_t)P.maxops;
    unsigned char* base = smem + align_up((int64_t)per_warp, 16) * wid;
    int8_t* sq = (int8_t*)base;
    int8_t* sr = sq + QP;
    int32_t* hb = (int32_t*)(sr + RP);
    int32_t* eb = hb + P.wmax;
    int32_t* hc = eb + P.wmax;
    char* ops = (char*)(hc + P.wmax);
    int32_t* lens = (int32_t*)(ops + align_up(P.maxops, 16));
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint8_t* dir = P.base + warp0 * P.per_warp;
    const int64_t nwork = *work_count;
    for (int64_t wi = warp0; wi < nwork; wi += nwarps) {
        const int64_t e = worklist[wi];
        hrm_alignment o = out[e];
        char* cig = cigars + e * cigar_pitch;
        int ql, rl, ml;
        __syncwarp();
        if (o.flag != 1 && o.sw_score > 0 && o.ref_begin >= 0 && src.load(e, sq, sr, lane, 32, ql, rl, ml)) {
            __syncwarp();
            const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.query_end - o.query_begin + 1;
            int band = 0, width_d = 0, nops = 0, flag = o.flag;
            const int rc = warp_banded(sr + o.ref_begin, sq + o.query_begin, refLen, readLen, o.sw_score, hb, eb, hc, dir,
                                       P.per_warp, lane, band, width_d);
            __syncwarp();
            if (lane == 0) {
                SwAlignment al;
                al.sw_score = o.sw_score;
                al.sw_score_next_best = o.sw_score_next_best;
                al.ref_begin = o.ref_begin;
                al.ref_end = o.ref_end;
                al.query_begin = o.query_begin;
                al.query_end = o.query_end;
                al.ref_end_next_best = o.ref_end_next_best;
                al.mismatches = 0;
                al.cigar_len = 0;
                if (rc == 0) nops = sw_traceback(DirLinear{dir}, width_d, width_d, band, refLen, readLen, ops, lens, P.maxops);
                else nops = -1;
                if (nops < 0) { // ref: banded_sw failed -> flag 1, empty path (ssw.c:910)
                    flag = 1;
                    nops = 0;
                }
                al.flag = flag;
                sw_emit_cigar(sq, ql, sr, &al, ops, lens, nops, cig, (int)cigar_pitch);
                o.mismatches = al.mismatches;
                o.flag = al.flag;
                o.cigar_len = al.cigar_len;
                out[e] = o;
            }
        }
        if (lane == 0 && o.cigar_len < cigar_pitch) cig[o.cigar_len] = 0;
    }
}

// ---- edit distance (function-level API and edlib mode) ------------------------------------------
__global__ void __launch_bounds__(128) edit_rows_kernel(const char* __restrict__ queries, int64_t qpitch,
                                                        const int32_t* __restrict__ qlen,
                                                        const char* __restrict__ targets, int64_t tpitch,
                                                        const int32_t* __restrict__ tlen, int64_t n,
                                                        int32_t* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int ql = qlen[e], tl = tlen[e];
        int d = -1;
        if (ql >= 0 && ql <= 64 * HRM_MYERS_MAX_BLOCKS && tl >= 0)
            d = myers_nw((const unsigned char*)queries + e * qpitch, ql, (const unsigned char*)targets + e * tpitch, tl);
        out[e] = d;
    }
}

__global__ void __launch_bounds__(128) edit_packed_kernel(PackedSrc src, int64_t n, int slice_bytes, int QP,
                                                          hrm_read_record* __restrict__ records)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int8_t* q = (int8_t*)(smem + (size_t)slice_bytes * threadIdx.x);
    int8_t* r = q + QP;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += stride) {
        int ql, rl, ml, d = -1;
        if (src.load(t, q, r, 0, 1, ql, rl, ml) && ql <= 64 * HRM_MYERS_MAX_BLOCKS) {
            for (int j = 0; j < ql; j++) q[j] = (int8_t)("ACGT"[q[j] & 3]);
            for (int j = 0; j < rl; j++) r[j] = (int8_t)("ACGT"[r[j] & 3]);
            d = myers_nw((const unsigned char*)q, ql, (const unsigned char*)r, rl);
        }
        records[t >> 1].edit_distance[t & 1] = d;
    }
}

// record header: mapped read, window length, mask length (+ zeroed alignments / distances)
__global__ void __launch_bounds__(256) record_header_kernel(VerifyParams VP, const int32_t* __restrict__ read_len,
                                                            const hrm_mapped_read* __restrict__ mapped, int64_t n,
                                                            hrm_read_record* __restrict__ records)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const hrm_mapped_read m = mapped[i];
        const int L = read_len[i];
        hrm_read_record rec;
        rec.mapped = m;
        zero_alignment(rec.alignments[0]);
        zero_alignment(rec.alignments[1]);
        rec.edit_distance[0] = rec.edit_distance[1] = -1;
        int wl = 0;
        if (m.orientation != HRM_ORIENT_NONE && m.pass >= 0 && m.pass < VP.num_passes) {
            const int64_t clen = VP.pass[m.pass].G.chrom_len[m.chromosome_id];
            wl = (int)((m.position + VP.w < clen) ? VP.w : clen - m.position);
        }
        rec.window_length = wl;
        rec.mask_len = L / 2 < 15 ? 15 : L / 2;
        records[i] = rec;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// kernel A2 launcher; returns false when the batch does not fit the paired kernel's frames
static bool launch_pair_passes(const PackedSrc& src, int64_t nreads, int QP, int RP, hrm_alignment* d_out, cudaStream_t s,
                               hrm_status& st)
{
    st = HRM_OK;
    const int maxQ = src.maxQ, w = src.maxR;
    int G = 0, R = 0; // frames: 4 x 40 rows (reads up to 152 bp), 8 x 32 (up to 248), 8 x 34 (up to 264: 250 bp reads)
    if (pair_fits(4 * 40, maxQ, w)) G = 4, R = 40;
    else if (pair_fits(8 * 32, maxQ, w)) G = 8, R = 32;
    else if (pair_fits(8 * 34, maxQ, w)) G = 8, R = 34;
    else return false;
    const int ppw = 32 / G;
    const size_t smem = ((size_t)QP + RP + 8 * (size_t)RP) * ppw * 4;
    if (smem > 200 * 1024) return false;
    int64_t blocks = HRM_SDIV(nreads, (int64_t)(ppw * 4));
    static const int per_sm = getenv("HRM_K7A_BLOCKS_PER_SM") ? atoi(getenv("HRM_K7A_BLOCKS_PER_SM")) : 4;
    const int64_t cap = (int64_t)num_sms() * (per_sm > 0 ? per_sm : 4);
    if (blocks > cap) blocks = cap;
    auto go = [&](auto kern) -> hrm_status {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        HRM_LAUNCH(kern, (unsigned)blocks, 128, smem, s, src, nreads, QP, RP, d_out);
        return HRM_OK;
    };
    st = G == 4 ? go(sw_pair_passes_kernel<4, 40>)
                : (R == 32 ? go(sw_pair_passes_kernel<8, 32>) : go(sw_pair_passes_kernel<8, 34>));
    return true;
}

template <class Src>
static bool try_pair_passes(const Src&, int64_t, int, int, hrm_alignment*, int64_t, int64_t, cudaStream_t, hrm_status&)
{
    return false;
}
template <>
bool try_pair_passes<PackedSrc>(const PackedSrc& src, int64_t n, int QP, int RP, hrm_alignment* d_out, int64_t out_stride,
                                int64_t out_offset, cudaStream_t s, hrm_status& st)
{
    if (out_stride != 1 || out_offset != 0 || (n & 1)) return false;
    return launch_pair_passes(src, n / 2, QP, RP, d_out, s, st);
}

template <class Src>
static hrm_status run_sw(const Src& src, int64_t n, int maxQ, int maxR, hrm_alignment* d_out, int64_t out_stride,
                         int64_t out_offset, char* d_cigars, int64_t cigar_pitch, cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    const int QP = (int)align_up(maxQ > 16 ? maxQ : 16, 16), RP = (int)align_up(maxR > 16 ? maxR : 16, 16);
    // kernel A: the paired SIMD kernel on the fused path, the generic warp-per-alignment kernel otherwise
    hrm_status pst = HRM_OK;
    if (try_pair_passes(src, n, QP, RP, d_out, out_stride, out_offset, s, pst)) {
        HRM_TRY(pst);
    } else {
        const size_t smemA = ((size_t)QP + RP + 4 * (size_t)RP) * 8;
        int64_t blocks = HRM_SDIV(n, (int64_t)8);
        const int64_t cap = (int64_t)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        const int rows = (int)align_up(maxQ, 16);
#define HRM_LAUNCH_A(RR)                                                                                        \
    do {                                                                                                        \
        auto kern = sw_passes_kernel<RR, Src>;                                                                  \
        if (smemA > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA); \
        HRM_LAUNCH(kern, (unsigned)blocks, 256, smemA, s, src, n, QP, RP, d_out, out_stride, out_offset);       \
    } while (0)
        if (rows <= 128) HRM_LAUNCH_A(4);
        else if (rows <= 160) HRM_LAUNCH_A(5);
        else if (rows <= 256) HRM_LAUNCH_A(8);
        else if (rows <= 512) HRM_LAUNCH_A(16);
        else {
            set_error("query longer than 512 bases");
            return HRM_ERR_INVALID;
        }
#undef HRM_LAUNCH_A
    }
    // kernel B1 (thread/alignment, band 1, on chip) -> B2 ladder (thread/alignment, one launch per band class)
    // -> B3 (warp/alignment, whatever is left)
    {
        HRM_REQUIRE(n < (1LL << 31), "too many alignments in one call");
        const int maxQq = maxQ > 16 ? maxQ : 16;
        const int maxLen = (maxQ > maxR ? maxQ : maxR) > 16 ? (maxQ > maxR ? maxQ : maxR) : 16;
        BandLists BL;
        BL.max_cls = sw_band_class(maxLen) < 8 ? sw_band_class(maxLen) : 8; // bands never exceed max(len)
        BL.cap = n;
        BL.dir_rows = maxQq;
        const int nlists = BL.max_cls + 2;
        Scratch wl;
        HRM_TRY(wl.alloc(sizeof(int32_t) * ((size_t)nlists * (size_t)n + 64), s));
        BL.counts = wl.as<int32_t>();
        BL.items = wl.as<int32_t>() + 64;
        HRM_CUDA(cudaMemsetAsync(BL.counts, 0, 64 * sizeof(int32_t), s));
        static const bool use_b1 = getenv("HRM_SW_B1") != nullptr; // experiment switch: band-1 kernel in front
        if (!use_b1) {
            int64_t blocks = HRM_SDIV(n, (int64_t)256);
            if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
            HRM_LAUNCH(sw_classify_kernel, (unsigned)blocks, 256, 0, s, d_out, n, d_cigars, cigar_pitch, BL);
        } else {
            int slice = (int)align_up(QP + RP + 3 * maxQq, 4) + 4; // odd number of words
            if (((slice / 4) & 1) == 0) slice += 4;
            const int threads = 128;
            const size_t smemS = (size_t)slice * threads;
            int64_t blocks = HRM_SDIV(n, (int64_t)threads);
            const int64_t cap = (int64_t)num_sms() * 8;
            if (blocks > cap) blocks = cap;
            auto kern = sw_finish_thread_kernel<Src, 1, 40, false, true>;
            if (smemS > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS);
            HRM_LAUNCH(kern, (unsigned)blocks, threads, smemS, s, src, n, (const int32_t*)nullptr,
                       (const int32_t*)nullptr, slice, QP, RP, maxQq, (uint8_t*)nullptr, d_out, d_cigars, cigar_pitch,
                       BL);
        }
        // B2 ladder: geometry per class, one direction pool shared by the launches
        struct ClassGeo {
            int band_max, nw, threads;
            size_t smem;
            int64_t blocks, words_per_warp;
        } geo[16];
        const int MW = HRM_SDIV(maxR > 16 ? maxR : 16, 32);
        const int64_t POOL_LIMIT = 768LL << 20;
        int64_t pool_bytes = 256;
        for (int c = 0; c <= BL.max_cls; c++) {
            ClassGeo& g = geo[c];
            g.band_max = (2 << c) - 1 < maxLen ? (2 << c) - 1 : maxLen;
            g.nw = (2 * g.band_max + 8) / 8;
            g.threads = 128;
            const size_t per_thread = sizeof(uint32_t) * (size_t)(2 * g.band_max + 3 + 4 * MW);
            while (per_thread * g.threads > 200 * 1024 && g.threads > 32) g.threads >>= 1;
            g.smem = per_thread * g.threads;
            HRM_REQUIRE(g.smem <= 200 * 1024, "band class does not fit shared memory");
            int resident = (int)((220 * 1024) / (g.smem + 1024));
            if (resident < 1) resident = 1;
            if (resident > 2048 / g.threads) resident = 2048 / g.threads;
            g.words_per_warp = (int64_t)BL.dir_rows * g.nw * 32;
            g.blocks = HRM_SDIV(n, (int64_t)g.threads);
            int64_t cap = (int64_t)num_sms() * resident;
            const int64_t fit = POOL_LIMIT / (g.words_per_warp * 4 * (g.threads / 32));
            if (cap > fit) cap = fit > 1 ? fit : 1;
            if (g.blocks > cap) g.blocks = cap;
            const int64_t need = g.blocks * (g.threads / 32) * g.words_per_warp * 4;
            if (need > pool_bytes) pool_bytes = need;
        }
        Scratch pool;
        HRM_TRY(pool.alloc((size_t)pool_bytes, s));
        for (int c = 0; c <= BL.max_cls; c++) {
            const ClassGeo& g = geo[c];
            auto kern = sw_finish_band_kernel<Src>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            HRM_LAUNCH(kern, (unsigned)g.blocks, g.threads, g.smem, s, src, BL, c, g.band_max, MW, g.nw, pool.as<uint32_t>(),
                       g.words_per_warp, d_out, d_cigars, cigar_pitch);
        }
        pool.release();
        int32_t* worklist = BL.items + (int64_t)(BL.max_cls + 1) * BL.cap;
        int32_t* work_count = BL.counts + BL.max_cls + 1;
        TracePool P;
        P.wmax = (int)align_up(2 * maxLen + 3 + 8 + 1, 4);
        P.maxops = 2 * maxLen + 8;
        P.per_warp = align_up((int64_t)(2 * maxLen + 1) * (maxQ > 16 ? maxQ : 16) + 16, 256);
        const size_t per_warp = (size_t)align_up((int64_t)((size_t)QP + RP + 3 * sizeof(int32_t) * (size_t)P.wmax +
                                                        align_up(P.maxops, 16) + sizeof(int32_t) * (size_t)P.maxops), 16);
        int warps = 8;
        while (per_warp * warps > 200 * 1024 && warps > 1) warps >>= 1;
        const size_t smemB = per_warp * warps;
        int resident = (int)((220 * 1024) / (smemB + 1024));
        if (resident < 1) resident = 1;
        if (resident > 8) resident = 8;
        int64_t blocks = HRM_SDIV(n, (int64_t)warps);
        const int64_t cap = (int64_t)num_sms() * resident; // one wave of resident CTAs, each loops over the list
        if (blocks > cap) blocks = cap;
        Scratch mem;
        HRM_TRY(mem.alloc((size_t)(blocks * warps * P.per_warp), s));
        P.base = mem.as<unsigned char>();
        auto kern = sw_finish_kernel<Src>;
        if (smemB > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB);
        HRM_LAUNCH(kern, (unsigned)blocks, warps * 32, smemB, s, src, worklist, work_count, QP, RP, P, d_out, d_cigars,
                   cigar_pitch);
    }
    return HRM_OK;
}

__global__ void __launch_bounds__(256) scatter_alignments_kernel(const hrm_alignment* __restrict__ al,
                                                                 hrm_read_record* __restrict__ records, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += stride)
        records[t >> 1].alignments[t & 1] = al[t];
}

hrm_status scatter_alignments(const hrm_alignment* al, hrm_read_record* records, int64_t n, cudaStream_t s)
{
    int64_t blocks = HRM_SDIV(2 * n, (int64_t)256);
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
    HRM_LAUNCH(scatter_alignments_kernel, (unsigned)blocks, 256, 0, s, al, records, n);
    return HRM_OK;
}

hrm_status verify_reads(const VerifyParams& VP, const int32_t* d_read_len, int64_t n, int max_read_len,
                        const hrm_mapped_read* d_mapped, hrm_read_record* d_records, char* d_cigars, int64_t cigar_pitch,
                        cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    int64_t hb = HRM_SDIV(n, (int64_t)256);
    if (hb > (int64_t)num_sms() * 16) hb = (int64_t)num_sms() * 16;
    HRM_LAUNCH(record_header_kernel, (unsigned)hb, 256, 0, s, VP, d_read_len, d_mapped, n, d_records);
    PackedSrc src;
    src.VP = VP;
    src.read_len = d_read_len;
    src.mapped = d_mapped;
    src.maxQ = max_read_len > 16 ? max_read_len : 16;
    src.maxR = VP.w;
    if (VP.mapper_type == HRM_MAPPER_SW) {
        // run the two alignments of every read as consecutive items into a dense temporary, then scatter
        Scratch tmp;
        HRM_TRY(tmp.alloc(sizeof(hrm_alignment) * (size_t)(2 * n), s));
        HRM_TRY(run_sw(src, 2 * n, src.maxQ, src.maxR, tmp.as<hrm_alignment>(), 1, 0, d_cigars, cigar_pitch, s));
        HRM_TRY(scatter_alignments(tmp.as<hrm_alignment>(), d_records, n, s));
    } else {
        const int QP = (int)align_up(src.maxQ, 16), RP = (int)align_up(src.maxR > 16 ? src.maxR : 16, 16);
        int slice = (int)align_up(QP + RP, 4) + 4;
        if (((slice / 4) & 1) == 0) slice += 4;
        const size_t smem = (size_t)slice * 128;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(edit_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int64_t blocks = HRM_SDIV(2 * n, (int64_t)128);
        if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
        HRM_LAUNCH(edit_packed_kernel, (unsigned)blocks, 128, smem, s, src, n, slice, QP, d_records);
    }
    return HRM_OK;
}

} // namespace hrm

using namespace hrm;

extern "C" hrm_status hrm_sw_align(const char* d_queries, int64_t query_pitch, const int32_t* d_query_len,
                                   const char* d_refs, int64_t ref_pitch, const int32_t* d_ref_len,
                                   const int32_t* d_mask_len, int64_t n, hrm_alignment* d_out, char* d_cigars,
                                   int64_t cigar_pitch, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && query_pitch > 0 && ref_pitch > 0 && cigar_pitch > 0, "sizes");
    if (n == 0) return HRM_OK;
    AsciiSrc src;
    src.queries = d_queries;
    src.qpitch = query_pitch;
    src.qlen = d_query_len;
    src.refs = d_refs;
    src.rpitch = ref_pitch;
    src.rlen = d_ref_len;
    src.mask_len = d_mask_len;
    src.maxQ = (int)(query_pitch < HRM_SW_MAX_QUERY ? query_pitch : HRM_SW_MAX_QUERY);
    src.maxR = (int)(ref_pitch < HRM_SW_MAX_REF ? ref_pitch : HRM_SW_MAX_REF);
    return run_sw(src, n, src.maxQ, src.maxR, d_out, 1, 0, d_cigars, cigar_pitch, as_stream(stream));
}

extern "C" hrm_status hrm_edit_distance(const char* d_queries, int64_t query_pitch, const int32_t* d_query_len,
                                        const char* d_targets, int64_t target_pitch, const int32_t* d_target_len,
                                        int64_t n, int32_t* d_distance, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && query_pitch > 0 && target_pitch > 0, "sizes");
    if (n == 0) return HRM_OK;
    int64_t blocks = HRM_SDIV(n, (int64_t)128);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    HRM_LAUNCH(edit_rows_kernel, (unsigned)blocks, 128, 0, as_stream(stream), d_queries, query_pitch, d_query_len,
               d_targets, target_pitch, d_target_len, n, d_distance);
    return HRM_OK;
}
