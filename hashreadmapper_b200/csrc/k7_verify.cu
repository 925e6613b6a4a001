// k7_verify.cu -- K6/K7: verification of mapped reads on the device.
// ref: Mappinghandler::CSSW src/gpu/mappinghandler.cu:383-600 (serial per-read prep on one host
//      thread, then 2 x Aligner::Align per read on a host ThreadPool), edlibAligner :841-1010,
//      StripedSmithWaterman::Aligner::Align src/ssw_cpp.cpp:361-400, ssw_align src/ssw.c:818-922,
//      edlibAlign src/edlib.cpp:1474-1476.
// Round-1 mapping: one thread per alignment running the scalar restatement of core_sw.cuh with its
// DP rows in a per-thread scratch slice (L1/L2 resident); inputs are built on the fly from the packed
// read and the packed genome (no ASCII materialisation, no host prep loop).  DESIGN.md lists the
// warp-wavefront version as the next optimisation of this kernel.
#include "pipeline.cuh"
#include "core_sw.cuh"

namespace hrm {

struct SwPool {
    unsigned char* base;
    int64_t per_thread;  // bytes per thread slice
    int maxQ, maxR, maxLen, maxops;
    int64_t dir_cap;
};

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static SwPool sw_pool_layout(int maxQ, int maxR)
{
    SwPool P;
    P.base = nullptr;
    P.maxQ = maxQ;
    P.maxR = maxR;
    P.maxLen = maxQ > maxR ? maxQ : maxR;
    P.maxops = 2 * P.maxLen + 8;
    P.dir_cap = (int64_t)(2 * P.maxLen + 1) * maxQ;
    int64_t b = 0;
    b += align_up(maxQ + 16, 16);                                  // q codes
    b += align_up(maxR + 16, 16);                                  // r codes
    b += align_up((int64_t)sizeof(int16_t) * (maxQ + 32), 16) * 2; // H, E
    b += align_up((int64_t)sizeof(int16_t) * (maxR + 16), 16);     // maxColumn
    b += align_up((int64_t)sizeof(int32_t) * (2 * P.maxLen + 32), 16) * 3; // hb eb hc
    b += align_up(P.dir_cap + 16, 16);                             // dir
    b += align_up(P.maxops, 16);                                   // ops
    b += align_up((int64_t)sizeof(int32_t) * P.maxops, 16);        // lens
    P.per_thread = b;
    return P;
}

__device__ __forceinline__ void sw_pool_carve(const SwPool& P, int64_t slot, int8_t*& q, int8_t*& r, SwScratch& S)
{
    unsigned char* p = P.base + slot * P.per_thread;
    q = (int8_t*)p;
    p += align_up(P.maxQ + 16, 16);
    r = (int8_t*)p;
    p += align_up(P.maxR + 16, 16);
    S.H = (int16_t*)p;
    p += align_up((int64_t)sizeof(int16_t) * (P.maxQ + 32), 16);
    S.E = (int16_t*)p;
    p += align_up((int64_t)sizeof(int16_t) * (P.maxQ + 32), 16);
    S.maxColumn = (int16_t*)p;
    p += align_up((int64_t)sizeof(int16_t) * (P.maxR + 16), 16);
    S.hb = (int32_t*)p;
    p += align_up((int64_t)sizeof(int32_t) * (2 * P.maxLen + 32), 16);
    S.eb = (int32_t*)p;
    p += align_up((int64_t)sizeof(int32_t) * (2 * P.maxLen + 32), 16);
    S.hc = (int32_t*)p;
    p += align_up((int64_t)sizeof(int32_t) * (2 * P.maxLen + 32), 16);
    S.dir = (uint8_t*)p;
    S.dir_cap = P.dir_cap;
    p += align_up(P.dir_cap + 16, 16);
    S.ops = (char*)p;
    p += align_up(P.maxops, 16);
    S.lens = (int32_t*)p;
    S.maxops = P.maxops;
}

// ---- function-level API: ASCII rows ------------------------------------------------------------
__global__ void __launch_bounds__(128) sw_rows_kernel(const char* __restrict__ queries, int64_t qpitch,
                                                      const int32_t* __restrict__ qlen,
                                                      const char* __restrict__ refs, int64_t rpitch,
                                                      const int32_t* __restrict__ rlen,
                                                      const int32_t* __restrict__ mask_len, int64_t n, SwPool P,
                                                      hrm_alignment* __restrict__ out, char* __restrict__ cigars,
                                                      int64_t cigar_pitch)
{
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nslots = (int64_t)gridDim.x * blockDim.x;
    int8_t *q, *r;
    SwScratch S;
    sw_pool_carve(P, slot, q, r, S);
    for (int64_t e = slot; e < n; e += nslots) {
        const int ql = qlen[e], rl = rlen[e];
        SwAlignment al;
        char* cig = cigars + e * cigar_pitch;
        if (ql < 0 || ql > P.maxQ || rl < 0 || rl > P.maxR) {
            al.sw_score = al.sw_score_next_best = al.ref_begin = al.ref_end = al.query_begin = al.query_end = 0;
            al.ref_end_next_best = al.mismatches = al.cigar_len = 0;
            al.flag = 1;
        } else {
            for (int i = 0; i < ql; i++) q[i] = sw_translate((unsigned char)queries[e * qpitch + i]);
            for (int i = 0; i < rl; i++) r[i] = sw_translate((unsigned char)refs[e * rpitch + i]);
            sw_align(q, ql, r, rl, mask_len[e], S, &al, cig, (int)cigar_pitch);
        }
        hrm_alignment o;
        o.sw_score = al.sw_score;
        o.sw_score_next_best = al.sw_score_next_best;
        o.ref_begin = al.ref_begin;
        o.ref_end = al.ref_end;
        o.query_begin = al.query_begin;
        o.query_end = al.query_end;
        o.ref_end_next_best = al.ref_end_next_best;
        o.mismatches = al.mismatches;
        o.flag = al.flag;
        o.cigar_len = al.cigar_len;
        out[e] = o;
        if (al.cigar_len < cigar_pitch) cig[al.cigar_len] = 0;
    }
}

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

// ---- fused verification: thread t = 2*read + a (a = 0: 3N(read'), a = 1: 3N(RC(read'))) -----------
__device__ __forceinline__ int conv_code(int c, int conv)
{
    if (conv == HRM_CONV_CT && c == 1) return 3;
    if (conv == HRM_CONV_GA && c == 2) return 0;
    return c;
}

__global__ void __launch_bounds__(128) verify_kernel(VerifyParams VP, const int32_t* __restrict__ read_len, int64_t n,
                                                     const hrm_mapped_read* __restrict__ mapped, SwPool P,
                                                     hrm_read_record* __restrict__ records, char* __restrict__ cigars,
                                                     int64_t cigar_pitch)
{
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nslots = (int64_t)gridDim.x * blockDim.x;
    int8_t *q, *r;
    SwScratch S;
    sw_pool_carve(P, slot, q, r, S);
    for (int64_t t = slot; t < 2 * n; t += nslots) {
        const int64_t rd = t >> 1;
        const int a = (int)(t & 1);
        const hrm_mapped_read m = mapped[rd];
        hrm_read_record* rec = records + rd;
        char* cig = cigars + t * cigar_pitch;
        const int L = read_len[rd];
        SwAlignment al;
        al.sw_score = al.sw_score_next_best = al.ref_begin = al.ref_end = al.query_begin = al.query_end = 0;
        al.ref_end_next_best = al.mismatches = al.flag = al.cigar_len = 0;
        int ed = -1, wl = 0;
        int maskLen = L / 2; // ref: mappinghandler.cu:453-454
        maskLen = maskLen < 15 ? 15 : maskLen;
        const bool ok = m.orientation != HRM_ORIENT_NONE && m.pass >= 0 && m.pass < VP.num_passes && L > 0 &&
                        L <= P.maxQ;
        if (ok) {
            const VerifyPass& VPp = VP.pass[m.pass];
            const uint32_t* rw = VPp.reads + rd * VPp.read_pitch;
            const int64_t clen = VPp.G.chrom_len[m.chromosome_id];
            const uint32_t* cw = VPp.G.chrom_words[m.chromosome_id];
            // ref: mappinghandler.cu:434-440 window length (strict <)
            wl = (int)((m.position + VP.w < clen) ? VP.w : clen - m.position);
            if (wl > P.maxR) wl = P.maxR;
            // readsequence = RC(read) if the SHD orientation was RC (:420-423); alignment 1 uses RC(readsequence)
            const bool rc = (m.orientation == HRM_ORIENT_REVCOMP) != (a == 1);
            for (int j = 0; j < L; j++) {
                const int c = rc ? 3 - (int)get_nuc(rw, L - 1 - j) : (int)get_nuc(rw, j);
                q[j] = (int8_t)conv_code(c, VPp.verify_conv);
            }
            for (int j = 0; j < wl; j++) r[j] = (int8_t)conv_code((int)get_nuc(cw, m.position + j), VPp.verify_conv);
            if (VP.mapper_type == HRM_MAPPER_SW) {
                sw_align(q, L, r, wl, maskLen, S, &al, cig, (int)cigar_pitch);
            } else {
                // edlib mode: global edit distance on the same strings (bytes "ACGT"[code])
                unsigned char* qa = (unsigned char*)S.dir;
                unsigned char* ta = qa + P.maxQ + 16;
                for (int j = 0; j < L; j++) qa[j] = (unsigned char)("ACGT"[q[j] & 3]);
                for (int j = 0; j < wl; j++) ta[j] = (unsigned char)("ACGT"[r[j] & 3]);
                ed = myers_nw(qa, L, ta, wl);
            }
        }
        if (al.cigar_len < cigar_pitch) cig[al.cigar_len] = 0;
        hrm_alignment o;
        o.sw_score = al.sw_score;
        o.sw_score_next_best = al.sw_score_next_best;
        o.ref_begin = al.ref_begin;
        o.ref_end = al.ref_end;
        o.query_begin = al.query_begin;
        o.query_end = al.query_end;
        o.ref_end_next_best = al.ref_end_next_best;
        o.mismatches = al.mismatches;
        o.flag = al.flag;
        o.cigar_len = al.cigar_len;
        rec->alignments[a] = o;
        rec->edit_distance[a] = ed;
        if (a == 0) {
            rec->mapped = m;
            rec->window_length = wl;
            rec->mask_len = maskLen;
        }
    }
}

static hrm_status sw_pool_make(SwPool& P, int64_t n_items, Scratch& mem, int& blocks, cudaStream_t s)
{
    // concurrency bounded by a scratch budget (B200: plenty of HBM, keep it modest anyway)
    const int64_t budget = 6LL << 30;
    int64_t slots = (int64_t)num_sms() * 128 * 2;
    if (slots > n_items) slots = n_items;
    if (slots * P.per_thread > budget) slots = budget / P.per_thread;
    if (slots < 128) slots = 128;
    blocks = (int)HRM_SDIV(slots, (int64_t)128);
    slots = (int64_t)blocks * 128;
    HRM_TRY(mem.alloc((size_t)(slots * P.per_thread), s));
    P.base = mem.as<unsigned char>();
    return HRM_OK;
}

hrm_status verify_reads(const VerifyParams& VP, const int32_t* d_read_len, int64_t n, int max_read_len,
                        const hrm_mapped_read* d_mapped, hrm_read_record* d_records, char* d_cigars, int64_t cigar_pitch,
                        cudaStream_t s)
{
    if (n == 0) return HRM_OK;
    SwPool P = sw_pool_layout(max_read_len > 16 ? max_read_len : 16, VP.w);
    Scratch mem;
    int blocks = 1;
    HRM_TRY(sw_pool_make(P, 2 * n, mem, blocks, s));
    HRM_LAUNCH(verify_kernel, blocks, 128, 0, s, VP, d_read_len, n, d_mapped, P, d_records, d_cigars, cigar_pitch);
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
    HRM_REQUIRE(query_pitch <= HRM_SW_MAX_QUERY + 16 && ref_pitch <= HRM_SW_MAX_REF + 16,
                "pitch exceeds HRM_SW_MAX_QUERY / HRM_SW_MAX_REF");
    if (n == 0) return HRM_OK;
    cudaStream_t s = as_stream(stream);
    SwPool P = sw_pool_layout((int)(query_pitch < HRM_SW_MAX_QUERY ? query_pitch : HRM_SW_MAX_QUERY),
                              (int)(ref_pitch < HRM_SW_MAX_REF ? ref_pitch : HRM_SW_MAX_REF));
    Scratch mem;
    int blocks = 1;
    HRM_TRY(sw_pool_make(P, n, mem, blocks, s));
    HRM_LAUNCH(sw_rows_kernel, blocks, 128, 0, s, d_queries, query_pitch, d_query_len, d_refs, ref_pitch, d_ref_len,
               d_mask_len, n, P, d_out, d_cigars, cigar_pitch);
    return HRM_OK;
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
