// ingest.cu -- read ingestion on the device: FASTQ / FASTA text -> ASCII rows ready for K1.
// ref: forEachReadInFile include/readlibraryio.hpp:288-326 (kseqpp parser, one host thread),
//      constructChunkedReadStorageFromFiles include/chunkedreadstorageconstruction.hpp:31-502: one parser thread
//      fills batches of fileParserMaxBatchsize = 65536 reads (:107), encoder threads run preprocessSequence (:70-95)
//      on every read of a batch -- a c g t -> upper case, every other character that is not A C G T is replaced by
//      "ACGT"[Ncount], Ncount = (Ncount + 1) % 4, with Ncount = 0 at the START OF EVERY BATCH (:277) -- and record
//      the ids of reads that contained such characters (:293-297).
// Here: the text of a chunk is in device memory; line starts come from two passes over the text (newlines per
// 1 KiB tile by byte-SIMD compare, scan of the tile counts, positions written tile by tile), a record
// is 4 lines (FASTQ, first byte '@') or a header line ('>') and every line up to the next header (FASTA), the cyclic
// replacement index of a read is an exclusive scan of the per-read counts of replaced characters, rebased at
// every 65536th read of the file.  One warp per read copies and normalises its sequence.
// Traffic: the text three times (count, positions, copy), rows once.
#include "runtime.cuh"

namespace hrm {

constexpr int INGEST_BATCH = 65536; // ref: fileParserMaxBatchsize chunkedreadstorageconstruction.hpp:107

__device__ __forceinline__ bool ingest_valid_base(unsigned char c)
{
    return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't';
}

constexpr int INGEST_TILE = 1024; // text bytes per warp-tile of the line index

// newlines of byte j of the 4 bytes in w, as 0x00 / 0xFF per byte
__device__ __forceinline__ uint32_t newline_mask4(uint32_t w) { return __vcmpeq4(w, 0x0A0A0A0Au); }

// bytes [t * TILE, (t + 1) * TILE) of the text: lane l owns bytes [32 l, 32 l + 32) of the tile
__device__ __forceinline__ int tile_lane_newlines(const char* __restrict__ text, int64_t n, int64_t tile, int lane,
                                                  uint32_t (&m)[8])
{
    const int64_t base = tile * INGEST_TILE + 32 * lane;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int64_t at = base + 4 * q;
        uint32_t w = 0u;
        if (at + 4 <= n) {
            w = *reinterpret_cast<const uint32_t*>(text + at); // the text buffer is 4-byte aligned (device allocation)
        } else {
            for (int j = 0; j < 4; j++)
                if (at + j < n) w |= (uint32_t)(unsigned char)text[at + j] << (8 * j);
        }
        m[q] = newline_mask4(w) & 0x01010101u;
        cnt += __popc(m[q]);
    }
    return cnt;
}

// pass 1: newlines per tile
__global__ void __launch_bounds__(256) tile_newlines_kernel(const char* __restrict__ text, int64_t n, int64_t ntiles,
                                                            int32_t* __restrict__ tile_count)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < ntiles; t += nwarps) {
        uint32_t m[8];
        const int c = __reduce_add_sync(0xffffffffu, tile_lane_newlines(text, n, t, lane, m));
        if (lane == 0) tile_count[t] = c;
    }
}

// pass 2: line_start[l + 1] = position after the l-th newline; line_start[0] = 0
__global__ void __launch_bounds__(256) line_starts_kernel(const char* __restrict__ text, int64_t n, int64_t ntiles,
                                                          const int32_t* __restrict__ tile_excl,
                                                          int64_t* __restrict__ line_start)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (warp0 == 0 && lane == 0) line_start[0] = 0;
    for (int64_t t = warp0; t < ntiles; t += nwarps) {
        uint32_t m[8];
        const int c = tile_lane_newlines(text, n, t, lane, m);
        int incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        int64_t line = (int64_t)tile_excl[t] + (incl - c); // newlines before this lane's bytes
        const int64_t base = t * INGEST_TILE + 32 * lane;
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (m[q] & (1u << (8 * j))) line_start[++line] = base + 4 * q + j + 1;
    }
}

// per read: sequence extent, length, number of characters that will be replaced; format errors
__global__ void __launch_bounds__(256) record_extents_kernel(const char* __restrict__ text, int64_t nbytes,
                                                             const int64_t* __restrict__ line_start, int64_t nlines,
                                                             int lines_per_record, int64_t nreads, int64_t pitch,
                                                             int64_t* __restrict__ seq_begin, int32_t* __restrict__ len,
                                                             int32_t* __restrict__ ninvalid, int* __restrict__ error)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const char head = lines_per_record == 4 ? '@' : '>';
    for (int64_t r = warp0; r < nreads; r += nwarps) {
        const int64_t l0 = r * lines_per_record;
        const int64_t hb = line_start[l0], sb = line_start[l0 + 1];
        int64_t se = (l0 + 2 <= nlines ? line_start[l0 + 2] - 1 : nbytes); // newline of the sequence line (or EOF)
        if (se > sb && text[se - 1] == '\r') se--;
        const int64_t L = se - sb;
        if (lane == 0) {
            if (text[hb] != head) atomicExch(error, 1);            // not a record start: multi-line or damaged input
            if (lines_per_record == 4 && (l0 + 2 >= nlines + 1 || text[line_start[l0 + 2]] != '+')) atomicExch(error, 1);
            if (L > pitch) atomicExch(error, 2);                    // row too short
            seq_begin[r] = sb;
            len[r] = (int32_t)L;
        }
        int bad = 0;
        for (int64_t i = sb + lane; i < se; i += 32) bad += !ingest_valid_base((unsigned char)text[i]);
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) ninvalid[r] = bad;
    }
}

__global__ void __launch_bounds__(256) copy_reads_kernel(const char* __restrict__ text,
                                                         const int64_t* __restrict__ seq_begin,
                                                         const int32_t* __restrict__ len,
                                                         const int32_t* __restrict__ ninvalid,
                                                         const int32_t* __restrict__ invalid_excl, int64_t nreads,
                                                         int64_t first_read_id, int carry, char* __restrict__ rows,
                                                         int64_t pitch, uint8_t* __restrict__ ambiguous)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < nreads; r += nwarps) {
        // replaced characters before this read inside its batch of 65536 reads of the FILE
        const int64_t in_batch = (first_read_id + r) % INGEST_BATCH;
        const int64_t batch_first = r - in_batch; // negative: the batch began in an earlier chunk (carry applies)
        int running = batch_first >= 0 ? invalid_excl[r] - invalid_excl[batch_first] : invalid_excl[r] + carry;
        const int64_t sb = seq_begin[r];
        const int L = len[r];
        char* row = rows + r * pitch;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            unsigned char c = i < L ? (unsigned char)text[sb + i] : (unsigned char)'A';
            const bool bad = !ingest_valid_base(c);
            const unsigned m = __ballot_sync(0xffffffffu, bad);
            if (bad) c = (unsigned char)("ACGT"[(running + __popc(m & ((1u << lane) - 1u))) & 3]);
            else if (c >= 'a') c = (unsigned char)(c - 32);
            if (i < L) row[i] = (char)c;
            running += __popc(m);
        }
        for (int64_t i = L + lane; i < pitch; i += 32) row[i] = 0;
        if (lane == 0 && ambiguous) ambiguous[r] = ninvalid[r] > 0 ? 1 : 0;
    }
}

// ---- FASTA with sequences over several lines (ref: kseqpp reads every line up to the next '>' into one sequence,
// include/kseqpp/kseqpp.hpp; readlibraryio.hpp:288-326 hands the joined sequence on) ---------------------------
// is_header[l] = 1 if line l starts with '>'
__global__ void __launch_bounds__(256) fasta_headers_kernel(const char* __restrict__ text, int64_t nbytes,
                                                            const int64_t* __restrict__ line_start, int64_t total_lines,
                                                            int32_t* __restrict__ is_header)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < total_lines; l += stride) {
        const int64_t b = line_start[l];
        is_header[l] = (b < nbytes && text[b] == '>') ? 1 : 0;
    }
}

// header_line[r] = line of the r-th header
__global__ void __launch_bounds__(256) fasta_record_lines_kernel(const int32_t* __restrict__ is_header,
                                                                 const int32_t* __restrict__ rec_excl, int64_t total_lines,
                                                                 int64_t* __restrict__ header_line)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < total_lines; l += stride)
        if (is_header[l]) header_line[rec_excl[l]] = l;
}

// text extent [b, e) of line l without its newline / carriage return
__device__ __forceinline__ void line_extent(const char* __restrict__ text, int64_t nbytes,
                                            const int64_t* __restrict__ line_start, int64_t nlines, int64_t l, int64_t& b,
                                            int64_t& e)
{
    b = line_start[l];
    e = l + 1 <= nlines ? line_start[l + 1] - 1 : nbytes;
    if (e > b && text[e - 1] == '\r') e--;
}

__global__ void __launch_bounds__(256) fasta_extents_kernel(const char* __restrict__ text, int64_t nbytes,
                                                            const int64_t* __restrict__ line_start, int64_t nlines,
                                                            int64_t total_lines, const int64_t* __restrict__ header_line,
                                                            int64_t nreads, int64_t pitch, int32_t* __restrict__ len,
                                                            int32_t* __restrict__ ninvalid, int* __restrict__ error)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < nreads; r += nwarps) {
        const int64_t l0 = header_line[r] + 1, l1 = r + 1 < nreads ? header_line[r + 1] : total_lines;
        int64_t L = 0;
        int bad = 0;
        for (int64_t l = l0; l < l1; l++) {
            int64_t b, e;
            line_extent(text, nbytes, line_start, nlines, l, b, e);
            L += e - b;
            for (int64_t i = b + lane; i < e; i += 32) bad += !ingest_valid_base((unsigned char)text[i]);
        }
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (L > pitch) atomicExch(error, 2);
            len[r] = (int32_t)(L > pitch ? pitch : L);
            ninvalid[r] = bad;
        }
    }
}

__global__ void __launch_bounds__(256) fasta_copy_kernel(const char* __restrict__ text, int64_t nbytes,
                                                         const int64_t* __restrict__ line_start, int64_t nlines,
                                                         int64_t total_lines, const int64_t* __restrict__ header_line,
                                                         const int32_t* __restrict__ ninvalid,
                                                         const int32_t* __restrict__ invalid_excl, int64_t nreads,
                                                         int64_t first_read_id, int carry, char* __restrict__ rows,
                                                         int64_t pitch, uint8_t* __restrict__ ambiguous)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < nreads; r += nwarps) {
        const int64_t in_batch = (first_read_id + r) % INGEST_BATCH;
        const int64_t batch_first = r - in_batch;
        int running = batch_first >= 0 ? invalid_excl[r] - invalid_excl[batch_first] : invalid_excl[r] + carry;
        const int64_t l0 = header_line[r] + 1, l1 = r + 1 < nreads ? header_line[r + 1] : total_lines;
        char* row = rows + r * pitch;
        int64_t at = 0;
        for (int64_t l = l0; l < l1; l++) {
            int64_t b, e;
            line_extent(text, nbytes, line_start, nlines, l, b, e);
            const int Ln = (int)(e - b);
            for (int i0 = 0; i0 < Ln; i0 += 32) {
                const int i = i0 + lane;
                unsigned char c = i < Ln ? (unsigned char)text[b + i] : (unsigned char)'A';
                const bool bad = !ingest_valid_base(c);
                const unsigned m = __ballot_sync(0xffffffffu, bad);
                if (bad) c = (unsigned char)("ACGT"[(running + __popc(m & ((1u << lane) - 1u))) & 3]);
                else if (c >= 'a') c = (unsigned char)(c - 32);
                if (i < Ln && at + i < pitch) row[at + i] = (char)c;
                running += __popc(m);
            }
            at += Ln;
        }
        for (int64_t i = at + lane; i < pitch; i += 32) row[i] = 0;
        if (lane == 0 && ambiguous) ambiguous[r] = ninvalid[r] > 0 ? 1 : 0;
    }
}

static unsigned igrid(int64_t items)
{
    int64_t g = HRM_SDIV(items, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

} // namespace hrm

using namespace hrm;

// replaced characters of the (unfinished) last batch of 65536 reads, for the next chunk of the same file
static hrm_status ingest_carry_out(const int32_t* d_invx, int64_t nreads, int64_t first_read_id, int32_t carry_replaced,
                                   int32_t* h_carry_replaced_out, cudaStream_t s)
{
    const int64_t last_in_batch = (first_read_id + nreads) % INGEST_BATCH; // reads of that batch inside this chunk
    int32_t tail[2] = {0, 0};
    const int64_t last_first = nreads - last_in_batch;
    HRM_CUDA(cudaMemcpyAsync(&tail[0], d_invx + nreads, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (last_first > 0) HRM_CUDA(cudaMemcpyAsync(&tail[1], d_invx + last_first, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    if (h_carry_replaced_out)
        *h_carry_replaced_out = last_in_batch == 0 ? 0 : (last_first >= 0 ? (tail[0] - tail[1]) & 3 : (tail[0] + carry_replaced) & 3);
    return HRM_OK;
}

static hrm_status ingest_fasta(const char* d_text, int64_t nbytes, int64_t ntiles, const int32_t* d_tile_excl, int64_t nlines,
                               int64_t total_lines, int64_t first_read_id, int32_t carry_replaced, char* d_rows,
                               int64_t pitch, int32_t* d_lengths, uint8_t* d_ambiguous, int64_t max_reads,
                               int64_t* h_num_reads, int32_t* h_carry_replaced_out, cudaStream_t s)
{
    Scratch lstart, ishdr, recx, tot, hline, ninv, invx, err;
    HRM_TRY(lstart.alloc(sizeof(int64_t) * ((size_t)nlines + 2), s));
    HRM_TRY(ishdr.alloc(sizeof(int32_t) * ((size_t)total_lines + 1), s));
    HRM_TRY(recx.alloc(sizeof(int32_t) * ((size_t)total_lines + 2), s));
    HRM_TRY(tot.alloc(sizeof(int64_t), s));
    HRM_LAUNCH(line_starts_kernel, igrid(ntiles * 32), 256, 0, s, d_text, nbytes, ntiles, d_tile_excl, lstart.as<int64_t>());
    HRM_LAUNCH(fasta_headers_kernel, igrid(total_lines), 256, 0, s, d_text, nbytes, lstart.as<int64_t>(), total_lines,
               ishdr.as<int32_t>());
    HRM_TRY(exclusive_scan_i32(ishdr.as<int32_t>(), recx.as<int32_t>(), total_lines, tot.as<int64_t>(), s));
    int64_t nreads = 0;
    HRM_CUDA(cudaMemcpyAsync(&nreads, tot.p, sizeof nreads, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    HRM_REQUIRE(nreads <= max_reads, "more reads in the text than max_reads");
    if (nreads == 0) return HRM_OK;
    HRM_TRY(hline.alloc(sizeof(int64_t) * (size_t)nreads, s));
    HRM_TRY(ninv.alloc(sizeof(int32_t) * (size_t)nreads, s));
    HRM_TRY(invx.alloc(sizeof(int32_t) * ((size_t)nreads + 1), s));
    HRM_TRY(err.alloc(sizeof(int) * 2, s));
    HRM_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int) * 2, s));
    HRM_LAUNCH(fasta_record_lines_kernel, igrid(total_lines), 256, 0, s, ishdr.as<int32_t>(), recx.as<int32_t>(), total_lines,
               hline.as<int64_t>());
    HRM_LAUNCH(fasta_extents_kernel, igrid(nreads * 32), 256, 0, s, d_text, nbytes, lstart.as<int64_t>(), nlines, total_lines,
               hline.as<int64_t>(), nreads, pitch, d_lengths, ninv.as<int32_t>(), err.as<int>());
    HRM_TRY(exclusive_scan_i32(ninv.as<int32_t>(), invx.as<int32_t>(), nreads, nullptr, s));
    int h_err = 0;
    HRM_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof h_err, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    if (h_err == 2) {
        set_error("a sequence is longer than the row pitch");
        return HRM_ERR_INVALID;
    }
    HRM_LAUNCH(fasta_copy_kernel, igrid(nreads * 32), 256, 0, s, d_text, nbytes, lstart.as<int64_t>(), nlines, total_lines,
               hline.as<int64_t>(), ninv.as<int32_t>(), invx.as<int32_t>(), nreads, first_read_id, carry_replaced & 3, d_rows,
               pitch, d_ambiguous);
    HRM_TRY(ingest_carry_out(invx.as<int32_t>(), nreads, first_read_id, carry_replaced, h_carry_replaced_out, s));
    *h_num_reads = nreads;
    return HRM_OK;
}

extern "C" hrm_status hrm_ingest_reads(const char* d_text, int64_t nbytes, int64_t first_read_id,
                                       int32_t carry_replaced, char* d_rows, int64_t pitch, int32_t* d_lengths,
                                       uint8_t* d_ambiguous, int64_t max_reads, int64_t* h_num_reads,
                                       int32_t* h_carry_replaced_out, hrm_stream stream)
{
    HRM_TRY(ensure_device());
    HRM_REQUIRE(nbytes >= 0 && nbytes < (1LL << 31), "text chunks are limited to 2 GiB");
    HRM_REQUIRE(pitch > 0 && max_reads >= 0 && h_num_reads != nullptr && first_read_id >= 0, "args");
    *h_num_reads = 0;
    if (h_carry_replaced_out) *h_carry_replaced_out = carry_replaced;
    if (nbytes == 0) return HRM_OK;
    HRM_REQUIRE(d_text != nullptr && d_rows != nullptr && d_lengths != nullptr, "buffers");
    cudaStream_t s = as_stream(stream);
    HRM_REQUIRE((reinterpret_cast<uintptr_t>(d_text) & 3) == 0, "d_text must be 4-byte aligned");
    Scratch tcount, texcl, tot;
    const int64_t ntiles = HRM_SDIV(nbytes, (int64_t)INGEST_TILE);
    HRM_TRY(tcount.alloc(sizeof(int32_t) * (size_t)ntiles, s));
    HRM_TRY(texcl.alloc(sizeof(int32_t) * ((size_t)ntiles + 1), s));
    HRM_TRY(tot.alloc(sizeof(int64_t) * 2, s));
    HRM_LAUNCH(tile_newlines_kernel, igrid(ntiles * 32), 256, 0, s, d_text, nbytes, ntiles, tcount.as<int32_t>());
    HRM_TRY(exclusive_scan_i32(tcount.as<int32_t>(), texcl.as<int32_t>(), ntiles, tot.as<int64_t>(), s));
    int64_t nlines = 0;
    char head[2] = {0, 0};
    HRM_CUDA(cudaMemcpyAsync(&nlines, tot.p, sizeof nlines, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaMemcpyAsync(head, d_text, 1, cudaMemcpyDeviceToHost, s));
    char last = 0;
    HRM_CUDA(cudaMemcpyAsync(&last, d_text + nbytes - 1, 1, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    HRM_REQUIRE(head[0] == '@' || head[0] == '>', "not FASTQ ('@') or FASTA ('>') text");
    const int lpr = head[0] == '@' ? 4 : 2;
    const int64_t total_lines = nlines + (last != '\n' ? 1 : 0); // a last line without newline still counts
    if (lpr == 2) // FASTA: records are delimited by their header lines, sequences may span any number of lines
        return ingest_fasta(d_text, nbytes, ntiles, texcl.as<int32_t>(), nlines, total_lines, first_read_id, carry_replaced,
                            d_rows, pitch, d_lengths, d_ambiguous, max_reads, h_num_reads, h_carry_replaced_out, s);
    HRM_REQUIRE(total_lines % lpr == 0, "incomplete FASTQ record");
    const int64_t nreads = total_lines / lpr;
    HRM_REQUIRE(nreads <= max_reads, "more reads in the text than max_reads");
    if (nreads == 0) return HRM_OK;
    Scratch lstart, sbeg, ninv, invx, err;
    HRM_TRY(lstart.alloc(sizeof(int64_t) * ((size_t)nlines + 2), s));
    HRM_TRY(sbeg.alloc(sizeof(int64_t) * (size_t)nreads, s));
    HRM_TRY(ninv.alloc(sizeof(int32_t) * (size_t)nreads, s));
    HRM_TRY(invx.alloc(sizeof(int32_t) * ((size_t)nreads + 1), s));
    HRM_TRY(err.alloc(sizeof(int) + sizeof(int64_t), s));
    HRM_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int) + sizeof(int64_t), s));
    HRM_LAUNCH(line_starts_kernel, igrid(ntiles * 32), 256, 0, s, d_text, nbytes, ntiles, texcl.as<int32_t>(),
               lstart.as<int64_t>());
    HRM_LAUNCH(record_extents_kernel, igrid(nreads * 32), 256, 0, s, d_text, nbytes, lstart.as<int64_t>(), nlines, lpr,
               nreads, pitch, sbeg.as<int64_t>(), d_lengths, ninv.as<int32_t>(), err.as<int>());
    HRM_TRY(exclusive_scan_i32(ninv.as<int32_t>(), invx.as<int32_t>(), nreads, nullptr, s));
    int h_err = 0;
    HRM_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof h_err, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    if (h_err == 1) {
        set_error("malformed record (header / '+' line missing, or multi-line sequences)");
        return HRM_ERR_INVALID;
    }
    if (h_err == 2) {
        set_error("a sequence is longer than the row pitch");
        return HRM_ERR_INVALID;
    }
    HRM_LAUNCH(copy_reads_kernel, igrid(nreads * 32), 256, 0, s, d_text, sbeg.as<int64_t>(), d_lengths, ninv.as<int32_t>(),
               invx.as<int32_t>(), nreads, first_read_id, carry_replaced & 3, d_rows, pitch, d_ambiguous);
    HRM_TRY(ingest_carry_out(invx.as<int32_t>(), nreads, first_read_id, carry_replaced, h_carry_replaced_out, s));
    *h_num_reads = nreads;
    return HRM_OK;
}

// ---- gzip'd read files (ref: kseqpp reads .gz through zlib's gzread, include/kseqpp/kseqpp.hpp; the reference's
// file reader accepts both, include/readlibraryio.hpp:288-326) -------------------------------------------------------
// Host-side inflate of a whole gzip stream (all members) into a host buffer whose text then goes to
// hrm_ingest_reads / hrm_mapper_stage_fastq.  *h_written = bytes produced; HRM_ERR_OVERFLOW when cap is too small
// (*h_written = bytes produced so far: grow and call again).  Decompression is I/O decoding, not part of the mapping path.
#include <zlib.h>
extern "C" hrm_status hrm_inflate_gzip(const void* h_in, int64_t nbytes, void* h_out, int64_t cap, int64_t* h_written)
{
    HRM_REQUIRE(h_in != nullptr && h_out != nullptr && h_written != nullptr && nbytes >= 0 && cap >= 0, "args");
    *h_written = 0;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) { // zlib or gzip header, detected
        set_error("inflateInit2 failed");
        return HRM_ERR_INVALID;
    }
    const unsigned char* in = (const unsigned char*)h_in;
    unsigned char* out = (unsigned char*)h_out;
    int64_t ipos = 0, opos = 0;
    hrm_status st = HRM_OK;
    bool ended = nbytes == 0; // the last member reached its end marker
    while (ipos < nbytes) {
        const int64_t ichunk = (nbytes - ipos) < (1LL << 30) ? (nbytes - ipos) : (1LL << 30);
        zs.next_in = const_cast<unsigned char*>(in + ipos);
        zs.avail_in = (uInt)ichunk;
        int rc = Z_OK;
        while (zs.avail_in > 0 && rc != Z_STREAM_END) {
            const int64_t ochunk = (cap - opos) < (1LL << 30) ? (cap - opos) : (1LL << 30);
            if (ochunk == 0) {
                st = HRM_ERR_OVERFLOW;
                break;
            }
            zs.next_out = out + opos;
            zs.avail_out = (uInt)ochunk;
            rc = inflate(&zs, Z_NO_FLUSH);
            opos += ochunk - (int64_t)zs.avail_out;
            if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) {
                set_error("inflate failed: %s", zs.msg ? zs.msg : "corrupt gzip data");
                st = HRM_ERR_INVALID;
                break;
            }
        }
        ipos += ichunk - (int64_t)zs.avail_in;
        if (st != HRM_OK) break;
        ended = rc == Z_STREAM_END;
        if (ended && ipos < nbytes) inflateReset(&zs); // next member of a multi-member file
    }
    inflateEnd(&zs);
    *h_written = opos;
    if (st == HRM_OK && !ended) {
        set_error("truncated gzip stream");
        return HRM_ERR_INVALID;
    }
    if (st == HRM_ERR_OVERFLOW) set_error("output buffer too small for the inflated text");
    return st;
}
