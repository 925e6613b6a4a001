// store.cu -- device-resident genome (S2) and read storage (S1).
// ref: struct Genome include/genome.hpp:84-446 (host std::map<int,std::string> + a whole
//      reverse-complement copy in RAM, ASCII slices shipped H2D per batch main_gpu.cu:642-656);
//      GpuReadStorage include/gpu/gpureadstorage.cuh:22-119, MultiGpuReadStorage
//      include/gpu/multigpureadstorage.cuh:655-905,1515-1532 (row gathers by read id).
// Here the genome is packed once (K1) and stays in HBM: 2 bits per base, 0.78 GB for 3.1 Gbp.
#include "store.cuh"
#include <string.h>

extern "C" hrm_status hrm_encode_2bit_contiguous(const char*, int64_t, int, uint32_t*, hrm_stream);
extern "C" hrm_status hrm_encode_2bit(const char*, int64_t, const int32_t*, int64_t, int, uint32_t*, int64_t, hrm_stream);

namespace hrm {

__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ rows, int64_t pitch,
                                                          const uint32_t* __restrict__ ids, uint32_t first_id,
                                                          int64_t n, int64_t nrows, uint32_t* __restrict__ out,
                                                          int64_t out_pitch)
{
    const int64_t total = n * out_pitch;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t e = t / out_pitch;
        const int wi = (int)(t - e * out_pitch);
        const int64_t id = ids ? (int64_t)ids[e] : (int64_t)first_id + e;
        out[t] = (wi < pitch && id < nrows) ? rows[id * pitch + wi] : 0u;
    }
}

__global__ void __launch_bounds__(256) gather_lengths_kernel(const int32_t* __restrict__ lengths,
                                                             const uint32_t* __restrict__ ids, int64_t n,
                                                             int64_t nrows, int32_t* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t id = ids[e];
        out[e] = id < nrows ? lengths[id] : 0;
    }
}

// per read: 1 if any of its first len characters is not A/C/G/T (ref: the reads ChunkedReadStorage records as
// ambiguous, chunkedreadstorageconstruction.hpp:70-95); thread per read over 16-byte pieces
__global__ void __launch_bounds__(256) flag_ambiguous_kernel(const char* __restrict__ ascii, int64_t pitch,
                                                             const int32_t* __restrict__ lengths, int64_t n,
                                                             uint8_t* __restrict__ flags,
                                                             unsigned long long* __restrict__ count)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int len = lengths[r];
        const uint4* row = reinterpret_cast<const uint4*>(ascii + r * pitch);
        bool bad = false;
        for (int c = 0; c * 16 < len && !bad; c++) {
            const uint4 v = row[c];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int t = c * 16 + j;
                const unsigned ch = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                if (t < len && ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') bad = true;
            }
        }
        flags[r] = bad ? 1 : 0;
        local += bad ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

__global__ void __launch_bounds__(256) gather_flags_kernel(const uint8_t* __restrict__ flags,
                                                           const uint32_t* __restrict__ ids, int64_t n, int64_t nrows,
                                                           uint8_t* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t id = ids[e];
        out[e] = (flags && id < nrows) ? flags[id] : 0;
    }
}

static unsigned sgrid(int64_t items)
{
    int64_t g = HRM_SDIV(items, (int64_t)256);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

} // namespace hrm

using namespace hrm;

// ------------------------------------------------------------------------------------------ genome
extern "C" hrm_status hrm_genome_create_from_ascii(hrm_genome** out, const char* h_ascii,
                                                   const int64_t* h_chrom_offsets, int n_chrom, int conversion,
                                                   hrm_stream stream)
{
    HRM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HRM_TRY(ensure_device());
    HRM_REQUIRE(h_ascii != nullptr && h_chrom_offsets != nullptr && n_chrom >= 1, "genome");
    HRM_REQUIRE(conversion >= 0 && conversion <= 2, "conversion");
    cudaStream_t s = as_stream(stream);
    auto g = std::unique_ptr<hrm_genome>(new hrm_genome);
    g->n_chrom = n_chrom;
    g->conversion = conversion;
    int64_t words = 0;
    std::vector<int64_t> word_off(n_chrom);
    for (int c = 0; c < n_chrom; c++) {
        const int64_t len = h_chrom_offsets[c + 1] - h_chrom_offsets[c];
        HRM_REQUIRE(len >= 0 && len < (1LL << 31), "chromosome length must be < 2^31 (ref: genome.hpp:92)");
        g->chrom_len.push_back(len);
        word_off[c] = words;
        words += (len + 15) / 16 + 4; // 4 words of zero padding keep stream reads in bounds
    }
    g->total_words = words;
    HRM_CUDA(cudaMalloc(&g->words, sizeof(uint32_t) * (size_t)words));
    HRM_CUDA(cudaMemsetAsync(g->words, 0, sizeof(uint32_t) * (size_t)words, s));
    // ship the ASCII through a bounded staging buffer and pack on the device
    const int64_t CHUNK = 64LL << 20; // bases per chunk (multiple of 16)
    char* d_stage = nullptr;
    HRM_CUDA(cudaMalloc(&d_stage, (size_t)CHUNK));
    hrm_status st = HRM_OK;
    for (int c = 0; c < n_chrom && st == HRM_OK; c++) {
        g->chrom_words.push_back(g->words + word_off[c]);
        const char* src = h_ascii + h_chrom_offsets[c];
        const int64_t len = g->chrom_len[c];
        for (int64_t at = 0; at < len && st == HRM_OK; at += CHUNK) {
            const int64_t m = (len - at) < CHUNK ? (len - at) : CHUNK;
            if (cudaMemcpyAsync(d_stage, src + at, (size_t)m, cudaMemcpyHostToDevice, s) != cudaSuccess) {
                set_error("H2D copy of the genome failed: %s", cudaGetErrorString(cudaGetLastError()));
                st = HRM_ERR_CUDA;
                break;
            }
            st = hrm_encode_2bit_contiguous(d_stage, m, conversion, g->chrom_words[c] + at / 16, stream);
            if (st == HRM_OK && cudaStreamSynchronize(s) != cudaSuccess) { // staging buffer is reused
                set_error("genome packing failed: %s", cudaGetErrorString(cudaGetLastError()));
                st = HRM_ERR_CUDA;
            }
        }
    }
    cudaFree(d_stage);
    if (st != HRM_OK) {
        cudaFree(g->words);
        return st;
    }
    HRM_CUDA(cudaMalloc(&g->d_chrom_words, sizeof(uint32_t*) * (size_t)n_chrom));
    HRM_CUDA(cudaMalloc(&g->d_chrom_len, sizeof(int64_t) * (size_t)n_chrom));
    HRM_CUDA(cudaMemcpy(g->d_chrom_words, g->chrom_words.data(), sizeof(uint32_t*) * (size_t)n_chrom,
                        cudaMemcpyHostToDevice));
    HRM_CUDA(cudaMemcpy(g->d_chrom_len, g->chrom_len.data(), sizeof(int64_t) * (size_t)n_chrom, cudaMemcpyHostToDevice));
    *out = g.release();
    return HRM_OK;
}

extern "C" void hrm_genome_destroy(hrm_genome* g)
{
    if (!g) return;
    if (g->words) cudaFree(g->words);
    if (g->d_chrom_words) cudaFree(g->d_chrom_words);
    if (g->d_chrom_len) cudaFree(g->d_chrom_len);
    delete g;
}

extern "C" int hrm_genome_num_chromosomes(const hrm_genome* g) { return g ? g->n_chrom : 0; }
extern "C" int64_t hrm_genome_chromosome_length(const hrm_genome* g, int chrom)
{
    return (g && chrom >= 0 && chrom < g->n_chrom) ? g->chrom_len[chrom] : -1;
}
extern "C" int64_t hrm_genome_num_windows_in_chromosome(const hrm_genome* g, int chrom, int k, int w)
{
    if (!g || chrom < 0 || chrom >= g->n_chrom || w < k || k < 1) return -1;
    const int64_t stride = w - k + 1;
    return (g->chrom_len[chrom] + stride - 1) / stride; // ref: genome.hpp:176-183
}
extern "C" int64_t hrm_genome_num_windows(const hrm_genome* g, int k, int w)
{
    if (!g) return -1;
    int64_t t = 0;
    for (int c = 0; c < g->n_chrom; c++) t += hrm_genome_num_windows_in_chromosome(g, c, k, w);
    return t;
}
extern "C" const uint32_t* hrm_genome_chromosome_2bit(const hrm_genome* g, int chrom)
{
    return (g && chrom >= 0 && chrom < g->n_chrom) ? g->chrom_words[chrom] : nullptr;
}
extern "C" hrm_status hrm_genome_window_info(const hrm_genome* g, int k, int w, int64_t gw, int32_t* chrom,
                                             int64_t* window_id, int64_t* position, int32_t* length)
{
    HRM_REQUIRE(g != nullptr && w >= k && k >= 1 && gw >= 0, "args");
    const int64_t stride = w - k + 1;
    int64_t base = 0;
    for (int c = 0; c < g->n_chrom; c++) {
        const int64_t nw = (g->chrom_len[c] + stride - 1) / stride;
        if (gw < base + nw) {
            const int64_t id = gw - base, pos = id * stride;
            if (chrom) *chrom = c;
            if (window_id) *window_id = id;
            if (position) *position = pos;
            if (length) *length = (int32_t)((pos + w <= g->chrom_len[c]) ? w : g->chrom_len[c] - pos);
            return HRM_OK;
        }
        base += nw;
    }
    set_error("window id out of range");
    return HRM_ERR_INVALID;
}

// -------------------------------------------------------------------------------------- read store
static hrm_status readstore_finish(hrm_readstore* rs, const int32_t* h_lengths)
{
    rs->len_min = 0;
    rs->len_max = 0;
    if (rs->n > 0 && h_lengths) {
        int32_t lo = h_lengths[0], hi = h_lengths[0];
        for (int64_t i = 1; i < rs->n; i++) {
            lo = h_lengths[i] < lo ? h_lengths[i] : lo;
            hi = h_lengths[i] > hi ? h_lengths[i] : hi;
        }
        rs->len_min = lo;
        rs->len_max = hi;
    }
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_create_from_ascii(hrm_readstore** out, const char* h_ascii, int64_t ascii_pitch,
                                                      const int32_t* h_lengths, int64_t n, int conversion,
                                                      hrm_stream stream)
{
    HRM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && ascii_pitch > 0 && ascii_pitch % 16 == 0, "ascii_pitch must be a positive multiple of 16");
    HRM_REQUIRE(n < (1LL << 32), "read ids are 32 bit (ref: config.hpp:8)");
    cudaStream_t s = as_stream(stream);
    auto rs = std::unique_ptr<hrm_readstore>(new hrm_readstore);
    rs->n = n;
    int32_t maxlen = 0;
    for (int64_t i = 0; i < n; i++) {
        HRM_REQUIRE(h_lengths[i] >= 0 && h_lengths[i] <= ascii_pitch, "read length exceeds pitch");
        maxlen = h_lengths[i] > maxlen ? h_lengths[i] : maxlen;
    }
    rs->pitch_words = (maxlen + 15) / 16 > 0 ? (maxlen + 15) / 16 : 1;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    HRM_CUDA(cudaMalloc(&rs->rows, sizeof(uint32_t) * nn * (size_t)rs->pitch_words));
    HRM_CUDA(cudaMalloc(&rs->lengths, sizeof(int32_t) * nn));
    HRM_CUDA(cudaMalloc(&rs->ambig, nn + 8));
    unsigned long long* d_count = nullptr;
    HRM_CUDA(cudaMalloc(&d_count, sizeof(unsigned long long)));
    HRM_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
    if (n > 0) {
        HRM_CUDA(cudaMemcpyAsync(rs->lengths, h_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        const int64_t CH = 1 << 20; // reads per staging chunk
        char* d_stage = nullptr;
        const int64_t chunk_reads = n < CH ? n : CH;
        HRM_CUDA(cudaMalloc(&d_stage, (size_t)(chunk_reads * ascii_pitch)));
        hrm_status st = HRM_OK;
        for (int64_t at = 0; at < n && st == HRM_OK; at += chunk_reads) {
            const int64_t m = (n - at) < chunk_reads ? (n - at) : chunk_reads;
            if (cudaMemcpyAsync(d_stage, h_ascii + at * ascii_pitch, (size_t)(m * ascii_pitch), cudaMemcpyHostToDevice,
                                s) != cudaSuccess) {
                set_error("H2D copy of reads failed");
                st = HRM_ERR_CUDA;
                break;
            }
            flag_ambiguous_kernel<<<sgrid(m), 256, 0, s>>>(d_stage, ascii_pitch, rs->lengths + at, m, rs->ambig + at, d_count);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            st = hrm_encode_2bit(d_stage, ascii_pitch, rs->lengths + at, m, conversion,
                                 rs->rows + at * rs->pitch_words, rs->pitch_words, stream);
            if (st == HRM_OK && cudaStreamSynchronize(s) != cudaSuccess) {
                set_error("read packing failed");
                st = HRM_ERR_CUDA;
            }
        }
        cudaFree(d_stage);
        if (st != HRM_OK) {
            cudaFree(rs->rows);
            cudaFree(rs->lengths);
            cudaFree(rs->ambig);
            cudaFree(d_count);
            return st;
        }
    }
    unsigned long long with_n = 0;
    HRM_CUDA(cudaMemcpyAsync(&with_n, d_count, sizeof with_n, cudaMemcpyDeviceToHost, s));
    HRM_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_count);
    rs->with_n = (int64_t)with_n;
    readstore_finish(rs.get(), h_lengths);
    *out = rs.release();
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_create_from_2bit(hrm_readstore** out, const uint32_t* d_seq2bit,
                                                     int64_t pitch_words, const int32_t* d_lengths, int64_t n,
                                                     hrm_stream stream)
{
    HRM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    HRM_TRY(ensure_device());
    HRM_REQUIRE(n >= 0 && pitch_words > 0 && n < (1LL << 32), "sizes");
    cudaStream_t s = as_stream(stream);
    auto rs = std::unique_ptr<hrm_readstore>(new hrm_readstore);
    rs->n = n;
    rs->pitch_words = pitch_words;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    HRM_CUDA(cudaMalloc(&rs->rows, sizeof(uint32_t) * nn * (size_t)pitch_words));
    HRM_CUDA(cudaMalloc(&rs->lengths, sizeof(int32_t) * nn));
    std::vector<int32_t> h_len((size_t)n);
    if (n > 0) {
        HRM_CUDA(cudaMemcpyAsync(rs->rows, d_seq2bit, sizeof(uint32_t) * (size_t)n * pitch_words,
                                 cudaMemcpyDeviceToDevice, s));
        HRM_CUDA(cudaMemcpyAsync(rs->lengths, d_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        HRM_CUDA(cudaMemcpyAsync(h_len.data(), d_lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        HRM_CUDA(cudaStreamSynchronize(s));
    }
    readstore_finish(rs.get(), h_len.data());
    *out = rs.release();
    return HRM_OK;
}

extern "C" void hrm_readstore_destroy(hrm_readstore* rs)
{
    if (!rs) return;
    if (rs->rows) cudaFree(rs->rows);
    if (rs->lengths) cudaFree(rs->lengths);
    if (rs->ambig) cudaFree(rs->ambig);
    delete rs;
}

extern "C" int hrm_readstore_handle_create(hrm_readstore* rs)
{
    if (!rs) return HRM_ERR_INVALID;
    std::lock_guard<std::mutex> lk(rs->mtx);
    rs->handles.push_back(true);
    return (int)rs->handles.size() - 1;
}

extern "C" hrm_status hrm_readstore_handle_destroy(hrm_readstore* rs, int handle)
{
    HRM_REQUIRE(rs != nullptr, "readstore");
    std::lock_guard<std::mutex> lk(rs->mtx);
    HRM_REQUIRE(handle >= 0 && handle < (int)rs->handles.size() && rs->handles[handle], "handle");
    rs->handles[handle] = false;
    return HRM_OK;
}

static bool rs_handle_ok(const hrm_readstore* rs, int handle)
{
    auto* m = const_cast<hrm_readstore*>(rs);
    std::lock_guard<std::mutex> lk(m->mtx);
    return handle >= 0 && handle < (int)rs->handles.size() && rs->handles[handle];
}

extern "C" hrm_status hrm_readstore_gather(const hrm_readstore* rs, int handle, uint32_t* d_out,
                                           int64_t out_pitch_words, const uint32_t* d_ids, int64_t n,
                                           hrm_stream stream)
{
    HRM_REQUIRE(rs != nullptr && rs_handle_ok(rs, handle), "readstore/handle");
    HRM_REQUIRE(n >= 0 && out_pitch_words > 0, "args");
    if (n == 0) return HRM_OK;
    HRM_REQUIRE(d_ids != nullptr && d_out != nullptr, "buffers");
    HRM_LAUNCH(gather_rows_kernel, sgrid(n * out_pitch_words), 256, 0, as_stream(stream), rs->rows, rs->pitch_words,
               d_ids, 0u, n, rs->n, d_out, out_pitch_words);
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_gather_contiguous(const hrm_readstore* rs, int handle, uint32_t* d_out,
                                                      int64_t out_pitch_words, uint32_t first_id, int64_t n,
                                                      hrm_stream stream)
{
    HRM_REQUIRE(rs != nullptr && rs_handle_ok(rs, handle), "readstore/handle");
    HRM_REQUIRE(n >= 0 && out_pitch_words > 0, "args");
    if (n == 0) return HRM_OK;
    HRM_LAUNCH(gather_rows_kernel, sgrid(n * out_pitch_words), 256, 0, as_stream(stream), rs->rows, rs->pitch_words,
               (const uint32_t*)nullptr, first_id, n, rs->n, d_out, out_pitch_words);
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_gather_lengths(const hrm_readstore* rs, int handle, int32_t* d_lengths,
                                                   const uint32_t* d_ids, int64_t n, hrm_stream stream)
{
    HRM_REQUIRE(rs != nullptr && rs_handle_ok(rs, handle), "readstore/handle");
    HRM_REQUIRE(n >= 0, "args");
    if (n == 0) return HRM_OK;
    HRM_REQUIRE(d_ids != nullptr && d_lengths != nullptr, "buffers");
    HRM_LAUNCH(gather_lengths_kernel, sgrid(n), 256, 0, as_stream(stream), rs->lengths, d_ids, n, rs->n, d_lengths);
    return HRM_OK;
}

// ref: GpuReadStorage::areSequencesAmbiguous gpureadstorage.cuh:31-37 (impl multigpureadstorage.cuh:623-653)
extern "C" hrm_status hrm_readstore_are_ambiguous(const hrm_readstore* rs, int handle, uint8_t* d_result,
                                                  const uint32_t* d_ids, int64_t n, hrm_stream stream)
{
    HRM_REQUIRE(rs != nullptr && rs_handle_ok(rs, handle), "readstore/handle");
    HRM_REQUIRE(n >= 0, "args");
    if (n == 0) return HRM_OK;
    HRM_REQUIRE(d_ids != nullptr && d_result != nullptr, "buffers");
    HRM_LAUNCH(gather_flags_kernel, sgrid(n), 256, 0, as_stream(stream), rs->ambig, d_ids, n, rs->n, d_result);
    return HRM_OK;
}

// carries per-read ambiguity flags (hrm_ingest_reads' d_ambiguous) into a store made from packed rows
extern "C" hrm_status hrm_readstore_set_ambiguous(hrm_readstore* rs, const uint8_t* d_flags, hrm_stream stream)
{
    HRM_REQUIRE(rs != nullptr && d_flags != nullptr, "args");
    cudaStream_t s = as_stream(stream);
    const size_t nn = (size_t)(rs->n > 0 ? rs->n : 1);
    if (!rs->ambig) HRM_CUDA(cudaMalloc(&rs->ambig, nn + 8));
    std::vector<uint8_t> h((size_t)rs->n);
    if (rs->n > 0) {
        HRM_CUDA(cudaMemcpyAsync(rs->ambig, d_flags, (size_t)rs->n, cudaMemcpyDeviceToDevice, s));
        HRM_CUDA(cudaMemcpyAsync(h.data(), d_flags, (size_t)rs->n, cudaMemcpyDeviceToHost, s));
        HRM_CUDA(cudaStreamSynchronize(s));
    }
    int64_t c = 0;
    for (uint8_t f : h) c += f != 0;
    rs->with_n = c;
    return HRM_OK;
}

// ref: GpuReadStorage::getIdsOfAmbiguousReads gpureadstorage.cuh:89-91: the ids, ascending, into h_ids
// (getNumberOfReadsWithN() entries = hrm_readstore_info().num_reads_with_n)
extern "C" hrm_status hrm_readstore_ambiguous_ids(const hrm_readstore* rs, uint32_t* h_ids)
{
    HRM_REQUIRE(rs != nullptr && h_ids != nullptr, "args");
    if (rs->n == 0 || !rs->ambig || rs->with_n == 0) return HRM_OK;
    std::vector<uint8_t> h((size_t)rs->n);
    HRM_CUDA(cudaMemcpy(h.data(), rs->ambig, (size_t)rs->n, cudaMemcpyDeviceToHost));
    int64_t at = 0;
    for (int64_t i = 0; i < rs->n; i++)
        if (h[(size_t)i]) h_ids[at++] = (uint32_t)i;
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_info(const hrm_readstore* rs, hrm_readstore_info_t* out)
{
    HRM_REQUIRE(rs != nullptr && out != nullptr, "args");
    memset(out, 0, sizeof *out);
    out->num_reads = rs->n;
    out->length_lower_bound = rs->len_min;
    out->length_upper_bound = rs->len_max;
    out->num_reads_with_n = rs->with_n;
    out->pitch_words = (int32_t)rs->pitch_words;
    out->is_paired_end = 0;
    out->device_bytes = rs->n * (rs->pitch_words * 4 + 4);
    return HRM_OK;
}

// ---- ChunkedReadStorage dump format (--save-preprocessedreads-to / --load-preprocessedreads-from) -------
// ref: ChunkedReadStorage::saveToFile include/chunkedreadstorage.hpp:246-400, loadFromFile :160-243,
//      LengthStore<uint32_t>::writeToStream / readFromStream include/lengthstorage.hpp:164-204 (lengths minus the
//      minimum, bitsPerLength bits each, MSB first in 32-bit words, :115-160, :215-290).
// Layout: u64 numReads | i32 lengthUpperBound | i32 lengthLowerBound | u8 hasQualities | i32 qualityBits |
//         u64 lengthsBytes, sequencesBytes, qualitiesBytes, ambigBytes | LengthStore | u64 pitchInInts, u64 numInts,
//         u32 rows[numReads * pitch] | (no qualities: u64 0, u64 0) | u64 numAmbiguous, u32 ids[].
// The packed rows are the reference's own 2-bit layout, so a dump written here loads in the reference and back.
namespace {
struct ByteWriter {
    char* p;
    int64_t cap, at = 0;
    void put(const void* src, size_t n)
    {
        if (p && at + (int64_t)n <= cap) memcpy(p + at, src, n);
        at += (int64_t)n;
    }
    template <class T> void val(T v) { put(&v, sizeof v); }
};
int length_bits(int diff) // ref: LengthStore(int, int, int64) lengthstorage.hpp:24-55
{
    if (diff == 0) return 0;
    uint32_t n = (uint32_t)diff;
    n |= n >> 1;
    n |= n >> 2;
    n |= n >> 4;
    n |= n >> 8;
    n |= n >> 16;
    n += 1;
    int b = 0;
    while ((1u << b) < n) b++;
    return b;
}
} // namespace

extern "C" hrm_status hrm_readstore_write_reference_format(const hrm_readstore* rs, void* h_buf, int64_t* h_size)
{
    HRM_REQUIRE(rs != nullptr && h_size != nullptr, "args");
    const int64_t n = rs->n, pitch = rs->pitch_words;
    const int bits = length_bits(rs->len_max - rs->len_min);
    const uint64_t len_words = ((uint64_t)bits * (uint64_t)n + 31) / 32;
    const uint64_t lengthsBytes = 4 * 4 + 4 + 8 + 8 + 8 + len_words * 4;
    const uint64_t seqBytes = 8 + 8 + (uint64_t)n * pitch * 4;
    const uint64_t qualBytes = 16;
    const uint64_t ambigBytes = 8 + (uint64_t)rs->with_n * 4;
    const int64_t total = 8 + 4 + 4 + 1 + 4 + 32 + (int64_t)(lengthsBytes + seqBytes + qualBytes + ambigBytes);
    if (!h_buf) {
        *h_size = total;
        return HRM_OK;
    }
    HRM_REQUIRE(*h_size >= total, "buffer too small");
    ByteWriter w{(char*)h_buf, *h_size};
    w.val<uint64_t>((uint64_t)n);
    w.val<int32_t>(rs->len_max); // saveToFile writes the upper bound first (:252-259)
    w.val<int32_t>(rs->len_min);
    w.val<uint8_t>(0);
    w.val<int32_t>(8);
    w.val<uint64_t>(lengthsBytes);
    w.val<uint64_t>(seqBytes);
    w.val<uint64_t>(qualBytes);
    w.val<uint64_t>(ambigBytes);
    // LengthStore
    std::vector<int32_t> lens((size_t)n);
    if (n > 0) HRM_CUDA(cudaMemcpy(lens.data(), rs->lengths, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> packed((size_t)len_words, 0u);
    for (int64_t i = 0; bits > 0 && i < n; i++) { // MSB first
        const uint32_t v = (uint32_t)(lens[(size_t)i] - rs->len_min);
        const uint64_t first = (uint64_t)bits * (uint64_t)i;
        for (int b = 0; b < bits; b++) {
            const uint64_t pos = first + (uint64_t)b;
            if ((v >> (bits - 1 - b)) & 1u) packed[(size_t)(pos >> 5)] |= 1u << (31 - (int)(pos & 31));
        }
    }
    w.val<int32_t>(32);
    w.val<int32_t>(bits);
    w.val<int32_t>(rs->len_min);
    w.val<int32_t>(rs->len_max);
    w.val<uint32_t>(bits > 0 ? (uint32_t)((1ull << bits) - 1) : 0u);
    w.val<int64_t>(n);
    w.val<uint64_t>(len_words);
    w.val<uint64_t>(len_words * 4);
    w.put(packed.data(), (size_t)len_words * 4);
    // sequences: straight from the device into the buffer
    w.val<uint64_t>((uint64_t)pitch);
    w.val<uint64_t>((uint64_t)n * pitch);
    if (n > 0) HRM_CUDA(cudaMemcpy((char*)h_buf + w.at, rs->rows, sizeof(uint32_t) * (size_t)(n * pitch), cudaMemcpyDeviceToHost));
    w.at += (int64_t)sizeof(uint32_t) * n * pitch;
    w.val<uint64_t>(0); // no quality scores: pitch 0, 0 elements (:330-373)
    w.val<uint64_t>(0);
    std::vector<uint32_t> ids((size_t)rs->with_n);
    if (rs->with_n > 0) HRM_TRY(hrm_readstore_ambiguous_ids(rs, ids.data()));
    w.val<uint64_t>((uint64_t)rs->with_n);
    w.put(ids.data(), ids.size() * 4);
    *h_size = w.at;
    return HRM_OK;
}

extern "C" hrm_status hrm_readstore_read_reference_format(hrm_readstore** out, const void* h_buf, int64_t size,
                                                          hrm_stream stream)
{
    HRM_REQUIRE(out != nullptr && h_buf != nullptr && size >= 53, "args");
    *out = nullptr;
    HRM_TRY(ensure_device());
    const char* p = (const char*)h_buf;
    const char* end = p + size;
    auto need = [&](uint64_t n) { return (uint64_t)(end - p) >= n; };
    auto rd64 = [&]() { uint64_t v; memcpy(&v, p, 8); p += 8; return v; };
    auto rd32 = [&]() { int32_t v; memcpy(&v, p, 4); p += 4; return v; };
    const uint64_t n = rd64();
    rd32(); // the two length bounds (written upper first, read lower first by the reference: unused)
    rd32();
    p += 1; // hasQualities
    rd32(); // quality bits
    const uint64_t lengthsBytes = rd64(), seqBytes = rd64(), qualBytes = rd64(), ambigBytes = rd64();
    HRM_REQUIRE(n < (1ull << 32), "read ids are 32 bit");
    HRM_REQUIRE(need(lengthsBytes) && lengthsBytes >= 44, "truncated length section");
    const char* after_len = p + lengthsBytes;
    const int dtb = rd32(), bits = rd32(), minL = rd32(), maxL = rd32();
    p += 4; // bitsMask
    const uint64_t ne = rd64(), rawElems = rd64(), rawBytes = rd64();
    HRM_REQUIRE(dtb == 32 && bits >= 0 && bits <= 31 && minL >= 0 && minL <= maxL && ne == n && rawBytes == rawElems * 4 &&
                    rawBytes <= lengthsBytes - 44 && rawElems >= ((uint64_t)bits * n + 31) / 32,
                "length store");
    std::vector<int32_t> lens((size_t)n, minL);
    for (uint64_t i = 0; bits > 0 && i < n; i++) {
        uint32_t v = 0;
        const uint64_t first = (uint64_t)bits * i;
        for (int b = 0; b < bits; b++) {
            const uint64_t pos = first + (uint64_t)b;
            uint32_t wv;
            memcpy(&wv, p + (pos >> 5) * 4, 4);
            v = (v << 1) | ((wv >> (31 - (int)(pos & 31))) & 1u);
        }
        lens[(size_t)i] = minL + (int32_t)v;
        HRM_REQUIRE(lens[(size_t)i] <= maxL, "length beyond the stored maximum");
    }
    p = after_len;
    HRM_REQUIRE(need(seqBytes) && seqBytes >= 16, "truncated sequence section");
    const uint64_t pitch = rd64(), numInts = rd64();
    HRM_REQUIRE(pitch >= 1 && pitch <= 4096 && numInts == n * pitch && seqBytes == 16 + numInts * 4, "sequence section");
    HRM_REQUIRE(pitch * 16 >= (uint64_t)maxL, "pitch too small for the longest read");
    const char* seq = p;
    p += numInts * 4;
    HRM_REQUIRE(need(qualBytes), "truncated quality section");
    p += qualBytes;
    HRM_REQUIRE(need(8), "truncated ambiguity section");
    const uint64_t na = rd64();
    HRM_REQUIRE(na <= n && need(na * 4) && ambigBytes == 8 + na * 4, "ambiguity section");
    cudaStream_t s = as_stream(stream);
    auto rs = std::unique_ptr<hrm_readstore>(new hrm_readstore);
    rs->n = (int64_t)n;
    rs->pitch_words = (int64_t)pitch;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    HRM_CUDA(cudaMalloc(&rs->rows, sizeof(uint32_t) * nn * (size_t)pitch));
    HRM_CUDA(cudaMalloc(&rs->lengths, sizeof(int32_t) * nn));
    HRM_CUDA(cudaMalloc(&rs->ambig, nn + 8));
    std::vector<uint8_t> flags((size_t)n, 0);
    for (uint64_t i = 0; i < na; i++) {
        uint32_t id;
        memcpy(&id, p + i * 4, 4);
        HRM_REQUIRE(id < n, "ambiguous read id out of range");
        flags[id] = 1;
    }
    if (n > 0) {
        HRM_CUDA(cudaMemcpyAsync(rs->rows, seq, (size_t)numInts * 4, cudaMemcpyHostToDevice, s));
        HRM_CUDA(cudaMemcpyAsync(rs->lengths, lens.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        HRM_CUDA(cudaMemcpyAsync(rs->ambig, flags.data(), (size_t)n, cudaMemcpyHostToDevice, s));
        HRM_CUDA(cudaStreamSynchronize(s));
    }
    int64_t c = 0;
    for (uint8_t f : flags) c += f;
    rs->with_n = c;
    readstore_finish(rs.get(), lens.data());
    *out = rs.release();
    return HRM_OK;
}
