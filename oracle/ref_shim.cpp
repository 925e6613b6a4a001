// ref_shim.cpp -- thin extern "C" wrapper around the REFERENCE'S OWN code, compiled from the
// sources where they lie under /root/reference (never copied into this repo).
//
// TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/libhrm_ref.so (git-ignored).  It is
// used (a) to pin oracle/hrm_oracle.c, (b) to generate tests/golden/*, and (c) as the
// "reference" CPU baseline of bench.py.  Nothing in the product links it.
//
// Everything here is glue: argument marshalling and the loops that the reference keeps in
// its driver (src/gpu/main_gpu.cu) -- the arithmetic is the reference's.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <algorithm>
#include <numeric>
#include <omp.h>

#include <config.hpp>
#include <sequencehelpers.hpp>
#include <helpers/hashers.cuh>
#include <fstream>
#include <cpuhashtable.hpp>
#include <readlibraryio.hpp>
#include <groupbykey.hpp>
#include <ssw_cpp.h>
#include <edlib.h>
#include <cigar.hpp>

using care::CpuReadOnlyMultiValueHashTable;
using care::GroupByKeyCpu;

extern "C" {

// ---- P1 -------------------------------------------------------------------------------
void ref_encode_2bit(uint32_t* out, const char* seq, int len)
{
    SequenceHelpers::encodeSequence2Bit(out, seq, len);
}
void ref_decode_2bit(char* out, const uint32_t* enc, int len)
{
    SequenceHelpers::decode2BitSequence(out, enc, len);
}
void ref_revcomp_2bit(uint32_t* out, const uint32_t* in, int len)
{
    SequenceHelpers::reverseComplementSequence2Bit(out, in, len);
}
void ref_revcomp_ascii(char* out, const char* in, int len)
{
    std::string s = SequenceHelpers::reverseComplementSequenceDecoded(in, len);
    std::memcpy(out, s.data(), (size_t)len);
}

// ---- H1/H2: host twin of minhashSignatures3264Kernel (gpusequencehasher.cuh:116-169) ----
uint64_t ref_murmur64(uint64_t x) { return hashers::MurmurHash<std::uint64_t>::hash(x); }

int ref_canonical_kmers(const uint32_t* enc, int len, int k, uint64_t* out)
{
    int n = 0;
    SequenceHelpers::forEachEncodedCanonicalKmerFromEncodedSequence(
        enc, len, k, [&](std::uint64_t kmer, int /*pos*/) { out[n++] = kmer; });
    return n;
}

void ref_minhash(const uint32_t* enc, int len, int k, int H, uint64_t* sig, uint8_t* valid)
{
    constexpr int maximum_kmer_length = max_k<std::uint64_t>::value;
    const std::uint64_t kmer_mask = std::numeric_limits<std::uint64_t>::max() >> ((maximum_kmer_length - k) * 2);
    for (int j = 0; j < H; j++) {
        const int hashFuncId = j;
        if (len >= k) {
            std::uint64_t minHashValue = std::numeric_limits<std::uint64_t>::max();
            SequenceHelpers::forEachEncodedCanonicalKmerFromEncodedSequence(
                enc, len, k, [&](std::uint64_t kmer, int /*pos*/) {
                    using hasher = hashers::MurmurHash<std::uint64_t>;
                    const auto hashvalue = hasher::hash(kmer + hashFuncId);
                    minHashValue = std::min(minHashValue, hashvalue);
                });
            sig[j] = minHashValue & kmer_mask;
            valid[j] = 1;
        } else {
            sig[j] = std::numeric_limits<std::uint64_t>::max();
            valid[j] = 0;
        }
    }
}

void ref_minhash_batch(const uint32_t* enc, int64_t pitch_words, const int32_t* lens, int n, int k,
                       int H, uint64_t* sigs, uint8_t* valid)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++)
        ref_minhash(enc + (int64_t)i * pitch_words, lens[i], k, H, sigs + (int64_t)i * H,
                    valid + (int64_t)i * H);
}

// ---- H3 / H3q: the reference's CPU tables (cpuhashtable.hpp:465-679, groupbykey.hpp:49-229),
// driven as FakeGpuMinhasher does (fakegpuminhasher.cuh:268-293, 360-376, 394-442, 692-720) ----
struct RefTables {
    using Table = CpuReadOnlyMultiValueHashTable<kmer_type, read_number>;
    int H;
    int maxResultsPerMap;
    std::vector<std::unique_ptr<Table>> t;
};

void* ref_tables_build(const uint64_t* sigs, const uint8_t* valid, const uint32_t* ids, int64_t n,
                       int H, int max_results_per_map, float loadfactor, int stable)
{
    auto* T = new RefTables;
    T->H = H;
    T->maxResultsPerMap = max_results_per_map;
    for (int j = 0; j < H; j++) {
        auto tab = std::make_unique<RefTables::Table>((std::uint64_t)n, loadfactor);
        std::vector<kmer_type> keys;
        std::vector<read_number> vals;
        keys.reserve(n);
        vals.reserve(n);
        for (int64_t i = 0; i < n; i++) {
            if (!valid[i * H + j]) continue;
            keys.push_back(sigs[i * H + j]);
            vals.push_back(ids ? ids[i] : (read_number)i);
        }
        if (!keys.empty()) tab->insert(keys.data(), vals.data(), (int)keys.size());
        auto groupByKey = [&](auto& k_, auto& v_, auto& countsPrefixSum) {
            // stable == 0: exactly the call FakeGpuMinhasher::compact makes on its CPU path
            //   (fakegpuminhasher.cuh:405-421: valuesOfSameKeyMustBeSorted = false).
            // stable == 1: the stable branch (groupbykey.hpp:113-119); values of a key ascend in
            //   insertion order, which is also what the default GPU group-by (radix sort + iota
            //   values, groupbykey.hpp:347-358) yields.
            GroupByKeyCpu<kmer_type, read_number, read_number> op(false, max_results_per_map, 1);
            if (stable) {
                op.valuesOfSameKeyMustBeSorted = true;
                op.executeWithIotaValues(k_, v_, countsPrefixSum);
            } else {
                op.execute(k_, v_, countsPrefixSum);
            }
        };
        tab->finalize(groupByKey, nullptr);
        T->t.push_back(std::move(tab));
    }
    return T;
}

void ref_tables_free(void* p) { delete (RefTables*)p; }

int64_t ref_tables_query(void* p, const uint64_t* qsigs, const uint8_t* qvalid, int64_t nq,
                         int32_t* num_per_seq, int64_t* offsets, uint32_t* values)
{
    auto* T = (RefTables*)p;
    const int H = T->H;
    // pass 1: counts (determineNumValues)
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; i++) {
        int numValues = 0;
        for (int map = 0; map < H; ++map) {
            if (!qvalid[i * H + map]) continue;
            auto qr = T->t[map]->query(qsigs[i * H + map]);
            if (qr.numValues <= T->maxResultsPerMap) numValues += qr.numValues;
        }
        num_per_seq[i] = numValues;
    }
    int64_t total = 0;
    for (int64_t i = 0; i < nq; i++) {
        offsets[i] = total;
        total += num_per_seq[i];
    }
    offsets[nq] = total;
    if (values) {
        // pass 2: copy ranges in map order (retrieveValues)
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t i = 0; i < nq; i++) {
            int64_t w = offsets[i];
            for (int map = 0; map < H; ++map) {
                if (!qvalid[i * H + map]) continue;
                auto qr = T->t[map]->query(qsigs[i * H + map]);
                if (qr.numValues <= T->maxResultsPerMap && qr.numValues > 0) {
                    std::copy(qr.valuesBegin, qr.valuesBegin + qr.numValues, values + w);
                    w += qr.numValues;
                }
            }
        }
    }
    return total;
}

// ---- hash-table files (--save-hashtables-to / --load-hashtables-from): the per-table stream functions are the
// reference's own (cpuhashtable.hpp:624-646 -> :217-244); the four header fields are written as
// FakeGpuMinhasher::writeToStream does (fakegpuminhasher.cuh:498-510; that class needs a GPU build) ----
int ref_tables_save(void* p, const char* path, int kmerSize, float loadfactor)
{
    auto* T = (RefTables*)p;
    std::ofstream os(path, std::ios::binary);
    if (!os) return -1;
    const int thr = T->maxResultsPerMap, numTables = T->H;
    os.write(reinterpret_cast<const char*>(&kmerSize), sizeof(int));
    os.write(reinterpret_cast<const char*>(&thr), sizeof(int));
    os.write(reinterpret_cast<const char*>(&loadfactor), sizeof(float));
    os.write(reinterpret_cast<const char*>(&numTables), sizeof(int));
    for (const auto& t : T->t) t->writeToStream(os);
    return os.good() ? 0 : -2;
}

void* ref_tables_load(const char* path, int* kmerSize, float* loadfactor)
{
    std::ifstream is(path, std::ios::binary);
    if (!is) return nullptr;
    auto* T = new RefTables;
    int numMaps = 0;
    is.read(reinterpret_cast<char*>(kmerSize), sizeof(int));
    is.read(reinterpret_cast<char*>(&T->maxResultsPerMap), sizeof(int));
    is.read(reinterpret_cast<char*>(loadfactor), sizeof(float));
    is.read(reinterpret_cast<char*>(&numMaps), sizeof(int));
    T->H = numMaps;
    for (int i = 0; i < numMaps; i++) {
        auto ptr = std::make_unique<RefTables::Table>();
        ptr->loadFromStream(is);
        T->t.push_back(std::move(ptr));
    }
    if (!is.good()) {
        delete T;
        return nullptr;
    }
    return T;
}

// ---- read files: the reference's own parser (forEachReadInFile include/readlibraryio.hpp:288-326, kseqpp) ----
// rows: cap rows of `pitch` bytes; returns the number of reads in the file (rows beyond cap are not written)
int64_t ref_read_file(const char* path, char* rows, int64_t pitch, int32_t* lens, int64_t cap)
{
    int64_t n = 0;
    care::forEachReadInFile(std::string(path), [&](auto /*readNumber*/, auto& read) {
        if (n < cap) {
            const int64_t L = (int64_t)read.sequence.size();
            lens[n] = (int32_t)L;
            memcpy(rows + n * pitch, read.sequence.data(), (size_t)(L < pitch ? L : pitch));
        }
        n++;
    });
    return n;
}

// ---- V2: the reference's SSW (src/ssw.c, src/ssw_cpp.cpp) ----
struct ref_alignment {
    int32_t sw_score, sw_score_next_best, ref_begin, ref_end, query_begin, query_end,
        ref_end_next_best, mismatches, flag, cigar_len;
};

void ref_ssw_align(const char* query, int qlen, const char* ref, int rlen, int maskLen,
                   ref_alignment* al, char* cigar, int cigar_cap)
{
    static thread_local StripedSmithWaterman::Aligner aligner;
    StripedSmithWaterman::Filter filter;
    StripedSmithWaterman::Alignment a;
    std::string q(query, (size_t)qlen); // Align() uses strlen(query)
    std::string r(ref, (size_t)rlen);
    const uint16_t flag = aligner.Align(q.c_str(), r.c_str(), rlen, filter, &a, maskLen);
    al->sw_score = a.sw_score;
    al->sw_score_next_best = a.sw_score_next_best;
    al->ref_begin = a.ref_begin;
    al->ref_end = a.ref_end;
    al->query_begin = a.query_begin;
    al->query_end = a.query_end;
    al->ref_end_next_best = a.ref_end_next_best;
    al->mismatches = a.mismatches;
    al->flag = flag;
    al->cigar_len = (int)a.cigar_string.size();
    if (cigar_cap > 0) {
        const size_t n = std::min((size_t)cigar_cap - 1, a.cigar_string.size());
        std::memcpy(cigar, a.cigar_string.data(), n);
        cigar[n] = 0;
    }
}

// batch form used by the CPU baseline: 2 alignments per read, OpenMP over reads
void ref_ssw_align_batch(const char* queries, int q_pitch, const int32_t* qlens, const char* refs,
                         int r_pitch, const int32_t* rlens, const int32_t* maskLens, int64_t n,
                         ref_alignment* out, char* cigars, int cigar_pitch)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++)
        ref_ssw_align(queries + i * q_pitch, qlens[i], refs + i * r_pitch, rlens[i], maskLens[i],
                      out + i, cigars + i * cigar_pitch, cigar_pitch);
}

// ---- V3: edlib as called by the reference (mappinghandler.cu:968-987) ----
int ref_edit_distance_nw(const char* q, int qlen, const char* t, int tlen)
{
    EdlibAlignResult result = edlibAlign(q, qlen, t, tlen, edlibDefaultAlignConfig());
    int d = result.status == EDLIB_STATUS_OK ? result.editDistance : -1;
    edlibFreeAlignResult(result);
    return d;
}

void ref_edit_distance_nw_batch(const char* queries, int q_pitch, const int32_t* qlens,
                                const char* refs, int r_pitch, const int32_t* rlens, int64_t n,
                                int32_t* out)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++)
        out[i] = ref_edit_distance_nw(queries + i * q_pitch, qlens[i], refs + i * r_pitch, rlens[i]);
}

// ---- cigar parser (src/cigar.cpp) : returns number of entries, ops as chars ----
int ref_cigar_parse(const char* s, char* ops, int32_t* lens, int cap)
{
    Cigar c{std::string(s)};
    int n = 0;
    for (const auto& e : c.getEntries()) {
        if (n < cap) {
            ops[n] = Cigar::opToChar(e.first);
            lens[n] = e.second;
        }
        n++;
    }
    return n;
}

int ref_num_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1; the caller sets the thread count to its CPU affinity explicitly
void ref_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

} // extern "C"

// ---------------------------------------------------------------------------------------------
// Whole CPU path, reference direction, assembled from the reference's own functions with the
// reference's own parallelisation (OpenMP over sequences / a parallel loop over reads).  This is
// the CPU baseline that bench.py times.  Glue only: the loops mirror
// gpuminhasherconstruction.cu:168-214 (build), main_gpu.cu:471-854 (window batches) and
// mappinghandler.cu:397-595 (verification).  The frequency filter has no CPU implementation in the
// reference (CUDA only, cuda_unique_by_count.cuh); it is restated with std::sort.
// ---------------------------------------------------------------------------------------------
extern "C" void ref_window_location(int, int, int, int, int, int*, int*, int*, int*);
extern "C" void ref_shd(const uint32_t*, int, const uint32_t*, int, float, int*, int*, int*);

struct ref_mapped_read {
    int32_t orientation, hammingDistance, shift, chromosomeId;
    int64_t position;
};

#include <chrono>
static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

// times[0] = read side (encode + minhash + table build), times[1] = window streaming (encode,
// minhash, query, filter), times[2] = extended windows + SHD + arg-min, times[3] = verification.
// conv_* are applied up front on the ASCII (the reference sees pre-converted input).
// mapper_type 0: 2 x SSW per mapped read, 1: 2 x edlib.  sw_out may be NULL.
void ref_cpu_pipeline(const char* genome, const int64_t* chrom_off, int nchrom, const char* reads, int read_pitch,
                      const int32_t* read_len, int64_t n_reads, int k, int w, int H, int min_hits,
                      int max_results_per_map, float rate, int batchsize, int mapper_type,
                      ref_mapped_read* out, ref_alignment* sw_out, int32_t* ed_out, double* times)
{
    double t0 = now_s();
    int maxReadLen = 0;
    for (int64_t r = 0; r < n_reads; r++) maxReadLen = std::max(maxReadLen, (int)read_len[r]);
    const int rpw = SequenceHelpers::getEncodedNumInts2Bit(std::max(maxReadLen, 1));
    std::vector<unsigned int> renc((size_t)std::max<int64_t>(n_reads, 1) * rpw, 0u);
    std::vector<uint64_t> rs((size_t)std::max<int64_t>(n_reads, 1) * H);
    std::vector<uint8_t> rv((size_t)std::max<int64_t>(n_reads, 1) * H);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n_reads; r++) {
        SequenceHelpers::encodeSequence2Bit(renc.data() + r * rpw, reads + r * (int64_t)read_pitch, read_len[r]);
        ref_minhash(renc.data() + r * rpw, read_len[r], k, H, rs.data() + r * H, rv.data() + r * H);
        out[r].orientation = 3;
        out[r].hammingDistance = 0;
        out[r].shift = 0;
        out[r].chromosomeId = 0;
        out[r].position = 0;
    }
    void* T = ref_tables_build(rs.data(), rv.data(), nullptr, n_reads, H, max_results_per_map, 0.8f, 0);
    times[0] = now_s() - t0;
    times[1] = times[2] = times[3] = 0;

    const int stride = w - k + 1;
    const int wpw = SequenceHelpers::getEncodedNumInts2Bit(w);
    const int xpw = SequenceHelpers::getEncodedNumInts2Bit(w + maxReadLen) + 1;
    for (int c = 0; c < nchrom; c++) {
        const char* chr = genome + chrom_off[c];
        const int64_t clen = chrom_off[c + 1] - chrom_off[c];
        const int64_t nwin = (clen + stride - 1) / stride;
        for (int64_t b0 = 0; b0 < nwin; b0 += batchsize) {
            double t1 = now_s();
            const int64_t b1 = std::min<int64_t>(b0 + batchsize, nwin);
            const int nb = (int)(b1 - b0);
            int64_t secB = b0 * stride - maxReadLen / 2, secE = (b1 - 1) * stride + w + maxReadLen / 2;
            secB = std::max<int64_t>(secB, 0);
            secE = std::min<int64_t>(secE, clen);
            std::vector<unsigned int> wenc((size_t)nb * wpw, 0u);
            std::vector<int32_t> wlen(nb);
            std::vector<uint64_t> ws((size_t)nb * H);
            std::vector<uint8_t> wv((size_t)nb * H);
#pragma omp parallel for schedule(dynamic, 16)
            for (int i = 0; i < nb; i++) {
                const int64_t p = (b0 + i) * stride;
                wlen[i] = (int)(p + w <= clen ? w : clen - p);
                SequenceHelpers::encodeSequence2Bit(wenc.data() + (size_t)i * wpw, chr + p, wlen[i]);
                ref_minhash(wenc.data() + (size_t)i * wpw, wlen[i], k, H, ws.data() + (size_t)i * H, wv.data() + (size_t)i * H);
            }
            std::vector<int32_t> num(nb);
            std::vector<int64_t> off(nb + 1);
            int64_t total = ref_tables_query(T, ws.data(), wv.data(), nb, num.data(), off.data(), nullptr);
            std::vector<uint32_t> vals((size_t)std::max<int64_t>(total, 1));
            if (total > 0) ref_tables_query(T, ws.data(), wv.data(), nb, num.data(), off.data(), vals.data());
            // frequency filter per window
            std::vector<int64_t> foff(nb + 1, 0);
            std::vector<uint32_t> flt((size_t)std::max<int64_t>(total, 1));
            std::vector<int32_t> fcnt(nb, 0);
#pragma omp parallel for schedule(dynamic, 16)
            for (int i = 0; i < nb; i++) {
                uint32_t* b = vals.data() + off[i];
                uint32_t* e = vals.data() + off[i + 1];
                std::sort(b, e);
                int kept = 0;
                for (uint32_t* p = b; p < e;) {
                    uint32_t* q = p;
                    while (q < e && *q == *p) ++q;
                    if (q - p >= min_hits || min_hits <= 1) b[kept++] = *p;
                    p = q;
                }
                fcnt[i] = kept;
            }
            for (int i = 0; i < nb; i++) foff[i + 1] = foff[i] + fcnt[i];
            for (int i = 0; i < nb; i++) std::copy(vals.data() + off[i], vals.data() + off[i] + fcnt[i], flt.data() + foff[i]);
            const int64_t ncand = foff[nb];
            times[1] += now_s() - t1;
            t1 = now_s();
            std::vector<int> c_shift(std::max<int64_t>(ncand, 1)), c_score(std::max<int64_t>(ncand, 1)),
                c_ori(std::max<int64_t>(ncand, 1)), c_left(std::max<int64_t>(ncand, 1)), c_win(std::max<int64_t>(ncand, 1));
            for (int i = 0; i < nb; i++)
                for (int64_t e = foff[i]; e < foff[i + 1]; e++) c_win[e] = i;
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t e = 0; e < ncand; e++) {
                const uint32_t rid = flt[e];
                const int64_t p = (b0 + c_win[e]) * stride;
                int left, right, len, sp;
                ref_window_location((int)secB, (int)secE, (int)p, w, read_len[rid] / 2, &left, &right, &len, &sp);
                std::vector<unsigned int> x(xpw, 0u);
                SequenceHelpers::encodeSequence2Bit(x.data(), chr + secB + sp, len);
                ref_shd(x.data(), len, renc.data() + (size_t)rid * rpw, read_len[rid], rate, &c_shift[e], &c_score[e], &c_ori[e]);
                c_left[e] = left;
            }
            for (int64_t e = 0; e < ncand; e++) { // ref: main_gpu.cu:777-821, serial, window order
                if (c_ori[e] == 3) continue;
                ref_mapped_read& br = out[flt[e]];
                if (br.orientation == 3 || br.hammingDistance > c_score[e]) {
                    br.orientation = c_ori[e];
                    br.hammingDistance = c_score[e];
                    br.shift = c_shift[e] - c_left[e];
                    br.chromosomeId = c;
                    br.position = (b0 + c_win[e]) * stride;
                }
            }
            times[2] += now_s() - t1;
        }
    }
    ref_tables_free(T);

    // verification (mappinghandler.cu:397-595): 3N(read'), 3N(RC(read')) vs 3N(window), C->T
    double t2 = now_s();
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < n_reads; r++) {
        if (sw_out) {
            std::memset(&sw_out[2 * r], 0, 2 * sizeof(ref_alignment));
        }
        if (ed_out) ed_out[2 * r] = ed_out[2 * r + 1] = -1;
        if (out[r].orientation == 3) continue;
        const int L = read_len[r];
        std::vector<unsigned int> enc(renc.begin() + r * rpw, renc.begin() + (r + 1) * rpw);
        if (out[r].orientation == 2) SequenceHelpers::reverseComplementSequenceInplace2Bit(enc.data(), L);
        std::string rd = SequenceHelpers::get2BitString(enc.data(), L);
        std::string rc = SequenceHelpers::reverseComplementSequenceDecoded(rd.data(), L);
        const char* chr = genome + chrom_off[out[r].chromosomeId];
        const int64_t clen = chrom_off[out[r].chromosomeId + 1] - chrom_off[out[r].chromosomeId];
        const int64_t wl = out[r].position + w < clen ? w : clen - out[r].position;
        std::string win(chr + out[r].position, (size_t)wl);
        for (auto* s : {&rd, &rc, &win})
            for (auto& ch : *s)
                if (ch == 'C') ch = 'T';
        const int maskLen = std::max(15, L / 2);
        char cig[8];
        ref_alignment a0, a1;
        if (mapper_type == 0) {
            ref_ssw_align(rd.data(), L, win.data(), (int)wl, maskLen, &a0, cig, 0);
            ref_ssw_align(rc.data(), L, win.data(), (int)wl, maskLen, &a1, cig, 0);
            if (sw_out) {
                sw_out[2 * r] = a0;
                sw_out[2 * r + 1] = a1;
            }
        } else {
            const int d0 = ref_edit_distance_nw(rd.data(), L, win.data(), (int)wl);
            const int d1 = ref_edit_distance_nw(rc.data(), L, win.data(), (int)wl);
            if (ed_out) {
                ed_out[2 * r] = d0;
                ed_out[2 * r + 1] = d1;
            }
        }
    }
    times[3] = now_s() - t2;
}

} // extern "C"
