// ref_shim_sam.cpp -- runs the REFERENCE'S OWN verification + SAM stage (Mappinghandler::go ->
// CSSW -> recalculation -> printtoSAM, src/gpu/mappinghandler.cu:67,383-766,184-293) on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/libhrm_ref_sam.so (git-ignored, its own
// library so that the macro below cannot meet another translation unit).  Nothing in the product
// links it.  The reference's file is compiled WHERE IT LIES, unmodified (#include of the path).
//
// The stage depends on undefined behaviour.  It is made well-defined here with the
// two patches of SURVEY's parity contract, both applied from OUTSIDE the reference's source:
//   (1) "own the two query strings": AlignerArguments::query / rc_query / ref / rc_ref are
//       std::string_views into a loop-local std::string and into a temporary
//       (mappinghandler.cu:458-463).  This translation unit is compiled with
//       `#define string_view string` after every standard header has been included, which turns those
//       members (and only text of the reference: no standard header sees the macro) into owning
//       std::strings.  Every other line of the reference compiles unchanged.
//   (2) "rc_ref in bounds": rc_ref = genomeRC[chr].data() + (size - pos - 1), length w
//       (mappinghandler.cu:447-449), i.e. rc_ref[i] = complement(genome[chr][pos - i]); it runs past the
//       end of the chromosome string whenever pos < w - 1.  The Genome object handed over as genomeRC is
//       built by the reference's own copy constructor (genome.hpp:152-163) and every chromosome string then
//       gets w + 64 bytes of NUL-filled spare capacity, so the reads past the end see NUL (which equals
//       none of 'A','C','G','T').  For pos >= w - 1 nothing is patched.
// Everything else is kept: the inverted `!h`, the 82-base limit, negative basesLeft, MAPQ through an
// out-of-range double -> uint32_t conversion (as compiled for x86-64).
//
// `private` is made public to read Mappinghandler::mappingout back (recalculated scores per read).
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <regex>
#include <set>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <omp.h>
#include <zlib.h>

#define string_view string
#define private public
#ifndef HRM_REFERENCE_ROOT
#define HRM_REFERENCE_ROOT /root/reference
#endif
#define HRM_STR2(x) #x
#define HRM_STR(x) HRM_STR2(x)
#include HRM_STR(HRM_REFERENCE_ROOT/src/gpu/mappinghandler.cu)
#undef private

struct ref_mapped_read_s {
    int32_t orientation, hammingDistance, shift, chromosomeId;
    int64_t position;
};

extern "C" {

// Writes "<outprefix>.SAM" (the reference's own file name rule, mappinghandler.cu:198).
// per_read (may be NULL): n_reads rows of 8 ints after the recalculation:
//   sw_score0, next_best0, sw_score1, next_best1, num_conversions0, num_conversions1, flag, flag_rc
// Returns 0, or -1 on an exception.
int ref_mapping_sam(const char* genome, const int64_t* chrom_off, int nchrom, const char* const* names,
                    const char* reads, int read_pitch, const int32_t* read_len, int64_t n_reads,
                    const ref_mapped_read_s* mapped, int windowSize, int threads, const char* outprefix,
                    int32_t* per_read)
{
    try {
        // Genome has a file constructor only (genome.hpp:121-149)
        const std::string fa = std::string(outprefix) + ".genome.fa";
        {
            std::ofstream f(fa);
            for (int c = 0; c < nchrom; c++) {
                f << ">" << names[c] << "\n";
                f.write(genome + chrom_off[c], chrom_off[c + 1] - chrom_off[c]);
                f << "\n";
            }
        }
        Genome g(fa);
        std::remove(fa.c_str());
        if ((int)g.names.size() != nchrom) return -2;
        Genome grc(g); // ref: the reverse-complement copy, main_gpu.cu (Genome(const Genome&))
        for (auto& kv : grc.data) { // patch (2): NUL-filled spare capacity behind every chromosome
            std::string& s = kv.second;
            const size_t n = s.size();
            s.resize(n + (size_t)windowSize + 64, '\0');
            s.resize(n);
        }

        int maxlen = 0;
        for (int64_t r = 0; r < n_reads; r++) maxlen = std::max(maxlen, (int)read_len[r]);
        const int pitchInts = SequenceHelpers::getEncodedNumInts2Bit(std::max(maxlen, 1));
        auto storage = std::make_unique<ChunkedReadStorage>(false, false, 8);
        {
            std::vector<int> lens(read_len, read_len + n_reads);
            std::vector<unsigned int> enc((size_t)n_reads * pitchInts, 0u);
            for (int64_t r = 0; r < n_reads; r++)
                SequenceHelpers::encodeSequence2Bit(enc.data() + r * pitchInts, reads + r * (int64_t)read_pitch, read_len[r]);
            storage->appendConsecutiveReads(0, (int)n_reads, std::move(lens), std::move(enc), pitchInts, {}, 0);
            storage->appendingFinished(std::size_t(1) << 40);
        }

        std::vector<MappedRead> results((size_t)n_reads);
        for (int64_t r = 0; r < n_reads; r++) {
            results[r].orientation = (AlignmentOrientation)mapped[r].orientation;
            results[r].hammingDistance = mapped[r].hammingDistance;
            results[r].shift = mapped[r].shift;
            results[r].chromosomeId = (std::size_t)mapped[r].chromosomeId;
            results[r].position = (std::size_t)mapped[r].position;
        }
        ProgramOptions po;
        po.windowSize = windowSize;
        po.threads = threads;
        po.outputfile = outprefix;
        po.mappType = MapperType::SW;

        // the reference narrates on std::cout; keep the test output clean
        std::ostringstream sink;
        std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
        Mappinghandler mh(&po, &g, &grc, &results);
        mh.go(storage);
        std::cout.rdbuf(old);

        if (per_read) {
            for (int64_t r = 0; r < n_reads; r++) {
                const auto& a = mh.mappingout.at((size_t)r);
                int32_t* o = per_read + 8 * r;
                o[0] = a.alignments[0].sw_score;
                o[1] = a.alignments[0].sw_score_next_best;
                o[2] = a.alignments[1].sw_score;
                o[3] = a.alignments[1].sw_score_next_best;
                o[4] = a.num_conversions[0];
                o[5] = a.num_conversions[1];
                o[6] = a.flag;
                o[7] = a.flag_rc;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_mapping_sam: %s\n", e.what());
        return -1;
    }
}

} // extern "C"

// ---- the reference's preprocessed-reads dump (ChunkedReadStorage::saveToFile / loadFromFile) -----------------------
extern "C" {

// builds a ChunkedReadStorage from ASCII rows the way constructChunkedReadStorageFromFiles appends batches and saves it
int ref_readstorage_save(const char* path, const char* reads, int read_pitch, const int32_t* read_len, int64_t n_reads,
                         const uint32_t* ambig_ids, int64_t n_ambig)
{
    try {
        int maxlen = 0;
        for (int64_t r = 0; r < n_reads; r++) maxlen = std::max(maxlen, (int)read_len[r]);
        const int pitchInts = SequenceHelpers::getEncodedNumInts2Bit(std::max(maxlen, 1));
        ChunkedReadStorage st(false, false, 8);
        const int64_t half = n_reads / 2; // two append batches, like the reference's 65536-read parser batches
        for (int part = 0; part < 2; part++) {
            const int64_t b = part == 0 ? 0 : half, e = part == 0 ? half : n_reads;
            if (e <= b) continue;
            std::vector<int> lens(read_len + b, read_len + e);
            std::vector<unsigned int> enc((size_t)(e - b) * pitchInts, 0u);
            for (int64_t r = b; r < e; r++)
                SequenceHelpers::encodeSequence2Bit(enc.data() + (r - b) * pitchInts, reads + r * (int64_t)read_pitch, read_len[r]);
            st.appendConsecutiveReads((read_number)b, (int)(e - b), std::move(lens), std::move(enc), pitchInts, {}, 0);
        }
        st.appendAmbiguousReadIds(std::vector<read_number>(ambig_ids, ambig_ids + n_ambig));
        st.appendingFinished(std::size_t(1) << 40);
        st.saveToFile(path);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_readstorage_save: %s\n", e.what());
        return -1;
    }
}

// loads a dump with the reference's loadFromFile and hands every read back (2-bit rows, lengths, ambiguous ids)
int64_t ref_readstorage_load(const char* path, uint32_t* rows, int64_t pitch_ints, int32_t* lens, int64_t cap,
                             uint32_t* ambig_ids, int64_t* n_ambig)
{
    try {
        ChunkedReadStorage st(false, false, 8);
        st.loadFromFile(path);
        const int64_t n = (int64_t)st.getNumberOfReads();
        if (n > cap) return -2;
        std::vector<read_number> ids((size_t)n);
        for (int64_t i = 0; i < n; i++) ids[(size_t)i] = (read_number)i;
        if (n > 0) {
            st.gatherSequences(rows, (std::size_t)pitch_ints, ids.data(), (int)n);
            st.gatherSequenceLengths(lens, ids.data(), (int)n);
        }
        *n_ambig = st.getNumberOfReadsWithN();
        st.getIdsOfAmbiguousReads(ambig_ids);
        return n;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_readstorage_load: %s\n", e.what());
        return -1;
    }
}

} // extern "C"
