// sam.cu -- O1: SAM-like text output (host side, no device work).
// ref: Mappinghandler::printtoSAM src/gpu/mappinghandler.cu:196-293 (SW mode), mapqfkt :184-193.
// Reproduced literally: "@HD", one "@SQ SN:<readId> LN:<windowlength>" per read, "@PG...@CO..." on
// one line, then per read 12 tab-separated columns + trailing tab; the window bases go to the RNEXT
// column; unmapped records (flag & 0x4) print the numeric flag in the TAG column.
// NOT reproduced: the score "recalculation" of mappinghandler.cu:601-766, which reads freed memory
// and out-of-bounds genome.  Scores are the raw SSW scores, Yf:i:<n> prints 0.
#include "mapper.hpp"
#include <math.h>
#include <string.h>
#include <string>

namespace {

// ref: mapqfkt mappinghandler.cu:184-193; guarded for s1 == 0 (reference: NaN -> UB)
uint32_t mapq_of(int s1, int s2)
{
    if (s1 <= 0) return 0;
    const double v = -4.343 * log(1.0 - (double)abs(s1 - s2) / (double)s1);
    uint32_t m = (uint32_t)v;
    m = (uint32_t)(m + 4.99);
    return m < 254 ? m : 254;
}

char comp(char c)
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'T'; // non-ACGT packs as A (ref: sequencehelpers.hpp:195-211), whose complement is T
    }
}
char canon(char c) { return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : 'A'; }

} // namespace

extern "C" hrm_status hrm_sam_format(const hrm_mapper* m, const hrm_read_record* h_records, const char* h_cigars,
                                     int64_t cigar_pitch, const char* h_reads_ascii, int64_t ascii_pitch,
                                     const int32_t* h_lengths, int64_t n, uint32_t first_read_id,
                                     const char* const* h_chrom_names, int with_header, char* h_out, int64_t cap,
                                     int64_t* h_written)
{
    using hrm::set_error;
    HRM_REQUIRE(m != nullptr && h_records != nullptr && h_cigars != nullptr && h_reads_ascii != nullptr &&
                    h_lengths != nullptr && h_written != nullptr,
                "args");
    HRM_REQUIRE(n >= 0 && cigar_pitch > 0 && ascii_pitch > 0, "sizes");
    std::string out;
    out.reserve((size_t)n * 640 + 256);
    if (with_header) {
        out += "@HD\tVN:1.4\n";
        for (int64_t i = 0; i < n; i++) {
            out += "@SQ\tSN:";
            out += std::to_string(first_read_id + (uint32_t)i);
            out += "\tLN:";
            out += std::to_string(h_records[i].window_length);
            out += "\n";
        }
        out += "@PG\tHashreadmapper\tID:1.0";
        out += "@CO: QNAME\tFLAG\tRNAME\tPOS\tMAPQ\tCIGAR\tRNEXT\tPNEXT\tTLEN\tSEQ\tQUAL\tTAG\n";
    }
    const std::string& G = m->host_genome;
    const int w = m->cfg.window_size;
    for (int64_t i = 0; i < n; i++) {
        const hrm_read_record& R = h_records[i];
        const bool mapped = R.mapped.orientation != HRM_ORIENT_NONE;
        const int a = R.alignments[0].sw_score >= R.alignments[1].sw_score ? 0 : 1;
        const hrm_alignment& A = R.alignments[a];
        const int chrom = mapped ? R.mapped.chromosome_id : 0;
        const int64_t wpos = mapped ? R.mapped.position : 0;
        const int64_t clen = m->chrom_off[chrom + 1] - m->chrom_off[chrom];
        const int64_t wl = (wpos + w < clen) ? w : clen - wpos; // ref: mappinghandler.cu:434-440
        const char* win = G.data() + (m->chrom_off[chrom] - m->chrom_off[0]) + wpos;
        const int L = h_lengths[i];
        const char* rd = h_reads_ascii + i * ascii_pitch;
        std::string seq((size_t)L, 'A');
        if (mapped && R.mapped.orientation == HRM_ORIENT_REVCOMP)
            for (int t = 0; t < L; t++) seq[t] = comp(rd[L - 1 - t]);
        else
            for (int t = 0; t < L; t++) seq[t] = canon(rd[t]);
        const uint32_t flag = mapped ? (uint32_t)A.flag : 0x4u;
        const uint32_t mapq = mapped ? mapq_of(A.sw_score, A.sw_score_next_best) : 0u;
        const int64_t pos = wpos + (mapped ? A.query_begin : 0); // ref: POS = windowPos + query_begin (:236,:250)
        out += std::to_string(first_read_id + (uint32_t)i);
        out += '\t';
        out += std::to_string(flag);
        out += '\t';
        out += h_chrom_names ? h_chrom_names[chrom] : std::to_string(chrom).c_str();
        out += '\t';
        out += std::to_string(pos);
        out += '\t';
        out += std::to_string(mapq);
        out += '\t';
        if (mapped) {
            const char* c = h_cigars + (2 * i + a) * cigar_pitch;
            const int64_t cl = A.cigar_len < cigar_pitch ? A.cigar_len : cigar_pitch;
            out.append(c, (size_t)cl);
        }
        out += '\t';
        out.append(win, (size_t)wl);
        out += "\t\t0\t";
        out += seq;
        out += "\t*\t";
        if (mapped) {
            out += "Yf:i:<0>";
            out += a == 0 ? "YZ:A:<+>" : "YZ:A:<->";
        } else {
            out += std::to_string(flag);
        }
        out += "\t\n";
    }
    *h_written = (int64_t)out.size();
    if (h_out && cap > 0) {
        const size_t c = out.size() < (size_t)cap ? out.size() : (size_t)cap;
        memcpy(h_out, out.data(), c);
    }
    return HRM_OK;
}
