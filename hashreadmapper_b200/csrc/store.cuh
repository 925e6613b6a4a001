// store.cuh -- device-resident genome and read storage (internal interface).
#pragma once
#include "runtime.cuh"
#include <vector>
#include <memory>
#include <mutex>

namespace hrm {
// what kernels see of a genome
struct GenomeDev {
    const uint32_t* const* chrom_words; // device array [n_chrom] of device pointers
    const int64_t* chrom_len;           // device array [n_chrom]
    int32_t n_chrom;
};
} // namespace hrm

struct hrm_genome {
    int n_chrom = 0;
    int conversion = 0;
    std::vector<int64_t> chrom_len;      // host
    std::vector<uint32_t*> chrom_words;  // host array of device pointers (into `words`)
    uint32_t* words = nullptr;           // one allocation, chromosomes word-aligned back to back (+ pad)
    int64_t total_words = 0;
    // device mirrors
    const uint32_t** d_chrom_words = nullptr;
    int64_t* d_chrom_len = nullptr;
    hrm::GenomeDev dev() const { return hrm::GenomeDev{d_chrom_words, d_chrom_len, n_chrom}; }
    int64_t device_bytes() const { return total_words * 4 + (int64_t)n_chrom * 16; }
};

struct hrm_readstore {
    int64_t n = 0;
    int64_t pitch_words = 0;
    uint32_t* rows = nullptr;
    int32_t* lengths = nullptr;
    uint8_t* ambig = nullptr;  // per read: a non-ACGT character was replaced (NULL: none known)
    int32_t len_min = 0, len_max = 0;
    int64_t with_n = 0;
    std::mutex mtx;
    std::vector<bool> handles;
};
