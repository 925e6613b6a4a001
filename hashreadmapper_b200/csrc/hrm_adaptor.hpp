// hrm_adaptor.hpp -- the reference-side binding: two C++ classes that derive from the reference's abstract
// care::gpu::GpuMinhasher (include/gpu/gpuminhasher.cuh:20-110) and care::gpu::GpuReadStorage
// (include/gpu/gpureadstorage.cuh:22-119) and forward every virtual to the C ABI of libhrm_b200.so.  Drop this header next to the reference's sources, compile it with the reference
// (nvcc, -I<reference>/include, rmm on the include path as in the reference's Makefile:23-31), link
// -lhrm_b200, and construct B200Minhasher where constructGpuMinhasherFromGpuReadStorage
// (src/gpu/gpuminhasherconstruction.cu:256-330) constructs FakeGpuMinhasher / SingleGpuMinhasher.
// Nothing else of the reference changes: WindowBatchProcessor (src/gpu/main_gpu.cu:431-856) holds the
// minhasher through `const GpuMinhasher*` and keeps calling determineNumValues / retrieveValues.
//
// Error behaviour mirrors CUDACHECK (include/gpu/cudaerrorcheck.cuh:42-58): a non-zero hrm_status is
// re-thrown as std::runtime_error carrying hrm_last_error().
//
// tests/test_host_logic.py::test_adaptor_compiles_against_reference compiles this header against the
// reference's own headers when /root/reference is present.
#pragma once
#include <gpu/gpuminhasher.cuh>   // the reference's interfaces
#include <gpu/gpureadstorage.cuh>

#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <iterator>

#include "hrm_b200.h"

namespace hrm_b200 {

inline void hrm_check(hrm_status st)
{
    if (st != HRM_OK) throw std::runtime_error(std::string("libhrm_b200: ") + hrm_last_error());
}

class B200Minhasher : public care::gpu::GpuMinhasher {
public:
    // ref: FakeGpuMinhasher(int maxNumKeys, int maxValuesPerKey, int k, float loadfactor)
    //      include/gpu/fakegpuminhasher.cuh:150-153
    B200Minhasher(int maxNumKeys, int maxValuesPerKey, int k, float loadfactor)
    {
        hrm_check(hrm_minhasher_create(&mh_, maxNumKeys, maxValuesPerKey, k, loadfactor));
    }
    ~B200Minhasher() override { hrm_minhasher_destroy(mh_); }
    B200Minhasher(const B200Minhasher&) = delete;
    B200Minhasher& operator=(const B200Minhasher&) = delete;

    care::MinhasherHandle makeMinhasherHandle() const override
    {
        const int id = hrm_minhasher_handle_create(mh_);
        if (id < 0) hrm_check(id);
        return constructHandle(id);
    }
    void destroyHandle(care::MinhasherHandle& handle) const override
    {
        hrm_check(hrm_minhasher_handle_destroy(mh_, handle.getId()));
        handle = constructHandle(std::numeric_limits<int>::max());
    }

    void setHostMemoryLimitForConstruction(std::size_t) override {}   // tables live on the device
    void setDeviceMemoryLimitsForConstruction(const std::vector<std::size_t>&) override {}
    void setThreadPool(care::ThreadPool*) override {}                // no host threads needed
    int addHashTables(int numAdditionalTables, const int* hashFunctionIds, cudaStream_t stream) override
    {
        return hrm_minhasher_add_tables(mh_, numAdditionalTables, hashFunctionIds, stream);
    }
    void compact(cudaStream_t stream) override { hrm_check(hrm_minhasher_compact(mh_, stream)); }
    void constructionIsFinished(cudaStream_t stream) override { hrm_check(hrm_minhasher_finish(mh_, stream)); }

    void insert(const unsigned int* d_sequenceData2Bit, int numSequences, const int* d_sequenceLengths,
                std::size_t encodedSequencePitchInInts, const read_number* d_readIds, const read_number* /*h_readIds*/,
                int firstHashfunction, int numHashfunctions, const int* /*h_hashFunctionNumbers*/, cudaStream_t stream,
                rmm::mr::device_memory_resource* /*mr*/) override
    {
        hrm_check(hrm_minhasher_insert(mh_, d_sequenceData2Bit, (int64_t)encodedSequencePitchInInts, d_sequenceLengths,
                                       numSequences, d_readIds, 0u, firstHashfunction, numHashfunctions, stream));
    }
    int checkInsertionErrors(int firstHashfunction, int numHashfunctions, cudaStream_t stream) override
    {
        return hrm_minhasher_check_insertion_errors(mh_, firstHashfunction, numHashfunctions, stream);
    }

    void determineNumValues(care::MinhasherHandle& queryHandle, const unsigned int* d_sequenceData2Bit,
                            std::size_t encodedSequencePitchInInts, const int* d_sequenceLengths, int numSequences,
                            int* d_numValuesPerSequence, int& totalNumValues, cudaStream_t stream,
                            rmm::mr::device_memory_resource* /*mr*/) const override
    {
        int64_t total = 0;
        hrm_check(hrm_minhasher_count(mh_, queryHandle.getId(), d_sequenceData2Bit, (int64_t)encodedSequencePitchInInts,
                                      d_sequenceLengths, numSequences, d_numValuesPerSequence, &total, stream));
        totalNumValues = (int)total;
    }
    void retrieveValues(care::MinhasherHandle& queryHandle, int numSequences, int totalNumValues, read_number* d_values,
                        const int* d_numValuesPerSequence, int* d_offsets, cudaStream_t stream,
                        rmm::mr::device_memory_resource* /*mr*/) const override
    {
        hrm_check(hrm_minhasher_retrieve(mh_, queryHandle.getId(), numSequences, totalNumValues, d_values,
                                         d_numValuesPerSequence, d_offsets, stream));
    }

    ::MemoryUsage getMemoryInfo() const noexcept override
    {
        ::MemoryUsage mu{};
        hrm_minhasher_info_t info;
        if (hrm_minhasher_info(mh_, &info) == HRM_OK) {
            int dev = 0;
            cudaGetDevice(&dev);
            mu.device[dev] = (std::size_t)info.device_bytes;
        }
        return mu;
    }
    ::MemoryUsage getMemoryInfo(const care::MinhasherHandle&) const noexcept override { return ::MemoryUsage{}; }
    int getNumResultsPerMapThreshold() const noexcept override { return info().max_results_per_map; }
    int getNumberOfMaps() const noexcept override { return info().num_tables; }
    int getKmerSize() const noexcept override { return info().k; }
    bool hasGpuTables() const noexcept override { return true; }

    // the reference's own file format (fakegpuminhasher.cuh:498-532): files are interchangeable with
    // FakeGpuMinhasher's --save-hashtables-to / --load-hashtables-from
    void writeToStream(std::ostream& os) const override
    {
        int64_t size = 0;
        hrm_check(hrm_minhasher_write_reference_format(mh_, nullptr, &size));
        std::vector<char> buf((std::size_t)size);
        hrm_check(hrm_minhasher_write_reference_format(mh_, buf.data(), &size));
        os.write(buf.data(), size);
    }
    int loadFromStream(std::ifstream& is, int numMapsUpperLimit) override
    {
        const std::vector<char> buf((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        hrm_minhasher* fresh = nullptr;
        hrm_check(hrm_minhasher_read_reference_format(&fresh, buf.data(), (int64_t)buf.size(), numMapsUpperLimit));
        hrm_minhasher_destroy(mh_);
        mh_ = fresh;
        return info().num_tables;
    }
    bool canWriteToStream() const noexcept override { return true; }
    bool canLoadFromStream() const noexcept override { return true; }

private:
    hrm_minhasher_info_t info() const noexcept
    {
        hrm_minhasher_info_t i{};
        hrm_minhasher_info(mh_, &i);
        return i;
    }
    hrm_minhasher* mh_ = nullptr;
};

// ref: class GpuReadStorage include/gpu/gpureadstorage.cuh:22-119 (the reference's implementation:
// MultiGpuReadStorage include/gpu/multigpureadstorage.cuh:655-905,1515-1560).  Construct it where performMappingGpu
// builds its MultiGpuReadStorage from the ChunkedReadStorage (src/gpu/main_gpu.cu:1010-1040); WindowBatchProcessor
// keeps calling gatherSequences / gatherSequenceLengths through `const GpuReadStorage*` (main_gpu.cu:591-607).
// Reads live 2-bit packed in HBM; quality scores are not stored (the mapping path never reads them).
class B200ReadStorage : public care::gpu::GpuReadStorage {
public:
    // host ASCII rows (what the reference's encoder threads are given, chunkedreadstorageconstruction.hpp:273-314)
    B200ReadStorage(const char* h_ascii, std::int64_t ascii_pitch, const int* h_lengths, std::int64_t n,
                    cudaStream_t stream = nullptr)
    {
        hrm_check(hrm_readstore_create_from_ascii(&rs_, h_ascii, ascii_pitch, h_lengths, n, HRM_CONV_NONE, stream));
    }
    // already packed device rows (what ChunkedReadStorage holds, chunkedreadstorage.hpp:44-80) + optional
    // per-read ambiguity flags on the device
    B200ReadStorage(const unsigned int* d_seq2bit, std::size_t pitchInInts, const int* d_lengths, std::int64_t n,
                    const std::uint8_t* d_ambiguous, cudaStream_t stream = nullptr)
    {
        hrm_check(hrm_readstore_create_from_2bit(&rs_, d_seq2bit, (std::int64_t)pitchInInts, d_lengths, n, stream));
        if (d_ambiguous) hrm_check(hrm_readstore_set_ambiguous(rs_, d_ambiguous, stream));
    }
    ~B200ReadStorage() override { hrm_readstore_destroy(rs_); }
    B200ReadStorage(const B200ReadStorage&) = delete;
    B200ReadStorage& operator=(const B200ReadStorage&) = delete;

    care::ReadStorageHandle makeHandle() const override
    {
        const int id = hrm_readstore_handle_create(rs_);
        if (id < 0) hrm_check(id);
        return constructHandle(id);
    }
    void destroyHandle(care::ReadStorageHandle& handle) const override
    {
        hrm_check(hrm_readstore_handle_destroy(rs_, handle.getId()));
        handle = constructHandle(std::numeric_limits<int>::max());
    }
    void areSequencesAmbiguous(care::ReadStorageHandle& handle, bool* d_result, const read_number* d_readIds,
                               int numSequences, cudaStream_t stream) const override
    {
        static_assert(sizeof(bool) == 1, "bool flags are bytes");
        hrm_check(hrm_readstore_are_ambiguous(rs_, handle.getId(), reinterpret_cast<std::uint8_t*>(d_result), d_readIds,
                                              numSequences, stream));
    }
    void gatherSequences(care::ReadStorageHandle& handle, unsigned int* d_sequence_data, size_t outSequencePitchInInts,
                         const AsyncConstBufferWrapper<read_number> /*h_readIds*/, const read_number* d_readIds,
                         int numSequences, cudaStream_t stream, rmm::mr::device_memory_resource* /*mr*/) const override
    {
        hrm_check(hrm_readstore_gather(rs_, handle.getId(), d_sequence_data, (std::int64_t)outSequencePitchInInts,
                                       d_readIds, numSequences, stream));
    }
    void gatherContiguousSequences(care::ReadStorageHandle& handle, unsigned int* d_sequence_data,
                                   size_t outSequencePitchInInts, read_number firstIndex, int numSequences,
                                   cudaStream_t stream, rmm::mr::device_memory_resource* /*mr*/) const override
    {
        hrm_check(hrm_readstore_gather_contiguous(rs_, handle.getId(), d_sequence_data,
                                                  (std::int64_t)outSequencePitchInInts, firstIndex, numSequences, stream));
    }
    // quality scores are not part of the mapping path (canUseQualityScores() == false): contract violation
    void gatherQualities(care::ReadStorageHandle&, char*, size_t, const AsyncConstBufferWrapper<read_number>,
                         const read_number*, int, cudaStream_t, rmm::mr::device_memory_resource*) const override
    {
        throw std::runtime_error("libhrm_b200: B200ReadStorage stores no quality scores");
    }
    void gatherContiguousQualities(care::ReadStorageHandle&, char*, size_t, read_number, int, cudaStream_t,
                                   rmm::mr::device_memory_resource*) const override
    {
        throw std::runtime_error("libhrm_b200: B200ReadStorage stores no quality scores");
    }
    void gatherSequenceLengths(care::ReadStorageHandle& handle, int* d_lengths, const read_number* d_readIds,
                               int numSequences, cudaStream_t stream) const override
    {
        hrm_check(hrm_readstore_gather_lengths(rs_, handle.getId(), d_lengths, d_readIds, numSequences, stream));
    }
    void getIdsOfAmbiguousReads(read_number* ids) const override { hrm_check(hrm_readstore_ambiguous_ids(rs_, ids)); }
    std::int64_t getNumberOfReadsWithN() const override { return info().num_reads_with_n; }
    ::MemoryUsage getMemoryInfo() const override
    {
        ::MemoryUsage mu{};
        int dev = 0;
        cudaGetDevice(&dev);
        mu.device[dev] = (std::size_t)info().device_bytes;
        return mu;
    }
    ::MemoryUsage getMemoryInfo(const care::ReadStorageHandle&) const override { return ::MemoryUsage{}; }
    read_number getNumberOfReads() const override { return (read_number)info().num_reads; }
    bool canUseQualityScores() const override { return false; }
    int getSequenceLengthLowerBound() const override { return info().length_lower_bound; }
    int getSequenceLengthUpperBound() const override { return info().length_upper_bound; }
    bool isPairedEnd() const override { return info().is_paired_end != 0; }

private:
    hrm_readstore_info_t info() const
    {
        hrm_readstore_info_t i{};
        hrm_readstore_info(rs_, &i);
        return i;
    }
    hrm_readstore* rs_ = nullptr;
};

} // namespace hrm_b200
