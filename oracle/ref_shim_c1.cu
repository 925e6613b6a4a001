// ref_shim_c1.cu -- the REFERENCE'S OWN candidate filter (C1) as a callable: GpuSegmentedUniqueByCount::unique
// (include/gpu/cuda_unique_by_count.cuh:33-215: segmented radix sort, run-length encode, threshold, compaction), the
// function GpuMinhashQueryFilter::keepDistinctByFrequency drives (include/gpu/minhashqueryfilter.cuh:239-278).
// It exists only as CUDA in the reference, so this checker needs a GPU: it is compiled here (build container, CUDA 12.9's
// bundled CCCL instead of the vendored cub/thrust 1.16, see oracle/Makefile) from the header where it lies and RUN on
// the GPU box by tests/test_gpu_kernels.py::test_k4_against_reference_unique_by_count.
//
// TEST INFRASTRUCTURE ONLY.  Output: oracle/_ref/libhrm_ref_c1.so (git-ignored).  Nothing in the product links it.
#include <cstdint>
#include <vector>
#include <hpc_helpers.cuh> // D2D / H2D aliases the reference's includers provide
#include <gpu/cuda_unique_by_count.cuh>

extern "C" {

// host arrays in, host arrays out.  out must hold numItems entries; out_lengths numSegments.  The reference writes the
// surviving ids of all segments back to back (segment order, ascending inside a segment): sum(out_lengths) entries.
int ref_unique_by_count(const uint32_t* h_items, int numItems, int numSegments, const int* h_offsets, int minimumCount,
                        uint32_t* h_out, int* h_out_lengths)
{
    try {
        cudaStream_t stream = 0;
        uint32_t *d_items = nullptr, *d_unique = nullptr;
        int *d_off = nullptr, *d_len = nullptr;
        const size_t ni = (size_t)(numItems > 0 ? numItems : 1), ns = (size_t)(numSegments > 0 ? numSegments : 1);
        if (cudaMalloc(&d_items, sizeof(uint32_t) * ni) != cudaSuccess) return -1;
        if (cudaMalloc(&d_unique, sizeof(uint32_t) * ni) != cudaSuccess) return -1;
        if (cudaMalloc(&d_off, sizeof(int) * (ns + 1)) != cudaSuccess) return -1;
        if (cudaMalloc(&d_len, sizeof(int) * ns) != cudaSuccess) return -1;
        cudaMemcpy(d_items, h_items, sizeof(uint32_t) * (size_t)numItems, cudaMemcpyHostToDevice);
        cudaMemcpy(d_off, h_offsets, sizeof(int) * ((size_t)numSegments + 1), cudaMemcpyHostToDevice);
        cudaMemset(d_unique, 0, sizeof(uint32_t) * ni);
        GpuSegmentedUniqueByCount::unique<uint32_t>(d_items, numItems, d_unique, d_len, numSegments, d_off, minimumCount,
                                                    stream);
        if (cudaStreamSynchronize(stream) != cudaSuccess) return -2;
        cudaMemcpy(h_out, d_unique, sizeof(uint32_t) * (size_t)numItems, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_out_lengths, d_len, sizeof(int) * (size_t)numSegments, cudaMemcpyDeviceToHost);
        cudaFree(d_items);
        cudaFree(d_unique);
        cudaFree(d_off);
        cudaFree(d_len);
        return 0;
    } catch (...) {
        return -3;
    }
}

} // extern "C"
