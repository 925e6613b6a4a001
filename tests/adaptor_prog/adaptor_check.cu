// adaptor_check.cu -- TEST PROGRAM: drives libhrm_b200.so through the REFERENCE'S OWN abstract interfaces
// (care::gpu::GpuMinhasher include/gpu/gpuminhasher.cuh:20-110, care::gpu::GpuReadStorage
// include/gpu/gpureadstorage.cuh:22-119) via the two adaptor classes of hashreadmapper_b200/csrc/hrm_adaptor.hpp,
// the way constructGpuMinhasherFromReadStorage (src/gpu/gpuminhasherconstruction.cu:168-214) and
// WindowBatchProcessor (src/gpu/main_gpu.cu:534-607) call them.  Compiled in the build container against the
// reference's headers where they lie (oracle/Makefile target `adaptor`), run on the GPU box by
// tests/test_gpu_adaptor.py, which compares its output with the ctypes path and with numpy.
//
//   adaptor_check <in.bin> <out.bin>
// in : i64 n, pitch, nq, ng, k, H, maxResultsPerMap | i32 lens[n] | u8 rows[n*pitch] | i32 qlens[nq] | u8 qrows[nq*pitch]
//      | u32 gather_ids[ng]
// out: i64 pw, total, nAmbig | u32 gathered[ng*pw] | i32 glens[ng] | u8 gamb[ng] | u32 ambig_ids[nAmbig]
//      | u32 contiguous[n*pw] | i32 num[nq] | i32 off[nq+1] | u32 values[total] | i64 info[6]
#include "hrm_adaptor.hpp"

#include <rmm/mr/device/per_device_resource.hpp>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return 3;                                                                \
        }                                                                            \
    } while (0)

template <class T> static void rd(FILE* f, std::vector<T>& v, size_t n)
{
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) std::abort();
}
template <class T> static void wr(FILE* f, const std::vector<T>& v)
{
    if (!v.empty()) std::fwrite(v.data(), sizeof(T), v.size(), f);
}

int main(int argc, char** argv)
{
    if (argc != 3) return 2;
    FILE* fi = std::fopen(argv[1], "rb");
    if (!fi) return 2;
    std::vector<int64_t> hd;
    rd(fi, hd, 7);
    const int64_t n = hd[0], pitch = hd[1], nq = hd[2], ng = hd[3];
    const int k = (int)hd[4], H = (int)hd[5], maxRes = (int)hd[6];
    std::vector<int> lens, qlens;
    std::vector<char> rows, qrows;
    std::vector<read_number> gids;
    rd(fi, lens, n);
    rd(fi, rows, n * pitch);
    rd(fi, qlens, nq);
    rd(fi, qrows, nq * pitch);
    rd(fi, gids, ng);
    std::fclose(fi);
    try {
        cudaStream_t stream = cudaStreamPerThread;
        auto* mr = rmm::mr::get_current_device_resource();
        // ---- the read storage, through the abstract interface only
        std::unique_ptr<care::gpu::GpuReadStorage> rs(new hrm_b200::B200ReadStorage(rows.data(), pitch, lens.data(), n, stream));
        std::unique_ptr<care::gpu::GpuReadStorage> qs(new hrm_b200::B200ReadStorage(qrows.data(), pitch, qlens.data(), nq, stream));
        care::ReadStorageHandle rh = rs->makeHandle(), qh = qs->makeHandle();
        const int64_t pw = (rs->getSequenceLengthUpperBound() + 15) / 16;
        const int64_t qpw = (qs->getSequenceLengthUpperBound() + 15) / 16;
        read_number* d_gids;
        unsigned int *d_g, *d_all, *d_q;
        int *d_glen, *d_len, *d_qlen;
        bool* d_amb;
        CK(cudaMalloc(&d_gids, sizeof(read_number) * (ng + 1)));
        CK(cudaMalloc(&d_g, sizeof(unsigned) * (ng * pw + 1)));
        CK(cudaMalloc(&d_all, sizeof(unsigned) * (n * pw + 1)));
        CK(cudaMalloc(&d_q, sizeof(unsigned) * (nq * qpw + 1)));
        CK(cudaMalloc(&d_glen, sizeof(int) * (ng + 1)));
        CK(cudaMalloc(&d_len, sizeof(int) * (n + 1)));
        CK(cudaMalloc(&d_qlen, sizeof(int) * (nq + 1)));
        CK(cudaMalloc(&d_amb, ng + 1));
        CK(cudaMemcpyAsync(d_gids, gids.data(), sizeof(read_number) * ng, cudaMemcpyHostToDevice, stream));
        AsyncConstBufferWrapper<read_number> h_ids_wrap = makeAsyncConstBufferWrapper(gids.data());
        rs->gatherSequences(rh, d_g, (size_t)pw, h_ids_wrap, d_gids, (int)ng, stream, mr);
        rs->gatherSequenceLengths(rh, d_glen, d_gids, (int)ng, stream);
        rs->areSequencesAmbiguous(rh, d_amb, d_gids, (int)ng, stream);
        rs->gatherContiguousSequences(rh, d_all, (size_t)pw, 0, (int)n, stream, mr);
        qs->gatherContiguousSequences(qh, d_q, (size_t)qpw, 0, (int)nq, stream, mr);
        // lengths of all reads / queries by id 0..n-1
        std::vector<read_number> iota((size_t)std::max(n, nq));
        for (size_t i = 0; i < iota.size(); i++) iota[i] = (read_number)i;
        read_number* d_iota;
        CK(cudaMalloc(&d_iota, sizeof(read_number) * (iota.size() + 1)));
        CK(cudaMemcpyAsync(d_iota, iota.data(), sizeof(read_number) * iota.size(), cudaMemcpyHostToDevice, stream));
        rs->gatherSequenceLengths(rh, d_len, d_iota, (int)n, stream);
        qs->gatherSequenceLengths(qh, d_qlen, d_iota, (int)nq, stream);
        const int64_t nAmbig = rs->getNumberOfReadsWithN();
        std::vector<read_number> ambig((size_t)nAmbig);
        rs->getIdsOfAmbiguousReads(ambig.data());

        // ---- the minhasher (ref: gpuminhasherconstruction.cu:128-214: add tables, insert batches, compact)
        std::unique_ptr<care::gpu::GpuMinhasher> mh(new hrm_b200::B200Minhasher((int)n, maxRes, k, 0.8f));
        std::vector<int> fn(H);
        for (int j = 0; j < H; j++) fn[j] = j;
        const int added = mh->addHashTables(H, fn.data(), stream);
        if (added != H) return 4;
        const int half = (int)(n / 2); // two insert batches, ids continue
        mh->insert(d_all, half, d_len, (size_t)pw, d_iota, iota.data(), 0, H, fn.data(), stream, mr);
        mh->insert(d_all + (size_t)half * pw, (int)n - half, d_len + half, (size_t)pw, d_iota + half, iota.data() + half, 0, H,
                   fn.data(), stream, mr);
        if (mh->checkInsertionErrors(0, H, stream) != 0) return 5;
        mh->compact(stream);
        mh->constructionIsFinished(stream);
        care::MinhasherHandle mhh = mh->makeMinhasherHandle();
        int* d_num;
        int* d_off;
        CK(cudaMalloc(&d_num, sizeof(int) * (nq + 1)));
        CK(cudaMalloc(&d_off, sizeof(int) * (nq + 2)));
        int total = 0;
        mh->determineNumValues(mhh, d_q, (size_t)qpw, d_qlen, (int)nq, d_num, total, stream, mr);
        read_number* d_vals;
        CK(cudaMalloc(&d_vals, sizeof(read_number) * ((size_t)total + 1)));
        mh->retrieveValues(mhh, (int)nq, total, d_vals, d_num, d_off, stream, mr);
        CK(cudaStreamSynchronize(stream));

        std::vector<unsigned> g((size_t)(ng * pw)), all((size_t)(n * pw));
        std::vector<int> glen((size_t)ng), num((size_t)nq), off((size_t)nq + 1);
        std::vector<unsigned char> gamb((size_t)ng);
        std::vector<read_number> vals((size_t)total);
        CK(cudaMemcpy(g.data(), d_g, sizeof(unsigned) * g.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(all.data(), d_all, sizeof(unsigned) * all.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(glen.data(), d_glen, sizeof(int) * glen.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gamb.data(), d_amb, gamb.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(num.data(), d_num, sizeof(int) * num.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(off.data(), d_off, sizeof(int) * off.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(vals.data(), d_vals, sizeof(read_number) * vals.size(), cudaMemcpyDeviceToHost));
        std::vector<int64_t> info = {(int64_t)rs->getNumberOfReads(), rs->getSequenceLengthLowerBound(),
                                     rs->getSequenceLengthUpperBound(), mh->getNumberOfMaps(), mh->getKmerSize(),
                                     mh->getNumResultsPerMapThreshold()};
        mh->destroyHandle(mhh);
        rs->destroyHandle(rh);
        qs->destroyHandle(qh);
        FILE* fo = std::fopen(argv[2], "wb");
        if (!fo) return 2;
        std::vector<int64_t> oh = {pw, (int64_t)total, nAmbig};
        wr(fo, oh);
        wr(fo, g);
        wr(fo, glen);
        wr(fo, gamb);
        wr(fo, ambig);
        wr(fo, all);
        wr(fo, num);
        wr(fo, off);
        wr(fo, vals);
        wr(fo, info);
        std::fclose(fo);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "adaptor_check: %s\n", e.what());
        return 1;
    }
    return 0;
}
