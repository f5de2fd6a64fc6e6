// streaming.cu — StreamingMatrix (ref src/matrix/optimized.rs:451-561): a matrix kept as row chunks sized from a memory
// limit, multiplied chunk by chunk with a callback per chunk.
//
// B200 form: the chunks live in PINNED HOST memory (the matrix may exceed what the caller wants resident in HBM); the
// device holds x and two chunk-sized staging sets. multiply_vector_streaming double-buffers: the copy stream uploads chunk
// k + 1 while the compute stream runs the warp-stream SpMV kernel (csrc/kernels.cu, the same row-ordered sums as every
// other SpMV of the library) on chunk k and returns its y slice; the callback sees the chunks in row order, like the
// reference's loop. Chunk size rule and triplet semantics (stable sort by row, zeros dropped, duplicates kept, stable
// (row, col) order inside a chunk) follow StreamingMatrix::from_triplets (:467-528).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "matrix.hpp"

using namespace sb200;

struct sb200_streaming_matrix {
    int device = 0;
    uint64_t total_rows = 0, total_cols = 0, chunk_size = 0, memory_limit = 0, nnz = 0;
    struct Chunk {
        uint64_t rows = 0, nnz = 0;
        uint32_t *row_ptr = nullptr;  // pinned, rows + 1
        uint32_t *cols = nullptr;     // pinned, nnz
        double *vals = nullptr;       // pinned, nnz
    };
    std::vector<Chunk> chunks;
    uint64_t max_chunk_nnz = 0;
    ~sb200_streaming_matrix() {
        for (auto &c : chunks) {
            if (c.row_ptr) cudaFreeHost(c.row_ptr);
            if (c.cols) cudaFreeHost(c.cols);
            if (c.vals) cudaFreeHost(c.vals);
        }
    }
};

extern "C" {

int32_t sb200_streaming_matrix_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t nt,
                                             uint64_t nrows, uint64_t ncols, uint64_t memory_limit_mb,
                                             sb200_streaming_matrix **out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    if (nrows >= 0xFFFFFFF0ull || ncols >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_INVALID_INPUT, "dimension exceeds the u32 IndexType of the reference (src/types.rs:22)");
    SB_TRY(validate_triplets(rows, cols, vals, nt, nrows, ncols));  // COOStorage::from_triplets per chunk would reject them
    std::unique_ptr<sb200_streaming_matrix> sm(new sb200_streaming_matrix());
    sm->device = current_device();
    SB_TRY(require_device(sm->device));
    DeviceGuard guard(sm->device);
    sm->total_rows = nrows;
    sm->total_cols = ncols;
    sm->memory_limit = memory_limit_mb * 1048576ull;  // :468
    // chunk size (:470-482): average entries per row from the triplet count, 12 bytes per entry + 4 per row, factor 2
    const uint64_t avg = nrows > 0 ? nt / nrows : 0;
    const uint64_t bytes_per_row = avg * (8 + 4) + 4;
    const uint64_t target = bytes_per_row > 0 ? std::max<uint64_t>(sm->memory_limit / (bytes_per_row * 2), 1) : 1000;
    sm->chunk_size = std::min(target, nrows);
    if (nrows == 0) return (*out = sm.release(), SB200_OK);
    const uint64_t nchunks = (nrows + sm->chunk_size - 1) / sm->chunk_size;

    // global CSR on the host: counting sort by row (stable), zeros dropped, stable sort of each row by column
    std::vector<uint64_t> rp(nrows + 1, 0);
    for (uint64_t i = 0; i < nt; i++)
        if (vals[i] != 0.0) rp[rows[i] + 1]++;
    for (uint64_t r = 0; r < nrows; r++) rp[r + 1] += rp[r];
    const uint64_t nnz = rp[nrows];
    sm->nnz = nnz;
    std::vector<uint32_t> ci(nnz);
    std::vector<double> cv(nnz);
    {
        std::vector<uint64_t> cur(rp.begin(), rp.end() - 1);
        for (uint64_t i = 0; i < nt; i++)
            if (vals[i] != 0.0) {
                const uint64_t p = cur[rows[i]]++;
                ci[p] = (uint32_t)cols[i];
                cv[p] = vals[i];
            }
    }
#pragma omp parallel
    {
        std::vector<std::pair<uint32_t, double>> tmp;
#pragma omp for schedule(dynamic, 1024)
        for (long long r = 0; r < (long long)nrows; r++) {
            const uint64_t s = rp[r], e = rp[r + 1];
            bool sorted = true;
            for (uint64_t k = s + 1; k < e; k++)
                if (ci[k - 1] > ci[k]) { sorted = false; break; }
            if (sorted) continue;
            tmp.resize(e - s);
            for (uint64_t k = s; k < e; k++) tmp[k - s] = {ci[k], cv[k]};
            std::stable_sort(tmp.begin(), tmp.end(),
                             [](const std::pair<uint32_t, double> &a, const std::pair<uint32_t, double> &b) { return a.first < b.first; });
            for (uint64_t k = s; k < e; k++) { ci[k] = tmp[k - s].first; cv[k] = tmp[k - s].second; }
        }
    }
    sm->chunks.resize(nchunks);
    for (uint64_t c = 0; c < nchunks; c++) {
        const uint64_t r0 = c * sm->chunk_size, r1 = std::min(nrows, r0 + sm->chunk_size);
        auto &ch = sm->chunks[c];
        ch.rows = r1 - r0;
        ch.nnz = rp[r1] - rp[r0];
        if (ch.nnz >= 0xFFFFFFF0ull) return fail(SB200_ERR_MEMORY_ALLOCATION, "chunk %llu holds too many entries", (unsigned long long)c);
        sm->max_chunk_nnz = std::max(sm->max_chunk_nnz, ch.nnz);
        SB_CUDA(cudaHostAlloc((void **)&ch.row_ptr, (ch.rows + 1) * sizeof(uint32_t), cudaHostAllocDefault));
        SB_CUDA(cudaHostAlloc((void **)&ch.cols, std::max<uint64_t>(ch.nnz, 1) * sizeof(uint32_t), cudaHostAllocDefault));
        SB_CUDA(cudaHostAlloc((void **)&ch.vals, std::max<uint64_t>(ch.nnz, 1) * sizeof(double), cudaHostAllocDefault));
        for (uint64_t r = r0; r <= r1; r++) ch.row_ptr[r - r0] = (uint32_t)(rp[r] - rp[r0]);
        if (ch.nnz) {
            memcpy(ch.cols, ci.data() + rp[r0], ch.nnz * sizeof(uint32_t));
            memcpy(ch.vals, cv.data() + rp[r0], ch.nnz * sizeof(double));
        }
    }
    *out = sm.release();
    return SB200_OK;
}

int32_t sb200_streaming_matrix_info(const sb200_streaming_matrix *sm, uint64_t *total_rows, uint64_t *total_cols,
                                    uint64_t *chunk_size, uint64_t *num_chunks, uint64_t *memory_usage) {
    if (!sm) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (total_rows) *total_rows = sm->total_rows;
    if (total_cols) *total_cols = sm->total_cols;
    if (chunk_size) *chunk_size = sm->chunk_size;
    if (num_chunks) *num_chunks = sm->chunks.size();
    if (memory_usage) {  // memory_usage (:551-558): nnz * 12 + rows * 4 per chunk
        uint64_t b = 0;
        for (auto &c : sm->chunks) b += c.nnz * 12 + c.rows * 4;
        *memory_usage = b;
    }
    return SB200_OK;
}

void sb200_streaming_matrix_free(sb200_streaming_matrix *sm) {
    if (!sm) return;
    DeviceGuard guard(sm->device);
    delete sm;
}

// multiply_vector_streaming (:531-548): callback(start_row, result slice) per chunk, in row order. A non-zero return of
// the callback stops the walk (extension; the reference's closure returns nothing).
int32_t sb200_streaming_matrix_multiply_vector(const sb200_streaming_matrix *sm, const double *x, uint64_t xlen,
                                               sb200_chunk_callback callback, void *user) {
    clear_error();
    if (!sm || !callback) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (xlen != sm->total_cols)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in matrix_vector_multiply",
                    (unsigned long long)sm->total_cols, (unsigned long long)xlen);
    if (xlen && !x) return fail(SB200_ERR_INVALID_INPUT, "null vector");
    if (sm->chunks.empty()) return SB200_OK;
    DeviceGuard guard(sm->device);
    SB_TRY(require_device(sm->device));
    struct Res {
        cudaStream_t copy = nullptr, comp = nullptr;
        cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
        double *h_y[2] = {nullptr, nullptr};
        ~Res() {
            for (int i = 0; i < 2; i++) {
                if (up[i]) cudaEventDestroy(up[i]);
                if (done[i]) cudaEventDestroy(done[i]);
                if (freed[i]) cudaEventDestroy(freed[i]);
                if (h_y[i]) cudaFreeHost(h_y[i]);
            }
            if (copy) cudaStreamDestroy(copy);
            if (comp) cudaStreamDestroy(comp);
        }
    } R;
    SB_CUDA(cudaStreamCreateWithFlags(&R.copy, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&R.comp, cudaStreamNonBlocking));
    DevBuf<double> d_x, d_vals[2], d_y[2];
    DevBuf<uint32_t> d_cols[2], d_rp[2];
    const size_t pad = ((sm->max_chunk_nnz + 3) & ~(size_t)3) + 8;  // the stream loads of the kernel may over-read
    SB_TRY(d_x.alloc(xlen));
    for (int i = 0; i < 2; i++) {
        SB_TRY(d_vals[i].alloc(pad));
        SB_TRY(d_cols[i].alloc(pad));
        SB_TRY(d_rp[i].alloc(sm->chunk_size + 1));
        SB_TRY(d_y[i].alloc(sm->chunk_size));
        SB_CUDA(cudaMemsetAsync(d_vals[i].p, 0, pad * sizeof(double), R.copy));
        SB_CUDA(cudaMemsetAsync(d_cols[i].p, 0, pad * sizeof(uint32_t), R.copy));
        SB_CUDA(cudaEventCreateWithFlags(&R.up[i], cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&R.done[i], cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&R.freed[i], cudaEventDisableTiming));
        SB_CUDA(cudaHostAlloc((void **)&R.h_y[i], std::max<uint64_t>(sm->chunk_size, 1) * sizeof(double), cudaHostAllocDefault));
    }
    SB_TRY(copy_h2d(d_x.p, x, xlen * 8, R.copy));
    SB_CUDA(cudaStreamSynchronize(R.copy));
    auto upload = [&](uint64_t c) -> int32_t {
        const int b = (int)(c & 1);
        const auto &ch = sm->chunks[c];
        SB_CUDA(cudaStreamWaitEvent(R.copy, R.freed[b], 0));  // the kernel that last read this staging set has finished
        SB_CUDA(cudaMemcpyAsync(d_rp[b].p, ch.row_ptr, (ch.rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, R.copy));
        if (ch.nnz) {
            SB_CUDA(cudaMemcpyAsync(d_cols[b].p, ch.cols, ch.nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, R.copy));
            SB_CUDA(cudaMemcpyAsync(d_vals[b].p, ch.vals, ch.nnz * sizeof(double), cudaMemcpyHostToDevice, R.copy));
        }
        SB_CUDA(cudaEventRecord(R.up[b], R.copy));
        return SB200_OK;
    };
    const uint64_t nchunks = sm->chunks.size();
    SB_TRY(upload(0));
    for (uint64_t c = 0; c < nchunks; c++) {
        const int b = (int)(c & 1);
        const auto &ch = sm->chunks[c];
        if (c + 1 < nchunks) SB_TRY(upload(c + 1));  // overlaps the kernel below
        SB_CUDA(cudaStreamWaitEvent(R.comp, R.up[b], 0));
        TileKernelArgs a{};
        a.vals = d_vals[b].p;
        a.cols = d_cols[b].p;
        a.row_ptr = d_rp[b].p;
        a.nrows = (uint32_t)ch.rows;
        a.xin = d_x.p;
        a.xin_own = d_x.p;
        a.xin_len = xlen;
        a.out = d_y[b].p;
        SB_TRY(launch_tile_kernel(-1, EPI_SPMV, a, R.comp));  // nslabs = 0, no SELL copy: the warp-stream CSR kernel
        SB_CUDA(cudaEventRecord(R.freed[b], R.comp));
        SB_CUDA(cudaMemcpyAsync(R.h_y[b], d_y[b].p, ch.rows * sizeof(double), cudaMemcpyDeviceToHost, R.comp));
        SB_CUDA(cudaEventRecord(R.done[b], R.comp));
        SB_CUDA(cudaEventSynchronize(R.done[b]));
        if (callback(c * sm->chunk_size, R.h_y[b], ch.rows, user) != 0) break;
    }
    SB_CUDA(cudaStreamSynchronize(R.copy));
    SB_CUDA(cudaStreamSynchronize(R.comp));
    return SB200_OK;
}

}  // extern "C"
