// ingest.cu — COO -> CSR and the PageRank system builder on the device.
//
// SparseMatrix::from_triplets (ref src/matrix/mod.rs:160-199) -> COOStorage::from_triplets (sparse.rs:528-548: exact
// zeros dropped) -> CSRStorage::from_coo (sparse.rs:80-132: STABLE sort by (row, col), duplicates kept as separate
// entries) used to run on the host (counting sort + per-row stable sorts: 13 s for the 110 M triplets of the C3 PageRank
// system). Here: one 64-bit key (row << 32 | col) per triplet, dropped zeros keyed past every real key, one stable LSD
// radix sort of (key, value) pairs (cub::DeviceRadixSort — library code on a one-off ingest step, not on the iteration
// path), row_ptr from the sorted keys, and for SB200_DUP_SUM one thread per distinct (row, col) adding its run in triplet
// order (the order the host loop added them in: results are bit-identical to the host path and to the oracle).
// computePageRank (ref src/core/solver.ts:664-722) builds its triplets on the device too: out-degrees, then
// S = I - alpha P^T as (i, i, 1) followed by (dst, src, -(alpha (w / outdeg[src]))).
// Validation (bounds, finite values) stays on the host, in triplet order, so that input errors are reported without a GPU.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstring>

#include "matrix.hpp"

namespace sb200 {

namespace {

constexpr int kT = 256;
constexpr unsigned long long kDropped = ~0ull;  // key of a dropped (exact zero) triplet: sorts behind every real key

unsigned grid_of(uint64_t n) {
    uint64_t g = (n + kT - 1) / kT;
    return (unsigned)(g > 148ull * 32 ? 148ull * 32 : (g ? g : 1));
}

__global__ void __launch_bounds__(kT) make_keys_kernel(const uint64_t *__restrict__ rows, const uint64_t *__restrict__ cols,
                                                       const double *__restrict__ vals, uint64_t nt,
                                                       unsigned long long *__restrict__ keys, unsigned long long *nzeros) {
    unsigned long long z = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i < nt; i += (uint64_t)gridDim.x * kT) {
        const bool zero = vals[i] == 0.0;  // COOStorage::from_triplets keeps `value != 0` (sparse.rs:535-541)
        keys[i] = zero ? kDropped : ((unsigned long long)rows[i] << 32 | (unsigned long long)cols[i]);
        z += zero;
    }
    if (z) atomicAdd(nzeros, z);
}

// row_ptr[r] = first sorted entry whose row is >= r; entry i closes the rows (row(i-1), row(i)]
__global__ void __launch_bounds__(kT) row_ptr_kernel(const unsigned long long *__restrict__ keys, uint64_t nnz, uint64_t nrows,
                                                     uint32_t *__restrict__ row_ptr) {
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i <= nnz; i += (uint64_t)gridDim.x * kT) {
        const uint64_t r_prev = i == 0 ? 0 : (uint64_t)(keys[i - 1] >> 32) + 1;
        const uint64_t r_cur = i == nnz ? nrows : (uint64_t)(keys[i] >> 32);
        for (uint64_t r = r_prev; r <= r_cur; r++) row_ptr[r] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(kT) split_keys_kernel(const unsigned long long *__restrict__ keys, uint64_t nnz,
                                                        uint32_t *__restrict__ cols) {
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * kT)
        cols[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
}

// SB200_DUP_SUM: the head of every run of equal keys adds the run left to right (= triplet order: the sort is stable);
// flag = 1 for heads whose sum is not an exact zero
__global__ void __launch_bounds__(kT) dup_sum_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
                                                     uint64_t nnz, double *__restrict__ sums, uint32_t *__restrict__ flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * kT) {
        uint32_t keep = 0;
        if (i == 0 || keys[i] != keys[i - 1]) {
            const unsigned long long k = keys[i];
            double acc = vals[i];
            for (uint64_t j = i + 1; j < nnz && keys[j] == k; j++) acc += vals[j];
            sums[i] = acc;
            keep = acc != 0.0;
        }
        flag[i] = keep;
    }
}

__global__ void __launch_bounds__(kT) compact_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ sums,
                                                     const uint32_t *__restrict__ flag, const uint32_t *__restrict__ slot,
                                                     uint64_t nnz, unsigned long long *__restrict__ okeys,
                                                     double *__restrict__ ovals) {
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i < nnz; i += (uint64_t)gridDim.x * kT)
        if (flag[i]) {
            okeys[slot[i]] = keys[i];
            ovals[slot[i]] = sums[i];
        }
}

// sorted (key, value) pairs on the device -> handle. keys / vals: nt entries, the first nnz are real.
int32_t finish_from_sorted(unsigned long long *keys, double *vals, uint64_t nnz, uint64_t nrows, uint64_t ncols,
                           int dup_policy, cudaStream_t st, sb200_matrix **out) {
    DevBuf<unsigned long long> ckeys;
    DevBuf<double> cvals, sums;
    if (dup_policy == SB200_DUP_SUM && nnz > 0) {
        if (nnz >= 0xFFFFFFF0ull) return fail(SB200_ERR_MEMORY_ALLOCATION, "too many triplets for 32-bit slots");
        DevBuf<uint32_t> flag, slot;
        SB_TRY(sums.alloc(nnz));
        SB_TRY(flag.alloc(nnz + 1));
        SB_TRY(slot.alloc(nnz + 1));
        dup_sum_kernel<<<grid_of(nnz), kT, 0, st>>>(keys, vals, nnz, sums.p, flag.p);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaMemcpyAsync(slot.p, flag.p, nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        uint64_t kept = 0;
        SB_TRY(device_exclusive_scan_u32(slot.p, nnz, &kept, st));
        SB_TRY(ckeys.alloc(kept));
        SB_TRY(cvals.alloc(kept));
        compact_kernel<<<grid_of(nnz), kT, 0, st>>>(keys, sums.p, flag.p, slot.p, nnz, ckeys.p, cvals.p);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaStreamSynchronize(st));
        keys = ckeys.p;
        vals = cvals.p;
        nnz = kept;
    }
    if (nnz >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_MEMORY_ALLOCATION,
                    "nnz %llu does not fit the u32 row_ptr of CSRStorage (src/matrix/sparse.rs:22); "
                    "row-partition the system across GPUs",
                    (unsigned long long)nnz);
    DevBuf<uint32_t> row_ptr, cols;
    SB_TRY(row_ptr.alloc(nrows + 1));
    SB_TRY(cols.alloc(nnz));
    row_ptr_kernel<<<grid_of(nnz + 1), kT, 0, st>>>(keys, nnz, nrows, row_ptr.p);
    split_keys_kernel<<<grid_of(nnz), kT, 0, st>>>(keys, nnz, cols.p);
    SB_CUDA(cudaGetLastError());
    return matrix_from_device_csr(row_ptr.p, cols.p, vals, nrows, ncols, nnz, st, out);
}

// stable sort of nt (key, value) pairs; the real entries end up in front of the dropped ones
int32_t sort_pairs(DevBuf<unsigned long long> &keys, DevBuf<double> &vals, uint64_t nt, cudaStream_t st) {
    if (nt == 0) return SB200_OK;
    if (nt >= 0x7FFFFFF0ull) return fail(SB200_ERR_MEMORY_ALLOCATION, "too many triplets for one device sort (%llu)", (unsigned long long)nt);
    DevBuf<unsigned long long> k2;
    DevBuf<double> v2;
    DevBuf<char> tmp;
    SB_TRY(k2.alloc(nt));
    SB_TRY(v2.alloc(nt));
    size_t bytes = 0;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, k2.p, vals.p, v2.p, (int)nt, 0, 64, st));
    SB_TRY(tmp.alloc(bytes));
    SB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, k2.p, vals.p, v2.p, (int)nt, 0, 64, st));
    SB_CUDA(cudaStreamSynchronize(st));
    std::swap(keys.p, k2.p);
    std::swap(keys.n, k2.n);
    std::swap(vals.p, v2.p);
    std::swap(vals.n, v2.n);
    return SB200_OK;
}

struct StreamGuard {
    cudaStream_t s = nullptr;
    ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};

// ---- PageRank triplets ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) outdeg_unit_kernel(const uint64_t *__restrict__ src, uint64_t ne, double *outdeg) {
    // unit weights: the sums are small integers, exact in any order
    for (uint64_t e = blockIdx.x * (uint64_t)kT + threadIdx.x; e < ne; e += (uint64_t)gridDim.x * kT) atomicAdd(outdeg + src[e], 1.0);
}

// weighted edges: (src, w) sorted by src (stable), the head of every run adds it in edge order
__global__ void __launch_bounds__(kT) outdeg_runs_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ w,
                                                         uint64_t ne, double *__restrict__ outdeg) {
    for (uint64_t i = blockIdx.x * (uint64_t)kT + threadIdx.x; i < ne; i += (uint64_t)gridDim.x * kT)
        if (i == 0 || keys[i] != keys[i - 1]) {
            const unsigned long long k = keys[i];
            double acc = 0.0;
            for (uint64_t j = i; j < ne && keys[j] == k; j++) acc += w[j];
            outdeg[k] = acc;
        }
}

__global__ void __launch_bounds__(kT) src_keys_kernel(const uint64_t *__restrict__ src, uint64_t ne, unsigned long long *keys) {
    for (uint64_t e = blockIdx.x * (uint64_t)kT + threadIdx.x; e < ne; e += (uint64_t)gridDim.x * kT) keys[e] = src[e];
}

__global__ void __launch_bounds__(kT) pagerank_triplets_kernel(const uint64_t *__restrict__ src, const uint64_t *__restrict__ dst,
                                                               const double *__restrict__ w, uint64_t ne, uint64_t n,
                                                               const double *__restrict__ outdeg, double alpha,
                                                               unsigned long long *__restrict__ keys, double *__restrict__ vals,
                                                               unsigned long long *nzeros) {
    unsigned long long z = 0;
    for (uint64_t t = blockIdx.x * (uint64_t)kT + threadIdx.x; t < n + ne; t += (uint64_t)gridDim.x * kT) {
        if (t < n) {  // S = I first (solver.ts:686-688): the diagonal precedes a self loop's contribution in the sum
            keys[t] = (unsigned long long)t << 32 | (unsigned long long)t;
            vals[t] = 1.0;
            continue;
        }
        const uint64_t e = t - n, j = src[e], i = dst[e];
        const double od = outdeg[j];
        // S[i][j] -= alpha * adj[j][i] / outdeg[j] where outdeg[j] > 0 (:689-698)
        const double v = od > 0.0 ? -(alpha * ((w ? w[e] : 1.0) / od)) : 0.0;
        const bool zero = v == 0.0;
        keys[t] = zero ? kDropped : ((unsigned long long)i << 32 | (unsigned long long)j);
        vals[t] = v;
        z += zero;
    }
    if (z) atomicAdd(nzeros, z);
}

}  // namespace

// validate in triplet order (src/matrix/mod.rs:166-187): the first offending triplet decides the error
int32_t validate_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t nt, uint64_t nrows,
                          uint64_t ncols) {
    if (nt && (!rows || !cols || !vals)) return fail(SB200_ERR_INVALID_INPUT, "null triplet slice");
    long long first = (long long)nt;
#pragma omp parallel for reduction(min : first) schedule(static)
    for (long long i = 0; i < (long long)nt; i++)
        if (rows[i] >= nrows || cols[i] >= ncols || !std::isfinite(vals[i])) first = std::min(first, i);
    if (first >= (long long)nt) return SB200_OK;
    const uint64_t i = (uint64_t)first;
    if (rows[i] >= nrows)
        return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "row index %llu out of bounds (max %llu) in triplet %llu",
                    (unsigned long long)rows[i], (unsigned long long)(nrows ? nrows - 1 : 0), (unsigned long long)i);
    if (cols[i] >= ncols)
        return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "column index %llu out of bounds (max %llu) in triplet %llu",
                    (unsigned long long)cols[i], (unsigned long long)(ncols ? ncols - 1 : 0), (unsigned long long)i);
    return fail(SB200_ERR_INVALID_INPUT, "Non-finite value %g at (%llu, %llu)", vals[i], (unsigned long long)rows[i],
                (unsigned long long)cols[i]);
}

int32_t matrix_from_triplets_device(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t nt,
                                    uint64_t nrows, uint64_t ncols, int dup_policy, sb200_matrix **out) {
    if (nrows >= 0xFFFFFFF0ull || ncols >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_INVALID_INPUT, "dimension exceeds the u32 IndexType of the reference (src/types.rs:22)");
    const int dev = current_device();
    SB_TRY(require_device(dev));
    DeviceGuard guard(dev);
    StreamGuard sg;
    SB_CUDA(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
    cudaStream_t st = sg.s;
    DevBuf<unsigned long long> keys, nz;
    DevBuf<double> dv;
    SB_TRY(keys.alloc(nt));
    SB_TRY(dv.alloc(nt));
    SB_TRY(nz.alloc(1));
    SB_CUDA(cudaMemsetAsync(nz.p, 0, 8, st));
    {
        DevBuf<uint64_t> dr, dc;  // the raw index slices only live until the keys exist
        SB_TRY(dr.alloc(nt));
        SB_TRY(dc.alloc(nt));
        SB_TRY(copy_h2d(dr.p, rows, nt * 8, st));
        SB_TRY(copy_h2d(dc.p, cols, nt * 8, st));
        SB_TRY(copy_h2d(dv.p, vals, nt * 8, st));
        make_keys_kernel<<<grid_of(nt), kT, 0, st>>>(dr.p, dc.p, dv.p, nt, keys.p, nz.p);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaStreamSynchronize(st));
    }
    unsigned long long nzeros = 0;
    SB_CUDA(cudaMemcpy(&nzeros, nz.p, 8, cudaMemcpyDeviceToHost));
    SB_TRY(sort_pairs(keys, dv, nt, st));
    return finish_from_sorted(keys.p, dv.p, nt - nzeros, nrows, ncols, dup_policy, st, out);
}

}  // namespace sb200

using namespace sb200;

extern "C" {

// computePageRank (ref src/core/solver.ts:664-722): outdeg[i] = sum_j adj[i][j] (:679-684); S = I, then
// S[i][j] -= alpha * adj[j][i] / outdeg[j] where outdeg[j] > 0 (:689-698) — dangling rows contribute nothing;
// rhs = (1 - alpha)/n (:708).  adj is dense in the reference (one number per pair), so repeated edges and a
// self loop's contribution to the diagonal are merged by summation (SB200_DUP_SUM), in edge order.
int32_t sb200_pagerank_system(const uint64_t *src, const uint64_t *dst, const double *w, uint64_t nedges, uint64_t n,
                              double alpha, sb200_matrix **S, double *rhs) {
    clear_error();
    if (!S) return fail(SB200_ERR_INVALID_INPUT, "S is null");
    *S = nullptr;
    if (!(alpha >= 0.0 && alpha <= 1.0)) return fail(SB200_ERR_INVALID_INPUT, "damping must be in [0, 1]");  // validateRange (:666)
    if (nedges && (!src || !dst)) return fail(SB200_ERR_INVALID_INPUT, "null edge list");
    if (n >= 0xFFFFFFF0ull) return fail(SB200_ERR_INVALID_INPUT, "dimension exceeds the u32 IndexType of the reference (src/types.rs:22)");
    long long first = (long long)nedges;
#pragma omp parallel for reduction(min : first) schedule(static)
    for (long long e = 0; e < (long long)nedges; e++)
        if (src[e] >= n || dst[e] >= n || (w && !std::isfinite(w[e]))) first = std::min(first, e);
    if (first < (long long)nedges) {
        const uint64_t e = (uint64_t)first;
        if (src[e] >= n || dst[e] >= n)
            return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "edge %llu (%llu -> %llu) out of bounds for %llu nodes",
                        (unsigned long long)e, (unsigned long long)src[e], (unsigned long long)dst[e], (unsigned long long)n);
        return fail(SB200_ERR_INVALID_INPUT, "non-finite weight on edge %llu", (unsigned long long)e);
    }
    const int dev = current_device();
    SB_TRY(require_device(dev));
    DeviceGuard guard(dev);
    StreamGuard sg;
    SB_CUDA(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
    cudaStream_t st = sg.s;
    const uint64_t nt = n + nedges;
    DevBuf<unsigned long long> keys, nz;
    DevBuf<double> vals;
    SB_TRY(keys.alloc(nt));
    SB_TRY(vals.alloc(nt));
    SB_TRY(nz.alloc(1));
    SB_CUDA(cudaMemsetAsync(nz.p, 0, 8, st));
    {
        DevBuf<uint64_t> ds, dd;
        DevBuf<double> dw, outdeg;
        SB_TRY(ds.alloc(nedges));
        SB_TRY(dd.alloc(nedges));
        SB_TRY(outdeg.alloc(n));
        SB_TRY(copy_h2d(ds.p, src, nedges * 8, st));
        SB_TRY(copy_h2d(dd.p, dst, nedges * 8, st));
        SB_CUDA(cudaMemsetAsync(outdeg.p, 0, std::max<uint64_t>(n, 1) * 8, st));
        if (w) {
            // weighted: the row sums are added in edge order (the reference's loop order), not by atomics
            SB_TRY(dw.alloc(nedges));
            SB_TRY(copy_h2d(dw.p, w, nedges * 8, st));
            DevBuf<unsigned long long> sk;
            DevBuf<double> sw;
            SB_TRY(sk.alloc(nedges));
            SB_TRY(sw.alloc(nedges));
            src_keys_kernel<<<grid_of(nedges), kT, 0, st>>>(ds.p, nedges, sk.p);
            SB_CUDA(cudaMemcpyAsync(sw.p, dw.p, nedges * 8, cudaMemcpyDeviceToDevice, st));
            SB_CUDA(cudaGetLastError());
            SB_TRY(sort_pairs(sk, sw, nedges, st));
            outdeg_runs_kernel<<<grid_of(nedges), kT, 0, st>>>(sk.p, sw.p, nedges, outdeg.p);
        } else {
            outdeg_unit_kernel<<<grid_of(nedges), kT, 0, st>>>(ds.p, nedges, outdeg.p);
        }
        SB_CUDA(cudaGetLastError());
        pagerank_triplets_kernel<<<grid_of(nt), kT, 0, st>>>(ds.p, dd.p, w ? dw.p : nullptr, nedges, n, outdeg.p, alpha, keys.p,
                                                             vals.p, nz.p);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaStreamSynchronize(st));
    }
    unsigned long long nzeros = 0;
    SB_CUDA(cudaMemcpy(&nzeros, nz.p, 8, cudaMemcpyDeviceToHost));
    SB_TRY(sort_pairs(keys, vals, nt, st));
    SB_TRY(finish_from_sorted(keys.p, vals.p, nt - nzeros, n, n, SB200_DUP_SUM, st, S));
    if (rhs)
        for (uint64_t i = 0; i < n; i++) rhs[i] = (1.0 - alpha) / (double)n;
    return SB200_OK;
}

}  // extern "C"
