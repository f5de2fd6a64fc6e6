// matrix.hpp — the device-resident SparseMatrix handle (internal).
#pragma once

#include <atomic>
#include <functional>
#include <memory>

#include "common.hpp"

namespace sb200 {

constexpr unsigned long long kNone = ~0ull;

// Per-solve device vectors + pinned control mirror; pooled on the matrix so that concurrent solves on one
// handle each get their own (the reference's solve(&self, ..) is re-entrant: all mutable state is per call).
struct Workspace {
    uint64_t n_local = 0, n_full = 0;
    DevBuf<double> t[2];      // term ping-pong (full length n_full when distributed)
    DevBuf<double> x;         // solution (local rows)
    DevBuf<double> c;         // D^-1 b
    DevBuf<double> b;         // staged right-hand side (host API)
    DevBuf<double> tmp;       // A*x0 / initial guess staging
    DevBuf<double> partials;  // CTA partial sums
    DevBuf<double> norm_log;  // per-term norms (bare recurrence)
    DevBuf<LoopCtl> ctl;
    LoopCtl *h_ctl = nullptr;  // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~Workspace();
    int32_t ensure(uint64_t n_local, uint64_t n_full, size_t npartials);
    uint64_t bytes() const;
};

}  // namespace sb200

struct sb200_matrix {
    std::atomic<int> refcount{1};  // the creator's handle + one per live solver state (matrix_retain / matrix_release)
    int device = 0;
    uint64_t nrows = 0, ncols = 0, nnz = 0;
    int tile_cfg = 0;
    uint32_t ntiles = 0;
    std::vector<uint32_t> h_row_ptr;  // host copy, nrows + 1
    sb200::DevBuf<double> d_vals;
    sb200::DevBuf<uint32_t> d_cols;
    sb200::DevBuf<uint32_t> d_row_ptr;
    sb200::DevBuf<sb200::TileDesc> d_tiles;
    // SELL-32 copy of the same entries for the hot kernels (kernels.cu "the SELL-32 kernel"); empty when the padded
    // layout would cost more than 25 % extra slots (power-law rows) or a TMA tile configuration is selected
    sb200::DevBuf<uint32_t> d_sell_ptr;   // nblocks + 1 slab offsets
    sb200::DevBuf<uint32_t> d_sell_cols;  // sell_slabs * 32
    sb200::DevBuf<double> d_sell_vals;    // sell_slabs * 32
    uint64_t sell_slabs = 0;
    bool use_sell = false;
    // hub rows (> kLongRow entries): list + chunk table for the pre-pass of launch_tile_kernel (kernels.cu long_rows_*)
    uint32_t nlong = 0, nlong_chunks = 0;
    sb200::DevBuf<uint32_t> d_long_rows, d_long_first;
    sb200::DevBuf<uint2> d_long_chunks;
    // column-slab layout for the hot kernels (kernels_slab.cu): when the gather source (8 * ncols bytes) does not fit the
    // L2 partition of a die, the entries are regrouped into nslabs column ranges of slab_width columns, stored slab-major
    // in one pair of arrays (CSR order inside a slab); per slab a u32 entry offset per 32-row block and a u16 length per
    // row. One fused launch walks the slabs, row sums carried over in column order (bit-identical results)
    int nslabs = 0;
    uint32_t slab_width = 0;
    sb200::DevBuf<double> d_slab_vals;
    sb200::DevBuf<uint32_t> d_slab_cols;
    sb200::DevBuf<uint32_t> d_slab_blk;   // nslabs * (nblocks + 1)
    sb200::DevBuf<uint16_t> d_slab_rel;    // nslabs * slab_rel_stride
    uint64_t slab_rel_stride = 0;
    uint64_t slab_entries = 0;            // entries stored in the slab arrays (incl. alignment padding between slabs)
    cudaStream_t stream = nullptr;  // for host-pointer entry points

    // distributed: this handle holds rows [row_base, row_base + nrows) of an n_global-square system
    uint64_t row_base = 0;
    uint64_t n_global = 0;
    bool distributed = false;

    // lazily computed analysis (K4), cached: the matrix is immutable apart from scale()
    std::mutex mu;
    bool analysed[2] = {false, false};  // per solve mode (diagonal extraction differs on duplicates)
    bool col_analysed = false;
    sb200::DevBuf<double> d_dinv[2];
    unsigned long long first_bad_dd = sb200::kNone, first_bad_diag[2] = {sb200::kNone, sb200::kNone};
    unsigned long long first_bad_col = sb200::kNone;
    double min_factor = 0.0;
    bool has_factor = false;

    std::vector<std::unique_ptr<sb200::Workspace>> pool;

    // columns of A + diagonal for the A x = b forward push (csrc/push.cu), built on first use
    void *axb_cache = nullptr;
    void (*axb_cache_free)(void *) = nullptr;

    ~sb200_matrix();
};

namespace sb200 {

// ingest helpers (matrix.cu)
int32_t matrix_from_host_csr(const uint64_t *row_ptr64, const uint32_t *row_ptr32, const uint32_t *cols,
                             const double *vals, uint64_t nrows, uint64_t ncols, uint64_t nnz, bool validate,
                             sb200_matrix **out, bool allow_slabs = true);
// reduce_cols (row-partitioned handles): sums the per-column |diagonal| and off-diagonal accumulators over all ranks
// before the column-dominance test, so that every rank sees whole columns
using ColReduce = std::function<int32_t(double *col_diag, double *col_off, uint64_t ncols, cudaStream_t st)>;
int32_t matrix_finish(sb200_matrix *m, bool allow_slabs);
int32_t matrix_from_device_csr(const uint32_t *d_row_ptr, const uint32_t *d_cols, const double *d_vals, uint64_t nrows,
                               uint64_t ncols, uint64_t nnz, cudaStream_t src_stream, sb200_matrix **out);
// COO -> CSR on the device (csrc/ingest.cu): SparseMatrix::from_triplets semantics (zeros dropped, stable (row, col) order,
// duplicates kept or summed in triplet order). Host slices, already validated.
int32_t matrix_from_triplets_device(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t nt,
                                    uint64_t nrows, uint64_t ncols, int dup_policy, sb200_matrix **out);
int32_t validate_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t nt, uint64_t nrows,
                          uint64_t ncols);
int32_t matrix_analyse(sb200_matrix *m, int mode, bool need_cols, const ColReduce &reduce_cols = ColReduce());
int32_t matrix_spmv_dev(const sb200_matrix *m, const double *x_dev, double *y_dev, int accumulate, cudaStream_t st);
void matrix_retain(sb200_matrix *m);
void matrix_release(sb200_matrix *m);  // deletes the handle when the last owner lets go
std::unique_ptr<Workspace> matrix_acquire_ws(sb200_matrix *m);
void matrix_release_ws(sb200_matrix *m, std::unique_ptr<Workspace> ws);
void fill_tile_args(const sb200_matrix *m, TileKernelArgs &a);
// kernel launches of one pass over the matrix: the hub-row pre-pass (2 launches) precedes the row kernel where it applies
inline uint64_t launches_per_pass(const sb200_matrix *m) { return m->nlong > 0 && !m->use_sell ? 3 : 1; }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace sb200
