// Internal declarations shared by the translation units of libsublinear_b200.so.
// Nothing here is part of the ABI (include/sublinear_b200.h is).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sublinear_b200.h"

namespace sb200 {

// ---- error plumbing: SolverError code + message, per thread (no exceptions cross the ABI) --------------
int32_t fail(int32_t code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();

#define SB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::sb200::fail(_e == cudaErrorMemoryAllocation ? SB200_ERR_MEMORY_ALLOCATION          \
                                                                 : SB200_ERR_ALGORITHM,                 \
                                 "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), __FILE__, __LINE__, \
                                 cudaGetErrorString(_e));                                               \
    } while (0)

#define SB_TRY(expr)                  \
    do {                              \
        int32_t _rc = (expr);         \
        if (_rc != SB200_OK) return _rc; \
    } while (0)

int current_device();           // device new handles are created on (thread-local, env default)
int32_t require_device(int dev); // cudaSetDevice + "no CPU fallback" check

// ---- device buffers --------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int32_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(SB200_ERR_MEMORY_ALLOCATION, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T),
                        cudaGetErrorString(e));
        }
        n = count;
        return SB200_OK;
    }
};

// host<->device copies that take pageable OR pinned host pointers (pageable staged through a pinned ring)
int32_t copy_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);
int32_t copy_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);

// pinned result buffers are recycled (runtime.cu): returns null on allocation failure
void *pinned_pool_get(size_t bytes);
void pinned_pool_put(void *p);

// ---- kernel-side contracts -------------------------------------------------------------------------------
// Tile table entry: tile t covers rows [desc[t].row0, desc[t+1].row0) and nnz [desc[t].nnz0, desc[t+1].nnz0).
struct TileDesc {
    uint32_t row0;
    uint32_t nnz0;
};

// Tile geometry (one instantiation of the kernel template each).
struct TileCfg {
    int threads;  // CTA size
    int rows;     // max rows per tile (<= threads)
    int cap;      // max nnz staged per tile; a single row above it is a "long row" tile
};
constexpr int kNumTileCfgs = 6;
extern const TileCfg kTileCfgs[kNumTileCfgs];
int default_tile_cfg();

// Device-resident loop state, mirrors the scalars of NeumannState (src/solver/neumann.rs:97-135) that the
// control flow of NeumannSolver::solve (:469-555) reads. Updated by the last CTA of each kernel.
struct LoopCtl {
    double red[2];       // row-partitioned runs: this rank's partial sums, all-reduced in place
    double term_norm2;   // ||t_k||_2^2 of the latest term
    double aux_norm2;    // ||D o t_k||^2 (identity residual)
    double res_norm2;    // latest ||A x - rhs||_2^2
    double res_norm;     // sqrt(res_norm2)  (+inf before the first residual, neumann.rs:236)
    double rhs_norm2;    // ||D^-1 b||^2 (error bounds, neumann.rs:329-331)
    double tolerance, series_tolerance;
    uint32_t max_terms, max_iterations;
    uint32_t alive;      // loop still running
    uint32_t sconv;      // series_converged
    uint32_t nonfinite;  // NumericalInstability seen
    uint32_t terms;      // terms_computed
    uint32_t iterations;
    uint32_t ticket;     // last-CTA election
    uint32_t xchg;       // row-partitioned P2P runs: number of exchanges consumed so far (drives epoch + slot parity)
    uint32_t peer_timeout;  // a peer never signalled (the loop is stopped and reported as AlgorithmError)
    // conjugate gradient on the same kernels (csrc/cg.cu; ref src/optimized_solver.rs:182-295): the scalars of the loop
    double cg_rsold;     // r^T r of the current residual
    double cg_pap;       // p^T A p of the latest SpMV
    double cg_alpha, cg_beta;
    double cg_tol_sq;    // tolerance^2 (:210)
    uint32_t cg_converged, cg_breakdown;  // `rsold <= tolerance_sq` seen (:218-221) / |p^T A p| < 1e-16 (:234-236)
    uint32_t cg_matvecs, cg_pad;
};

// Row-partitioned runs over peer memory (csrc/dist.cu): every rank maps every other rank's exchange arena (CUDA IPC);
// a kernel stores its slice of the new term (and, when a residual check follows, of the solution) straight into all
// peers' buffers over NVLink, its last CTA publishes the rank's partial sums + a flag to all peers, and a one-warp
// wait kernel sums the partials in rank order (bitwise identical on every rank) and takes the loop decision.
constexpr int kMaxPeers = 8;
struct PeerExchange {               // all zero = not used (single GPU, or the NCCL path)
    int world, rank;
    unsigned long long epoch_base;  // exchange e of this solve signals epoch_base + e
    double *t_out[kMaxPeers];       // peer p's term buffer being written this launch (global row indexing)
    double *x_out[kMaxPeers];       // peer p's full-length solution mirror, or null when x is not published
    double *slots[kMaxPeers];       // peer p's partial-sum slots [2 parities][world][2]
    unsigned long long *flags[kMaxPeers];  // peer p's flag array [world]
};

constexpr int kMaxSlabs = 8;  // column slabs of the gather source (matrix.cu build_slabs)
constexpr uint32_t kLongRow = 1024;    // rows above this length leave the ordered per-lane sums (warp-stream kernel)
constexpr uint32_t kLongChunk = 8192;  // entries of a long row summed by one CTA of the pre-pass

enum Epilogue { EPI_SPMV = 0, EPI_PUSH = 1, EPI_RESID = 2, EPI_CG = 3 };  // EPI_CG: out = A p and sum p_i (A p)_i (warp-stream kernel only)

struct TileKernelArgs {
    // CSR (device)
    const double *vals;
    const uint32_t *cols;
    const uint32_t *row_ptr;
    const TileDesc *tiles;
    uint32_t ntiles;
    uint32_t nrows;
    // the same matrix in SELL-32 slabs (kernels.cu "the SELL-32 kernel"); sell_ptr == null: CSR kernels only
    const uint32_t *sell_ptr;   // nblocks + 1 slab offsets (slab = 32 slots, one per row of the block)
    const uint32_t *sell_cols;
    const double *sell_vals;
    // the same matrix regrouped into column slabs (matrix.cu build_slabs): nslabs > 1 -> launch_tile_kernel runs the fused
    // slab kernel (kernels_slab.cu): one launch walks slab after slab, the row sums carry over through `acc`.
    // Entries are stored slab-major in one pair of arrays; inside a slab in CSR order. Per slab s and 32-row block b:
    // slab_blk[s * (nblocks + 1) + b] = index of the block's first entry, slab_rel[s * slab_rel_stride + row] = offset of
    // the row's first entry inside its block (u16; bit 15 marks a hub row, summed by the long_rows_* pre-pass)
    int nslabs;
    const double *slab_vals;
    const uint32_t *slab_cols;
    const uint32_t *slab_blk;
    const uint16_t *slab_rel;
    uint64_t slab_rel_stride;
    int acc_keep;               // the carried row sums fit the L2 next to the slab of the vector: do not mark them evict-first
    // rows with more than kLongRow entries (hub rows of power-law graphs): launch_tile_kernel lets the whole grid compute
    // their sums first (kernels.cu long_rows_*), the row-block kernel only looks them up
    uint32_t nlong;             // number of long rows of this matrix (0: none)
    uint32_t nlong_chunks;
    const uint32_t *long_rows;  // their local row ids, ascending
    const uint32_t *long_first; // nlong + 1: first chunk of every long row
    const uint2 *long_chunks;   // {begin, end} entry ranges of at most kLongChunk entries
    const double *long_sum;     // set by the launcher: (A xin)_row of the long rows, in list order
    double *acc;           // n doubles of scratch for the partial row sums; null: `out` carries them
    // vectors
    const double *xin;    // gather source (term / solution / x)
    const double *xin_own; // value of xin for local row i is xin_own[i] (== xin + row_base)
    uint32_t row_base;    // global column index of local row 0 (0 unless row-partitioned)
    uint64_t xin_len;     // length of the gather source (matrix columns)
    double *out;          // SPMV: y ; PUSH: new term (indexed by local row)
    double *sol;          // PUSH: solution (read+write)
    const double *dinv;   // PUSH
    const double *rhs;    // RESID: subtracted vector
    int accumulate;       // SPMV: y += A x
    // loop control
    LoopCtl *ctl;         // may be null for SPMV
    double *partials;     // gridDim doubles (x2 when aux)
    uint32_t it;          // iteration index this launch belongs to
    int last_in_iter;     // run end-of-iteration logic in the tail
    int force;            // ignore ctl->alive (final residual / bare recurrence)
    int identity_res;     // PUSH: also accumulate ||D o t'||^2
    int defer_tail;       // distributed: only publish the local sums; a later kernel runs the loop logic
    double *norm_log;     // optional: norm_log[it] = ||t_it||^2 (bare recurrence)
    unsigned long long *phase_log;  // debug ($SUBLINEAR_B200_PHASE_LOG=1): per-phase cycles of thread 0, summed over CTAs
    PeerExchange px;      // row-partitioned P2P exchange (warp-stream kernel only)
};

// launchers (kernels.cu; ingest passes in kernels_ingest.cu, vector passes in kernels_vec.cu). grid = 0 -> persistent grid sized from occupancy.
int32_t launch_tile_kernel(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream);
int tile_kernel_max_grid(int cfg, Epilogue epi);
// setup pass over the CSR (K4): per-row dominance, diagonal and its inverse
struct SetupOut {
    double *dinv;                  // n
    unsigned long long *first_bad_dd;    // first row violating row dominance (or ~0)
    unsigned long long *first_bad_diag;  // first row with missing / ~zero diagonal (or ~0)
    double *col_diag, *col_off;    // optional column accumulators (n each) for column dominance
    double *min_factor_bits;       // optional: min diag/off ratio (as double, atomicMin on ordered bits)
};
int32_t launch_setup_rows(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                          uint32_t row_base, int compat_diag, SetupOut out, cudaStream_t stream);
// column-slab split at ingest (warp per 32-row block): per (row, slab) the u16 offset of the row's first entry inside its
// block (bit 15 = hub row, left to the pre-pass) and per (block, slab) the u32 entry count; flags[0] receives 1 if some row is not sorted by column (then the split would change the
// accumulation order and is not used), flags[1] the number of entries in rows that reach into more than one slab. After the
// prefix sums over blk (per slab, plus the slab's base), the ordered fill.
int32_t launch_slab_count(const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows, uint32_t slab_width, int nslabs,
                          uint32_t long_row, uint16_t *rel, uint64_t rel_stride, uint32_t *blk, unsigned long long *flags,
                          cudaStream_t stream);
int32_t launch_slab_fill(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                         uint32_t slab_width, int nslabs, const uint16_t *rel, uint64_t rel_stride, const uint32_t *blk,
                         uint32_t *slab_cols, double *slab_vals, cudaStream_t stream);
int32_t launch_add_u32(uint32_t *data, uint64_t n, uint32_t v, cudaStream_t stream);
// exclusive prefix sum in place over data[0..n); data[n] and *total receive the sum (multi-CTA: tile sums, scan of the
// tile sums, tile-local scan)
int32_t device_exclusive_scan_u32(uint32_t *data, uint64_t n, uint64_t *total, cudaStream_t stream);
// the fused column-slab kernel (kernels_slab.cu)
int32_t launch_slab_kernel(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out);
int32_t launch_csr_to_sell(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                           const uint32_t *sell_ptr, uint32_t *sell_cols, double *sell_vals, cudaStream_t stream);
int32_t launch_col_dominance(const double *col_diag, const double *col_off, uint32_t n, unsigned long long *first_bad,
                             cudaStream_t stream);
// iteration 0 (neumann.rs:191-211 + first compute_next_term): c = b*dinv; t = c (or (b - Ax0)*dinv); x = base + t
struct InitArgs {
    const double *b, *dinv;
    const double *x0;      // initial guess or null
    const double *ax0;     // A*x0 (correct mode with guess) or null
    double *c_out;         // D^-1 b (kept for ref_compat residual / error bounds); may be null
    double *t_out, *x_out;
    uint32_t n;
    int compat;            // ref_compat: x = (x0 or c) + c ; correct: x = (x0 or 0) + t0
    LoopCtl *ctl;
    double *partials;
    int last_in_iter;
    int identity_res;
    int defer_tail;
    int skip_term0;        // max_terms == 0 / max_iterations == 0: x = base, no term accumulated
    double *norm_log;
    uint32_t row_base;     // global index of local row 0 (P2P publishing)
    PeerExchange px;
};
int32_t launch_init_state(const InitArgs &a, cudaStream_t stream);
int init_state_grid();
int32_t launch_scale(double *v, uint64_t n, double factor, cudaStream_t stream);
// SolverAlgorithm state interface (state.cu): op 0 = accumulate term 0 (x += t, ||t||^2 as a term), op 1 = ||t||^2 -> red[0]
int32_t launch_state_vec(int op, const double *t, double *x, uint64_t n, LoopCtl *ctl, double *partials, cudaStream_t stream);
int32_t launch_update_rhs(const uint64_t *idx, const double *delta, uint64_t count, const double *dinv, double *b,
                          double *rhs, double *also, cudaStream_t stream);
// distributed: finish the loop logic after the partial norms were all-reduced (kind: 1 = term, 2 = residual)
int32_t launch_dist_tail(LoopCtl *ctl, int kind, uint32_t it, int last_in_iter, int identity_res, int force,
                         double *norm_log, cudaStream_t stream);
// conjugate gradient vector passes (kernels_vec.cu). phase 0: x = 0, r = p = b, rsold = b.b ; phase 1: x += alpha p,
// r -= alpha ap, rsnew = r.r (then beta, rsold, iteration count and the loop decision in the tail) ; phase 2: p = r + beta p
struct CgVecArgs {
    const double *b;   // phase 0
    double *x, *r, *p;
    const double *ap;  // phase 1
    uint64_t n;
    LoopCtl *ctl;
    double *partials;
    int phase;
};
int32_t launch_cg_vec(const CgVecArgs &a, cudaStream_t stream);
int cg_vec_grid();

// P2P exchange: copy this rank's slice src[0..n) to dst[p] + offset on every rank, then signal (kind 0: no sums)
int32_t launch_peer_publish(const double *src, uint64_t n, uint64_t offset, double *const *dst, LoopCtl *ctl,
                            const PeerExchange &px, int force, cudaStream_t stream);

}  // namespace sb200
