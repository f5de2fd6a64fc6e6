/*
 * sublinear_b200.h — C ABI of the B200-native Neumann-series / push-iteration path.
 *
 * This is the drop-in boundary for ONE hot path of ruvnet/sublinear-time-solver: what a Rust
 * `extern "C"` block (or cgo / ctypes / N-API) binds in place of the crate's CPU implementation of
 *   SparseMatrix::from_triplets / Matrix::multiply_vector        (src/matrix/mod.rs:160-199, 415-439)
 *   NeumannSolver::{new,default,fast,high_precision,solve}        (src/solver/neumann.rs:36-80, 469-555)
 *   SolverOptions / SolverResult / SolverError                    (src/solver/mod.rs:22-195, src/error.rs:16-138)
 * plus the entry-estimation and PageRank front doors that only exist in the TS package
 *   SublinearSolver.estimateEntry / computePageRank               (src/core/solver.ts:550-659, 664-722).
 * The reference has no C FFI of its own; its only FFI is wasm-bindgen (src/wasm_iface.rs:45-243), whose
 * handle lifecycle (new ... dispose) and flat-slice arguments this header follows.
 *
 * Conventions
 *  - plain pointers + sizes, no C++/torch types; every function returns int32_t: 0 or an SB200_ERR_* code
 *    equal to the 1-based position of the variant in `enum SolverError` (src/error.rs:16-138);
 *    sb200_last_error() returns the message of the calling thread's last failure.
 *  - all pointers are caller-owned HOST pointers unless the name ends in `_dev` (device pointers on the
 *    handle's GPU) ; sb200_result.solution is library-owned until sb200_result_free().
 *  - the matrix handle owns a device-resident CSR copy (values f64 / col_indices u32 / row_ptr u32 —
 *    CSRStorage, src/matrix/sparse.rs:16-23); solve calls are re-entrant on one handle, like
 *    `solve(&self, &dyn Matrix, ..)` in the reference (workspaces are per call).
 *  - there is no CPU fallback: every compute entry point fails with SB200_ERR_ALGORITHM if no CUDA device
 *    is usable.
 */
#ifndef SUBLINEAR_B200_H
#define SUBLINEAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 1

/* SolverError variants (src/error.rs:16-138), 1-based in declaration order. */
typedef enum sb200_status {
    SB200_OK = 0,
    SB200_ERR_MATRIX_NOT_DIAGONALLY_DOMINANT = 1,
    SB200_ERR_NUMERICAL_INSTABILITY = 2,
    SB200_ERR_CONVERGENCE_FAILURE = 3,
    SB200_ERR_INVALID_INPUT = 4,
    SB200_ERR_DIMENSION_MISMATCH = 5,
    SB200_ERR_UNSUPPORTED_MATRIX_FORMAT = 6,
    SB200_ERR_MEMORY_ALLOCATION = 7,
    SB200_ERR_INDEX_OUT_OF_BOUNDS = 8,
    SB200_ERR_INVALID_SPARSE_MATRIX = 9,
    SB200_ERR_ALGORITHM = 10,      /* also: CUDA / NCCL failures, missing device */
    SB200_ERR_WASM_BINDING = 11,   /* never produced; kept so codes line up with the enum */
    SB200_ERR_IO = 12,
    SB200_ERR_SERIALIZATION = 13
} sb200_status;

/* ConvergenceMode / NormType (src/types.rs:30-55). The Neumann loop itself only uses the L2 residual
 * (src/solver/neumann.rs:316, 422-430); the fields are carried for API fidelity. */
enum { SB200_CONV_RESIDUAL_NORM = 0, SB200_CONV_RELATIVE_RESIDUAL = 1, SB200_CONV_SOLUTION_CHANGE = 2,
       SB200_CONV_RELATIVE_SOLUTION_CHANGE = 3, SB200_CONV_COMBINED = 4 };
enum { SB200_NORM_L1 = 0, SB200_NORM_L2 = 1, SB200_NORM_LINF = 2, SB200_NORM_WEIGHTED = 3 };

/* Which semantics sb200_solve follows. */
enum {
    SB200_MODE_CORRECT = 0,    /* x = sum_k (-D^-1 R)^k D^-1 b ; residual = ||A x - b||_2 (the documented maths) */
    SB200_MODE_REF_COMPAT = 1  /* literal neumann.rs control flow incl. its quirks: x starts at D^-1 b, residual vs D^-1 b */
};
/* Dominance accepted by the solver: Rust checks rows only (src/matrix/mod.rs:467-485); the TS analyser
 * accepts row OR column dominance (src/core/matrix.ts:343-345), which PageRank systems need. */
enum { SB200_DOMINANCE_ROW = 0, SB200_DOMINANCE_ROW_OR_COL = 1 };
/* Residual cadence: the reference recomputes ||A x - rhs|| every 5th iteration (neumann.rs:489-491);
 * IDENTITY takes ||b - A x_k|| = ||D o t_{k+1}|| from the push kernel for free (SURVEY.md F12;
 * MODE_CORRECT only) and runs one verifying residual SpMV at the end. */
enum { SB200_RESIDUAL_EVERY_5 = 0, SB200_RESIDUAL_IDENTITY = 1 };
/* Duplicate (row,col) policy at ingest: Rust keeps them as separate entries. */
enum { SB200_DUP_KEEP = 0, SB200_DUP_SUM = 1 };

typedef struct sb200_matrix sb200_matrix; /* SparseMatrix in CSR form, device resident */
typedef struct sb200_solver sb200_solver; /* NeumannSolver configuration */
typedef struct sb200_comm sb200_comm;     /* one rank of a row-partitioned multi-GPU job */

/* SolverOptions (src/solver/mod.rs:22-45) as a POD, followed by this library's extensions. */
typedef struct sb200_options {
    double tolerance;              /* 1e-6  */
    uint64_t max_iterations;       /* 1000  */
    int32_t convergence_mode;      /* SB200_CONV_RESIDUAL_NORM */
    int32_t norm_type;             /* SB200_NORM_L2 */
    int32_t collect_stats;         /* 0 */
    uint64_t streaming_interval;   /* 0 (carried, unused by NeumannSolver::solve) */
    const double *initial_guess;   /* NULL = None */
    uint64_t initial_guess_len;
    int32_t compute_error_bounds;  /* 0 */
    double error_bounds_tolerance; /* 1e-8 */
    int32_t enable_profiling;      /* 0 */
    int32_t has_random_seed;       /* 0 = None */
    uint64_t random_seed;
    /* extensions */
    int32_t mode;                  /* SB200_MODE_CORRECT */
    int32_t dominance;             /* SB200_DOMINANCE_ROW */
    int32_t residual_check;        /* SB200_RESIDUAL_EVERY_5 */
    int32_t reserved;
} sb200_options;

/* SolverResult (src/solver/mod.rs:121-138) + SolverStats / ErrorBounds / MemoryInfo fields the path fills. */
typedef struct sb200_result {
    double *solution;              /* library-owned, solution_len doubles (NULL for *_dev solves) */
    uint64_t solution_len;
    double residual_norm;
    uint64_t iterations;
    int32_t converged;
    int32_t has_error_bounds;      /* ErrorBounds::upper_bound_only (neumann.rs:340-343) */
    double error_upper_bound;
    int32_t has_stats;             /* SolverStats when options.collect_stats (neumann.rs:539-548) */
    double total_time_ms;
    uint64_t matvec_count;
    uint64_t memory_bytes;         /* MemoryInfo.current_usage_bytes: device bytes of the per-solve workspace */
    /* extensions */
    uint64_t terms_computed;
    int32_t series_converged;
    double last_term_norm;
    double device_time_ms;         /* CUDA-event time of the iteration loop (kernels only) */
    uint64_t kernel_launches;      /* launches of this library's kernels during the call */
    uint64_t h2d_bytes, d2h_bytes; /* bytes moved across PCIe during the call */
    /* filled when options.enable_profiling (ProfileData in the reference, src/types.rs:236-251, is never filled):
     * CUDA-event time of the launches that did work, per kernel kind */
    double push_kernel_ms;         /* fused push kernels (one per term after the first) */
    uint64_t push_kernel_count;
    double resid_kernel_ms;        /* residual kernels (every 5th iteration + final) */
    uint64_t resid_kernel_count;
} sb200_result;

/* ---------------------------------------------------------------------------------------------- */
/* runtime                                                                                        */
/* ---------------------------------------------------------------------------------------------- */
int32_t sb200_abi_version(void);
size_t sb200_last_error(char *buf, size_t cap);
int32_t sb200_device_count(int32_t *count);
/* Device used by handles created afterwards on the calling thread (default 0 or $SUBLINEAR_B200_DEVICE). */
int32_t sb200_set_device(int32_t device);
int32_t sb200_get_device(int32_t *device);
/* Pinned host memory for callers that want PCIe line rate on sb200_solve's b / solution buffers.
 * Pageable pointers are accepted everywhere and staged through an internal pinned ring. */
int32_t sb200_host_alloc(uint64_t bytes, void **out);
int32_t sb200_host_free(void *ptr);

/* ---------------------------------------------------------------------------------------------- */
/* SparseMatrix                                                                                   */
/* ---------------------------------------------------------------------------------------------- */
/* SparseMatrix::from_triplets (src/matrix/mod.rs:160-199): bounds + finiteness validation in triplet
 * order, exact zeros dropped (sparse.rs:535-541), stable sort by (row,col), duplicates kept (sparse.rs:95). */
int32_t sb200_matrix_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals,
                                   uint64_t ntriplets, uint64_t nrows, uint64_t ncols, sb200_matrix **out);
int32_t sb200_matrix_from_triplets_ex(const uint64_t *rows, const uint64_t *cols, const double *vals,
                                      uint64_t ntriplets, uint64_t nrows, uint64_t ncols, int32_t dup_policy,
                                      sb200_matrix **out);
/* Ingest the three CSRStorage slices as they are (src/matrix/sparse.rs:16-23; the 5-slice kernel signature of
 * src/simd_ops.rs:20-26). Validates monotone row_ptr, column bounds and finite values; rows need not be sorted. */
int32_t sb200_matrix_from_csr(const uint32_t *row_ptr, const uint32_t *col_indices, const double *values,
                              uint64_t nrows, uint64_t ncols, uint64_t nnz, sb200_matrix **out);
int32_t sb200_matrix_from_csr64(const uint64_t *row_ptr, const uint32_t *col_indices, const double *values,
                                uint64_t nrows, uint64_t ncols, uint64_t nnz, sb200_matrix **out);
/* SparseMatrix::from_dense / identity / diagonal (src/matrix/mod.rs:202-239). */
int32_t sb200_matrix_from_dense(const double *data, uint64_t nrows, uint64_t ncols, sb200_matrix **out);
int32_t sb200_matrix_identity(uint64_t size, sb200_matrix **out);
int32_t sb200_matrix_diagonal(const double *diag, uint64_t size, sb200_matrix **out);
void sb200_matrix_free(sb200_matrix *m);

/* Matrix::rows / cols / nnz / get / is_diagonally_dominant / diagonal_dominance_factor (src/matrix/mod.rs:25-104). */
int32_t sb200_matrix_rows(const sb200_matrix *m, uint64_t *out);
int32_t sb200_matrix_cols(const sb200_matrix *m, uint64_t *out);
int32_t sb200_matrix_nnz(const sb200_matrix *m, uint64_t *out);
int32_t sb200_matrix_get(const sb200_matrix *m, uint64_t row, uint64_t col, double *value, int32_t *present);
int32_t sb200_matrix_is_diagonally_dominant(const sb200_matrix *m, int32_t dominance, int32_t *out);
int32_t sb200_matrix_diagonal_dominance_factor(const sb200_matrix *m, double *factor, int32_t *present);
/* Device layout the hot kernels read (extension; the CSRStorage slices stay the ingest / export format):
 * layout 0 = the CSR slices as uploaded, 1 = an additional SELL-32 copy (blocks of 32 rows, element k of row r at
 * slab*32 + k*32 + r, zero-padded to the longest row of the block) chosen when it costs <= 25 % extra slots,
 * 2 = an additional copy regrouped into 2..4 column slabs (one kernel pass per slab; chosen when the gathered vector
 * does not fit the L2 partition of a die, i.e. 8*ncols > 48 MB). Results are bit-identical in all three.
 * slots = value slots the kernels stream per SpMV (= nnz for layouts 0 and 2); device_bytes = all matrix arrays. */
enum { SB200_LAYOUT_CSR = 0, SB200_LAYOUT_SELL32 = 1, SB200_LAYOUT_CSR_SLABS = 2 };
int32_t sb200_matrix_storage_info(const sb200_matrix *m, int32_t *layout, uint64_t *slots, uint64_t *device_bytes);
/* CSRStorage::to_triplets / SparseMatrix::as_csr (src/matrix/sparse.rs:210-227, mod.rs:313-319): copy out.
 * Any output pointer may be NULL. row_ptr has nrows+1 entries. */
int32_t sb200_matrix_export_csr(const sb200_matrix *m, uint64_t *row_ptr, uint32_t *col_indices, double *values);
/* Matrix::multiply_vector / multiply_vector_add (src/matrix/mod.rs:415-465): y = A x, y += A x. */
int32_t sb200_matrix_multiply_vector(const sb200_matrix *m, const double *x, uint64_t xlen, double *y, uint64_t ylen);
int32_t sb200_matrix_multiply_vector_add(const sb200_matrix *m, const double *x, uint64_t xlen, double *y, uint64_t ylen);
/* Same on device pointers, enqueued on `stream` (a cudaStream_t, NULL = default stream); no host sync. */
int32_t sb200_matrix_multiply_vector_dev(const sb200_matrix *m, const double *x_dev, uint64_t xlen, double *y_dev,
                                         uint64_t ylen, int32_t accumulate, void *stream);
/* SparseMatrix::scale / add_diagonal (src/matrix/mod.rs:345-372). */
int32_t sb200_matrix_scale(sb200_matrix *m, double factor);

/* ---------------------------------------------------------------------------------------------- */
/* NeumannSolver + SolverOptions                                                                  */
/* ---------------------------------------------------------------------------------------------- */
/* NeumannSolver::new(max_terms, series_tolerance) / default() / high_precision() / fast() (neumann.rs:48-80). */
int32_t sb200_neumann_new(uint64_t max_terms, double series_tolerance, sb200_solver **out);
int32_t sb200_neumann_default(sb200_solver **out);
int32_t sb200_neumann_high_precision(sb200_solver **out);
int32_t sb200_neumann_fast(sb200_solver **out);
/* with_adaptive_truncation / with_power_caching (neumann.rs:83-92). */
int32_t sb200_neumann_with_adaptive_truncation(sb200_solver *s, int32_t enable);
int32_t sb200_neumann_with_power_caching(sb200_solver *s, int32_t enable);
int32_t sb200_neumann_config(const sb200_solver *s, uint64_t *max_terms, double *series_tolerance,
                             int32_t *adaptive_truncation, int32_t *cache_powers);
/* SolverAlgorithm::algorithm_name (neumann.rs:464-466): "neumann". */
const char *sb200_solver_algorithm_name(const sb200_solver *s);
void sb200_solver_free(sb200_solver *s);

/* SolverOptions::default / high_precision / fast / streaming(interval) (src/solver/mod.rs:47-116). */
void sb200_options_default(sb200_options *o);
void sb200_options_high_precision(sb200_options *o);
void sb200_options_fast(sb200_options *o);
void sb200_options_streaming(sb200_options *o, uint64_t interval);

/* NeumannSolver::solve(&matrix, &b, &options) -> Result<SolverResult> (src/solver/neumann.rs:469-555).
 * On SB200_ERR_CONVERGENCE_FAILURE / SB200_ERR_NUMERICAL_INSTABILITY `out` is still filled (iterations,
 * residual_norm = the fields of the Err variant, plus the last iterate). */
int32_t sb200_solve(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                    const sb200_options *opt, sb200_result *out);
/* Same, writing the solution into a caller buffer of blen doubles (out->solution stays NULL). */
int32_t sb200_solve_into(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                         const sb200_options *opt, double *x_out, sb200_result *out);
/* Inputs/outputs already resident in HBM: b_dev, x_dev (and opt->initial_guess, if set) are device
 * pointers on the matrix's GPU; work is enqueued on `stream`; the call returns after the loop finished. */
int32_t sb200_solve_dev(const sb200_solver *s, const sb200_matrix *m, const double *b_dev, uint64_t blen,
                        const sb200_options *opt, double *x_dev, void *stream, sb200_result *out);
void sb200_result_free(sb200_result *r);

/* The bare recurrence, for measurement and per-term parity: from t = D^-1 b, x = t run exactly `nterms`
 * push iterations t <- t - D^-1 (A t); x += t (neumann.rs:280-299 + 264-266 + 271) with no convergence logic.
 * Device pointers; x_dev / t_dev / term_norms (host, nterms doubles) may be NULL. elapsed_ms = CUDA-event
 * time of the nterms kernel launches alone. */
int32_t sb200_push_iterations_dev(const sb200_matrix *m, const double *b_dev, uint64_t blen, uint64_t nterms,
                                  double *x_dev, double *t_dev, double *term_norms, void *stream,
                                  float *elapsed_ms);

/* ---------------------------------------------------------------------------------------------- */
/* SolverAlgorithm state interface and streaming                                */
/* ---------------------------------------------------------------------------------------------- */
/* trait SolverAlgorithm { initialize, step, is_converged, extract_solution, update_rhs } (src/solver/mod.rs:223-252)
 * and trait SolverState { residual_norm, matvec_count, error_bounds, memory_usage, reset } (:336-352) for
 * NeumannSolver / NeumannState (src/solver/neumann.rs:97-135, 350-462). The state shares ownership of its matrix
 * handle: sb200_matrix_free and sb200_state_free may be called in either order. */
typedef struct sb200_state sb200_state;
enum { SB200_STEP_CONTINUE = 0, SB200_STEP_CONVERGED = 1 }; /* StepResult (src/solver/mod.rs:355-363) */
typedef struct sb200_state_info_t {
    uint64_t dimension;
    double residual_norm;      /* +inf until the first step */
    uint64_t matvec_count;
    uint64_t terms_computed;
    int32_t series_converged;
    double last_term_norm;
    int32_t has_error_bounds;  /* estimate_error_bounds (neumann.rs:321-347), when adaptive_truncation */
    double error_upper_bound;
    uint64_t memory_bytes;
} sb200_state_info_t;
/* initialize = NeumannState::new (neumann.rs:139-249): same checks and errors as sb200_solve; no term is added. */
int32_t sb200_neumann_initialize(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                                 const sb200_options *opt, sb200_state **out);
/* step: the reference's own step() fails for lack of a matrix reference (neumann.rs:390-403); this runs the body it
 * left commented out (:404-418): compute_next_term, update_residual, estimate_error_bounds;
 * *step_result = CONVERGED once series_converged or terms_computed >= max_terms.
 * SB200_ERR_NUMERICAL_INSTABILITY on a non-finite residual (solver/mod.rs:271-279). */
int32_t sb200_state_step(sb200_state *st, int32_t *step_result);
int32_t sb200_state_is_converged(const sb200_state *st, int32_t *out);             /* neumann.rs:422-430 */
int32_t sb200_state_extract_solution(const sb200_state *st, double *x, uint64_t xlen); /* :432-434 */
/* update_rhs(&mut state, &[(index, delta)]) (neumann.rs:436-462). SB200_MODE_REF_COMPAT: the literal code (the scaled
 * delta goes into rhs AND the solution, the series restarts from the whole new rhs). SB200_MODE_CORRECT: the
 * incremental solve its comment asks for — the series restarts from D^-1 delta_b, so further steps add
 * A^-1 delta_b to the solution already held. SB200_ERR_INDEX_OUT_OF_BOUNDS leaves the state untouched. */
int32_t sb200_state_update_rhs(sb200_state *st, const uint64_t *indices, const double *deltas, uint64_t count);
int32_t sb200_state_reset(sb200_state *st);                                         /* neumann.rs:367-378 */
int32_t sb200_state_info(const sb200_state *st, sb200_state_info_t *info);
void sb200_state_free(sb200_state *st);

/* PartialSolution (src/solver/mod.rs:198-217), handed to the callback every options.streaming_interval iterations of
 * sb200_solve_streaming; `solution` points to library-owned pinned host memory valid during the call only. */
typedef struct sb200_partial_solution {
    uint64_t iteration;
    const double *solution;
    uint64_t solution_len;
    double residual_norm;          /* latest evaluated residual (+inf before the first evaluation) */
    int32_t converged;
    int32_t has_estimated_remaining;
    uint64_t estimated_remaining;  /* terms until ||t|| < series_tolerance at the observed contraction rate */
    double timestamp_ms;           /* since the call started */
} sb200_partial_solution;
/* return non-zero to stop the solve after this partial solution */
typedef int32_t (*sb200_stream_callback)(const sb200_partial_solution *partial, void *user);
/* NeumannSolver::solve with SolverOptions::streaming(interval) (src/solver/mod.rs:100-116); the callback shape follows
 * WasmSublinearSolver::solve_stream (src/wasm_iface.rs:119-166). streaming_interval = 0 -> plain sb200_solve. */
int32_t sb200_solve_streaming(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                              const sb200_options *opt, sb200_stream_callback callback, void *user, sb200_result *out);

/* ---------------------------------------------------------------------------------------------- */
/* conjugate gradient on the same SpMV kernel                          */
/* ---------------------------------------------------------------------------------------------- */
/* OptimizedSolverConfig (src/optimized_solver.rs:108-127). */
typedef struct sb200_cg_config {
    uint64_t max_iterations;   /* 1000 */
    double tolerance;          /* 1e-6 */
    int32_t enable_profiling;  /* 0; here: CUDA-event time of every SpMV launch */
    int32_t reserved;
} sb200_cg_config;

/* OptimizedSolverResult + OptimizedSolverStats (src/optimized_solver.rs:130-166). */
typedef struct sb200_cg_result {
    double *solution;              /* library-owned until sb200_cg_result_free (NULL for *_into / *_dev solves) */
    uint64_t solution_len;
    double residual_norm;          /* sqrt(r.r) of the recurrence at exit (:275) */
    uint64_t iterations;
    int32_t converged;             /* `rsold <= tolerance^2` seen while iteration < max_iterations (:217-221) */
    int32_t breakdown;             /* extension: the loop left through |p.Ap| < 1e-16 (:234-236) */
    double computation_time_ms;
    uint64_t matvec_count;
    uint64_t dot_product_count;
    uint64_t axpy_count;
    uint64_t total_flops;          /* matvec_count*nnz*2 + iterations*rows*6 (:278-279) */
    double average_bandwidth_gbs;  /* the reference's own formulas (:281-285) */
    double average_gflops;
    /* extensions */
    double device_time_ms;         /* CUDA-event time of the whole loop */
    uint64_t kernel_launches;
    uint64_t h2d_bytes, d2h_bytes;
    double spmv_kernel_ms;         /* enable_profiling: CUDA-event time of the SpMV launches that did work */
    uint64_t spmv_kernel_count;
} sb200_cg_result;

void sb200_cg_config_default(sb200_cg_config *c);
/* OptimizedConjugateGradientSolver::solve(&matrix, &b) (src/optimized_solver.rs:182-295); the same loop as
 * FastConjugateGradient::solve (src/fast_solver.rs:126-178) and UltraFastCG::solve (src/ultra_fast.rs:116-158).
 * x0 = 0; no dominance or symmetry check (the reference has none). "Matrix must be square" ->
 * SB200_ERR_INVALID_INPUT, length mismatch -> SB200_ERR_DIMENSION_MISMATCH. Not converging is not an error
 * (converged = 0), as in the reference. */
int32_t sb200_cg_solve(const sb200_matrix *m, const double *b, uint64_t blen, const sb200_cg_config *cfg,
                       sb200_cg_result *out);
int32_t sb200_cg_solve_into(const sb200_matrix *m, const double *b, uint64_t blen, const sb200_cg_config *cfg,
                            double *x_out, sb200_cg_result *out);
/* b_dev / x_dev resident in HBM on the matrix's GPU; work is enqueued on `stream`, the call returns after the loop. */
int32_t sb200_cg_solve_dev(const sb200_matrix *m, const double *b_dev, uint64_t blen, const sb200_cg_config *cfg,
                           double *x_dev, void *stream, sb200_cg_result *out);
void sb200_cg_result_free(sb200_cg_result *r);

/* ---------------------------------------------------------------------------------------------- */
/* forward / backward push                                                      */
/* ---------------------------------------------------------------------------------------------- */
/* PushGraph (src/graph/adjacency.rs:199-277): adjacency + transpose + out-/in-degrees (row / column weight sums). */
typedef struct sb200_push_graph sb200_push_graph;
/* ForwardPushConfig = BackwardPushConfig (src/solver/forward_push.rs:25-50, backward_push.rs:25-50). */
typedef struct sb200_push_config {
    double alpha;               /* 0.15 restart probability */
    double epsilon;             /* 1e-6: a node is pushed while residual >= epsilon * max(degree, 1) */
    uint64_t max_pushes;        /* 1 000 000; never exceeded (the round that reaches it pushes the lowest node ids only) */
    double queue_threshold;     /* 1e-8: admission test residual / max(degree, 1) >= threshold */
    int32_t adaptive_threshold; /* 1: threshold x1.1 / x0.9 every 1 000 pushes by queue length (src/graph/mod.rs:204-212) */
    int32_t reserved;
} sb200_push_config;
/* ForwardPushResult / BackwardPushResult (forward_push.rs:10-22) minus the two vectors (caller buffers). */
typedef struct sb200_push_stats {
    uint64_t push_count;
    uint64_t nodes_visited;
    double residual_norm;       /* L2 norm of the residual vector (forward_push.rs:217-219) */
    /* extensions */
    uint64_t rounds;            /* rounds: every queued node above its threshold is pushed in the same round */
    uint64_t kernel_launches;
    double device_time_ms;
    uint64_t dense_rounds;      /* rounds run as select + SpMV over the whole graph (frontier above nnz / 4 edges) */
    uint64_t edges_touched;     /* edges read by the pushes: the work of the walk (nnz per dense round) */
} sb200_push_stats;
void sb200_push_config_default(sb200_push_config *c);
/* PushGraph::from_matrix(&CompressedSparseRow) (adjacency.rs:211-224): row u = out-edges of u, weights >= 0. */
int32_t sb200_push_graph_from_csr(const uint64_t *row_ptr, const uint32_t *col_indices, const double *weights, uint64_t n,
                                  sb200_push_graph **out);
/* PushGraph::from_edges(num_nodes, &[(from, to, weight)]) (adjacency.rs:227-239): out-of-range edges are dropped. */
int32_t sb200_push_graph_from_edges(uint64_t n, const uint64_t *from, const uint64_t *to, const double *weights,
                                    uint64_t nedges, sb200_push_graph **out);
int32_t sb200_push_graph_info(const sb200_push_graph *g, uint64_t *num_nodes, uint64_t *num_edges);
int32_t sb200_push_graph_degrees(const sb200_push_graph *g, uint64_t node, double *out_degree, double *in_degree);
void sb200_push_graph_free(sb200_push_graph *g);
/* ForwardPushSolver::solve_single_source (nsources = 1: unit mass; an out-of-range source gives the all-zero result) /
 * solve_multi_source (mass 1/nsources per listed source) (forward_push.rs:66-177). estimate / residual: n doubles each.
 * The reference pushes one node at a time in priority order; here every queued node above its threshold is pushed in
 * the same round and only the frontier's edges are read (sparse frontier: candidate list -> flag / scan -> expand ->
 * stable sort by neighbour -> ordered reduce; deterministic, no floating-point atomics). Same push rule, same stopping
 * condition, same invariants: estimate, residual >= 0, sum(estimate) + sum(residual) = 1 for the forward direction,
 * residual < epsilon * max(degree, 1) everywhere at exit. */
int32_t sb200_forward_push(const sb200_push_graph *g, const sb200_push_config *cfg, const uint64_t *sources,
                           uint64_t nsources, double *estimate, double *residual, sb200_push_stats *stats);
/* BackwardPushSolver::solve_single_target / solve_multi_target (backward_push.rs:66-220): mass moves to the
 * predecessors with weight / max(out_degree(predecessor), 1); thresholds use the in-degree. */
int32_t sb200_backward_push(const sb200_push_graph *g, const sb200_push_config *cfg, const uint64_t *targets,
                            uint64_t ntargets, double *estimate, double *residual, sb200_push_stats *stats);
/* ForwardPushSolver::solve_with_target (forward_push.rs:234-290): stops once estimate[target] > target_precision and
 * residual[target] < 0.1 target_precision (tested once per round here, before every pop in the reference). */
int32_t sb200_forward_push_with_target(const sb200_push_graph *g, const sb200_push_config *cfg, uint64_t source,
                                       uint64_t target, double target_precision, double *estimate, double *residual,
                                       sb200_push_stats *stats);
/* BackwardPushSolver::solve_with_source (backward_push.rs:238-290). */
int32_t sb200_backward_push_with_source(const sb200_push_graph *g, const sb200_push_config *cfg, uint64_t source,
                                        uint64_t target, double source_precision, double *estimate, double *residual,
                                        sb200_push_stats *stats);
/* BackwardPushSolver::combine_with_forward (backward_push.rs:312-330). */
int32_t sb200_push_combine_with_forward(double alpha, const double *backward_estimate, const double *backward_residual,
                                        uint64_t nbackward, const double *forward_estimate, const double *forward_residual,
                                        uint64_t nforward, double *out);
/* BidirectionalPushSolver::solve_bidirectional / adaptive_solve (backward_push.rs:362-420). */
int32_t sb200_bidirectional_push(const sb200_push_graph *g, const sb200_push_config *forward_cfg,
                                 const sb200_push_config *backward_cfg, uint64_t source, uint64_t target, double *out);
int32_t sb200_bidirectional_adaptive_push(const sb200_push_graph *g, const sb200_push_config *forward_cfg,
                                          const sb200_push_config *backward_cfg, uint64_t source, uint64_t target, double *out);

/* SublinearSolver.solveForwardPush (src/core/solver.ts:437-522) for A x = b: the residual r = b - A x is kept exact, a push
 * of node i sets x_i += r_i / a_ii, r_i = 0, r_j -= a_ji r_i / a_ii down column i. The reference pushes the node of largest
 * |r| per iteration; here every queued node with |r_i| >= epsilon is pushed per round (sparse frontier, as above).
 * Converged: max |r| < epsilon. `iterations` counts node pushes; reaching max_iterations without convergence returns
 * SB200_ERR_CONVERGENCE_FAILURE, a zero diagonal under a pushed node SB200_ERR_NUMERICAL_INSTABILITY (x_out and stats
 * are filled in both cases). */
typedef struct sb200_axb_push_stats {
    uint64_t iterations;
    uint64_t rounds;
    double residual_norm;       /* ||b - A x||_2 (VectorOperations.norm2(residual)) */
    double max_residual;        /* max |r_i| */
    int32_t converged;
    int32_t reserved;
} sb200_axb_push_stats;
int32_t sb200_forward_push_solve(const sb200_matrix *m, const double *b, uint64_t blen, double epsilon,
                                 uint64_t max_iterations, double *x_out, sb200_axb_push_stats *stats);

/* ---------------------------------------------------------------------------------------------- */
/* StreamingMatrix (src/matrix/optimized.rs:451-561)                             */
/* ---------------------------------------------------------------------------------------------- */
/* Row chunks sized from a memory limit (chunk_size = min(rows, max(1, limit / (2 (12 nt / rows + 4))))), kept in pinned
 * host memory; multiply_vector_streaming uploads chunk k + 1 while chunk k is multiplied on the GPU and hands every
 * chunk's slice of y to the callback in row order. Triplet semantics as SparseMatrix::from_triplets. */
typedef struct sb200_streaming_matrix sb200_streaming_matrix;
/* start_row, the chunk's slice of A x, its length; a non-zero return stops the walk (extension). */
typedef int32_t (*sb200_chunk_callback)(uint64_t start_row, const double *result, uint64_t len, void *user);
int32_t sb200_streaming_matrix_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals, uint64_t ntriplets,
                                             uint64_t nrows, uint64_t ncols, uint64_t memory_limit_mb,
                                             sb200_streaming_matrix **out);
int32_t sb200_streaming_matrix_info(const sb200_streaming_matrix *sm, uint64_t *total_rows, uint64_t *total_cols,
                                    uint64_t *chunk_size, uint64_t *num_chunks, uint64_t *memory_usage);
int32_t sb200_streaming_matrix_multiply_vector(const sb200_streaming_matrix *sm, const double *x, uint64_t xlen,
                                               sb200_chunk_callback callback, void *user);
void sb200_streaming_matrix_free(sb200_streaming_matrix *sm);

/* ---------------------------------------------------------------------------------------------- */
/* single-entry estimation and PageRank (TS-only front doors)               */
/* ---------------------------------------------------------------------------------------------- */
/* Batched estimate of x[rows[q]] for A x = b by absorbing random walks (Ulam-von Neumann estimator,
 * SURVEY.md Appendix C).  nwalks = 0 -> max(100, ceil(1/eps^2)) (src/core/solver.ts:587);
 * max_steps = 0 -> 1000 (src/core/solver.ts:399).  est / var: nqueries doubles each (var may be NULL). */
int32_t sb200_solve_entry(const sb200_matrix *m, const double *b, uint64_t blen, const uint64_t *rows,
                          uint64_t nqueries, double eps, uint64_t nwalks, uint64_t max_steps, uint64_t seed,
                          double *est, double *var);
/* The same batch over several GPUs (replicas): replicas[r] holds the whole matrix on its own device
 * (sb200_set_device + sb200_matrix_from_* once per GPU); the queries are cut into contiguous slices, one host thread per
 * replica. Estimates are identical to one sb200_solve_entry call for any number of replicas (the walk keys use the
 * position in the whole batch). */
int32_t sb200_solve_entry_replicas(const sb200_matrix *const *replicas, int32_t nreplicas, const double *b, uint64_t blen,
                                   const uint64_t *rows, uint64_t nqueries, double eps, uint64_t nwalks, uint64_t max_steps,
                                   uint64_t seed, double *est, double *var);
/* computePageRank's system (src/core/solver.ts:664-722): S = I - alpha P^T with dangling mass dropped,
 * rhs = (1-alpha)/n.  Edge e: src[e] -> dst[e] with weight w[e] (w NULL = 1). rhs: n doubles (may be NULL). */
int32_t sb200_pagerank_system(const uint64_t *src, const uint64_t *dst, const double *w, uint64_t nedges,
                              uint64_t n, double alpha, sb200_matrix **S, double *rhs);

/* ---------------------------------------------------------------------------------------------- */
/* synthetic inputs (the reference's own generators)                               */
/* ---------------------------------------------------------------------------------------------- */
/* create_test_matrix + create_test_rhs (benches/performance_benchmarks.rs:12-43), rows [row0,row1) of the
 * size x size system, emitted in CSR (= from_triplets of the generated triplets; row_ptr local, 64-bit).
 * Pass NULL outputs to query sizes: *nnz_out receives the nnz of the row range. */
int32_t sb200_gen_bench_csr(uint64_t size, double sparsity, uint64_t row0, uint64_t row1, uint64_t *row_ptr,
                            uint32_t *col_indices, double *values, double *b, uint64_t *nnz_out);

/* ---------------------------------------------------------------------------------------------- */
/* multi-GPU: contiguous row blocks, one process (rank) per GPU, one exchange per iteration       */
/* ---------------------------------------------------------------------------------------------- */
#define SB200_UNIQUE_ID_BYTES 128
/* Rank 0 creates the id and ships it to the other ranks by any host-side means (torch.distributed
 * broadcast, MPI, a file); every rank then calls sb200_comm_init on its own GPU. */
int32_t sb200_comm_unique_id(uint8_t id[SB200_UNIQUE_ID_BYTES]);
int32_t sb200_comm_init(int32_t rank, int32_t world, const uint8_t id[SB200_UNIQUE_ID_BYTES], int32_t device,
                        sb200_comm **out);
void sb200_comm_free(sb200_comm *c);
/* Row partition used by every rank: rank r owns rows [r*ceil(n/world), min(n,(r+1)*ceil(n/world))) — the
 * reference's own chunking rule (src/simd_ops.rs:219, src/matrix/optimized.rs:485-517). */
int32_t sb200_partition_rows(uint64_t nrows, int32_t world, int32_t rank, uint64_t *row0, uint64_t *row1);
/* The rank's row block [row0,row1) of an n x n system: local row_ptr (64-bit), GLOBAL column indices. */
int32_t sb200_dist_matrix_from_csr(sb200_comm *c, uint64_t n_global, uint64_t row0, uint64_t row1,
                                   const uint64_t *row_ptr, const uint32_t *col_indices, const double *values,
                                   sb200_matrix **out);
/* Distributed NeumannSolver::solve: b_local / x_local hold the rank's rows only (host pointers). Per term:
 * fused push kernel on the local rows, then one allgather of the updated term slice + a 1-double allreduce. */
int32_t sb200_dist_solve(sb200_comm *c, const sb200_solver *s, const sb200_matrix *m_local, const double *b_local,
                         uint64_t nlocal, const sb200_options *opt, double *x_local, sb200_result *out);
/* Bare distributed recurrence for measurement (cf. sb200_push_iterations_dev); device pointers, local rows. */
int32_t sb200_dist_push_iterations_dev(sb200_comm *c, const sb200_matrix *m_local, const double *b_local_dev,
                                       uint64_t nlocal, uint64_t nterms, double *x_local_dev, double *term_norms,
                                       float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* SUBLINEAR_B200_H */
