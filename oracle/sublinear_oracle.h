/*
 * oracle/sublinear_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's Rust CPU path for the Neumann-series / push
 * iteration (ruvnet/sublinear-time-solver, commit 6e0dd66).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (sublinear-time-solver_b200/) never links, imports or calls it.
 *
 * Why a restatement: the reference is Rust (+TypeScript); this image has neither cargo/rustc nor
 * node, and the crate's own build system is out of bounds, so oracle/_ref (the real reference
 * compiled) cannot exist here.  Parity pinning: the oracle is checked against every known-answer
 * test the reference holds for this path (tests/test_oracle_golden.py) and against
 * golden vectors produced by the one reference-authored implementation that does run here, the
 * numpy Jacobi in scripts/linear_systems/iterative_solvers.py:17-105 (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows.  Arithmetic is IEEE f64 with no FMA
 * contraction (compile with -ffp-contract=off: rustc never fuses a*b+c on its own).
 */
#ifndef SUBLINEAR_ORACLE_H
#define SUBLINEAR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes = 1-based position of the variant in `enum SolverError` (src/error.rs:16-138). */
enum {
    ORC_OK = 0,
    ORC_ERR_NOT_DIAGONALLY_DOMINANT = 1,
    ORC_ERR_NUMERICAL_INSTABILITY = 2,
    ORC_ERR_CONVERGENCE_FAILURE = 3,
    ORC_ERR_INVALID_INPUT = 4,
    ORC_ERR_DIMENSION_MISMATCH = 5,
    ORC_ERR_UNSUPPORTED_FORMAT = 6,
    ORC_ERR_MEMORY_ALLOCATION = 7,
    ORC_ERR_INDEX_OUT_OF_BOUNDS = 8,
    ORC_ERR_INVALID_SPARSE_MATRIX = 9,
    ORC_ERR_ALGORITHM = 10
};

/* CSRStorage (src/matrix/sparse.rs:16-23): values f64, col_indices u32, row_ptr u32. */
typedef struct {
    uint64_t nrows, ncols, nnz;
    double *values;
    uint32_t *col_indices;
    uint32_t *row_ptr; /* nrows + 1 */
} orc_csr;

void orc_csr_free(orc_csr *m);

/* SparseMatrix::from_triplets (src/matrix/mod.rs:160-199) -> COOStorage::from_triplets
 * (src/matrix/sparse.rs:528-548) -> CSRStorage::from_coo (src/matrix/sparse.rs:80-132). */
int orc_csr_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals,
                          uint64_t ntrip, uint64_t nrows, uint64_t ncols, orc_csr *out);

/* CSRStorage::get (src/matrix/sparse.rs:142-155): binary search in the row; returns 1 if present. */
int orc_csr_get(const orc_csr *m, uint64_t row, uint64_t col, double *out);

/* SpMV variants.  All compute y = A*x. */
enum { ORC_SPMV_SCALAR = 0, ORC_SPMV_SIMD4 = 1, ORC_SPMV_PARALLEL = 2 };
/* CSRStorage::multiply_vector (src/matrix/sparse.rs:187-203). */
void orc_spmv_scalar(const orc_csr *m, const double *x, double *y);
/* matrix_vector_multiply_simd, `simd` feature on (src/simd_ops.rs:20-88). */
void orc_spmv_simd4(const orc_csr *m, const double *x, double *y);
/* parallel_matrix_vector_multiply (src/simd_ops.rs:202-239); nthreads<=0 -> all cores. */
void orc_spmv_parallel(const orc_csr *m, const double *x, double *y, int nthreads);
/* Matrix::multiply_vector (src/matrix/mod.rs:415-439): dimension checks, then dispatch. */
int orc_multiply_vector(const orc_csr *m, const double *x, uint64_t xlen, double *y, uint64_t ylen,
                        int variant, int nthreads);

/* Matrix::is_diagonally_dominant (src/matrix/mod.rs:467-485), row-wise, equality allowed.
 * first_bad_row (optional) receives the first violating row or UINT64_MAX. */
int orc_is_diagonally_dominant(const orc_csr *m, uint64_t *first_bad_row);
/* Column-wise dominance: accepted by the TS analyzeMatrix (src/core/matrix.ts:343-345). */
int orc_is_col_diagonally_dominant(const orc_csr *m);

/* utils::l2_norm / l1_norm / linf_norm (src/solver/mod.rs:369-381). */
double orc_l2_norm(const double *v, uint64_t n);
double orc_l1_norm(const double *v, uint64_t n);
double orc_linf_norm(const double *v, uint64_t n);
/* dot_product_simd / axpy_simd (src/simd_ops.rs:116-147, 158-189). */
double orc_dot_simd4(const double *x, const double *y, uint64_t n);
void orc_axpy_simd4(double alpha, const double *x, double *y, uint64_t n);

/* Solve modes. */
enum {
    ORC_MODE_CORRECT = 0,   /* x = sum_k (-D^-1 R)^k D^-1 b, residual ||Ax-b||_2 (documented maths, neumann.rs:16-22) */
    ORC_MODE_REF_COMPAT = 1 /* literal control flow + quirks of neumann.rs:139-249,469-555 (x starts at c, residual vs c) */
};
enum { ORC_DOM_ROW = 0, ORC_DOM_ROW_OR_COL = 1 };

/* SolverOptions (src/solver/mod.rs:22-63) + NeumannSolver config (src/solver/neumann.rs:24-80). */
typedef struct {
    double tolerance;          /* 1e-6 */
    uint64_t max_iterations;   /* 1000 */
    const double *initial_guess; /* NULL = none */
    uint64_t initial_guess_len;
    int compute_error_bounds;  /* 0 */
    uint64_t max_terms;        /* NeumannSolver::default(): 50 */
    double series_tolerance;   /* 1e-8 */
    int adaptive_truncation;   /* 1 */
    int mode;                  /* ORC_MODE_* */
    int dominance;             /* ORC_DOM_* (Rust: row only) */
    int spmv_variant;          /* ORC_SPMV_* used inside the loop */
    int nthreads;              /* for ORC_SPMV_PARALLEL */
} orc_options;

void orc_options_default(orc_options *o);

/* SolverResult (src/solver/mod.rs:121-138) + SolverStats.matvec_count. */
typedef struct {
    double *solution; /* caller-provided buffer of n doubles */
    double residual_norm;
    uint64_t iterations;
    uint64_t terms_computed;
    uint64_t matvec_count;
    int converged;
    int series_converged;
    int has_error_bound;
    double error_bound;
    double last_term_norm;
    double total_time_ms;
} orc_result;

/* NeumannSolver::solve (src/solver/neumann.rs:469-555) incl. NeumannState::new (:139-249),
 * compute_next_term (:252-277), apply_iteration_matrix (:280-299), update_residual (:302-318),
 * estimate_error_bounds (:321-347), is_converged (:422-430). Returns ORC_OK or an ORC_ERR_*. */
int orc_neumann_solve(const orc_csr *m, const double *b, uint64_t blen, const orc_options *opt,
                      orc_result *res);

/* ---- the SolverAlgorithm state interface: initialize / step / is_converged / extract_solution / update_rhs
 * (src/solver/mod.rs:223-252) and SolverState::reset (neumann.rs:367-378). `m` must outlive the state. ---- */
typedef struct orc_state orc_state;
int orc_state_new(const orc_csr *m, const double *b, uint64_t blen, const orc_options *opt, orc_state **out);
/* step as the reference intends it (the body commented out at neumann.rs:404-418): *step_result 0 = Continue,
 * 1 = Converged. */
int orc_state_step(orc_state *st, int *step_result);
int orc_state_is_converged(const orc_state *st);
void orc_state_solution(const orc_state *st, double *x);
/* update_rhs (neumann.rs:436-462): literal in ORC_MODE_REF_COMPAT; in ORC_MODE_CORRECT the series restarts from
 * D^-1 delta_b so that further steps add A^-1 delta_b to the solution held. */
int orc_state_update_rhs(orc_state *st, const uint64_t *idx, const double *delta, uint64_t count);
void orc_state_reset(orc_state *st);
void orc_state_info(const orc_state *st, double *residual_norm, uint64_t *matvec_count, uint64_t *terms_computed,
                    int *series_converged, double *term_norm, int *has_bound, double *bound);
void orc_state_free(orc_state *st);

/* Run exactly `nterms` push iterations t <- t - dinv.(A t); x += t from t = c, x = c (no control
 * flow) and return wall seconds: used as the timed CPU baseline and for per-term parity. */
double orc_push_iterations(const orc_csr *m, const double *b, uint64_t nterms, int spmv_variant,
                           int nthreads, double *x_out, double *t_out, double *term_norms);

/* ---- forward / backward push: ForwardPushSolver::{solve_single_source, solve_multi_source}
 * (src/solver/forward_push.rs:66-216) and BackwardPushSolver::{solve_single_target, solve_multi_target}
 * (src/solver/backward_push.rs:66-220) over PushGraph::from_matrix (src/graph/adjacency.rs:211-224) with the
 * WorkQueue / VisitedTracker of src/graph/mod.rs:130-260.
 * `adj` is the adjacency CSR (row u = out-edges of u with weights); degrees are row sums, reverse degrees column sums.
 * Pop order: WorkItem only derives PartialOrd (src/graph/mod.rs:141-147) and has no Ord impl, so BinaryHeap<WorkItem>
 * does not compile in the reference; the restatement orders items by (priority, node_id), what the derive would give.
 * Pinning: the reference holds no golden vectors for this path and cannot be built (no Ord impl, no Rust toolchain
 * here), but (priority, node_id) is a TOTAL order, so the pop sequence does not depend on the heap's internals: the
 * restatement is pinned bit for bit to an independent pure-Python restatement (tests/test_push.py::py_forward_push,
 * heapq) on the graphs of tests/rust/push_tests.rs:15-59, with the vectors committed (tests/golden/push_*.npz,
 * make_golden_push.py), and held to the reference's own tests restated (reference_push_tests) and to the exact
 * personalised PageRank. Not pinned to an output of reference code: none exists. est / res: n doubles each. */
typedef struct {
    double alpha;            /* 0.15 */
    double epsilon;          /* 1e-6 */
    uint64_t max_pushes;     /* 1 000 000 */
    double queue_threshold;  /* 1e-8 */
    int adaptive_threshold;  /* 1 */
} orc_push_config;
void orc_push_config_default(orc_push_config *c);
typedef struct {
    uint64_t push_count, nodes_visited;
    double residual_norm;
} orc_push_stats;
int orc_forward_push(const orc_csr *adj, const orc_push_config *cfg, const uint64_t *sources, uint64_t nsources,
                     double *est, double *res, orc_push_stats *stats);
int orc_backward_push(const orc_csr *adj, const orc_push_config *cfg, const uint64_t *targets, uint64_t ntargets,
                      double *est, double *res, orc_push_stats *stats);
/* solve_with_target (forward_push.rs:234-290) / solve_with_source (backward_push.rs:238-290) / combine_with_forward
 * (backward_push.rs:312-330) */
int orc_forward_push_with_target(const orc_csr *adj, const orc_push_config *cfg, uint64_t source, uint64_t target,
                                 double target_precision, double *est, double *res, orc_push_stats *stats);
int orc_backward_push_with_source(const orc_csr *adj, const orc_push_config *cfg, uint64_t source, uint64_t target,
                                  double source_precision, double *est, double *res, orc_push_stats *stats);
double orc_push_combine_with_forward(double alpha, const double *best, const double *bres, uint64_t nb, const double *fest,
                                     const double *fres, uint64_t nf);
/* SublinearSolver.solveForwardPush (src/core/solver.ts:437-522) */
int orc_ts_forward_push(const orc_csr *a, const double *b, uint64_t blen, double epsilon, uint64_t max_iterations,
                        double *x, uint64_t *iterations, double *residual_norm, int *converged);

/* ---- conjugate gradient on the same SpMV ----
 * OptimizedConjugateGradientSolver::solve (src/optimized_solver.rs:182-295); the same loop is
 * FastConjugateGradient::solve (src/fast_solver.rs:126-178) and UltraFastCG::solve
 * (src/ultra_fast.rs:116-158), which only differ in the summation order of their dot products. */
enum {
    ORC_DOT_SEQUENTIAL = 0, /* `for .. { s += a*b }` (optimized_solver.rs:211-214, 229-232, 248-251) */
    ORC_DOT_CHUNK4 = 1,     /* sum += (p0+p1+p2+p3) per chunk of 4 (fast_solver.rs:180-200) */
    ORC_DOT_CHUNK8 = 2      /* sum += (p0+..+p7) per chunk of 8 (ultra_fast.rs:161-185) */
};
typedef struct {
    double *solution;      /* caller-provided buffer of n doubles */
    double residual_norm;  /* sqrt(rsold) at exit (optimized_solver.rs:275) */
    uint64_t iterations;
    int converged;
    uint64_t matvec_count;
    uint64_t total_flops;  /* matvec_count*nnz*2 + iterations*rows*6 (optimized_solver.rs:278-279) */
} orc_cg_result;
/* Returns ORC_OK, ORC_ERR_INVALID_INPUT ("Matrix must be square") or ORC_ERR_DIMENSION_MISMATCH. */
int orc_cg_solve(const orc_csr *m, const double *b, uint64_t blen, uint64_t max_iterations,
                 double tolerance, int spmv_variant, int dot_variant, int nthreads, orc_cg_result *res);

/* ---- synthetic inputs ---- */
/* create_test_matrix + create_test_rhs (benches/performance_benchmarks.rs:12-43): rows
 * [row0,row1) of the size x size system, emitted as CSR (= from_triplets of the generated triplets). */
uint64_t orc_gen_bench_k(uint64_t size, double sparsity);
int orc_gen_bench_csr(uint64_t size, double sparsity, uint64_t row0, uint64_t row1, orc_csr *out, double *b);
/* the same generator, as raw triplets in generation order (to exercise from_triplets). */
int64_t orc_gen_bench_triplets(uint64_t size, double sparsity, uint64_t *rows, uint64_t *cols,
                               double *vals, uint64_t cap);
/* generate_test_matrix (src/ultra_fast.rs:221-248): sequential LCG 12345, b = 1. */
int64_t orc_gen_ultra_triplets(uint64_t size, double sparsity, uint64_t *rows, uint64_t *cols,
                               double *vals, uint64_t cap);

/* ---- PageRank system (src/core/solver.ts:664-722): S = I - alpha P^T, dangling mass dropped ---- */
int orc_pagerank_system(const uint64_t *src, const uint64_t *dst, const double *w, uint64_t nedges,
                        uint64_t n, double alpha, orc_csr *S, double *rhs);

/* ---- single-entry estimation (spec sources src/core/solver.ts:390-432,550-659) ----
 * Ulam-von Neumann absorbing walk; counter-based RNG (splitmix64 of seed, query, walk, step). */
int orc_solve_entry(const orc_csr *m, const double *b, const uint64_t *rows, uint64_t nq,
                    uint64_t nwalks, uint64_t max_steps, uint64_t seed, double *est, double *var);
/* createSeededRandom (src/core/utils.ts:161-168): 32-bit LCG; returns next state, *u = state/2^32. */
uint32_t orc_ts_lcg_next(uint32_t state, double *u);

#ifdef __cplusplus
}
#endif
#endif
