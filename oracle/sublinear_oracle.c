/*
 * oracle/sublinear_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See sublinear_oracle.h for scope, pinning and the rule about who may call this.
 * Reference paths are relative to /root/reference (ruvnet/sublinear-time-solver @ 6e0dd66).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 */
#include "sublinear_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void orc_csr_free(orc_csr *m) {
    if (!m) return;
    free(m->values);
    free(m->col_indices);
    free(m->row_ptr);
    memset(m, 0, sizeof(*m));
}

/* ------------------------------------------------------------------------------------------ */
/* CSR construction                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t row;
    uint32_t col;
    double val;
} coo_entry;

/* Stable merge sort by (row, col): `sorted_entries.sort_by(|a,b| a.0.cmp(&b.0).then(a.1.cmp(&b.1)))`
 * (src/matrix/sparse.rs:95) — Rust's sort_by is stable, so equal (row,col) keep input order. */
static void merge_sort_entries(coo_entry *a, coo_entry *tmp, uint64_t n) {
    if (n < 2) return;
    uint64_t h = n / 2;
    merge_sort_entries(a, tmp, h);
    merge_sort_entries(a + h, tmp, n - h);
    uint64_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        int take_right = (a[j].row < a[i].row) || (a[j].row == a[i].row && a[j].col < a[i].col);
        tmp[k++] = take_right ? a[j++] : a[i++];
    }
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(coo_entry));
}

int orc_csr_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals,
                          uint64_t ntrip, uint64_t nrows, uint64_t ncols, orc_csr *out) {
    memset(out, 0, sizeof(*out));
    /* validation, in triplet order (src/matrix/mod.rs:166-187) */
    for (uint64_t i = 0; i < ntrip; i++) {
        if (rows[i] >= nrows) return ORC_ERR_INDEX_OUT_OF_BOUNDS;
        if (cols[i] >= ncols) return ORC_ERR_INDEX_OUT_OF_BOUNDS;
        if (!isfinite(vals[i])) return ORC_ERR_INVALID_INPUT;
    }
    /* COOStorage::from_triplets drops exact zeros (src/matrix/sparse.rs:535-541) */
    coo_entry *e = (coo_entry *)malloc((ntrip ? ntrip : 1) * sizeof(coo_entry));
    coo_entry *tmp = (coo_entry *)malloc((ntrip ? ntrip : 1) * sizeof(coo_entry));
    if (!e || !tmp) { free(e); free(tmp); return ORC_ERR_MEMORY_ALLOCATION; }
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < ntrip; i++) {
        if (vals[i] != 0.0) {
            e[cnt].row = rows[i];
            e[cnt].col = (uint32_t)cols[i];
            e[cnt].val = vals[i];
            cnt++;
        }
    }
    out->nrows = nrows;
    out->ncols = ncols;
    out->row_ptr = (uint32_t *)calloc(nrows + 1, sizeof(uint32_t));
    out->values = (double *)malloc((cnt ? cnt : 1) * sizeof(double));
    out->col_indices = (uint32_t *)malloc((cnt ? cnt : 1) * sizeof(uint32_t));
    if (!out->row_ptr || !out->values || !out->col_indices) {
        free(e); free(tmp); orc_csr_free(out);
        return ORC_ERR_MEMORY_ALLOCATION;
    }
    /* CSRStorage::from_coo (src/matrix/sparse.rs:80-132); empty -> all-zero row_ptr (:81-87) */
    merge_sort_entries(e, tmp, cnt);
    uint64_t current_row = 0, nnz = 0;
    for (uint64_t i = 0; i < cnt; i++) {
        while (current_row < e[i].row) { /* :104-107 */
            current_row++;
            out->row_ptr[current_row] = (uint32_t)nnz;
        }
        out->values[nnz] = e[i].val;
        out->col_indices[nnz] = e[i].col;
        nnz++;
    }
    while (current_row < nrows) { /* :115-118 */
        current_row++;
        out->row_ptr[current_row] = (uint32_t)nnz;
    }
    out->nnz = nnz;
    free(e);
    free(tmp);
    return ORC_OK;
}

/* CSRStorage::get (src/matrix/sparse.rs:142-155).  Rust's slice::binary_search leaves the choice
 * among duplicate keys unspecified; this restatement returns the probe hit of the classic
 * lo/hi bisection (documented in DESIGN.md as "any duplicate"). */
int orc_csr_get(const orc_csr *m, uint64_t row, uint64_t col, double *out) {
    if (row >= m->nrows || col >= m->ncols) return 0; /* src/matrix/mod.rs:395-397 */
    uint64_t lo = m->row_ptr[row], hi = m->row_ptr[row + 1];
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        uint32_t c = m->col_indices[mid];
        if (c == (uint32_t)col) { if (out) *out = m->values[mid]; return 1; }
        if (c < (uint32_t)col) lo = mid + 1; else hi = mid;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* SpMV variants                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* src/matrix/sparse.rs:187-203: fill(0.0) then `*row_sum += values[i] * x[col]` left to right. */
void orc_spmv_scalar(const orc_csr *m, const double *x, double *y) {
    for (uint64_t row = 0; row < m->nrows; row++) {
        double acc = 0.0;
        for (uint64_t i = m->row_ptr[row]; i < m->row_ptr[row + 1]; i++)
            acc += m->values[i] * x[m->col_indices[i]];
        y[row] = acc;
    }
}

/* src/simd_ops.rs:20-88: rows with nnz >= 8 use four lane accumulators over chunks of 4,
 * horizontal sum ((l0+l1)+l2)+l3, then the scalar tail is added into y[row]; short rows scalar. */
void orc_spmv_simd4(const orc_csr *m, const double *x, double *y) {
    for (uint64_t row = 0; row < m->nrows; row++) {
        uint64_t start = m->row_ptr[row], end = m->row_ptr[row + 1];
        if (end <= start) { y[row] = 0.0; continue; }
        const double *v = m->values + start;
        const uint32_t *c = m->col_indices + start;
        uint64_t nnz = end - start;
        if (nnz >= 8) {
            uint64_t chunks = nnz / 4;
            double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
            for (uint64_t ch = 0; ch < chunks; ch++) {
                uint64_t i = ch * 4;
                l0 = l0 + v[i] * x[c[i]];
                l1 = l1 + v[i + 1] * x[c[i + 1]];
                l2 = l2 + v[i + 2] * x[c[i + 2]];
                l3 = l3 + v[i + 3] * x[c[i + 3]];
            }
            double acc = l0 + l1 + l2 + l3;
            for (uint64_t i = chunks * 4; i < nnz; i++) acc += v[i] * x[c[i]];
            y[row] = acc;
        } else {
            double acc = 0.0;
            for (uint64_t i = 0; i < nnz; i++) acc += v[i] * x[c[i]];
            y[row] = acc;
        }
    }
}

/* src/simd_ops.rs:202-239: contiguous chunks of ceil(rows/threads) rows, one task per chunk,
 * scalar accumulation inside (rayon par_chunks_mut -> OpenMP, one chunk per thread). */
void orc_spmv_parallel(const orc_csr *m, const double *x, double *y, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    uint64_t rows = m->nrows;
    if (rows == 0) return;
    uint64_t chunk = (rows + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    int64_t nchunks = (int64_t)((rows + chunk - 1) / chunk);
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int64_t ci = 0; ci < nchunks; ci++) {
        uint64_t r0 = (uint64_t)ci * chunk, r1 = r0 + chunk;
        if (r1 > rows) r1 = rows;
        for (uint64_t row = r0; row < r1; row++) {
            double acc = 0.0;
            for (uint64_t i = m->row_ptr[row]; i < m->row_ptr[row + 1]; i++)
                acc += m->values[i] * x[m->col_indices[i]];
            y[row] = acc;
        }
    }
}

static void spmv_dispatch(const orc_csr *m, const double *x, double *y, int variant, int nthreads) {
    if (variant == ORC_SPMV_SIMD4) orc_spmv_simd4(m, x, y);
    else if (variant == ORC_SPMV_PARALLEL) orc_spmv_parallel(m, x, y, nthreads);
    else orc_spmv_scalar(m, x, y);
}

/* src/matrix/mod.rs:415-439 */
int orc_multiply_vector(const orc_csr *m, const double *x, uint64_t xlen, double *y, uint64_t ylen,
                        int variant, int nthreads) {
    if (xlen != m->ncols) return ORC_ERR_DIMENSION_MISMATCH;
    if (ylen != m->nrows) return ORC_ERR_DIMENSION_MISMATCH;
    spmv_dispatch(m, x, y, variant, nthreads);
    return ORC_OK;
}

/* src/matrix/mod.rs:467-485: `diagonal = value.abs()` is overwritten by each diagonal entry (the
 * last duplicate wins), a missing diagonal is 0, failure iff diagonal < off_diagonal_sum. */
int orc_is_diagonally_dominant(const orc_csr *m, uint64_t *first_bad_row) {
    if (first_bad_row) *first_bad_row = UINT64_MAX;
    for (uint64_t row = 0; row < m->nrows; row++) {
        double diagonal = 0.0, off = 0.0;
        for (uint64_t i = m->row_ptr[row]; i < m->row_ptr[row + 1]; i++) {
            if ((uint64_t)m->col_indices[i] == row) diagonal = fabs(m->values[i]);
            else off += fabs(m->values[i]);
        }
        if (diagonal < off) {
            if (first_bad_row) *first_bad_row = row;
            return 0;
        }
    }
    return 1;
}

/* Column-wise analogue (TS analyzeMatrix accepts row OR column dominance, src/core/matrix.ts:343-345). */
int orc_is_col_diagonally_dominant(const orc_csr *m) {
    uint64_t n = m->ncols;
    double *diag = (double *)calloc(n ? n : 1, sizeof(double));
    double *off = (double *)calloc(n ? n : 1, sizeof(double));
    for (uint64_t row = 0; row < m->nrows; row++)
        for (uint64_t i = m->row_ptr[row]; i < m->row_ptr[row + 1]; i++) {
            uint32_t c = m->col_indices[i];
            if ((uint64_t)c == row) diag[c] = fabs(m->values[i]);
            else off[c] += fabs(m->values[i]);
        }
    int ok = 1;
    for (uint64_t c = 0; c < n; c++)
        if (diag[c] < off[c]) { ok = 0; break; }
    free(diag);
    free(off);
    return ok;
}

/* src/solver/mod.rs:369-381 */
double orc_l2_norm(const double *v, uint64_t n) {
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s += v[i] * v[i];
    return sqrt(s);
}
double orc_l1_norm(const double *v, uint64_t n) {
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s += fabs(v[i]);
    return s;
}
double orc_linf_norm(const double *v, uint64_t n) {
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s = fmax(s, fabs(v[i]));
    return s;
}

/* src/simd_ops.rs:116-147 */
double orc_dot_simd4(const double *x, const double *y, uint64_t n) {
    uint64_t chunks = n / 4;
    double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
    for (uint64_t ch = 0; ch < chunks; ch++) {
        uint64_t i = ch * 4;
        l0 = l0 + x[i] * y[i];
        l1 = l1 + x[i + 1] * y[i + 1];
        l2 = l2 + x[i + 2] * y[i + 2];
        l3 = l3 + x[i + 3] * y[i + 3];
    }
    double r = l0 + l1 + l2 + l3;
    for (uint64_t i = chunks * 4; i < n; i++) r += x[i] * y[i];
    return r;
}

/* src/simd_ops.rs:158-189: y = alpha*x + y */
void orc_axpy_simd4(double alpha, const double *x, double *y, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) y[i] = (alpha * x[i]) + y[i];
}

/* ------------------------------------------------------------------------------------------ */
/* Neumann solver                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* SolverOptions::default (src/solver/mod.rs:47-63), NeumannSolver::default (src/solver/neumann.rs:58-60) */
void orc_options_default(orc_options *o) {
    memset(o, 0, sizeof(*o));
    o->tolerance = 1e-6;
    o->max_iterations = 1000;
    o->max_terms = 50;
    o->series_tolerance = 1e-8;
    o->adaptive_truncation = 1;
    o->mode = ORC_MODE_CORRECT;
    o->dominance = ORC_DOM_ROW;
    o->spmv_variant = ORC_SPMV_SCALAR;
    o->nthreads = 0;
}

typedef struct {
    const orc_csr *m;
    uint64_t n;
    double *solution, *rhs, *residual, *dinv, *term, *temp;
    double *b;               /* own copy of the right-hand side (update_rhs changes it) */
    int mode;
    const double *resid_rhs; /* c in ref_compat, b in correct mode */
    double residual_norm, term_norm, rhs_norm;
    uint64_t terms, matvec, max_terms;
    int series_converged;
    double tolerance, series_tolerance;
    int variant, nthreads;
    int has_bound;
    double bound;
} nstate;

/* compute_next_term (src/solver/neumann.rs:252-277) + apply_iteration_matrix (:280-299) */
static void compute_next_term(nstate *s) {
    if (s->terms >= s->max_terms) return; /* :253-255 */
    if (s->terms > 0) {
        spmv_dispatch(s->m, s->term, s->temp, s->variant, s->nthreads); /* :285 */
        s->matvec++;
        for (uint64_t i = 0; i < s->n; i++) s->temp[i] *= s->dinv[i];  /* :289-291 */
        for (uint64_t i = 0; i < s->n; i++) s->term[i] -= s->temp[i];  /* :294-296 */
    }
    for (uint64_t i = 0; i < s->n; i++) s->solution[i] += s->term[i];  /* :264-266 */
    s->terms++;
    s->term_norm = orc_l2_norm(s->term, s->n);                          /* :271 */
    if (s->term_norm < s->series_tolerance) s->series_converged = 1;
}

/* update_residual (src/solver/neumann.rs:302-318): r = A*x - rhs, rhs = D^-1 b in the reference
 * (its own comment at :312-314 admits it); correct mode subtracts the true b. */
static void update_residual(nstate *s) {
    spmv_dispatch(s->m, s->solution, s->residual, s->variant, s->nthreads);
    s->matvec++;
    for (uint64_t i = 0; i < s->n; i++) s->residual[i] = s->residual[i] - s->resid_rhs[i];
    s->residual_norm = orc_l2_norm(s->residual, s->n);
}

/* estimate_error_bounds (src/solver/neumann.rs:321-347) */
static void estimate_error_bounds(nstate *s) {
    if (!s->series_converged || s->terms == 0) return;
    double est = 0.0;
    if (s->terms > 1) {
        double ratio = s->term_norm / s->rhs_norm;
        est = pow(ratio, 1.0 / (double)(s->terms - 1));
    }
    if (est < 1.0) {
        double remaining = pow(est, (double)(int)s->terms) / (1.0 - est);
        s->bound = remaining * s->rhs_norm;
        s->has_bound = 1;
    }
}

/* is_converged (src/solver/neumann.rs:422-430) */
static int is_converged(const nstate *s) {
    int residual_converged = s->residual_norm <= s->tolerance;
    int max_terms_reached = s->terms >= s->max_terms;
    return residual_converged || (s->series_converged && !max_terms_reached);
}

/* NeumannState::new (src/solver/neumann.rs:139-249). Allocates; the caller releases with state_release. */
static int state_init(nstate *sp, const orc_csr *m, const double *b, uint64_t blen, const orc_options *opt) {
    uint64_t n = m->nrows;
    memset(sp, 0, sizeof(*sp));
    if (m->nrows != m->ncols) return ORC_ERR_INVALID_INPUT;          /* :147-152 */
    if (blen != n) return ORC_ERR_DIMENSION_MISMATCH;                /* :154-160 */
    int dd = orc_is_diagonally_dominant(m, NULL);                    /* :163-169 */
    if (!dd && opt->dominance == ORC_DOM_ROW_OR_COL) dd = orc_is_col_diagonally_dominant(m);
    if (!dd) return ORC_ERR_NOT_DIAGONALLY_DOMINANT;

    nstate s;
    memset(&s, 0, sizeof(s));
    s.m = m;
    s.n = n;
    s.mode = opt->mode;
    size_t bytes = (n ? n : 1) * sizeof(double);
    s.solution = (double *)malloc(bytes);
    s.rhs = (double *)malloc(bytes);
    s.residual = (double *)calloc(n ? n : 1, sizeof(double));
    s.dinv = (double *)calloc(n ? n : 1, sizeof(double));
    s.term = (double *)malloc(bytes);
    s.temp = (double *)malloc(bytes);
    s.b = (double *)malloc(bytes);
    int rc = ORC_OK;
    if (!s.solution || !s.rhs || !s.residual || !s.dinv || !s.term || !s.temp || !s.b) {
        rc = ORC_ERR_MEMORY_ALLOCATION;
        goto fail;
    }
    memcpy(s.b, b, n * sizeof(double));
    for (uint64_t i = 0; i < n; i++) { /* :172-188 */
        double d;
        if (orc_csr_get(m, i, i, &d)) {
            if (opt->mode == ORC_MODE_CORRECT) {
                /* duplicated diagonal entries: the SpMV sums them, so the correct D does too
                 * (SURVEY.md Appendix B rule 1); identical to get() when the entry is unique. */
                double dsum = 0.0;
                for (uint64_t k = m->row_ptr[i]; k < m->row_ptr[i + 1]; k++)
                    if ((uint64_t)m->col_indices[k] == i) dsum += m->values[k];
                d = dsum;
            }
            if (fabs(d) < 1e-14) { rc = ORC_ERR_INVALID_SPARSE_MATRIX; goto fail; }
            s.dinv[i] = 1.0 / d;
        } else {
            rc = ORC_ERR_INVALID_SPARSE_MATRIX;
            goto fail;
        }
    }
    for (uint64_t i = 0; i < n; i++) s.rhs[i] = b[i] * s.dinv[i]; /* :191-194 */
    s.rhs_norm = orc_l2_norm(s.rhs, n);
    if (opt->initial_guess) {                                       /* :197-206 */
        if (opt->initial_guess_len != n) { rc = ORC_ERR_DIMENSION_MISMATCH; goto fail; }
        memcpy(s.solution, opt->initial_guess, n * sizeof(double));
    } else if (opt->mode == ORC_MODE_REF_COMPAT) {
        memcpy(s.solution, s.rhs, n * sizeof(double));              /* x_0 = D^-1 b (F4 quirk) */
    } else {
        memset(s.solution, 0, n * sizeof(double));
    }
    memcpy(s.term, s.rhs, n * sizeof(double));                      /* :211 */
    s.variant = opt->spmv_variant;
    s.nthreads = opt->nthreads;
    if (opt->mode == ORC_MODE_CORRECT && opt->initial_guess) {
        /* correct mode with x0 != 0: t0 = D^-1 (b - A x0) (last paragraph) */
        spmv_dispatch(m, s.solution, s.temp, s.variant, s.nthreads);
        s.matvec++;
        for (uint64_t i = 0; i < n; i++) s.term[i] = (b[i] - s.temp[i]) * s.dinv[i];
    }
    s.resid_rhs = (opt->mode == ORC_MODE_REF_COMPAT) ? s.rhs : s.b;
    s.residual_norm = INFINITY;                                     /* :236 */
    s.tolerance = opt->tolerance;
    s.max_terms = opt->max_terms;
    s.series_tolerance = opt->series_tolerance;
    *sp = s;
    return ORC_OK;
fail:
    free(s.solution); free(s.rhs); free(s.residual); free(s.dinv); free(s.term); free(s.temp); free(s.b);
    return rc;
}

static void state_release(nstate *s) {
    free(s->solution); free(s->rhs); free(s->residual); free(s->dinv); free(s->term); free(s->temp); free(s->b);
    memset(s, 0, sizeof(*s));
}

int orc_neumann_solve(const orc_csr *m, const double *b, uint64_t blen, const orc_options *opt,
                      orc_result *res) {
    double t_start = now_s();
    uint64_t n = m->nrows;
    nstate s;
    int rc = state_init(&s, m, b, blen, opt);
    if (rc != ORC_OK) {
        res->total_time_ms = (now_s() - t_start) * 1e3;
        return rc;
    }

    /* NeumannSolver::solve (src/solver/neumann.rs:469-555) */
    uint64_t iterations = 0;
    while (!is_converged(&s) && iterations < opt->max_iterations) { /* :481 */
        compute_next_term(&s);                                      /* :486 */
        if (iterations % 5 == 0) update_residual(&s);               /* :489-491 */
        if (opt->compute_error_bounds && opt->adaptive_truncation) estimate_error_bounds(&s); /* :494-496 */
        iterations++;                                               /* :498 */
        if (!isfinite(s.residual_norm)) { rc = ORC_ERR_NUMERICAL_INSTABILITY; break; } /* :501-507 */
        if (s.series_converged) break;                              /* :510-512 */
    }
    if (rc == ORC_OK) {
        update_residual(&s);                                        /* :516 */
        int converged = is_converged(&s);                           /* :518 */
        if (!converged && iterations >= opt->max_iterations) rc = ORC_ERR_CONVERGENCE_FAILURE; /* :523-530 */
        res->converged = converged;
    } else {
        res->converged = 0;
    }
    if (res->solution) memcpy(res->solution, s.solution, n * sizeof(double));
    res->residual_norm = s.residual_norm;
    res->iterations = iterations;
    res->terms_computed = s.terms;
    res->matvec_count = s.matvec;
    res->series_converged = s.series_converged;
    res->has_error_bound = opt->compute_error_bounds ? s.has_bound : 0;
    res->error_bound = s.bound;
    res->last_term_norm = s.term_norm;
    res->total_time_ms = (now_s() - t_start) * 1e3;
    state_release(&s);
    return rc;
}

/* ---- the SolverAlgorithm state interface (src/solver/mod.rs:223-252; neumann.rs:350-462) ---- */
struct orc_state {
    nstate s;
    int adaptive_truncation;
};

/* SolverAlgorithm::initialize (neumann.rs:381-388) */
int orc_state_new(const orc_csr *m, const double *b, uint64_t blen, const orc_options *opt, orc_state **out) {
    orc_state *st = (orc_state *)calloc(1, sizeof(orc_state));
    if (!st) return ORC_ERR_MEMORY_ALLOCATION;
    int rc = state_init(&st->s, m, b, blen, opt);
    if (rc != ORC_OK) { free(st); return rc; }
    st->adaptive_truncation = opt->adaptive_truncation;
    *out = st;
    return ORC_OK;
}

/* SolverAlgorithm::step as the reference intends it (the body it left commented out for lack of a matrix
 * reference, neumann.rs:404-418): next term, residual, error bounds; Converged (1) once the series converged or
 * max_terms is reached, else Continue (0). */
int orc_state_step(orc_state *st, int *step_result) {
    compute_next_term(&st->s);
    update_residual(&st->s);
    if (st->adaptive_truncation) estimate_error_bounds(&st->s);
    *step_result = (st->s.series_converged || st->s.terms >= st->s.max_terms) ? 1 : 0;
    return isfinite(st->s.residual_norm) ? ORC_OK : ORC_ERR_NUMERICAL_INSTABILITY;
}

int orc_state_is_converged(const orc_state *st) { return is_converged(&st->s); }

void orc_state_solution(const orc_state *st, double *x) { memcpy(x, st->s.solution, st->s.n * sizeof(double)); }

/* SolverAlgorithm::update_rhs (neumann.rs:436-462). ref_compat: the literal code (scaled delta added to rhs AND to
 * the solution, series restarted from the whole new rhs). correct mode: the incremental solve the comment at
 * :451-453 asks for — b and rhs take the delta, the series restarts from t = D^-1 delta_b only, so the following steps
 * add A^-1 delta_b to the solution already held. */
int orc_state_update_rhs(orc_state *st, const uint64_t *idx, const double *delta, uint64_t count) {
    nstate *s = &st->s;
    for (uint64_t k = 0; k < count; k++)
        if (idx[k] >= s->n) return ORC_ERR_INDEX_OUT_OF_BOUNDS;  /* :439-445 (checked up front: no partial update) */
    if (s->mode == ORC_MODE_CORRECT) memset(s->term, 0, s->n * sizeof(double));
    for (uint64_t k = 0; k < count; k++) {
        double scaled = delta[k] * s->dinv[idx[k]];               /* :448 */
        s->rhs[idx[k]] += scaled;                                 /* :449 */
        s->b[idx[k]] += delta[k];
        if (s->mode == ORC_MODE_CORRECT) s->term[idx[k]] += scaled;
        else s->solution[idx[k]] += scaled;                       /* :453 */
    }
    if (s->mode != ORC_MODE_CORRECT) memcpy(s->term, s->rhs, s->n * sizeof(double)); /* :457 */
    s->terms = 0;                                                 /* :458 */
    s->series_converged = 0;                                      /* :459 */
    s->rhs_norm = orc_l2_norm(s->rhs, s->n);
    return ORC_OK;
}

/* SolverState::reset (neumann.rs:367-378) */
void orc_state_reset(orc_state *st) {
    nstate *s = &st->s;
    memset(s->solution, 0, s->n * sizeof(double));
    memset(s->residual, 0, s->n * sizeof(double));
    s->residual_norm = INFINITY;
    memcpy(s->term, s->rhs, s->n * sizeof(double));
    s->terms = 0;
    s->matvec = 0;
    s->series_converged = 0;
    s->has_bound = 0;
}

void orc_state_info(const orc_state *st, double *residual_norm, uint64_t *matvec_count, uint64_t *terms_computed,
                    int *series_converged, double *term_norm, int *has_bound, double *bound) {
    if (residual_norm) *residual_norm = st->s.residual_norm;
    if (matvec_count) *matvec_count = st->s.matvec;
    if (terms_computed) *terms_computed = st->s.terms;
    if (series_converged) *series_converged = st->s.series_converged;
    if (term_norm) *term_norm = st->s.term_norm;
    if (has_bound) *has_bound = st->s.has_bound;
    if (bound) *bound = st->s.bound;
}

void orc_state_free(orc_state *st) {
    if (!st) return;
    state_release(&st->s);
    free(st);
}

/* The bare push recurrence (neumann.rs:280-299 + :264-266 + :271) for `nterms` terms after term 0,
 * from t = c = D^-1 b, x = c.  Returns seconds spent in the nterms iterations only. */
double orc_push_iterations(const orc_csr *m, const double *b, uint64_t nterms, int spmv_variant,
                           int nthreads, double *x_out, double *t_out, double *term_norms) {
    uint64_t n = m->nrows;
    double *dinv = (double *)malloc((n ? n : 1) * sizeof(double));
    double *t = (double *)malloc((n ? n : 1) * sizeof(double));
    double *x = (double *)malloc((n ? n : 1) * sizeof(double));
    double *tmp = (double *)malloc((n ? n : 1) * sizeof(double));
    for (uint64_t i = 0; i < n; i++) {
        double d = 0.0;
        for (uint64_t k = m->row_ptr[i]; k < m->row_ptr[i + 1]; k++)
            if ((uint64_t)m->col_indices[k] == i) d += m->values[k];
        dinv[i] = 1.0 / d;
        t[i] = b[i] * dinv[i];
        x[i] = t[i];
    }
    double t0 = now_s();
    for (uint64_t k = 0; k < nterms; k++) {
        spmv_dispatch(m, t, tmp, spmv_variant, nthreads);
        for (uint64_t i = 0; i < n; i++) tmp[i] *= dinv[i];
        for (uint64_t i = 0; i < n; i++) t[i] -= tmp[i];
        for (uint64_t i = 0; i < n; i++) x[i] += t[i];
        double nrm = orc_l2_norm(t, n);
        if (term_norms) term_norms[k] = nrm;
    }
    double dt = now_s() - t0;
    if (x_out) memcpy(x_out, x, n * sizeof(double));
    if (t_out) memcpy(t_out, t, n * sizeof(double));
    free(dinv); free(t); free(x); free(tmp);
    return dt;
}

/* ------------------------------------------------------------------------------------------ */
/* forward / backward push (src/solver/forward_push.rs, src/solver/backward_push.rs)           */
/* ------------------------------------------------------------------------------------------ */
void orc_push_config_default(orc_push_config *c) { /* forward_push.rs:40-50 = backward_push.rs:40-50 */
    c->alpha = 0.15;
    c->epsilon = 1e-6;
    c->max_pushes = 1000000;
    c->queue_threshold = 1e-8;
    c->adaptive_threshold = 1;
}

/* WorkQueue (src/graph/mod.rs:130-212): binary max-heap of (priority, node) + in_queue bit set + threshold */
typedef struct { double pr; uint64_t node; } witem;
typedef struct {
    witem *heap;
    uint64_t len, cap;
    uint8_t *in_queue;
    double threshold;
} wqueue;

static int witem_less(const witem *a, const witem *b) { /* derive(PartialOrd): priority, then node_id */
    if (a->pr != b->pr) return a->pr < b->pr;
    return a->node < b->node;
}

static void wq_push_if_threshold(wqueue *q, uint64_t node, double residual, double degree) { /* :171-181 */
    double priority = degree > 0.0 ? residual / degree : residual;
    if (!(priority >= q->threshold) || q->in_queue[node]) return;
    if (q->len == q->cap) {
        q->cap = q->cap ? q->cap * 2 : 64;
        q->heap = (witem *)realloc(q->heap, q->cap * sizeof(witem));
    }
    uint64_t i = q->len++;
    witem it = {priority, node};
    while (i > 0) { /* sift up */
        uint64_t parent = (i - 1) / 2;
        if (!witem_less(&q->heap[parent], &it)) break;
        q->heap[i] = q->heap[parent];
        i = parent;
    }
    q->heap[i] = it;
    q->in_queue[node] = 1;
}

static int wq_pop(wqueue *q, uint64_t *node) { /* :184-191 */
    if (q->len == 0) return 0;
    witem top = q->heap[0];
    witem last = q->heap[--q->len];
    uint64_t i = 0;
    for (;;) { /* sift down */
        uint64_t l = 2 * i + 1, r = l + 1, big = i;
        const witem *cur = &last;
        if (l < q->len && witem_less(cur, &q->heap[l])) { big = l; cur = &q->heap[l]; }
        if (r < q->len && witem_less(cur, &q->heap[r])) { big = r; }
        if (big == i) break;
        q->heap[i] = q->heap[big];
        i = big;
    }
    if (q->len) q->heap[i] = last;
    q->in_queue[top.node] = 0;
    *node = top.node;
    return 1;
}

static void wq_adaptive(wqueue *q, uint64_t max_size, uint64_t min_size) { /* :204-212 */
    if (q->len > max_size) q->threshold *= 1.1;
    else if (q->len < min_size && q->threshold > 1e-12) q->threshold *= 0.9;
}

/* direction 0: forward (mass moves along out-edges, thresholds on the out-degree);
 * direction 1: backward (mass moves to predecessors with weight / max(out_degree(pred), 1), thresholds on the in-degree) */
static int push_run(const orc_csr *adj, const orc_push_config *cfg, const uint64_t *seeds, uint64_t nseeds,
                    int direction, int watch, uint64_t watch_node, double watch_precision, double *est, double *res,
                    orc_push_stats *stats) {
    if (adj->nrows != adj->ncols) return ORC_ERR_INVALID_INPUT;
    uint64_t n = adj->nrows;
    double *deg = (double *)calloc(n ? n : 1, sizeof(double));   /* row sums (adjacency.rs:214) */
    double *rdeg = (double *)calloc(n ? n : 1, sizeof(double));  /* column sums (adjacency.rs:215) */
    uint8_t *visited = (uint8_t *)calloc(n ? n : 1, 1);
    wqueue q = {NULL, 0, 0, (uint8_t *)calloc(n ? n : 1, 1), cfg->queue_threshold};
    /* transpose for the backward direction (mod.rs:93-127) */
    uint32_t *tptr = NULL, *tcol = NULL;
    double *tval = NULL;
    for (uint64_t u = 0; u < n; u++)
        for (uint64_t k = adj->row_ptr[u]; k < adj->row_ptr[u + 1]; k++) {
            deg[u] += adj->values[k];
            rdeg[adj->col_indices[k]] += adj->values[k];
        }
    if (direction == 1) {
        tptr = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
        tcol = (uint32_t *)malloc((adj->nnz ? adj->nnz : 1) * sizeof(uint32_t));
        tval = (double *)malloc((adj->nnz ? adj->nnz : 1) * sizeof(double));
        for (uint64_t k = 0; k < adj->nnz; k++) tptr[adj->col_indices[k] + 1]++;
        for (uint64_t i = 0; i < n; i++) tptr[i + 1] += tptr[i];
        uint32_t *pos = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        memcpy(pos, tptr, n * sizeof(uint32_t));
        for (uint64_t u = 0; u < n; u++)
            for (uint64_t k = adj->row_ptr[u]; k < adj->row_ptr[u + 1]; k++) {
                uint32_t c = adj->col_indices[k];
                tcol[pos[c]] = (uint32_t)u;
                tval[pos[c]] = adj->values[k];
                pos[c]++;
            }
        free(pos);
    }
    const double *tdeg = direction ? rdeg : deg; /* the degree the thresholds use */
    for (uint64_t i = 0; i < n; i++) est[i] = res[i] = 0.0;
    uint64_t push_count = 0, nvisited = 0;
    int any = 0;
    if (nseeds == 1) { /* solve_single_source / solve_single_target: unit mass, out of range -> zero result */
        if (seeds[0] < n) { res[seeds[0]] = 1.0; any = 1; }
    } else if (nseeds > 1) { /* solve_multi_*: 1/len mass per listed seed, out-of-range ones are skipped */
        double mass = 1.0 / (double)nseeds;
        for (uint64_t s = 0; s < nseeds; s++) if (seeds[s] < n) { res[seeds[s]] += mass; any = 1; }
    }
    if (any)
        for (uint64_t s = 0; s < nseeds; s++)
            if (seeds[s] < n) wq_push_if_threshold(&q, seeds[s], res[seeds[s]], fmax(tdeg[seeds[s]], 1.0));
    uint64_t node;
    while (q.len > 0 && push_count < cfg->max_pushes) {
        /* solve_with_target (forward_push.rs:260-263) / solve_with_source (backward_push.rs:263-266): before every pop */
        if (watch && est[watch_node] > watch_precision && res[watch_node] < watch_precision * 0.1) break;
        if (!wq_pop(&q, &node)) break;
        if (res[node] < cfg->epsilon * fmax(tdeg[node], 1.0)) continue;
        /* push_node / backward_push_node */
        if (res[node] > 0.0) {
            est[node] += cfg->alpha * res[node];
            double remaining = (1.0 - cfg->alpha) * res[node];
            res[node] = 0.0;
            if (tdeg[node] > 0.0) {
                if (direction == 0) {
                    for (uint64_t k = adj->row_ptr[node]; k < adj->row_ptr[node + 1]; k++) {
                        uint64_t v = adj->col_indices[k];
                        res[v] += remaining * adj->values[k] / deg[node];
                        wq_push_if_threshold(&q, v, res[v], fmax(deg[v], 1.0));
                    }
                } else {
                    for (uint64_t k = tptr[node]; k < tptr[node + 1]; k++) {
                        uint64_t p = tcol[k];
                        double transition = tval[k] / fmax(deg[p], 1.0);
                        res[p] += remaining * transition;
                        wq_push_if_threshold(&q, p, res[p], fmax(rdeg[p], 1.0));
                    }
                }
            } else { /* no edges in the push direction: the mass stays on the node */
                res[node] += remaining;
                wq_push_if_threshold(&q, node, res[node], 1.0);
            }
        }
        if (!visited[node]) { visited[node] = 1; nvisited++; }
        push_count++;
        if (cfg->adaptive_threshold && push_count % 1000 == 0) wq_adaptive(&q, 10000, 100);
    }
    double nrm = 0.0;
    for (uint64_t i = 0; i < n; i++) nrm += res[i] * res[i];
    stats->push_count = push_count;
    stats->nodes_visited = nvisited;
    stats->residual_norm = any ? sqrt(nrm) : 0.0;
    free(deg); free(rdeg); free(visited); free(q.heap); free(q.in_queue); free(tptr); free(tcol); free(tval);
    return ORC_OK;
}

int orc_forward_push(const orc_csr *adj, const orc_push_config *cfg, const uint64_t *sources, uint64_t nsources,
                     double *est, double *res, orc_push_stats *stats) {
    return push_run(adj, cfg, sources, nsources, 0, 0, 0, 0.0, est, res, stats);
}

int orc_backward_push(const orc_csr *adj, const orc_push_config *cfg, const uint64_t *targets, uint64_t ntargets,
                      double *est, double *res, orc_push_stats *stats) {
    return push_run(adj, cfg, targets, ntargets, 1, 0, 0, 0.0, est, res, stats);
}

/* ForwardPushSolver::solve_with_target (forward_push.rs:234-290): out-of-range source or target -> all-zero result */
int orc_forward_push_with_target(const orc_csr *adj, const orc_push_config *cfg, uint64_t source, uint64_t target,
                                 double target_precision, double *est, double *res, orc_push_stats *stats) {
    uint64_t seed = (source >= adj->nrows || target >= adj->nrows) ? adj->nrows : source;
    return push_run(adj, cfg, &seed, 1, 0, seed < adj->nrows, target, target_precision, est, res, stats);
}

/* BackwardPushSolver::solve_with_source (backward_push.rs:238-290) */
int orc_backward_push_with_source(const orc_csr *adj, const orc_push_config *cfg, uint64_t source, uint64_t target,
                                  double source_precision, double *est, double *res, orc_push_stats *stats) {
    uint64_t seed = (source >= adj->nrows || target >= adj->nrows) ? adj->nrows : target;
    return push_run(adj, cfg, &seed, 1, 1, seed < adj->nrows, source, source_precision, est, res, stats);
}

/* BackwardPushSolver::combine_with_forward (backward_push.rs:312-330) */
double orc_push_combine_with_forward(double alpha, const double *best, const double *bres, uint64_t nb, const double *fest,
                                     const double *fres, uint64_t nf) {
    double total = 0.0;
    uint64_t m = nb < nf ? nb : nf;
    for (uint64_t i = 0; i < m; i++) {
        total += best[i] * fest[i];
        total += bres[i] * fest[i] * alpha;
        total += best[i] * fres[i] * alpha;
    }
    return total;
}

/* SublinearSolver.solveForwardPush (src/core/solver.ts:437-522): Gauss-Southwell push on the residual of A x = b.
 * The column walk `for j != maxNode: residual[j] -= getEntry(j, maxNode) * pushValue` runs over the transposed copy
 * (entries of column maxNode in ascending row order; duplicate entries are applied one after the other, getEntry's
 * first-match rule differs only for duplicated coordinates, which this path does not produce).
 * Returns ORC_OK, ORC_ERR_CONVERGENCE_FAILURE (maxIterations exhausted, :505-511) or ORC_ERR_NUMERICAL_INSTABILITY
 * (|diagonal| < 1e-15 under the pushed node, :468-471). */
int orc_ts_forward_push(const orc_csr *a, const double *b, uint64_t blen, double epsilon, uint64_t max_iterations,
                        double *x, uint64_t *iterations, double *residual_norm, int *converged) {
    if (a->nrows != a->ncols || blen != a->nrows) return ORC_ERR_DIMENSION_MISMATCH;
    uint64_t n = a->nrows;
    double *r = (double *)malloc((n ? n : 1) * sizeof(double));
    double *diag = (double *)calloc(n ? n : 1, sizeof(double));
    uint32_t *tptr = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
    uint32_t *trow = (uint32_t *)malloc((a->nnz ? a->nnz : 1) * sizeof(uint32_t));
    double *tval = (double *)malloc((a->nnz ? a->nnz : 1) * sizeof(double));
    for (uint64_t k = 0; k < a->nnz; k++) tptr[a->col_indices[k] + 1]++;
    for (uint64_t i = 0; i < n; i++) tptr[i + 1] += tptr[i];
    {
        uint32_t *pos = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        memcpy(pos, tptr, n * sizeof(uint32_t));
        for (uint64_t i = 0; i < n; i++)
            for (uint64_t k = a->row_ptr[i]; k < a->row_ptr[i + 1]; k++) {
                uint32_t c = a->col_indices[k];
                trow[pos[c]] = (uint32_t)i;
                tval[pos[c]] = a->values[k];
                pos[c]++;
                if (c == i) diag[i] += a->values[k];
            }
        free(pos);
    }
    for (uint64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; }
    int conv = 0, rc = ORC_OK;
    uint64_t it = 0;
    for (uint64_t iter = 0; iter < max_iterations; iter++) {
        double max_res = 0.0;
        int64_t max_node = -1;
        for (uint64_t i = 0; i < n; i++)  /* first strict maximum of |residual| (:455-461) */
            if (fabs(r[i]) > max_res) { max_res = fabs(r[i]); max_node = (int64_t)i; }
        if (max_res < epsilon) { conv = 1; break; }
        if (fabs(diag[max_node]) < 1e-15) { rc = ORC_ERR_NUMERICAL_INSTABILITY; break; }
        double push = r[max_node] / diag[max_node];
        x[max_node] += push;
        r[max_node] = 0.0;
        for (uint64_t k = tptr[max_node]; k < tptr[max_node + 1]; k++)
            if (trow[k] != (uint64_t)max_node) r[trow[k]] -= tval[k] * push;
        it = iter + 1;
    }
    if (!conv && rc == ORC_OK) {
        /* the loop may also end exactly converged after the last allowed push: the reference only learns that at the top
         * of the next iteration, which does not exist -> CONVERGENCE_FAILED (:505-511) */
        rc = ORC_ERR_CONVERGENCE_FAILURE;
    }
    double nrm = 0.0;
    for (uint64_t i = 0; i < n; i++) nrm += r[i] * r[i];
    if (iterations) *iterations = it;
    if (residual_norm) *residual_norm = sqrt(nrm);
    if (converged) *converged = conv;
    free(r); free(diag); free(tptr); free(trow); free(tval);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Conjugate gradient (src/optimized_solver.rs:182-295; fast_solver.rs:126-178; ultra_fast.rs:116-158) */
/* ------------------------------------------------------------------------------------------ */
static double cg_dot(const double *x, const double *y, uint64_t n, int variant) {
    double sum = 0.0;
    if (variant == ORC_DOT_CHUNK4) { /* fast_solver.rs:180-200 */
        uint64_t chunks = n / 4;
        for (uint64_t c = 0; c < chunks; c++) {
            uint64_t i = c * 4;
            sum += x[i] * y[i] + x[i + 1] * y[i + 1] + x[i + 2] * y[i + 2] + x[i + 3] * y[i + 3];
        }
        for (uint64_t i = chunks * 4; i < n; i++) sum += x[i] * y[i];
    } else if (variant == ORC_DOT_CHUNK8) { /* ultra_fast.rs:161-185 */
        uint64_t chunks = n / 8;
        for (uint64_t c = 0; c < chunks; c++) {
            uint64_t i = c * 8;
            sum += x[i] * y[i] + x[i + 1] * y[i + 1] + x[i + 2] * y[i + 2] + x[i + 3] * y[i + 3] +
                   x[i + 4] * y[i + 4] + x[i + 5] * y[i + 5] + x[i + 6] * y[i + 6] + x[i + 7] * y[i + 7];
        }
        for (uint64_t i = chunks * 8; i < n; i++) sum += x[i] * y[i];
    } else { /* optimized_solver.rs:211-214 */
        for (uint64_t i = 0; i < n; i++) sum += x[i] * y[i];
    }
    return sum;
}

int orc_cg_solve(const orc_csr *m, const double *b, uint64_t blen, uint64_t max_iterations,
                 double tolerance, int spmv_variant, int dot_variant, int nthreads, orc_cg_result *res) {
    if (m->nrows != m->ncols) return ORC_ERR_INVALID_INPUT;      /* :188-190 */
    if (blen != m->nrows) return ORC_ERR_DIMENSION_MISMATCH;     /* :191-193 */
    uint64_t n = m->nrows;
    double *x = res->solution;
    double *r = (double *)malloc((n ? n : 1) * sizeof(double));
    double *p = (double *)malloc((n ? n : 1) * sizeof(double));
    double *ap = (double *)calloc(n ? n : 1, sizeof(double));
    if (!r || !p || !ap) { free(r); free(p); free(ap); return ORC_ERR_MEMORY_ALLOCATION; }
    for (uint64_t i = 0; i < n; i++) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; } /* :202-208, :215 */
    uint64_t iteration = 0, matvecs = 0;
    double tol_sq = tolerance * tolerance;
    int converged = 0;
    double rsold = cg_dot(r, r, n, dot_variant);
    while (iteration < max_iterations) {
        if (rsold <= tol_sq) { converged = 1; break; }            /* :218-221 */
        spmv_dispatch(m, p, ap, spmv_variant, nthreads);          /* :224 */
        matvecs++;
        double pap = cg_dot(p, ap, n, dot_variant);               /* :228-232 */
        if (fabs(pap) < 1e-16) break;                             /* :234-236 */
        double alpha = rsold / pap;
        for (uint64_t i = 0; i < n; i++) x[i] += alpha * p[i];    /* :241-243 */
        for (uint64_t i = 0; i < n; i++) r[i] -= alpha * ap[i];   /* :246-248 */
        double rsnew = cg_dot(r, r, n, dot_variant);
        double beta = rsnew / rsold;
        for (uint64_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i]; /* :258-260 */
        rsold = rsnew;
        iteration++;
    }
    res->residual_norm = sqrt(rsold);
    res->iterations = iteration;
    res->converged = converged;
    res->matvec_count = matvecs;
    res->total_flops = matvecs * m->nnz * 2 + iteration * n * 6;
    free(r); free(p); free(ap);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Generators                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* nnz_per_row = ((size as f64 * sparsity).max(3.0) as usize).min(size)  (benches/performance_benchmarks.rs:14) */
uint64_t orc_gen_bench_k(uint64_t size, double sparsity) {
    double k = fmax((double)size * sparsity, 3.0);
    uint64_t ku = (uint64_t)k;
    return ku < size ? ku : size;
}

#define LCG_A 1664525ULL
#define LCG_C 1013904223ULL
#define TWO64 18446744073709551616.0 /* `u64::MAX as f64` rounds to 2^64 */

/* One row of create_test_matrix (benches/performance_benchmarks.rs:17-35), in generation order. */
static uint64_t gen_bench_row(uint64_t size, uint64_t k, uint64_t i, uint32_t *cols, double *vals) {
    double diagonal_value = 10.0 + ((double)i * 0.01);
    uint64_t cnt = 0;
    cols[cnt] = (uint32_t)i;
    vals[cnt] = diagonal_value;
    cnt++;
    double max_off = diagonal_value / ((double)k * 2.0);
    uint64_t rng = i * LCG_A + LCG_C;
    for (uint64_t j = 1; j < k; j++) {
        rng = rng * LCG_A + LCG_C;
        uint64_t col = rng % size;
        if (col != i) {
            rng = rng * LCG_A + LCG_C;
            double value = ((double)rng / TWO64) * max_off;
            cols[cnt] = (uint32_t)col;
            vals[cnt] = value;
            cnt++;
        }
    }
    return cnt;
}

int64_t orc_gen_bench_triplets(uint64_t size, double sparsity, uint64_t *rows, uint64_t *cols,
                               double *vals, uint64_t cap) {
    uint64_t k = orc_gen_bench_k(size, sparsity);
    uint32_t *c = (uint32_t *)malloc((k + 1) * sizeof(uint32_t));
    double *v = (double *)malloc((k + 1) * sizeof(double));
    uint64_t total = 0;
    for (uint64_t i = 0; i < size; i++) {
        uint64_t cnt = gen_bench_row(size, k, i, c, v);
        for (uint64_t j = 0; j < cnt; j++) {
            if (total >= cap) { free(c); free(v); return -1; }
            rows[total] = i;
            cols[total] = c[j];
            vals[total] = v[j];
            total++;
        }
    }
    free(c);
    free(v);
    return (int64_t)total;
}

/* CSR of rows [row0,row1) = what from_triplets yields for those rows: zero values dropped
 * (sparse.rs:535-541), entries stably sorted by column (sparse.rs:95). row_ptr is local (starts at 0). */
int orc_gen_bench_csr(uint64_t size, double sparsity, uint64_t row0, uint64_t row1, orc_csr *out, double *b) {
    memset(out, 0, sizeof(*out));
    if (row1 > size || row0 > row1) return ORC_ERR_INVALID_INPUT;
    uint64_t k = orc_gen_bench_k(size, sparsity);
    uint64_t nloc = row1 - row0;
    uint64_t cap = nloc * k + 1;
    out->nrows = nloc;
    out->ncols = size;
    out->row_ptr = (uint32_t *)calloc(nloc + 1, sizeof(uint32_t));
    out->values = (double *)malloc(cap * sizeof(double));
    out->col_indices = (uint32_t *)malloc(cap * sizeof(uint32_t));
    uint32_t *c = (uint32_t *)malloc((k + 1) * sizeof(uint32_t));
    double *v = (double *)malloc((k + 1) * sizeof(double));
    if (!out->row_ptr || !out->values || !out->col_indices || !c || !v) {
        free(c); free(v); orc_csr_free(out);
        return ORC_ERR_MEMORY_ALLOCATION;
    }
    uint64_t nnz = 0;
    for (uint64_t i = row0; i < row1; i++) {
        uint64_t cnt = gen_bench_row(size, k, i, c, v);
        /* stable insertion sort by column */
        for (uint64_t a = 1; a < cnt; a++) {
            uint32_t cc = c[a];
            double vv = v[a];
            uint64_t p = a;
            while (p > 0 && c[p - 1] > cc) { c[p] = c[p - 1]; v[p] = v[p - 1]; p--; }
            c[p] = cc;
            v[p] = vv;
        }
        for (uint64_t a = 0; a < cnt; a++) {
            if (v[a] != 0.0) {
                out->col_indices[nnz] = c[a];
                out->values[nnz] = v[a];
                nnz++;
            }
        }
        out->row_ptr[i - row0 + 1] = (uint32_t)nnz;
        if (b) b[i - row0] = 1.0 + ((double)i * 0.001); /* create_test_rhs (:41-43) */
    }
    out->nnz = nnz;
    free(c);
    free(v);
    return ORC_OK;
}

/* generate_test_matrix (src/ultra_fast.rs:221-248): one LCG stream across all rows; the same
 * `rng_state` both picks the column and (un-advanced) supplies the value. */
int64_t orc_gen_ultra_triplets(uint64_t size, double sparsity, uint64_t *rows, uint64_t *cols,
                               double *vals, uint64_t cap) {
    uint64_t state = 12345ULL, total = 0;
    double kf = fmax((double)size * sparsity, 1.0);
    uint64_t k = (uint64_t)kf;
    if (k > 10) k = 10;
    for (uint64_t i = 0; i < size; i++) {
        if (total >= cap) return -1;
        rows[total] = i; cols[total] = i; vals[total] = 10.0 + (double)i * 0.01;
        total++;
        for (uint64_t d = 0; d < k; d++) {
            state = state * 1103515245ULL + 12345ULL;
            uint64_t j = state % size;
            if (i != j) {
                if (total >= cap) return -1;
                rows[total] = i; cols[total] = j;
                vals[total] = ((double)state / TWO64) * 0.1;
                total++;
            }
        }
    }
    return (int64_t)total;
}

/* ------------------------------------------------------------------------------------------ */
/* PageRank system                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* computePageRank (src/core/solver.ts:664-722): outdeg[i] = sum_j adj[i][j] (:679-684);
 * S[i][i] = 1; S[i][j] -= alpha * adj[j][i] / outdeg[j] when outdeg[j] > 0 (:689-698), so dangling
 * nodes contribute nothing; rhs = (1-alpha)/n (:708).  Edge (src -> dst, w) is adj[src][dst] = w;
 * repeated edges accumulate (a dense adjacency holds one number per pair). */
int orc_pagerank_system(const uint64_t *src, const uint64_t *dst, const double *w, uint64_t nedges,
                        uint64_t n, double alpha, orc_csr *S, double *rhs) {
    double *outdeg = (double *)calloc(n ? n : 1, sizeof(double));
    uint64_t nt = nedges + n;
    uint64_t *tr = (uint64_t *)malloc(nt * sizeof(uint64_t));
    uint64_t *tc = (uint64_t *)malloc(nt * sizeof(uint64_t));
    double *tv = (double *)malloc(nt * sizeof(double));
    if (!outdeg || !tr || !tc || !tv) { free(outdeg); free(tr); free(tc); free(tv); return ORC_ERR_MEMORY_ALLOCATION; }
    for (uint64_t e = 0; e < nedges; e++) {
        if (src[e] >= n || dst[e] >= n) { free(outdeg); free(tr); free(tc); free(tv); return ORC_ERR_INDEX_OUT_OF_BOUNDS; }
        outdeg[src[e]] += w ? w[e] : 1.0;
    }
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++) { tr[k] = i; tc[k] = i; tv[k] = 1.0; k++; }
    for (uint64_t e = 0; e < nedges; e++) {
        uint64_t j = src[e], i = dst[e];
        if (outdeg[j] > 0.0) {
            tr[k] = i; tc[k] = j;
            tv[k] = -(alpha * ((w ? w[e] : 1.0) / outdeg[j]));
            k++;
        }
    }
    orc_csr raw;
    int rc = orc_csr_from_triplets(tr, tc, tv, k, n, n, &raw);
    free(tr); free(tc); free(tv); free(outdeg);
    if (rc != ORC_OK) return rc;
    /* a dense matrix has one value per (i,j): merge duplicates by summation, in CSR order */
    memset(S, 0, sizeof(*S));
    S->nrows = n; S->ncols = n;
    S->row_ptr = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
    S->values = (double *)malloc((raw.nnz ? raw.nnz : 1) * sizeof(double));
    S->col_indices = (uint32_t *)malloc((raw.nnz ? raw.nnz : 1) * sizeof(uint32_t));
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t p = raw.row_ptr[i], e = raw.row_ptr[i + 1];
        while (p < e) {
            uint32_t c = raw.col_indices[p];
            double acc = raw.values[p];
            p++;
            while (p < e && raw.col_indices[p] == c) { acc += raw.values[p]; p++; }
            if (acc != 0.0) { S->col_indices[nnz] = c; S->values[nnz] = acc; nnz++; }
        }
        S->row_ptr[i + 1] = (uint32_t)nnz;
    }
    S->nnz = nnz;
    orc_csr_free(&raw);
    if (rhs) for (uint64_t i = 0; i < n; i++) rhs[i] = (1.0 - alpha) / (double)n;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Single-entry estimation                                                                    */
/* ------------------------------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* src/core/utils.ts:161-168 */
uint32_t orc_ts_lcg_next(uint32_t state, double *u) {
    uint64_t s = ((uint64_t)state * 1664525ULL + 1013904223ULL) % 0x100000000ULL;
    if (u) *u = (double)s / 4294967296.0;
    return (uint32_t)s;
}

/* Ulam-von Neumann absorbing walk.  M_sj = -(a_sj * dinv_s), j != s;
 * one uniform u per step: scan row s in CSR order accumulating |M_sj|; move to the first j whose
 * running sum exceeds u, multiplying the weight by sign(M_sj); if none does, the walk is absorbed.
 * Each visited state adds W * c_s, c = D^-1 b.  Walk w of query q draws
 * u_d = (splitmix64(key + d) >> 11) * 2^-53, key = splitmix64(splitmix64(seed ^ (q+1)*0xA0761D6478BD642F) + w).
 * numSamples / maxSteps defaults come from src/core/solver.ts:587 and :399. */
int orc_solve_entry(const orc_csr *m, const double *b, const uint64_t *rows, uint64_t nq,
                    uint64_t nwalks, uint64_t max_steps, uint64_t seed, double *est, double *var) {
    uint64_t n = m->nrows;
    if (m->nrows != m->ncols) return ORC_ERR_INVALID_INPUT;
    double *dinv = (double *)malloc((n ? n : 1) * sizeof(double));
    double *c = (double *)malloc((n ? n : 1) * sizeof(double));
    for (uint64_t i = 0; i < n; i++) {
        double d = 0.0;
        int present = 0;
        for (uint64_t k = m->row_ptr[i]; k < m->row_ptr[i + 1]; k++)
            if ((uint64_t)m->col_indices[k] == i) { d += m->values[k]; present = 1; }
        if (!present || fabs(d) < 1e-14) { free(dinv); free(c); return ORC_ERR_INVALID_SPARSE_MATRIX; }
        dinv[i] = 1.0 / d;
        c[i] = b[i] * dinv[i];
    }
    for (uint64_t q = 0; q < nq; q++) {
        if (rows[q] >= n) { free(dinv); free(c); return ORC_ERR_INDEX_OUT_OF_BOUNDS; }
        uint64_t qkey = splitmix64(seed ^ ((q + 1) * 0xA0761D6478BD642FULL));
        double sum = 0.0, sumsq = 0.0;
        for (uint64_t w = 0; w < nwalks; w++) {
            uint64_t key = splitmix64(qkey + w);
            uint64_t s = rows[q];
            double W = 1.0, acc = 0.0;
            for (uint64_t step = 0; step < max_steps; step++) {
                acc += W * c[s];
                double u = (double)(splitmix64(key + step) >> 11) * (1.0 / 9007199254740992.0);
                double cum = 0.0, ds = dinv[s];
                int moved = 0;
                for (uint64_t k = m->row_ptr[s]; k < m->row_ptr[s + 1]; k++) {
                    uint32_t j = m->col_indices[k];
                    if ((uint64_t)j == s) continue;
                    double mv = -(m->values[k] * ds);
                    cum += fabs(mv);
                    if (cum > u) {
                        if (mv < 0.0) W = -W;
                        s = j;
                        moved = 1;
                        break;
                    }
                }
                if (!moved) break;
            }
            sum += acc;
            sumsq += acc * acc;
        }
        double mean = sum / (double)nwalks;
        est[q] = mean;
        if (var) var[q] = nwalks > 1 ? (sumsq - (double)nwalks * mean * mean) / (double)(nwalks - 1) : 0.0;
    }
    free(dinv);
    free(c);
    return ORC_OK;
}
