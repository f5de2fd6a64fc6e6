// solver.cu — NeumannSolver / SolverOptions / SolverResult over the tile kernels.
//
// Host control flow of NeumannSolver::solve (ref src/solver/neumann.rs:469-555).  The per-iteration decisions
// (series_converged, residual <= tolerance, NumericalInstability, max_iterations) are taken ON THE DEVICE by the
// last CTA of each kernel (LoopCtl, device_util.cuh); the host only enqueues iterations in batches and reads the
// loop state back once per batch, so a term costs one kernel launch and no host synchronisation.
#include <chrono>
#include <cmath>
#include <cstring>

#include "matrix.hpp"
#include "solver.hpp"

using namespace sb200;

namespace sb200 {

static double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}


int32_t validate_options(const sb200_options *opt) {
    if (!opt) return fail(SB200_ERR_INVALID_INPUT, "options is null");
    if (opt->mode != SB200_MODE_CORRECT && opt->mode != SB200_MODE_REF_COMPAT)
        return fail(SB200_ERR_INVALID_INPUT, "unknown mode %d", opt->mode);
    if (opt->residual_check == SB200_RESIDUAL_IDENTITY && opt->mode != SB200_MODE_CORRECT)
        return fail(SB200_ERR_INVALID_INPUT,
                    "SB200_RESIDUAL_IDENTITY needs SB200_MODE_CORRECT (the identity b - A x_k = D t_{k+1} does not hold "
                    "for the reference's double-counted iterate)");
    return SB200_OK;
}

// The solve on device-resident vectors. b_dev: n doubles; x0_dev: initial guess or null; x_out_dev: n doubles.
int32_t solve_device(const sb200_solver *s, sb200_matrix *m, const double *b_dev, const double *x0_dev,
                     const sb200_options *opt, double *x_out_dev, cudaStream_t st, Workspace &ws, SolveStats &stats,
                     const StreamHook *hook) {
    const uint64_t n = m->nrows;
    const int cfg = m->tile_cfg;
    const bool compat = opt->mode == SB200_MODE_REF_COMPAT;
    const bool identity = opt->residual_check == SB200_RESIDUAL_IDENTITY;
    const uint64_t max_it = opt->max_iterations, max_terms = s->max_terms;
    if (max_it >= 0xFFFFFFFFull || max_terms >= 0xFFFFFFFFull)
        return fail(SB200_ERR_INVALID_INPUT, "max_iterations / max_terms must fit 32 bits");

    const size_t npart = 2 * (size_t)std::max(std::max(tile_kernel_max_grid(cfg, EPI_PUSH), tile_kernel_max_grid(cfg, EPI_RESID)),
                                              init_state_grid()) + 2;
    SB_TRY(ws.ensure(n, n, npart));

    LoopCtl h{};
    h.res_norm = INFINITY;  // neumann.rs:236
    h.tolerance = opt->tolerance;
    h.series_tolerance = s->series_tolerance;
    h.max_terms = (uint32_t)max_terms;
    h.max_iterations = (uint32_t)max_it;
    h.alive = 1;
    *ws.h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws.ctl.p, ws.h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));

    uint64_t launches = 0, extra_matvec = 0;
    const uint64_t kpl = launches_per_pass(m);
    const double *dinv = m->d_dinv[opt->mode].p;
    const double *resid_rhs = compat ? ws.c.p : b_dev;  // update_residual subtracts D^-1 b in the reference (:308-314)

    // correct mode with an initial guess: t0 = D^-1 (b - A x0), one extra SpMV
    const double *ax0 = nullptr;
    if (!compat && x0_dev) {
        SB_TRY(matrix_spmv_dev(m, x0_dev, ws.tmp.p, 0, st));
        launches += kpl;
        extra_matvec++;
        ax0 = ws.tmp.p;
    }

    TileKernelArgs base{};
    fill_tile_args(m, base);
    base.ctl = ws.ctl.p;
    base.partials = ws.partials.p;
    base.identity_res = identity;
    base.acc = ws.tmp.p;  // column-slab passes: partial row sums (A x0, the only other use of tmp, is consumed by the init kernel)

    // optional per-launch events (options.enable_profiling): which launches did work is known after the loop
    struct ProfEv {
        cudaEvent_t e0, e1;
        int kind;  // 1 push, 2 residual
        uint64_t it;
        int force;
    };
    std::vector<ProfEv> prof;
    const bool profiling = opt->enable_profiling != 0;
    struct ProfCleanup {
        std::vector<ProfEv> &v;
        ~ProfCleanup() {
            for (auto &p : v) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
        }
    } prof_cleanup{prof};
    auto prof_begin = [&](int kind, uint64_t it, int force) -> int32_t {
        if (!profiling) return SB200_OK;
        ProfEv p{nullptr, nullptr, kind, it, force};
        SB_CUDA(cudaEventCreate(&p.e0));
        SB_CUDA(cudaEventCreate(&p.e1));
        SB_CUDA(cudaEventRecord(p.e0, st));
        prof.push_back(p);
        return SB200_OK;
    };
    auto prof_end = [&]() -> int32_t {
        if (!profiling) return SB200_OK;
        SB_CUDA(cudaEventRecord(prof.back().e1, st));
        return SB200_OK;
    };

    auto enqueue_resid = [&](uint64_t it, int last, int force) -> int32_t {
        TileKernelArgs a = base;
        a.xin = x_out_dev;
        a.xin_own = x_out_dev;
        a.rhs = resid_rhs;
        a.it = (uint32_t)it;
        a.last_in_iter = last;
        a.force = force;
        a.identity_res = 0;
        launches += kpl;
        SB_TRY(prof_begin(2, it, force));
        SB_TRY(launch_tile_kernel(cfg, EPI_RESID, a, st));
        return prof_end();
    };
    auto read_ctl = [&]() -> int32_t {
        SB_CUDA(cudaMemcpyAsync(ws.h_ctl, ws.ctl.p, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return SB200_OK;
    };

    SB_CUDA(cudaEventRecord(ws.ev0, st));
    // ---- iteration 0 state + first term (NeumannState::new :191-211, compute_next_term with terms == 0) ----
    const bool loop_runs = max_it > 0;
    {
        InitArgs ia{};
        ia.b = b_dev;
        ia.dinv = dinv;
        ia.x0 = x0_dev;
        ia.ax0 = ax0;
        ia.c_out = ws.c.p;
        ia.t_out = ws.t[0].p;
        ia.x_out = x_out_dev;
        ia.n = (uint32_t)n;
        ia.compat = compat;
        ia.ctl = ws.ctl.p;
        ia.partials = ws.partials.p;
        ia.identity_res = identity;
        ia.skip_term0 = !(loop_runs && max_terms > 0);
        if (ia.skip_term0) {
            // the loop body never adds a term: only materialise c / t / x = base. The kernel's tail would count a
            // term, so run it against a scratch control block (h_ctl is re-uploaded right after).
            ia.last_in_iter = 0;
            SB_TRY(launch_init_state(ia, st));
            launches++;
            SB_CUDA(cudaMemcpyAsync(ws.ctl.p, ws.h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));
        } else {
            const bool resid_due = !identity;  // iteration 0: 0 % 5 == 0 (:489-491)
            ia.last_in_iter = !resid_due;
            SB_TRY(launch_init_state(ia, st));
            launches++;
            if (resid_due) SB_TRY(enqueue_resid(0, 1, 0));
        }
    }

    // ---- iterations 1 .. : one fused push kernel per term, residual kernel every 5th iteration ----
    uint64_t it = 1;
    const uint64_t push_end = std::min(max_it, max_terms);  // iterations [1, push_end) compute a term
    const bool streaming = hook && hook->fn && hook->interval > 0;
    const uint64_t kBatch = streaming ? hook->interval : 8;  // the loop state is read back once per batch
    bool stopped_by_callback = false;
    // PartialSolution (solver/mod.rs:198-217) after the batch that ended at iteration `upto`
    auto emit_partial = [&](uint64_t upto) -> int32_t {
        const LoopCtl &c = *ws.h_ctl;
        SB_CUDA(cudaMemcpyAsync(hook->host_x, x_out_dev, n * 8, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        sb200_partial_solution p{};
        p.iteration = std::min<uint64_t>(c.iterations, upto);
        p.solution = hook->host_x;
        p.solution_len = n;
        p.residual_norm = c.res_norm;
        p.converged = (c.res_norm <= opt->tolerance) || (c.sconv && c.terms < max_terms);
        const double tn = std::sqrt(c.term_norm2), t0n = std::sqrt(c.rhs_norm2);
        if (c.terms > 1 && tn > 0.0 && t0n > 0.0 && tn < t0n) {
            const double rho = std::pow(tn / t0n, 1.0 / (double)(c.terms - 1));  // observed contraction per term
            if (rho < 1.0 && rho > 0.0) {
                const double need = tn < s->series_tolerance ? 0.0 : std::ceil(std::log(s->series_tolerance / tn) / std::log(rho));
                p.has_estimated_remaining = 1;
                p.estimated_remaining = (uint64_t)std::max(0.0, need);
            }
        }
        p.timestamp_ms = wall_ms() - hook->t0_ms;
        if (hook->fn(&p, hook->user) != 0) stopped_by_callback = true;
        return SB200_OK;
    };
    bool alive = loop_runs && max_terms > 0;
    if (alive) {
        // small systems finish within the first batch; read the state only after real work was queued
        while (alive && it < push_end) {
            const uint64_t end = std::min(push_end, it + kBatch);
            for (; it < end; it++) {
                const bool resid_due = !identity && (it % 5 == 0);
                TileKernelArgs a = base;
                a.xin = ws.t[(it - 1) & 1].p;
                a.xin_own = a.xin;
                a.out = ws.t[it & 1].p;
                a.sol = x_out_dev;
                a.dinv = dinv;
                a.it = (uint32_t)it;
                a.last_in_iter = !resid_due;
                SB_TRY(prof_begin(1, it, 0));
                SB_TRY(launch_tile_kernel(cfg, EPI_PUSH, a, st));
                SB_TRY(prof_end());
                launches += kpl;
                if (resid_due) SB_TRY(enqueue_resid(it, 1, 0));
            }
            SB_TRY(read_ctl());
            alive = ws.h_ctl->alive != 0;
            if (streaming) {
                SB_TRY(emit_partial(it));
                if (stopped_by_callback && alive) {  // freeze the loop where the caller stopped it
                    ws.h_ctl->alive = 0;
                    SB_CUDA(cudaMemcpyAsync(ws.ctl.p, ws.h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));
                    alive = false;
                }
            }
        }
        if (it <= 1) {  // push_end <= 1: nothing was read back yet
            SB_TRY(read_ctl());
            alive = ws.h_ctl->alive != 0;
        }
    }

    uint64_t iterations = ws.h_ctl->iterations;
    uint64_t resid_in_loop = 0;
    // ---- "spin" phase: terms == max_terms, series not converged. compute_next_term is
    // a no-op (:253-255), the solution no longer changes, the loop keeps re-evaluating the residual every 5th
    // iteration until it is <= tolerance or max_iterations is hit. One evaluation tells all of them. ----
    if (!stopped_by_callback && loop_runs && (max_terms == 0 || (alive && iterations >= max_terms)) && iterations < max_it) {
        uint64_t cur = iterations;
        if (identity) {
            // identity mode evaluates nothing further: the residual estimate is frozen
            iterations = (ws.h_ctl->res_norm <= opt->tolerance) ? cur : max_it;
        } else {
            const uint64_t first = (cur + 4) / 5 * 5;  // first iteration index >= cur with it % 5 == 0
            if (first < max_it) {
                SB_TRY(enqueue_resid(first, 0, 1));
                SB_TRY(read_ctl());
                resid_in_loop++;
                const double r = ws.h_ctl->res_norm;
                if (!std::isfinite(r)) {
                    ws.h_ctl->nonfinite = 1;
                    iterations = first + 1;
                } else if (r <= opt->tolerance) {
                    iterations = first + 1;  // loop condition fails at the top of the next iteration
                } else {
                    for (uint64_t k = first + 5; k < max_it; k += 5) resid_in_loop++;  // identical re-evaluations
                    iterations = max_it;
                }
            } else {
                iterations = max_it;
            }
        }
    }

    // ---- final residual (:516) ----
    SB_TRY(enqueue_resid(iterations, 0, 1));
    SB_CUDA(cudaEventRecord(ws.ev1, st));
    SB_TRY(read_ctl());
    float dev_ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&dev_ms, ws.ev0, ws.ev1));

    const LoopCtl &c = *ws.h_ctl;
    for (auto &p : prof) {  // launches past the end of the loop were no-ops: leave them out of the averages
        const bool live = p.force || (p.kind == 1 ? p.it < c.terms : p.it < c.iterations);
        if (!live) continue;
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, p.e0, p.e1));
        if (p.kind == 1) { stats.push_ms += ms; stats.push_count++; }
        else { stats.resid_ms += ms; stats.resid_count++; }
    }
    stats.iterations = iterations;
    stats.terms = c.terms;
    stats.series_converged = c.sconv != 0;
    stats.residual_norm = c.res_norm;
    stats.last_term_norm = std::sqrt(c.term_norm2);
    stats.rhs_norm = std::sqrt(c.rhs_norm2);
    stats.nonfinite = c.nonfinite != 0;
    stats.device_ms = dev_ms;
    stats.launches = launches;
    // matvec_count (:286, :305): one per term after the first, one per in-loop residual, one final (+ A x0)
    uint64_t loop_resids = 0;
    if (!identity) {
        const uint64_t counted = std::min<uint64_t>(c.iterations, iterations);
        loop_resids = counted ? (counted - 1) / 5 + 1 : 0;  // it = 0,5,10,.. < counted
        if (max_terms == 0 || !loop_runs) loop_resids = 0;   // nothing was enqueued by the normal phase
    }
    stats.matvec = (c.terms > 0 ? c.terms - 1 : 0) + loop_resids + resid_in_loop + 1 + extra_matvec;
    // is_converged (:422-430) after the final residual (:518)
    stats.converged = (c.res_norm <= opt->tolerance) || (stats.series_converged && c.terms < max_terms);
    return SB200_OK;
}

}  // namespace sb200

// -------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------
static int32_t new_solver(uint64_t max_terms, double tol, int adaptive, int cache, sb200_solver **out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    sb200_solver *s = new sb200_solver();
    s->max_terms = max_terms;
    s->series_tolerance = tol;
    s->adaptive_truncation = adaptive;
    s->cache_powers = cache;
    *out = s;
    return SB200_OK;
}

// shared front half of the solve entry points: argument checks in the reference's order (NeumannState::new :147-206)
int32_t sb200::solve_precheck(const sb200_solver *s, const sb200_matrix *m, uint64_t blen, const sb200_options *opt) {
    if (!s || !m) return fail(SB200_ERR_INVALID_INPUT, "null solver or matrix");
    SB_TRY(validate_options(opt));
    if (m->distributed) return fail(SB200_ERR_INVALID_INPUT, "row-block matrix: use sb200_dist_solve");
    if (m->nrows != m->ncols)
        return fail(SB200_ERR_INVALID_INPUT, "Matrix must be square for Neumann series (%llu x %llu)",
                    (unsigned long long)m->nrows, (unsigned long long)m->ncols);
    if (blen != m->nrows)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in neumann_initialization",
                    (unsigned long long)m->nrows, (unsigned long long)blen);
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    const bool cols = opt->dominance == SB200_DOMINANCE_ROW_OR_COL;
    SB_TRY(matrix_analyse(mm, opt->mode, cols));
    const bool dd = m->first_bad_dd == kNone || (cols && m->first_bad_col == kNone);
    if (!dd)
        return fail(SB200_ERR_MATRIX_NOT_DIAGONALLY_DOMINANT, "matrix is not diagonally dominant (first violating row %llu)",
                    (unsigned long long)m->first_bad_dd);
    if (m->first_bad_diag[opt->mode] != kNone)
        return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "Missing or near-zero diagonal element at position %llu",
                    (unsigned long long)m->first_bad_diag[opt->mode]);
    if (opt->initial_guess && opt->initial_guess_len != m->nrows)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in initial_guess",
                    (unsigned long long)m->nrows, (unsigned long long)opt->initial_guess_len);
    return SB200_OK;
}

static void fill_result(const sb200_solver *s, const sb200_options *opt, const SolveStats &st, const Workspace &ws,
                        double total_ms, sb200_result *out) {
    out->residual_norm = st.residual_norm;
    out->iterations = st.iterations;
    out->converged = st.converged;
    out->terms_computed = st.terms;
    out->series_converged = st.series_converged;
    out->last_term_norm = st.last_term_norm;
    out->device_time_ms = st.device_ms;
    out->kernel_launches = st.launches;
    out->memory_bytes = ws.bytes();
    out->matvec_count = st.matvec;
    out->total_time_ms = total_ms;
    out->has_stats = opt->collect_stats != 0;
    out->push_kernel_ms = st.push_ms;
    out->push_kernel_count = st.push_count;
    out->resid_kernel_ms = st.resid_ms;
    out->resid_kernel_count = st.resid_count;
    out->has_error_bounds = 0;
    out->error_upper_bound = 0.0;
    // estimate_error_bounds (:321-347), evaluated on the final state (it only fires once series_converged)
    if (opt->compute_error_bounds && s->adaptive_truncation && st.series_converged && st.terms > 0) {
        double est = 0.0;
        if (st.terms > 1) est = std::pow(st.last_term_norm / st.rhs_norm, 1.0 / (double)(st.terms - 1));
        if (est < 1.0) {
            out->has_error_bounds = 1;
            out->error_upper_bound = std::pow(est, (double)(int)st.terms) / (1.0 - est) * st.rhs_norm;
        }
    }
}

static int32_t classify(const sb200_options *opt, const SolveStats &st) {
    if (st.nonfinite)  // :501-507
        return fail(SB200_ERR_NUMERICAL_INSTABILITY, "Non-finite residual norm at iteration %llu",
                    (unsigned long long)st.iterations);
    if (!st.converged && st.iterations >= opt->max_iterations)  // :523-530
        return fail(SB200_ERR_CONVERGENCE_FAILURE,
                    "neumann failed to converge: %llu iterations, residual %.6e, tolerance %.6e",
                    (unsigned long long)st.iterations, st.residual_norm, opt->tolerance);
    return SB200_OK;
}

static int32_t solve_host(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                          const sb200_options *opt, double *x_out, bool own_solution, sb200_result *out,
                          StreamHook *hook = nullptr) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "result is null");
    memset(out, 0, sizeof(*out));
    const double t0 = wall_ms();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    DeviceGuard g(m->device);
    SB_TRY(solve_precheck(s, m, blen, opt));
    if (blen && !b) return fail(SB200_ERR_INVALID_INPUT, "b is null");
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    const uint64_t n = m->nrows;
    auto ws = matrix_acquire_ws(mm);
    struct Release {
        sb200_matrix *m;
        std::unique_ptr<Workspace> &ws;
        ~Release() { matrix_release_ws(m, std::move(ws)); }
    } rel{mm, ws};
    SB_TRY(ws->ensure(n, n, 4));
    cudaStream_t st = m->stream;
    SB_TRY(copy_h2d(ws->b.p, b, n * 8, st));
    uint64_t h2d = n * 8;
    DevBuf<double> x0;
    if (opt->initial_guess) {
        SB_TRY(x0.alloc(n));
        SB_TRY(copy_h2d(x0.p, opt->initial_guess, n * 8, st));
        h2d += n * 8;
    }
    SolveStats stats{};
    struct HookBuf {
        double *p = nullptr;
        ~HookBuf() { if (p) pinned_pool_put(p); }
    } hook_buf;
    if (hook) {
        hook_buf.p = (double *)pinned_pool_get(n * 8);
        if (!hook_buf.p) return fail(SB200_ERR_MEMORY_ALLOCATION, "pinned allocation of %llu bytes failed", (unsigned long long)(n * 8));
        hook->host_x = hook_buf.p;
        hook->t0_ms = t0;
    }
    SB_TRY(solve_device(s, mm, ws->b.p, x0.p, opt, ws->x.p, st, *ws, stats, hook));
    double *dst = x_out;
    if (own_solution) {
        dst = (double *)pinned_pool_get(n * 8);
        if (!dst) return fail(SB200_ERR_MEMORY_ALLOCATION, "pinned allocation of %llu bytes failed", (unsigned long long)(n * 8));
        out->solution = dst;
        out->solution_len = n;
    }
    if (dst) SB_TRY(copy_d2h(dst, ws->x.p, n * 8, st));
    SB_CUDA(cudaStreamSynchronize(st));
    fill_result(s, opt, stats, *ws, wall_ms() - t0, out);
    out->h2d_bytes = h2d;
    out->d2h_bytes = dst ? n * 8 : 0;
    return classify(opt, stats);
}

extern "C" {

int32_t sb200_neumann_new(uint64_t max_terms, double series_tolerance, sb200_solver **out) {
    return new_solver(max_terms, series_tolerance, 1, 1, out);  // neumann.rs:48-55
}
int32_t sb200_neumann_default(sb200_solver **out) { return new_solver(50, 1e-8, 1, 1, out); }          // :58-60
int32_t sb200_neumann_high_precision(sb200_solver **out) { return new_solver(100, 1e-12, 1, 1, out); }  // :63-70
int32_t sb200_neumann_fast(sb200_solver **out) { return new_solver(20, 1e-6, 0, 0, out); }              // :73-80

int32_t sb200_neumann_with_adaptive_truncation(sb200_solver *s, int32_t enable) {
    if (!s) return fail(SB200_ERR_INVALID_INPUT, "null solver");
    s->adaptive_truncation = enable != 0;
    return SB200_OK;
}
int32_t sb200_neumann_with_power_caching(sb200_solver *s, int32_t enable) {
    if (!s) return fail(SB200_ERR_INVALID_INPUT, "null solver");
    s->cache_powers = enable != 0;  // carried only: the reference never fills matrix_powers either (neumann.rs:213-217)
    return SB200_OK;
}
int32_t sb200_neumann_config(const sb200_solver *s, uint64_t *max_terms, double *series_tolerance,
                             int32_t *adaptive_truncation, int32_t *cache_powers) {
    if (!s) return fail(SB200_ERR_INVALID_INPUT, "null solver");
    if (max_terms) *max_terms = s->max_terms;
    if (series_tolerance) *series_tolerance = s->series_tolerance;
    if (adaptive_truncation) *adaptive_truncation = s->adaptive_truncation;
    if (cache_powers) *cache_powers = s->cache_powers;
    return SB200_OK;
}
const char *sb200_solver_algorithm_name(const sb200_solver *) { return "neumann"; }
void sb200_solver_free(sb200_solver *s) { delete s; }

// SolverOptions presets (src/solver/mod.rs:47-116)
void sb200_options_default(sb200_options *o) {
    memset(o, 0, sizeof(*o));
    o->tolerance = 1e-6;
    o->max_iterations = 1000;
    o->convergence_mode = SB200_CONV_RESIDUAL_NORM;
    o->norm_type = SB200_NORM_L2;
    o->error_bounds_tolerance = 1e-8;
    o->mode = SB200_MODE_CORRECT;
    o->dominance = SB200_DOMINANCE_ROW;
    o->residual_check = SB200_RESIDUAL_EVERY_5;
}
void sb200_options_high_precision(sb200_options *o) {
    sb200_options_default(o);
    o->tolerance = 1e-12;
    o->max_iterations = 5000;
    o->convergence_mode = SB200_CONV_COMBINED;
    o->collect_stats = 1;
    o->compute_error_bounds = 1;
    o->error_bounds_tolerance = 1e-14;
}
void sb200_options_fast(sb200_options *o) {
    sb200_options_default(o);
    o->tolerance = 1e-3;
    o->max_iterations = 100;
    o->error_bounds_tolerance = 1e-4;
}
void sb200_options_streaming(sb200_options *o, uint64_t interval) {
    sb200_options_default(o);
    o->tolerance = 1e-4;
    o->collect_stats = 1;
    o->streaming_interval = interval;
    o->error_bounds_tolerance = 1e-6;
    o->enable_profiling = 1;
}

int32_t sb200_solve(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                    const sb200_options *opt, sb200_result *out) {
    return solve_host(s, m, b, blen, opt, nullptr, true, out);
}

int32_t sb200_solve_into(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                         const sb200_options *opt, double *x_out, sb200_result *out) {
    if (blen && !x_out) return fail(SB200_ERR_INVALID_INPUT, "x_out is null");
    return solve_host(s, m, b, blen, opt, x_out, false, out);
}

int32_t sb200_solve_dev(const sb200_solver *s, const sb200_matrix *m, const double *b_dev, uint64_t blen,
                        const sb200_options *opt, double *x_dev, void *stream, sb200_result *out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "result is null");
    memset(out, 0, sizeof(*out));
    const double t0 = wall_ms();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    DeviceGuard g(m->device);
    SB_TRY(solve_precheck(s, m, blen, opt));
    if (blen && (!b_dev || !x_dev)) return fail(SB200_ERR_INVALID_INPUT, "null device vector");
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    auto ws = matrix_acquire_ws(mm);
    struct Release {
        sb200_matrix *m;
        std::unique_ptr<Workspace> &ws;
        ~Release() { matrix_release_ws(m, std::move(ws)); }
    } rel{mm, ws};
    SolveStats stats{};
    SB_TRY(solve_device(s, mm, b_dev, opt->initial_guess, opt, x_dev, (cudaStream_t)stream, *ws, stats));
    fill_result(s, opt, stats, *ws, wall_ms() - t0, out);
    return classify(opt, stats);
}

int32_t sb200_solve_streaming(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                              const sb200_options *opt, sb200_stream_callback callback, void *user, sb200_result *out) {
    if (!opt || !callback || opt->streaming_interval == 0) return solve_host(s, m, b, blen, opt, nullptr, true, out);
    StreamHook hook;
    hook.interval = opt->streaming_interval;
    hook.fn = callback;
    hook.user = user;
    return solve_host(s, m, b, blen, opt, nullptr, true, out, &hook);
}

void sb200_result_free(sb200_result *r) {
    if (!r) return;
    if (r->solution) pinned_pool_put(r->solution);
    r->solution = nullptr;
    r->solution_len = 0;
}

// The bare recurrence: term 0 then `nterms` fused push launches, no convergence logic.
int32_t sb200_push_iterations_dev(const sb200_matrix *m, const double *b_dev, uint64_t blen, uint64_t nterms,
                                  double *x_dev, double *t_dev, double *term_norms, void *stream, float *elapsed_ms) {
    clear_error();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    if (m->distributed) return fail(SB200_ERR_INVALID_INPUT, "row-block matrix: use sb200_dist_push_iterations_dev");
    if (m->nrows != m->ncols) return fail(SB200_ERR_INVALID_INPUT, "matrix must be square");
    if (blen != m->nrows) return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu", (unsigned long long)m->nrows, (unsigned long long)blen);
    DeviceGuard g(m->device);
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    SB_TRY(matrix_analyse(mm, SB200_MODE_CORRECT, false));
    if (m->first_bad_diag[0] != kNone)
        return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "Missing or near-zero diagonal element at position %llu",
                    (unsigned long long)m->first_bad_diag[0]);
    const uint64_t n = m->nrows;
    cudaStream_t st = (cudaStream_t)stream;
    auto ws = matrix_acquire_ws(mm);
    struct Release {
        sb200_matrix *m;
        std::unique_ptr<Workspace> &ws;
        ~Release() { matrix_release_ws(m, std::move(ws)); }
    } rel{mm, ws};
    const int cfg = m->tile_cfg;
    const size_t npart = 2 * (size_t)std::max(tile_kernel_max_grid(cfg, EPI_PUSH), init_state_grid()) + 2;
    SB_TRY(ws->ensure(n, n, npart));
    if (ws->norm_log.n < nterms + 1) SB_TRY(ws->norm_log.alloc(nterms + 1));
    LoopCtl h{};
    h.res_norm = INFINITY;
    h.alive = 1;
    h.max_terms = 0xFFFFFFFFu;
    h.max_iterations = 0xFFFFFFFFu;
    *ws->h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws->ctl.p, ws->h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));
    double *x = x_dev ? x_dev : ws->x.p;
    InitArgs ia{};
    ia.b = b_dev;
    ia.dinv = m->d_dinv[0].p;
    ia.c_out = nullptr;
    ia.t_out = ws->t[0].p;
    ia.x_out = x;
    ia.n = (uint32_t)n;
    ia.ctl = ws->ctl.p;
    ia.partials = ws->partials.p;
    ia.norm_log = ws->norm_log.p;
    SB_TRY(launch_init_state(ia, st));
    TileKernelArgs base{};
    fill_tile_args(m, base);
    base.ctl = ws->ctl.p;
    base.partials = ws->partials.p;
    base.acc = ws->tmp.p;
    base.force = 1;
    base.norm_log = ws->norm_log.p;
    base.sol = x;
    base.dinv = m->d_dinv[0].p;
    // debug aid: $SUBLINEAR_B200_PHASE_LOG=<thread index> prints the per-phase cycle breakdown of the push kernel
    DevBuf<unsigned long long> plog;
    if (const char *e = getenv("SUBLINEAR_B200_PHASE_LOG")) {
        SB_TRY(plog.alloc(64));
        unsigned long long init[64] = {0};
        init[63] = (unsigned long long)atoll(e);
        SB_CUDA(cudaMemcpyAsync(plog.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        base.phase_log = plog.p;
    }
    SB_CUDA(cudaEventRecord(ws->ev0, st));
    for (uint64_t it = 1; it <= nterms; it++) {
        TileKernelArgs a = base;
        a.xin = ws->t[(it - 1) & 1].p;
        a.xin_own = a.xin;
        a.out = ws->t[it & 1].p;
        a.it = (uint32_t)it;
        SB_TRY(launch_tile_kernel(cfg, EPI_PUSH, a, st));
    }
    SB_CUDA(cudaEventRecord(ws->ev1, st));
    if (t_dev) SB_CUDA(cudaMemcpyAsync(t_dev, ws->t[nterms & 1].p, n * 8, cudaMemcpyDeviceToDevice, st));
    std::vector<double> log(nterms + 1);
    SB_CUDA(cudaMemcpyAsync(log.data(), ws->norm_log.p, (nterms + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (elapsed_ms) SB_CUDA(cudaEventElapsedTime(elapsed_ms, ws->ev0, ws->ev1));
    if (term_norms)
        for (uint64_t k = 0; k < nterms; k++) term_norms[k] = std::sqrt(log[k + 1]);
    if (plog.p) {
        unsigned long long h[64];
        SB_CUDA(cudaMemcpy(h, plog.p, sizeof(h), cudaMemcpyDeviceToHost));
        const char *names[] = {"ring+tma_issue", "wait_cols+gather_issue", "load_rows", "cp_async_wait", "wait_vals",
                               "barrier1", "product", "barrier2", "sum+epilogue", "fence+barrier3"};
        unsigned long long tot = 0;
        for (int i = 0; i < 10; i++) tot += h[i];
        fprintf(stderr, "[sublinear_b200] push-kernel phase cycles (thread %llu, %llu CTA-launches, %llu tiles):\n", h[63], h[32],
                (unsigned long long)m->ntiles * nterms);
        for (int i = 0; i < 10; i++)
            fprintf(stderr, "  %-24s %6.1f %%  %8.0f cycles/tile\n", names[i], 100.0 * h[i] / (double)tot,
                    (double)h[i] / ((double)m->ntiles * nterms));
    }
    return SB200_OK;
}

}  // extern "C"
