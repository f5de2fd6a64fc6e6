// state.cu — the SolverAlgorithm state interface of the Neumann path:
//   initialize / step / is_converged / extract_solution / update_rhs   (ref src/solver/mod.rs:223-252)
//   SolverState::{residual_norm, matvec_count, error_bounds, reset}     (ref src/solver/neumann.rs:350-378)
// and the streaming solve (SolverOptions.streaming_interval + PartialSolution, ref src/solver/mod.rs:33-34,198-217;
// the callback shape follows WasmSublinearSolver::solve_stream, ref src/wasm_iface.rs:119-166).
//
// The reference's NeumannSolver::step returns an error because its state holds no matrix reference
// (neumann.rs:390-403); the body it left commented out (:404-418) is what runs here: next term, residual, error
// bounds, Converged once the series converged or max_terms is reached. A state handle keeps its matrix handle alive
// and owns the device vectors of NeumannState (:97-135): solution, rhs = D^-1 b,
// current_term, plus b itself. Every step is the same fused push / residual kernels the batch solve uses. The state
// shares ownership of the matrix handle (reference count), so the two may be freed in any order.
#include <cmath>
#include <cstring>

#include "matrix.hpp"
#include "solver.hpp"

using namespace sb200;

struct sb200_state {
    sb200_solver solver;
    sb200_options opt;
    sb200_matrix *m = nullptr;
    std::unique_ptr<Workspace> ws;
    DevBuf<uint64_t> d_idx;
    DevBuf<double> d_delta;
    uint64_t n = 0;
    int cur = 0;  // ws->t[cur] holds current_term
    uint64_t terms = 0, matvec = 0;
    bool sconv = false;
    double residual_norm = INFINITY, term_norm = 0.0, rhs_norm = 0.0;
    bool has_bound = false;
    double bound = 0.0;

    ~sb200_state() {
        if (!m) return;
        if (ws) {
            DeviceGuard g(m->device);
            matrix_release_ws(m, std::move(ws));
        }
        matrix_release(m);  // shared ownership: the caller may have freed its matrix handle first
    }
};

namespace {

int32_t upload_ctl(sb200_state *st) {
    LoopCtl h{};
    h.res_norm = st->residual_norm;
    h.tolerance = st->opt.tolerance;
    h.series_tolerance = st->solver.series_tolerance;
    h.max_terms = 0xFFFFFFFFu;       // the loop decisions are taken on the host, one step at a time
    h.max_iterations = 0xFFFFFFFFu;
    h.alive = 1;
    h.terms = (uint32_t)st->terms;
    *st->ws->h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(st->ws->ctl.p, st->ws->h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st->m->stream));
    return SB200_OK;
}

int32_t read_ctl(sb200_state *st) {
    SB_CUDA(cudaMemcpyAsync(st->ws->h_ctl, st->ws->ctl.p, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st->m->stream));
    SB_CUDA(cudaStreamSynchronize(st->m->stream));
    return SB200_OK;
}

// ||rhs||_2 (the error bound's scale, neumann.rs:329-343)
int32_t refresh_rhs_norm(sb200_state *st) {
    SB_TRY(launch_state_vec(1, st->ws->c.p, nullptr, st->n, st->ws->ctl.p, st->ws->partials.p, st->m->stream));
    SB_TRY(read_ctl(st));
    st->rhs_norm = std::sqrt(st->ws->h_ctl->red[0]);
    return SB200_OK;
}

// estimate_error_bounds (neumann.rs:321-347)
void estimate_error_bounds(sb200_state *st) {
    if (!st->sconv || st->terms == 0) return;
    double est = 0.0;
    if (st->terms > 1) est = std::pow(st->term_norm / st->rhs_norm, 1.0 / (double)(st->terms - 1));
    if (est < 1.0) {
        st->has_bound = true;
        st->bound = std::pow(est, (double)(int)st->terms) / (1.0 - est) * st->rhs_norm;
    }
}

bool state_converged(const sb200_state *st) {  // is_converged (neumann.rs:422-430)
    return st->residual_norm <= st->opt.tolerance || (st->sconv && st->terms < st->solver.max_terms);
}

}  // namespace

extern "C" {

// SolverAlgorithm::initialize -> NeumannState::new (neumann.rs:139-249, 381-388)
int32_t sb200_neumann_initialize(const sb200_solver *s, const sb200_matrix *m, const double *b, uint64_t blen,
                                 const sb200_options *opt, sb200_state **out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    DeviceGuard g(m->device);
    SB_TRY(solve_precheck(s, m, blen, opt));
    if (blen && !b) return fail(SB200_ERR_INVALID_INPUT, "b is null");
    if (s->max_terms >= 0xFFFFFFFFull) return fail(SB200_ERR_INVALID_INPUT, "max_terms must fit 32 bits");
    std::unique_ptr<sb200_state> st(new sb200_state());
    st->solver = *s;
    st->opt = *opt;
    st->opt.initial_guess = nullptr;  // consumed below; the caller's buffer is not retained
    st->opt.initial_guess_len = 0;
    st->m = const_cast<sb200_matrix *>(m);
    matrix_retain(st->m);
    st->n = m->nrows;
    st->ws = matrix_acquire_ws(st->m);
    const uint64_t n = st->n;
    const size_t npart = 2 * (size_t)std::max(std::max(tile_kernel_max_grid(m->tile_cfg, EPI_PUSH),
                                                       tile_kernel_max_grid(m->tile_cfg, EPI_RESID)),
                                              init_state_grid()) + 2;
    SB_TRY(st->ws->ensure(n, n, npart));
    Workspace &ws = *st->ws;
    cudaStream_t stream = m->stream;
    const bool compat = opt->mode == SB200_MODE_REF_COMPAT;
    SB_TRY(copy_h2d(ws.b.p, b, n * 8, stream));
    const double *x0 = nullptr, *ax0 = nullptr;
    DevBuf<double> guess;
    if (opt->initial_guess) {
        SB_TRY(guess.alloc(n));
        SB_TRY(copy_h2d(guess.p, opt->initial_guess, n * 8, stream));
        x0 = guess.p;
        if (!compat) {  // correct mode: t0 = D^-1 (b - A x0), one SpMV
            SB_TRY(matrix_spmv_dev(m, x0, ws.tmp.p, 0, stream));
            st->matvec++;
            ax0 = ws.tmp.p;
        }
    }
    SB_TRY(upload_ctl(st.get()));
    InitArgs ia{};
    ia.b = ws.b.p;
    ia.dinv = m->d_dinv[opt->mode].p;
    ia.x0 = x0;
    ia.ax0 = ax0;
    ia.c_out = ws.c.p;
    ia.t_out = ws.t[0].p;
    ia.x_out = ws.x.p;
    ia.n = (uint32_t)n;
    ia.compat = compat;
    ia.ctl = ws.ctl.p;
    ia.partials = ws.partials.p;
    ia.skip_term0 = 1;  // NeumannState::new adds no term: solution = initial_guess or rhs (compat) / 0 (correct)
    SB_TRY(launch_init_state(ia, stream));
    SB_TRY(upload_ctl(st.get()));  // the init kernel's tail counted a term on the scratch state
    SB_TRY(refresh_rhs_norm(st.get()));
    SB_TRY(upload_ctl(st.get()));
    SB_CUDA(cudaStreamSynchronize(stream));
    *out = st.release();
    return SB200_OK;
}

// SolverAlgorithm::step (neumann.rs:390-419, the intended body :404-418)
int32_t sb200_state_step(sb200_state *st, int32_t *step_result) {
    clear_error();
    if (!st || !step_result) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    sb200_matrix *m = st->m;
    DeviceGuard g(m->device);
    Workspace &ws = *st->ws;
    cudaStream_t stream = m->stream;
    const bool compat = st->opt.mode == SB200_MODE_REF_COMPAT;
    bool term_added = false;
    if (st->terms < st->solver.max_terms) {  // compute_next_term (:252-277)
        if (st->terms == 0) {
            SB_TRY(launch_state_vec(0, ws.t[st->cur].p, ws.x.p, st->n, ws.ctl.p, ws.partials.p, stream));
        } else {
            TileKernelArgs a{};
            fill_tile_args(m, a);
            a.ctl = ws.ctl.p;
            a.partials = ws.partials.p;
            a.acc = ws.tmp.p;
            a.xin = ws.t[st->cur].p;
            a.xin_own = a.xin;
            a.out = ws.t[st->cur ^ 1].p;
            a.sol = ws.x.p;
            a.dinv = m->d_dinv[st->opt.mode].p;
            a.it = (uint32_t)st->terms;
            a.force = 1;
            SB_TRY(launch_tile_kernel(m->tile_cfg, EPI_PUSH, a, stream));
            st->cur ^= 1;
            st->matvec++;
        }
        term_added = true;
    }
    {  // update_residual (:302-318)
        TileKernelArgs a{};
        fill_tile_args(m, a);
        a.ctl = ws.ctl.p;
        a.partials = ws.partials.p;
        a.acc = ws.tmp.p;
        a.xin = ws.x.p;
        a.xin_own = a.xin;
        a.rhs = compat ? ws.c.p : ws.b.p;
        a.force = 1;
        SB_TRY(launch_tile_kernel(m->tile_cfg, EPI_RESID, a, stream));
        st->matvec++;
    }
    SB_TRY(read_ctl(st));
    const LoopCtl &c = *ws.h_ctl;
    if (term_added) {
        st->terms += 1;
        st->term_norm = std::sqrt(c.term_norm2);
        if (st->term_norm < st->solver.series_tolerance) st->sconv = true;  // :271-274
    }
    st->residual_norm = c.res_norm;
    if (st->solver.adaptive_truncation) estimate_error_bounds(st);          // :410-412
    *step_result = (st->sconv || st->terms >= st->solver.max_terms) ? SB200_STEP_CONVERGED : SB200_STEP_CONTINUE;
    if (!std::isfinite(st->residual_norm))  // SolverAlgorithm::solve's check after a step (solver/mod.rs:271-279)
        return fail(SB200_ERR_NUMERICAL_INSTABILITY, "Non-finite residual norm after %llu terms", (unsigned long long)st->terms);
    return SB200_OK;
}

int32_t sb200_state_is_converged(const sb200_state *st, int32_t *out) {
    if (!st || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = state_converged(st);
    return SB200_OK;
}

// SolverAlgorithm::extract_solution (neumann.rs:432-434)
int32_t sb200_state_extract_solution(const sb200_state *st, double *x, uint64_t xlen) {
    clear_error();
    if (!st || (xlen && !x)) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (xlen != st->n)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in extract_solution", (unsigned long long)st->n,
                    (unsigned long long)xlen);
    DeviceGuard g(st->m->device);
    SB_TRY(copy_d2h(x, st->ws->x.p, st->n * 8, st->m->stream));
    SB_CUDA(cudaStreamSynchronize(st->m->stream));
    return SB200_OK;
}

// SolverAlgorithm::update_rhs (neumann.rs:436-462)
int32_t sb200_state_update_rhs(sb200_state *st, const uint64_t *indices, const double *deltas, uint64_t count) {
    clear_error();
    if (!st || (count && (!indices || !deltas))) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    for (uint64_t k = 0; k < count; k++)
        if (indices[k] >= st->n)  // :439-445, checked before anything is changed
            return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "index %llu out of bounds (max %llu) in rhs_update",
                        (unsigned long long)indices[k], (unsigned long long)(st->n ? st->n - 1 : 0));
    sb200_matrix *m = st->m;
    DeviceGuard g(m->device);
    Workspace &ws = *st->ws;
    cudaStream_t stream = m->stream;
    const bool compat = st->opt.mode == SB200_MODE_REF_COMPAT;
    if (count) {
        if (st->d_idx.n < count) SB_TRY(st->d_idx.alloc(count));
        if (st->d_delta.n < count) SB_TRY(st->d_delta.alloc(count));
        SB_TRY(copy_h2d(st->d_idx.p, indices, count * 8, stream));
        SB_TRY(copy_h2d(st->d_delta.p, deltas, count * 8, stream));
    }
    double *term = ws.t[st->cur].p;
    if (compat) {
        // the literal code: rhs and the solution take the scaled delta, the series restarts from the whole rhs (:448-459)
        SB_TRY(launch_update_rhs(st->d_idx.p, st->d_delta.p, count, m->d_dinv[st->opt.mode].p, ws.b.p, ws.c.p, ws.x.p, stream));
        SB_CUDA(cudaMemcpyAsync(term, ws.c.p, st->n * 8, cudaMemcpyDeviceToDevice, stream));
    } else {
        // the incremental solve the comment at :451-453 asks for: restart the series from D^-1 delta_b alone; further
        // steps add A^-1 delta_b to the solution held
        SB_CUDA(cudaMemsetAsync(term, 0, st->n * 8, stream));
        SB_TRY(launch_update_rhs(st->d_idx.p, st->d_delta.p, count, m->d_dinv[st->opt.mode].p, ws.b.p, ws.c.p, term, stream));
    }
    st->terms = 0;      // :458
    st->sconv = false;  // :459
    SB_TRY(upload_ctl(st));
    SB_TRY(refresh_rhs_norm(st));
    SB_TRY(upload_ctl(st));
    SB_CUDA(cudaStreamSynchronize(stream));
    return SB200_OK;
}

// SolverState::reset (neumann.rs:367-378)
int32_t sb200_state_reset(sb200_state *st) {
    clear_error();
    if (!st) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    DeviceGuard g(st->m->device);
    Workspace &ws = *st->ws;
    cudaStream_t stream = st->m->stream;
    SB_CUDA(cudaMemsetAsync(ws.x.p, 0, st->n * 8, stream));
    SB_CUDA(cudaMemcpyAsync(ws.t[st->cur].p, ws.c.p, st->n * 8, cudaMemcpyDeviceToDevice, stream));
    st->residual_norm = INFINITY;
    st->terms = 0;
    st->matvec = 0;
    st->sconv = false;
    st->has_bound = false;
    SB_TRY(upload_ctl(st));
    SB_CUDA(cudaStreamSynchronize(stream));
    return SB200_OK;
}

// SolverState::{residual_norm, matvec_count, error_bounds, memory_usage} (neumann.rs:350-365) + the series counters
int32_t sb200_state_info(const sb200_state *st, sb200_state_info_t *info) {
    if (!st || !info) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    memset(info, 0, sizeof(*info));
    info->dimension = st->n;
    info->residual_norm = st->residual_norm;
    info->matvec_count = st->matvec;
    info->terms_computed = st->terms;
    info->series_converged = st->sconv;
    info->last_term_norm = st->term_norm;
    info->has_error_bounds = st->has_bound;
    info->error_upper_bound = st->bound;
    info->memory_bytes = st->ws->bytes();
    return SB200_OK;
}

void sb200_state_free(sb200_state *st) { delete st; }

}  // extern "C"
