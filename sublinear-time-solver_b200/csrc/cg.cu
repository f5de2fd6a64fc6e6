// cg.cu — conjugate gradient on the push path's SpMV kernel.
//
// OptimizedConjugateGradientSolver::solve (ref src/optimized_solver.rs:182-295); FastConjugateGradient::solve
// (src/fast_solver.rs:126-178) and UltraFastCG::solve (src/ultra_fast.rs:116-158) run the same loop and differ only in
// the summation order of their dot products, so one device implementation serves the three.
//
// Per iteration three launches, no host synchronisation:
//   warp_kernel<EPI_CG>  ap = A p, partial p.ap ; its last CTA sets alpha = rsold / p.ap  (or stops on |p.ap| < 1e-16)
//   cg_vec phase 1       x += alpha p ; r -= alpha ap ; partial r.r ; last CTA: beta, rsold, iteration count, loop test
//   cg_vec phase 2       p = r + beta p
// The loop decisions (`iteration < max_iterations`, `rsold <= tolerance^2`, the p.ap guard) are taken on the device by
// the last CTA of each reducing kernel (LoopCtl, device_util.cuh); the host enqueues iterations in batches and reads the
// 1-cache-line loop state back once per batch. Dot products are fixed-order two-stage reductions (CTA tree, then the
// CTA partials in index order): reproducible run to run, different from the reference's sequential sums in the last
// bits only (the reference's own three CG variants differ from each other in the same way).
#include <chrono>
#include <cmath>
#include <cstring>

#include "matrix.hpp"

using namespace sb200;

namespace {

double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct CgStats {
    uint64_t iterations = 0, matvecs = 0, launches = 0;
    bool converged = false, breakdown = false;
    double rsold = 0.0;
    float device_ms = 0.f;
    double spmv_ms = 0.0;
    uint64_t spmv_count = 0;
};

int32_t cg_precheck(const sb200_matrix *m, uint64_t blen, const sb200_cg_config *cfg) {
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    if (!cfg) return fail(SB200_ERR_INVALID_INPUT, "config is null");
    if (m->distributed) return fail(SB200_ERR_INVALID_INPUT, "row-block matrix: the CG path is single-GPU");
    if (m->nrows != m->ncols)  // optimized_solver.rs:188-190
        return fail(SB200_ERR_INVALID_INPUT, "Matrix must be square (%llu x %llu)", (unsigned long long)m->nrows,
                    (unsigned long long)m->ncols);
    if (blen != m->nrows)  // optimized_solver.rs:191-193
        return fail(SB200_ERR_DIMENSION_MISMATCH, "Right-hand side vector length must match matrix size: expected %llu, actual %llu",
                    (unsigned long long)m->nrows, (unsigned long long)blen);
    if (cfg->max_iterations >= 0xFFFFFFFFull) return fail(SB200_ERR_INVALID_INPUT, "max_iterations must fit 32 bits");
    return SB200_OK;
}

// the loop on device-resident vectors: b_dev (n) -> x_dev (n); r, p, ap live in the pooled workspace
int32_t cg_device(sb200_matrix *m, const double *b_dev, const sb200_cg_config *cfg, double *x_dev, cudaStream_t st,
                  Workspace &ws, CgStats &stats) {
    const uint64_t n = m->nrows;
    const size_t npart = 2 * (size_t)std::max(tile_kernel_max_grid(-1, EPI_CG), cg_vec_grid()) + 2;
    SB_TRY(ws.ensure(n, n, npart));
    double *r = ws.t[0].p, *p = ws.t[1].p, *ap = ws.c.p;

    LoopCtl h{};
    h.alive = 1;
    h.max_iterations = (uint32_t)cfg->max_iterations;
    h.cg_tol_sq = cfg->tolerance * cfg->tolerance;  // :210
    *ws.h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws.ctl.p, ws.h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));

    struct ProfEv {
        cudaEvent_t e0, e1;
        uint64_t it;
    };
    std::vector<ProfEv> prof;
    struct ProfCleanup {
        std::vector<ProfEv> &v;
        ~ProfCleanup() {
            for (auto &q : v) { cudaEventDestroy(q.e0); cudaEventDestroy(q.e1); }
        }
    } prof_cleanup{prof};
    const bool profiling = cfg->enable_profiling != 0;

    TileKernelArgs spmv{};
    fill_tile_args(m, spmv);
    spmv.xin = p;
    spmv.xin_own = p;
    spmv.out = ap;
    spmv.ctl = ws.ctl.p;
    spmv.partials = ws.partials.p;
    spmv.acc = ws.tmp.p;

    CgVecArgs va{};
    va.b = b_dev;
    va.x = x_dev;
    va.r = r;
    va.p = p;
    va.ap = ap;
    va.n = n;
    va.ctl = ws.ctl.p;
    va.partials = ws.partials.p;

    auto read_ctl = [&]() -> int32_t {
        SB_CUDA(cudaMemcpyAsync(ws.h_ctl, ws.ctl.p, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return SB200_OK;
    };

    uint64_t launches = 0;
    SB_CUDA(cudaEventRecord(ws.ev0, st));
    va.phase = 0;  // x = 0, r = p = b, rsold = b.b (:202-215)
    SB_TRY(launch_cg_vec(va, st));
    launches++;

    const uint64_t max_it = cfg->max_iterations;
    const uint64_t kBatch = 8;
    uint64_t it = 0;
    bool alive = max_it > 0;
    while (alive && it < max_it) {
        const uint64_t end = std::min(max_it, it + kBatch);
        for (; it < end; it++) {
            if (profiling) {
                ProfEv q{nullptr, nullptr, it};
                SB_CUDA(cudaEventCreate(&q.e0));
                SB_CUDA(cudaEventCreate(&q.e1));
                SB_CUDA(cudaEventRecord(q.e0, st));
                prof.push_back(q);
            }
            SB_TRY(launch_tile_kernel(-1, EPI_CG, spmv, st));
            if (profiling) SB_CUDA(cudaEventRecord(prof.back().e1, st));
            va.phase = 1;
            SB_TRY(launch_cg_vec(va, st));
            va.phase = 2;
            SB_TRY(launch_cg_vec(va, st));
            launches += 2 + launches_per_pass(m);
        }
        SB_TRY(read_ctl());
        alive = ws.h_ctl->alive != 0;
    }
    SB_CUDA(cudaEventRecord(ws.ev1, st));
    SB_TRY(read_ctl());
    SB_CUDA(cudaEventElapsedTime(&stats.device_ms, ws.ev0, ws.ev1));

    const LoopCtl &c = *ws.h_ctl;
    for (auto &q : prof) {  // launches past the end of the loop were no-ops
        if (q.it >= c.cg_matvecs) continue;
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, q.e0, q.e1));
        stats.spmv_ms += ms;
        stats.spmv_count++;
    }
    stats.iterations = c.iterations;
    stats.matvecs = c.cg_matvecs;
    stats.converged = c.cg_converged != 0;
    stats.breakdown = c.cg_breakdown != 0;
    stats.rsold = c.cg_rsold;
    stats.launches = launches;
    return SB200_OK;
}

void cg_fill_result(const sb200_matrix *m, const CgStats &st, double total_ms, sb200_cg_result *out) {
    out->residual_norm = std::sqrt(st.rsold);  // :275
    out->iterations = st.iterations;
    out->converged = st.converged;
    out->breakdown = st.breakdown;
    out->computation_time_ms = total_ms;
    out->matvec_count = st.matvecs;
    // the reference's solve() never bumps these two counters although its unit test expects them non-zero
    // (optimized_solver.rs:297-320 vs :421-435); they report the reductions / vector updates actually performed
    out->dot_product_count = 1 + st.matvecs + st.iterations;
    out->axpy_count = 3 * st.iterations;
    out->total_flops = st.matvecs * m->nnz * 2 + st.iterations * m->nrows * 6;  // :278-279
    out->average_bandwidth_gbs = 0.0;
    out->average_gflops = 0.0;
    if (total_ms > 0.0) {  // :281-285 (the reference's own formulas)
        out->average_bandwidth_gbs = (double)(out->total_flops * 8) / 1e9 / (total_ms / 1000.0);
        out->average_gflops = (double)out->total_flops / (total_ms * 1e6);
    }
    out->device_time_ms = st.device_ms;
    out->kernel_launches = st.launches;
    out->spmv_kernel_ms = st.spmv_ms;
    out->spmv_kernel_count = st.spmv_count;
}

struct WsLease {
    sb200_matrix *m;
    std::unique_ptr<Workspace> ws;
    explicit WsLease(sb200_matrix *mm) : m(mm), ws(matrix_acquire_ws(mm)) {}
    ~WsLease() { matrix_release_ws(m, std::move(ws)); }
};

int32_t cg_host(const sb200_matrix *m, const double *b, uint64_t blen, const sb200_cg_config *cfg, double *x_out,
                bool own_solution, sb200_cg_result *out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "result is null");
    memset(out, 0, sizeof(*out));
    const double t0 = wall_ms();
    SB_TRY(cg_precheck(m, blen, cfg));
    if (blen && !b) return fail(SB200_ERR_INVALID_INPUT, "b is null");
    DeviceGuard g(m->device);
    SB_TRY(require_device(m->device));
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    const uint64_t n = m->nrows;
    WsLease lease(mm);
    Workspace &ws = *lease.ws;
    SB_TRY(ws.ensure(n, n, 4));
    cudaStream_t st = m->stream;
    SB_TRY(copy_h2d(ws.b.p, b, n * 8, st));
    CgStats stats{};
    SB_TRY(cg_device(mm, ws.b.p, cfg, ws.x.p, st, ws, stats));
    double *dst = x_out;
    if (own_solution) {
        dst = (double *)pinned_pool_get(n * 8);
        if (!dst) return fail(SB200_ERR_MEMORY_ALLOCATION, "pinned allocation of %llu bytes failed", (unsigned long long)(n * 8));
        out->solution = dst;
        out->solution_len = n;
    }
    if (dst) SB_TRY(copy_d2h(dst, ws.x.p, n * 8, st));
    SB_CUDA(cudaStreamSynchronize(st));
    cg_fill_result(m, stats, wall_ms() - t0, out);
    out->h2d_bytes = n * 8;
    out->d2h_bytes = dst ? n * 8 : 0;
    return SB200_OK;
}

}  // namespace

extern "C" {

// OptimizedSolverConfig::default (optimized_solver.rs:119-127)
void sb200_cg_config_default(sb200_cg_config *c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->max_iterations = 1000;
    c->tolerance = 1e-6;
    c->enable_profiling = 0;
}

int32_t sb200_cg_solve(const sb200_matrix *m, const double *b, uint64_t blen, const sb200_cg_config *cfg,
                       sb200_cg_result *out) {
    return cg_host(m, b, blen, cfg, nullptr, true, out);
}

int32_t sb200_cg_solve_into(const sb200_matrix *m, const double *b, uint64_t blen, const sb200_cg_config *cfg,
                            double *x_out, sb200_cg_result *out) {
    if (blen && !x_out) return fail(SB200_ERR_INVALID_INPUT, "x_out is null");
    return cg_host(m, b, blen, cfg, x_out, false, out);
}

int32_t sb200_cg_solve_dev(const sb200_matrix *m, const double *b_dev, uint64_t blen, const sb200_cg_config *cfg,
                           double *x_dev, void *stream, sb200_cg_result *out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "result is null");
    memset(out, 0, sizeof(*out));
    const double t0 = wall_ms();
    SB_TRY(cg_precheck(m, blen, cfg));
    if (blen && (!b_dev || !x_dev)) return fail(SB200_ERR_INVALID_INPUT, "null device vector");
    DeviceGuard g(m->device);
    SB_TRY(require_device(m->device));
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    WsLease lease(mm);
    CgStats stats{};
    SB_TRY(cg_device(mm, b_dev, cfg, x_dev, (cudaStream_t)stream, *lease.ws, stats));
    cg_fill_result(m, stats, wall_ms() - t0, out);
    return SB200_OK;
}

void sb200_cg_result_free(sb200_cg_result *r) {
    if (!r) return;
    if (r->solution) pinned_pool_put(r->solution);
    r->solution = nullptr;
    r->solution_len = 0;
}

}  // extern "C"
