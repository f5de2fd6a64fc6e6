// solver.hpp — NeumannSolver handle and the device-side solve core (internal).
#pragma once

#include "matrix.hpp"

struct sb200_solver {
    uint64_t max_terms = 50;          // NeumannSolver.max_terms          (ref src/solver/neumann.rs:26)
    double series_tolerance = 1e-8;   // NeumannSolver.series_tolerance   (:28)
    int adaptive_truncation = 1;      // (:30)
    int cache_powers = 1;             // (:32) carried, unused: the reference never fills matrix_powers
};

namespace sb200 {

struct SolveStats {
    uint64_t iterations = 0, terms = 0, matvec = 0, launches = 0;
    bool converged = false, series_converged = false, nonfinite = false;
    double residual_norm = 0, last_term_norm = 0, rhs_norm = 0;
    float device_ms = 0;
    double push_ms = 0, resid_ms = 0;   // per-kind CUDA-event totals (enable_profiling)
    uint64_t push_count = 0, resid_count = 0;
};

int32_t validate_options(const sb200_options *opt);
// argument checks of NeumannState::new in the reference's order (neumann.rs:147-206) + cached matrix analysis
int32_t solve_precheck(const sb200_solver *s, const sb200_matrix *m, uint64_t blen, const sb200_options *opt);
// streaming (sb200_solve_streaming): every `interval` iterations the loop state is read back anyway; the hook then
// copies the iterate into `host_x` (pinned, n doubles) and calls `fn`. A non-zero return stops the loop.
struct StreamHook {
    uint64_t interval = 0;
    sb200_stream_callback fn = nullptr;
    void *user = nullptr;
    double *host_x = nullptr;
    double t0_ms = 0.0;
};
int32_t solve_device(const sb200_solver *s, sb200_matrix *m, const double *b_dev, const double *x0_dev,
                     const sb200_options *opt, double *x_out_dev, cudaStream_t st, Workspace &ws, SolveStats &stats,
                     const StreamHook *hook = nullptr);

}  // namespace sb200
