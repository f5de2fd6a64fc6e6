// sublinear.hpp — C++17 host-side mirror of the reference crate's API for the Neumann / push path, over the C ABI
// (include/sublinear_b200.h). Same names, argument meaning and error behaviour as the Rust types:
//   sublinear::SparseMatrix   <-> SparseMatrix  (src/matrix/mod.rs:123-372) + trait Matrix (:25-104)
//   sublinear::NeumannSolver  <-> NeumannSolver (src/solver/neumann.rs:24-93, 469-555) / trait SolverAlgorithm
//   sublinear::SolverOptions  <-> SolverOptions (src/solver/mod.rs:22-116)
//   sublinear::SolverResult   <-> SolverResult  (src/solver/mod.rs:121-195)
//   sublinear::SolverError    <-> SolverError   (src/error.rs:16-138), thrown where Rust returns Err(..)
//   sublinear::OptimizedSparseMatrix / OptimizedConjugateGradientSolver / OptimizedSolverConfig / OptimizedSolverResult
//                             <-> src/optimized_solver.rs:16-340 (the CG consumer of the same SpMV kernel)
//   sublinear::NeumannState + NeumannSolver::{initialize, step, is_converged, extract_solution, update_rhs}
//                             <-> trait SolverAlgorithm / SolverState (src/solver/mod.rs:223-352, neumann.rs:350-462)
//   sublinear::PushGraph / ForwardPushSolver / BackwardPushSolver
//                             <-> src/graph/adjacency.rs:199-277, src/solver/forward_push.rs, backward_push.rs
// Header only; link against libsublinear_b200.so. The Rust toolchain is not available in the build image, so this is
// the compiled-language host layer (the Rust shim in ../rust/ is the same mapping, shipped unbuilt).
#pragma once

#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/sublinear_b200.h"

namespace sublinear {

using Precision = double;   // src/types.rs:19
using IndexType = uint32_t; // src/types.rs:22

enum class ErrorKind : int32_t {
    MatrixNotDiagonallyDominant = 1, NumericalInstability, ConvergenceFailure, InvalidInput, DimensionMismatch,
    UnsupportedMatrixFormat, MemoryAllocationError, IndexOutOfBounds, InvalidSparseMatrix, AlgorithmError,
    WasmBindingError, IoError, SerializationError
};

class SolverError : public std::runtime_error {
public:
    SolverError(int32_t code, std::string msg) : std::runtime_error(std::move(msg)), kind(static_cast<ErrorKind>(code)) {}
    ErrorKind kind;
    // fields of ConvergenceFailure / NumericalInstability (src/error.rs:28-47)
    uint64_t iterations = 0;
    double residual_norm = 0.0;
    // SolverError::is_recoverable (src/error.rs:141-152)
    bool is_recoverable() const {
        return kind == ErrorKind::ConvergenceFailure || kind == ErrorKind::NumericalInstability ||
               kind == ErrorKind::MatrixNotDiagonallyDominant;
    }
};

namespace detail {
inline std::string last_error() {
    char buf[1024];
    sb200_last_error(buf, sizeof(buf));
    return buf;
}
inline void check(int32_t rc) {
    if (rc != SB200_OK) throw SolverError(rc, last_error());
}
}  // namespace detail

enum class SolveMode : int32_t { Correct = SB200_MODE_CORRECT, RefCompat = SB200_MODE_REF_COMPAT };

struct SolverOptions {
    Precision tolerance = 1e-6;
    uint64_t max_iterations = 1000;
    int32_t convergence_mode = SB200_CONV_RESIDUAL_NORM;
    int32_t norm_type = SB200_NORM_L2;
    bool collect_stats = false;
    uint64_t streaming_interval = 0;
    std::optional<std::vector<Precision>> initial_guess;
    bool compute_error_bounds = false;
    Precision error_bounds_tolerance = 1e-8;
    bool enable_profiling = false;
    std::optional<uint64_t> random_seed;
    SolveMode mode = SolveMode::Correct;
    int32_t dominance = SB200_DOMINANCE_ROW;
    int32_t residual_check = SB200_RESIDUAL_EVERY_5;

    static SolverOptions from(const sb200_options &o) {
        SolverOptions s;
        s.tolerance = o.tolerance; s.max_iterations = o.max_iterations; s.convergence_mode = o.convergence_mode;
        s.norm_type = o.norm_type; s.collect_stats = o.collect_stats; s.streaming_interval = o.streaming_interval;
        s.compute_error_bounds = o.compute_error_bounds; s.error_bounds_tolerance = o.error_bounds_tolerance;
        s.enable_profiling = o.enable_profiling;
        return s;
    }
    static SolverOptions high_precision() { sb200_options o; sb200_options_high_precision(&o); return from(o); }
    static SolverOptions fast() { sb200_options o; sb200_options_fast(&o); return from(o); }
    static SolverOptions streaming(uint64_t interval) { sb200_options o; sb200_options_streaming(&o, interval); return from(o); }

    sb200_options to_c() const {
        sb200_options o;
        sb200_options_default(&o);
        o.tolerance = tolerance; o.max_iterations = max_iterations; o.convergence_mode = convergence_mode;
        o.norm_type = norm_type; o.collect_stats = collect_stats; o.streaming_interval = streaming_interval;
        if (initial_guess) { o.initial_guess = initial_guess->data(); o.initial_guess_len = initial_guess->size(); }
        o.compute_error_bounds = compute_error_bounds; o.error_bounds_tolerance = error_bounds_tolerance;
        o.enable_profiling = enable_profiling; o.has_random_seed = random_seed.has_value();
        o.random_seed = random_seed.value_or(0); o.mode = static_cast<int32_t>(mode); o.dominance = dominance;
        o.residual_check = residual_check;
        return o;
    }
};

struct SolverStats {  // src/types.rs SolverStats: the fields this path fills
    double total_time_ms = 0.0;
    uint64_t matvec_count = 0;
};

struct SolverResult {
    std::vector<Precision> solution;
    Precision residual_norm = 0.0;
    uint64_t iterations = 0;
    bool converged = false;
    std::optional<Precision> error_upper_bound;
    std::optional<SolverStats> stats;
    uint64_t terms_computed = 0;
    bool series_converged = false;
    // SolverResult::meets_quality_criteria (src/solver/mod.rs:192-194)
    bool meets_quality_criteria(Precision tol) const { return converged && residual_norm <= tol; }
};

class SparseMatrix {
public:
    // SparseMatrix::from_triplets(Vec<(usize,usize,f64)>, rows, cols) (src/matrix/mod.rs:160-199)
    static SparseMatrix from_triplets(const std::vector<std::tuple<size_t, size_t, Precision>> &t, size_t rows, size_t cols) {
        std::vector<uint64_t> r(t.size()), c(t.size());
        std::vector<double> v(t.size());
        for (size_t i = 0; i < t.size(); i++) { r[i] = std::get<0>(t[i]); c[i] = std::get<1>(t[i]); v[i] = std::get<2>(t[i]); }
        sb200_matrix *h = nullptr;
        detail::check(sb200_matrix_from_triplets(r.data(), c.data(), v.data(), t.size(), rows, cols, &h));
        return SparseMatrix(h);
    }
    static SparseMatrix from_csr(const std::vector<uint32_t> &row_ptr, const std::vector<uint32_t> &col_indices,
                                 const std::vector<Precision> &values, size_t rows, size_t cols) {
        sb200_matrix *h = nullptr;
        detail::check(sb200_matrix_from_csr(row_ptr.data(), col_indices.data(), values.data(), rows, cols, values.size(), &h));
        return SparseMatrix(h);
    }
    static SparseMatrix from_dense(const std::vector<Precision> &data, size_t rows, size_t cols) {
        if (data.size() != rows * cols) throw SolverError(SB200_ERR_DIMENSION_MISMATCH, "dense_to_sparse_conversion");
        sb200_matrix *h = nullptr;
        detail::check(sb200_matrix_from_dense(data.data(), rows, cols, &h));
        return SparseMatrix(h);
    }
    static SparseMatrix identity(size_t n) { sb200_matrix *h = nullptr; detail::check(sb200_matrix_identity(n, &h)); return SparseMatrix(h); }
    static SparseMatrix diagonal(const std::vector<Precision> &d) {
        sb200_matrix *h = nullptr;
        detail::check(sb200_matrix_diagonal(d.data(), d.size(), &h));
        return SparseMatrix(h);
    }
    SparseMatrix(SparseMatrix &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    SparseMatrix &operator=(SparseMatrix &&o) noexcept { if (this != &o) { sb200_matrix_free(h_); h_ = o.h_; o.h_ = nullptr; } return *this; }
    SparseMatrix(const SparseMatrix &) = delete;
    SparseMatrix &operator=(const SparseMatrix &) = delete;
    ~SparseMatrix() { sb200_matrix_free(h_); }

    // trait Matrix (src/matrix/mod.rs:25-104)
    size_t rows() const { uint64_t v; detail::check(sb200_matrix_rows(h_, &v)); return v; }
    size_t cols() const { uint64_t v; detail::check(sb200_matrix_cols(h_, &v)); return v; }
    size_t nnz() const { uint64_t v; detail::check(sb200_matrix_nnz(h_, &v)); return v; }
    bool is_square() const { return rows() == cols(); }
    std::optional<Precision> get(size_t row, size_t col) const {
        double v; int32_t p;
        detail::check(sb200_matrix_get(h_, row, col, &v, &p));
        return p ? std::optional<Precision>(v) : std::nullopt;
    }
    bool is_diagonally_dominant() const { int32_t o; detail::check(sb200_matrix_is_diagonally_dominant(h_, SB200_DOMINANCE_ROW, &o)); return o; }
    std::optional<Precision> diagonal_dominance_factor() const {
        double f; int32_t p;
        detail::check(sb200_matrix_diagonal_dominance_factor(h_, &f, &p));
        return p ? std::optional<Precision>(f) : std::nullopt;
    }
    void multiply_vector(const std::vector<Precision> &x, std::vector<Precision> &result) const {
        detail::check(sb200_matrix_multiply_vector(h_, x.data(), x.size(), result.data(), result.size()));
    }
    void multiply_vector_add(const std::vector<Precision> &x, std::vector<Precision> &result) const {
        detail::check(sb200_matrix_multiply_vector_add(h_, x.data(), x.size(), result.data(), result.size()));
    }
    void scale(Precision factor) { detail::check(sb200_matrix_scale(h_, factor)); }
    const char *format_name() const { return "CSR"; }
    const sb200_matrix *handle() const { return h_; }

private:
    explicit SparseMatrix(sb200_matrix *h) : h_(h) {}
    sb200_matrix *h_ = nullptr;
};

enum class StepResult { Continue, Converged };  // src/solver/mod.rs:355-363 (Failed(String) surfaces as SolverError)

class NeumannSolver {
public:
    NeumannSolver(size_t max_terms, Precision series_tolerance) { detail::check(sb200_neumann_new(max_terms, series_tolerance, &h_)); }
    static NeumannSolver default_() { sb200_solver *h; detail::check(sb200_neumann_default(&h)); return NeumannSolver(h); }
    static NeumannSolver high_precision() { sb200_solver *h; detail::check(sb200_neumann_high_precision(&h)); return NeumannSolver(h); }
    static NeumannSolver fast() { sb200_solver *h; detail::check(sb200_neumann_fast(&h)); return NeumannSolver(h); }
    NeumannSolver(NeumannSolver &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    NeumannSolver(const NeumannSolver &) = delete;
    ~NeumannSolver() { sb200_solver_free(h_); }
    NeumannSolver &with_adaptive_truncation(bool e) { detail::check(sb200_neumann_with_adaptive_truncation(h_, e)); return *this; }
    NeumannSolver &with_power_caching(bool e) { detail::check(sb200_neumann_with_power_caching(h_, e)); return *this; }
    const char *algorithm_name() const { return sb200_solver_algorithm_name(h_); }

    // SolverAlgorithm::solve(&self, matrix, b, options) -> Result<SolverResult> (src/solver/neumann.rs:469-555)
    SolverResult solve(const SparseMatrix &matrix, const std::vector<Precision> &b, const SolverOptions &options = {}) const {
        sb200_options o = options.to_c();
        sb200_result r;
        SolverResult out;
        out.solution.resize(b.size());
        const int32_t rc = sb200_solve_into(h_, matrix.handle(), b.data(), b.size(), &o, out.solution.data(), &r);
        out.residual_norm = r.residual_norm; out.iterations = r.iterations; out.converged = r.converged;
        out.terms_computed = r.terms_computed; out.series_converged = r.series_converged;
        if (r.has_error_bounds) out.error_upper_bound = r.error_upper_bound;
        if (r.has_stats) out.stats = SolverStats{r.total_time_ms, r.matvec_count};
        if (rc != SB200_OK) {
            SolverError e(rc, detail::last_error());
            e.iterations = r.iterations;
            e.residual_norm = r.residual_norm;
            throw e;
        }
        return out;
    }

    // ---- trait SolverAlgorithm, the stepping half (src/solver/mod.rs:223-252) ----
    // initialize(&self, matrix, b, options) -> Result<State>
    inline class NeumannState initialize(const SparseMatrix &matrix, const std::vector<Precision> &b,
                                         const SolverOptions &options = {}) const;
    // step(&self, &mut state) -> Result<StepResult>: the body the reference left commented out (neumann.rs:404-418)
    inline StepResult step(class NeumannState &state) const;
    inline bool is_converged(const class NeumannState &state) const;
    inline std::vector<Precision> extract_solution(const class NeumannState &state) const;
    inline void update_rhs(class NeumannState &state, const std::vector<std::pair<size_t, Precision>> &delta_b) const;

private:
    explicit NeumannSolver(sb200_solver *h) : h_(h) {}
    sb200_solver *h_ = nullptr;
};

// NeumannState behind trait SolverState (src/solver/mod.rs:336-352, neumann.rs:350-378). It shares ownership of the
// matrix handle, so it may outlive the SparseMatrix it was initialised from.
class NeumannState {
public:
    NeumannState(NeumannState &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    NeumannState(const NeumannState &) = delete;
    ~NeumannState() { sb200_state_free(h_); }
    Precision residual_norm() const { return info().residual_norm; }
    size_t matvec_count() const { return info().matvec_count; }
    std::optional<Precision> error_bounds() const {
        const auto i = info();
        return i.has_error_bounds ? std::optional<Precision>(i.error_upper_bound) : std::nullopt;
    }
    size_t memory_usage() const { return info().memory_bytes; }
    size_t terms_computed() const { return info().terms_computed; }
    bool series_converged() const { return info().series_converged != 0; }
    void reset() { detail::check(sb200_state_reset(h_)); }
    sb200_state *handle() const { return h_; }

private:
    friend class NeumannSolver;
    explicit NeumannState(sb200_state *h) : h_(h) {}
    sb200_state_info_t info() const { sb200_state_info_t i; detail::check(sb200_state_info(h_, &i)); return i; }
    sb200_state *h_ = nullptr;
};

inline NeumannState NeumannSolver::initialize(const SparseMatrix &matrix, const std::vector<Precision> &b,
                                              const SolverOptions &options) const {
    sb200_options o = options.to_c();
    sb200_state *st = nullptr;
    detail::check(sb200_neumann_initialize(h_, matrix.handle(), b.data(), b.size(), &o, &st));
    return NeumannState(st);
}
inline StepResult NeumannSolver::step(NeumannState &state) const {
    int32_t r = 0;
    detail::check(sb200_state_step(state.handle(), &r));
    return r == SB200_STEP_CONVERGED ? StepResult::Converged : StepResult::Continue;
}
inline bool NeumannSolver::is_converged(const NeumannState &state) const {
    int32_t c = 0;
    detail::check(sb200_state_is_converged(state.handle(), &c));
    return c != 0;
}
inline std::vector<Precision> NeumannSolver::extract_solution(const NeumannState &state) const {
    sb200_state_info_t i;
    detail::check(sb200_state_info(state.handle(), &i));
    std::vector<Precision> x(i.dimension);
    detail::check(sb200_state_extract_solution(state.handle(), x.data(), x.size()));
    return x;
}
inline void NeumannSolver::update_rhs(NeumannState &state, const std::vector<std::pair<size_t, Precision>> &delta_b) const {
    std::vector<uint64_t> idx(delta_b.size());
    std::vector<double> dl(delta_b.size());
    for (size_t k = 0; k < delta_b.size(); k++) { idx[k] = delta_b[k].first; dl[k] = delta_b[k].second; }
    detail::check(sb200_state_update_rhs(state.handle(), idx.data(), dl.data(), dl.size()));
}

// ---- conjugate gradient on the same SpMV kernel (src/optimized_solver.rs) ---------------------------------------

// OptimizedSparseMatrix (src/optimized_solver.rs:16-107): CSR store + matvec / byte counters
class OptimizedSparseMatrix {
public:
    // from_triplets(Vec<(usize,usize,f64)>, rows, cols) -> Result<Self, String> (:40-56)
    static OptimizedSparseMatrix from_triplets(const std::vector<std::tuple<size_t, size_t, Precision>> &t, size_t rows,
                                               size_t cols) {
        return OptimizedSparseMatrix(SparseMatrix::from_triplets(t, rows, cols));
    }
    std::pair<size_t, size_t> dimensions() const { return {m_.rows(), m_.cols()}; }  // :59-61
    size_t nnz() const { return m_.nnz(); }                                         // :64-66
    // multiply_vector(&self, x, y) (:69-91): asserts the lengths, bumps matvec_count and bytes_processed
    void multiply_vector(const std::vector<Precision> &x, std::vector<Precision> &y) const {
        if (x.size() != m_.cols() || y.size() != m_.rows()) throw SolverError(SB200_ERR_DIMENSION_MISMATCH, "assert_eq!(x.len(), cols)");
        matvec_count_ += 1;
        bytes_processed_ += m_.nnz() * 8 + x.size() * 8 + y.size() * 8;  // :74
        m_.multiply_vector(x, y);
    }
    std::pair<size_t, size_t> get_performance_stats() const { return {matvec_count_, bytes_processed_}; }  // :94-99
    void reset_stats() const { matvec_count_ = 0; bytes_processed_ = 0; }                                 // :102-105
    const SparseMatrix &inner() const { return m_; }

private:
    explicit OptimizedSparseMatrix(SparseMatrix m) : m_(std::move(m)) {}
    SparseMatrix m_;
    mutable size_t matvec_count_ = 0, bytes_processed_ = 0;
};

struct OptimizedSolverConfig {  // :108-127
    size_t max_iterations = 1000;
    Precision tolerance = 1e-6;
    bool enable_profiling = false;
};

struct OptimizedSolverStats {  // :150-166
    size_t matvec_count = 0, dot_product_count = 0, axpy_count = 0, total_flops = 0;
    double average_bandwidth_gbs = 0.0, average_gflops = 0.0;
};

struct OptimizedSolverResult {  // :130-147
    std::vector<Precision> solution;
    Precision residual_norm = 0.0;
    size_t iterations = 0;
    bool converged = false;
    double computation_time_ms = 0.0;
    OptimizedSolverStats performance_stats;
    const std::vector<Precision> &data() const { return solution; }  // :336-340
};

class OptimizedConjugateGradientSolver {
public:
    explicit OptimizedConjugateGradientSolver(OptimizedSolverConfig config = {}) : config_(config) {}
    static OptimizedConjugateGradientSolver new_(OptimizedSolverConfig config) { return OptimizedConjugateGradientSolver(config); }

    // solve(&mut self, matrix, b) -> Result<OptimizedSolverResult, String> (:182-295); Err(String) -> SolverError
    OptimizedSolverResult solve(const OptimizedSparseMatrix &matrix, const std::vector<Precision> &b) {
        sb200_cg_config c;
        sb200_cg_config_default(&c);
        c.max_iterations = config_.max_iterations;
        c.tolerance = config_.tolerance;
        c.enable_profiling = config_.enable_profiling;
        sb200_cg_result r;
        OptimizedSolverResult out;
        out.solution.resize(b.size());
        detail::check(sb200_cg_solve_into(matrix.inner().handle(), b.data(), b.size(), &c, out.solution.data(), &r));
        out.residual_norm = r.residual_norm; out.iterations = r.iterations; out.converged = r.converged;
        out.computation_time_ms = r.computation_time_ms;
        out.performance_stats = {size_t(r.matvec_count), size_t(r.dot_product_count), size_t(r.axpy_count),
                                 size_t(r.total_flops), r.average_bandwidth_gbs, r.average_gflops};
        stats_ = out.performance_stats;
        return out;
    }
    size_t get_last_iteration_count() const { return stats_.matvec_count; }  // :323-325 (returns matvec_count)

private:
    OptimizedSolverConfig config_;
    OptimizedSolverStats stats_;
};

// ---- forward / backward push (src/graph/adjacency.rs, src/solver/forward_push.rs, backward_push.rs) ---------------

class PushGraph {
public:
    // PushGraph::from_matrix(&CompressedSparseRow) (adjacency.rs:211-224)
    static PushGraph from_matrix(const std::vector<uint64_t> &row_ptr, const std::vector<uint32_t> &col_indices,
                                 const std::vector<Precision> &values) {
        sb200_push_graph *h = nullptr;
        detail::check(sb200_push_graph_from_csr(row_ptr.data(), col_indices.data(), values.data(), row_ptr.size() - 1, &h));
        return PushGraph(h);
    }
    // PushGraph::from_edges(num_nodes, &[(from, to, weight)]) (adjacency.rs:227-239)
    static PushGraph from_edges(size_t num_nodes, const std::vector<std::tuple<size_t, size_t, Precision>> &edges) {
        std::vector<uint64_t> f(edges.size()), t(edges.size());
        std::vector<double> w(edges.size());
        for (size_t i = 0; i < edges.size(); i++) { f[i] = std::get<0>(edges[i]); t[i] = std::get<1>(edges[i]); w[i] = std::get<2>(edges[i]); }
        sb200_push_graph *h = nullptr;
        detail::check(sb200_push_graph_from_edges(num_nodes, f.data(), t.data(), w.data(), w.size(), &h));
        return PushGraph(h);
    }
    PushGraph(PushGraph &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    PushGraph(const PushGraph &) = delete;
    ~PushGraph() { sb200_push_graph_free(h_); }
    size_t num_nodes() const { uint64_t n; detail::check(sb200_push_graph_info(h_, &n, nullptr)); return n; }
    size_t num_edges() const { uint64_t e; detail::check(sb200_push_graph_info(h_, nullptr, &e)); return e; }
    Precision out_degree(size_t node) const { double d; detail::check(sb200_push_graph_degrees(h_, node, &d, nullptr)); return d; }
    Precision in_degree(size_t node) const { double d; detail::check(sb200_push_graph_degrees(h_, node, nullptr, &d)); return d; }
    const sb200_push_graph *handle() const { return h_; }

private:
    explicit PushGraph(sb200_push_graph *h) : h_(h) {}
    sb200_push_graph *h_ = nullptr;
};

struct PushConfig {  // ForwardPushConfig / BackwardPushConfig (forward_push.rs:25-50)
    Precision alpha = 0.15, epsilon = 1e-6;
    size_t max_pushes = 1000000;
    Precision queue_threshold = 1e-8;
    bool adaptive_threshold = true;
    sb200_push_config to_c() const {
        sb200_push_config c;
        sb200_push_config_default(&c);
        c.alpha = alpha; c.epsilon = epsilon; c.max_pushes = max_pushes; c.queue_threshold = queue_threshold;
        c.adaptive_threshold = adaptive_threshold;
        return c;
    }
};

struct PushResult {  // ForwardPushResult / BackwardPushResult (forward_push.rs:10-22)
    std::vector<Precision> estimate, residual;
    size_t push_count = 0, nodes_visited = 0;
    Precision residual_norm = 0.0;
};

namespace detail {
template <typename Fn>
inline PushResult run_push(Fn fn, const PushGraph &g, const PushConfig &cfg, const std::vector<size_t> &seeds) {
    const sb200_push_config c = cfg.to_c();
    std::vector<uint64_t> s(seeds.begin(), seeds.end());
    PushResult r;
    r.estimate.assign(g.num_nodes(), 0.0);
    r.residual.assign(g.num_nodes(), 0.0);
    sb200_push_stats st;
    check(fn(g.handle(), &c, s.data(), s.size(), r.estimate.data(), r.residual.data(), &st));
    r.push_count = st.push_count; r.nodes_visited = st.nodes_visited; r.residual_norm = st.residual_norm;
    return r;
}
template <typename Fn>
inline PushResult run_push_watch(Fn fn, const PushGraph &g, const PushConfig &cfg, size_t source, size_t target,
                                 Precision precision) {
    const sb200_push_config c = cfg.to_c();
    PushResult r;
    r.estimate.assign(g.num_nodes(), 0.0);
    r.residual.assign(g.num_nodes(), 0.0);
    sb200_push_stats st;
    check(fn(g.handle(), &c, source, target, precision, r.estimate.data(), r.residual.data(), &st));
    r.push_count = st.push_count; r.nodes_visited = st.nodes_visited; r.residual_norm = st.residual_norm;
    return r;
}
}  // namespace detail

class ForwardPushSolver {  // forward_push.rs:52-328
public:
    ForwardPushSolver(PushGraph graph, PushConfig config = {}) : graph_(std::move(graph)), config_(config) {}
    PushResult solve_single_source(size_t source) const { return detail::run_push(sb200_forward_push, graph_, config_, {source}); }
    PushResult solve_multi_source(const std::vector<size_t> &sources) const { return detail::run_push(sb200_forward_push, graph_, config_, sources); }
    Precision query_single_entry(size_t source, size_t target) const {
        const PushResult r = solve_single_source(source);
        return target < r.estimate.size() ? r.estimate[target] : 0.0;
    }
    std::vector<Precision> extrapolated_solution(const PushResult &result) const {  // :317-327
        std::vector<Precision> x = result.estimate;
        for (size_t i = 0; i < x.size(); i++) x[i] += config_.alpha * result.residual[i];
        return x;
    }
    PushResult solve_with_target(size_t source, size_t target, Precision target_precision) const {  // :234-290
        return detail::run_push_watch(sb200_forward_push_with_target, graph_, config_, source, target, target_precision);
    }
    const PushGraph &graph() const { return graph_; }
    const PushConfig &config() const { return config_; }

private:
    PushGraph graph_;
    PushConfig config_;
};

class BackwardPushSolver {  // backward_push.rs:52-330
public:
    BackwardPushSolver(PushGraph graph, PushConfig config = {}) : graph_(std::move(graph)), config_(config) {}
    PushResult solve_single_target(size_t target) const { return detail::run_push(sb200_backward_push, graph_, config_, {target}); }
    PushResult solve_multi_target(const std::vector<size_t> &targets) const { return detail::run_push(sb200_backward_push, graph_, config_, targets); }
    Precision query_transition_probability(size_t source, size_t target) const {
        const PushResult r = solve_single_target(target);
        return source < r.estimate.size() ? r.estimate[source] : 0.0;
    }
    PushResult solve_with_source(size_t source, size_t target, Precision source_precision) const {  // :238-290
        return detail::run_push_watch(sb200_backward_push_with_source, graph_, config_, source, target, source_precision);
    }
    std::vector<Precision> extrapolated_solution(const PushResult &result) const {  // :300-309
        std::vector<Precision> x = result.estimate;
        for (size_t i = 0; i < x.size(); i++) x[i] += config_.alpha * result.residual[i];
        return x;
    }
    std::vector<Precision> reachability_probabilities(size_t target) const {  // :294-297
        return extrapolated_solution(solve_single_target(target));
    }
    Precision combine_with_forward(const PushResult &backward_result, const std::vector<Precision> &forward_estimate,
                                   const std::vector<Precision> &forward_residual) const {  // :312-330
        double out = 0.0;
        detail::check(sb200_push_combine_with_forward(config_.alpha, backward_result.estimate.data(), backward_result.residual.data(),
                                                      backward_result.estimate.size(), forward_estimate.data(),
                                                      forward_residual.data(), forward_estimate.size(), &out));
        return out;
    }
    const PushGraph &graph() const { return graph_; }

private:
    PushGraph graph_;
    PushConfig config_;
};

// BidirectionalPushSolver (backward_push.rs:338-420). The reference clones the graph into two solvers; here both
// directions run on the one device-resident PushGraph.
class BidirectionalPushSolver {
public:
    BidirectionalPushSolver(PushGraph graph, PushConfig forward_config = {}, PushConfig backward_config = {})
        : graph_(std::move(graph)), forward_(forward_config), backward_(backward_config) {}
    Precision solve_bidirectional(size_t source, size_t target) const {
        const sb200_push_config f = forward_.to_c(), b = backward_.to_c();
        double out = 0.0;
        detail::check(sb200_bidirectional_push(graph_.handle(), &f, &b, source, target, &out));
        return out;
    }
    Precision adaptive_solve(size_t source, size_t target) const {
        const sb200_push_config f = forward_.to_c(), b = backward_.to_c();
        double out = 0.0;
        detail::check(sb200_bidirectional_adaptive_push(graph_.handle(), &f, &b, source, target, &out));
        return out;
    }

private:
    PushGraph graph_;
    PushConfig forward_, backward_;
};

// SublinearSolver.solveForwardPush of the TS package (src/core/solver.ts:437-522): A x = b by residual pushes
struct ForwardPushSolveResult {
    std::vector<Precision> solution;
    size_t iterations = 0;
    Precision residual = 0.0;
    bool converged = false;
};
inline ForwardPushSolveResult forward_push_solve(const SparseMatrix &matrix, const std::vector<Precision> &b,
                                                 Precision epsilon = 1e-6, size_t max_iterations = 1000) {
    ForwardPushSolveResult r;
    r.solution.assign(b.size(), 0.0);
    sb200_axb_push_stats st;
    detail::check(sb200_forward_push_solve(matrix.handle(), b.data(), b.size(), epsilon, max_iterations, r.solution.data(), &st));
    r.iterations = st.iterations;
    r.residual = st.residual_norm;
    r.converged = st.converged != 0;
    return r;
}

// StreamingMatrix (src/matrix/optimized.rs:451-561): row chunks in pinned host memory, streamed through the GPU
class StreamingMatrix {
public:
    static StreamingMatrix from_triplets(const std::vector<std::tuple<size_t, size_t, Precision>> &triplets, size_t rows,
                                         size_t cols, size_t memory_limit_mb) {
        std::vector<uint64_t> r(triplets.size()), c(triplets.size());
        std::vector<double> v(triplets.size());
        for (size_t i = 0; i < triplets.size(); i++) { r[i] = std::get<0>(triplets[i]); c[i] = std::get<1>(triplets[i]); v[i] = std::get<2>(triplets[i]); }
        sb200_streaming_matrix *h = nullptr;
        detail::check(sb200_streaming_matrix_from_triplets(r.data(), c.data(), v.data(), v.size(), rows, cols, memory_limit_mb, &h));
        return StreamingMatrix(h);
    }
    StreamingMatrix(StreamingMatrix &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    StreamingMatrix(const StreamingMatrix &) = delete;
    ~StreamingMatrix() { sb200_streaming_matrix_free(h_); }
    // multiply_vector_streaming(x, |start_row, result| ..)
    template <typename F>
    void multiply_vector_streaming(const std::vector<Precision> &x, F callback) const {
        auto tramp = [](uint64_t start, const double *res, uint64_t len, void *user) -> int32_t {
            (*static_cast<F *>(user))(static_cast<size_t>(start), res, static_cast<size_t>(len));
            return 0;
        };
        detail::check(sb200_streaming_matrix_multiply_vector(h_, x.data(), x.size(), tramp, &callback));
    }
    size_t memory_usage() const {
        uint64_t b = 0;
        detail::check(sb200_streaming_matrix_info(h_, nullptr, nullptr, nullptr, nullptr, &b));
        return b;
    }
    size_t num_chunks() const {
        uint64_t k = 0;
        detail::check(sb200_streaming_matrix_info(h_, nullptr, nullptr, nullptr, &k, nullptr));
        return k;
    }

private:
    explicit StreamingMatrix(sb200_streaming_matrix *h) : h_(h) {}
    sb200_streaming_matrix *h_ = nullptr;
};

}  // namespace sublinear
