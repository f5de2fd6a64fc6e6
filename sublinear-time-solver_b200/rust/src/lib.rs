//! Rust shim over include/sublinear_b200.h — UNBUILT here (no Rust toolchain in the build image); it is the 1:1
//! mapping a maintainer would add to route `NeumannSolver::solve` / `SparseMatrix` to the B200 library.
//! Types and traits come from the reference crate (`sublinear`): src/matrix/mod.rs:25-104 (trait Matrix),
//! src/solver/mod.rs:223-333 (trait SolverAlgorithm), src/solver/mod.rs:22-195, src/error.rs:16-138.
#![allow(non_camel_case_types)]
use std::os::raw::c_char;

use sublinear::error::{Result, SolverError};
use sublinear::solver::{SolverOptions, SolverResult, SolverState, StepResult};
use sublinear::types::{ErrorBoundMethod, ErrorBounds, MemoryInfo, Precision, SolverStats};

#[repr(C)] pub struct sb200_matrix { _p: [u8; 0] }
#[repr(C)] pub struct sb200_solver { _p: [u8; 0] }
#[repr(C)] pub struct sb200_state { _p: [u8; 0] }
#[repr(C)] pub struct sb200_push_graph { _p: [u8; 0] }

#[repr(C)]
pub struct sb200_push_config {
    pub alpha: f64, pub epsilon: f64, pub max_pushes: u64, pub queue_threshold: f64, pub adaptive_threshold: i32,
    pub reserved: i32,
}

#[repr(C)]
pub struct sb200_push_stats {
    pub push_count: u64, pub nodes_visited: u64, pub residual_norm: f64, pub rounds: u64, pub kernel_launches: u64,
    pub device_time_ms: f64, pub dense_rounds: u64, pub edges_touched: u64,
}

#[repr(C)]
pub struct sb200_axb_push_stats {
    pub iterations: u64, pub rounds: u64, pub residual_norm: f64, pub max_residual: f64, pub converged: i32, pub reserved: i32,
}

#[repr(C)] pub struct sb200_streaming_matrix { _private: [u8; 0] }

#[repr(C)]
pub struct sb200_state_info_t {
    pub dimension: u64, pub residual_norm: f64, pub matvec_count: u64, pub terms_computed: u64,
    pub series_converged: i32, pub last_term_norm: f64, pub has_error_bounds: i32, pub error_upper_bound: f64,
    pub memory_bytes: u64,
}

#[repr(C)]
pub struct sb200_cg_config { pub max_iterations: u64, pub tolerance: f64, pub enable_profiling: i32, pub reserved: i32 }

#[repr(C)]
pub struct sb200_cg_result {
    pub solution: *mut f64, pub solution_len: u64, pub residual_norm: f64, pub iterations: u64, pub converged: i32,
    pub breakdown: i32, pub computation_time_ms: f64, pub matvec_count: u64, pub dot_product_count: u64,
    pub axpy_count: u64, pub total_flops: u64, pub average_bandwidth_gbs: f64, pub average_gflops: f64,
    pub device_time_ms: f64, pub kernel_launches: u64, pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub spmv_kernel_ms: f64, pub spmv_kernel_count: u64,
}

#[repr(C)]
pub struct sb200_options {
    pub tolerance: f64, pub max_iterations: u64, pub convergence_mode: i32, pub norm_type: i32,
    pub collect_stats: i32, pub streaming_interval: u64, pub initial_guess: *const f64, pub initial_guess_len: u64,
    pub compute_error_bounds: i32, pub error_bounds_tolerance: f64, pub enable_profiling: i32,
    pub has_random_seed: i32, pub random_seed: u64,
    pub mode: i32, pub dominance: i32, pub residual_check: i32, pub reserved: i32,
}

#[repr(C)]
pub struct sb200_result {
    pub solution: *mut f64, pub solution_len: u64, pub residual_norm: f64, pub iterations: u64, pub converged: i32,
    pub has_error_bounds: i32, pub error_upper_bound: f64, pub has_stats: i32, pub total_time_ms: f64,
    pub matvec_count: u64, pub memory_bytes: u64, pub terms_computed: u64, pub series_converged: i32,
    pub last_term_norm: f64, pub device_time_ms: f64, pub kernel_launches: u64, pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub push_kernel_ms: f64, pub push_kernel_count: u64, pub resid_kernel_ms: f64, pub resid_kernel_count: u64,
}

extern "C" {
    fn sb200_last_error(buf: *mut c_char, cap: usize) -> usize;
    fn sb200_matrix_from_triplets(rows: *const u64, cols: *const u64, vals: *const f64, n: u64, nrows: u64, ncols: u64,
                                  out: *mut *mut sb200_matrix) -> i32;
    fn sb200_matrix_free(m: *mut sb200_matrix);
    fn sb200_matrix_multiply_vector(m: *const sb200_matrix, x: *const f64, xlen: u64, y: *mut f64, ylen: u64) -> i32;
    fn sb200_neumann_new(max_terms: u64, series_tolerance: f64, out: *mut *mut sb200_solver) -> i32;
    fn sb200_solver_free(s: *mut sb200_solver);
    fn sb200_options_default(o: *mut sb200_options);
    fn sb200_solve_into(s: *const sb200_solver, m: *const sb200_matrix, b: *const f64, blen: u64,
                        opt: *const sb200_options, x_out: *mut f64, out: *mut sb200_result) -> i32;
    // trait SolverAlgorithm / SolverState (src/solver/mod.rs:223-352)
    fn sb200_neumann_initialize(s: *const sb200_solver, m: *const sb200_matrix, b: *const f64, blen: u64,
                                opt: *const sb200_options, out: *mut *mut sb200_state) -> i32;
    fn sb200_state_step(st: *mut sb200_state, step_result: *mut i32) -> i32;
    fn sb200_state_is_converged(st: *const sb200_state, out: *mut i32) -> i32;
    fn sb200_state_extract_solution(st: *const sb200_state, x: *mut f64, xlen: u64) -> i32;
    fn sb200_state_update_rhs(st: *mut sb200_state, indices: *const u64, deltas: *const f64, count: u64) -> i32;
    fn sb200_state_reset(st: *mut sb200_state) -> i32;
    fn sb200_state_info(st: *const sb200_state, info: *mut sb200_state_info_t) -> i32;
    fn sb200_state_free(st: *mut sb200_state);
    // PushGraph / ForwardPushSolver / BackwardPushSolver (src/graph/adjacency.rs:199-277, src/solver/forward_push.rs,
    // src/solver/backward_push.rs)
    fn sb200_push_config_default(c: *mut sb200_push_config);
    fn sb200_push_graph_from_csr(row_ptr: *const u64, col_indices: *const u32, weights: *const f64, n: u64,
                                 out: *mut *mut sb200_push_graph) -> i32;
    fn sb200_push_graph_free(g: *mut sb200_push_graph);
    fn sb200_forward_push(g: *const sb200_push_graph, cfg: *const sb200_push_config, sources: *const u64, nsources: u64,
                          estimate: *mut f64, residual: *mut f64, stats: *mut sb200_push_stats) -> i32;
    fn sb200_backward_push(g: *const sb200_push_graph, cfg: *const sb200_push_config, targets: *const u64, ntargets: u64,
                           estimate: *mut f64, residual: *mut f64, stats: *mut sb200_push_stats) -> i32;
    // solve_with_target / solve_with_source / combine_with_forward / BidirectionalPushSolver
    // (src/solver/forward_push.rs:234-290, backward_push.rs:238-420)
    fn sb200_forward_push_with_target(g: *const sb200_push_graph, cfg: *const sb200_push_config, source: u64, target: u64,
                                      target_precision: f64, estimate: *mut f64, residual: *mut f64,
                                      stats: *mut sb200_push_stats) -> i32;
    fn sb200_backward_push_with_source(g: *const sb200_push_graph, cfg: *const sb200_push_config, source: u64, target: u64,
                                       source_precision: f64, estimate: *mut f64, residual: *mut f64,
                                       stats: *mut sb200_push_stats) -> i32;
    fn sb200_push_combine_with_forward(alpha: f64, backward_estimate: *const f64, backward_residual: *const f64, nbackward: u64,
                                       forward_estimate: *const f64, forward_residual: *const f64, nforward: u64,
                                       out: *mut f64) -> i32;
    fn sb200_bidirectional_push(g: *const sb200_push_graph, forward_cfg: *const sb200_push_config,
                                backward_cfg: *const sb200_push_config, source: u64, target: u64, out: *mut f64) -> i32;
    fn sb200_bidirectional_adaptive_push(g: *const sb200_push_graph, forward_cfg: *const sb200_push_config,
                                         backward_cfg: *const sb200_push_config, source: u64, target: u64, out: *mut f64) -> i32;
    // SublinearSolver.solveForwardPush of the TS package (src/core/solver.ts:437-522)
    fn sb200_forward_push_solve(m: *const sb200_matrix, b: *const f64, blen: u64, epsilon: f64, max_iterations: u64,
                                x_out: *mut f64, stats: *mut sb200_axb_push_stats) -> i32;
    // StreamingMatrix (src/matrix/optimized.rs:451-561)
    fn sb200_streaming_matrix_from_triplets(rows: *const u64, cols: *const u64, vals: *const f64, ntriplets: u64, nrows: u64,
                                            ncols: u64, memory_limit_mb: u64, out: *mut *mut sb200_streaming_matrix) -> i32;
    fn sb200_streaming_matrix_info(sm: *const sb200_streaming_matrix, total_rows: *mut u64, total_cols: *mut u64,
                                   chunk_size: *mut u64, num_chunks: *mut u64, memory_usage: *mut u64) -> i32;
    fn sb200_streaming_matrix_multiply_vector(sm: *const sb200_streaming_matrix, x: *const f64, xlen: u64,
                                              callback: extern "C" fn(u64, *const f64, u64, *mut c_void) -> i32,
                                              user: *mut c_void) -> i32;
    fn sb200_streaming_matrix_free(sm: *mut sb200_streaming_matrix);
    // OptimizedConjugateGradientSolver (src/optimized_solver.rs:168-295)
    fn sb200_cg_config_default(c: *mut sb200_cg_config);
    fn sb200_cg_solve_into(m: *const sb200_matrix, b: *const f64, blen: u64, cfg: *const sb200_cg_config,
                           x_out: *mut f64, out: *mut sb200_cg_result) -> i32;
}

fn last_error() -> String {
    let mut buf = vec![0u8; 1024];
    unsafe { sb200_last_error(buf.as_mut_ptr() as *mut c_char, buf.len()) };
    String::from_utf8_lossy(&buf).trim_end_matches('\0').to_string()
}

/// Status code (1-based variant index of `SolverError`, src/error.rs:16-138) -> the Rust error value.
fn to_error(code: i32, r: &sb200_result, tolerance: f64) -> SolverError {
    let message = last_error();
    match code {
        1 => SolverError::MatrixNotDiagonallyDominant { row: 0, diagonal: 0.0, off_diagonal_sum: 0.0 }, // neumann.rs:164-168
        2 => SolverError::NumericalInstability { reason: message, iteration: r.iterations as usize, residual_norm: r.residual_norm },
        3 => SolverError::ConvergenceFailure { iterations: r.iterations as usize, residual_norm: r.residual_norm, tolerance,
                                               algorithm: "neumann".to_string() },
        5 => SolverError::DimensionMismatch { expected: 0, actual: 0, operation: message },
        8 => SolverError::IndexOutOfBounds { index: 0, max_index: 0, context: message },
        9 => SolverError::InvalidSparseMatrix { reason: message, position: None },
        4 => SolverError::InvalidInput { message, parameter: None },
        _ => SolverError::AlgorithmError { algorithm: "neumann".to_string(), message, context: vec![] },
    }
}

/// `SparseMatrix` resident on the B200 (same constructor signature as src/matrix/mod.rs:160-164).
pub struct B200Matrix { h: *mut sb200_matrix, rows: usize, cols: usize }
unsafe impl Send for B200Matrix {}
unsafe impl Sync for B200Matrix {}

impl B200Matrix {
    pub fn from_triplets(triplets: Vec<(usize, usize, Precision)>, rows: usize, cols: usize) -> Result<Self> {
        let r: Vec<u64> = triplets.iter().map(|t| t.0 as u64).collect();
        let c: Vec<u64> = triplets.iter().map(|t| t.1 as u64).collect();
        let v: Vec<f64> = triplets.iter().map(|t| t.2).collect();
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sb200_matrix_from_triplets(r.as_ptr(), c.as_ptr(), v.as_ptr(), v.len() as u64, rows as u64, cols as u64, &mut h) };
        if rc != 0 { return Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)); }
        Ok(Self { h, rows, cols })
    }
    /// `Matrix::multiply_vector` (src/matrix/mod.rs:415-439)
    pub fn multiply_vector(&self, x: &[Precision], result: &mut [Precision]) -> Result<()> {
        let rc = unsafe { sb200_matrix_multiply_vector(self.h, x.as_ptr(), x.len() as u64, result.as_mut_ptr(), result.len() as u64) };
        if rc != 0 { Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)) } else { Ok(()) }
    }
    pub fn rows(&self) -> usize { self.rows }
    pub fn cols(&self) -> usize { self.cols }
}
impl Drop for B200Matrix { fn drop(&mut self) { unsafe { sb200_matrix_free(self.h) } } }

/// `NeumannSolver` whose `solve` runs on the B200 (src/solver/neumann.rs:469-555).
pub struct B200NeumannSolver { h: *mut sb200_solver }
unsafe impl Send for B200NeumannSolver {}
unsafe impl Sync for B200NeumannSolver {}

impl B200NeumannSolver {
    pub fn new(max_terms: usize, series_tolerance: Precision) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { sb200_neumann_new(max_terms as u64, series_tolerance, &mut h) };
        Self { h }
    }
    pub fn default() -> Self { Self::new(50, 1e-8) }

    pub fn solve(&self, matrix: &B200Matrix, b: &[Precision], options: &SolverOptions) -> Result<SolverResult> {
        let mut o: sb200_options = unsafe { std::mem::zeroed() };
        unsafe { sb200_options_default(&mut o) };
        o.tolerance = options.tolerance;
        o.max_iterations = options.max_iterations as u64;
        o.collect_stats = options.collect_stats as i32;
        o.compute_error_bounds = options.compute_error_bounds as i32;
        if let Some(ref g) = options.initial_guess { o.initial_guess = g.as_ptr(); o.initial_guess_len = g.len() as u64; }
        let mut x = vec![0.0; b.len()];
        let mut r: sb200_result = unsafe { std::mem::zeroed() };
        let rc = unsafe { sb200_solve_into(self.h, matrix.h, b.as_ptr(), b.len() as u64, &o, x.as_mut_ptr(), &mut r) };
        if rc != 0 { return Err(to_error(rc, &r, options.tolerance)); }
        let mut res = if r.converged != 0 { SolverResult::success(x, r.residual_norm, r.iterations as usize) }
                      else { SolverResult::failure(x, r.residual_norm, r.iterations as usize) };
        if r.has_stats != 0 {
            let mut s = SolverStats::new();
            s.total_time_ms = r.total_time_ms;
            s.matvec_count = r.matvec_count as usize;
            res.stats = Some(s);
        }
        Ok(res)
    }
}
impl Drop for B200NeumannSolver { fn drop(&mut self) { unsafe { sb200_solver_free(self.h) } } }

/// `NeumannState` on the device (src/solver/neumann.rs:97-135). The C state shares ownership of the matrix handle, so
/// it may outlive the `B200Matrix` it was initialised from.
pub struct B200NeumannState { h: *mut sb200_state }
unsafe impl Send for B200NeumannState {}

impl B200NeumannState {
    fn info(&self) -> sb200_state_info_t {
        let mut i: sb200_state_info_t = unsafe { std::mem::zeroed() };
        unsafe { sb200_state_info(self.h, &mut i) };
        i
    }
}
impl Drop for B200NeumannState { fn drop(&mut self) { unsafe { sb200_state_free(self.h) } } }

/// trait SolverState (src/solver/mod.rs:336-352; NeumannState impl src/solver/neumann.rs:350-378)
impl SolverState for B200NeumannState {
    fn residual_norm(&self) -> Precision { self.info().residual_norm }
    fn matvec_count(&self) -> usize { self.info().matvec_count as usize }
    fn error_bounds(&self) -> Option<ErrorBounds> {
        let i = self.info();
        if i.has_error_bounds != 0 { Some(ErrorBounds::upper_bound_only(i.error_upper_bound, ErrorBoundMethod::NeumannTruncation)) } else { None }
    }
    fn memory_usage(&self) -> MemoryInfo {
        let b = self.info().memory_bytes as usize;
        MemoryInfo { current_usage_bytes: b, peak_usage_bytes: b, matrix_memory_bytes: 0, vector_memory_bytes: b,
                     workspace_memory_bytes: 0, allocation_count: 1, deallocation_count: 0 }
    }
    fn reset(&mut self) { unsafe { sb200_state_reset(self.h) }; }
}

/// The stepping half of trait SolverAlgorithm (src/solver/mod.rs:223-252) for the B200 solver. `initialize` takes the
/// device-resident matrix; a blanket `impl SolverAlgorithm` would first convert `&dyn Matrix` with `to_triplets()`.
impl B200NeumannSolver {
    pub fn initialize(&self, matrix: &B200Matrix, b: &[Precision], options: &SolverOptions) -> Result<B200NeumannState> {
        let mut o: sb200_options = unsafe { std::mem::zeroed() };
        unsafe { sb200_options_default(&mut o) };
        o.tolerance = options.tolerance;
        o.max_iterations = options.max_iterations as u64;
        if let Some(ref g) = options.initial_guess { o.initial_guess = g.as_ptr(); o.initial_guess_len = g.len() as u64; }
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sb200_neumann_initialize(self.h, matrix.h, b.as_ptr(), b.len() as u64, &o, &mut h) };
        if rc != 0 { return Err(to_error(rc, &unsafe { std::mem::zeroed() }, options.tolerance)); }
        Ok(B200NeumannState { h })
    }
    /// The body the reference left commented out (src/solver/neumann.rs:404-418).
    pub fn step(&self, state: &mut B200NeumannState) -> Result<StepResult> {
        let mut r = 0i32;
        let rc = unsafe { sb200_state_step(state.h, &mut r) };
        if rc != 0 { return Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)); }
        Ok(if r == 1 { StepResult::Converged } else { StepResult::Continue })
    }
    pub fn is_converged(&self, state: &B200NeumannState) -> bool {
        let mut c = 0i32;
        unsafe { sb200_state_is_converged(state.h, &mut c) };
        c != 0
    }
    pub fn extract_solution(&self, state: &B200NeumannState) -> Vec<Precision> {
        let mut x = vec![0.0; state.info().dimension as usize];
        unsafe { sb200_state_extract_solution(state.h, x.as_mut_ptr(), x.len() as u64) };
        x
    }
    pub fn update_rhs(&self, state: &mut B200NeumannState, delta_b: &[(usize, Precision)]) -> Result<()> {
        let idx: Vec<u64> = delta_b.iter().map(|d| d.0 as u64).collect();
        let dl: Vec<f64> = delta_b.iter().map(|d| d.1).collect();
        let rc = unsafe { sb200_state_update_rhs(state.h, idx.as_ptr(), dl.as_ptr(), dl.len() as u64) };
        if rc != 0 { Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)) } else { Ok(()) }
    }
    pub fn algorithm_name(&self) -> &'static str { "neumann" }
}

/// `OptimizedConjugateGradientSolver::solve` (src/optimized_solver.rs:182-295) on the same SpMV kernel.
/// Returns (solution, residual_norm, iterations, converged) — the fields of `OptimizedSolverResult`.
pub fn b200_cg_solve(matrix: &B200Matrix, b: &[Precision], max_iterations: usize, tolerance: Precision)
                     -> std::result::Result<(Vec<Precision>, Precision, usize, bool), String> {
    let mut c: sb200_cg_config = unsafe { std::mem::zeroed() };
    unsafe { sb200_cg_config_default(&mut c) };
    c.max_iterations = max_iterations as u64;
    c.tolerance = tolerance;
    let mut x = vec![0.0; b.len()];
    let mut r: sb200_cg_result = unsafe { std::mem::zeroed() };
    let rc = unsafe { sb200_cg_solve_into(matrix.h, b.as_ptr(), b.len() as u64, &c, x.as_mut_ptr(), &mut r) };
    if rc != 0 { return Err(last_error()); }   // "Matrix must be square" / length mismatch (:188-193)
    Ok((x, r.residual_norm, r.iterations as usize, r.converged != 0))
}

/// `PushGraph` + `ForwardPushSolver::solve_single_source` / `BackwardPushSolver::solve_single_target` on the device
/// (src/solver/forward_push.rs:66-121, src/solver/backward_push.rs:66-121). Returns the fields of
/// `ForwardPushResult`: (estimate, residual, push_count, nodes_visited, residual_norm).
pub struct B200PushGraph { h: *mut sb200_push_graph, n: usize }
unsafe impl Send for B200PushGraph {}
unsafe impl Sync for B200PushGraph {}

impl B200PushGraph {
    /// `PushGraph::from_matrix(&CompressedSparseRow)` (src/graph/adjacency.rs:211-224)
    pub fn from_matrix(row_ptr: &[usize], col_indices: &[usize], values: &[f64]) -> Result<Self> {
        let rp: Vec<u64> = row_ptr.iter().map(|&v| v as u64).collect();
        let ci: Vec<u32> = col_indices.iter().map(|&v| v as u32).collect();
        let mut h = std::ptr::null_mut();
        let n = row_ptr.len() - 1;
        let rc = unsafe { sb200_push_graph_from_csr(rp.as_ptr(), ci.as_ptr(), values.as_ptr(), n as u64, &mut h) };
        if rc != 0 { return Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)); }
        Ok(Self { h, n })
    }
    fn run(&self, backward: bool, alpha: f64, epsilon: f64, seed: usize) -> Result<(Vec<f64>, Vec<f64>, usize, usize, f64)> {
        let mut c: sb200_push_config = unsafe { std::mem::zeroed() };
        unsafe { sb200_push_config_default(&mut c) };
        c.alpha = alpha;
        c.epsilon = epsilon;
        let (mut est, mut res) = (vec![0.0; self.n], vec![0.0; self.n]);
        let mut st: sb200_push_stats = unsafe { std::mem::zeroed() };
        let s = [seed as u64];
        let rc = unsafe {
            if backward { sb200_backward_push(self.h, &c, s.as_ptr(), 1, est.as_mut_ptr(), res.as_mut_ptr(), &mut st) }
            else { sb200_forward_push(self.h, &c, s.as_ptr(), 1, est.as_mut_ptr(), res.as_mut_ptr(), &mut st) }
        };
        if rc != 0 { return Err(to_error(rc, &unsafe { std::mem::zeroed() }, 0.0)); }
        Ok((est, res, st.push_count as usize, st.nodes_visited as usize, st.residual_norm))
    }
    pub fn forward_push_single_source(&self, alpha: f64, epsilon: f64, source: usize) -> Result<(Vec<f64>, Vec<f64>, usize, usize, f64)> {
        self.run(false, alpha, epsilon, source)
    }
    pub fn backward_push_single_target(&self, alpha: f64, epsilon: f64, target: usize) -> Result<(Vec<f64>, Vec<f64>, usize, usize, f64)> {
        self.run(true, alpha, epsilon, target)
    }
}
impl Drop for B200PushGraph { fn drop(&mut self) { unsafe { sb200_push_graph_free(self.h) } } }
