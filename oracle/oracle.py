"""ctypes view of the CPU ORACLE (oracle/sublinear_oracle.c) — test infrastructure, NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates the reference's Rust CPU path (see sublinear_oracle.h for the citations);
the reference itself (Rust/TypeScript) cannot be built in this image, so there is no oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OK = 0
ERR_NOT_DIAGONALLY_DOMINANT = 1
ERR_NUMERICAL_INSTABILITY = 2
ERR_CONVERGENCE_FAILURE = 3
ERR_INVALID_INPUT = 4
ERR_DIMENSION_MISMATCH = 5
ERR_INDEX_OUT_OF_BOUNDS = 8
ERR_INVALID_SPARSE_MATRIX = 9

MODE_CORRECT, MODE_REF_COMPAT = 0, 1
DOM_ROW, DOM_ROW_OR_COL = 0, 1
SPMV_SCALAR, SPMV_SIMD4, SPMV_PARALLEL = 0, 1, 2
DOT_SEQUENTIAL, DOT_CHUNK4, DOT_CHUNK8 = 0, 1, 2


class _Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint64), ("ncols", C.c_uint64), ("nnz", C.c_uint64),
                ("values", C.POINTER(C.c_double)), ("col_indices", C.POINTER(C.c_uint32)),
                ("row_ptr", C.POINTER(C.c_uint32))]


class _Options(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("max_iterations", C.c_uint64),
                ("initial_guess", C.POINTER(C.c_double)), ("initial_guess_len", C.c_uint64),
                ("compute_error_bounds", C.c_int), ("max_terms", C.c_uint64),
                ("series_tolerance", C.c_double), ("adaptive_truncation", C.c_int),
                ("mode", C.c_int), ("dominance", C.c_int), ("spmv_variant", C.c_int),
                ("nthreads", C.c_int)]


class _Result(C.Structure):
    _fields_ = [("solution", C.POINTER(C.c_double)), ("residual_norm", C.c_double),
                ("iterations", C.c_uint64), ("terms_computed", C.c_uint64),
                ("matvec_count", C.c_uint64), ("converged", C.c_int), ("series_converged", C.c_int),
                ("has_error_bound", C.c_int), ("error_bound", C.c_double),
                ("last_term_norm", C.c_double), ("total_time_ms", C.c_double)]


class _PushConfig(C.Structure):
    _fields_ = [("alpha", C.c_double), ("epsilon", C.c_double), ("max_pushes", C.c_uint64),
                ("queue_threshold", C.c_double), ("adaptive_threshold", C.c_int)]


class _PushStats(C.Structure):
    _fields_ = [("push_count", C.c_uint64), ("nodes_visited", C.c_uint64), ("residual_norm", C.c_double)]


class _CgResult(C.Structure):
    _fields_ = [("solution", C.POINTER(C.c_double)), ("residual_norm", C.c_double), ("iterations", C.c_uint64),
                ("converged", C.c_int), ("matvec_count", C.c_uint64), ("total_flops", C.c_uint64)]


def build(fast: bool = False, out_dir: str | None = None) -> str:
    """Compile the oracle with the system gcc (the image's $CC wrapper lacks libgomp)."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    out = os.path.join(out_dir or _HERE, name)
    src = os.path.join(_HERE, "sublinear_oracle.c")
    hdr = os.path.join(_HERE, "sublinear_oracle.h")
    if os.path.exists(out) and out_dir is None and \
            os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return out
    opt = ["-O3", "-march=native"] if fast else ["-O2"]
    cmd = ["/usr/bin/gcc", *opt, "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-o", out, src, "-lm"]
    subprocess.run(cmd, check=True)
    return out


_libs: dict = {}


def lib(fast: bool = False, out_dir: str | None = None):
    key = (fast, out_dir)
    if key in _libs:
        return _libs[key]
    L = C.CDLL(build(fast, out_dir))
    u64p, f64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    L.orc_csr_free.argtypes = [C.POINTER(_Csr)]
    L.orc_csr_free.restype = None
    L.orc_csr_from_triplets.argtypes = [u64p, u64p, f64p, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_Csr)]
    L.orc_csr_get.argtypes = [C.POINTER(_Csr), C.c_uint64, C.c_uint64, f64p]
    for f in (L.orc_spmv_scalar, L.orc_spmv_simd4):
        f.argtypes = [C.POINTER(_Csr), f64p, f64p]
        f.restype = None
    L.orc_spmv_parallel.argtypes = [C.POINTER(_Csr), f64p, f64p, C.c_int]
    L.orc_spmv_parallel.restype = None
    L.orc_multiply_vector.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, f64p, C.c_uint64, C.c_int, C.c_int]
    L.orc_is_diagonally_dominant.argtypes = [C.POINTER(_Csr), u64p]
    L.orc_is_col_diagonally_dominant.argtypes = [C.POINTER(_Csr)]
    for f in (L.orc_l2_norm, L.orc_l1_norm, L.orc_linf_norm):
        f.argtypes = [f64p, C.c_uint64]
        f.restype = C.c_double
    L.orc_dot_simd4.argtypes = [f64p, f64p, C.c_uint64]
    L.orc_dot_simd4.restype = C.c_double
    L.orc_axpy_simd4.argtypes = [C.c_double, f64p, f64p, C.c_uint64]
    L.orc_axpy_simd4.restype = None
    L.orc_options_default.argtypes = [C.POINTER(_Options)]
    L.orc_options_default.restype = None
    L.orc_neumann_solve.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, C.POINTER(_Options), C.POINTER(_Result)]
    L.orc_push_iterations.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, C.c_int, C.c_int, f64p, f64p, f64p]
    L.orc_push_iterations.restype = C.c_double
    L.orc_state_new.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, C.POINTER(_Options), C.POINTER(C.c_void_p)]
    L.orc_state_step.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.orc_state_is_converged.argtypes = [C.c_void_p]
    L.orc_state_solution.argtypes = [C.c_void_p, f64p]
    L.orc_state_solution.restype = None
    L.orc_state_update_rhs.argtypes = [C.c_void_p, u64p, f64p, C.c_uint64]
    L.orc_state_reset.argtypes = [C.c_void_p]
    L.orc_state_reset.restype = None
    L.orc_state_info.argtypes = [C.c_void_p, f64p, u64p, u64p, C.POINTER(C.c_int), f64p, C.POINTER(C.c_int), f64p]
    L.orc_state_info.restype = None
    L.orc_state_free.argtypes = [C.c_void_p]
    L.orc_state_free.restype = None
    L.orc_push_config_default.argtypes = [C.POINTER(_PushConfig)]
    L.orc_push_config_default.restype = None
    for f in (L.orc_forward_push, L.orc_backward_push):
        f.argtypes = [C.POINTER(_Csr), C.POINTER(_PushConfig), u64p, C.c_uint64, f64p, f64p, C.POINTER(_PushStats)]
    for f in (L.orc_forward_push_with_target, L.orc_backward_push_with_source):
        f.argtypes = [C.POINTER(_Csr), C.POINTER(_PushConfig), C.c_uint64, C.c_uint64, C.c_double, f64p, f64p,
                      C.POINTER(_PushStats)]
    L.orc_push_combine_with_forward.argtypes = [C.c_double, f64p, f64p, C.c_uint64, f64p, f64p, C.c_uint64]
    L.orc_push_combine_with_forward.restype = C.c_double
    L.orc_ts_forward_push.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, C.c_double, C.c_uint64, f64p, u64p, f64p,
                                      C.POINTER(C.c_int)]
    L.orc_cg_solve.argtypes = [C.POINTER(_Csr), f64p, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int,
                               C.POINTER(_CgResult)]
    L.orc_gen_bench_k.argtypes = [C.c_uint64, C.c_double]
    L.orc_gen_bench_k.restype = C.c_uint64
    L.orc_gen_bench_csr.argtypes = [C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, C.POINTER(_Csr), f64p]
    L.orc_gen_bench_triplets.argtypes = [C.c_uint64, C.c_double, u64p, u64p, f64p, C.c_uint64]
    L.orc_gen_bench_triplets.restype = C.c_int64
    L.orc_gen_ultra_triplets.argtypes = [C.c_uint64, C.c_double, u64p, u64p, f64p, C.c_uint64]
    L.orc_gen_ultra_triplets.restype = C.c_int64
    L.orc_pagerank_system.argtypes = [u64p, u64p, f64p, C.c_uint64, C.c_uint64, C.c_double, C.POINTER(_Csr), f64p]
    L.orc_solve_entry.argtypes = [C.POINTER(_Csr), f64p, u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, f64p, f64p]
    L.orc_ts_lcg_next.argtypes = [C.c_uint32, f64p]
    L.orc_ts_lcg_next.restype = C.c_uint32
    _libs[key] = L
    return L


class OracleError(Exception):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"oracle error code {code} {what}")
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Csr:
    """CSRStorage restated (values f64 / col_indices u32 / row_ptr u32), numpy-backed."""

    def __init__(self, nrows, ncols, values, col_indices, row_ptr):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.values = _f64(values)
        self.col_indices = np.ascontiguousarray(col_indices, dtype=np.uint32)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        assert self.row_ptr.shape[0] == self.nrows + 1
        self.nnz = int(self.values.shape[0])

    def c(self) -> _Csr:
        return _Csr(self.nrows, self.ncols, self.nnz, _p(self.values, C.c_double),
                    _p(self.col_indices, C.c_uint32), _p(self.row_ptr, C.c_uint32))

    @staticmethod
    def _take(raw: _Csr, L) -> "Csr":
        nnz, n = int(raw.nnz), int(raw.nrows)
        vals = np.ctypeslib.as_array(raw.values, shape=(max(nnz, 1),))[:nnz].copy()
        cols = np.ctypeslib.as_array(raw.col_indices, shape=(max(nnz, 1),))[:nnz].copy()
        rp = np.ctypeslib.as_array(raw.row_ptr, shape=(n + 1,)).copy()
        out = Csr(n, int(raw.ncols), vals, cols, rp)
        L.orc_csr_free(C.byref(raw))
        return out

    @staticmethod
    def from_triplets(rows, cols, vals, nrows, ncols) -> "Csr":
        L = lib()
        r, c, v = _u64(rows), _u64(cols), _f64(vals)
        raw = _Csr()
        rc = L.orc_csr_from_triplets(_p(r, C.c_uint64), _p(c, C.c_uint64), _p(v, C.c_double),
                                     len(v), nrows, ncols, C.byref(raw))
        if rc != OK:
            raise OracleError(rc, "from_triplets")
        return Csr._take(raw, L)

    @staticmethod
    def from_dense(a) -> "Csr":
        """SparseMatrix::from_dense (src/matrix/mod.rs:202-223): row-major scan, zeros filtered."""
        a = np.asarray(a, dtype=np.float64)
        r, c = np.nonzero(a)
        return Csr.from_triplets(r, c, a[r, c], a.shape[0], a.shape[1])

    def get(self, row, col):
        out = C.c_double()
        m = self.c()
        return out.value if lib().orc_csr_get(C.byref(m), row, col, C.byref(out)) else None

    def multiply_vector(self, x, variant=SPMV_SCALAR, nthreads=0, ylen=None):
        x = _f64(x)
        y = np.zeros(self.nrows if ylen is None else ylen)
        m = self.c()
        rc = lib().orc_multiply_vector(C.byref(m), _p(x, C.c_double), len(x), _p(y, C.c_double), len(y),
                                       variant, nthreads)
        if rc != OK:
            raise OracleError(rc, "multiply_vector")
        return y

    def is_diagonally_dominant(self):
        m = self.c()
        bad = C.c_uint64()
        return bool(lib().orc_is_diagonally_dominant(C.byref(m), C.byref(bad)))

    def first_non_dominant_row(self):
        m = self.c()
        bad = C.c_uint64()
        lib().orc_is_diagonally_dominant(C.byref(m), C.byref(bad))
        return None if bad.value == 2 ** 64 - 1 else bad.value

    def is_col_diagonally_dominant(self):
        m = self.c()
        return bool(lib().orc_is_col_diagonally_dominant(C.byref(m)))

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.values, self.col_indices.astype(np.int64), self.row_ptr.astype(np.int64)),
                             shape=(self.nrows, self.ncols))


@dataclass
class Result:
    solution: np.ndarray
    residual_norm: float
    iterations: int
    terms_computed: int
    matvec_count: int
    converged: bool
    series_converged: bool
    error_bound: float | None
    last_term_norm: float
    total_time_ms: float
    status: int = OK


def neumann_solve(m: Csr, b, *, tolerance=1e-6, max_iterations=1000, initial_guess=None,
                  compute_error_bounds=False, max_terms=50, series_tolerance=1e-8,
                  adaptive_truncation=True, mode=MODE_CORRECT, dominance=DOM_ROW,
                  spmv_variant=SPMV_SCALAR, nthreads=0, fast=False, raise_on_error=True) -> Result:
    L = lib(fast)
    b = _f64(b)
    o = _Options()
    L.orc_options_default(C.byref(o))
    o.tolerance, o.max_iterations = tolerance, max_iterations
    o.compute_error_bounds, o.max_terms = int(compute_error_bounds), max_terms
    o.series_tolerance, o.adaptive_truncation = series_tolerance, int(adaptive_truncation)
    o.mode, o.dominance, o.spmv_variant, o.nthreads = mode, dominance, spmv_variant, nthreads
    ig = None
    if initial_guess is not None:
        ig = _f64(initial_guess)
        o.initial_guess, o.initial_guess_len = _p(ig, C.c_double), len(ig)
    x = np.zeros(m.nrows)
    r = _Result()
    r.solution = _p(x, C.c_double)
    mc = m.c()
    rc = L.orc_neumann_solve(C.byref(mc), _p(b, C.c_double), len(b), C.byref(o), C.byref(r))
    if rc != OK and (raise_on_error or rc not in (ERR_CONVERGENCE_FAILURE, ERR_NUMERICAL_INSTABILITY)):
        raise OracleError(rc, "neumann_solve")
    return Result(x, r.residual_norm, int(r.iterations), int(r.terms_computed), int(r.matvec_count),
                  bool(r.converged), bool(r.series_converged),
                  r.error_bound if r.has_error_bound else None, r.last_term_norm, r.total_time_ms, rc)


class NeumannState:
    """SolverAlgorithm::{initialize, step, is_converged, extract_solution, update_rhs} + SolverState::reset restated
    (src/solver/mod.rs:223-252, src/solver/neumann.rs:350-462); step() is the body the reference left commented out."""

    def __init__(self, m: Csr, b, *, tolerance=1e-6, initial_guess=None, max_terms=50, series_tolerance=1e-8,
                 adaptive_truncation=True, mode=MODE_CORRECT, dominance=DOM_ROW, spmv_variant=SPMV_SCALAR):
        L = lib()
        self._L, self._m, self._mc = L, m, m.c()     # keep the arrays the C state points into alive
        b = _f64(b)
        o = _Options()
        L.orc_options_default(C.byref(o))
        o.tolerance, o.max_terms, o.series_tolerance = tolerance, max_terms, series_tolerance
        o.adaptive_truncation, o.mode, o.dominance, o.spmv_variant = int(adaptive_truncation), mode, dominance, spmv_variant
        if initial_guess is not None:
            ig = _f64(initial_guess)
            o.initial_guess, o.initial_guess_len = _p(ig, C.c_double), len(ig)
        self._h = C.c_void_p()
        rc = L.orc_state_new(C.byref(self._mc), _p(b, C.c_double), len(b), C.byref(o), C.byref(self._h))
        if rc != OK:
            self._h = None
            raise OracleError(rc, "state_new")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_state_free(self._h)
            self._h = None

    def step(self):
        out = C.c_int()
        rc = self._L.orc_state_step(self._h, C.byref(out))
        if rc != OK:
            raise OracleError(rc, "state_step")
        return out.value

    def is_converged(self):
        return bool(self._L.orc_state_is_converged(self._h))

    def extract_solution(self):
        x = np.zeros(self._m.nrows)
        self._L.orc_state_solution(self._h, _p(x, C.c_double))
        return x

    def update_rhs(self, delta_b):
        idx = _u64([i for i, _ in delta_b])
        dl = _f64([d for _, d in delta_b])
        rc = self._L.orc_state_update_rhs(self._h, _p(idx, C.c_uint64), _p(dl, C.c_double), len(dl))
        if rc != OK:
            raise OracleError(rc, "update_rhs")

    def reset(self):
        self._L.orc_state_reset(self._h)

    def info(self):
        rn, tn, bd = C.c_double(), C.c_double(), C.c_double()
        mv, tc = C.c_uint64(), C.c_uint64()
        sc, hb = C.c_int(), C.c_int()
        self._L.orc_state_info(self._h, C.byref(rn), C.byref(mv), C.byref(tc), C.byref(sc), C.byref(tn), C.byref(hb), C.byref(bd))
        return {"residual_norm": rn.value, "matvec_count": mv.value, "terms_computed": tc.value,
                "series_converged": bool(sc.value), "last_term_norm": tn.value,
                "error_upper_bound": bd.value if hb.value else None}


@dataclass
class PushResult:
    """ForwardPushResult / BackwardPushResult (src/solver/forward_push.rs:10-22, backward_push.rs:10-22)"""
    estimate: np.ndarray
    residual: np.ndarray
    push_count: int
    nodes_visited: int
    residual_norm: float

    def extrapolated_solution(self, alpha):
        """ForwardPushSolver::extrapolated_solution (forward_push.rs:317-327)"""
        return self.estimate + alpha * self.residual


def _push(fn_name, adj: Csr, seeds, alpha, epsilon, max_pushes, queue_threshold, adaptive_threshold):
    L = lib()
    c = _PushConfig()
    L.orc_push_config_default(C.byref(c))
    c.alpha, c.epsilon, c.max_pushes = alpha, epsilon, max_pushes
    c.queue_threshold, c.adaptive_threshold = queue_threshold, int(adaptive_threshold)
    s = _u64(np.atleast_1d(seeds))
    est, res = np.zeros(adj.nrows), np.zeros(adj.nrows)
    st = _PushStats()
    a = adj.c()
    rc = getattr(L, fn_name)(C.byref(a), C.byref(c), _p(s, C.c_uint64), len(s), _p(est, C.c_double), _p(res, C.c_double),
                             C.byref(st))
    if rc != OK:
        raise OracleError(rc, fn_name)
    return PushResult(est, res, int(st.push_count), int(st.nodes_visited), st.residual_norm)


def forward_push(adj: Csr, sources, alpha=0.15, epsilon=1e-6, max_pushes=1_000_000, queue_threshold=1e-8,
                 adaptive_threshold=True) -> PushResult:
    """ForwardPushSolver::solve_single_source / solve_multi_source (src/solver/forward_push.rs:66-177) restated."""
    return _push("orc_forward_push", adj, sources, alpha, epsilon, max_pushes, queue_threshold, adaptive_threshold)


def backward_push(adj: Csr, targets, alpha=0.15, epsilon=1e-6, max_pushes=1_000_000, queue_threshold=1e-8,
                  adaptive_threshold=True) -> PushResult:
    """BackwardPushSolver::solve_single_target / solve_multi_target (src/solver/backward_push.rs:66-177) restated."""
    return _push("orc_backward_push", adj, targets, alpha, epsilon, max_pushes, queue_threshold, adaptive_threshold)


def _push_watch(fn_name, adj: Csr, source, target, precision, alpha, epsilon, max_pushes, queue_threshold,
                adaptive_threshold):
    L = lib()
    c = _PushConfig()
    L.orc_push_config_default(C.byref(c))
    c.alpha, c.epsilon, c.max_pushes = alpha, epsilon, max_pushes
    c.queue_threshold, c.adaptive_threshold = queue_threshold, int(adaptive_threshold)
    est, res = np.zeros(adj.nrows), np.zeros(adj.nrows)
    st = _PushStats()
    a = adj.c()
    rc = getattr(L, fn_name)(C.byref(a), C.byref(c), source, target, precision, _p(est, C.c_double), _p(res, C.c_double),
                             C.byref(st))
    if rc != OK:
        raise OracleError(rc, fn_name)
    return PushResult(est, res, int(st.push_count), int(st.nodes_visited), st.residual_norm)


def forward_push_with_target(adj: Csr, source, target, target_precision, alpha=0.15, epsilon=1e-6, max_pushes=1_000_000,
                             queue_threshold=1e-8, adaptive_threshold=True) -> PushResult:
    """ForwardPushSolver::solve_with_target (src/solver/forward_push.rs:234-290) restated."""
    return _push_watch("orc_forward_push_with_target", adj, source, target, target_precision, alpha, epsilon, max_pushes,
                       queue_threshold, adaptive_threshold)


def backward_push_with_source(adj: Csr, source, target, source_precision, alpha=0.15, epsilon=1e-6, max_pushes=1_000_000,
                              queue_threshold=1e-8, adaptive_threshold=True) -> PushResult:
    """BackwardPushSolver::solve_with_source (src/solver/backward_push.rs:238-290) restated."""
    return _push_watch("orc_backward_push_with_source", adj, source, target, source_precision, alpha, epsilon, max_pushes,
                       queue_threshold, adaptive_threshold)


def push_combine_with_forward(alpha, backward: PushResult, forward_estimate, forward_residual) -> float:
    """BackwardPushSolver::combine_with_forward (src/solver/backward_push.rs:312-330) restated."""
    be, br, fe, fr = _f64(backward.estimate), _f64(backward.residual), _f64(forward_estimate), _f64(forward_residual)
    return float(lib().orc_push_combine_with_forward(alpha, _p(be, C.c_double), _p(br, C.c_double), len(be),
                                                     _p(fe, C.c_double), _p(fr, C.c_double), len(fe)))


@dataclass
class TsPushResult:
    solution: np.ndarray
    iterations: int
    residual: float
    converged: bool
    status: int


def ts_forward_push(a: Csr, b, epsilon=1e-6, max_iterations=1000) -> TsPushResult:
    """SublinearSolver.solveForwardPush (src/core/solver.ts:437-522) restated (Gauss-Southwell, one node per iteration)."""
    b = _f64(b)
    x = np.zeros(a.nrows)
    it, res, conv = C.c_uint64(), C.c_double(), C.c_int()
    ac = a.c()
    rc = lib().orc_ts_forward_push(C.byref(ac), _p(b, C.c_double), len(b), epsilon, max_iterations, _p(x, C.c_double),
                                   C.byref(it), C.byref(res), C.byref(conv))
    if rc not in (OK, ERR_CONVERGENCE_FAILURE, ERR_NUMERICAL_INSTABILITY):
        raise OracleError(rc, "ts_forward_push")
    return TsPushResult(x, int(it.value), res.value, bool(conv.value), rc)


@dataclass
class CgResult:
    solution: np.ndarray
    residual_norm: float
    iterations: int
    converged: bool
    matvec_count: int
    total_flops: int


def cg_solve(m: Csr, b, *, max_iterations=1000, tolerance=1e-6, spmv_variant=SPMV_SCALAR,
             dot_variant=DOT_SEQUENTIAL, nthreads=0, fast=False) -> CgResult:
    """OptimizedConjugateGradientSolver::solve (src/optimized_solver.rs:182-295) / FastConjugateGradient /
    UltraFastCG restated; dot_variant selects the summation order of the three reference variants."""
    L = lib(fast)
    b = _f64(b)
    x = np.zeros(m.nrows)
    r = _CgResult()
    r.solution = _p(x, C.c_double)
    mc = m.c()
    rc = L.orc_cg_solve(C.byref(mc), _p(b, C.c_double), len(b), max_iterations, tolerance, spmv_variant, dot_variant,
                        nthreads, C.byref(r))
    if rc != OK:
        raise OracleError(rc, "cg_solve")
    return CgResult(x, r.residual_norm, int(r.iterations), bool(r.converged), int(r.matvec_count), int(r.total_flops))


def push_iterations(m: Csr, b, nterms, spmv_variant=SPMV_SCALAR, nthreads=0, fast=False):
    """x, t, per-term norms and seconds for `nterms` bare push iterations after term 0."""
    L = lib(fast)
    b = _f64(b)
    x, t, norms = np.zeros(m.nrows), np.zeros(m.nrows), np.zeros(max(nterms, 1))
    mc = m.c()
    secs = L.orc_push_iterations(C.byref(mc), _p(b, C.c_double), nterms, spmv_variant, nthreads,
                                 _p(x, C.c_double), _p(t, C.c_double), _p(norms, C.c_double))
    return x, t, norms[:nterms], secs


def gen_bench_k(size, sparsity):
    return int(lib().orc_gen_bench_k(size, sparsity))


def gen_bench_csr(size, sparsity, row0=0, row1=None, fast=False):
    """create_test_matrix/create_test_rhs (benches/performance_benchmarks.rs:12-43), rows [row0,row1)."""
    L = lib(fast)
    row1 = size if row1 is None else row1
    raw = _Csr()
    b = np.zeros(row1 - row0)
    rc = L.orc_gen_bench_csr(size, sparsity, row0, row1, C.byref(raw), _p(b, C.c_double))
    if rc != OK:
        raise OracleError(rc, "gen_bench_csr")
    return Csr._take(raw, L), b


def gen_bench_triplets(size, sparsity):
    L = lib()
    cap = size * gen_bench_k(size, sparsity) + 1
    r, c, v = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap)
    n = L.orc_gen_bench_triplets(size, sparsity, _p(r, C.c_uint64), _p(c, C.c_uint64), _p(v, C.c_double), cap)
    assert n >= 0
    b = 1.0 + np.arange(size, dtype=np.float64) * 0.001
    return r[:n], c[:n], v[:n], b


def gen_ultra_triplets(size, sparsity):
    L = lib()
    cap = size * 11 + 1
    r, c, v = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap)
    n = L.orc_gen_ultra_triplets(size, sparsity, _p(r, C.c_uint64), _p(c, C.c_uint64), _p(v, C.c_double), cap)
    assert n >= 0
    return r[:n], c[:n], v[:n], np.ones(size)


def pagerank_system(src, dst, n, alpha=0.85, weights=None):
    L = lib()
    s, d = _u64(src), _u64(dst)
    w = None if weights is None else _f64(weights)
    raw = _Csr()
    rhs = np.zeros(n)
    rc = L.orc_pagerank_system(_p(s, C.c_uint64), _p(d, C.c_uint64),
                               None if w is None else _p(w, C.c_double), len(s), n, alpha,
                               C.byref(raw), _p(rhs, C.c_double))
    if rc != OK:
        raise OracleError(rc, "pagerank_system")
    return Csr._take(raw, L), rhs


def solve_entry(m: Csr, b, rows, nwalks, max_steps=1000, seed=0):
    L = lib()
    b, q = _f64(b), _u64(rows)
    est, var = np.zeros(len(q)), np.zeros(len(q))
    mc = m.c()
    rc = L.orc_solve_entry(C.byref(mc), _p(b, C.c_double), _p(q, C.c_uint64), len(q), nwalks, max_steps, seed,
                           _p(est, C.c_double), _p(var, C.c_double))
    if rc != OK:
        raise OracleError(rc, "solve_entry")
    return est, var


def ts_lcg(seed, count):
    """createSeededRandom (src/core/utils.ts:161-168) stream."""
    L = lib()
    s, out = C.c_uint32(seed & 0xFFFFFFFF).value, []
    for _ in range(count):
        u = C.c_double()
        s = L.orc_ts_lcg_next(s, C.byref(u))
        out.append(u.value)
    return out
