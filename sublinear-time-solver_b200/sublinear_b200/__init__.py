"""ctypes binding of libsublinear_b200.so — the Python-side mirror of the reference crate's API for the
Neumann / push-iteration path (names follow the Rust crate: SparseMatrix, NeumannSolver, SolverOptions,
SolverResult, SolverError).  Used by tests/ and bench.py; it adds no compute of its own: every call goes
through the C ABI declared in include/sublinear_b200.h, and there is no CPU fallback — a missing library or
GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_DIR, "libsublinear_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "sublinear_b200.h")

OK = 0
ERROR_NAMES = {
    1: "MatrixNotDiagonallyDominant", 2: "NumericalInstability", 3: "ConvergenceFailure", 4: "InvalidInput",
    5: "DimensionMismatch", 6: "UnsupportedMatrixFormat", 7: "MemoryAllocationError", 8: "IndexOutOfBounds",
    9: "InvalidSparseMatrix", 10: "AlgorithmError", 11: "WasmBindingError", 12: "IoError", 13: "SerializationError",
}
MODE_CORRECT, MODE_REF_COMPAT = 0, 1
DOMINANCE_ROW, DOMINANCE_ROW_OR_COL = 0, 1
RESIDUAL_EVERY_5, RESIDUAL_IDENTITY = 0, 1
DUP_KEEP, DUP_SUM = 0, 1
LAYOUT_CSR, LAYOUT_SELL32, LAYOUT_CSR_SLABS = 0, 1, 2
STEP_CONTINUE, STEP_CONVERGED = 0, 1
UNIQUE_ID_BYTES = 128


class SolverError(Exception):
    """`enum SolverError` (src/error.rs:16-138): .code is the 1-based variant index, .variant its name."""

    def __init__(self, code: int, message: str, result=None):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.variant = ERROR_NAMES.get(code, str(code))
        self.message = message
        self.result = result  # iterations / residual_norm of ConvergenceFailure / NumericalInstability


class _Options(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("max_iterations", C.c_uint64), ("convergence_mode", C.c_int32),
                ("norm_type", C.c_int32), ("collect_stats", C.c_int32), ("streaming_interval", C.c_uint64),
                ("initial_guess", C.c_void_p), ("initial_guess_len", C.c_uint64),
                ("compute_error_bounds", C.c_int32), ("error_bounds_tolerance", C.c_double),
                ("enable_profiling", C.c_int32), ("has_random_seed", C.c_int32), ("random_seed", C.c_uint64),
                ("mode", C.c_int32), ("dominance", C.c_int32), ("residual_check", C.c_int32), ("reserved", C.c_int32)]


class _StateInfo(C.Structure):
    _fields_ = [("dimension", C.c_uint64), ("residual_norm", C.c_double), ("matvec_count", C.c_uint64),
                ("terms_computed", C.c_uint64), ("series_converged", C.c_int32), ("last_term_norm", C.c_double),
                ("has_error_bounds", C.c_int32), ("error_upper_bound", C.c_double), ("memory_bytes", C.c_uint64)]


class _Partial(C.Structure):
    _fields_ = [("iteration", C.c_uint64), ("solution", C.POINTER(C.c_double)), ("solution_len", C.c_uint64),
                ("residual_norm", C.c_double), ("converged", C.c_int32), ("has_estimated_remaining", C.c_int32),
                ("estimated_remaining", C.c_uint64), ("timestamp_ms", C.c_double)]


_STREAM_CB = C.CFUNCTYPE(C.c_int32, C.POINTER(_Partial), C.c_void_p)
_CHUNK_CB = C.CFUNCTYPE(C.c_int32, C.c_uint64, C.POINTER(C.c_double), C.c_uint64, C.c_void_p)


class _PushConfig(C.Structure):
    _fields_ = [("alpha", C.c_double), ("epsilon", C.c_double), ("max_pushes", C.c_uint64),
                ("queue_threshold", C.c_double), ("adaptive_threshold", C.c_int32), ("reserved", C.c_int32)]


class _PushStats(C.Structure):
    _fields_ = [("push_count", C.c_uint64), ("nodes_visited", C.c_uint64), ("residual_norm", C.c_double),
                ("rounds", C.c_uint64), ("kernel_launches", C.c_uint64), ("device_time_ms", C.c_double),
                ("dense_rounds", C.c_uint64), ("edges_touched", C.c_uint64)]


class _AxbPushStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("rounds", C.c_uint64), ("residual_norm", C.c_double),
                ("max_residual", C.c_double), ("converged", C.c_int32), ("reserved", C.c_int32)]


class _CgConfig(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("tolerance", C.c_double), ("enable_profiling", C.c_int32),
                ("reserved", C.c_int32)]


class _CgResult(C.Structure):
    _fields_ = [("solution", C.POINTER(C.c_double)), ("solution_len", C.c_uint64), ("residual_norm", C.c_double),
                ("iterations", C.c_uint64), ("converged", C.c_int32), ("breakdown", C.c_int32),
                ("computation_time_ms", C.c_double), ("matvec_count", C.c_uint64), ("dot_product_count", C.c_uint64),
                ("axpy_count", C.c_uint64), ("total_flops", C.c_uint64), ("average_bandwidth_gbs", C.c_double),
                ("average_gflops", C.c_double), ("device_time_ms", C.c_double), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("spmv_kernel_ms", C.c_double),
                ("spmv_kernel_count", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("solution", C.POINTER(C.c_double)), ("solution_len", C.c_uint64), ("residual_norm", C.c_double),
                ("iterations", C.c_uint64), ("converged", C.c_int32), ("has_error_bounds", C.c_int32),
                ("error_upper_bound", C.c_double), ("has_stats", C.c_int32), ("total_time_ms", C.c_double),
                ("matvec_count", C.c_uint64), ("memory_bytes", C.c_uint64), ("terms_computed", C.c_uint64),
                ("series_converged", C.c_int32), ("last_term_norm", C.c_double), ("device_time_ms", C.c_double),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("push_kernel_ms", C.c_double), ("push_kernel_count", C.c_uint64),
                ("resid_kernel_ms", C.c_double), ("resid_kernel_count", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", PKG_DIR, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", PKG_DIR, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libsublinear_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32, f64 = C.c_void_p, C.c_uint64, C.c_int32, C.c_double
    P = C.POINTER
    sig = {
        "sb200_abi_version": ([], i32),
        "sb200_last_error": ([C.c_char_p, C.c_size_t], C.c_size_t),
        "sb200_device_count": ([P(i32)], i32),
        "sb200_set_device": ([i32], i32),
        "sb200_get_device": ([P(i32)], i32),
        "sb200_host_alloc": ([u64, P(vp)], i32),
        "sb200_host_free": ([vp], i32),
        "sb200_matrix_from_triplets": ([vp, vp, vp, u64, u64, u64, P(vp)], i32),
        "sb200_matrix_from_triplets_ex": ([vp, vp, vp, u64, u64, u64, i32, P(vp)], i32),
        "sb200_matrix_from_csr": ([vp, vp, vp, u64, u64, u64, P(vp)], i32),
        "sb200_matrix_from_csr64": ([vp, vp, vp, u64, u64, u64, P(vp)], i32),
        "sb200_matrix_from_dense": ([vp, u64, u64, P(vp)], i32),
        "sb200_matrix_identity": ([u64, P(vp)], i32),
        "sb200_matrix_diagonal": ([vp, u64, P(vp)], i32),
        "sb200_matrix_free": ([vp], None),
        "sb200_matrix_rows": ([vp, P(u64)], i32),
        "sb200_matrix_cols": ([vp, P(u64)], i32),
        "sb200_matrix_nnz": ([vp, P(u64)], i32),
        "sb200_matrix_get": ([vp, u64, u64, P(f64), P(i32)], i32),
        "sb200_matrix_is_diagonally_dominant": ([vp, i32, P(i32)], i32),
        "sb200_matrix_diagonal_dominance_factor": ([vp, P(f64), P(i32)], i32),
        "sb200_matrix_storage_info": ([vp, P(i32), P(u64), P(u64)], i32),
        "sb200_matrix_export_csr": ([vp, vp, vp, vp], i32),
        "sb200_matrix_multiply_vector": ([vp, vp, u64, vp, u64], i32),
        "sb200_matrix_multiply_vector_add": ([vp, vp, u64, vp, u64], i32),
        "sb200_matrix_multiply_vector_dev": ([vp, vp, u64, vp, u64, i32, vp], i32),
        "sb200_matrix_scale": ([vp, f64], i32),
        "sb200_neumann_new": ([u64, f64, P(vp)], i32),
        "sb200_neumann_default": ([P(vp)], i32),
        "sb200_neumann_high_precision": ([P(vp)], i32),
        "sb200_neumann_fast": ([P(vp)], i32),
        "sb200_neumann_with_adaptive_truncation": ([vp, i32], i32),
        "sb200_neumann_with_power_caching": ([vp, i32], i32),
        "sb200_neumann_config": ([vp, P(u64), P(f64), P(i32), P(i32)], i32),
        "sb200_solver_algorithm_name": ([vp], C.c_char_p),
        "sb200_solver_free": ([vp], None),
        "sb200_options_default": ([P(_Options)], None),
        "sb200_options_high_precision": ([P(_Options)], None),
        "sb200_options_fast": ([P(_Options)], None),
        "sb200_options_streaming": ([P(_Options), u64], None),
        "sb200_solve": ([vp, vp, vp, u64, P(_Options), P(_Result)], i32),
        "sb200_solve_into": ([vp, vp, vp, u64, P(_Options), vp, P(_Result)], i32),
        "sb200_solve_dev": ([vp, vp, vp, u64, P(_Options), vp, vp, P(_Result)], i32),
        "sb200_result_free": ([P(_Result)], None),
        "sb200_push_iterations_dev": ([vp, vp, u64, u64, vp, vp, vp, vp, P(C.c_float)], i32),
        "sb200_neumann_initialize": ([vp, vp, vp, u64, P(_Options), P(vp)], i32),
        "sb200_state_step": ([vp, P(i32)], i32),
        "sb200_state_is_converged": ([vp, P(i32)], i32),
        "sb200_state_extract_solution": ([vp, vp, u64], i32),
        "sb200_state_update_rhs": ([vp, vp, vp, u64], i32),
        "sb200_state_reset": ([vp], i32),
        "sb200_state_info": ([vp, P(_StateInfo)], i32),
        "sb200_state_free": ([vp], None),
        "sb200_solve_streaming": ([vp, vp, vp, u64, P(_Options), _STREAM_CB, vp, P(_Result)], i32),
        "sb200_push_config_default": ([P(_PushConfig)], None),
        "sb200_push_graph_from_csr": ([vp, vp, vp, u64, P(vp)], i32),
        "sb200_push_graph_from_edges": ([u64, vp, vp, vp, u64, P(vp)], i32),
        "sb200_push_graph_info": ([vp, P(u64), P(u64)], i32),
        "sb200_push_graph_degrees": ([vp, u64, P(f64), P(f64)], i32),
        "sb200_push_graph_free": ([vp], None),
        "sb200_forward_push": ([vp, P(_PushConfig), vp, u64, vp, vp, P(_PushStats)], i32),
        "sb200_backward_push": ([vp, P(_PushConfig), vp, u64, vp, vp, P(_PushStats)], i32),
        "sb200_forward_push_with_target": ([vp, P(_PushConfig), u64, u64, f64, vp, vp, P(_PushStats)], i32),
        "sb200_backward_push_with_source": ([vp, P(_PushConfig), u64, u64, f64, vp, vp, P(_PushStats)], i32),
        "sb200_push_combine_with_forward": ([f64, vp, vp, u64, vp, vp, u64, P(f64)], i32),
        "sb200_bidirectional_push": ([vp, P(_PushConfig), P(_PushConfig), u64, u64, P(f64)], i32),
        "sb200_bidirectional_adaptive_push": ([vp, P(_PushConfig), P(_PushConfig), u64, u64, P(f64)], i32),
        "sb200_forward_push_solve": ([vp, vp, u64, f64, u64, vp, P(_AxbPushStats)], i32),
        "sb200_streaming_matrix_from_triplets": ([vp, vp, vp, u64, u64, u64, u64, P(vp)], i32),
        "sb200_streaming_matrix_info": ([vp, P(u64), P(u64), P(u64), P(u64), P(u64)], i32),
        "sb200_streaming_matrix_multiply_vector": ([vp, vp, u64, _CHUNK_CB, vp], i32),
        "sb200_streaming_matrix_free": ([vp], None),
        "sb200_cg_config_default": ([P(_CgConfig)], None),
        "sb200_cg_solve": ([vp, vp, u64, P(_CgConfig), P(_CgResult)], i32),
        "sb200_cg_solve_into": ([vp, vp, u64, P(_CgConfig), vp, P(_CgResult)], i32),
        "sb200_cg_solve_dev": ([vp, vp, u64, P(_CgConfig), vp, vp, P(_CgResult)], i32),
        "sb200_cg_result_free": ([P(_CgResult)], None),
        "sb200_solve_entry": ([vp, vp, u64, vp, u64, f64, u64, u64, u64, vp, vp], i32),
        "sb200_solve_entry_replicas": ([vp, C.c_int32, vp, u64, vp, u64, f64, u64, u64, u64, vp, vp], i32),
        "sb200_pagerank_system": ([vp, vp, vp, u64, u64, f64, P(vp), vp], i32),
        "sb200_gen_bench_csr": ([u64, f64, u64, u64, vp, vp, vp, vp, P(u64)], i32),
        "sb200_comm_unique_id": ([vp], i32),
        "sb200_comm_init": ([i32, i32, vp, i32, P(vp)], i32),
        "sb200_comm_free": ([vp], None),
        "sb200_partition_rows": ([u64, i32, i32, P(u64), P(u64)], i32),
        "sb200_dist_matrix_from_csr": ([vp, u64, u64, u64, vp, vp, vp, P(vp)], i32),
        "sb200_dist_solve": ([vp, vp, vp, vp, u64, P(_Options), vp, P(_Result)], i32),
        "sb200_dist_push_iterations_dev": ([vp, vp, vp, u64, u64, vp, vp, P(C.c_float)], i32),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)  # AttributeError here = the library does not export what the header declares
        f.argtypes = args
        f.restype = res
    _lib = L
    return L


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    lib().sb200_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def _check(rc: int, result=None):
    if rc != OK:
        raise SolverError(rc, last_error(), result)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def device_count() -> int:
    n = C.c_int32()
    _check(lib().sb200_device_count(C.byref(n)))
    return n.value


def set_device(dev: int):
    _check(lib().sb200_set_device(dev))


class PinnedArray:
    """float64 numpy view over sb200_host_alloc memory (PCIe line-rate source / destination)."""

    def __init__(self, n: int):
        self._p = C.c_void_p()
        _check(lib().sb200_host_alloc(max(n, 1) * 8, C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_double)), shape=(max(n, 1),))[:n]

    def __del__(self):
        if getattr(self, "_p", None) and _lib is not None:
            _lib.sb200_host_free(self._p)
            self._p = None


@dataclass
class SolverOptions:
    """`SolverOptions` (src/solver/mod.rs:22-45) + this library's extensions (mode / dominance / residual_check)."""
    tolerance: float = 1e-6
    max_iterations: int = 1000
    convergence_mode: int = 0
    norm_type: int = 1
    collect_stats: bool = False
    streaming_interval: int = 0
    initial_guess: np.ndarray | None = None
    compute_error_bounds: bool = False
    error_bounds_tolerance: float = 1e-8
    enable_profiling: bool = False
    random_seed: int | None = None
    mode: int = MODE_CORRECT
    dominance: int = DOMINANCE_ROW
    residual_check: int = RESIDUAL_EVERY_5

    @staticmethod
    def _preset(fn, *a) -> "SolverOptions":
        o = _Options()
        fn(C.byref(o), *a)
        return SolverOptions(o.tolerance, o.max_iterations, o.convergence_mode, o.norm_type, bool(o.collect_stats),
                             o.streaming_interval, None, bool(o.compute_error_bounds), o.error_bounds_tolerance,
                             bool(o.enable_profiling), None, o.mode, o.dominance, o.residual_check)

    @staticmethod
    def default():
        return SolverOptions._preset(lib().sb200_options_default)

    @staticmethod
    def high_precision():
        return SolverOptions._preset(lib().sb200_options_high_precision)

    @staticmethod
    def fast():
        return SolverOptions._preset(lib().sb200_options_fast)

    @staticmethod
    def streaming(interval: int):
        return SolverOptions._preset(lib().sb200_options_streaming, interval)

    def _c(self, guess_ptr=None, guess_len=0):
        o = _Options()
        lib().sb200_options_default(C.byref(o))
        o.tolerance, o.max_iterations = self.tolerance, self.max_iterations
        o.convergence_mode, o.norm_type = self.convergence_mode, self.norm_type
        o.collect_stats, o.streaming_interval = int(self.collect_stats), self.streaming_interval
        o.compute_error_bounds, o.error_bounds_tolerance = int(self.compute_error_bounds), self.error_bounds_tolerance
        o.enable_profiling = int(self.enable_profiling)
        o.has_random_seed, o.random_seed = int(self.random_seed is not None), self.random_seed or 0
        o.mode, o.dominance, o.residual_check = self.mode, self.dominance, self.residual_check
        if guess_ptr is not None:
            o.initial_guess, o.initial_guess_len = guess_ptr, guess_len
        return o


@dataclass
class SolverResult:
    """`SolverResult` (src/solver/mod.rs:121-138) + SolverStats / extension fields."""
    solution: np.ndarray | None
    residual_norm: float
    iterations: int
    converged: bool
    error_upper_bound: float | None
    matvec_count: int
    total_time_ms: float
    memory_bytes: int
    terms_computed: int
    series_converged: bool
    last_term_norm: float
    device_time_ms: float
    kernel_launches: int
    h2d_bytes: int
    d2h_bytes: int
    push_kernel_ms: float = 0.0
    push_kernel_count: int = 0
    resid_kernel_ms: float = 0.0
    resid_kernel_count: int = 0

    @staticmethod
    def _from(r: _Result, solution):
        return SolverResult(solution, r.residual_norm, int(r.iterations), bool(r.converged),
                            r.error_upper_bound if r.has_error_bounds else None, int(r.matvec_count),
                            r.total_time_ms, int(r.memory_bytes), int(r.terms_computed), bool(r.series_converged),
                            r.last_term_norm, r.device_time_ms, int(r.kernel_launches), int(r.h2d_bytes),
                            int(r.d2h_bytes), r.push_kernel_ms, int(r.push_kernel_count), r.resid_kernel_ms,
                            int(r.resid_kernel_count))


class SparseMatrix:
    """`SparseMatrix` in CSR form (src/matrix/mod.rs:123-372), resident in HBM."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_matrix_free(self._h)
            self._h = None

    # constructors ------------------------------------------------------------------------------
    @staticmethod
    def from_triplets(rows, cols, vals, nrows, ncols, dup_policy=DUP_KEEP) -> "SparseMatrix":
        r, c, v = _u64(rows), _u64(cols), _f64(vals)
        h = C.c_void_p()
        _check(lib().sb200_matrix_from_triplets_ex(_ptr(r), _ptr(c), _ptr(v), len(v), nrows, ncols, dup_policy,
                                                   C.byref(h)))
        return SparseMatrix(h)

    @staticmethod
    def from_csr(row_ptr, col_indices, values, nrows, ncols) -> "SparseMatrix":
        ci = np.ascontiguousarray(col_indices, dtype=np.uint32)
        v = _f64(values)
        h = C.c_void_p()
        rp = np.asarray(row_ptr)
        if rp.dtype == np.uint32:
            rp = np.ascontiguousarray(rp)
            _check(lib().sb200_matrix_from_csr(_ptr(rp), _ptr(ci), _ptr(v), nrows, ncols, len(v), C.byref(h)))
        else:
            rp = _u64(rp)
            _check(lib().sb200_matrix_from_csr64(_ptr(rp), _ptr(ci), _ptr(v), nrows, ncols, len(v), C.byref(h)))
        return SparseMatrix(h)

    @staticmethod
    def from_dense(a) -> "SparseMatrix":
        a = _f64(a)
        h = C.c_void_p()
        _check(lib().sb200_matrix_from_dense(_ptr(a), a.shape[0], a.shape[1], C.byref(h)))
        return SparseMatrix(h)

    @staticmethod
    def identity(n) -> "SparseMatrix":
        h = C.c_void_p()
        _check(lib().sb200_matrix_identity(n, C.byref(h)))
        return SparseMatrix(h)

    @staticmethod
    def diagonal(d) -> "SparseMatrix":
        d = _f64(d)
        h = C.c_void_p()
        _check(lib().sb200_matrix_diagonal(_ptr(d), len(d), C.byref(h)))
        return SparseMatrix(h)

    @staticmethod
    def pagerank_system(src, dst, n, alpha=0.85, weights=None):
        """computePageRank's system (src/core/solver.ts:664-722): returns (S, rhs)."""
        s, d = _u64(src), _u64(dst)
        w = None if weights is None else _f64(weights)
        rhs = np.zeros(n)
        h = C.c_void_p()
        _check(lib().sb200_pagerank_system(_ptr(s), _ptr(d), _ptr(w), len(s), n, alpha, C.byref(h), _ptr(rhs)))
        return SparseMatrix(h), rhs

    # Matrix trait ------------------------------------------------------------------------------
    def _dim(self, fn):
        out = C.c_uint64()
        _check(fn(self._h, C.byref(out)))
        return out.value

    def rows(self):
        return self._dim(lib().sb200_matrix_rows)

    def cols(self):
        return self._dim(lib().sb200_matrix_cols)

    def nnz(self):
        return self._dim(lib().sb200_matrix_nnz)

    def is_square(self):
        return self.rows() == self.cols()

    def get(self, row, col):
        v, p = C.c_double(), C.c_int32()
        _check(lib().sb200_matrix_get(self._h, row, col, C.byref(v), C.byref(p)))
        return v.value if p.value else None

    def is_diagonally_dominant(self, dominance=DOMINANCE_ROW):
        out = C.c_int32()
        _check(lib().sb200_matrix_is_diagonally_dominant(self._h, dominance, C.byref(out)))
        return bool(out.value)

    def diagonal_dominance_factor(self):
        f, p = C.c_double(), C.c_int32()
        _check(lib().sb200_matrix_diagonal_dominance_factor(self._h, C.byref(f), C.byref(p)))
        return f.value if p.value else None

    def multiply_vector(self, x, ylen=None):
        x = _f64(x)
        y = np.zeros(self.rows() if ylen is None else ylen)
        _check(lib().sb200_matrix_multiply_vector(self._h, _ptr(x), len(x), _ptr(y), len(y)))
        return y

    def multiply_vector_add(self, x, y):
        x, y = _f64(x), _f64(y).copy()
        _check(lib().sb200_matrix_multiply_vector_add(self._h, _ptr(x), len(x), _ptr(y), len(y)))
        return y

    def multiply_vector_dev(self, x_ptr, xlen, y_ptr, ylen, accumulate=False, stream=0):
        _check(lib().sb200_matrix_multiply_vector_dev(self._h, x_ptr, xlen, y_ptr, ylen, int(accumulate), stream))

    def scale(self, factor):
        _check(lib().sb200_matrix_scale(self._h, factor))

    def storage_info(self):
        """Device layout the hot kernels read: {'layout': LAYOUT_CSR | LAYOUT_SELL32 | LAYOUT_CSR_SLABS, 'slots',
        'device_bytes'}."""
        lay, slots, nbytes = C.c_int32(), C.c_uint64(), C.c_uint64()
        _check(lib().sb200_matrix_storage_info(self._h, C.byref(lay), C.byref(slots), C.byref(nbytes)))
        return {"layout": lay.value, "slots": slots.value, "device_bytes": nbytes.value}

    def to_csr(self):
        n, nnz = self.rows(), self.nnz()
        rp, ci, v = np.zeros(n + 1, np.uint64), np.zeros(nnz, np.uint32), np.zeros(nnz)
        _check(lib().sb200_matrix_export_csr(self._h, _ptr(rp), _ptr(ci), _ptr(v)))
        return rp, ci, v

    def to_triplets(self):
        rp, ci, v = self.to_csr()
        rows = np.repeat(np.arange(self.rows(), dtype=np.uint64), np.diff(rp).astype(np.int64))
        return rows, ci.astype(np.uint64), v


class NeumannSolver:
    """`NeumannSolver` (src/solver/neumann.rs:24-93, 469-555)."""

    def __init__(self, max_terms: int = 50, series_tolerance: float = 1e-8, _handle=None):
        if _handle is None:
            _handle = C.c_void_p()
            _check(lib().sb200_neumann_new(max_terms, series_tolerance, C.byref(_handle)))
        self._h = _handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_solver_free(self._h)
            self._h = None

    @staticmethod
    def _mk(fn):
        h = C.c_void_p()
        _check(fn(C.byref(h)))
        return NeumannSolver(_handle=h)

    @staticmethod
    def new(max_terms, series_tolerance):
        return NeumannSolver(max_terms, series_tolerance)

    @staticmethod
    def default():
        return NeumannSolver._mk(lib().sb200_neumann_default)

    @staticmethod
    def high_precision():
        return NeumannSolver._mk(lib().sb200_neumann_high_precision)

    @staticmethod
    def fast():
        return NeumannSolver._mk(lib().sb200_neumann_fast)

    def with_adaptive_truncation(self, enable):
        _check(lib().sb200_neumann_with_adaptive_truncation(self._h, int(enable)))
        return self

    def with_power_caching(self, enable):
        _check(lib().sb200_neumann_with_power_caching(self._h, int(enable)))
        return self

    def config(self):
        mt, tol, ad, cp = C.c_uint64(), C.c_double(), C.c_int32(), C.c_int32()
        _check(lib().sb200_neumann_config(self._h, C.byref(mt), C.byref(tol), C.byref(ad), C.byref(cp)))
        return {"max_terms": mt.value, "series_tolerance": tol.value, "adaptive_truncation": bool(ad.value),
                "cache_powers": bool(cp.value)}

    def algorithm_name(self):
        return lib().sb200_solver_algorithm_name(self._h).decode()

    def solve(self, matrix: SparseMatrix, b, options: SolverOptions | None = None, out=None) -> SolverResult:
        """`solver.solve(&matrix, &b, &options)`: host vectors in, host solution out.
        `out` (optional float64 array, e.g. PinnedArray.array) receives the solution in place."""
        options = options or SolverOptions()
        b = _f64(b)
        guess = None if options.initial_guess is None else _f64(options.initial_guess)
        o = options._c(None if guess is None else guess.ctypes.data, 0 if guess is None else len(guess))
        r = _Result()
        if out is not None:
            rc = lib().sb200_solve_into(self._h, matrix._h, _ptr(b), len(b), C.byref(o), _ptr(out), C.byref(r))
            res = SolverResult._from(r, out)
        else:
            rc = lib().sb200_solve(self._h, matrix._h, _ptr(b), len(b), C.byref(o), C.byref(r))
            sol = None
            if r.solution:
                sol = np.ctypeslib.as_array(r.solution, shape=(max(int(r.solution_len), 1),))[:int(r.solution_len)].copy()
            res = SolverResult._from(r, sol)
            lib().sb200_result_free(C.byref(r))
        _check(rc, res)
        return res

    def initialize(self, matrix: SparseMatrix, b, options: SolverOptions | None = None) -> "NeumannState":
        """`SolverAlgorithm::initialize` (src/solver/neumann.rs:381-388): the stepping interface."""
        options = options or SolverOptions()
        b = _f64(b)
        guess = None if options.initial_guess is None else _f64(options.initial_guess)
        o = options._c(None if guess is None else guess.ctypes.data, 0 if guess is None else len(guess))
        h = C.c_void_p()
        _check(lib().sb200_neumann_initialize(self._h, matrix._h, _ptr(b), len(b), C.byref(o), C.byref(h)))
        return NeumannState(h, matrix)

    def solve_streaming(self, matrix: SparseMatrix, b, options: SolverOptions, callback) -> SolverResult:
        """`solve` with `SolverOptions::streaming(interval)`: `callback(partial: dict) -> bool | None` receives a
        PartialSolution (src/solver/mod.rs:198-217) every `streaming_interval` iterations; return True to stop."""
        b = _f64(b)
        guess = None if options.initial_guess is None else _f64(options.initial_guess)
        o = options._c(None if guess is None else guess.ctypes.data, 0 if guess is None else len(guess))

        def _cb(pp, _user):
            p = pp.contents
            sol = np.ctypeslib.as_array(p.solution, shape=(max(int(p.solution_len), 1),))[:int(p.solution_len)].copy()
            stop = callback({"iteration": int(p.iteration), "solution": sol, "residual_norm": p.residual_norm,
                             "converged": bool(p.converged),
                             "estimated_remaining": int(p.estimated_remaining) if p.has_estimated_remaining else None,
                             "timestamp_ms": p.timestamp_ms})
            return 1 if stop else 0

        cb = _STREAM_CB(_cb)
        r = _Result()
        rc = lib().sb200_solve_streaming(self._h, matrix._h, _ptr(b), len(b), C.byref(o), cb, None, C.byref(r))
        sol = None
        if r.solution:
            sol = np.ctypeslib.as_array(r.solution, shape=(max(int(r.solution_len), 1),))[:int(r.solution_len)].copy()
        res = SolverResult._from(r, sol)
        lib().sb200_result_free(C.byref(r))
        _check(rc, res)
        return res

    def solve_dev(self, matrix: SparseMatrix, b_ptr: int, n: int, x_ptr: int, options: SolverOptions | None = None,
                  stream: int = 0, guess_ptr: int | None = None) -> SolverResult:
        """Inputs/outputs already resident in HBM (raw device pointers, e.g. torch tensor.data_ptr())."""
        options = options or SolverOptions()
        o = options._c(guess_ptr, n if guess_ptr else 0)
        r = _Result()
        rc = lib().sb200_solve_dev(self._h, matrix._h, b_ptr, n, C.byref(o), x_ptr, stream, C.byref(r))
        res = SolverResult._from(r, None)
        _check(rc, res)
        return res


@dataclass
class OptimizedSolverConfig:
    """`OptimizedSolverConfig` (src/optimized_solver.rs:108-127)."""
    max_iterations: int = 1000
    tolerance: float = 1e-6
    enable_profiling: bool = False

    def _c(self):
        c = _CgConfig()
        lib().sb200_cg_config_default(C.byref(c))
        c.max_iterations, c.tolerance, c.enable_profiling = self.max_iterations, self.tolerance, int(self.enable_profiling)
        return c


@dataclass
class OptimizedSolverResult:
    """`OptimizedSolverResult` + `OptimizedSolverStats` (src/optimized_solver.rs:130-166) + extension fields."""
    solution: np.ndarray | None
    residual_norm: float
    iterations: int
    converged: bool
    breakdown: bool
    computation_time_ms: float
    matvec_count: int
    dot_product_count: int
    axpy_count: int
    total_flops: int
    average_bandwidth_gbs: float
    average_gflops: float
    device_time_ms: float
    kernel_launches: int
    h2d_bytes: int
    d2h_bytes: int
    spmv_kernel_ms: float
    spmv_kernel_count: int

    @staticmethod
    def _from(r: _CgResult, solution):
        return OptimizedSolverResult(solution, r.residual_norm, int(r.iterations), bool(r.converged), bool(r.breakdown),
                                     r.computation_time_ms, int(r.matvec_count), int(r.dot_product_count),
                                     int(r.axpy_count), int(r.total_flops), r.average_bandwidth_gbs, r.average_gflops,
                                     r.device_time_ms, int(r.kernel_launches), int(r.h2d_bytes), int(r.d2h_bytes),
                                     r.spmv_kernel_ms, int(r.spmv_kernel_count))

    def data(self):
        return self.solution


class OptimizedConjugateGradientSolver:
    """`OptimizedConjugateGradientSolver` (src/optimized_solver.rs:168-295); the loop shared with FastConjugateGradient
    (src/fast_solver.rs:111-178) and UltraFastCG (src/ultra_fast.rs:100-158), on the push path's SpMV kernel."""

    def __init__(self, config: OptimizedSolverConfig | None = None):
        self.config = config or OptimizedSolverConfig()
        self._last = None

    @staticmethod
    def new(config: OptimizedSolverConfig):
        return OptimizedConjugateGradientSolver(config)

    def solve(self, matrix: SparseMatrix, b, out=None) -> OptimizedSolverResult:
        b = _f64(b)
        c = self.config._c()
        r = _CgResult()
        if out is not None:
            rc = lib().sb200_cg_solve_into(matrix._h, _ptr(b), len(b), C.byref(c), _ptr(out), C.byref(r))
            res = OptimizedSolverResult._from(r, out)
        else:
            rc = lib().sb200_cg_solve(matrix._h, _ptr(b), len(b), C.byref(c), C.byref(r))
            sol = None
            if r.solution:
                sol = np.ctypeslib.as_array(r.solution, shape=(max(int(r.solution_len), 1),))[:int(r.solution_len)].copy()
            res = OptimizedSolverResult._from(r, sol)
            lib().sb200_cg_result_free(C.byref(r))
        _check(rc, res)
        self._last = res
        return res

    def solve_dev(self, matrix: SparseMatrix, b_ptr: int, n: int, x_ptr: int, stream: int = 0) -> OptimizedSolverResult:
        c = self.config._c()
        r = _CgResult()
        rc = lib().sb200_cg_solve_dev(matrix._h, b_ptr, n, C.byref(c), x_ptr, stream or None, C.byref(r))
        res = OptimizedSolverResult._from(r, None)
        _check(rc, res)
        self._last = res
        return res

    def get_last_iteration_count(self):
        """`get_last_iteration_count` returns stats.matvec_count in the reference (src/optimized_solver.rs:323-325)."""
        return self._last.matvec_count if self._last else 0


class NeumannState:
    """`NeumannState` behind `SolverAlgorithm::{step, is_converged, extract_solution, update_rhs}` and
    `SolverState::{residual_norm, matvec_count, error_bounds, reset}` (src/solver/neumann.rs:97-135, 350-462)."""

    def __init__(self, handle, matrix):
        self._h = handle
        self._matrix = matrix   # the state points into the matrix handle: keep it alive

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_state_free(self._h)
            self._h = None

    def step(self) -> int:
        out = C.c_int32()
        _check(lib().sb200_state_step(self._h, C.byref(out)))
        return out.value

    def is_converged(self) -> bool:
        out = C.c_int32()
        _check(lib().sb200_state_is_converged(self._h, C.byref(out)))
        return bool(out.value)

    def info(self) -> dict:
        i = _StateInfo()
        _check(lib().sb200_state_info(self._h, C.byref(i)))
        return {"dimension": i.dimension, "residual_norm": i.residual_norm, "matvec_count": i.matvec_count,
                "terms_computed": i.terms_computed, "series_converged": bool(i.series_converged),
                "last_term_norm": i.last_term_norm,
                "error_upper_bound": i.error_upper_bound if i.has_error_bounds else None, "memory_bytes": i.memory_bytes}

    def residual_norm(self) -> float:
        return self.info()["residual_norm"]

    def matvec_count(self) -> int:
        return self.info()["matvec_count"]

    def extract_solution(self) -> np.ndarray:
        x = np.zeros(self.info()["dimension"])
        _check(lib().sb200_state_extract_solution(self._h, _ptr(x), len(x)))
        return x

    def update_rhs(self, delta_b):
        """`update_rhs(&mut state, &[(index, delta)])`"""
        idx = _u64([i for i, _ in delta_b])
        dl = _f64([d for _, d in delta_b])
        _check(lib().sb200_state_update_rhs(self._h, _ptr(idx), _ptr(dl), len(dl)))

    def reset(self):
        _check(lib().sb200_state_reset(self._h))


@dataclass
class PushConfig:
    """`ForwardPushConfig` / `BackwardPushConfig` (src/solver/forward_push.rs:25-50, backward_push.rs:25-50)."""
    alpha: float = 0.15
    epsilon: float = 1e-6
    max_pushes: int = 1_000_000
    queue_threshold: float = 1e-8
    adaptive_threshold: bool = True

    def _c(self):
        c = _PushConfig()
        lib().sb200_push_config_default(C.byref(c))
        c.alpha, c.epsilon, c.max_pushes = self.alpha, self.epsilon, self.max_pushes
        c.queue_threshold, c.adaptive_threshold = self.queue_threshold, int(self.adaptive_threshold)
        return c


@dataclass
class PushResult:
    """`ForwardPushResult` / `BackwardPushResult` (src/solver/forward_push.rs:10-22) + extension fields."""
    estimate: np.ndarray
    residual: np.ndarray
    push_count: int
    nodes_visited: int
    residual_norm: float
    rounds: int
    kernel_launches: int
    device_time_ms: float
    dense_rounds: int = 0
    edges_touched: int = 0


class PushGraph:
    """`PushGraph` (src/graph/adjacency.rs:199-277), resident in HBM."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_push_graph_free(self._h)
            self._h = None

    @staticmethod
    def from_matrix(row_ptr, col_indices, values, n) -> "PushGraph":
        """`PushGraph::from_matrix(&CompressedSparseRow)`: row u holds the out-edges of u."""
        rp, ci, v = _u64(row_ptr), np.ascontiguousarray(col_indices, dtype=np.uint32), _f64(values)
        h = C.c_void_p()
        _check(lib().sb200_push_graph_from_csr(_ptr(rp), _ptr(ci), _ptr(v), n, C.byref(h)))
        return PushGraph(h)

    @staticmethod
    def from_edges(num_nodes, edges) -> "PushGraph":
        """`PushGraph::from_edges(num_nodes, &[(from, to, weight)])`"""
        f = _u64([e[0] for e in edges])
        t = _u64([e[1] for e in edges])
        w = _f64([e[2] for e in edges])
        h = C.c_void_p()
        _check(lib().sb200_push_graph_from_edges(num_nodes, _ptr(f), _ptr(t), _ptr(w), len(w), C.byref(h)))
        return PushGraph(h)

    def num_nodes(self):
        n = C.c_uint64()
        _check(lib().sb200_push_graph_info(self._h, C.byref(n), None))
        return n.value

    def num_edges(self):
        e = C.c_uint64()
        _check(lib().sb200_push_graph_info(self._h, None, C.byref(e)))
        return e.value

    def out_degree(self, node):
        d = C.c_double()
        _check(lib().sb200_push_graph_degrees(self._h, node, C.byref(d), None))
        return d.value

    def in_degree(self, node):
        d = C.c_double()
        _check(lib().sb200_push_graph_degrees(self._h, node, None, C.byref(d)))
        return d.value


class _PushSolver:
    _fn = None

    def __init__(self, graph: PushGraph, config: PushConfig | None = None):
        self.graph, self.config = graph, config or PushConfig()

    def _run(self, seeds) -> PushResult:
        s = _u64(np.atleast_1d(seeds))
        n = self.graph.num_nodes()
        est, res = np.zeros(n), np.zeros(n)
        c = self.config._c()
        st = _PushStats()
        _check(getattr(lib(), self._fn)(self.graph._h, C.byref(c), _ptr(s), len(s), _ptr(est), _ptr(res), C.byref(st)))
        return self._result(est, res, st)

    @staticmethod
    def _result(est, res, st) -> PushResult:
        return PushResult(est, res, int(st.push_count), int(st.nodes_visited), st.residual_norm, int(st.rounds),
                          int(st.kernel_launches), st.device_time_ms, int(st.dense_rounds), int(st.edges_touched))

    def _run_watch(self, fn, source, target, precision) -> PushResult:
        n = self.graph.num_nodes()
        est, res = np.zeros(n), np.zeros(n)
        c = self.config._c()
        st = _PushStats()
        _check(getattr(lib(), fn)(self.graph._h, C.byref(c), source, target, precision, _ptr(est), _ptr(res), C.byref(st)))
        return self._result(est, res, st)


class ForwardPushSolver(_PushSolver):
    """`ForwardPushSolver` (src/solver/forward_push.rs:52-328) as frontier-synchronous rounds on the device."""
    _fn = "sb200_forward_push"

    def solve_single_source(self, source) -> PushResult:
        return self._run([source])

    def solve_multi_source(self, sources) -> PushResult:
        return self._run(list(sources))

    def query_single_entry(self, source, target) -> float:
        r = self.solve_single_source(source)
        return float(r.estimate[target]) if target < len(r.estimate) else 0.0

    def solve_with_target(self, source, target, target_precision) -> PushResult:
        """`solve_with_target` (forward_push.rs:234-290): early stop once the target's estimate is settled"""
        return self._run_watch("sb200_forward_push_with_target", source, target, target_precision)

    def extrapolated_solution(self, result: PushResult) -> np.ndarray:
        return result.estimate + self.config.alpha * result.residual


class BackwardPushSolver(_PushSolver):
    """`BackwardPushSolver` (src/solver/backward_push.rs:52-330)."""
    _fn = "sb200_backward_push"

    def solve_single_target(self, target) -> PushResult:
        return self._run([target])

    def solve_multi_target(self, targets) -> PushResult:
        return self._run(list(targets))

    def query_transition_probability(self, source, target) -> float:
        r = self.solve_single_target(target)
        return float(r.estimate[source]) if source < len(r.estimate) else 0.0

    def solve_with_source(self, source, target, source_precision) -> PushResult:
        """`solve_with_source` (backward_push.rs:238-290)"""
        return self._run_watch("sb200_backward_push_with_source", source, target, source_precision)

    def reachability_probabilities(self, target) -> np.ndarray:
        return self.extrapolated_solution(self.solve_single_target(target))

    def extrapolated_solution(self, result: PushResult) -> np.ndarray:
        return result.estimate + self.config.alpha * result.residual

    def combine_with_forward(self, backward_result: PushResult, forward_estimate, forward_residual) -> float:
        """`combine_with_forward` (backward_push.rs:312-330)"""
        fe, fr = _f64(forward_estimate), _f64(forward_residual)
        be, br = _f64(backward_result.estimate), _f64(backward_result.residual)
        out = C.c_double()
        _check(lib().sb200_push_combine_with_forward(self.config.alpha, _ptr(be), _ptr(br), len(be), _ptr(fe), _ptr(fr),
                                                     len(fe), C.byref(out)))
        return out.value


class BidirectionalPushSolver:
    """`BidirectionalPushSolver` (src/solver/backward_push.rs:338-420)."""

    def __init__(self, graph: PushGraph, forward_config: PushConfig | None = None, backward_config: PushConfig | None = None):
        self.graph = graph
        self.forward_config, self.backward_config = forward_config or PushConfig(), backward_config or PushConfig()

    def _call(self, fn, source, target) -> float:
        fc, bc = self.forward_config._c(), self.backward_config._c()
        out = C.c_double()
        _check(getattr(lib(), fn)(self.graph._h, C.byref(fc), C.byref(bc), source, target, C.byref(out)))
        return out.value

    def solve_bidirectional(self, source, target) -> float:
        return self._call("sb200_bidirectional_push", source, target)

    def adaptive_solve(self, source, target) -> float:
        return self._call("sb200_bidirectional_adaptive_push", source, target)


@dataclass
class ForwardPushSolveResult:
    """result of the TS solver's `solveForwardPush` (src/core/solver.ts:513-521)"""
    solution: np.ndarray
    iterations: int
    residual: float
    converged: bool
    rounds: int
    max_residual: float
    method: str = "forward-push"


def forward_push_solve(matrix: "SparseMatrix", b, epsilon=1e-6, max_iterations=1000) -> ForwardPushSolveResult:
    """`SublinearSolver.solveForwardPush` (src/core/solver.ts:437-522) for A x = b (defaults: SolverConfig of the TS
    package). Raises SolverError ConvergenceFailure / NumericalInstability like the reference throws."""
    b = _f64(b)
    x = np.zeros(len(b))
    st = _AxbPushStats()
    rc = lib().sb200_forward_push_solve(matrix._h, _ptr(b), len(b), epsilon, max_iterations, _ptr(x), C.byref(st))
    res = ForwardPushSolveResult(x, int(st.iterations), st.residual_norm, bool(st.converged), int(st.rounds), st.max_residual)
    _check(rc, res)
    return res


class StreamingMatrix:
    """`StreamingMatrix` (src/matrix/optimized.rs:451-561): row chunks sized from a memory limit, held in pinned host
    memory and streamed through the GPU."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_streaming_matrix_free(self._h)
            self._h = None

    @staticmethod
    def from_triplets(rows, cols, vals, nrows, ncols, memory_limit_mb) -> "StreamingMatrix":
        r, c, v = _u64(rows), _u64(cols), _f64(vals)
        h = C.c_void_p()
        _check(lib().sb200_streaming_matrix_from_triplets(_ptr(r), _ptr(c), _ptr(v), len(v), nrows, ncols, memory_limit_mb,
                                                          C.byref(h)))
        return StreamingMatrix(h)

    def info(self) -> dict:
        v = [C.c_uint64() for _ in range(5)]
        _check(lib().sb200_streaming_matrix_info(self._h, *[C.byref(t) for t in v]))
        return dict(zip(("total_rows", "total_cols", "chunk_size", "num_chunks", "memory_usage"), (t.value for t in v)))

    def memory_usage(self) -> int:
        return self.info()["memory_usage"]

    def multiply_vector_streaming(self, x, callback) -> None:
        """`multiply_vector_streaming(x, |start_row, result| ..)`; callback(start_row, result) may return True to stop"""
        x = _f64(x)

        def _cb(start, ptr, n, _user):
            stop = callback(int(start), np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),))[:int(n)].copy())
            return 1 if stop else 0

        cb = _CHUNK_CB(_cb)
        _check(lib().sb200_streaming_matrix_multiply_vector(self._h, _ptr(x), len(x), cb, None))

    def multiply_vector(self, x) -> np.ndarray:
        y = np.zeros(self.info()["total_rows"])

        def put(start, part):
            y[start:start + len(part)] = part

        self.multiply_vector_streaming(x, put)
        return y


def push_iterations_dev(matrix: SparseMatrix, b_ptr: int, n: int, nterms: int, x_ptr: int = 0, t_ptr: int = 0,
                        stream: int = 0):
    """Bare recurrence on device pointers; returns (per-term norms, CUDA-event ms of the nterms launches)."""
    norms = np.zeros(max(nterms, 1))
    ms = C.c_float()
    _check(lib().sb200_push_iterations_dev(matrix._h, b_ptr, n, nterms, x_ptr or None, t_ptr or None, _ptr(norms),
                                           stream or None, C.byref(ms)))
    return norms[:nterms], ms.value


def solve_entry(matrix: SparseMatrix, b, rows, eps=0.01, nwalks=0, max_steps=0, seed=0):
    """Batched single-entry estimates x[rows[q]] by absorbing random walks; returns (estimate, variance)."""
    b, q = _f64(b), _u64(rows)
    est, var = np.zeros(len(q)), np.zeros(len(q))
    _check(lib().sb200_solve_entry(matrix._h, _ptr(b), len(b), _ptr(q), len(q), eps, nwalks, max_steps, seed,
                                   _ptr(est), _ptr(var)))
    return est, var


def solve_entry_replicas(matrices, b, rows, eps=0.01, nwalks=0, max_steps=0, seed=0):
    """`solve_entry` over several GPUs: `matrices` = one handle of the same matrix per device. Same estimates as
    `solve_entry` on one of them."""
    b, q = _f64(b), _u64(rows)
    est, var = np.zeros(len(q)), np.zeros(len(q))
    arr = (C.c_void_p * len(matrices))(*[m._h for m in matrices])
    _check(lib().sb200_solve_entry_replicas(arr, len(matrices), _ptr(b), len(b), _ptr(q), len(q), eps, nwalks, max_steps,
                                            seed, _ptr(est), _ptr(var)))
    return est, var


def gen_bench_csr(size, sparsity, row0=0, row1=None):
    """create_test_matrix / create_test_rhs (benches/performance_benchmarks.rs:12-43) as CSR, rows [row0,row1)."""
    row1 = size if row1 is None else row1
    nnz = C.c_uint64()
    _check(lib().sb200_gen_bench_csr(size, sparsity, row0, row1, None, None, None, None, C.byref(nnz)))
    rp = np.zeros(row1 - row0 + 1, np.uint64)
    ci, v, b = np.zeros(nnz.value, np.uint32), np.zeros(nnz.value), np.zeros(row1 - row0)
    _check(lib().sb200_gen_bench_csr(size, sparsity, row0, row1, _ptr(rp), _ptr(ci), _ptr(v), _ptr(b), C.byref(nnz)))
    return rp, ci, v, b


def partition_rows(nrows: int, world: int, rank: int):
    r0, r1 = C.c_uint64(), C.c_uint64()
    _check(lib().sb200_partition_rows(nrows, world, rank, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value


class Comm:
    """One rank of a row-partitioned multi-GPU job (NCCL communicator inside the library)."""

    def __init__(self, rank: int, world: int, unique_id: bytes, device: int):
        assert len(unique_id) == UNIQUE_ID_BYTES
        self.rank, self.world, self.device = rank, world, device
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        _check(lib().sb200_comm_init(rank, world, buf, device, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        _check(lib().sb200_comm_unique_id(buf))
        return buf.raw

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sb200_comm_free(self._h)
            self._h = None

    def matrix_from_csr(self, n_global, row0, row1, row_ptr, col_indices, values) -> SparseMatrix:
        rp, ci, v = _u64(row_ptr), np.ascontiguousarray(col_indices, dtype=np.uint32), _f64(values)
        h = C.c_void_p()
        _check(lib().sb200_dist_matrix_from_csr(self._h, n_global, row0, row1, _ptr(rp), _ptr(ci), _ptr(v), C.byref(h)))
        return SparseMatrix(h)

    def solve(self, solver: NeumannSolver, m_local: SparseMatrix, b_local, options: SolverOptions | None = None):
        options = options or SolverOptions()
        b = _f64(b_local)
        x = np.zeros(len(b))
        guess = None if options.initial_guess is None else _f64(options.initial_guess)   # this rank's rows of the guess
        o = options._c(None if guess is None else guess.ctypes.data, 0 if guess is None else len(guess))
        r = _Result()
        rc = lib().sb200_dist_solve(self._h, solver._h, m_local._h, _ptr(b), len(b), C.byref(o), _ptr(x), C.byref(r))
        res = SolverResult._from(r, x)
        _check(rc, res)
        return res

    def push_iterations_dev(self, m_local: SparseMatrix, b_ptr: int, nlocal: int, nterms: int, x_ptr: int = 0):
        norms = np.zeros(max(nterms, 1))
        ms = C.c_float()
        _check(lib().sb200_dist_push_iterations_dev(self._h, m_local._h, b_ptr, nlocal, nterms, x_ptr or None,
                                                    _ptr(norms), C.byref(ms)))
        return norms[:nterms], ms.value
