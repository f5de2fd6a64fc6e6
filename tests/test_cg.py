"""Conjugate gradient on the push path's SpMV.

CPU part (`-m "not gpu"`): the oracle's restatement of OptimizedConjugateGradientSolver::solve
(src/optimized_solver.rs:182-295) against the reference's unit test and against golden vectors produced by the
reference-authored numpy CG (tests/golden/make_golden_cg.py).
GPU part: sb200_cg_solve* through the C ABI against the oracle on the same inputs.

Tolerances: CG's dot products are sums over n terms; the device reduces them in a fixed tree, the oracle sequentially,
numpy through BLAS.  Iterates therefore agree to rounding amplified by the recurrence: rtol 1e-9 on early iterates and
on the converged solution (all systems here have condition numbers < 1e4), identical iteration counts.
"""
import glob
import os

import numpy as np
import pytest

import sublinear_b200 as sb

CG_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cg_*.npz")))


def load(O, path):
    g = np.load(path)
    n = int(g["n"])
    return g, O.Csr.from_triplets(g["rows"], g["cols"], g["vals"], n, n)


def sym_dd(n, k, seed):
    """symmetric, strictly diagonally dominant sparse test system (SPD), ~2k off-diagonals per row"""
    rng = np.random.default_rng(seed)
    r = np.repeat(np.arange(n), k)
    c = rng.integers(0, n, n * k)
    v = rng.uniform(-1, 1, n * k)
    keep = r != c
    r, c, v = r[keep], c[keep], v[keep]
    rows = np.concatenate([r, c])
    cols = np.concatenate([c, r])
    vals = np.concatenate([v, v])
    offsum = np.bincount(rows, weights=np.abs(vals), minlength=n)
    rows = np.concatenate([rows, np.arange(n)])
    cols = np.concatenate([cols, np.arange(n)])
    vals = np.concatenate([vals, 1.5 * offsum + 1.0])
    return rows, cols, vals, rng.uniform(-5, 5, n)


# ---- oracle vs the reference's own answers (CPU) --------------------------------------------------------------

def test_oracle_cg_rust_unit_test(oracle):
    # src/optimized_solver.rs:398-418: converged, residual < 1e-6, iterations > 0, ||A x - b|| < 1e-10
    O = oracle
    A = O.Csr.from_dense(np.array([[4., 1.], [1., 3.]]))
    for dot in (O.DOT_SEQUENTIAL, O.DOT_CHUNK4, O.DOT_CHUNK8):
        r = O.cg_solve(A, [1., 2.], dot_variant=dot)
        assert r.converged and r.residual_norm < 1e-6 and r.iterations > 0
        assert np.linalg.norm(A.multiply_vector(r.solution) - [1., 2.]) < 1e-10
        assert r.matvec_count == r.iterations == 2 and r.total_flops == 2 * 4 * 2 + 2 * 2 * 6
    # "Matrix must be square" / length mismatch (:188-193)
    with pytest.raises(O.OracleError) as e:
        O.cg_solve(O.Csr.from_triplets([0], [1], [1.0], 2, 3), [1., 2.])
    assert e.value.code == O.ERR_INVALID_INPUT
    with pytest.raises(O.OracleError) as e:
        O.cg_solve(A, [1., 2., 3.])
    assert e.value.code == O.ERR_DIMENSION_MISMATCH


def golden_expectations(g, tol):
    """pass count the Rust loop reaches for `tol` (first k with ||r_k|| <= tol, checked at the top of pass k+1) and the
    passes that are safely above the |p.Ap| < 1e-16 guard (||r|| > 1e-6)"""
    hist = g["residual_history"]
    k_stop = int(np.argmax(hist <= tol)) + 1 if (hist <= tol).any() else None
    safe = [int(k) for k in g["kept_passes"] if hist[int(k) - 1] > 1e-6 or int(k) == k_stop]
    return k_stop, safe


@pytest.mark.parametrize("path", CG_GOLDEN, ids=[os.path.basename(p)[:-4] for p in CG_GOLDEN])
def test_oracle_cg_matches_reference_numpy_cg(oracle, path):
    O = oracle
    g, A = load(O, path)
    b = g["b"]
    k_stop, safe = golden_expectations(g, 1e-6)
    for k in safe:                                                     # x after k passes, tolerance 0 = never stop early
        xk = g["kept_iterates"][k - 1]
        r = O.cg_solve(A, b, max_iterations=k, tolerance=0.0)
        assert r.iterations == k
        np.testing.assert_allclose(r.solution, xk, rtol=1e-9, atol=1e-9 * np.abs(xk).max())
        np.testing.assert_allclose(r.residual_norm, g["residual_history"][k - 1], rtol=1e-6, atol=1e-15)
    r = O.cg_solve(A, b, tolerance=1e-6)                               # OptimizedSolverConfig::default
    assert r.converged and r.iterations == k_stop
    xs = g["kept_iterates"][k_stop - 1]
    np.testing.assert_allclose(r.solution, xs, rtol=1e-9, atol=1e-9 * np.abs(xs).max())


def test_oracle_cg_breakdown_guard_limits_accuracy(oracle):
    """Reference quirk worth a regression test: the absolute guard `pap.abs() < 1e-16 -> break`
    (src/optimized_solver.rs:234-236) fires once ||p||^2 lambda < 1e-16, i.e. at ||r|| ~ 1e-8..1e-9, so a tolerance
    below that is never reached: the loop leaves through the guard with converged = false."""
    O = oracle
    g, A = load(O, [p for p in CG_GOLDEN if "dd_symmetric_n100" in p][0])
    r = O.cg_solve(A, g["b"], tolerance=1e-10)
    assert not r.converged and r.iterations < 1000 and r.matvec_count == r.iterations + 1
    assert 1e-10 < r.residual_norm < 1e-7
    assert int(g["iterations"]) > r.iterations                          # the unguarded numpy CG went on to 1e-10


def test_oracle_cg_loop_exits(oracle):
    O = oracle
    A = O.Csr.from_dense(np.array([[4., 1.], [1., 3.]]))
    r = O.cg_solve(A, [1., 2.], max_iterations=1)                       # runs out of iterations: not converged (:217)
    assert not r.converged and r.iterations == 1 and r.residual_norm > 0
    r = O.cg_solve(A, [0., 0.])                                         # rsold = 0 <= tol^2 before any work
    assert r.converged and r.iterations == 0 and r.matvec_count == 0 and (r.solution == 0).all()
    r = O.cg_solve(A, [1., 2.], max_iterations=0)
    assert not r.converged and r.iterations == 0
    Z = O.Csr.from_triplets([0], [1], [1.0], 2, 2)                      # p.Ap = 0 -> `break` (:234-236)
    r = O.cg_solve(Z, [1., 0.])
    assert not r.converged and r.iterations == 0 and r.matvec_count == 1 and (r.solution == 0).all()


def test_oracle_cg_vs_scipy(oracle):
    import scipy.sparse.linalg as spla
    O = oracle
    rows, cols, vals, b = sym_dd(2000, 4, 3)
    A = O.Csr.from_triplets(rows, cols, vals, 2000, 2000)
    r = O.cg_solve(A, b, tolerance=1e-7)
    x = spla.spsolve(A.to_scipy().tocsc(), b)
    assert r.converged
    np.testing.assert_allclose(r.solution, x, rtol=1e-6, atol=1e-7)


# ---- the device path vs the oracle (GPU) -----------------------------------------------------------------------

def to_gpu(A):
    return sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, A.nrows, A.ncols)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CG_GOLDEN, ids=[os.path.basename(p)[:-4] for p in CG_GOLDEN])
def test_gpu_cg_matches_oracle_and_golden(oracle, path):
    O = oracle
    g, A = load(O, path)
    b = g["b"]
    m = to_gpu(A)
    k_stop, safe = golden_expectations(g, 1e-6)
    for k in safe:
        xk = g["kept_iterates"][k - 1]
        r = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(k, 0.0)).solve(m, b)
        o = O.cg_solve(A, b, max_iterations=k, tolerance=0.0)
        assert r.iterations == o.iterations == k and not r.converged
        scale = np.abs(xk).max()
        np.testing.assert_allclose(r.solution, o.solution, rtol=1e-9, atol=1e-9 * scale)
        np.testing.assert_allclose(r.solution, xk, rtol=1e-9, atol=1e-9 * scale)
        np.testing.assert_allclose(r.residual_norm, o.residual_norm, rtol=1e-6, atol=1e-15)
    r = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig()).solve(m, b)     # 1000 / 1e-6
    o = O.cg_solve(A, b)
    assert r.converged and r.iterations == o.iterations == k_stop
    assert r.matvec_count == o.matvec_count and r.total_flops == o.total_flops
    xs = g["kept_iterates"][k_stop - 1]
    np.testing.assert_allclose(r.solution, xs, rtol=1e-9, atol=1e-9 * np.abs(xs).max())
    # the |p.Ap| < 1e-16 guard (reference quirk): same exit as the oracle when the tolerance is out of reach
    r = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(1000, 1e-12)).solve(m, b)
    o = O.cg_solve(A, b, tolerance=1e-12)
    if not o.converged:
        assert not r.converged and r.breakdown and abs(r.iterations - o.iterations) <= 1


@pytest.mark.gpu
def test_gpu_cg_rust_unit_tests_and_errors(oracle):
    m = sb.SparseMatrix.from_dense([[4., 1.], [1., 3.]])                # src/optimized_solver.rs:380-435
    s = sb.OptimizedConjugateGradientSolver.new(sb.OptimizedSolverConfig())
    r = s.solve(m, [1., 2.])
    assert r.converged and r.residual_norm < 1e-6 and r.iterations > 0
    assert np.linalg.norm(m.multiply_vector(r.solution) - [1., 2.]) < 1e-10
    assert r.matvec_count > 0 and r.dot_product_count > 0 and r.total_flops > 0
    assert s.get_last_iteration_count() == r.matvec_count == 2
    assert r.kernel_launches >= 1 + 3 * r.iterations
    # the loop's three other exits
    r = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(max_iterations=1)).solve(m, [1., 2.])
    assert not r.converged and r.iterations == 1
    r = s.solve(m, [0., 0.])
    assert r.converged and r.iterations == 0 and r.matvec_count == 0 and (r.solution == 0).all()
    r = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(max_iterations=0)).solve(m, [1., 2.])
    assert not r.converged and r.iterations == 0 and (r.solution == 0).all()
    z = sb.SparseMatrix.from_triplets([0], [1], [1.0], 2, 2)
    r = s.solve(z, [1., 0.])
    assert not r.converged and r.breakdown and r.iterations == 0 and r.matvec_count == 1 and (r.solution == 0).all()
    with pytest.raises(sb.SolverError) as e:
        s.solve(sb.SparseMatrix.from_triplets([0], [1], [1.0], 2, 3), [1., 2.])
    assert e.value.variant == "InvalidInput" and "square" in str(e.value)
    with pytest.raises(sb.SolverError) as e:
        s.solve(m, [1., 2., 3.])
    assert e.value.variant == "DimensionMismatch"


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["0", "1"])
@pytest.mark.parametrize("n,k", [(1000, 3), (50_000, 5), (300_000, 4)])
def test_gpu_cg_scaled_vs_oracle(oracle, monkeypatch, n, k, layout):
    O = oracle
    monkeypatch.setenv("SUBLINEAR_B200_SELL", layout)
    rows, cols, vals, b = sym_dd(n, k, n + k)
    A = O.Csr.from_triplets(rows, cols, vals, n, n)
    m = to_gpu(A)
    assert m.storage_info()["layout"] == int(layout)
    cfg = sb.OptimizedSolverConfig(1000, 1e-6, enable_profiling=True)
    r = sb.OptimizedConjugateGradientSolver(cfg).solve(m, b)
    o = O.cg_solve(A, b, tolerance=1e-6)
    assert r.converged and o.converged and r.iterations == o.iterations
    assert r.spmv_kernel_count == r.matvec_count == o.matvec_count
    np.testing.assert_allclose(r.solution, o.solution, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r.residual_norm, o.residual_norm, rtol=1e-3, atol=1e-14)
    res = np.linalg.norm(A.multiply_vector(r.solution) - b)
    assert res < 2e-6
    # the other two reference variants only reorder the dot products: same iterate to rounding
    o4 = O.cg_solve(A, b, tolerance=1e-6, dot_variant=O.DOT_CHUNK4, spmv_variant=O.SPMV_SIMD4)
    assert o4.iterations == r.iterations
    np.testing.assert_allclose(r.solution, o4.solution, rtol=1e-8, atol=1e-11)


@pytest.mark.gpu
def test_gpu_cg_device_pointers_match_host_call(oracle):
    import torch
    rows, cols, vals, b = sym_dd(20_000, 4, 11)
    m = sb.SparseMatrix.from_triplets(rows, cols, vals, 20_000, 20_000)
    s = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(500, 1e-6))
    r_host = s.solve(m, b)
    bd = torch.tensor(b, device="cuda")
    xd = torch.empty_like(bd)
    r_dev = s.solve_dev(m, bd.data_ptr(), len(b), xd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert r_dev.iterations == r_host.iterations and r_dev.converged
    assert np.array_equal(xd.cpu().numpy(), r_host.solution)            # same kernels, same reduction order: bit-equal
