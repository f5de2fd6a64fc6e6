"""Pin the CPU oracle (oracle/) against every known answer the reference holds for the Neumann / push
path and against golden vectors produced by reference-authored code
(tests/golden/make_golden.py, scripts/linear_systems/iterative_solvers.py:17-105).  CPU only."""
import glob
import os

import numpy as np
import pytest


def csr(O, dense):
    return O.Csr.from_dense(np.asarray(dense, dtype=np.float64))


# ---- reference unit tests restated as known-answer tests ---------------------------------------

def test_spmv_known_answers(oracle):
    O = oracle
    # sparse.rs:923-933, matrix/mod.rs:590-600, simd_ops.rs:259-268, optimized.rs:586-599
    for variant in (O.SPMV_SCALAR, O.SPMV_SIMD4, O.SPMV_PARALLEL):
        assert csr(O, [[2, 1], [1, 3]]).multiply_vector([1, 2], variant).tolist() == [4.0, 7.0]
        # optimized_solver.rs:389-396 ; fast_solver.rs:260-272
        assert csr(O, [[4, 1], [1, 3]]).multiply_vector([1, 2], variant).tolist() == [6.0, 7.0]
        assert csr(O, [[4, 1], [2, 3]]).multiply_vector([1, 2], variant).tolist() == [6.0, 8.0]


def test_dot_axpy_known_answers(oracle):
    import ctypes as C
    L = oracle.lib()
    x = np.array([1., 2., 3., 4., 5.]); y = np.array([2., 3., 4., 5., 6.])
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert L.orc_dot_simd4(p(x), p(y), 5) == 70.0          # simd_ops.rs:271-277
    xx = np.array([1., 2., 3., 4.]); yy = np.ones(4)
    L.orc_axpy_simd4(2.0, p(xx), p(yy), 4)                  # simd_ops.rs:279-285
    assert yy.tolist() == [3., 5., 7., 9.]


def test_csr_get_and_nnz(oracle):
    # sparse.rs:910-920
    m = oracle.Csr.from_triplets([0, 0, 1, 2, 2], [0, 2, 1, 0, 2], [1., 2., 3., 4., 5.], 3, 3)
    assert m.nnz == 5
    assert m.get(0, 0) == 1.0 and m.get(0, 2) == 2.0 and m.get(1, 1) == 3.0
    assert m.get(0, 1) is None and m.get(5, 0) is None
    assert m.row_ptr.tolist() == [0, 2, 3, 5]


def test_from_triplets_semantics(oracle):
    O = oracle
    # zeros dropped (sparse.rs:535-541); stable (row,col) sort keeps duplicates in input order (sparse.rs:95)
    m = O.Csr.from_triplets([1, 0, 1, 1, 0], [1, 1, 0, 1, 0], [5., 0., 2., 7., 3.], 2, 2)
    assert m.row_ptr.tolist() == [0, 1, 4]
    assert m.col_indices.tolist() == [0, 0, 1, 1]
    assert m.values.tolist() == [3., 2., 5., 7.]
    assert m.multiply_vector([1., 1.]).tolist() == [3., 14.]           # SpMV sums duplicates
    # empty (sparse.rs:81-87)
    e = O.Csr.from_triplets([], [], [], 3, 3)
    assert e.nnz == 0 and e.row_ptr.tolist() == [0, 0, 0, 0]
    assert e.multiply_vector([1., 2., 3.]).tolist() == [0., 0., 0.]
    # validation (matrix/mod.rs:166-187)
    with pytest.raises(O.OracleError) as ei:
        O.Csr.from_triplets([2], [0], [1.], 2, 2)
    assert ei.value.code == O.ERR_INDEX_OUT_OF_BOUNDS
    with pytest.raises(O.OracleError) as ei:
        O.Csr.from_triplets([0], [2], [1.], 2, 2)
    assert ei.value.code == O.ERR_INDEX_OUT_OF_BOUNDS
    with pytest.raises(O.OracleError) as ei:
        O.Csr.from_triplets([0], [0], [float("nan")], 2, 2)
    assert ei.value.code == O.ERR_INVALID_INPUT


def test_multiply_vector_dimension_checks(oracle):
    O = oracle
    m = O.Csr.from_triplets([0, 1], [0, 2], [1., 2.], 2, 3)      # 2x3 like matrix/mod.rs:617-619
    with pytest.raises(O.OracleError) as ei:
        m.multiply_vector([1., 2.])
    assert ei.value.code == O.ERR_DIMENSION_MISMATCH
    with pytest.raises(O.OracleError) as ei:
        m.multiply_vector([1., 2., 3.], ylen=3)
    assert ei.value.code == O.ERR_DIMENSION_MISMATCH
    assert m.multiply_vector([1., 2., 3.]).tolist() == [1., 6.]


def test_diagonal_dominance(oracle):
    O = oracle
    assert csr(O, [[5, 1], [2, 7]]).is_diagonally_dominant()        # matrix/mod.rs:603-607
    assert not csr(O, [[1, 3], [2, 2]]).is_diagonally_dominant()    # matrix/mod.rs:609-613
    assert csr(O, [[4, 1], [2, 5]]).is_diagonally_dominant()        # matrix/mod.rs:577-587
    assert csr(O, [[2, -2], [1, 1]]).is_diagonally_dominant()       # equality allowed (mod.rs:480)
    assert not csr(O, [[0, 1], [1, 2]]).is_diagonally_dominant()    # missing diagonal counts as 0
    assert csr(O, [[1, 3], [2, 2]]).first_non_dominant_row() == 0


def test_neumann_rejects_non_dd(oracle):
    O = oracle
    m = csr(O, [[1, 3], [2, 1]])                                    # neumann.rs:609-631
    for mode in (O.MODE_CORRECT, O.MODE_REF_COMPAT):
        with pytest.raises(O.OracleError) as ei:
            O.neumann_solve(m, [4., 3.], max_terms=20, mode=mode)
        assert ei.value.code == O.ERR_NOT_DIAGONALLY_DOMINANT


def test_neumann_state_initialisation(oracle):
    O = oracle
    # neumann.rs:633-648: diag(2,3), b=[4,6] -> dinv=[0.5,1/3], c=[2,2]; a diagonal system is solved by term 0
    m = O.Csr.from_triplets([0, 1], [0, 1], [2., 3.], 2, 2)
    r = O.neumann_solve(m, [4., 6.], mode=O.MODE_CORRECT)
    assert r.solution.tolist() == [2.0, 2.0]
    x, t, norms, _ = O.push_iterations(m, [4., 6.], 0)
    assert t.tolist() == [4. * 0.5, 6. * (1. / 3.)]


def test_neumann_simple_system_both_modes(oracle):
    O = oracle
    m = csr(O, [[4, 1], [1, 3]])                                    # neumann.rs:576-607
    r = O.neumann_solve(m, [5., 4.], max_terms=20, mode=O.MODE_CORRECT)
    assert r.converged and abs(r.solution[0] - 1.0) < 1e-7 and abs(r.solution[1] - 1.0) < 1e-7
    assert abs(r.solution[0] - 1.0) < 0.1 and abs(r.solution[1] - 1.0) < 0.1   # the reference's own assert
    # ref_compat reproduces the literal Rust control flow: x_true + D^-1 b, 17 terms, 21 matvecs
    c = O.neumann_solve(m, [5., 4.], max_terms=20, mode=O.MODE_REF_COMPAT)
    assert c.converged and c.series_converged
    assert c.terms_computed == 17 and c.iterations == 17 and c.matvec_count == 21
    np.testing.assert_allclose(c.solution, [1 + 5 / 4, 1 + 4 / 3], rtol=0, atol=1e-8)
    assert abs(c.residual_norm - 12.8198) < 1e-3


def test_neumann_missing_or_zero_diagonal(oracle):
    O = oracle
    m = O.Csr.from_triplets([0, 1], [0, 1], [1., 1e-15], 2, 2)
    with pytest.raises(O.OracleError) as ei:
        O.neumann_solve(m, [1., 1.])
    assert ei.value.code == O.ERR_INVALID_SPARSE_MATRIX
    m = O.Csr.from_triplets([0], [0], [1.], 2, 2)   # row 1 empty: DD holds (0>=0), diagonal missing
    with pytest.raises(O.OracleError) as ei:
        O.neumann_solve(m, [1., 1.])
    assert ei.value.code == O.ERR_INVALID_SPARSE_MATRIX
    m = O.Csr.from_triplets([0, 1], [0, 1], [1., 1.], 2, 2)
    with pytest.raises(O.OracleError) as ei:
        O.neumann_solve(m, [1., 1., 1.])
    assert ei.value.code == O.ERR_DIMENSION_MISMATCH
    m = O.Csr.from_triplets([0, 1], [0, 1], [1., 1.], 2, 3)
    with pytest.raises(O.OracleError) as ei:
        O.neumann_solve(m, [1., 1.])
    assert ei.value.code == O.ERR_INVALID_INPUT


def test_max_terms_quirk_and_convergence_failure(oracle):
    O = oracle
    # Appendix A quirk 3: terms stop at max_terms, the loop spins to max_iterations, then errors
    m = csr(O, [[4, 1], [1, 3]])
    r = O.neumann_solve(m, [5., 4.], max_terms=3, max_iterations=40, mode=O.MODE_REF_COMPAT, raise_on_error=False)
    assert r.status == O.ERR_CONVERGENCE_FAILURE and r.terms_computed == 3 and r.iterations == 40
    assert r.matvec_count == 2 + 8 + 1        # 2 term SpMVs + residual at it=0,5,..,35 + final


def test_norm_known_answers(oracle):
    import ctypes as C
    L = oracle.lib()
    v = np.array([3., -4.])                                        # solver/mod.rs:584-595
    p = v.ctypes.data_as(C.POINTER(C.c_double))
    assert L.orc_l1_norm(p, 2) == 7.0 and L.orc_l2_norm(p, 2) == 5.0 and L.orc_linf_norm(p, 2) == 4.0


def test_ts_lcg_stream(oracle):
    # src/core/utils.ts:161-168: state = (state*1664525 + 1013904223) mod 2^32
    s, exp = 42, []
    for _ in range(4):
        s = (s * 1664525 + 1013904223) % 2 ** 32
        exp.append(s / 2 ** 32)
    assert oracle.ts_lcg(42, 4) == exp


# ---- golden vectors from reference-authored Jacobi ---------------------------------------------

GOLDEN_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "jacobi_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def test_oracle_matches_reference_jacobi(oracle, path):
    O = oracle
    g = np.load(path)
    n = int(g["n"])
    A = O.Csr.from_triplets(g["rows"], g["cols"], g["vals"], n, n)
    b = g["b"]
    # k Jacobi sweeps from 0 == Neumann partial sum of k terms == term 0 + (k-1) push iterations
    for sweeps, xk in zip(g["kept_sweeps"], g["kept_iterates"]):
        for variant in (O.SPMV_SCALAR, O.SPMV_SIMD4, O.SPMV_PARALLEL):
            x, _, _, _ = O.push_iterations(A, b, int(sweeps) - 1, variant)
            np.testing.assert_allclose(x, xk, rtol=1e-12, atol=1e-13 * np.abs(xk).max())
    # residual history: ||A x_k - b||_2 after each sweep
    hist = g["residual_history"]
    k = len(hist)
    x, _, _, _ = O.push_iterations(A, b, k - 1)
    r = np.linalg.norm(A.to_scipy() @ x - b)
    assert abs(r - hist[-1]) <= 1e-6 * hist[0] * 1e-6 + 1e-3 * hist[-1] + 1e-14
    np.testing.assert_allclose(x, g["solution"], rtol=1e-12, atol=1e-13 * np.abs(x).max())
    # full solver, correct mode, tight enough to pass the golden's tolerance
    res = O.neumann_solve(A, b, mode=O.MODE_CORRECT, series_tolerance=1e-14, tolerance=1e-13, max_terms=500)
    np.testing.assert_allclose(res.solution, g["solution"], rtol=0, atol=50 * float(g["tol"]))
    assert np.linalg.norm(A.to_scipy() @ res.solution - b) <= 10 * float(g["tol"])


def test_c1_fixture_solve_matches_survey_scratch(oracle, golden_dir):
    """Config C1: tests/data/test-matrix.json, b = 1 (SURVEY.md §8c last bullet)."""
    O = oracle
    g = np.load(os.path.join(golden_dir, "jacobi_c1_test_matrix_ones.npz"))
    n = int(g["n"])
    A = O.Csr.from_triplets(g["rows"], g["cols"], g["vals"], n, n)
    assert (n, A.nnz) == (1000, 300650) and A.is_diagonally_dominant()
    b = g["b"]
    c = O.neumann_solve(A, b, mode=O.MODE_REF_COMPAT)
    assert (c.terms_computed, c.matvec_count) == (6, 8)
    import scipy.sparse.linalg as spl
    xt = spl.spsolve(A.to_scipy().tocsc(), b)
    dinv = 1.0 / A.to_scipy().diagonal()
    assert np.abs(c.solution - (xt + dinv * b)).max() <= 1.5e-11
    r = O.neumann_solve(A, b, mode=O.MODE_CORRECT)
    assert r.converged
    assert np.linalg.norm(A.to_scipy() @ r.solution - b) / np.linalg.norm(b) < 5e-9
    np.testing.assert_allclose(r.solution, xt, rtol=0, atol=1e-10)


# ---- generators & cross-checks -----------------------------------------------------------------

def test_gen_bench_matches_python_restatement(oracle):
    """Independent pure-Python restatement of benches/performance_benchmarks.rs:12-43."""
    O = oracle
    size, sparsity = 200, 0.05
    k = min(size, int(max(size * sparsity, 3.0)))
    assert O.gen_bench_k(size, sparsity) == k
    M = 2 ** 64
    rows, cols, vals = [], [], []
    for i in range(size):
        d = 10.0 + i * 0.01
        rows.append(i); cols.append(i); vals.append(d)
        mo = d / (k * 2.0)
        rng = (i * 1664525 + 1013904223) % M
        for _ in range(1, k):
            rng = (rng * 1664525 + 1013904223) % M
            col = rng % size
            if col != i:
                rng = (rng * 1664525 + 1013904223) % M
                rows.append(i); cols.append(col); vals.append((float(rng) / float(2 ** 64 - 1)) * mo)
    r, c, v, b = O.gen_bench_triplets(size, sparsity)
    assert r.tolist() == rows and c.tolist() == cols and v.tolist() == vals
    assert b.tolist() == [1.0 + i * 0.001 for i in range(size)]
    A, b2 = O.gen_bench_csr(size, sparsity)
    A2 = O.Csr.from_triplets(r, c, v, size, size)
    assert A.values.tolist() == A2.values.tolist() and A.col_indices.tolist() == A2.col_indices.tolist()
    assert A.row_ptr.tolist() == A2.row_ptr.tolist() and b2.tolist() == b.tolist()
    # row slices are the same rows
    S, bs = O.gen_bench_csr(size, sparsity, 50, 120)
    lo, hi = A.row_ptr[50], A.row_ptr[120]
    assert S.values.tolist() == A.values[lo:hi].tolist() and bs.tolist() == b[50:120].tolist()
    assert S.ncols == size and S.nrows == 70


@pytest.mark.parametrize("gen", ["bench", "ultra"])
def test_oracle_vs_scipy_spsolve(oracle, gen):
    import scipy.sparse.linalg as spl
    O = oracle
    n = 2000
    if gen == "bench":
        A, b = O.gen_bench_csr(n, 0.005)
    else:
        r, c, v, b = O.gen_ultra_triplets(n, 0.004)
        A = O.Csr.from_triplets(r, c, v, n, n)
    xt = spl.spsolve(A.to_scipy().tocsc(), b)
    for variant in (O.SPMV_SCALAR, O.SPMV_SIMD4, O.SPMV_PARALLEL):
        res = O.neumann_solve(A, b, spmv_variant=variant)
        assert res.converged
        assert np.linalg.norm(A.to_scipy() @ res.solution - b) / np.linalg.norm(b) < 1e-6
        np.testing.assert_allclose(res.solution, xt, rtol=1e-7)
    c = O.neumann_solve(A, b, mode=O.MODE_REF_COMPAT)
    dinv = 1.0 / A.to_scipy().diagonal()
    np.testing.assert_allclose(c.solution, xt + dinv * b, rtol=1e-7)


def test_initial_guess_modes(oracle):
    import scipy.sparse.linalg as spl
    O = oracle
    A, b = O.gen_bench_csr(500, 0.02)
    xt = spl.spsolve(A.to_scipy().tocsc(), b)
    x0 = xt + 0.01 * np.sin(np.arange(500))
    r = O.neumann_solve(A, b, initial_guess=x0, mode=O.MODE_CORRECT)
    np.testing.assert_allclose(r.solution, xt, rtol=1e-7)
    # reference quirk: x0 + sum M^k c (Appendix A)
    c = O.neumann_solve(A, b, initial_guess=x0, mode=O.MODE_REF_COMPAT)
    np.testing.assert_allclose(c.solution, x0 + xt, rtol=1e-7)
    with pytest.raises(O.OracleError) as ei:
        O.neumann_solve(A, b, initial_guess=x0[:10])
    assert ei.value.code == O.ERR_DIMENSION_MISMATCH


def test_error_bounds(oracle):
    O = oracle
    A, b = O.gen_bench_csr(500, 0.02)
    r = O.neumann_solve(A, b, compute_error_bounds=True, mode=O.MODE_REF_COMPAT)
    assert r.series_converged and r.error_bound is not None and 0 <= r.error_bound < 1e-6


def test_pagerank_system_matches_dense_restatement(oracle):
    """Dense restatement of computePageRank (src/core/solver.ts:664-722) on a small digraph with a dangling node."""
    O = oracle
    rng = np.random.default_rng(7)
    n, alpha = 40, 0.85
    adj = (rng.random((n, n)) < 0.1).astype(np.float64)
    adj[5, :] = 0.0                       # dangling
    adj[3, 3] = 1.0                       # self loop
    out = adj.sum(axis=1)
    S = np.eye(n)
    for i in range(n):
        for j in range(n):
            if out[j] > 0:
                S[i, j] -= alpha * adj[j, i] / out[j]
    src, dst = np.nonzero(adj)
    M, rhs = O.pagerank_system(src, dst, n, alpha)
    np.testing.assert_allclose(M.to_scipy().toarray(), S, rtol=0, atol=1e-15)
    assert np.allclose(rhs, (1 - alpha) / n)
    assert M.is_col_diagonally_dominant()
    res = O.neumann_solve(M, rhs, dominance=O.DOM_ROW_OR_COL, max_terms=200, series_tolerance=1e-12, tolerance=1e-10)
    np.testing.assert_allclose(res.solution, np.linalg.solve(S, rhs), rtol=1e-8)
    # the documented 4-node example graph shape (tests/mcp/mcp-tool-tests.js) stays a valid probability-like vector
    assert (res.solution > 0).all()


def test_solve_entry_unbiased(oracle):
    import scipy.sparse.linalg as spl
    O = oracle
    A, b = O.gen_bench_csr(300, 0.03)
    xt = spl.spsolve(A.to_scipy().tocsc(), b)
    rows = np.array([0, 17, 299])
    est, var = O.solve_entry(A, b, rows, nwalks=20000, seed=3)
    se = np.sqrt(var / 20000)
    assert (np.abs(est - xt[rows]) < 5 * se + 1e-12).all()
    est2, _ = O.solve_entry(A, b, rows, nwalks=20000, seed=3)
    assert est.tolist() == est2.tolist()
    # mixed-sign system (the documented 3x3 MCP example): true x_1 = 0.7561
    m = csr(O, [[4, -1, 0], [-1, 4, -1], [0, -1, 3]])
    est, var = O.solve_entry(m, [1., 2., 1.], [1], nwalks=200000, seed=11)
    assert abs(est[0] - 0.75609756) < 5 * np.sqrt(var[0] / 200000)
