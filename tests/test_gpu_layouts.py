"""The two device layouts of the hot kernels — the CSR slices (warp-stream kernel) and the SELL-32 copy (zero
shared memory, lane-owns-row) — must give the SAME BITS as the scalar oracle: both accumulate every row left to right
like CSRStorage::multiply_vector (src/matrix/sparse.rs:193-203), with FMA contraction off on both sides.
$SUBLINEAR_B200_SELL = 1 / 0 forces / forbids the SELL copy at ingest."""
import numpy as np
import pytest

import sublinear_b200 as sb

pytestmark = pytest.mark.gpu


def to_gpu(A):
    return sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, A.nrows, A.ncols)


def ragged(O, n, k, seed, dd=False):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = rng.integers(0, n, n * k)
    vals = rng.standard_normal(n * k)
    drop = rng.random(n * k) < 0.3                        # ragged: varying row lengths, some empty rows
    rows, cols, vals = rows[~drop], cols[~drop], vals[~drop]
    if dd:                                                # strictly row-dominant: |a_ii| = 2 sum|off| + 1
        off = rows != cols
        rows, cols, vals = rows[off], cols[off], vals[off]
        s = np.bincount(rows, weights=np.abs(vals), minlength=n)
        rows = np.concatenate([rows, np.arange(n)])
        cols = np.concatenate([cols, np.arange(n)])
        vals = np.concatenate([vals, 2.0 * s + 1.0])
    return O.Csr.from_triplets(rows, cols, vals, n, n)


@pytest.mark.parametrize("layout", ["0", "1"])
@pytest.mark.parametrize("n,k", [(1, 1), (7, 3), (31, 5), (32, 5), (33, 5), (255, 9), (257, 9), (4096, 40), (20000, 11),
                                 (3000, 400), (100_003, 10)])
def test_spmv_bit_exact_in_both_layouts(oracle, monkeypatch, n, k, layout):
    O = oracle
    monkeypatch.setenv("SUBLINEAR_B200_SELL", layout)
    A = ragged(O, n, k, n + k)
    m = to_gpu(A)
    info = m.storage_info()
    assert info["layout"] == int(layout) and (info["slots"] >= A.nnz if layout == "1" else info["slots"] == A.nnz)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    y = m.multiply_vector(x)
    assert np.array_equal(y, A.multiply_vector(x, O.SPMV_SCALAR))
    y0 = rng.standard_normal(n)
    acc = y0.copy()                                       # multiply_vector_add: y[row] += products, left to right
    for i in range(min(n, 500)):
        a = y0[i]
        for q in range(A.row_ptr[i], A.row_ptr[i + 1]):
            a += A.values[q] * x[A.col_indices[q]]
        acc[i] = a
    assert np.array_equal(m.multiply_vector_add(x, y0)[:500], acc[:500])
    rp, ci, v = m.to_csr()                                # export is the CSRStorage slices in either layout
    assert np.array_equal(rp, A.row_ptr) and np.array_equal(ci, A.col_indices) and np.array_equal(v, A.values)


@pytest.mark.parametrize("layout", ["0", "1"])
def test_rectangular_and_nonfinite_inputs(oracle, monkeypatch, layout):
    """ncols != nrows (Matrix::multiply_vector is not square-only) and a NaN / inf in x must only reach the rows that
    reference it: padding slots of the SELL copy are never gathered."""
    O = oracle
    monkeypatch.setenv("SUBLINEAR_B200_SELL", layout)
    rng = np.random.default_rng(2)
    nr, nc, nt = 70, 45, 400
    A = O.Csr.from_triplets(rng.integers(0, nr, nt), rng.integers(1, nc, nt), rng.standard_normal(nt), nr, nc)
    m = to_gpu(A)
    x = rng.standard_normal(nc)
    x[0] = np.nan                                          # column 0 is referenced by no entry (and is the padding column)
    y = m.multiply_vector(x)
    assert np.isfinite(y).all() and np.array_equal(y, A.multiply_vector(x))
    x[7] = np.inf
    y = m.multiply_vector(x)
    ref = A.multiply_vector(x)
    assert np.array_equal(np.isfinite(y), np.isfinite(ref)) and np.array_equal(y[np.isfinite(ref)], ref[np.isfinite(ref)])


@pytest.mark.parametrize("layout", ["0", "1"])
@pytest.mark.parametrize("mode", [sb.MODE_CORRECT, sb.MODE_REF_COMPAT])
def test_solve_identical_to_oracle_in_both_layouts(oracle, monkeypatch, layout, mode):
    O = oracle
    monkeypatch.setenv("SUBLINEAR_B200_SELL", layout)
    for n, k in [(33, 4), (5000, 12), (60_000, 9)]:
        A = ragged(O, n, k, 3 * n + k, dd=True)
        b = np.random.default_rng(n).uniform(-5, 5, n)
        m = to_gpu(A)
        assert m.storage_info()["layout"] == int(layout)
        for resid in (sb.RESIDUAL_EVERY_5, sb.RESIDUAL_IDENTITY):
            if resid == sb.RESIDUAL_IDENTITY and mode != sb.MODE_CORRECT:
                continue
            r = sb.NeumannSolver.default().solve(m, b, sb.SolverOptions(mode=mode, residual_check=resid, collect_stats=True))
            if resid == sb.RESIDUAL_EVERY_5:
                o = O.neumann_solve(A, b, mode=mode)
                assert (r.iterations, r.terms_computed, r.matvec_count, r.converged) == \
                       (o.iterations, o.terms_computed, o.matvec_count, o.converged)
                assert np.array_equal(r.solution, o.solution)
                np.testing.assert_allclose(r.residual_norm, o.residual_norm, rtol=1e-9, atol=1e-18)
            else:
                assert r.converged
                assert np.linalg.norm(A.multiply_vector(r.solution) - b) <= 1e-6


def test_layout_selection_rule_and_scale(oracle, monkeypatch):
    O = oracle
    monkeypatch.delenv("SUBLINEAR_B200_SELL", raising=False)
    rp, ci, v, b = sb.gen_bench_csr(200_000, 5e-5)         # every row ~10 entries: SELL costs no padding
    m = sb.SparseMatrix.from_csr(rp, ci, v, 200_000, 200_000)
    info = m.storage_info()
    assert info["layout"] == sb.LAYOUT_SELL32 and info["slots"] <= 1.01 * len(v) + 2048
    rng = np.random.default_rng(9)                         # power-law rows: padding would multiply the stream
    n = 50_000
    dst = np.minimum((rng.pareto(1.1, 400_000) * 20).astype(np.int64), n - 1)
    src = rng.integers(0, n, 400_000)
    S, _ = sb.SparseMatrix.pagerank_system(src, dst, n, 0.85)
    assert S.storage_info()["layout"] == sb.LAYOUT_CSR
    # SparseMatrix::scale (src/matrix/mod.rs:345-357) must reach both copies
    x = rng.standard_normal(200_000)
    y = m.multiply_vector(x)
    m.scale(0.5)
    assert np.array_equal(m.multiply_vector(x), 0.5 * y)
    r = sb.NeumannSolver.default().solve(m, b)
    A, _ = O.gen_bench_csr(200_000, 5e-5)
    A.values *= 0.5
    assert np.array_equal(r.solution, O.neumann_solve(A, b).solution)
