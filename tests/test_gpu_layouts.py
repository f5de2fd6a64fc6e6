"""The device layouts of the hot kernels — the CSR slices (warp-stream kernel), the SELL-32 copy (zero shared memory,
lane-owns-row) and the column-slab split (one warp-stream pass per slab of the gathered vector, row sums carried over) —
must give the SAME BITS as the scalar oracle: all accumulate every row left to right like CSRStorage::multiply_vector
(src/matrix/sparse.rs:193-203), with FMA contraction off on both sides.
$SUBLINEAR_B200_SELL = 1 / 0 forces / forbids the SELL copy at ingest, $SUBLINEAR_B200_SLABS = 2..4 / 0 the slab split."""
import numpy as np
import pytest

import sublinear_b200 as sb

pytestmark = pytest.mark.gpu


LAYOUTS = ["csr", "sell", "slabs2", "slabs3", "slabs4"]


def set_layout(monkeypatch, layout, ncols=None):
    """force a device layout through the environment; returns the layout id storage_info must report"""
    if layout.startswith("slabs"):
        k = int(layout[5:])
        monkeypatch.setenv("SUBLINEAR_B200_SLABS", str(k))
        monkeypatch.setenv("SUBLINEAR_B200_SELL", "0")
        return sb.LAYOUT_CSR_SLABS if ncols is None or ncols >= k else sb.LAYOUT_CSR
    monkeypatch.setenv("SUBLINEAR_B200_SLABS", "0")
    monkeypatch.setenv("SUBLINEAR_B200_SELL", "1" if layout == "sell" else "0")
    return sb.LAYOUT_SELL32 if layout == "sell" else sb.LAYOUT_CSR


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


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("n,k", [(1, 1), (7, 3), (31, 5), (32, 5), (33, 5), (255, 9), (257, 9), (4096, 40), (20000, 11),
                                 (3000, 400), (100_003, 10)])
def test_spmv_bit_exact_in_all_layouts(oracle, monkeypatch, n, k, layout):
    O = oracle
    want = set_layout(monkeypatch, layout, n)
    A = ragged(O, n, k, n + k)
    m = to_gpu(A)
    info = m.storage_info()
    assert info["layout"] == want and (info["slots"] >= A.nnz if layout == "sell" else info["slots"] == A.nnz)
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


@pytest.mark.parametrize("layout", LAYOUTS)
def test_rectangular_and_nonfinite_inputs(oracle, monkeypatch, layout):
    """ncols != nrows (Matrix::multiply_vector is not square-only) and a NaN / inf in x must only reach the rows that
    reference it: padding slots of the SELL copy are never gathered."""
    O = oracle
    set_layout(monkeypatch, layout)
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


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("mode", [sb.MODE_CORRECT, sb.MODE_REF_COMPAT])
def test_solve_identical_to_oracle_in_all_layouts(oracle, monkeypatch, layout, mode):
    O = oracle
    want = set_layout(monkeypatch, layout)
    for n, k in [(33, 4), (5000, 12), (60_000, 9)]:
        A = ragged(O, n, k, 3 * n + k, dd=True)
        b = np.random.default_rng(n).uniform(-5, 5, n)
        m = to_gpu(A)
        assert m.storage_info()["layout"] == want
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


@pytest.mark.parametrize("layout", ["slabs3", "sell"])
def test_initial_guess_state_and_cg_in_slab_layout(oracle, monkeypatch, layout):
    """the entry points that need scratch for the partial row sums: A x0 at setup, residual kernels, stepping, CG"""
    O = oracle
    set_layout(monkeypatch, layout)
    n = 20_000
    A = ragged(O, n, 9, 77, dd=True)
    b = np.random.default_rng(3).uniform(-5, 5, n)
    x0 = np.random.default_rng(4).standard_normal(n)
    m = to_gpu(A)
    for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
        r = sb.NeumannSolver.default().solve(m, b, sb.SolverOptions(mode=mode, initial_guess=x0, collect_stats=True))
        o = O.neumann_solve(A, b, mode=mode, initial_guess=x0)
        assert (r.iterations, r.terms_computed, r.matvec_count) == (o.iterations, o.terms_computed, o.matvec_count)
        assert np.array_equal(r.solution, o.solution)
    st = sb.NeumannSolver.default().initialize(m, b)
    ost = O.NeumannState(A, b)
    for _ in range(5):
        assert st.step() == ost.step()
        assert np.array_equal(st.extract_solution(), ost.extract_solution())
        np.testing.assert_allclose(st.residual_norm(), ost.info()["residual_norm"], rtol=1e-9)
    S = O.Csr.from_triplets(*_sym(n, 3), n, n)
    ms = to_gpu(S)
    rc = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(200, 1e-6)).solve(ms, b)
    oc = O.cg_solve(S, b, max_iterations=200, tolerance=1e-6)
    assert rc.converged and rc.iterations == oc.iterations
    np.testing.assert_allclose(rc.solution, oc.solution, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("layout", ["csr", "slabs2", "slabs3"])
def test_hub_rows_in_every_layout(oracle, monkeypatch, layout):
    """rows above 1024 entries (hub rows of power-law graphs) are summed by the grid-wide pre-pass (lane-strided order:
    tolerance-level parity); the other rows of the same 32-row blocks must stay bit-exact, in the slab layout too (hub
    rows are only marked there: bit 15 of their row offset)"""
    O = oracle
    set_layout(monkeypatch, layout)
    rng = np.random.default_rng(17)
    n = 40_000
    hubs = np.array([3, 64, 65, 9000, 39_999])
    rows = np.concatenate([np.repeat(hubs, [5000, 1500, 30_000, 1025, 2000]), rng.integers(0, n, 300_000)])
    cols = rng.integers(0, n, len(rows))
    vals = rng.standard_normal(len(rows))
    A = O.Csr.from_triplets(rows, cols, vals, n, n)
    m = to_gpu(A)
    x = rng.standard_normal(n)
    y, ref = m.multiply_vector(x), A.multiply_vector(x)
    lens = np.diff(A.row_ptr.astype(np.int64))
    short = lens <= 1024
    assert (~short).sum() == len(hubs)
    exact = short.copy()
    if layout == "csr":       # the warp-stream kernel sums every row of a hub's 32-row block lane-strided
        for h in np.nonzero(~short)[0]:
            exact[(h // 32) * 32:(h // 32) * 32 + 32] = False
    assert np.array_equal(y[exact], ref[exact])
    np.testing.assert_allclose(y[~exact], ref[~exact], rtol=1e-11, atol=1e-11)
    y0 = rng.standard_normal(n)
    ya = m.multiply_vector_add(x, y0)
    np.testing.assert_allclose(ya, y0 + ref, rtol=1e-11, atol=1e-10)


def _sym(n, k):
    rng = np.random.default_rng(n + k)
    r = np.repeat(np.arange(n), k)
    c = rng.integers(0, n, n * k)
    v = rng.uniform(-1, 1, n * k)
    keep = r != c
    r, c, v = r[keep], c[keep], v[keep]
    rows, cols, vals = np.concatenate([r, c]), np.concatenate([c, r]), np.concatenate([v, v])
    s = np.bincount(rows, weights=np.abs(vals), minlength=n)
    return (np.concatenate([rows, np.arange(n)]), np.concatenate([cols, np.arange(n)]), np.concatenate([vals, 1.5 * s + 1.0]))


def test_unsorted_rows_drop_the_slab_split(oracle, monkeypatch):
    """from_csr accepts rows that are not sorted by column; splitting them by column would reorder the sums, so the
    split is dropped and the result still equals the oracle's (which adds in the given order)"""
    O = oracle
    monkeypatch.setenv("SUBLINEAR_B200_SLABS", "3")
    monkeypatch.setenv("SUBLINEAR_B200_SELL", "0")
    rng = np.random.default_rng(8)
    n, k = 5000, 7
    cols = rng.integers(0, n, (n, k)).astype(np.uint32)              # unsorted within the row
    vals = rng.standard_normal((n, k))
    rp = np.arange(n + 1, dtype=np.uint32) * k
    A = O.Csr(n, n, vals.ravel(), cols.ravel(), rp)
    m = sb.SparseMatrix.from_csr(rp, cols.ravel(), vals.ravel(), n, n)
    assert m.storage_info()["layout"] == sb.LAYOUT_CSR
    x = rng.standard_normal(n)
    assert np.array_equal(m.multiply_vector(x), A.multiply_vector(x))
    cs = np.sort(cols, axis=1)                                        # sorted: the split is used
    ms = sb.SparseMatrix.from_csr(rp, cs.ravel(), vals.ravel(), n, n)
    assert ms.storage_info()["layout"] == sb.LAYOUT_CSR_SLABS
    assert np.array_equal(ms.multiply_vector(x), O.Csr(n, n, vals.ravel(), cs.ravel(), rp).multiply_vector(x))


def test_layout_selection_rule_and_scale(oracle, monkeypatch):
    O = oracle
    monkeypatch.delenv("SUBLINEAR_B200_SELL", raising=False)
    monkeypatch.delenv("SUBLINEAR_B200_SLABS", raising=False)
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
    monkeypatch.setenv("SUBLINEAR_B200_SLABS", "3")                   # scale reaches the slab copies too
    m3 = sb.SparseMatrix.from_csr(rp, ci, v, 200_000, 200_000)
    assert m3.storage_info()["layout"] == sb.LAYOUT_CSR_SLABS
    m3.scale(0.5)
    assert np.array_equal(m3.multiply_vector(x), 0.5 * y)
