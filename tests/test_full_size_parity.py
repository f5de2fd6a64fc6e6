"""Parity on exactly what bench.py measures: the headline workload gen_bench(10 000 000, 1e-6) (BASELINE.json configs[2]
shape, the `metric` size) and config C2 gen_bench(1 000 000, 1e-5) in their DEFAULT device layouts, against the CPU
oracle's full solve (oracle/sublinear_oracle.c: orc_neumann_solve with the row-chunk parallel SpMV of
src/simd_ops.rs:202-239 — same left-to-right order inside a row as the scalar loop of src/matrix/sparse.rs:193-203, so
the comparison is bit for bit), plus the row blocks the multi-GPU path builds at N = 2, 4, 8.

Tolerances: iterations / terms / matvec_count / converged identical; solution np.array_equal (bit-exact f64); residual
norm rtol 1e-9 (the norm is a fixed-shape tree on the GPU, a sequential sum in the oracle)."""
import numpy as np
import pytest

import sublinear_b200 as sb

pytestmark = pytest.mark.gpu


def _check_solves(O, A, b, m):
    x = np.random.default_rng(1).standard_normal(A.ncols)
    assert np.array_equal(m.multiply_vector(x), A.multiply_vector(x, O.SPMV_PARALLEL)), "SpMV differs from the oracle"
    for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
        r = sb.NeumannSolver.default().solve(m, b, sb.SolverOptions(mode=mode, collect_stats=True))
        o = O.neumann_solve(A, b, mode=mode, spmv_variant=O.SPMV_PARALLEL)
        assert (r.iterations, r.terms_computed, r.matvec_count, r.converged) == \
               (o.iterations, o.terms_computed, o.matvec_count, o.converged), (mode, r, o)
        assert np.array_equal(r.solution, o.solution), f"mode {mode}: solution differs from the oracle"
        np.testing.assert_allclose(r.residual_norm, o.residual_norm, rtol=1e-9, atol=1e-18)
        if mode == sb.MODE_CORRECT:    # the contract of BASELINE.json: ||Ax-b||/||b|| from an independent (oracle) SpMV
            res = np.linalg.norm(A.multiply_vector(r.solution, O.SPMV_PARALLEL) - b)
            np.testing.assert_allclose(r.residual_norm, res, rtol=1e-6, atol=1e-12)
            assert res / np.linalg.norm(b) < 1e-6


def test_headline_n10M_default_layout_vs_oracle(oracle, monkeypatch):
    """bench.py's workload c3_n10M_nnz100M: the default layout must be the column-slab one, and the solve must equal the
    oracle's bit for bit in both modes."""
    O = oracle
    for k in ("SUBLINEAR_B200_SLABS", "SUBLINEAR_B200_SELL", "SUBLINEAR_B200_TILE_CFG"):
        monkeypatch.delenv(k, raising=False)
    n = 10_000_000
    A, b = O.gen_bench_csr(n, 1e-6)
    assert 99_000_000 < A.nnz <= 100_000_000
    m = sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, n, n)
    assert m.storage_info()["layout"] == sb.LAYOUT_CSR_SLABS
    _check_solves(O, A, b, m)


def test_c2_n1M_default_layout_vs_oracle(oracle, monkeypatch):
    """config C2 (n = 1 M, nnz = 10 M) at full size: 8 MB vector, SELL-32 layout by default."""
    O = oracle
    for k in ("SUBLINEAR_B200_SLABS", "SUBLINEAR_B200_SELL", "SUBLINEAR_B200_TILE_CFG"):
        monkeypatch.delenv(k, raising=False)
    n = 1_000_000
    A, b = O.gen_bench_csr(n, 1e-5)
    m = sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, n, n)
    assert m.storage_info()["layout"] == sb.LAYOUT_SELL32
    _check_solves(O, A, b, m)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_blocks_of_the_headline_system_vs_oracle(oracle, monkeypatch, world):
    """the row block rank `world - 1` holds at N = world (rows x 10 M columns, gathers from the full-length vector), in
    whatever layout the default rule picks for it: SpMV and accumulate bit-exact against the oracle's row block."""
    O = oracle
    for k in ("SUBLINEAR_B200_SLABS", "SUBLINEAR_B200_SELL", "SUBLINEAR_B200_TILE_CFG"):
        monkeypatch.delenv(k, raising=False)
    n = 10_000_000
    r0, r1 = sb.partition_rows(n, world, world - 1)
    A, _ = O.gen_bench_csr(n, 1e-6, r0, r1)
    m = sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, r1 - r0, n)
    rng = np.random.default_rng(world)
    x = rng.standard_normal(n)
    y = m.multiply_vector(x)
    ref = A.multiply_vector(x, O.SPMV_PARALLEL, ylen=r1 - r0)
    assert np.array_equal(y, ref), (world, m.storage_info())
    y0 = rng.standard_normal(r1 - r0)
    ya = m.multiply_vector_add(x, y0)
    k = 4096                                                   # y0 + products, left to right (first rows by hand)
    acc = y0[:k].copy()
    for i in range(k):
        a = acc[i]
        for q in range(A.row_ptr[i], A.row_ptr[i + 1]):
            a += A.values[q] * x[A.col_indices[q]]
        acc[i] = a
    assert np.array_equal(ya[:k], acc)
