"""The SolverAlgorithm state interface (initialize / step / is_converged / extract_solution / update_rhs / reset,
src/solver/mod.rs:223-252, src/solver/neumann.rs:350-462) and the streaming solve (SolverOptions::streaming +
PartialSolution, src/solver/mod.rs:100-116, 198-217) — SURVEY.md §8f.4.

CPU part: the oracle's restatement against the reference's own unit test of the state and the documented semantics.
GPU part: the device path step by step against the oracle (solution bit-exact: same left-to-right row sums)."""
import numpy as np
import pytest

import sublinear_b200 as sb


def dd_system(O, n, k, seed):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = rng.integers(0, n, n * k)
    vals = rng.uniform(-1, 1, n * k)
    off = rows != cols
    rows, cols, vals = rows[off], cols[off], vals[off]
    s = np.bincount(rows, weights=np.abs(vals), minlength=n)
    rows = np.concatenate([rows, np.arange(n)])
    cols = np.concatenate([cols, np.arange(n)])
    vals = np.concatenate([vals, 2.0 * s + 1.0])
    return O.Csr.from_triplets(rows, cols, vals, n, n), rng.uniform(-5, 5, n)


# ---- oracle (CPU) ------------------------------------------------------------------------------------------------

def test_oracle_state_initialisation_matches_rust_unit_test(oracle):
    # neumann.rs:633-648: diag(2,3), b = [4,6] -> dimension 2, diagonal_inv [0.5, 1/3], rhs [2,2], terms_computed 0
    O = oracle
    A = O.Csr.from_dense(np.diag([2.0, 3.0]))
    st = O.NeumannState(A, [4.0, 6.0], mode=O.MODE_REF_COMPAT)
    i = st.info()
    assert i["terms_computed"] == 0 and i["matvec_count"] == 0 and np.isinf(i["residual_norm"]) and not st.is_converged()
    assert st.extract_solution().tolist() == [2.0, 2.0]                 # ref_compat: solution = rhs.clone() (:207)
    assert O.NeumannState(A, [4.0, 6.0]).extract_solution().tolist() == [0.0, 0.0]


def test_oracle_stepping_equals_batch_solve_terms(oracle):
    """k steps hold the k-term partial sum: the same terms the batch solve accumulates (its residual cadence differs)."""
    O = oracle
    A, b = dd_system(O, 400, 6, 1)
    for mode in (O.MODE_CORRECT, O.MODE_REF_COMPAT):
        st = O.NeumannState(A, b, mode=mode)
        steps = 0
        while True:
            res = st.step()
            steps += 1
            if res == 1:
                break
        ref = O.neumann_solve(A, b, mode=mode, tolerance=0.0)            # tolerance 0: stops on the series criterion
        assert steps == ref.terms_computed and st.info()["series_converged"]
        assert np.array_equal(st.extract_solution(), ref.solution)
        assert st.info()["matvec_count"] == 2 * steps - 1                # (steps-1) term SpMVs + steps residual SpMVs
        assert st.is_converged()


def test_oracle_update_rhs_semantics(oracle):
    O = oracle
    A, b = dd_system(O, 300, 5, 2)
    delta = [(3, 1.0), (17, -2.5), (3, 0.25)]                            # a repeated index accumulates in order
    b2 = b.copy()
    for i, d in delta:
        b2[i] += d
    # correct mode: converge, update, converge again -> the solution of A x = b + delta_b
    st = O.NeumannState(A, b, series_tolerance=1e-13, max_terms=200)
    while st.step() == 0:
        pass
    st.update_rhs(delta)
    assert st.info()["terms_computed"] == 0 and not st.info()["series_converged"]
    while st.step() == 0:
        pass
    x = st.extract_solution()
    assert np.linalg.norm(A.multiply_vector(x) - b2) <= 1e-10 * np.linalg.norm(b2)
    fresh = O.neumann_solve(A, b2, series_tolerance=1e-13, max_terms=200, tolerance=0.0).solution
    np.testing.assert_allclose(x, fresh, rtol=1e-10, atol=1e-12)
    # ref_compat: the literal code adds the scaled delta to rhs and solution and restarts from the whole rhs (:446-459)
    st = O.NeumannState(A, b, mode=O.MODE_REF_COMPAT)
    x0 = st.extract_solution()
    st.update_rhs(delta)
    dinv = 1.0 / np.array([A.get(i, i) for i in range(A.nrows)])
    want = x0.copy()
    for i, d in delta:
        want[i] += d * dinv[i]
    assert np.array_equal(st.extract_solution(), want)
    with pytest.raises(O.OracleError) as e:                              # IndexOutOfBounds, state untouched (:439-445)
        st.update_rhs([(0, 1.0), (300, 1.0)])
    assert e.value.code == O.ERR_INDEX_OUT_OF_BOUNDS and np.array_equal(st.extract_solution(), want)
    st.reset()                                                           # SolverState::reset (:367-378)
    i = st.info()
    assert (st.extract_solution() == 0).all() and i["terms_computed"] == 0 and i["matvec_count"] == 0
    assert np.isinf(i["residual_norm"])


# ---- device path (GPU) ---------------------------------------------------------------------------------------------

def to_gpu(A):
    return sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, A.nrows, A.ncols)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [sb.MODE_CORRECT, sb.MODE_REF_COMPAT])
def test_gpu_state_steps_match_oracle(oracle, mode):
    O = oracle
    A, b = dd_system(O, 3000, 7, 5)
    m = to_gpu(A)
    solver = sb.NeumannSolver.new(30, 1e-9)
    st = solver.initialize(m, b, sb.SolverOptions(mode=mode))
    ost = O.NeumannState(A, b, mode=mode, max_terms=30, series_tolerance=1e-9)
    assert np.array_equal(st.extract_solution(), ost.extract_solution())
    assert st.info()["terms_computed"] == 0 and np.isinf(st.residual_norm()) and not st.is_converged()
    for k in range(1, 40):
        r, ro = st.step(), ost.step()
        i, io = st.info(), ost.info()
        assert r == ro and st.is_converged() == ost.is_converged()
        assert (i["terms_computed"], i["matvec_count"], i["series_converged"]) == \
               (io["terms_computed"], io["matvec_count"], io["series_converged"])
        assert np.array_equal(st.extract_solution(), ost.extract_solution())
        np.testing.assert_allclose(i["residual_norm"], io["residual_norm"], rtol=1e-9, atol=1e-18)
        np.testing.assert_allclose(i["last_term_norm"], io["last_term_norm"], rtol=1e-11)
        if io["error_upper_bound"] is None:
            assert i["error_upper_bound"] is None
        else:
            np.testing.assert_allclose(i["error_upper_bound"], io["error_upper_bound"], rtol=1e-6)
        if r == sb.STEP_CONVERGED:
            break
    assert r == sb.STEP_CONVERGED and 5 < k < 30


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [sb.MODE_CORRECT, sb.MODE_REF_COMPAT])
def test_gpu_update_rhs_and_reset_match_oracle(oracle, mode):
    O = oracle
    A, b = dd_system(O, 2000, 6, 8)
    m = to_gpu(A)
    st = sb.NeumannSolver.new(200, 1e-13).initialize(m, b, sb.SolverOptions(mode=mode))
    ost = O.NeumannState(A, b, mode=mode, max_terms=200, series_tolerance=1e-13)
    while ost.step() == 0:
        assert st.step() == 0
    assert st.step() == 1
    delta = [(3, 1.0), (1999, -2.5), (3, 0.25), (500, 4.0)]
    st.update_rhs(delta)
    ost.update_rhs(delta)
    assert np.array_equal(st.extract_solution(), ost.extract_solution())
    assert st.info()["terms_computed"] == 0 and not st.info()["series_converged"]
    while True:
        r, ro = st.step(), ost.step()
        assert r == ro
        assert np.array_equal(st.extract_solution(), ost.extract_solution())
        if r == 1:
            break
    if mode == sb.MODE_CORRECT:                                          # the incremental solve converges to A^-1 (b + delta_b)
        b2 = b.copy()
        for i, d in delta:
            b2[i] += d
        assert np.linalg.norm(A.multiply_vector(st.extract_solution()) - b2) <= 1e-10 * np.linalg.norm(b2)
        np.testing.assert_allclose(st.residual_norm(), ost.info()["residual_norm"], rtol=1e-6, atol=1e-13)
    with pytest.raises(sb.SolverError) as e:
        st.update_rhs([(0, 1.0), (2000, 1.0)])
    assert e.value.variant == "IndexOutOfBounds" and np.array_equal(st.extract_solution(), ost.extract_solution())
    st.reset()
    ost.reset()
    assert (st.extract_solution() == 0).all() and st.info()["matvec_count"] == 0 and np.isinf(st.residual_norm())
    for _ in range(3):
        assert st.step() == ost.step()
    assert np.array_equal(st.extract_solution(), ost.extract_solution())


@pytest.mark.gpu
def test_gpu_initialize_errors_and_initial_guess(oracle):
    O = oracle
    with pytest.raises(sb.SolverError) as e:                             # neumann.rs:609-631
        sb.NeumannSolver.default().initialize(sb.SparseMatrix.from_dense([[1., 3.], [2., 1.]]), [4., 3.])
    assert e.value.variant == "MatrixNotDiagonallyDominant"
    A, b = dd_system(O, 500, 5, 3)
    m = to_gpu(A)
    with pytest.raises(sb.SolverError) as e:
        sb.NeumannSolver.default().initialize(m, b[:-1])
    assert e.value.variant == "DimensionMismatch"
    x0 = np.random.default_rng(0).standard_normal(500)
    for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
        st = sb.NeumannSolver.default().initialize(m, b, sb.SolverOptions(mode=mode, initial_guess=x0))
        ost = O.NeumannState(A, b, mode=mode, initial_guess=x0)
        assert np.array_equal(st.extract_solution(), x0)
        for _ in range(4):
            assert st.step() == ost.step()
            assert np.array_equal(st.extract_solution(), ost.extract_solution())
        assert st.info()["matvec_count"] == ost.info()["matvec_count"]


@pytest.mark.gpu
def test_gpu_streaming_partials(oracle):
    O = oracle
    A, b = dd_system(O, 4000, 6, 12)
    m = to_gpu(A)
    solver = sb.NeumannSolver.new(60, 1e-11)
    opts = sb.SolverOptions.streaming(4)                                  # src/solver/mod.rs:100-116
    assert opts.streaming_interval == 4 and opts.tolerance == 1e-4
    opts.tolerance = 0.0                                                  # run until the series criterion
    seen = []
    r = solver.solve_streaming(m, b, opts, lambda p: seen.append(p) and None)
    plain = solver.solve(m, b, sb.SolverOptions(tolerance=0.0))
    assert np.array_equal(r.solution, plain.solution) and r.iterations == plain.iterations
    assert len(seen) == -(-(r.iterations - 1) // 4) and [p["iteration"] for p in seen[:-1]] == [5 + 4 * i for i in range(len(seen) - 1)]
    for p in seen[:-1]:                                                   # a partial after k iterations = the k-term sum
        xk, _, _, _ = O.push_iterations(A, b, p["iteration"] - 1)         # x = c + t_1 + .. : the k-term partial sum
        assert np.array_equal(p["solution"], xk)
        assert p["estimated_remaining"] is not None and not p["converged"]
    assert seen[-1]["converged"] and np.array_equal(seen[-1]["solution"], r.solution)
    ts = [p["timestamp_ms"] for p in seen]
    assert ts == sorted(ts)
    # the callback can stop the solve: the iterate at that point is returned, not converged, no error
    stopped = solver.solve_streaming(m, b, opts, lambda p: True)
    assert stopped.iterations == 5 and not stopped.converged and np.array_equal(stopped.solution, seen[0]["solution"])
    # streaming_interval = 0 -> plain solve, callback never called
    calls = []
    r0 = solver.solve_streaming(m, b, sb.SolverOptions(tolerance=0.0), lambda p: calls.append(p))
    assert not calls and np.array_equal(r0.solution, plain.solution)


@pytest.mark.gpu
def test_gpu_state_outlives_matrix_handle(oracle):
    """The state shares ownership of the matrix handle: sb200_matrix_free before sb200_state_free is legal (a garbage
    collector finalises the two wrappers in arbitrary order), and the state stays usable."""
    O = oracle
    A, b = dd_system(O, 1000, 5, 21)
    m = to_gpu(A)
    st = sb.NeumannSolver.default().initialize(m, b)
    st.step()
    sb.lib().sb200_matrix_free(m._h)                     # the caller lets go of its handle first
    m._h = None
    ost = O.NeumannState(A, b)
    ost.step()
    for _ in range(3):
        assert st.step() == ost.step()
    assert np.array_equal(st.extract_solution(), ost.extract_solution())
    del st                                               # the last owner frees the matrix


@pytest.mark.gpu
def test_streaming_matrix_row_chunks(oracle):
    """StreamingMatrix (src/matrix/optimized.rs:451-561): chunk size from the memory limit, chunks visited in row order,
    every slice bit-identical to the oracle's SpMV over the whole matrix; memory_usage = sum(nnz * 12 + rows * 4)"""
    O = oracle
    rng = np.random.default_rng(21)
    nr, nc, nt = 50_000, 40_000, 600_000
    rows, cols = rng.integers(0, nr, nt), rng.integers(0, nc, nt)
    vals = rng.standard_normal(nt)
    vals[::97] = 0.0                                                     # exact zeros are dropped
    A = O.Csr.from_triplets(rows, cols, vals, nr, nc)
    sm = sb.StreamingMatrix.from_triplets(rows, cols, vals, nr, nc, memory_limit_mb=1)
    info = sm.info()
    avg = nt // nr
    want_chunk = min(nr, max(1, (1 << 20) // ((avg * 12 + 4) * 2)))      # optimized.rs:470-482
    assert info["chunk_size"] == want_chunk and info["num_chunks"] == -(-nr // want_chunk) and info["num_chunks"] > 5
    assert info["memory_usage"] == A.nnz * 12 + nr * 4
    x = rng.standard_normal(nc)
    seen = []
    y = np.full(nr, np.nan)

    def cb(start, part):
        seen.append((start, len(part)))
        y[start:start + len(part)] = part

    sm.multiply_vector_streaming(x, cb)
    assert [s for s, _ in seen] == [k * want_chunk for k in range(info["num_chunks"])]
    assert sum(n for _, n in seen) == nr
    assert np.array_equal(y, A.multiply_vector(x))
    assert np.array_equal(sm.multiply_vector(x), y)
    stops = []
    sm.multiply_vector_streaming(x, lambda s, p: stops.append(s) or len(stops) == 3)      # the callback may stop the walk
    assert len(stops) == 3
    one = sb.StreamingMatrix.from_triplets(rows, cols, vals, nr, nc, memory_limit_mb=4096)  # everything fits: one chunk
    assert one.info()["num_chunks"] == 1 and np.array_equal(one.multiply_vector(x), y)
    with pytest.raises(sb.SolverError) as ei:
        sm.multiply_vector(np.ones(nc + 1))
    assert ei.value.variant == "DimensionMismatch"
    with pytest.raises(sb.SolverError) as ei:
        sb.StreamingMatrix.from_triplets([nr], [0], [1.0], nr, nc, 1)
    assert ei.value.variant == "IndexOutOfBounds"
