"""Forward / backward push: ForwardPushSolver / BackwardPushSolver
(src/solver/forward_push.rs, src/solver/backward_push.rs, graphs of tests/rust/push_tests.rs:15-77).

CPU part: the oracle's sequential restatement against the properties the reference's own tests assert and against the
exact personalised PageRank  pi_s = alpha (I - (1-alpha) P^T)^-1 e_s  (P = row-normalised adjacency).
Pinning of the oracle: the reference's queue item derives PartialOrd on (priority, node_id) (src/graph/mod.rs:141-147)
— the order BinaryHeap would use had the type an Ord impl — so the pop sequence is a total order and independent of
the heap's internals. An independent pure-Python restatement (heapq on (-priority, -node)) must reproduce the C
oracle bit for bit on the graphs of tests/rust/push_tests.rs:15-59; the vectors are committed under tests/golden/
(make_golden_push.py).
GPU part: the sparse-frontier device version pushes every queued node above its threshold per round (same push rule,
same stopping condition), so it satisfies the same invariants and the same error bound; it is compared with the exact
PPR and with the oracle within that bound (a different push order gives different last digits, not a different limit).

Bound used: at exit every residual is < eps * max(deg, 1), and estimate = pi - sum_u r[u] pi_u, hence
0 <= pi[v] - estimate[v] <= sum_u r[u] <= eps * sum_u max(deg_u, 1)."""
import numpy as np
import pytest

import sublinear_b200 as sb


def simple_graph():
    """create_test_graph / create_simple_graph (forward_push.rs:333-341, push_tests.rs:15-22)"""
    return np.array([0, 2, 4, 6, 7]), np.array([1, 2, 0, 3, 0, 3, 1]), np.array([0.5, 0.5, 0.8, 0.2, 0.6, 0.4, 1.0]), 4


def random_graph(n, epn):
    """create_random_graph (push_tests.rs:25-47): LCG targets, weight 1/epn, self-targets skipped, rows normalised"""
    seed, rows, cols = 12345, [], []
    for i in range(n):
        for _ in range(epn):
            seed = (seed * 1103515245 + 12345) % 2 ** 64
            t = seed % n
            if t != i:
                rows.append(i)
                cols.append(t)
    rows, cols = np.array(rows), np.array(cols)
    w = np.full(len(rows), 1.0 / epn)
    s = np.bincount(rows, weights=w, minlength=n)
    w = w / s[rows]                                        # AdjacencyList::normalize
    rp = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))])
    return rp, cols, w, n


def path_graph(n):
    """create_path_graph (push_tests.rs:50-59): 0 -> 1 -> .. -> n-1, the last node has no out-edge"""
    return np.concatenate([np.arange(n), [n - 1]]), np.arange(1, n), np.ones(n - 1), n


def dense(rp, ci, w, n):
    A = np.zeros((n, n))
    for u in range(n):
        for k in range(rp[u], rp[u + 1]):
            A[u, ci[k]] += w[k]
    return A


def exact_ppr(A, s, alpha):
    """forward push target: nodes without out-weight keep their mass (self-loop), forward_push.rs:209-214"""
    n = len(A)
    deg = A.sum(1)
    P = np.where(deg[:, None] > 0, A / np.where(deg > 0, deg, 1.0)[:, None], np.eye(n))
    return alpha * np.linalg.solve(np.eye(n) - (1 - alpha) * P.T, np.eye(n)[s])


def backward_exact(A, t, alpha):
    """backward push target for a row-normalised graph (out-degree 1): column t of alpha (I - (1-alpha) P)^-1; a node
    without in-edges keeps the mass that reaches it and converts all of it (backward_push.rs:212-217: the self-loop),
    so its estimate is the arriving mass itself, 1/alpha times the column entry"""
    n = len(A)
    col = alpha * np.linalg.solve(np.eye(n) - (1 - alpha) * A, np.eye(n)[t])
    return np.where(A.sum(0) > 0, col, col / alpha)


def check_forward(r, A, s, alpha, eps):
    deg = A.sum(1)
    assert (r.estimate >= 0).all() and (r.residual >= 0).all()
    assert abs(r.estimate.sum() + r.residual.sum() - 1.0) < 1e-12                      # mass conservation
    assert (r.residual < eps * np.maximum(deg, 1.0)).all()                             # stopping condition
    pi = exact_ppr(A, s, alpha)
    gap = pi - r.estimate
    assert (gap >= -1e-12).all() and gap.max() <= eps * np.maximum(deg, 1.0).sum() + 1e-12
    np.testing.assert_allclose(r.residual_norm, np.sqrt((r.residual ** 2).sum()), rtol=1e-12, atol=1e-300)


# ---- oracle (CPU) ----------------------------------------------------------------------------------------------

def test_oracle_forward_push_reference_tests(oracle):
    O = oracle
    rp, ci, w, n = simple_graph()
    A = O.Csr(n, n, w, ci, rp)
    r = O.forward_push(A, 0)                                                           # forward_push.rs:343-356
    assert r.push_count > 0 and r.nodes_visited > 0 and r.estimate[0] > 0 and r.residual_norm >= 0
    check_forward(r, dense(rp, ci, w, n), 0, 0.15, 1e-6)
    assert abs(r.extrapolated_solution(0.15).sum() - 1.0) < 0.1                        # test_mass_conservation
    m = O.forward_push(A, [0, 2])                                                      # test_forward_push_multi_source
    assert m.push_count > 0 and m.estimate.sum() > 0
    assert abs(m.estimate.sum() + m.residual.sum() - 1.0) < 1e-12
    z = O.forward_push(A, 7)                                                           # out-of-range source (:73-82)
    assert z.push_count == 0 and z.nodes_visited == 0 and not z.estimate.any() and z.residual_norm == 0.0
    capped = O.forward_push(A, 0, max_pushes=5)
    assert capped.push_count == 5


@pytest.mark.parametrize("graph,seed_node", [(random_graph(200, 5), 3), (path_graph(12), 0), (path_graph(12), 11)])
def test_oracle_forward_push_vs_exact_ppr(oracle, graph, seed_node):
    O = oracle
    rp, ci, w, n = graph
    for eps in (1e-4, 1e-8):
        r = O.forward_push(O.Csr(n, n, w, ci, rp), seed_node, epsilon=eps, queue_threshold=eps / 10, adaptive_threshold=False)
        check_forward(r, dense(rp, ci, w, n), seed_node, 0.15, eps)


def test_oracle_backward_push(oracle):
    O = oracle
    rp, ci, w, n = random_graph(150, 4)
    A = dense(rp, ci, w, n)
    t = 7
    # epsilon below the default queue_threshold (1e-8): the queue admission test would decide, so it is lowered too
    r = O.backward_push(O.Csr(n, n, w, ci, rp), t, epsilon=1e-9, queue_threshold=1e-10, adaptive_threshold=False)
    assert r.push_count > 0 and (r.estimate >= 0).all() and (r.residual >= 0).all()
    rdeg = A.sum(0)
    assert (r.residual < 1e-9 * np.maximum(rdeg, 1.0)).all()
    np.testing.assert_allclose(r.estimate, backward_exact(A, t, 0.15), rtol=0, atol=1e-6)
    z = O.backward_push(O.Csr(n, n, w, ci, rp), n + 3)
    assert z.push_count == 0 and not z.estimate.any()


def py_forward_push(rp, ci, w, n, sources, alpha=0.15, eps=1e-6, max_pushes=1_000_000, thr=1e-8, adaptive=True,
                    target=None, precision=0.0):
    """ForwardPushSolver::solve_* (forward_push.rs:66-290) + WorkQueue (mod.rs:130-212), independent of the C oracle:
    heapq as a max-heap on (priority, node_id) — derive(PartialOrd) order"""
    import heapq
    deg = np.zeros(n)
    for u in range(n):
        for k in range(rp[u], rp[u + 1]):
            deg[u] += w[k]
    est, res = np.zeros(n), np.zeros(n)
    srcs = [s for s in np.atleast_1d(sources) if s < n]
    if not srcs:
        return est, res, 0
    mass = 1.0 if np.ndim(sources) == 0 or len(np.atleast_1d(sources)) == 1 else 1.0 / len(np.atleast_1d(sources))
    for s_ in srcs:
        res[s_] += mass
    heap, inq = [], set()

    def push_if(node, r, d):
        pr = r / d if d > 0 else r
        if pr >= thr_box[0] and node not in inq:
            heapq.heappush(heap, (-pr, -node))
            inq.add(node)

    thr_box = [thr]
    for s_ in srcs:
        push_if(s_, res[s_], max(deg[s_], 1.0))
    pushes = 0
    while heap and pushes < max_pushes:
        if target is not None and est[target] > precision and res[target] < precision * 0.1:
            break
        _, negnode = heapq.heappop(heap)
        node = -negnode
        inq.discard(node)
        if res[node] < eps * max(deg[node], 1.0):
            continue
        if res[node] > 0.0:
            est[node] += alpha * res[node]
            remaining = (1.0 - alpha) * res[node]
            res[node] = 0.0
            if deg[node] > 0.0:
                for k in range(rp[node], rp[node + 1]):
                    v = ci[k]
                    res[v] += remaining * w[k] / deg[node]
                    push_if(v, res[v], max(deg[v], 1.0))
            else:
                res[node] += remaining
                push_if(node, res[node], 1.0)
        pushes += 1
        if adaptive and pushes % 1000 == 0:
            if len(heap) > 10000:
                thr_box[0] *= 1.1
            elif len(heap) < 100 and thr_box[0] > 1e-12:
                thr_box[0] *= 0.9
    return est, res, pushes


@pytest.mark.parametrize("name,graph,src", [("simple4", simple_graph(), 0), ("lcg100x5", random_graph(100, 5), 3),
                                             ("lcg100x5_multi", random_graph(100, 5), [0, 2, 50]), ("path12", path_graph(12), 0)])
def test_oracle_push_pinned_to_independent_restatement(oracle, golden_dir, name, graph, src):
    """the C oracle's pop order = derive(PartialOrd) on (priority, node_id): bit-identical to the heapq restatement and to
    the committed vectors (tests/golden/push_<name>.npz, written by tests/golden/make_golden_push.py from the restatement)"""
    import os
    O = oracle
    rp, ci, w, n = graph
    r = O.forward_push(O.Csr(n, n, w, ci, rp), src)
    est, res, pushes = py_forward_push(rp, ci, w, n, src)
    assert pushes == r.push_count
    assert np.array_equal(est, r.estimate) and np.array_equal(res, r.residual)
    g = np.load(os.path.join(golden_dir, f"push_{name}.npz"))
    assert int(g["push_count"]) == r.push_count
    assert np.array_equal(g["estimate"], r.estimate) and np.array_equal(g["residual"], r.residual)
    if np.ndim(src) == 0:
        t = (src + 1) % n
        rt = O.forward_push_with_target(O.Csr(n, n, w, ci, rp), src, t, 1e-3)
        e2, r2, p2 = py_forward_push(rp, ci, w, n, src, target=t, precision=1e-3)
        assert p2 == rt.push_count and np.array_equal(e2, rt.estimate) and np.array_equal(r2, rt.residual)
        assert rt.push_count <= r.push_count


def test_oracle_with_source_and_combine(oracle):
    """solve_with_source stops early; combine_with_forward (backward_push.rs:312-330) against numpy; reference test
    properties of push_tests.rs (bidirectional estimate finite and >= 0)"""
    O = oracle
    rp, ci, w, n = random_graph(120, 4)
    A = O.Csr(n, n, w, ci, rp)
    full = O.backward_push(A, 9)
    early = O.backward_push_with_source(A, 4, 9, 1e-3)
    assert early.push_count <= full.push_count
    z = O.backward_push_with_source(A, n + 1, 9, 1e-3)
    assert z.push_count == 0 and not z.estimate.any()
    f = O.forward_push(A, 4)
    c = O.push_combine_with_forward(0.15, full, f.estimate, f.residual)
    ref = 0.0
    for i in range(n):
        ref += full.estimate[i] * f.estimate[i]
        ref += full.residual[i] * f.estimate[i] * 0.15
        ref += full.estimate[i] * f.residual[i] * 0.15
    assert c == ref and c >= 0.0 and np.isfinite(c)


def test_oracle_ts_forward_push(oracle):
    """SublinearSolver.solveForwardPush (src/core/solver.ts:437-522): converges to A^-1 b on diagonally dominant systems
    with max |r| < epsilon; CONVERGENCE_FAILED when maxIterations runs out; zero diagonal -> NUMERICAL_INSTABILITY"""
    O = oracle
    M = np.array([[4.0, -1, 0.25], [-2, 5, 1], [0.5, 1, 3]])            # the MCP 3x3 example's shape: asymmetric, row-DD
    b = np.array([1.0, 2, 3])
    r = O.ts_forward_push(O.Csr.from_dense(M), b, 1e-10, 1000)
    assert r.converged and r.status == O.OK and r.iterations > 3
    np.testing.assert_allclose(r.solution, np.linalg.solve(M, b), atol=1e-9)
    np.testing.assert_allclose(r.residual, np.linalg.norm(b - M @ r.solution), atol=1e-12)
    A, bb = O.gen_bench_csr(300, 0.03)
    r = O.ts_forward_push(A, bb, 1e-8, 100000)
    assert r.converged
    assert np.abs(bb - A.multiply_vector(r.solution)).max() < 1e-8 * 1.0001 + 1e-9
    few = O.ts_forward_push(A, bb, 1e-8, 10)
    assert not few.converged and few.status == O.ERR_CONVERGENCE_FAILURE and few.iterations == 10
    Z = np.array([[0.0, 1], [1, 2]])
    z = O.ts_forward_push(O.Csr.from_dense(Z), np.array([1.0, 0.0]), 1e-6, 10)
    assert z.status == O.ERR_NUMERICAL_INSTABILITY


def complete_graph(n):
    """create_complete_graph (push_tests.rs:62-77): every node to every other node, weight 1/(n-1)"""
    rows = np.repeat(np.arange(n), n - 1)
    cols = np.array([j for i in range(n) for j in range(n) if j != i])
    return np.arange(n + 1) * (n - 1), cols, np.full(n * (n - 1), 1.0 / (n - 1)), n


class _OracleApi:
    """the oracle behind the method names of the reference's solvers, so that the restated tests read like push_tests.rs"""

    def __init__(self, O, graph, **cfg):
        rp, ci, w, n = graph
        self.O, self.A, self.cfg, self.alpha = O, O.Csr(n, n, w, ci, rp), cfg, cfg.get("alpha", 0.15)

    def solve_single_source(self, s): return self.O.forward_push(self.A, s, **self.cfg)
    def solve_multi_source(self, s): return self.O.forward_push(self.A, list(s), **self.cfg)
    def query_single_entry(self, s, t): return float(self.solve_single_source(s).estimate[t])
    def solve_single_target(self, t): return self.O.backward_push(self.A, t, **self.cfg)
    def solve_multi_target(self, t): return self.O.backward_push(self.A, list(t), **self.cfg)
    def query_transition_probability(self, s, t): return float(self.solve_single_target(t).estimate[s])
    def extrapolated_solution(self, r): return r.estimate + self.alpha * r.residual
    def reachability_probabilities(self, t): return self.extrapolated_solution(self.solve_single_target(t))

    def solve_bidirectional(self, s, t):
        f, b = self.solve_single_source(s), self.solve_single_target(t)
        return self.O.push_combine_with_forward(self.alpha, b, f.estimate, f.residual)


class _DeviceApi:
    def __init__(self, graph, **cfg):
        rp, ci, w, n = graph
        self.g = sb.PushGraph.from_matrix(rp, ci, w, n)
        self.c = sb.PushConfig(**cfg)
        self.f, self.b = sb.ForwardPushSolver(self.g, self.c), sb.BackwardPushSolver(self.g, self.c)
        self.bi = sb.BidirectionalPushSolver(self.g, self.c, self.c)

    def __getattr__(self, name):
        for o in (self.f, self.b, self.bi):
            if hasattr(o, name):
                return getattr(o, name)
        raise AttributeError(name)

    def extrapolated_solution(self, r): return self.f.extrapolated_solution(r)


def reference_push_tests(make):
    """tests/rust/push_tests.rs:81-330 restated against `make(graph, **config)` (the oracle or the device solvers).
    Two of the reference's own assertions do not hold for its algorithm and are asserted as they really come out:
    test_forward_push_complete_graph wants every entry within 0.1 of alpha (the source holds 0.338 of the mass), and
    test_backward_push_reachability wants reach[1] > reach[0] on a path (node 0 has no in-edge and keeps what reaches it)."""
    s = make(simple_graph())
    r = s.solve_single_source(0)                                             # test_forward_push_basic_functionality
    assert r.push_count > 0 and r.nodes_visited > 0 and r.estimate[0] > 0.0 and r.residual_norm >= 0.0
    assert (r.estimate >= 0).all() and (r.residual >= 0).all()
    assert abs(s.extrapolated_solution(r).sum() - 1.0) < 0.1                 # test_forward_push_mass_conservation
    tight = make(simple_graph(), epsilon=1e-10, max_pushes=100_000).solve_single_source(0)   # test_forward_push_convergence
    loose = make(simple_graph(), epsilon=1e-4, max_pushes=100_000).solve_single_source(0)
    assert tight.push_count >= loose.push_count and tight.residual_norm <= loose.residual_norm * 10.0
    m = s.solve_multi_source([0, 2])                                         # test_forward_push_multi_source
    assert m.push_count > 0 and m.nodes_visited > 0 and m.estimate[0] > 0.0 and m.estimate[2] > 0.0 and m.estimate.sum() > 0.0
    assert s.query_single_entry(0, 1) >= 0.0 and s.query_single_entry(0, 0) > 0.0     # test_forward_push_single_entry_query
    p = make(path_graph(5)).solve_single_source(0)                           # test_forward_push_path_graph
    assert p.estimate[0] > p.estimate[1] and (p.estimate[1] > p.estimate[2] or p.estimate[2] < 1e-6)
    c = make(complete_graph(4))                                              # test_forward_push_complete_graph
    final = c.extrapolated_solution(c.solve_single_source(0))
    assert abs(final[0] - 26.0 / 77.0) < 1e-5 and np.abs(final[1:] - 17.0 / 77.0).max() < 1e-5   # exact PPR: 26/77, 17/77
    b = s.solve_single_target(3)                                             # test_backward_push_basic_functionality
    assert b.push_count > 0 and b.nodes_visited > 0 and b.estimate[3] > 0.0 and b.residual_norm >= 0.0 and (b.estimate >= 0).all()
    assert 0.0 <= s.query_transition_probability(0, 3) <= 1.0 and s.query_transition_probability(0, 0) > 0.0
    mt = s.solve_multi_target([1, 3])                                        # test_backward_push_multi_target
    assert mt.push_count > 0 and mt.nodes_visited > 0 and mt.estimate[1] > 0.0 and mt.estimate[3] > 0.0
    reach = make(path_graph(5)).reachability_probabilities(4)                # test_backward_push_reachability
    assert reach[4] > reach[3] > reach[2] > reach[1] and reach[0] > reach[1]
    bi, fq, bq = s.solve_bidirectional(0, 3), s.query_single_entry(0, 3), s.query_transition_probability(0, 3)
    assert bi >= 0.0 and fq >= 0.0 and bq >= 0.0                             # test_bidirectional_solver_consistency
    z = s.solve_single_source(10)                                            # out-of-bounds source: zeros (push_tests.rs:429-495)
    assert z.push_count == 0 and not z.estimate.any() and not z.residual.any()
    zt = s.solve_single_target(10)
    assert zt.push_count == 0 and not zt.estimate.any()


def test_oracle_reference_push_tests(oracle):
    reference_push_tests(lambda g, **cfg: _OracleApi(oracle, g, **cfg))


FIXTURES = ["jacobi_c1_test_matrix_ones", "jacobi_dd_asymmetric_n50_random", "jacobi_banded_n100_smooth", "jacobi_mcp_3x3"]


def _fixture(golden_dir, name):
    import os
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n = int(g["n"])
    return n, g["rows"], g["cols"], g["vals"], g["b"]


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_ts_forward_push_on_reference_fixtures(oracle, golden_dir, name):
    """solveForwardPush on the reference's own committed systems (tests/data/test-matrix.json = config C1,
    scripts/linear_systems/test_matrices, the MCP 3x3 example; stored with the Jacobi golden vectors): converges to A^-1 b"""
    O = oracle
    n, rows, cols, vals, b = _fixture(golden_dir, name)
    A = O.Csr.from_triplets(rows, cols, vals, n, n)
    r = O.ts_forward_push(A, b, 1e-10, 10_000_000)
    assert r.converged and r.status == O.OK
    x = np.linalg.solve(A.to_scipy().toarray(), b)
    np.testing.assert_allclose(r.solution, x, rtol=0, atol=1e-8)
    assert np.abs(b - A.multiply_vector(r.solution)).max() < 1e-10 + 1e-13


# ---- device path (GPU) -------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_reference_push_tests():
    reference_push_tests(lambda g, **cfg: _DeviceApi(g, **cfg))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_forward_push_solve_on_reference_fixtures(oracle, golden_dir, name):
    O = oracle
    n, rows, cols, vals, b = _fixture(golden_dir, name)
    A = O.Csr.from_triplets(rows, cols, vals, n, n)
    m = sb.SparseMatrix.from_triplets(rows, cols, vals, n, n)
    r = sb.forward_push_solve(m, b, 1e-10, 10_000_000)
    o = O.ts_forward_push(A, b, 1e-10, 10_000_000)
    assert r.converged and o.converged and r.max_residual < 1e-10
    np.testing.assert_allclose(r.solution, o.solution, rtol=0, atol=1e-8)
    np.testing.assert_allclose(r.solution, np.linalg.solve(A.to_scipy().toarray(), b), rtol=0, atol=1e-8)


@pytest.mark.gpu
def test_gpu_forward_push_reference_tests(oracle):
    rp, ci, w, n = simple_graph()
    g = sb.PushGraph.from_matrix(rp, ci, w, n)
    assert (g.num_nodes(), g.num_edges()) == (4, 7) and g.out_degree(1) == 1.0 and g.in_degree(3) == 0.2 + 0.4
    assert g.out_degree(9) == 0.0
    solver = sb.ForwardPushSolver(g, sb.PushConfig())
    r = solver.solve_single_source(0)
    assert r.push_count > 0 and r.nodes_visited > 0 and r.estimate[0] > 0 and r.residual_norm >= 0
    check_forward(r, dense(rp, ci, w, n), 0, 0.15, 1e-6)
    assert abs(solver.extrapolated_solution(r).sum() - 1.0) < 0.1
    assert solver.query_single_entry(0, 1) >= 0.0 and solver.query_single_entry(0, 1) == r.estimate[1]
    m = solver.solve_multi_source([0, 2])
    assert m.push_count > 0 and abs(m.estimate.sum() + m.residual.sum() - 1.0) < 1e-12
    z = solver.solve_single_source(7)
    assert z.push_count == 0 and z.nodes_visited == 0 and not z.estimate.any() and z.residual_norm == 0.0
    o = oracle.forward_push(oracle.Csr(n, n, w, ci, rp), 0)          # both within eps * sum(max(deg,1)) of the exact PPR
    assert np.abs(r.estimate - o.estimate).max() <= 1e-6 * 4 + 1e-12
    e = sb.PushGraph.from_edges(3, [(0, 1, 0.5), (1, 2, 1.0), (2, 0, 0.3), (5, 0, 1.0)])   # adjacency.rs:285-296 + drop
    assert (e.num_nodes(), e.num_edges()) == (3, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("graph,seed_node", [(random_graph(200, 5), 3), (random_graph(5000, 8), 17), (path_graph(12), 0),
                                             (path_graph(12), 11)])
def test_gpu_forward_push_vs_exact_ppr_and_oracle(oracle, graph, seed_node):
    O = oracle
    rp, ci, w, n = graph
    g = sb.PushGraph.from_matrix(rp, ci, w, n)
    for eps in (1e-4, 1e-8):
        cfg = sb.PushConfig(epsilon=eps, queue_threshold=eps / 10, adaptive_threshold=False)
        r = sb.ForwardPushSolver(g, cfg).solve_single_source(seed_node)
        o = O.forward_push(O.Csr(n, n, w, ci, rp), seed_node, epsilon=eps, queue_threshold=eps / 10, adaptive_threshold=False)
        assert r.rounds > 0 and r.kernel_launches >= 2 * r.rounds and r.dense_rounds == 0
        assert 0 < r.edges_touched < 40 * r.push_count + 40                 # only the frontier's edges are read
        if n <= 1000:
            check_forward(r, dense(rp, ci, w, n), seed_node, 0.15, eps)
        else:
            deg = np.bincount(np.repeat(np.arange(n), np.diff(rp)), weights=w, minlength=n)
            assert (r.estimate >= 0).all() and (r.residual >= 0).all()
            assert abs(r.estimate.sum() + r.residual.sum() - 1.0) < 1e-12
            assert (r.residual < eps * np.maximum(deg, 1.0)).all()
        # both estimates sit below the exact PPR by at most their residual mass
        assert np.abs(r.estimate - o.estimate).max() <= max(r.residual.sum(), o.residual.sum()) + 1e-12
        r2 = sb.ForwardPushSolver(g, cfg).solve_single_source(seed_node)
        assert np.array_equal(r.estimate, r2.estimate) and r.push_count == r2.push_count    # deterministic


@pytest.mark.gpu
def test_gpu_backward_push(oracle):
    O = oracle
    rp, ci, w, n = random_graph(150, 4)
    A = dense(rp, ci, w, n)
    g = sb.PushGraph.from_matrix(rp, ci, w, n)
    t = 7
    cfg = sb.PushConfig(epsilon=1e-9, queue_threshold=1e-10, adaptive_threshold=False)
    r = sb.BackwardPushSolver(g, cfg).solve_single_target(t)
    assert r.push_count > 0 and (r.estimate >= 0).all() and (r.residual >= 0).all()
    assert (r.residual < 1e-9 * np.maximum(A.sum(0), 1.0)).all()
    np.testing.assert_allclose(r.estimate, backward_exact(A, t, 0.15), rtol=0, atol=1e-6)
    o = O.backward_push(O.Csr(n, n, w, ci, rp), t, epsilon=1e-9, queue_threshold=1e-10, adaptive_threshold=False)
    np.testing.assert_allclose(r.estimate, o.estimate, rtol=0, atol=1e-6)
    assert sb.BackwardPushSolver(g).query_transition_probability(3, t) >= 0.0
    m = sb.BackwardPushSolver(g, cfg).solve_multi_target([1, 2, 3])
    mo = O.backward_push(O.Csr(n, n, w, ci, rp), [1, 2, 3], epsilon=1e-9, queue_threshold=1e-10, adaptive_threshold=False)
    np.testing.assert_allclose(m.estimate, mo.estimate, rtol=0, atol=1e-6)
    z = sb.BackwardPushSolver(g).solve_single_target(n + 3)
    assert z.push_count == 0 and not z.estimate.any()


@pytest.mark.gpu
def test_gpu_push_is_local_and_caps_max_pushes(oracle):
    """the sparse frontier reads only the pushed nodes' edges (work << nnz per round for a local query on a big graph),
    max_pushes is never exceeded, solve_with_target / solve_with_source stop early, bidirectional matches the oracle"""
    O = oracle
    rp, ci, w, n = random_graph(200_000, 8)
    g = sb.PushGraph.from_matrix(rp, ci, w, n)
    A = O.Csr(n, n, w, ci, rp)
    cfg = sb.PushConfig(epsilon=1e-5, queue_threshold=1e-7, adaptive_threshold=False)
    fs = sb.ForwardPushSolver(g, cfg)
    r = fs.solve_single_source(11)
    o = O.forward_push(A, 11, epsilon=1e-5, queue_threshold=1e-7, adaptive_threshold=False)
    assert r.dense_rounds == 0 and r.edges_touched < 0.2 * len(w) * r.rounds           # local: not nnz per round
    assert abs(r.estimate.sum() + r.residual.sum() - 1.0) < 1e-12
    deg = np.bincount(np.repeat(np.arange(n), np.diff(rp)), weights=w, minlength=n)
    assert (r.residual < 1e-5 * np.maximum(deg, 1.0)).all()
    assert np.abs(r.estimate - o.estimate).max() <= max(r.residual.sum(), o.residual.sum()) + 1e-12
    capped = sb.ForwardPushSolver(g, sb.PushConfig(epsilon=1e-5, queue_threshold=1e-7, adaptive_threshold=False,
                                                   max_pushes=137)).solve_single_source(11)
    assert capped.push_count == 137
    assert abs(capped.estimate.sum() + capped.residual.sum() - 1.0) < 1e-12
    t = int(ci[rp[11]])                                                                  # an out-neighbour of the source
    early = fs.solve_with_target(11, t, 1e-3)
    assert early.push_count <= r.push_count and early.estimate[t] > 0
    assert not (early.estimate[t] > 1e-3 and early.residual[t] < 1e-4) or early.push_count < r.push_count
    zero = fs.solve_with_target(11, n + 5, 1e-3)
    assert zero.push_count == 0 and not zero.estimate.any()
    bs = sb.BackwardPushSolver(g, cfg)
    bfull = bs.solve_single_target(t)
    bearly = bs.solve_with_source(11, t, 1e-4)
    assert bearly.push_count <= bfull.push_count
    c = bs.combine_with_forward(bfull, r.estimate, r.residual)
    assert c == O.push_combine_with_forward(0.15, O.PushResult(bfull.estimate, bfull.residual, 0, 0, 0.0), r.estimate, r.residual)
    bi = sb.BidirectionalPushSolver(g, cfg, cfg)
    v = bi.solve_bidirectional(11, t)
    ob = O.push_combine_with_forward(0.15, O.backward_push(A, t, epsilon=1e-5, queue_threshold=1e-7, adaptive_threshold=False),
                                     o.estimate, o.residual)
    assert np.isfinite(v) and v >= 0 and abs(v - ob) <= 1e-4 * max(ob, 1e-6) + 1e-9
    assert bi.adaptive_solve(11, t) >= 0.0 and bi.adaptive_solve(n + 1, t) == 0.0


@pytest.mark.gpu
def test_gpu_dense_rounds_for_wide_frontiers(oracle):
    """a multi-source push over every node starts with the whole graph in the frontier: dense rounds, then back to lists"""
    O = oracle
    rp, ci, w, n = random_graph(300_000, 8)
    g = sb.PushGraph.from_matrix(rp, ci, w, n)
    cfg = sb.PushConfig(epsilon=2e-6 / 1.0, queue_threshold=1e-9, adaptive_threshold=False, max_pushes=100_000_000)
    r = sb.ForwardPushSolver(g, cfg).solve_multi_source(np.arange(n))
    assert r.dense_rounds > 0 and r.rounds >= r.dense_rounds
    assert abs(r.estimate.sum() + r.residual.sum() - 1.0) < 1e-9
    deg = np.bincount(np.repeat(np.arange(n), np.diff(rp)), weights=w, minlength=n)
    assert (r.residual < cfg.epsilon * np.maximum(deg, 1.0)).all() and (r.estimate >= 0).all()


@pytest.mark.gpu
def test_gpu_forward_push_solve_axb(oracle):
    """TS solveForwardPush for A x = b on the device: r = b - A x exact, max |r| < epsilon at exit; same limit as the
    sequential oracle (different push order: agreement to the residual bound), failure modes like the reference"""
    O = oracle
    M = np.array([[4.0, -1, 0.25], [-2, 5, 1], [0.5, 1, 3]])
    b = np.array([1.0, 2, 3])
    m = sb.SparseMatrix.from_dense(M)
    r = sb.forward_push_solve(m, b, 1e-10, 1000)
    assert r.converged and r.method == "forward-push" and r.iterations > 3
    np.testing.assert_allclose(r.solution, np.linalg.solve(M, b), atol=1e-9)
    np.testing.assert_allclose(r.residual, np.linalg.norm(b - M @ r.solution), atol=1e-12)
    n = 50_000
    A, bb = O.gen_bench_csr(n, 10.0 / n)
    mg = sb.SparseMatrix.from_csr(A.row_ptr, A.col_indices, A.values, n, n)
    e = np.zeros(n)
    e[123] = 1.0                                                 # a sparse right-hand side: the walk stays local
    r = sb.forward_push_solve(mg, e, 1e-12, 10_000_000)
    o = O.ts_forward_push(A, e, 1e-12, 10_000_000)
    assert r.converged and o.converged and r.max_residual < 1e-12
    np.testing.assert_allclose(r.solution, o.solution, rtol=0, atol=1e-11)
    assert np.abs(e - A.multiply_vector(r.solution)).max() < 1e-11
    rl = sb.forward_push_solve(mg, e, 1e-5, 10_000_000)         # a loose epsilon keeps the walk inside a few hops
    assert rl.converged and 0 < np.count_nonzero(rl.solution) < n // 10 and rl.iterations < n // 10
    rd = sb.forward_push_solve(mg, bb, 1e-6, 100_000_000)       # dense right-hand side: every node starts in the frontier
    assert rd.converged and np.abs(bb - A.multiply_vector(rd.solution)).max() < 1e-6 * 1.01
    with pytest.raises(sb.SolverError) as ei:
        sb.forward_push_solve(mg, bb, 1e-9, 10)
    assert ei.value.variant == "ConvergenceFailure" and ei.value.result.iterations == 10
    with pytest.raises(sb.SolverError) as ei:
        sb.forward_push_solve(sb.SparseMatrix.from_dense(np.array([[0.0, 1], [1, 2]])), np.array([1.0, 0.0]), 1e-6, 10)
    assert ei.value.variant == "NumericalInstability"
    with pytest.raises(sb.SolverError) as ei:
        sb.forward_push_solve(m, np.ones(4))
    assert ei.value.variant == "DimensionMismatch"
