"""Forward / backward push: ForwardPushSolver / BackwardPushSolver
(src/solver/forward_push.rs, src/solver/backward_push.rs, graphs of tests/rust/push_tests.rs:15-77).

CPU part: the oracle's sequential restatement against the properties the reference's own tests assert and against the
exact personalised PageRank  pi_s = alpha (I - (1-alpha) P^T)^-1 e_s  (P = row-normalised adjacency).
GPU part: the frontier-synchronous device version has the same push rule and stopping condition, so it satisfies the
same invariants and the same error bound; it is compared with the exact PPR and with the oracle within that bound
(the reference's pop order is undefined — its queue item has no Ord impl — so there is nothing bitwise to pin).

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


# ---- device path (GPU) -------------------------------------------------------------------------------------------

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
        assert r.rounds > 0 and r.kernel_launches >= 2 * r.rounds
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
