"""Diagnostic (single GPU): the sparse-frontier push at scale — forward / backward push on a 10 M-node graph with 80 M edges
and the TS forward push for A x = b on the n = 10 M gen_bench system (python tests/push_timing.py [n])."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import sublinear_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sb.set_device(0)
rng = np.random.default_rng(7)
k = 8
t0 = time.perf_counter()
cols = rng.integers(0, n, n * k, dtype=np.int64).astype(np.uint32)          # k uniform-random out-edges per node, weight 1/k
rp = np.arange(n + 1, dtype=np.uint64) * k
w = np.full(n * k, 1.0 / k)
t1 = time.perf_counter()
g = sb.PushGraph.from_matrix(rp, cols, w, n)
t2 = time.perf_counter()
for eps in (1e-5, 1e-6, 1e-7):
    cfg = sb.PushConfig(epsilon=eps, queue_threshold=eps / 100, adaptive_threshold=False, max_pushes=50_000_000)
    for name, solver, run in (("forward", sb.ForwardPushSolver(g, cfg), lambda s: s.solve_single_source(12345)),
                              ("backward", sb.BackwardPushSolver(g, cfg), lambda s: s.solve_single_target(12345))):
        run(solver)
        tw = time.perf_counter()
        r = run(solver)
        wall = time.perf_counter() - tw
        print(json.dumps({"config": f"{name} push", "n": n, "edges": n * k, "epsilon": eps, "push_count": r.push_count,
                          "nodes_visited": r.nodes_visited, "rounds": r.rounds, "dense_rounds": r.dense_rounds,
                          "edges_touched": r.edges_touched, "edges_touched_over_nnz": r.edges_touched / (n * k),
                          "device_ms": r.device_time_ms, "wall_ms_incl_2x80MB_result_copies": wall * 1e3,
                          "mass": float(r.estimate.sum() + r.residual.sum()) if name == "forward" else None,
                          "pushes_per_s_device": r.push_count / max(r.device_time_ms * 1e-3, 1e-9)}), flush=True)
print(json.dumps({"config": "push graph build", "n": n, "edges": n * k, "generate_s": t1 - t0, "build_s": t2 - t1}), flush=True)
del g
rpm, ci, v, b = sb.gen_bench_csr(n, 10.0 / n)
m = sb.SparseMatrix.from_csr(rpm, ci, v, n, n)
e = np.zeros(n)
e[4242] = 1.0
for eps in (1e-8, 1e-12):
    sb.forward_push_solve(m, e, eps, 100_000_000)
    tw = time.perf_counter()
    r = sb.forward_push_solve(m, e, eps, 100_000_000)
    wall = time.perf_counter() - tw
    print(json.dumps({"config": "TS forward push A x = e_i", "n": n, "nnz": len(v), "epsilon": eps, "iterations": r.iterations,
                      "rounds": r.rounds, "converged": r.converged, "max_residual": r.max_residual,
                      "nonzeros_in_x": int(np.count_nonzero(r.solution)), "wall_ms": wall * 1e3}), flush=True)
