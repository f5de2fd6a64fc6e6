"""Timing of the two SURVEY §8 configurations bench.py does not cover (diagnostic, single GPU):
  C3  PageRank (alpha = 0.85) on a synthetic power-law graph, n = 10 M / 100 M edges, eps = 1e-6
  C4  1 024 single-entry random-walk queries (eps = 0.01 -> 10 000 walks each) on the n = 10 M gen_bench system
python tests/config_timing.py [c3|c4|both] [n]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import sublinear_b200 as sb  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
sb.set_device(0)

if which in ("c3", "both"):
    rng = np.random.default_rng(42)
    ne, alpha = 10 * n, 0.85
    t0 = time.perf_counter()
    src = rng.integers(0, n, ne)
    dst = np.minimum((rng.pareto(1.1, ne) * 50).astype(np.int64), n - 1)      # heavy hubs (power-law in-degree)
    t1 = time.perf_counter()
    S, rhs = sb.SparseMatrix.pagerank_system(src, dst, n, alpha)
    t2 = time.perf_counter()
    del src, dst
    opt = sb.SolverOptions(dominance=sb.DOMINANCE_ROW_OR_COL, tolerance=1e-6, collect_stats=True, enable_profiling=True)
    solver = sb.NeumannSolver.new(200, 1e-9)
    r = solver.solve(S, rhs, opt)
    best = min((solver.solve(S, rhs, opt) for _ in range(3)), key=lambda q: q.device_time_ms)
    res = float(np.linalg.norm(S.multiply_vector(best.solution) - rhs))
    print(json.dumps({"config": "C3 pagerank", "n": n, "edges": ne, "nnz": S.nnz(), "layout": S.storage_info()["layout"],
                      "graph_s": t1 - t0, "system_build_s": t2 - t1, "converged": best.converged,
                      "iterations": best.iterations, "terms": best.terms_computed, "matvecs": best.matvec_count,
                      "device_ms": best.device_time_ms, "nnz_per_s": S.nnz() * best.matvec_count / (best.device_time_ms * 1e-3),
                      "push_kernel_avg_us": 1e3 * best.push_kernel_ms / max(best.push_kernel_count, 1),
                      "residual_norm": best.residual_norm, "residual_check": res, "mass": float(best.solution.sum()),
                      "min": float(best.solution.min())}), flush=True)
    del S

if which in ("c4", "both"):
    rp, ci, v, b = sb.gen_bench_csr(n, 10.0 / n)
    m = sb.SparseMatrix.from_csr(rp, ci, v, n, n)
    del rp, ci, v
    x = sb.NeumannSolver.default().solve(m, b).solution
    state, rows = 12345, []
    for _ in range(1024):                            # rows from the reference's 32-bit LCG (src/core/utils.ts:161-168)
        state = (state * 1664525 + 1013904223) % 2 ** 32
        rows.append(state % n)
    rows = np.asarray(rows)
    sb.solve_entry(m, b, rows[:8], eps=0.1, seed=1)  # warm-up (setup of the walk tables)
    t0 = time.perf_counter()
    est, var = sb.solve_entry(m, b, rows, eps=0.01, seed=7)
    dt = time.perf_counter() - t0
    se = np.sqrt(var / 10000)
    print(json.dumps({"config": "C4 solve_entry batch", "n": n, "queries": 1024, "walks_per_query": 10000,
                      "seconds": dt, "queries_per_s": 1024 / dt, "walks_per_s": 1024 * 10000 / dt,
                      "within_5_standard_errors": bool((np.abs(est - x[rows]) <= 5 * se + 1e-9).all()),
                      "max_abs_error": float(np.abs(est - x[rows]).max())}), flush=True)
