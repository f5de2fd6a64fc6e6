"""Generate golden vectors for the Neumann/push path from REFERENCE-AUTHORED code run in the build
container (the GPU box has no /root/reference; only the committed .npz/.json files travel).

Source of truth: IterativeSolvers.jacobi in /root/reference/scripts/linear_systems/iterative_solvers.py:17-105
(pure numpy, imported unmodified).  Jacobi from x0 = 0 after k sweeps equals the Neumann partial sum
sum_{j<k} (-D^-1 R)^j D^-1 b that NeumannSolver documents (src/solver/neumann.rs:16-22), so its iterates,
residual history and final solution pin the oracle's `correct` mode.

Inputs are the reference's own committed fixtures:
  scripts/linear_systems/test_matrices/n_*/{dd_asymmetric,banded,tridiagonal}.json (+ rhs_vectors)
  tests/data/test-matrix.json (config C1, 1000x1000, b = 1; the committed vector file is empty)
  docs/testing/test_matrix.json (3x3 of the MCP examples)

Run:  python tests/golden/make_golden.py        (needs /root/reference; takes ~1-2 min)
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "scripts", "linear_systems"))
from iterative_solvers import IterativeSolvers  # noqa: E402  (reference code, imported unmodified)


def run_case(name, A, b, tol, max_iter=200):
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    iterates = []
    res = IterativeSolvers().jacobi(A, b, max_iter=max_iter, tol=tol,
                                    callback=lambda it, x, r: iterates.append(x.copy()))
    assert res["success"], name
    r, c = np.nonzero(A)
    keep = [0, 1, 2, 4, len(iterates) - 1]
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        n=np.int64(A.shape[0]), rows=r.astype(np.uint32), cols=c.astype(np.uint32), vals=A[r, c], b=b,
        solution=res["solution"], iterations=np.int64(res["iterations"]),
        residual_history=np.asarray(res["convergence_history"]), tol=np.float64(tol),
        kept_sweeps=np.asarray([k + 1 for k in keep if k < len(iterates)], dtype=np.int64),
        kept_iterates=np.stack([iterates[k] for k in keep if k < len(iterates)]))
    print(f"{name}: n={A.shape[0]} nnz={len(r)} sweeps={res['iterations']} residual={res['residual']:.3e}")


def main():
    base = os.path.join(REF, "scripts", "linear_systems", "test_matrices")
    for n, kind, rhs in [(50, "dd_asymmetric", "ones"), (50, "dd_asymmetric", "random"),
                         (100, "banded", "smooth"), (100, "dd_symmetric", "random"),
                         (200, "tridiagonal", "ones")]:
        d = json.load(open(os.path.join(base, f"n_{n}", f"{kind}.json")))
        run_case(f"jacobi_{kind}_n{n}_{rhs}", d["matrix"], d["rhs_vectors"][rhs], tol=1e-10)
    d = json.load(open(os.path.join(REF, "docs", "testing", "test_matrix.json")))
    A = d["data"] if isinstance(d, dict) and "data" in d else d
    run_case("jacobi_docs_3x3", A, [1.0, 2.0, 1.0], tol=1e-12)
    # the documented MCP example A=[[4,-1,0],[-1,4,-1],[0,-1,3]], b=[1,2,1] (docs/reference/MCP_TOOL_TEST_RESULTS.md:49-55)
    run_case("jacobi_mcp_3x3", [[4, -1, 0], [-1, 4, -1], [0, -1, 3]], [1.0, 2.0, 1.0], tol=1e-12)
    d = json.load(open(os.path.join(REF, "tests", "data", "test-matrix.json")))
    A = np.asarray(d["data"], dtype=np.float64)
    run_case("jacobi_c1_test_matrix_ones", A, np.ones(A.shape[0]), tol=1e-8, max_iter=50)


if __name__ == "__main__":
    main()
