"""Generate golden vectors for the conjugate-gradient consumer of the push path's SpMV
from REFERENCE-AUTHORED code run in the build container (only the committed .npz files travel to the GPU box).

Source of truth: IterativeSolvers.conjugate_gradient in
/root/reference/scripts/linear_systems/iterative_solvers.py:289-370 (pure numpy, imported unmodified). It runs the same
recurrence as OptimizedConjugateGradientSolver::solve (src/optimized_solver.rs:182-295) from x0 = 0: after k passes
both hold the same x_k, r_k up to the summation order of the dot products (BLAS here, sequential loops in Rust).  Its
stopping test `sqrt(rsnew) < tol` after the update corresponds to the Rust `rsold <= tol^2` at the top of the next
pass, so for a given tolerance the pass counts follow from the stored residual history.  The numpy code has no
`|p.Ap| < 1e-16 -> break` guard (src/optimized_solver.rs:234-236); that guard stops the Rust loop once ||r|| falls to
~1e-8..1e-9, so the Rust solver cannot reach the 1e-10 these vectors were run to: tests compare pass by pass above
that level and check the guard's effect separately.

Inputs are the reference's own committed symmetric fixtures:
  scripts/linear_systems/test_matrices/n_*/{dd_symmetric,spd_well_conditioned,tridiagonal,laplacian_1d}.json

Run:  python tests/golden/make_golden_cg.py        (needs /root/reference)
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REF, "scripts", "linear_systems"))
from iterative_solvers import IterativeSolvers  # noqa: E402  (reference code, imported unmodified)


def run_case(name, A, b, tol, max_iter=1000):
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    iterates = []
    res = IterativeSolvers().conjugate_gradient(A, b, max_iter=max_iter, tol=tol,
                                                callback=lambda it, x, r: iterates.append(x.copy()))
    assert res["success"], name
    r, c = np.nonzero(A)
    keep = list(range(len(iterates)))  # every pass: the systems are small (n <= 200, <= 40 passes)
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        n=np.int64(A.shape[0]), rows=r.astype(np.uint32), cols=c.astype(np.uint32), vals=A[r, c], b=b,
        solution=res["solution"], iterations=np.int64(res["iterations"]),
        residual_history=np.asarray(res["convergence_history"]), tol=np.float64(tol),
        kept_passes=np.asarray([k + 1 for k in keep], dtype=np.int64),
        kept_iterates=np.stack([iterates[k] for k in keep]))
    print(f"{name}: n={A.shape[0]} nnz={len(r)} passes={res['iterations']} residual={res['residual']:.3e}")


def main():
    base = os.path.join(REF, "scripts", "linear_systems", "test_matrices")
    for n, kind, rhs in [(50, "dd_symmetric", "ones"), (100, "dd_symmetric", "random"),
                         (100, "spd_well_conditioned", "random"), (200, "tridiagonal", "smooth"),
                         (50, "laplacian_1d", "ones")]:
        d = json.load(open(os.path.join(base, f"n_{n}", f"{kind}.json")))
        run_case(f"cg_{kind}_n{n}_{rhs}", d["matrix"], d["rhs_vectors"][rhs], tol=1e-10)
    # the Rust unit-test system [[4,1],[1,3]] x = [1,2] (src/optimized_solver.rs:380-418, src/solver_core.rs:256-278)
    run_case("cg_rust_unit_2x2", [[4, 1], [1, 3]], [1.0, 2.0], tol=1e-10)


if __name__ == "__main__":
    main()
