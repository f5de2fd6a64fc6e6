"""Golden vectors for the forward-push oracle (tests/golden/push_*.npz).

The reference's push solvers cannot be run (src/graph/mod.rs does not compile: its heap item has no Ord impl; no Rust
toolchain in this image either), so the vectors come from an INDEPENDENT pure-Python restatement of
ForwardPushSolver::solve_single_source / solve_multi_source (src/solver/forward_push.rs:66-216) + WorkQueue
(src/graph/mod.rs:130-212) with the queue ordered by derive(PartialOrd) on (priority, node_id) — the function
py_forward_push of tests/test_push.py — on the graphs of tests/rust/push_tests.rs:15-59. The C oracle must reproduce them
bit for bit (tests/test_push.py::test_oracle_push_pinned_to_independent_restatement).
Run from the repository root: python tests/golden/make_golden_push.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import test_push as T  # noqa: E402

CASES = [("simple4", T.simple_graph(), 0), ("lcg100x5", T.random_graph(100, 5), 3),
         ("lcg100x5_multi", T.random_graph(100, 5), [0, 2, 50]), ("path12", T.path_graph(12), 0)]
for name, (rp, ci, w, n), src in CASES:
    est, res, pushes = T.py_forward_push(rp, ci, w, n, src)
    np.savez(os.path.join(HERE, f"push_{name}.npz"), estimate=est, residual=res, push_count=pushes)
    print(name, pushes, est.sum() + res.sum())
