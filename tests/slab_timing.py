"""Diagnostic: push-kernel time when the gathered columns are confined to a slab of the term vector
(python tests/slab_timing.py [n] [slab_cols ...]) — sizes the L2 working set a column-blocked SpMV would have."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sublinear_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
slabs = [int(a) for a in sys.argv[2:]] or [n, n // 2, n // 3, n // 4, n // 8]
sb.set_device(0)
k = 10
rng = np.random.default_rng(1)
for w in slabs:
    # row i: diagonal 10 + 0.01 i, nine positive off-diagonals at uniform-random columns of [0, w) (never the diagonal)
    cols = rng.integers(0, w, size=(n, k - 1), dtype=np.int64)
    i = np.arange(n, dtype=np.int64)[:, None]
    cols = np.where(cols == i, (cols + 1) % w, cols)
    diag = 10.0 + 0.01 * np.arange(n)
    vals = rng.random((n, k - 1)) * (diag / (2.0 * k))[:, None]
    c = np.concatenate([i, cols], axis=1)
    v = np.concatenate([diag[:, None], vals], axis=1)
    order = np.argsort(c, axis=1, kind="stable")
    c = np.take_along_axis(c, order, axis=1).astype(np.uint32).ravel()
    v = np.take_along_axis(v, order, axis=1).ravel()
    rp = np.arange(n + 1, dtype=np.uint64) * k
    m = sb.SparseMatrix.from_csr(rp, c, v, n, n)
    del cols, vals, order
    bd = torch.ones(n, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(3):
        norms, ms = sb.push_iterations_dev(m, bd.data_ptr(), n, 12)
        best = min(best, ms / 12 * 1e3)
    alg = 12 * len(v) + 44 * n + 4
    print(f"columns in [0,{w}) = {8 * w / 1e6:.0f} MB slab (+ the diagonal): layout={m.storage_info()['layout']} push {best:.1f} us "
          f"{alg / best / 1e3:.0f} GB/s frac {alg / best / 1e3 / 6547.5:.3f}", flush=True)
    del m, c, v
