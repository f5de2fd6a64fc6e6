"""Single-GPU timing of the bare push recurrence (diagnostic): python tests/kernel_timing.py [random|banded] [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sublinear_b200 as sb  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "random"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
sb.set_device(0)
if wl == "banded":
    import bench
    rp, ci, v, b = bench.gen_banded(n, 10, 64)
else:
    rp, ci, v, b = sb.gen_bench_csr(n, 10.0 / n)
m = sb.SparseMatrix.from_csr(rp, ci, v, n, n)
bd = torch.tensor(b, device="cuda")
best = 1e9
for _ in range(3):
    norms, ms = sb.push_iterations_dev(m, bd.data_ptr(), n, 12)
    best = min(best, ms / 12 * 1e3)
alg = 12 * len(v) + 44 * n + 4
env = {k: v_ for k, v_ in os.environ.items() if k.startswith("SUBLINEAR_B200")}
print(f"{wl} n={n} nnz={len(v)} layout={m.storage_info()['layout']} {env}: push {best:.1f} us  {alg / best / 1e3:.0f} GB/s  frac {alg / best / 1e3 / 6554.6:.3f}", flush=True)
