"""Single-GPU timing of the CG consumer of the push path's SpMV (diagnostic): python tests/cg_timing.py [n] [k]
Symmetric strictly-dominant system with ~2k+1 entries per row at uniform-random columns; OptimizedSolverConfig default."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sublinear_b200 as sb  # noqa: E402
from test_cg import sym_dd  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sb.set_device(0)
t0 = time.perf_counter()
rows, cols, vals, b = sym_dd(n, k, 1)
m = sb.SparseMatrix.from_triplets(rows, cols, vals, n, n)
nnz = m.nnz()
del rows, cols, vals
print(f"setup {time.perf_counter() - t0:.1f} s, n={n} nnz={nnz} layout={m.storage_info()['layout']}", flush=True)
bd = torch.tensor(b, device="cuda")
xd = torch.empty_like(bd)
s = sb.OptimizedConjugateGradientSolver(sb.OptimizedSolverConfig(1000, 1e-6, enable_profiling=True))
best = None
for _ in range(3):
    r = s.solve_dev(m, bd.data_ptr(), n, xd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if best is None or r.device_time_ms < best.device_time_ms:
        best = r
r = best
spmv_us = r.spmv_kernel_ms / r.spmv_kernel_count * 1e3
per_it_us = r.device_time_ms / r.iterations * 1e3
alg_spmv = 12 * nnz + 4 * (n + 1) + 24 * n          # stream + row_ptr + p gather source + own p + ap write
alg_vec = 72 * n                                     # x r/w, r r/w, p r (phase 1: 48 n) ; r, p r, p w (phase 2: 24 n)
print(f"cg: converged={r.converged} iterations={r.iterations} matvecs={r.matvec_count} residual={r.residual_norm:.3e} "
      f"device {r.device_time_ms:.2f} ms = {per_it_us:.1f} us/iteration; SpMV+dot kernel {spmv_us:.1f} us "
      f"({alg_spmv / spmv_us / 1e3:.0f} GB/s algorithmic), vector passes {per_it_us - spmv_us:.1f} us "
      f"({alg_vec / (per_it_us - spmv_us) / 1e3:.0f} GB/s); nnz/s whole solve {nnz * r.matvec_count / (r.device_time_ms * 1e-3):.3e}",
      flush=True)
