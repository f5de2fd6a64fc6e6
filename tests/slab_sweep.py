"""Diagnostic (single GPU): the fused column-slab kernel against the single-pass layouts on the headline system and on
the row blocks the multi-GPU path builds (python tests/slab_sweep.py [n] [what ...]; what in {full, blocks}).

  full   : bare push recurrence (sb200_push_iterations_dev) on gen_bench(n, 10/n), layouts SLABS = 0 (SELL-32), 2, 3, 4
  blocks : multiply_vector_dev (same gathers and stream, lighter epilogue) on the last row block of world = 2, 4, 8,
           default rule vs forced single pass vs forced slabs
The layout is chosen at ingest from $SUBLINEAR_B200_SLABS; CTAs per SM of the slab kernel from $SUBLINEAR_B200_SLAB_CTAS
(read once per process)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sublinear_b200 as sb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
what = sys.argv[2:] or ["full", "blocks"]
sb.set_device(0)
PEAK = 6547.5
tag = {k: v for k, v in os.environ.items() if k.startswith("SUBLINEAR_B200")}


def with_env(slabs):
    if slabs is None:
        os.environ.pop("SUBLINEAR_B200_SLABS", None)
    else:
        os.environ["SUBLINEAR_B200_SLABS"] = str(slabs)


if "full" in what:
    rp, ci, v, b = sb.gen_bench_csr(n, 10.0 / n)
    bd = torch.tensor(b, device="cuda")
    alg = 12 * len(v) + 44 * n + 4
    for slabs in (None, 0, 2, 3, 4):
        with_env(slabs)
        t0 = time.perf_counter()
        m = sb.SparseMatrix.from_csr(rp, ci, v, n, n)
        t_ingest = time.perf_counter() - t0
        best = 1e9
        for _ in range(3):
            norms, ms = sb.push_iterations_dev(m, bd.data_ptr(), n, 12)
            best = min(best, ms / 12 * 1e3)
        print(f"full n={n} nnz={len(v)} SLABS={slabs} layout={m.storage_info()['layout']} {tag}: push {best:.1f} us  "
              f"{alg / best / 1e3:.0f} GB/s  frac {alg / best / 1e3 / PEAK:.3f}  ingest {t_ingest:.2f} s", flush=True)
        del m
    del rp, ci, v

if "blocks" in what:
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    for world in (2, 4, 8):
        r0, r1 = sb.partition_rows(n, world, world - 1)
        rp, ci, v, b = sb.gen_bench_csr(n, 10.0 / n, r0, r1)
        nl = r1 - r0
        y = torch.empty(nl, dtype=torch.float64, device="cuda")
        alg = 12 * len(v) + 12 * nl + 8 * n
        for slabs in (None, 0, 2, 3, 4):
            with_env(slabs)
            m = sb.SparseMatrix.from_csr(rp, ci, v, nl, n)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            best = 1e9
            for rep in range(4):
                torch.cuda.synchronize()
                ev[0].record()
                for _ in range(10):
                    m.multiply_vector_dev(x.data_ptr(), n, y.data_ptr(), nl, False, torch.cuda.current_stream().cuda_stream)
                ev[1].record()
                torch.cuda.synchronize()
                if rep:
                    best = min(best, ev[0].elapsed_time(ev[1]) / 10 * 1e3)
            print(f"block world={world} rows={nl} nnz={len(v)} SLABS={slabs} layout={m.storage_info()['layout']} {tag}: "
                  f"spmv {best:.1f} us  {alg / best / 1e3:.0f} GB/s", flush=True)
            del m
