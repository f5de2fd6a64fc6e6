#!/usr/bin/env python
"""bench.py — push-iteration throughput (processed nnz/s) of the Neumann / forward-push path.

Workload (BASELINE.json `metric`): gen_bench(10 000 000, 1e-6) = the reference's own Criterion generator
(benches/performance_benchmarks.rs:12-43): n = 10 M, nnz ~ 100 M, uniform-random columns, b_i = 1 + 0.001 i.
A "step" is one full NeumannSolver::solve (default solver 50 terms / 1e-8, default options tolerance 1e-6,
MODE_CORRECT, the reference's residual-every-5th-iteration cadence).
  value : nnz x SpMV-equivalents (matvec_count) per second, inputs resident in HBM (sb200_solve_dev), CUDA-event timed
  e2e   : the same through the host-pointer C-ABI call a drop-in user makes (sb200_solve_into): pinned host b in,
          pinned host x out, H2D + D2H inside the timed region
  roofline : fused push kernel, algorithmic bytes 12 nnz + 44 n + 4 per launch / average launch time (CUDA events
          around every push launch inside the timed steps) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the oracle's restatement of the reference's rayon row-chunk SpMV (src/simd_ops.rs:202-239) inside
          the same recurrence, all host cores, bounded sample
  config.parity : the solution of the last timed step (gathered on rank 0 for N > 1) against the oracle's full solve on
          the same system — bit_exact / max_abs_err / counts_equal — and ||Ax-b||/||b|| from the ORACLE's SpMV
N > 1 (torchrun, one rank per GPU): STRONG scaling on the same 10 M system — contiguous row blocks, one exchange of
the term slice + two partial sums per term (sb200_dist_solve: fused peer-memory stores by default,
SUBLINEAR_B200_DIST=nccl for the allgather + allreduce flavour).
`--impl reference`: the reference's CPU path (oracle port; no Rust toolchain in this image) on the host cores, timing the
same step as the GPU arm (one full default solve, row-chunk parallel SpMV on all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sublinear-time-solver_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (size, sparsity)  — gen_bench(size, sparsity), SURVEY.md §8(d)
    "c3_n10M_nnz100M": (10_000_000, 1e-6),
    "c2_n1M_nnz10M": (1_000_000, 1e-5),
    "c5_n100M_nnz1B": (100_000_000, 1e-7),
    "tiny_n100k": (100_000, 1e-4),
    # best-case gather locality (SURVEY.md §8d asks for it next to the uniform-random headline): same row length and
    # values, columns within +-64 of the diagonal; single GPU only, generated with numpy
    "banded_n10M_nnz100M": (10_000_000, 1e-6),
}
METRIC = "push-iter nnz/sec at n=10M nnz=100M"
UNIT = "nnz/s"


def algorithmic_bytes_push(n, nnz):
    return 12 * nnz + 44 * n + 4        # SURVEY.md §8(d), BASELINE.md §3


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


class CpuPath:
    """The reference's CPU path restated (oracle/, rebuilt -O3 -march=native on THIS host): the rayon-style
    row-chunk parallel SpMV (src/simd_ops.rs:202-239) inside the push recurrence, all host cores."""

    def __init__(self, size, sparsity):
        import ctypes as C
        import tempfile
        import numpy as np
        from oracle import oracle as O
        self.C, self.O = C, O
        self.L = O.lib(fast=True, out_dir=tempfile.mkdtemp(prefix="sb200_oracle_"))
        self.raw = O._Csr()
        self.b = np.zeros(size)
        rc = self.L.orc_gen_bench_csr(size, sparsity, 0, size, C.byref(self.raw),
                                      self.b.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        self.nnz = int(self.raw.nnz)
        self.cores = os.cpu_count() or 1

    def run(self, iters):
        C = self.C
        secs = self.L.orc_push_iterations(C.byref(self.raw), self.b.ctypes.data_as(C.POINTER(C.c_double)), iters,
                                          self.O.SPMV_PARALLEL, self.cores, None, None, None)
        return self.nnz * iters / secs, secs

    def solve(self, mode=0):
        """one full NeumannSolver::default().solve (the step bench.py's GPU arm times) with the row-chunk parallel SpMV"""
        import numpy as np
        C, O = self.C, self.O
        o = O._Options()
        self.L.orc_options_default(C.byref(o))
        o.mode, o.spmv_variant, o.nthreads = mode, O.SPMV_PARALLEL, self.cores
        x = np.zeros(len(self.b))
        r = O._Result()
        r.solution = x.ctypes.data_as(C.POINTER(C.c_double))
        t0 = time.perf_counter()
        rc = self.L.orc_neumann_solve(C.byref(self.raw), self.b.ctypes.data_as(C.POINTER(C.c_double)), len(self.b),
                                      C.byref(o), C.byref(r))
        secs = time.perf_counter() - t0
        assert rc == 0, rc
        return int(r.matvec_count), int(r.terms_computed), secs

    def close(self):
        self.L.orc_csr_free(self.C.byref(self.raw))


def cpu_sample_shape(size, sparsity):
    # bounded: the CPU arm uses the workload itself up to 10M rows, a 10M-row system of the same generator above
    if size > 10_000_000:
        return 10_000_000, sparsity * size / 10_000_000
    return size, sparsity


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores. The reference is
    Rust (no toolchain here, so no oracle/_ref): the arm is the oracle port (kind "port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size, sparsity = WORKLOADS[args.workload]
    s_size, s_sparsity = cpu_sample_shape(size, sparsity)
    cpu = CpuPath(s_size, s_sparsity)
    # the same step as the GPU arm: one full default solve (term recurrence + residual every 5th iteration + final
    # residual), counted in SpMV-equivalents; warm-up bounded to one solve (each is ~1.5 s of all-core CPU work)
    mode = 0 if args.mode == "correct" else 1
    vals = []
    mv = terms = 0
    for i in range(min(args.warmup, 1) + args.steps):
        mv, terms, secs = cpu.solve(mode)
        if i >= min(args.warmup, 1):
            vals.append(secs)
    cpu.close()
    sample = (f"gen_bench({s_size}, {s_sparsity:g}) nnz={cpu.nnz}, one NeumannSolver::default().solve per step "
              f"({terms} terms, {mv} SpMV-equivalents: SpMV + diagonal scale + term/solution update + norms), OpenMP row "
              f"chunks on {cpu.cores} threads")
    total_s = sum(vals)
    value = cpu.nnz * mv * len(vals) / total_s
    ms = 1e3 * total_s / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": s_size, "nnz": cpu.nnz,
                       "generator": "gen_bench(size, sparsity) = benches/performance_benchmarks.rs:12-43, uniform-random columns",
                       "step": "one NeumannSolver::default().solve (50 terms / 1e-8, tolerance 1e-6, residual every 5th iteration)",
                       "mode": args.mode, "terms_per_step": terms, "matvecs_per_step": mv, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def gen_banded(n, k, w, seed=1):
    """Banded companion of gen_bench: a_ii = 10 + 0.01 i, k-1 positive off-diagonals of the same magnitude law
    (< a_ii / (2k)) at columns i + d, 1 <= |d| <= w (reflected at the borders); rows sorted by column."""
    import numpy as np
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.int64)[:, None]
    d = rng.integers(1, w + 1, size=(n, k - 1)) * rng.choice(np.array([-1, 1]), size=(n, k - 1))
    c = i + d
    c = np.where((c < 0) | (c >= n), i - d, c)          # fold back inside, never onto the diagonal
    diag = 10.0 + 0.01 * np.arange(n)
    vals = rng.random((n, k - 1)) * (diag / (2.0 * k))[:, None]
    cols = np.concatenate([i, c], axis=1)
    vals = np.concatenate([diag[:, None], vals], axis=1)
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1).astype(np.uint32).ravel()
    vals = np.take_along_axis(vals, order, axis=1).ravel()
    rp = np.arange(n + 1, dtype=np.uint64) * k
    return rp, cols, vals, 1.0 + 0.001 * np.arange(n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_n10M_nnz100M", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-iters", type=int, default=5)
    ap.add_argument("--cpu-baseline-iters", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed solution")
    ap.add_argument("--mode", default="correct", choices=["correct", "ref_compat"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import sublinear_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout must carry ONE JSON line, but NCCL prints its version banner (NCCL_DEBUG=VERSION) with a C-level write to
    # fd 1: keep a private handle on the real stdout for the result and point fd 1 at stderr for everything else
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    sb.set_device(local_rank)
    dist = world > 1
    if dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    size, sparsity = WORKLOADS[args.workload]
    mode = sb.MODE_CORRECT if args.mode == "correct" else sb.MODE_REF_COMPAT
    opts = sb.SolverOptions(mode=mode, enable_profiling=True, collect_stats=True)
    solver = sb.NeumannSolver.default()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist:
            td.barrier()
            torch.cuda.synchronize()

    t_gen = time.perf_counter()
    if args.workload.startswith("banded"):
        if dist:
            raise SystemExit("the banded workload is single-GPU only")
        rp, ci, v, b = gen_banded(size, 10, 64)
        nnz_total, n_local = len(v), size
        m = sb.SparseMatrix.from_csr(rp, ci, v, size, size)
        comm = None
    elif not dist:
        rp, ci, v, b = sb.gen_bench_csr(size, sparsity)
        nnz_total, n_local = len(v), size
        m = sb.SparseMatrix.from_csr(rp, ci, v, size, size)
        comm = None
    else:
        uid = [sb.Comm.unique_id() if rank == 0 else None]
        td.broadcast_object_list(uid, src=0)
        comm = sb.Comm(rank, world, uid[0], local_rank)
        r0, r1 = sb.partition_rows(size, world, rank)
        rp, ci, v, b = sb.gen_bench_csr(size, sparsity, r0, r1)      # rows are independently seeded: each rank builds its own
        n_local = r1 - r0
        m = comm.matrix_from_csr(size, r0, r1, rp, ci, v)
        t = torch.tensor([len(v)], dtype=torch.int64, device="cuda")
        td.all_reduce(t)
        nnz_total = int(t.item())
    del rp, ci, v
    t_gen = time.perf_counter() - t_gen

    # inputs: device-resident b for `value`, pinned host b / x for `e2e`
    b_dev = torch.tensor(b, device="cuda")
    x_dev = torch.empty_like(b_dev)
    b_pin = torch.from_numpy(b).pin_memory()
    x_pin = torch.empty(n_local, dtype=torch.float64).pin_memory()

    def step_dev():
        if not dist:
            return solver.solve_dev(m, b_dev.data_ptr(), n_local, x_dev.data_ptr(), opts, stream)
        return _dist_step(b_dev.data_ptr(), x_dev.data_ptr())

    def _dist_step(bp, xp):
        import ctypes as C
        o = opts._c()
        r = sb._Result()
        rc = sb.lib().sb200_dist_solve(comm._h, solver._h, m._h, bp, n_local, C.byref(o), xp, C.byref(r))
        res = sb.SolverResult._from(r, None)
        sb._check(rc, res)
        return res

    def step_e2e():
        if not dist:
            return solver.solve(m, b_pin.numpy(), opts, out=x_pin.numpy())
        return _dist_step(b_pin.data_ptr(), x_pin.data_ptr())

    # ---- value: device-resident, CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # nvidia-smi needs ~0.5 s to deliver its first sample: started before the warm-up steps,
        time.sleep(0.6)      # which put the same load on the GPU as the timed ones
    for _ in range(args.warmup):
        step_dev()
    if rank == 0:
        sampler.rows.clear() # keep the samples taken under load only
    for _ in range(args.warmup):
        step_dev()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    results = []
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        results.append(step_dev())
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    wall = (t_wall1 - t_wall0) * 1e3
    if dist:
        # the row-partitioned path runs on the communicator's own stream (torch events on the current stream do not
        # see it); its calls are synchronous, so the barrier-to-barrier wall clock, max over ranks, is the step time
        ms_total = wall
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_total = float(t.item())
    r_last = results[-1]
    matvecs = sum(r.matvec_count for r in results)
    value = nnz_total * matvecs / (ms_total * 1e-3)
    launches = sum(r.kernel_launches for r in results)

    # ---- roofline of the dominant kernel (fused push), live from the timed region ----
    push_ms = sum(r.push_kernel_ms for r in results)
    push_cnt = sum(r.push_kernel_count for r in results)
    res_ms = sum(r.resid_kernel_ms for r in results)
    res_cnt = sum(r.resid_kernel_count for r in results)
    nnz_local = m.nnz()
    peak, peak_src = measured_peak()
    roofline = None
    layout = m.storage_info()
    kernel_name = {sb.LAYOUT_SELL32: "sell_kernel<EPI_PUSH,256,8> (SELL-32 layout, no shared memory)",
                   sb.LAYOUT_CSR_SLABS: "slab_kernel<EPI_PUSH,256,3>: one fused launch walks the column slabs of the term "
                                        "vector (CSR order kept, row sums carried from slab to slab)",
                   sb.LAYOUT_CSR: "warp_kernel<EPI_PUSH,256,4> (CSR slices)"}[layout["layout"]]
    if push_cnt:
        t_push = push_ms / push_cnt * 1e-3
        alg = algorithmic_bytes_push(n_local, nnz_local) + (8 * (size - n_local) if dist else 0)
        ach = alg / t_push / 1e9
        # DRAM bytes per launch come from an ncu capture (profiles/scripts/r2_ncu_push.sh -> profiles/push_traffic.json);
        # the entry is used only if it was captured on the kernel that ran here, else null
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "push_traffic.json"))).get(args.workload)
            if tr and not dist and tr.get("kernel", "").split("<")[0] == kernel_name.split("<")[0]:
                traffic, traffic_src = tr["dram_bytes_per_launch"], tr.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "avg_launch_us": t_push * 1e6,
                    "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                    "frac_of_nominal_8TBs": ach / 8000.0, "push_share_of_step": push_ms / (ms_total if not dist else wall),
                    "resid_kernel_avg_us": (res_ms / res_cnt * 1e3) if res_cnt else None,
                    "nnz_per_s_push_kernel": nnz_local / t_push}

    # ---- e2e: host buffers through the C-ABI solve call, copies inside the timed region ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_results = [step_e2e() for _ in range(args.steps)]
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = nnz_total * sum(r.matvec_count for r in e2e_results) / e2e_s

    # the same call with PAGEABLE host buffers (what a Rust Vec<f64> caller hands over): staged through the library's
    # pinned ring
    e2e_pageable = None
    if not dist:
        b_pg, x_pg = np.array(b, copy=True), np.empty(n_local)
        solver.solve(m, b_pg, opts, out=x_pg)
        barrier()
        t0 = time.perf_counter()
        pg_results = [solver.solve(m, b_pg, opts, out=x_pg) for _ in range(args.steps)]
        barrier()
        e2e_pageable = nnz_total * sum(r.matvec_count for r in pg_results) / (time.perf_counter() - t0)

    # ---- parity of what was timed, against the CPU oracle (the checker, never the timed path): the solution of the last
    # timed step vs orc_neumann_solve on the same system (bit-exact expected), and ||Ax-b||/||b|| from the ORACLE's SpMV.
    # N > 1: the rank slices are gathered on rank 0 first.
    parity = None
    if not args.no_parity:
        x_full = x_dev
        if dist:
            per = -(-size // world)
            pad = torch.zeros(per, dtype=torch.float64, device="cuda")
            pad[:n_local] = x_dev
            parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            td.gather(pad, parts, dst=0)
            if rank == 0:
                x_full = torch.cat(parts)[:size]
        if rank == 0:
            from oracle import oracle as O
            t0 = time.perf_counter()
            xs = x_full.cpu().numpy()
            if args.workload.startswith("banded"):
                rp_, ci_, v_, bo = gen_banded(size, 10, 64)
                Ao = O.Csr(size, size, v_, ci_, rp_.astype(np.uint32))
            else:
                Ao, bo = O.gen_bench_csr(size, sparsity)
            o = O.neumann_solve(Ao, bo, mode=mode, spmv_variant=O.SPMV_PARALLEL)
            rres = float(np.linalg.norm(Ao.multiply_vector(xs, O.SPMV_PARALLEL) - bo) / np.linalg.norm(bo))
            parity = {"oracle": "oracle/sublinear_oracle.c orc_neumann_solve (row-chunk parallel SpMV, scalar order in a row)",
                      "max_abs_err": float(np.max(np.abs(xs - o.solution))),
                      "bit_exact": bool(np.array_equal(xs, o.solution)),
                      "counts_equal": (r_last.iterations, r_last.terms_computed, r_last.matvec_count, bool(r_last.converged)) ==
                                      (o.iterations, o.terms_computed, o.matvec_count, bool(o.converged)),
                      "residual_norm": r_last.residual_norm, "oracle_residual_norm": o.residual_norm,
                      "rel_residual_oracle_spmv": rres, "seconds": time.perf_counter() - t0}
            del Ao
    rel_res = parity["rel_residual_oracle_spmv"] if parity else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s_size, s_sparsity = cpu_sample_shape(size, sparsity)
        cpu = CpuPath(s_size, s_sparsity)
        cpu.run(2)
        v, secs = cpu.run(args.cpu_baseline_iters)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": "port",
                        "sample": f"oracle restatement of parallel_matrix_vector_multiply (src/simd_ops.rs:202-239) in the push "
                                  f"recurrence, gen_bench({s_size}, {s_sparsity:g}) nnz={cpu.nnz}, "
                                  f"{args.cpu_baseline_iters} iterations, {secs:.2f} s"}
        cpu.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": size, "nnz": nnz_total,
                       "generator": ("banded companion of gen_bench: columns within +-64 of the diagonal (best-case locality)"
                                     if args.workload.startswith("banded") else
                                     "gen_bench(size, sparsity) = benches/performance_benchmarks.rs:12-43, uniform-random columns"),
                       "step": "one NeumannSolver::default().solve (50 terms / 1e-8, tolerance 1e-6, residual every 5th iteration)",
                       "mode": args.mode, "terms_per_step": r_last.terms_computed, "matvecs_per_step": r_last.matvec_count,
                       "iterations_per_step": r_last.iterations, "converged": r_last.converged,
                       "l2_policy": "inputs larger than L2 (1.2 GB CSR stream per SpMV vs 126 MB L2), no flush",
                       "parallelism": "single GPU" if not dist else (f"row blocks x{world}, term slice exchanged per term: " +
                                                      ("NCCL allgather + allreduce" if os.environ.get("SUBLINEAR_B200_DIST") == "nccl"
                                                       else "fused peer-memory stores from the push kernel (CUDA IPC over NVLink)")),
                       "rel_residual": rel_res, "parity": parity, "setup_s": t_gen,
                       "device_layout": {sb.LAYOUT_SELL32: "SELL-32", sb.LAYOUT_CSR_SLABS: "CSR column slabs", sb.LAYOUT_CSR: "CSR"}[layout["layout"]],
                       "value_slots_streamed_per_spmv": layout["slots"], "matrix_device_bytes": layout["device_bytes"]},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local * world if dist else 8 * n_local,
                    "d2h_bytes_per_step": 8 * n_local * world if dist else 8 * n_local, "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "sb200_solve_into (pinned host b -> pinned host x)" if not dist else "sb200_dist_solve (host b_local -> host x_local)",
                    "pageable_value": e2e_pageable},
            "gpu_launches": launches, "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
