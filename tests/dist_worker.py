"""Worker for the multi-rank tests (launched by torch.distributed.run, one rank per process).

  backend nccl (GPU): row-partitioned sb200_dist_solve vs the single-process CPU oracle on the same system (both modes,
                      identity residual, bare recurrence, initial guess, a PageRank system under ROW_OR_COL dominance,
                      error agreement across ranks).
  backend nccl_big  : the same comparison at a size whose row blocks take the column-slab layout (n >= 6.5 M); the
                      oracle runs on rank 0 only and its solution is broadcast.
  backend gloo (CPU): the exchange protocol of csrc/dist.cu restated with gloo collectives around the oracle's
                      row-block SpMV — partition boundaries, padded in-place allgather layout, all-reduced norms
                      driving identical loop decisions on every rank.
Prints "RANK r OK" on success.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sublinear-time-solver_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import sublinear_b200 as sb  # noqa: E402
from oracle import oracle as O  # noqa: E402


def gloo_protocol(rank, world, n, sparsity, mode):
    """dist.cu's per-term protocol with gloo: local push on the rank's rows, allreduce(norm^2), in-place allgather of
    the padded term slice, residual every 5th iteration from an allgathered solution."""
    r0, r1 = sb.partition_rows(n, world, rank)
    per = -(-n // world)
    assert r0 == min(n, rank * per) and r1 == min(n, (rank + 1) * per)
    rp, ci, v, b = sb.gen_bench_csr(n, sparsity, r0, r1)            # the library's host-side generator, local rows
    A = O.Csr(r1 - r0, n, v, ci, rp.astype(np.uint32))
    nloc = r1 - r0
    dinv = np.array([1.0 / A.values[A.row_ptr[i]:A.row_ptr[i + 1]][A.col_indices[A.row_ptr[i]:A.row_ptr[i + 1]] == r0 + i].sum()
                     for i in range(nloc)])
    c = b * dinv
    compat = mode == sb.MODE_REF_COMPAT
    tol, stol, max_terms, max_it = 1e-6, 1e-8, 50, 1000

    def allgather(local):
        full = torch.zeros(per * world, dtype=torch.float64)
        full[rank * per: rank * per + nloc] = torch.from_numpy(local)
        td.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per].clone())   # equal padded counts
        return full.numpy()[:n]

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64)
        td.all_reduce(t)
        return float(t.item())

    def residual(x):
        xf = allgather(x)
        r = A.multiply_vector(xf) - (c if compat else b)
        return np.sqrt(allsum(float(np.dot(r, r))))

    t = c.copy()
    x = (c + c) if compat else c.copy()
    terms, it, sconv, res = 1, 0, False, np.inf
    if np.sqrt(allsum(float(np.dot(t, t)))) < stol:
        sconv = True
    res = residual(x)
    it = 1
    matvec = 1
    while not sconv and not (res <= tol) and it < max_it and terms < max_terms:
        tf = allgather(t)
        t = t - dinv * A.multiply_vector(tf)
        matvec += 1
        x = x + t
        terms += 1
        if np.sqrt(allsum(float(np.dot(t, t)))) < stol:
            sconv = True
        if it % 5 == 0:
            res = residual(x)
            matvec += 1
        it += 1
    res = residual(x)
    matvec += 1
    return r0, r1, x, it, terms, matvec, res


def main():
    backend = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    sparsity = float(sys.argv[3]) if len(sys.argv) > 3 else 5e-4
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if backend != "nccl_big" or rank == 0:
        Afull, bfull = O.gen_bench_csr(n, sparsity)
    if backend == "gloo":
        td.init_process_group("gloo")
        for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
            o = O.neumann_solve(Afull, bfull, mode=mode)
            r0, r1, x, it, terms, matvec, res = gloo_protocol(rank, world, n, sparsity, mode)
            assert (it, terms, matvec) == (o.iterations, o.terms_computed, o.matvec_count), (it, terms, matvec, o)
            assert np.array_equal(x, o.solution[r0:r1])      # row sums do not depend on the partition
            assert abs(res - o.residual_norm) <= 1e-9 * max(o.residual_norm, 1e-30)
    elif backend == "nccl_big":
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [sb.Comm.unique_id() if rank == 0 else None]
        td.broadcast_object_list(uid, src=0)
        comm = sb.Comm(rank, world, uid[0], local_rank)
        r0, r1 = sb.partition_rows(n, world, rank)
        rp, ci, v, b = sb.gen_bench_csr(n, sparsity, r0, r1)
        m = comm.matrix_from_csr(n, r0, r1, rp, ci, v)
        want_slabs = int(os.environ.get("EXPECT_SLABS", "1"))
        if want_slabs:
            assert m.storage_info()["layout"] == sb.LAYOUT_CSR_SLABS, m.storage_info()
        for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
            r = comm.solve(sb.NeumannSolver.default(), m, b, sb.SolverOptions(mode=mode))
            ref = torch.empty(n + 4, dtype=torch.float64, device="cuda")
            if rank == 0:
                o = O.neumann_solve(Afull, bfull, mode=mode, spmv_variant=O.SPMV_PARALLEL)
                ref[:n] = torch.from_numpy(o.solution)
                ref[n:] = torch.tensor([o.iterations, o.terms_computed, o.matvec_count, o.residual_norm], dtype=torch.float64)
            td.broadcast(ref, src=0)
            ref = ref.cpu().numpy()
            assert (r.iterations, r.terms_computed, r.matvec_count) == tuple(int(t) for t in ref[n:n + 3]), (rank, mode, r)
            assert np.array_equal(r.solution, ref[r0:r1]), "distributed solution differs from the oracle"
            assert abs(r.residual_norm - ref[n + 3]) <= 1e-9 * max(ref[n + 3], 1e-30)
    else:
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [sb.Comm.unique_id() if rank == 0 else None]
        td.broadcast_object_list(uid, src=0)
        comm = sb.Comm(rank, world, uid[0], local_rank)
        r0, r1 = sb.partition_rows(n, world, rank)
        rp, ci, v, b = sb.gen_bench_csr(n, sparsity, r0, r1)
        m = comm.matrix_from_csr(n, r0, r1, rp, ci, v)
        for mode in (sb.MODE_CORRECT, sb.MODE_REF_COMPAT):
            o = O.neumann_solve(Afull, bfull, mode=mode)
            r = comm.solve(sb.NeumannSolver.default(), m, b, sb.SolverOptions(mode=mode))
            assert (r.iterations, r.terms_computed, r.matvec_count, r.converged) == \
                   (o.iterations, o.terms_computed, o.matvec_count, o.converged), (rank, mode, r, o)
            assert np.array_equal(r.solution, o.solution[r0:r1]), "distributed solution differs from the oracle"
            assert abs(r.residual_norm - o.residual_norm) <= 1e-9 * max(o.residual_norm, 1e-30)
            # initial guess (local slice): ref_compat adds the series onto it, correct restarts from D^-1 (b - A x0)
            x0 = np.random.default_rng(11).standard_normal(n)
            o = O.neumann_solve(Afull, bfull, mode=mode, initial_guess=x0)
            r = comm.solve(sb.NeumannSolver.default(), m, b, sb.SolverOptions(mode=mode, initial_guess=x0[r0:r1]))
            assert (r.iterations, r.terms_computed, r.matvec_count, r.converged) == \
                   (o.iterations, o.terms_computed, o.matvec_count, o.converged), (rank, mode, r, o)
            assert np.array_equal(r.solution, o.solution[r0:r1]), "initial-guess solution differs from the oracle"
        # a wrong-length guess on ONE rank is rejected on EVERY rank (agreed before any collective of the solve)
        try:
            bad = x0[r0:r1 - 1] if rank == 0 else x0[r0:r1]
            comm.solve(sb.NeumannSolver.default(), m, b, sb.SolverOptions(initial_guess=bad))
            raise AssertionError("expected an error")
        except sb.SolverError as e:
            assert e.variant in ("DimensionMismatch", "InvalidInput"), e
        # identity-residual mode: no solution allgathers inside the loop
        r = comm.solve(sb.NeumannSolver.default(), m, b, sb.SolverOptions(residual_check=sb.RESIDUAL_IDENTITY))
        x, _, _, _ = O.push_iterations(Afull, bfull, r.terms_computed - 1)
        assert np.array_equal(r.solution, x[r0:r1]) and r.converged
        # bare recurrence used by bench.py
        bd = torch.tensor(b, device="cuda")
        xd = torch.empty_like(bd)
        norms, ms = comm.push_iterations_dev(m, bd.data_ptr(), r1 - r0, 6, xd.data_ptr())
        x6, _, on, _ = O.push_iterations(Afull, bfull, 6)
        assert np.array_equal(xd.cpu().numpy(), x6[r0:r1])
        np.testing.assert_allclose(norms, on, rtol=1e-12)
        # PageRank system (config C3's shape, scaled): S = I - alpha P^T is column- but not row-dominant, so the solve
        # needs SB200_DOMINANCE_ROW_OR_COL — whole columns = an all-reduce of the per-column sums of the row blocks
        rng = np.random.default_rng(5)
        npr, ne = 30_000, 300_000
        src = rng.integers(0, npr, ne)
        dst = np.minimum((rng.pareto(1.1, ne) * 20).astype(np.int64), npr - 1)
        S, rhs = O.pagerank_system(src, dst, npr, 0.85)
        p0, p1 = sb.partition_rows(npr, world, rank)
        srp = (S.row_ptr[p0:p1 + 1] - S.row_ptr[p0]).astype(np.uint64)
        sl = slice(int(S.row_ptr[p0]), int(S.row_ptr[p1]))
        mp = comm.matrix_from_csr(npr, p0, p1, srp, S.col_indices[sl], S.values[sl])
        solver = sb.NeumannSolver.new(200, 1e-9)
        try:
            comm.solve(solver, mp, rhs[p0:p1])                       # row dominance alone: rejected on every rank
            raise AssertionError("expected MatrixNotDiagonallyDominant")
        except sb.SolverError as e:
            assert e.variant == "MatrixNotDiagonallyDominant", e
        r = comm.solve(solver, mp, rhs[p0:p1], sb.SolverOptions(dominance=sb.DOMINANCE_ROW_OR_COL))
        o = O.neumann_solve(S, rhs, max_terms=200, series_tolerance=1e-9, dominance=O.DOM_ROW_OR_COL)
        assert (r.iterations, r.terms_computed, r.matvec_count, r.converged) == \
               (o.iterations, o.terms_computed, o.matvec_count, o.converged), (rank, r, o)
        np.testing.assert_allclose(r.solution, o.solution[p0:p1], rtol=1e-12, atol=1e-18)   # hub rows: lane-strided sums
        # a row block that is not diagonally dominant on ONE rank must fail on EVERY rank (no hang)
        if world > 1:
            v2 = v.copy()
            if rank == world - 1:
                v2[rp[0]:rp[1]] *= 1e-3
                v2[rp[0]:rp[1]][ci[rp[0]:rp[1]] != r0] = 1e3
            m2 = comm.matrix_from_csr(n, r0, r1, rp, ci, v2)
            try:
                comm.solve(sb.NeumannSolver.default(), m2, b)
                raise AssertionError("expected MatrixNotDiagonallyDominant")
            except sb.SolverError as e:
                assert e.variant == "MatrixNotDiagonallyDominant", e
    td.barrier()
    print(f"RANK {rank} OK", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:          # a failed rank must take the job down (torchrun kills the others) instead of
        import traceback           # leaving them blocked in a collective until the test's timeout
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
