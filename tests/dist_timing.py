"""Per-iteration timing of the row-partitioned bare recurrence (diagnostic; run under torch.distributed.run)."""
import os
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sublinear-time-solver_b200"))
import sublinear_b200 as sb  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
td.init_process_group("nccl", device_id=torch.device("cuda", lr))
uid = [sb.Comm.unique_id() if rank == 0 else None]
td.broadcast_object_list(uid, src=0)
comm = sb.Comm(rank, world, uid[0], lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sp = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-6
r0, r1 = sb.partition_rows(n, world, rank)
rp, ci, v, b = sb.gen_bench_csr(n, sp, r0, r1)
m = comm.matrix_from_csr(n, r0, r1, rp, ci, v)
bd = torch.tensor(b, device="cuda")
for tag in ("warm", "timed"):
    norms, ms = comm.push_iterations_dev(m, bd.data_ptr(), r1 - r0, 20)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    if rank == 0 and tag == "timed":
        print(f"world {world} mode {os.environ.get('SUBLINEAR_B200_DIST', 'p2p')} nostore {os.environ.get('SUBLINEAR_B200_DEBUG_NOSTORE')}: "
              f"{t.item() / 20 * 1e3:.1f} us per push iteration (max over ranks), nnz/s {len(v) * world * 20 / (t.item() * 1e-3):.3e}", flush=True)
td.barrier()
td.destroy_process_group()
