"""Multi-rank coverage of the row-partitioned path: world_size-2 gloo on CPU (host logic + exchange protocol) and
NCCL on GPUs when at least two are visible."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, backend, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(HERE, "dist_worker.py"),
           backend, *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        assert f"RANK {rank} OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,n", [(2, 3001), (3, 1000)])
def test_row_partition_protocol_gloo(world, n):
    launch(world, "gloo", n, 0.004)


@pytest.mark.gpu
def test_dist_solve_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    launch(2, "nccl", 20001, 5e-4)
