"""N>1 host logic on CPU: world_size-2 gloo process group (env sharding, per-rank keys, max-over-ranks timing)."""
import os
import sys
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401
from mujoco_mjx_lab_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    parallel.init("gloo")
    lo, hi = parallel.shard_range(n_total, rank, world)
    keys = parallel.rank_keys(42, rank, hi - lo)
    t = parallel.max_over_ranks(10.0 + rank, torch.device("cpu"))
    s = parallel.sum_over_ranks(float(hi - lo), torch.device("cpu"))
    parallel.barrier()
    out[rank] = (lo, hi, keys, t, s)
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, n_total = 2, 1001
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    (lo0, hi0, k0, t0, s0), (lo1, hi1, k1, t1, s1) = out[0], out[1]
    assert lo0 == 0 and hi0 == lo1 and hi1 == n_total and abs((hi0 - lo0) - (hi1 - lo1)) <= 1      # disjoint cover
    assert t0 == t1 == 11.0 and s0 == s1 == n_total                                                  # max / sum over ranks
    assert k0.shape == (hi0 - lo0, 2) and not np.array_equal(k0[: len(k1)], k1[: len(k0)])           # ranks draw different resets
    assert len(np.unique(np.concatenate([k0, k1]).view(np.uint64))) == n_total


def test_shard_range_properties():
    for n in (1, 7, 64, 1000, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.gpu
def test_peer_memory_allreduce_adam_two_gpus():
    """mjxb_allreduce_adam on 2 GPUs (one process each): the gradient sum over NVLink peer memory fused into the Adam kernel gives the
    parameters of NCCL all-reduce + torch.optim.Adam, identically on both ranks. Skipped on a box with one GPU."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_comm_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer comm ok" in r.stdout
