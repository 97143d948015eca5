"""Env-sharded data parallelism helpers: environments are independent units, so ranks own contiguous env
ranges and the physics path needs no collective.  torch.distributed is used only for the barrier / max-over-ranks timing
here, and for the learner's gradient all-reduce in the PPO driver."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import jax_random


def dist_env() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str) -> Tuple[int, int, int]:
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa(local_rank: int) -> int:
    """One process per GPU: restrict this process to the CPUs that are local to its GPU (NVML's ideal CPU affinity), so that pinned host
    buffers (first touch) and the launching thread sit on the GPU's NUMA node. A launcher such as torchrun does not bind its workers;
    with eight ranks the host-buffer path otherwise crosses the socket interconnect for part of the ranks. Call before the first CUDA /
    pinned allocation. Returns the number of CPUs bound to (0: left as it was -- NVML missing, or the mask is empty / not allowed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by `rank` (strong scaling: fixed total)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_keys(seed: int, rank: int, n_local: int, step: int = 0) -> np.ndarray:
    """Per-rank reset keys: fold the rank and the step into the seed key, then split per env (train_ppo.py:150-151 per shard)."""
    k = jax_random.PRNGKey(seed)
    k = jax_random.split(k, rank + 2)[rank + 1]
    k = jax_random.split(k, step + 2)[step + 1]
    return jax_random.split(k, n_local)


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
