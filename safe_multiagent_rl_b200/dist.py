"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded by contiguous global-id
ranges, no data-path collective.  The only exchange is the all-reduce of the additive stats
vector that feeds the meta-agent's lambda update."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world_size, local_rank); a plain single-process run returns (0, 1, 0)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Best effort: pin this process to the CPU cores NVML reports as local to its GPU, so that
    pinned host staging buffers (first touch) and the copy threads live on the GPU's NUMA node.
    Only matters for the host-buffer (PCIe) path when several ranks share one box."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
    except Exception:
        pass
    return False


def shard_range(n_envs_total: int, rank: int, world: int):
    """Contiguous env-id range [offset, offset + count) owned by ``rank``.  Ranges differ by at
    most one env and concatenate to [0, n_envs_total); env ids key the Philox noise stream, so
    results do not depend on ``world``."""
    base, rem = divmod(int(n_envs_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def allreduce_stats(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the additive stats vector over ranks (in place).  Every slot is a sum over envs
    (integer-valued cost sums / counts are exact in f64), so the result equals the
    single-GPU vector over all envs."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec
