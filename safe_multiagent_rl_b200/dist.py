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


class StatsComm:
    """The C-ABI communicator (include/smarl.h: smarl_comm_* / smarl_stats_allreduce): NCCL bound by libsmarl
    itself, so the whole closed-loop batch -- steps, accounting, the all-reduce, the lambda update -- is one
    capturable sequence of C-ABI calls and a host without PyTorch can shard the same way.  The 128-byte unique
    id travels over whatever the caller has; here that is torch.distributed's existing process group."""

    def __init__(self, rank: int, world: int, unique_id: bytes):
        import ctypes as C

        from . import _lib
        self.lib, self.rank, self.world = _lib.load(), int(rank), int(world)
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(self.lib.smarl_comm_init_from_unique_id(C.byref(self._h), buf, self.rank, self.world))

    @staticmethod
    def new_unique_id() -> bytes:
        import ctypes as C

        from . import _lib
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().smarl_comm_get_unique_id(buf))
        return buf.raw

    @classmethod
    def from_process_group(cls, group=None):
        """Collective over ``group``: rank 0 creates the id, everyone joins.  Single-process runs get a
        one-rank communicator (the all-reduce is then the identity, through the same code path)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            box = [cls.new_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            return cls(rank, world, box[0])
        return cls(0, 1, cls.new_unique_id())

    def allreduce(self, vec: torch.Tensor) -> torch.Tensor:
        from . import _lib
        assert vec.is_cuda and vec.dtype == torch.float64 and vec.is_contiguous()
        _lib.check(self.lib.smarl_stats_allreduce(self._h, vec.data_ptr(), vec.numel(), _lib.stream_ptr()))
        return vec

    def close(self):
        if self._h:
            self.lib.smarl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_stats(vec: torch.Tensor, group=None, comm: "StatsComm | None" = None) -> torch.Tensor:
    """Sum the additive stats vector over ranks (in place).  Every slot is a sum over envs
    (integer-valued cost sums / counts are exact in f64), so the result equals the
    single-GPU vector over all envs.  With ``comm`` the reduction goes through the C ABI
    (smarl_stats_allreduce), otherwise through torch.distributed."""
    if comm is not None:
        return comm.allreduce(vec)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec
