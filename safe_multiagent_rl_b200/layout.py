"""HBM layout helpers.  Every batched array is agent-major SoA ``[rows, ld]`` (env index
fastest, ``ld`` = n_envs rounded up to 16) so that the kernels' per-row accesses are
contiguous 128-bit vectors; see include/smarl.h "Conventions"."""
from __future__ import annotations

import torch

LD_ALIGN = 16


def pad_ld(n_envs: int) -> int:
    return (int(n_envs) + LD_ALIGN - 1) // LD_ALIGN * LD_ALIGN


def alloc(rows, n_envs, dtype, device, lead=()):
    """Zero-filled ``[*lead, rows, ld]`` tensor."""
    return torch.zeros(*lead, rows, pad_ld(n_envs), dtype=dtype, device=device)


def env_major(t: torch.Tensor, n_envs: int) -> torch.Tensor:
    """``[..., rows, ld]`` storage -> ``[..., n_envs, rows]`` view (no copy)."""
    return t[..., :n_envs].transpose(-1, -2)


def require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("safe_multiagent_rl_b200 runs on CUDA devices only (there is no CPU fallback)")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device available (there is no CPU fallback)")
    if device.index is None:                      # normalise "cuda" -> "cuda:<current>" so tensor.device compares equal
        device = torch.device("cuda", torch.cuda.current_device())
    return device
