"""BatchedMetaAgent -- the Lagrange-multiplier learner of safe_multi_agent_RL/meta_agent.py:4-42
for batches of episodes, with lambda resident on the device.

* ``act``'s penalty <lambda, c> is fused into the env step kernels (pass ``meta.lambdas`` to
  ``env.step`` / ``env.rollout``); ``act`` itself is kept for API compatibility.
* ``step`` consumes the stats vector of a finished batch of episodes (RolloutBuffer.finish /
  env.rollout) instead of per-step constraint lists; it records only while
  ``learning_cycle >= start_learning_cycle`` like the reference (meta_agent.py:19,:25-30).
* ``update`` all-reduces the recorded sums over the process group (one tiny NCCL all-reduce:
  the only inter-GPU traffic on this path) and applies meta_agent.py:32-39 on the device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class BatchedMetaAgent:
    def __init__(self, constraint_space, gamma, lr, thresholds, leq=True, start_learning_cycle=10, decay=1.0,
                 lambda_0=0, n_agents=None, device="cuda", process_group=None, comm=None):
        if not leq:
            raise NotImplementedError("only leq=True constraints (the reference's only use) are built")
        self.constraint_space = constraint_space
        self.K = int(sum(constraint_space))
        self.A = int(n_agents) if n_agents is not None else self.K
        self.gamma = gamma
        self.device = torch.device(device)
        self.thresholds = torch.as_tensor(np.asarray(thresholds, dtype=np.float64)[: self.K]).to(self.device)
        assert self.thresholds.numel() == self.K, "need one threshold per constraint"
        self.leq = leq
        self.lambdas = torch.full((self.K,), float(lambda_0), dtype=torch.float64, device=self.device)
        self.lr = float(lr)
        self.learning_cycle = 0
        self.start_learning_cycle = start_learning_cycle
        self.decay = decay
        self.process_group = process_group
        self.comm = comm           # optional dist.StatsComm: all-reduce through libsmarl's own NCCL binding
        self.lib = _lib.load()
        self._acc = torch.zeros(self.lib.smarl_stats_len(self.A, self.K), dtype=torch.float64, device=self.device)
        self._recorded = False
        self._pending = None       # compat path: per-env cost sums of the episode batch in flight, [E, K]

    def act(self, constraint, reward):
        """meta_agent.py:18-23 on tensors: ``[E, K]``, ``[E, A]`` -> modified reward ``[E, A]``.
        (Compatibility path; the fused path never materialises this.)  Like the reference, the step's
        constraints are recorded while the gate ``learning_cycle >= start_learning_cycle`` is open
        (meta_agent.py:19-20); a following ``step()`` without arguments folds them into the episode sums."""
        c = constraint.to(torch.float64)
        if self.learning_cycle >= self.start_learning_cycle:
            self._pending = c.clone() if self._pending is None else self._pending + c
        pen = c @ self.lambdas
        return reward.to(torch.float64) - pen[:, None]

    def step(self, stats=None):
        """End of an episode batch (meta_agent.py:25-30).  With ``stats`` (the vector of a finished batch from
        RolloutBuffer.finish / env.rollout) it is recorded if the gate is open; without, the constraints that
        ``act`` recorded since the last call become one episode per env (nothing recorded -> no-op, :26-27)."""
        if stats is None:
            if self._pending is not None:
                K, A = self.K, self.A
                self._acc[:K] += self._pending.sum(0)                       # sum_e sum_t c[t, e, k]
                self._acc[K:2 * K] += (self._pending > self.thresholds).to(torch.float64).sum(0)
                self._acc[2 * K + 2 * A] += self._pending.shape[0]          # episodes
                self._pending = None
                self._recorded = True
            return
        if self.learning_cycle >= self.start_learning_cycle:
            vec = stats.vec if hasattr(stats, "vec") else stats
            self._acc += vec
            self._recorded = True

    def global_stats(self):
        """Recorded sums over all ranks (sum_e C_k, violations, return sums, episode count): through the C-ABI
        communicator when one was given (``comm=StatsComm...``), else through torch.distributed."""
        acc = self._acc.clone()
        if self.comm is not None:
            self.comm.allreduce(acc)
        elif dist.is_available() and dist.is_initialized():
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.process_group)
        return acc

    def update(self):
        """meta_agent.py:32-39: lambda <- max(0, lambda + lr * (mean_e C - thr)); lr /= decay."""
        if self._recorded:
            acc = self.global_stats()
            _lib.check(self.lib.smarl_lambda_update(_lib.ptr(self.lambdas), _lib.ptr(acc), _lib.ptr(self.thresholds),
                                                    self.lr, self.A, self.K, _lib.stream_ptr()))
        self.lr = self.lr / self.decay
        self._acc.zero_()
        self._recorded = False
        self._pending = None
        self.learning_cycle = 0

    def increment_learning_cycle(self):
        self.learning_cycle += 1
