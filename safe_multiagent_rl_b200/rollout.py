"""Device rollout buffer and per-episode accounting.

Host-side mirror of safe_multi_agent_RL/buffer.py:7-48 (Buffer.append/step/mean_score) and of
the learners' compute_returns (agent.py:129-132, :200-206) for n_envs episodes at a time.
The *_step kernels write step t's reward / cost / done / penalty straight into slab t of this
buffer (a zero-copy ``Buffer.append``); ``finish`` runs one accounting kernel over it.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch

from . import _lib
from .layout import alloc, env_major, pad_ld

G_NONE, G_REWARD_TO_GO, G_DISCOUNTED_TERMS, G_PPO_STANDARDISED = 0, 1, 2, 3


class Stats:
    """Decoded statistics vector (include/smarl.h smarl_stats_len layout)."""

    def __init__(self, vec: torch.Tensor, n_agents: int, n_constraints: int):
        self.vec, self.A, self.K = vec, n_agents, n_constraints

    @property
    def cost_sum(self):
        return self.vec[: self.K]

    @property
    def violations(self):
        return self.vec[self.K: 2 * self.K]

    @property
    def return_sum(self):
        return self.vec[2 * self.K: 2 * self.K + self.A]

    @property
    def modified_return_sum(self):
        return self.vec[2 * self.K + self.A: 2 * self.K + 2 * self.A]

    @property
    def count(self):
        return self.vec[2 * self.K + 2 * self.A]


_THR_CACHE = {}


def device_thresholds(thresholds, device, n_constraints=None):
    """f64 device tensor for a thresholds list; cached by value so repeated calls (and calls made
    while a CUDA graph is being captured, after one warm-up call) do no host->device copy.
    ``n_constraints``: the kernels read exactly K thresholds, so any other count is rejected (the reference's
    CLI default ``[2]`` with CoverageDiscrete's K = n_agents would otherwise be read out of bounds)."""
    if thresholds is None:
        return None
    n = thresholds.numel() if isinstance(thresholds, torch.Tensor) else np.asarray(thresholds).size
    if n_constraints is not None and n != n_constraints:
        raise ValueError(f"need one threshold per constraint: got {n}, the env has {n_constraints}")
    if isinstance(thresholds, torch.Tensor) and thresholds.is_cuda and thresholds.dtype == torch.float64:
        return thresholds
    key = (str(device), tuple(float(x) for x in np.asarray(
        thresholds.detach().cpu() if isinstance(thresholds, torch.Tensor) else thresholds, dtype=np.float64).ravel()))
    if key not in _THR_CACHE:
        _THR_CACHE[key] = torch.tensor(key[1], dtype=torch.float64, device=device)
    return _THR_CACHE[key]


def make_accounting(gamma, n_steps, g_mode, thresholds_dev):
    return _lib.Accounting(float(gamma), int(n_steps), int(g_mode), _lib.ptr(thresholds_dev))


class RolloutBuffer:
    """[T, rows, ld] device slabs for one batch of n_envs episodes."""

    def __init__(self, n_steps, n_agents, n_constraints, n_envs, cost_dtype=torch.uint8, device="cuda",
                 g_mode=G_REWARD_TO_GO, shared_reward=False, weights=None, store_done=True):
        """shared_reward: store ONE reward row per env and step (the unweighted env reward; ``weights``
        [A] f32 device tensor or None are applied by the accounting kernel) instead of A per-agent rows.
        store_done: keep per-step done flags (envs whose agents never finish do not need them)."""
        self.T, self.A, self.K, self.E = int(n_steps), int(n_agents), int(n_constraints), int(n_envs)
        self.ld = pad_ld(n_envs)
        self.device = torch.device(device)
        self.g_mode = g_mode
        assert cost_dtype in (torch.uint8, torch.int32, torch.float32)
        self.cost_dtype = cost_dtype
        self.cost_code = {torch.uint8: _lib.COST_U8, torch.int32: _lib.COST_I32, torch.float32: _lib.COST_F32}[cost_dtype]
        T, A, K, E = self.T, self.A, self.K, self.E
        self.shared_reward, self.weights = bool(shared_reward), weights
        self.reward = alloc(1 if shared_reward else A, E, torch.float32, device, (T,))
        self.cost = alloc(K, E, cost_dtype, device, (T,))
        self.done = alloc(A, E, torch.uint8, device, (T,)) if store_done else None
        self.penalty = alloc(1, E, torch.float32, device, (T,))[:, 0]
        self.R = alloc(A, E, torch.float32, device)
        self.modR = alloc(A, E, torch.float32, device)
        self.Csum = alloc(K, E, torch.float32 if cost_dtype == torch.float32 else torch.int32, device)
        self.G = alloc(A, E, torch.float32, device, (T,)) if g_mode != G_NONE else None
        lib = _lib.load()
        self.stats_vec = torch.zeros(lib.smarl_stats_len(A, K), dtype=torch.float64, device=device)
        self.stats_scratch = torch.zeros(max(1, lib.smarl_stats_scratch_len(A, K, E)), dtype=torch.float64,
                                         device=device)

    def slot(self, t):
        """Step t's slabs and their reference-oriented views, built once and cached (tensor indexing
        and view construction would otherwise dominate the host cost of a small-batch step)."""
        try:
            return self._slots[t]
        except (AttributeError, KeyError):
            if not hasattr(self, "_slots"):
                self._slots = {}
            r, c, p = self.reward[t], self.cost[t], self.penalty[t]
            d = self.done[t] if self.done is not None else None
            self._slots[t] = (r, c, d, p, env_major(r, self.E), env_major(c, self.E),
                              env_major(d, self.E) if d is not None else None)
            return self._slots[t]

    # ---- views in the reference's orientation ([.., env, agent]) ------------------------------
    def rewards(self):
        """[T, E, A] per-agent rewards (materialised from the env reward and the weights if shared)."""
        if self.shared_reward:
            w = self.weights if self.weights is not None else torch.ones(self.A, device=self.reward.device)
            return self.reward[:, 0, : self.E, None] * w[None, None, :]
        return env_major(self.reward, self.E)

    def modified_rewards(self):
        return self.rewards() - self.penalty[:, : self.E, None]

    def costs(self):
        return env_major(self.cost, self.E)                  # [T, E, K]

    def finish(self, gamma, thresholds=None, g_mode=None, n_active=None):
        """Buffer.step for all envs: returns dict(R [E,A], modR [E,A], C [E,K], G [T,E,A] | None, stats).
        ``n_active`` (i32 [ld] episode lengths, e.g. the Collision env's ``episode_len``) is only needed by
        ``G_PPO_STANDARDISED`` when episodes can end early."""
        lib = _lib.load()
        g_mode = self.g_mode if g_mode is None else g_mode
        if g_mode != G_NONE and self.G is None:
            raise ValueError("buffer was created without G storage")
        thr = device_thresholds(thresholds, self.device, self.K)
        acc = make_accounting(gamma, self.T, g_mode, thr)
        if self.shared_reward:
            _lib.check(lib.smarl_rollout_returns_shared(
                C.byref(acc), _lib.ptr(self.reward), _lib.ptr(self.weights), _lib.ptr(self.cost), self.cost_code,
                _lib.ptr(self.penalty), _lib.ptr(n_active), _lib.ptr(self.R), _lib.ptr(self.modR), _lib.ptr(self.Csum),
                _lib.ptr(self.G if g_mode else None), _lib.ptr(self.stats_vec), _lib.ptr(self.stats_scratch), self.A,
                self.K, self.E, self.ld, _lib.stream_ptr()))
            return dict(R=env_major(self.R, self.E), modR=env_major(self.modR, self.E), C=env_major(self.Csum, self.E),
                        G=env_major(self.G, self.E) if g_mode else None, stats=Stats(self.stats_vec, self.A, self.K))
        _lib.check(lib.smarl_rollout_returns(
            C.byref(acc), _lib.ptr(self.reward), _lib.ptr(self.cost), self.cost_code, _lib.ptr(self.penalty),
            _lib.ptr(n_active), _lib.ptr(self.R), _lib.ptr(self.modR), _lib.ptr(self.Csum), _lib.ptr(self.G if g_mode else None),
            _lib.ptr(self.stats_vec), _lib.ptr(self.stats_scratch), self.A, self.K, self.E, self.ld,
            _lib.stream_ptr()))
        return dict(R=env_major(self.R, self.E), modR=env_major(self.modR, self.E),
                    C=env_major(self.Csum, self.E),
                    G=env_major(self.G, self.E) if g_mode else None,
                    stats=Stats(self.stats_vec, self.A, self.K))


class BatchedBuffer:
    """Reference ``Buffer`` protocol (buffer.py:7-48) fed with whole batches of episodes.

    ``scores``, ``modified_scores``, ``constraints`` are ``[episodes][A|K]`` like the reference's
    lists (materialised lazily from the device tensors); ``lambdas`` is ``[meta cycles][K]``.
    """

    def __init__(self, params=None, constrained=True, save_path=None):
        self.params, self.constrained, self.save_path = params, constrained, save_path
        self.lambdas = []
        self._R, self._modR, self._C = [], [], []
        self._steps = []          # per-step appends (compat path)

    # -- compat path: per-step tensors [E, A] / [E, K], exactly the reference's call sequence ----
    def append(self, reward, modified_reward, constraint):
        self._steps.append((reward, modified_reward, constraint))

    def step(self):
        gamma = self.params.gamma
        r = torch.stack([s[0] for s in self._steps]).double()
        m = torch.stack([s[1] for s in self._steps]).double()
        c = torch.stack([s[2] for s in self._steps]).double()
        disc = torch.tensor([gamma ** i for i in range(r.shape[0])], dtype=torch.float64, device=r.device)
        self._R.append((disc[:, None, None] * r).sum(0))
        self._modR.append((disc[:, None, None] * m).sum(0))
        self._C.append(c.sum(0))
        self._steps = []

    # -- fast path: the products of RolloutBuffer.finish / env.rollout -----------------------------
    def extend(self, R, modR, C):
        self._R.append(R.double())
        self._modR.append(modR.double())
        self._C.append(C.double())

    def append_lambdas(self, lambdas):
        self.lambdas.append(np.asarray(torch.as_tensor(lambdas).detach().cpu(), dtype=np.float64).copy())

    @staticmethod
    def _cat(chunks):
        return torch.cat(chunks, dim=0).cpu().numpy() if chunks else np.zeros((0, 0))

    @property
    def scores(self):
        return self._cat(self._R)

    @property
    def modified_scores(self):
        return self._cat(self._modR)

    @property
    def constraints(self):
        return self._cat(self._C)

    def mean_score(self, n=100):
        """buffer.py:45-48: means over the last n episodes; costs relative to the thresholds."""
        thr = np.asarray(self.params.thresholds, dtype=np.float64)
        return (self.scores[-n:].mean(axis=0), self.modified_scores[-n:].mean(axis=0),
                list(self.constraints[-n:].mean(axis=0) - thr))

    # -- persistence (buffer.py:50-65,141-153; the matplotlib figures are out of scope) ---------------
    @staticmethod
    def _batch_means(series, batch_size, first):
        """The per-batch series the reference plots (buffer.py:155-177): entry 0 is ``first``, then the
        means over episodes [0, b), [b, 2b), ... for every multiple of ``batch_size`` below the length."""
        out = [np.asarray(first, dtype=np.float64)]
        lo = 0
        for hi in range(batch_size, len(series), batch_size):
            out.append(series[lo:hi].mean(axis=0))
            lo = hi
        return np.stack(out)

    def batch_scores(self):
        """(scores, modified scores) averaged per batch of ``params.batch_size`` episodes, [batches, A]."""
        b = int(self.params.batch_size)
        sc, ms = self.scores, self.modified_scores
        return self._batch_means(sc, b, sc[0]), self._batch_means(ms, b, ms[0])

    def batch_constraints(self):
        """Per-batch mean cost minus threshold, [batches, K] (buffer.py:168-177)."""
        thr = np.asarray(self.params.thresholds, dtype=np.float64)
        c = self.constraints
        return self._batch_means(c - thr, int(self.params.batch_size), c[0] - thr)

    def _results_dir(self):
        if self.save_path is None:
            p = self.params
            stem = os.path.join("results", f"{p.environment}_s{p.size}_n{p.n_agents}_{p.numpy_seed}-{p.torch_seed}_{p.algo}_")
            if not self.constrained:
                stem += "unconstr_"
            i = 0
            while os.path.isdir(stem + str(i)):
                i += 1
            self.save_path = stem + str(i)
        os.makedirs(self.save_path, exist_ok=True)
        return self.save_path

    def save_results(self):
        """Write the files the reference's analysis scripts read (make_graphs.py:17-21):
        ``params.json``, ``constr<numpy_seed>.npy`` [episodes, K], ``scores<numpy_seed>.npy`` [episodes, A]
        and, when constrained, ``lambdas.npy`` [meta cycles, K]; same directory naming as buffer.py:51-64."""
        path = self._results_dir()
        with open(os.path.join(path, "params.json"), "w") as fh:
            json.dump(vars(self.params), fh)
        seed = str(self.params.numpy_seed)
        np.save(os.path.join(path, "constr" + seed + ".npy"), self.constraints)
        np.save(os.path.join(path, "scores" + seed + ".npy"), self.scores)
        if self.constrained:
            np.save(os.path.join(path, "lambdas.npy"), np.asarray(self.lambdas, dtype=np.float64))
        return path
