"""Time the UNMODIFIED reference's own driver loop (main.py:28-57 minus the policy networks) on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY: used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg.  The
reference is pure Python; ``__graft_entry__.build()`` copies its two packages (envs/, safe_multi_agent_RL/) from
/root/reference into the git-ignored ``oracle/_ref/`` so that they travel to the GPU box with the snapshot (nothing
of it is tracked in the repository).  Where that copy is absent, bench.py falls back to oracle/scalar_port.py
(``kind: "port"``).

The loop per episode, all reference code: ``env.reset()``; per step ``env.step(actions)`` (CoverageDiscrete /
Congestion / CollisionAvoidance), ``MetaAgent.act`` (meta_agent.py:18-23), the driver's per-agent reward lists
(main.py:46-47) and ``Buffer.append`` (buffer.py:22-25); per episode ``Buffer.step`` (:30-43), ``MetaAgent.step``
(:25-30) and ``ACAgent.compute_returns`` (agent.py:200-206) per agent.  Actions are pre-recorded random draws.
"""
from __future__ import annotations

import os
import sys
import time
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")


def root():
    """The reference tree to time: the in-repo copy, else the mounted original, else None."""
    for r in (REF_COPY, os.environ.get("SMARL_REFERENCE_ROOT", "/root/reference")):
        if os.path.isfile(os.path.join(r, "envs", "coverage.py")) and \
                os.path.isfile(os.path.join(r, "safe_multi_agent_RL", "buffer.py")):
            return r
    return None


_ns = None


def load():
    global _ns
    if _ns is not None:
        return _ns
    r = root()
    if r is None:
        raise RuntimeError("no reference tree (oracle/_ref or /root/reference)")
    sys.dont_write_bytecode = True
    if r not in sys.path:
        sys.path.insert(0, r)
    for name in ("matplotlib", "matplotlib.pyplot"):       # only Buffer.save_results plots (buffer.py:2)
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import envs.collision_avoidance as collision
    import envs.congestion as congestion
    import envs.coverage as coverage
    from safe_multi_agent_RL.agent import ACAgent
    from safe_multi_agent_RL.buffer import Buffer
    from safe_multi_agent_RL.meta_agent import MetaAgent
    _ns = types.SimpleNamespace(coverage=coverage, congestion=congestion, collision=collision, MetaAgent=MetaAgent,
                                Buffer=Buffer, ACAgent=ACAgent, root=r)
    return _ns


def make_workload(env_name, size, n_agents, max_t, seed):
    """Same seeded workload as oracle/scalar_port.make_workload, on the reference's own classes."""
    ref = load()
    rng = np.random.default_rng(seed)
    A = n_agents
    if env_name == "coverage":
        env = ref.coverage.CoverageDiscrete(size, A, shuffle=False, weights=(1.0 + np.arange(A) % 3).tolist())
        for ag, s in zip(env.agents, np.floor(rng.random((A, 2)) * size)):
            ag.start = [float(s[0]), float(s[1])]
        actions = rng.integers(0, 5, size=(max_t, A)).tolist()
    elif env_name == "congestion":
        env = ref.congestion.Congestion(size, A, noise=0.0)
        starts = np.floor(rng.random((A, 2)) * size); starts[0] = 0
        for ag, s in zip(env.agents, starts):
            ag.start = [float(s[0]), float(s[1])]
        env.demand_rate = rng.random((size + 1, size + 1)) * 8 + 2      # the shipped table is 4x4 (congestion.py:28)
        actions = rng.integers(0, 5, size=(max_t, A)).tolist()
    else:
        env = ref.collision.CollisionAvoidance(size, A, shuffle=False)
        for ag, s in zip(env.agents, rng.random((A, 2)) * size):
            ag.start = [float(s[0]), float(s[1])]
        env.landmarks = (rng.random((1, 2)) * size).tolist()
        actions = rng.normal(0, 0.5, size=(max_t, A, 1, 2)).astype(np.float32).astype(np.float64).tolist()
    return env, actions, int(sum(env.constraint_space))


def time_episodes(env_name, size, n_agents, max_t, gamma, seconds, seed=0):
    """Whole episodes for ~``seconds`` on this core; returns (agent_steps, elapsed_s)."""
    ref = load()
    env, actions, K = make_workload(env_name, size, n_agents, max_t, seed)
    meta = ref.MetaAgent(env.constraint_space, gamma, 0.002, [25.0] * K, start_learning_cycle=0, lambda_0=0.1)
    buf = ref.Buffer(types.SimpleNamespace(gamma=gamma, thresholds=[25.0] * K))
    steps, t0 = 0, time.perf_counter()
    while True:
        env.reset()
        per_agent = [[] for _ in range(n_agents)]
        n = 0
        for t in range(max_t):
            state, reward, constraint, done = env.step(actions[t])
            modified = meta.act(constraint, reward)
            for a, m in enumerate(modified):
                per_agent[a].append(m)
            buf.append(reward, modified, constraint)
            n += 1
            if np.all(done):
                break
        buf.step()
        meta.step()
        for r in per_agent:                                      # ACAgent.step's returns (agent.py:200-206, :209)
            ref.ACAgent.compute_returns(types.SimpleNamespace(rewards=r, gamma=gamma))
        if len(buf.scores) > 64:                                 # the reference keeps every episode; cap the lists
            del buf.scores[:], buf.modified_scores[:], buf.constraints[:], meta.constraint_values[:]
        steps += n * n_agents
        el = time.perf_counter() - t0
        if el >= seconds:
            return steps, el


def _worker(args):
    return time_episodes(*args)


def time_all_cores(env_name, size, n_agents, max_t, gamma, seconds, procs):
    """Independent single-env replicas of the reference, one per process; returns (agent_steps_per_s, procs)."""
    import multiprocessing as mp
    if procs <= 1:
        s, el = time_episodes(env_name, size, n_agents, max_t, gamma, seconds)
        return s / el, 1
    load()                                                       # import before forking
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(env_name, size, n_agents, max_t, gamma, seconds, i) for i in range(procs)])
    return sum(s / el for s, el in res), procs
