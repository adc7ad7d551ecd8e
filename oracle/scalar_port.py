"""One-env-at-a-time CPU port with the reference's execution model.  TEST INFRASTRUCTURE ONLY.

The reference cannot travel to the GPU box, so the CPU baseline that bench.py reports there
(``cpu_baseline.kind == "port"`` and ``--impl reference``) is this port: like the reference it
holds ONE environment as Python lists of per-agent state, steps it with per-agent Python loops,
calls scipy's ``distance_matrix`` for pair distances and does the per-step MetaAgent / Buffer
bookkeeping in Python -- the same work per env step, hence the same performance character.
Measured in the build container on one core (env.step + MetaAgent.act + Buffer.append/step,
recorded actions, agent-steps/s, reference -> port): Coverage S5 A3 T50 9.5e4 -> 1.1e5;
Coverage S32 A16 T50 1.8e5 -> 2.2e5; Collision S5 A3 T50 4.9e4 -> 4.4e4; Congestion S10 A8
T100 2.6e5 -> 4.0e5, i.e. the port is 0.9-1.55x the reference (a slightly STRONGER baseline,
because it skips some of the reference's per-step list copies and numpy fancy indexing).  It is NOT a
copy of the reference sources: every class below is written against the behaviour pinned by
oracle/numpy_oracle.py and is diff-tested against it (tests/test_scalar_port.py).

Citations are relative to /root/reference.
"""
from __future__ import annotations

import time
import warnings

import numpy as np

warnings.filterwarnings("ignore", category=DeprecationWarning)
from scipy.spatial import distance_matrix  # noqa: E402  (the reference's pair-distance routine)

MOVES = ((1, 0), (-1, 0), (0, -1), (0, 1), (0, 0))     # coverage.py:176 / congestion.py:55


def _clamp(v, hi):
    return max(0, min(hi, v))


class CoveragePort:
    """CoverageDiscrete.step (coverage.py:100-106,174-196,76-89)."""

    def __init__(self, size, n_agents, starts, weights=None, fieldview_size=None):
        self.size, self.n = size, n_agents
        self.fv = size / (np.sqrt(n_agents)) if fieldview_size is None else fieldview_size
        self.weights = weights
        self.starts = [[float(s[0]), float(s[1])] for s in starts]
        self.state = [list(s) for s in self.starts]

    def reset(self):
        self.state = [list(s) for s in self.starts]
        return [list(s) for s in self.state]

    def step(self, action):
        for i in range(self.n):
            dx, dy = MOVES[action[i]]
            x, y = self.state[i]
            self.state[i] = [_clamp(x + dx, self.size), _clamp(y + dy, self.size)]
        dist = distance_matrix(self.state, self.state)
        rew = 0
        for i in range(self.n):
            for j in range(i + 1, self.n):
                gap = self.fv - dist[i, j]
                if gap > 0:
                    rew -= gap ** 2
        reward = [rew] * self.n
        if self.weights is not None:
            reward = [r * w for r, w in zip(reward, self.weights)]
        cost = [0 if a == 4 else 1 for a in action]
        return [list(s) for s in self.state], reward, cost, [False] * self.n


class CongestionPort:
    """Congestion.step with the effective moves supplied (congestion.py:49-137)."""

    def __init__(self, size, n_agents, starts, demand):
        self.size, self.n, self.demand = size, n_agents, np.asarray(demand)
        self.starts = [[float(s[0]), float(s[1])] for s in starts]
        self.state = [list(s) for s in self.starts]
        self.edges = [[0, 0, 0, 0] for _ in range(n_agents)]

    def reset(self):
        self.state = [list(s) for s in self.starts]
        return [list(s) for s in self.state]

    def step(self, action, moves=None):
        moves = action if moves is None else moves
        for i in range(self.n):
            dx, dy = MOVES[moves[i]]
            x, y = self.state[i]
            nx, ny = _clamp(x + dx, self.size), _clamp(y + dy, self.size)
            self.edges[i] = [x, y, nx, ny]
            self.state[i] = [nx, ny]
        # Congestion._congestions (:113-137): classes of identical directed edges, opened by the
        # lowest-index agent that INTENDS to move; earlier intended-stayers are not counted.
        con = [0] * self.n
        for i in range(self.n):
            if con[i] or action[i] >= 4:
                continue
            group = [i] + [j for j in range(i + 1, self.n) if self.edges[j] == self.edges[i]]
            for a in group:
                con[a] = len(group) - 1
        reward = []
        for i in range(self.n):
            if action[i] < 4:
                reward.append(-4.0 - con[i] * 2.0)
            else:
                d = self.demand[int(self.state[i][0]), int(self.state[i][1])]
                reward.append(-30.0 * (con[i] + 1) / d + 7.5 - 4.0)
        at_origin = sum(1 for s in self.state if s == [0, 0])
        return [list(s) for s in self.state], reward, [max(0, self.n // 3 - at_origin)], [False] * self.n


class CollisionPort:
    """CollisionAvoidance.step, shuffle=False (collision_avoidance.py:103-162)."""

    def __init__(self, size, n_agents, starts, landmarks, agents_size=0.25):
        self.size, self.n, self.agents_size = size, n_agents, agents_size
        self.starts = [[float(s[0]), float(s[1])] for s in starts]
        self.landmarks = [[float(l[0]), float(l[1])] for l in landmarks]
        self.reset()

    def reset(self):
        self.state = [list(s) for s in self.starts]
        self.done = [False] * self.n
        return [list(s) for s in self.state]

    def step(self, action):
        for i in range(self.n):
            if self.done[i]:
                continue
            dx, dy = np.squeeze(action[i])
            norm = np.sqrt(dx ** 2 + dy ** 2)
            if norm > 1:
                dx, dy = dx / norm, dy / norm
            x, y = self.state[i]
            self.state[i] = [_clamp(x + dx, self.size), _clamp(y + dy, self.size)]
            for land in self.landmarks:
                if np.linalg.norm(np.array(self.state[i]) - np.array(land)) < self.agents_size:
                    self.done[i] = True
        rew = -np.sum(np.amin(distance_matrix(self.state, self.landmarks), axis=1))
        alive = [s for s, d in zip(self.state, self.done) if not d]
        if alive:
            dist = distance_matrix(alive, alive)
            cost = (np.sum(dist < 2 * self.agents_size) - len(alive)) / 2
        else:
            cost = 0
        return [list(s) for s in self.state], [rew] * self.n, [cost], list(self.done)


class MetaAgentPort:
    """MetaAgent.act / step / update (meta_agent.py:18-39), recording always on."""

    def __init__(self, lambdas, thresholds, lr):
        self.lambdas = np.array(lambdas, dtype=np.float64)
        self.thresholds = np.array(thresholds, dtype=np.float64)
        self.lr = lr
        self.batch, self.values = [], []

    def act(self, constraint, reward):
        self.batch.append(constraint)
        return (-np.inner(self.lambdas, np.array(constraint)) + np.array(reward)).tolist()

    def step(self):
        self.values.append([sum(c) for c in np.array(self.batch).T])
        self.batch = []

    def update(self):
        mean = np.array([np.mean(v) for v in np.array(self.values).T])
        self.lambdas = np.maximum(self.lambdas + self.lr * (mean - self.thresholds), 0.0)
        self.values = []


class BufferPort:
    """Buffer.append / step (buffer.py:22-43)."""

    def __init__(self, gamma):
        self.gamma = gamma
        self.r, self.m, self.c = [], [], []
        self.scores, self.modified_scores, self.constraints = [], [], []

    def append(self, reward, modified_reward, constraint):
        self.r.append(reward); self.m.append(modified_reward); self.c.append(constraint)

    def step(self):
        disc = [self.gamma ** i for i in range(len(self.r) + 1)]
        self.scores.append([sum(g * v for g, v in zip(disc, col)) for col in np.array(self.r).T])
        self.modified_scores.append([sum(g * v for g, v in zip(disc, col)) for col in np.array(self.m).T])
        self.constraints.append([sum(col) for col in np.array(self.c).T])
        self.r, self.m, self.c = [], [], []


def reward_to_go(rewards, gamma):
    """ACAgent.compute_returns (agent.py:200-206)."""
    run, out = 0, []
    for r in reversed(rewards):
        run = r + gamma * run
        out.insert(0, run)
    return out


def run_episode(env, meta, buf, actions, gamma, moves=None):
    """One episode of the driver loop, main.py:28-57, policies replaced by recorded actions."""
    env.reset()
    per_agent = [[] for _ in range(env.n)]
    for t in range(len(actions)):
        if moves is not None:
            state, reward, constraint, done = env.step(actions[t], moves[t])
        else:
            state, reward, constraint, done = env.step(actions[t])
        modified = meta.act(constraint, reward)
        for a, m in enumerate(modified):
            per_agent[a].append(m)
        buf.append(reward, modified, constraint)
        if np.all(done):
            break
    buf.step()
    meta.step()
    return [reward_to_go(r, gamma) for r in per_agent], len(per_agent[0])


def make_workload(env_name, size, n_agents, max_t, seed, n_landmarks=1):
    """Seeded synthetic env + recorded action sequence for the named workload."""
    rng = np.random.default_rng(seed)
    A = n_agents
    if env_name == "coverage":
        env = CoveragePort(size, A, np.floor(rng.random((A, 2)) * size), weights=(1.0 + np.arange(A) % 3).tolist())
        actions = rng.integers(0, 5, size=(max_t, A)).tolist()
        K = A
    elif env_name == "congestion":
        starts = np.floor(rng.random((A, 2)) * size); starts[0] = 0
        env = CongestionPort(size, A, starts, rng.random((size + 1, size + 1)) * 8 + 2)
        actions = rng.integers(0, 5, size=(max_t, A)).tolist()
        K = 1
    elif env_name == "collision":
        env = CollisionPort(size, A, rng.random((A, 2)) * size, rng.random((n_landmarks, 2)) * size)
        actions = rng.normal(0, 0.5, size=(max_t, A, 1, 2)).astype(np.float32).astype(np.float64).tolist()
        K = 1
    else:
        raise ValueError(env_name)
    return env, actions, K


def time_episodes(env_name, size, n_agents, max_t, gamma, seconds, seed=0):
    """Run whole episodes for ~``seconds`` on this core; returns (agent_steps, elapsed_s)."""
    env, actions, K = make_workload(env_name, size, n_agents, max_t, seed)
    meta = MetaAgentPort([0.1] * K, [25.0] * K, 0.002)
    buf = BufferPort(gamma)
    steps, t0 = 0, time.perf_counter()
    while True:
        _, n = run_episode(env, meta, buf, actions, gamma)
        steps += n * n_agents
        el = time.perf_counter() - t0
        if el >= seconds:
            return steps, el


def _worker(args):
    return time_episodes(*args)


def time_all_cores(env_name, size, n_agents, max_t, gamma, seconds, procs):
    """Independent single-env replicas, one per process; returns (agent_steps_per_s, procs)."""
    import multiprocessing as mp
    if procs <= 1:
        s, el = time_episodes(env_name, size, n_agents, max_t, gamma, seconds)
        return s / el, 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(env_name, size, n_agents, max_t, gamma, seconds, i) for i in range(procs)])
    return sum(s / el for s, el in res), procs
