"""Vectorised numpy restatement of the reference env-step hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Arrays are env-major
(``[E, A, ...]``) and float64/int64, i.e. the reference's own arithmetic types,
batched over E independent environment instances.  Citations are relative to
/root/reference.

Parity pin: checked against the live reference (tests/test_oracle_vs_reference.py)
and against tests/golden/*.npz generated from it.
"""
from __future__ import annotations

import numpy as np

# Shared direction table: envs/coverage.py:176, envs/congestion.py:55
#   action 0 -> (+1,0)  1 -> (-1,0)  2 -> (0,-1)  3 -> (0,+1)  4 -> stay
DIR_X = np.array([1, -1, 0, 0, 0], dtype=np.int64)
DIR_Y = np.array([0, 0, -1, 1, 0], dtype=np.int64)
# envs/coverage.py:221 (CoverageDiscretized: 9 actions, 8 = stay)
DIR9_X = np.array([1, -1, 0, 0, 1, 1, -1, -1, 0], dtype=np.int64)
DIR9_Y = np.array([0, 0, -1, 1, 1, -1, 1, -1, 0], dtype=np.int64)

# envs/congestion.py:7-10
HOURLY_COMPENSATION = 30.0
AVERAGE_RIDE_COMPENSATION = 7.5
AVERAGE_RIDE_COST = 4.0
CONGESTION_COST = 2.0


# --------------------------------------------------------------------------------------
# Coverage ("Explore") -- CoverageDiscrete
# --------------------------------------------------------------------------------------
def coverage_fieldview(size, n_agents, fieldview_size=None):
    """envs/coverage.py:15-18: fv = size / sqrt(n_agents) unless overridden."""
    if fieldview_size is None:
        return size / (np.sqrt(n_agents))
    return fieldview_size


def coverage_penalty_lut(size, fv, n=None):
    """Pairwise overlap penalty as a function of the integer squared distance q.

    envs/coverage.py:78-83: ``D = distance_matrix(states, states)`` (scipy:
    sqrt(dx*dx+dy*dy) in f64, exact for integer coordinates => D = sqrt(q)), then
    ``if fv - D > 0: rew -= (fv - D) ** 2`` with numpy *scalar* arithmetic.  The
    same scalar expression is evaluated here so the table is bit-identical.
    """
    if n is None:
        n = 2 * size * size + 1
    lut = np.zeros(n, dtype=np.float64)
    for q in range(n):
        d = np.sqrt(np.float64(q))
        if fv - d > 0:
            lut[q] = (fv - d) ** 2
    return lut


def pair_sqdist(pos):
    """Integer squared distances for all ordered pairs: [E, A, A]."""
    dx = pos[:, :, None, 0] - pos[:, None, :, 0]
    dy = pos[:, :, None, 1] - pos[:, None, :, 1]
    return dx * dx + dy * dy


def coverage_reward_scalar(pos, lut):
    """Un-weighted env reward ``rew`` (one per env), envs/coverage.py:79-83.

    Sequential f64 accumulation in the reference's i-major pair order, so the
    result is bit-identical to the reference's Python loop.
    """
    E, A, _ = pos.shape
    rew = np.zeros(E, dtype=np.float64)
    x = pos[:, :, 0].astype(np.int64)
    y = pos[:, :, 1].astype(np.int64)
    for i in range(A):
        for j in range(i + 1, A):
            dx = x[:, i] - x[:, j]
            dy = y[:, i] - y[:, j]
            rew = rew - lut[dx * dx + dy * dy]
    return rew


def coverage_discrete_step(pos, actions, size, lut, weights=None):
    """One CoverageDiscrete.step for E envs.

    transition envs/coverage.py:174-189, reward :76-89, constraint :191-196,
    check_done :97-98.

    pos      [E, A, 2] int   coordinates in 0..size (inclusive clamp, :185-186)
    actions  [E, A]    int   0..4
    returns  new_pos [E,A,2] int64, reward [E,A] f64, cost [E,A] int64, done [E,A] bool
    """
    pos = np.asarray(pos, dtype=np.int64)
    actions = np.asarray(actions, dtype=np.int64)
    E, A, _ = pos.shape
    new = np.empty_like(pos)
    new[:, :, 0] = np.clip(pos[:, :, 0] + DIR_X[actions], 0, size)
    new[:, :, 1] = np.clip(pos[:, :, 1] + DIR_Y[actions], 0, size)
    rew = coverage_reward_scalar(new, lut)
    reward = np.repeat(rew[:, None], A, axis=1)
    if weights is not None:
        # coverage.py:86-87 zips and therefore truncates; we require >= A weights.
        w = np.asarray(weights, dtype=np.float64)[:A]
        assert w.shape[0] == A, "need at least n_agents weights"
        reward = reward * w[None, :]
    cost = (actions != 4).astype(np.int64)       # travelled_distance = [1,1,1,1,0]
    done = np.zeros((E, A), dtype=bool)
    return new, reward, cost, done


# --------------------------------------------------------------------------------------
# Coverage on float positions -- CoverageContinuous / CoverageDiscretized
# --------------------------------------------------------------------------------------
_POW_UFUNC = None


def _pow2(v):
    """libm pow(v, 2.0) element-wise: what numpy's *scalar* ``x ** 2`` evaluates (coverage.py:83);
    it differs from x*x by one ulp on ~0.08 % of inputs."""
    global _POW_UFUNC
    if _POW_UFUNC is None:
        import ctypes
        import ctypes.util
        f = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6").pow
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double, ctypes.c_double]
        _POW_UFUNC = np.frompyfunc(lambda a: f(a, 2.0), 1, 1)
    return _POW_UFUNC(np.asarray(v, dtype=np.float64)).astype(np.float64)


def coverage_float_reward(pos, fv, weights=None, exact_pow=True):
    """CoverageContinuous.reward (envs/coverage.py:76-89) on float64 positions [E, A, 2].

    D_ij = sqrt(dx*dx + dy*dy) (scipy distance_matrix, array ops); i-major sequential sum of
    (fv - D_ij)**2 over pairs with fv - D_ij > 0.  ``exact_pow`` uses libm pow like the reference's
    numpy scalar power; False uses a multiply (what the CUDA kernel does; <= 1 ulp per term)."""
    pos = np.asarray(pos, dtype=np.float64)
    E, A, _ = pos.shape
    rew = np.zeros(E, dtype=np.float64)
    for i in range(A):
        for j in range(i + 1, A):
            dx = np.abs(pos[:, j, 0] - pos[:, i, 0])
            dy = np.abs(pos[:, j, 1] - pos[:, i, 1])
            gap = fv - np.sqrt(dx * dx + dy * dy)
            pen = _pow2(gap) if exact_pow else gap * gap
            rew = np.where(gap > 0, rew - pen, rew)
    reward = np.repeat(rew[:, None], A, axis=1)
    if weights is not None:
        reward = reward * np.asarray(weights, dtype=np.float64)[None, :A]
    return reward


def coverage_continuous_step(pos, actions, size, fv, weights=None, coarseness=None, exact_pow=True):
    """One CoverageContinuous.step (coverage.py:54-74 transition, :76-89 reward, :91-95 constraint).

    pos [E,A,2] f64, actions [E,A,2] f64 (fp32-origin).  With ``coarseness`` the move is rescaled to
    max_norm = sqrt(2)*size/coarseness when sqrt(norm) > max_norm (sic: the reference compares the
    square root of the norm, :67).  cost_a = np.linalg.norm(action_a) of the ORIGINAL action."""
    pos = np.asarray(pos, dtype=np.float64)
    actions = np.asarray(actions, dtype=np.float64)
    dx, dy = actions[:, :, 0].copy(), actions[:, :, 1].copy()
    if coarseness is not None:
        norm = np.sqrt(dx * dx + dy * dy)
        max_norm = np.sqrt(2) * size / coarseness
        big = np.sqrt(norm) > max_norm
        with np.errstate(invalid="ignore", divide="ignore"):
            dx = np.where(big, (dx / norm) * max_norm, dx)
            dy = np.where(big, (dy / norm) * max_norm, dy)
    new = np.empty_like(pos)
    new[:, :, 0] = np.maximum(0.0, np.minimum(float(size), pos[:, :, 0] + dx))
    new[:, :, 1] = np.maximum(0.0, np.minimum(float(size), pos[:, :, 1] + dy))
    reward = coverage_float_reward(new, fv, weights, exact_pow)
    ax, ay = actions[:, :, 0], actions[:, :, 1]
    cost = np.sqrt(_fma(ay, ay, ax * ax))                   # np.linalg.norm of the 2-vector
    return new, reward, cost, np.zeros(pos.shape[:2], dtype=bool)


def coverage_discretized_step(pos, actions, size, coarseness, fv, weights=None, exact_pow=True):
    """One CoverageDiscretized.step (coverage.py:219-241): 9 lattice moves on a 1/zoom grid,
    zoom = coarseness/size; state stays float: x' = max(0, min(size*zoom, x*zoom + dx)) / zoom."""
    pos = np.asarray(pos, dtype=np.float64)
    actions = np.asarray(actions, dtype=np.int64)
    zoom = coarseness / size
    hi = size * zoom
    new = np.empty_like(pos)
    new[:, :, 0] = np.maximum(0.0, np.minimum(hi, pos[:, :, 0] * zoom + DIR9_X[actions])) / zoom
    new[:, :, 1] = np.maximum(0.0, np.minimum(hi, pos[:, :, 1] * zoom + DIR9_Y[actions])) / zoom
    reward = coverage_float_reward(new, fv, weights, exact_pow)
    table = np.array([1, 1, 1, 1, np.sqrt(2), np.sqrt(2), np.sqrt(2), np.sqrt(2), 0]) * (size / coarseness)
    return new, reward, table[actions], np.zeros(pos.shape[:2], dtype=bool)


# --------------------------------------------------------------------------------------
# Congestion
# --------------------------------------------------------------------------------------
def congestion_noise_moves(actions, u1, u2, noise):
    """envs/congestion.py:64-67: ``move = a if u1 < 1 - noise else int(u2 * 5)``."""
    actions = np.asarray(actions, dtype=np.int64)
    repl = (np.asarray(u2) * 5).astype(np.int64)
    return np.where(np.asarray(u1) < 1 - noise, actions, repl)


def congestions_literal(actions, edges):
    """Literal single-env restatement of Congestion._congestions (congestion.py:113-137).

    actions: list of A intended actions; edges: list of A 4-tuples.  Kept literal
    (including the never-true ``== 5`` branch) to validate the closed form below.
    """
    A = len(actions)
    con = [0] * A
    for i in range(A):
        if con[i]:
            continue
        if actions[i] < 4:
            share = [i]
            for j in range(i + 1, A):
                if actions[j] < 5 and tuple(edges[j]) == tuple(edges[i]):
                    share.append(j)
            for a in share:
                con[a] = len(share) - 1
        else:
            share = [i]
            for j in range(i + 1, A):
                if actions[j] == 5 and tuple(edges[j][:2]) == tuple(edges[i][:2]):
                    share.append(j)
            for a in share:
                con[a] = len(share) - 1
    return con


def congestions_closed_form(actions, edge_key):
    """Vectorised closed form of Congestion._congestions.

    Agents are partitioned by exact directed-edge equality.  Inside a class, with
    L the lowest index whose *intended* action is a move (<4): members with index
    >= L get (#members >= L) - 1, members below L (intended stayers) get 0; classes
    without an intended mover get 0.

    actions [E, A] intended; edge_key [E, A] any integer key unique per (x,y,x',y').
    """
    actions = np.asarray(actions)
    E, A = actions.shape
    same = edge_key[:, :, None] == edge_key[:, None, :]            # [E, i, j]
    mover = actions < 4                                            # [E, A]
    lower_eq = np.tril(np.ones((A, A), dtype=bool))                # j <= i
    # active_i: some k <= i in i's class intends to move  <=>  i >= L(class)
    active = (same & lower_eq[None] & mover[:, None, :]).any(axis=2)
    count = (same & active[:, None, :]).sum(axis=2)                # includes i itself if active
    return np.where(active, count - 1, 0).astype(np.int64)


def congestion_step(pos, actions, moves, size, demand):
    """One Congestion.step for E envs with the effective moves given.

    transition envs/congestion.py:49-75 (``moves`` = the post-noise move actually
    applied, :64-69), reward :77-90 (branches on the *intended* action, indexes
    demand_rate with the *new* position :86), constraint :93-100, done :103-104.

    pos [E,A,2] int, actions/moves [E,A] int 0..4, demand [(size+1),(size+1)] numeric
    returns new_pos, reward [E,A] f64, cost [E,1] int64, done [E,A] bool, congestions [E,A]
    """
    pos = np.asarray(pos, dtype=np.int64)
    actions = np.asarray(actions, dtype=np.int64)
    moves = np.asarray(moves, dtype=np.int64)
    demand = np.asarray(demand)
    E, A, _ = pos.shape
    new = np.empty_like(pos)
    new[:, :, 0] = np.clip(pos[:, :, 0] + DIR_X[moves], 0, size)
    new[:, :, 1] = np.clip(pos[:, :, 1] + DIR_Y[moves], 0, size)
    W = size + 1
    key = ((pos[:, :, 0] * W + pos[:, :, 1]) * W + new[:, :, 0]) * W + new[:, :, 1]
    con = congestions_closed_form(actions, key)
    d = demand[new[:, :, 0], new[:, :, 1]]
    move_rew = -AVERAGE_RIDE_COST - con * CONGESTION_COST
    wait_rew = -HOURLY_COMPENSATION * (con + 1) / d + AVERAGE_RIDE_COMPENSATION - AVERAGE_RIDE_COST
    reward = np.where(actions < 4, move_rew, wait_rew).astype(np.float64)
    at_origin = ((new[:, :, 0] == 0) & (new[:, :, 1] == 0)).sum(axis=1)
    cost = np.maximum(0, A // 3 - at_origin).astype(np.int64)[:, None]
    done = np.zeros((E, A), dtype=bool)
    return new, reward, cost, done, con


# --------------------------------------------------------------------------------------
# CollisionAvoidance
# --------------------------------------------------------------------------------------
def _load_libm_fma():
    import ctypes
    import ctypes.util
    for name in (ctypes.util.find_library("m"), "libm.so.6"):
        if not name:
            continue
        try:
            f = ctypes.CDLL(name).fma
        except (OSError, AttributeError):
            continue
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double] * 3
        return np.frompyfunc(lambda a, b, c: f(a, b, c), 3, 1)
    raise RuntimeError("libm fma not found")


_FMA_UFUNC = None


def _fma(a, b, c):
    """Correctly-rounded fused a*b+c on float64 arrays (libm ``fma`` via ctypes; Python
    3.12 has no math.fma and numpy has no fma ufunc)."""
    global _FMA_UFUNC
    if _FMA_UFUNC is None:
        _FMA_UFUNC = _load_libm_fma()
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=np.float64),
                                  np.asarray(b, dtype=np.float64),
                                  np.asarray(c, dtype=np.float64))
    return _FMA_UFUNC(a, b, c).astype(np.float64)


def collision_step(pos, done, actions, landmarks, size, agents_size=0.25):
    """One CollisionAvoidance.step for E envs (shuffle=False state layout).

    transition envs/collision_avoidance.py:103-125, reward :127-130/:158-162,
    constraint :132-133/:150-156, check_done :135-136.  Envs whose agents were
    all done *before* this step are inactive (the reference driver has already
    broken out of the episode, main.py:51): state frozen, reward/cost 0.

    pos [E,A,2] f64, done [E,A] bool, actions [E,A,2] f64 (fp32-origin values:
    then ``dx**2`` == dx*dx exactly and positions are bit-exact; for arbitrary
    f64 actions the reference's libm pow may differ by 1 ulp in the clip norm),
    landmarks [E,L,2] f64.
    returns new_pos, reward [E,A] f64, cost [E,1] f64, new_done [E,A] bool, active [E] bool
    """
    pos = np.asarray(pos, dtype=np.float64)
    done = np.asarray(done, dtype=bool)
    actions = np.asarray(actions, dtype=np.float64)
    landmarks = np.asarray(landmarks, dtype=np.float64)
    E, A, _ = pos.shape
    active = ~done.all(axis=1)
    dx = actions[:, :, 0].copy()
    dy = actions[:, :, 1].copy()
    norm = np.sqrt(dx * dx + dy * dy)                      # :113
    clip = norm > 1                                        # :114-117
    with np.errstate(invalid="ignore", divide="ignore"):
        dx = np.where(clip, dx / norm, dx)
        dy = np.where(clip, dy / norm, dy)
    nx = np.maximum(0.0, np.minimum(float(size), pos[:, :, 0] + dx))   # :118
    ny = np.maximum(0.0, np.minimum(float(size), pos[:, :, 1] + dy))   # :119
    moving = (~done) & active[:, None]
    new = pos.copy()
    new[:, :, 0] = np.where(moving, nx, pos[:, :, 0])
    new[:, :, 1] = np.where(moving, ny, pos[:, :, 1])
    # :122-124  np.linalg.norm(state - land) == sqrt(fma(ddy, ddy, ddx*ddx)) [probed]
    ddx = new[:, :, None, 0] - landmarks[:, None, :, 0]    # [E, A, L]
    ddy = new[:, :, None, 1] - landmarks[:, None, :, 1]
    reach = (np.sqrt(_fma(ddy, ddy, ddx * ddx)) < agents_size).any(axis=2)
    new_done = done | (moving & reach)
    # reward :158-162: distance_matrix(states, landmarks) = sqrt((lx-px)^2 + (ly-py)^2)
    ldx = landmarks[:, None, :, 0] - new[:, :, None, 0]
    ldy = landmarks[:, None, :, 1] - new[:, :, None, 1]
    dist = np.sqrt(ldx * ldx + ldy * ldy)
    rew = -np.sum(np.ascontiguousarray(np.amin(dist, axis=2)), axis=1)
    reward = np.repeat(rew[:, None], A, axis=1)
    # constraint :150-156 over agents that are not done AFTER this step
    pdx = new[:, :, None, 0] - new[:, None, :, 0]
    pdy = new[:, :, None, 1] - new[:, None, :, 1]
    D = np.sqrt(pdx * pdx + pdy * pdy)
    alive = ~new_done
    close = (D < 2 * agents_size) & alive[:, :, None] & alive[:, None, :]
    iu = np.triu(np.ones((A, A), dtype=bool), k=1)
    cost = (close & iu[None]).sum(axis=(1, 2)).astype(np.float64)[:, None]
    reward = np.where(active[:, None], reward, 0.0)
    cost = np.where(active[:, None], cost, 0.0)
    return new, reward, cost, new_done, active


# --------------------------------------------------------------------------------------
# Rollout accounting: MetaAgent.act, Buffer.step, compute_returns, MetaAgent.update
# --------------------------------------------------------------------------------------
def modified_reward(reward, cost, lambdas):
    """safe_multi_agent_RL/meta_agent.py:21-22 (leq=True): r - <lambda, c>.

    reward [..., A], cost [..., K], lambdas [K] -> [..., A]
    """
    pen = np.asarray(cost, dtype=np.float64) @ np.asarray(lambdas, dtype=np.float64)
    return -pen[..., None] + np.asarray(reward, dtype=np.float64)


def episode_returns(rewards, gamma):
    """safe_multi_agent_RL/buffer.py:31-35: R_a = sum_t gamma**t * r[t, a] (left to right).

    rewards [T, E, A] (steps past an episode's end must be 0) -> [E, A]
    """
    rewards = np.asarray(rewards, dtype=np.float64)
    R = np.zeros(rewards.shape[1:], dtype=np.float64)
    for t in range(rewards.shape[0]):
        R = R + (gamma ** t) * rewards[t]
    return R


def episode_cost_sums(costs):
    """buffer.py:39 / meta_agent.py:28: C_k = sum_t c[t, k] (undiscounted). [T,E,K] -> [E,K]"""
    return np.asarray(costs).sum(axis=0)


def reward_to_go(mod_rewards, gamma):
    """safe_multi_agent_RL/agent.py:200-206: G_t = r_t + gamma * G_{t+1}. [T,E,A] -> [T,E,A]"""
    mod_rewards = np.asarray(mod_rewards, dtype=np.float64)
    G = np.zeros_like(mod_rewards)
    run = np.zeros(mod_rewards.shape[1:], dtype=np.float64)
    for t in reversed(range(mod_rewards.shape[0])):
        run = mod_rewards[t] + gamma * run
        G[t] = run
    return G


def discounted_terms(mod_rewards, gamma):
    """agent.py:129-132 (legacy/REINFORCE/PPO): gamma**t * r_t. [T,E,A] -> [T,E,A]"""
    mod_rewards = np.asarray(mod_rewards, dtype=np.float64)
    disc = np.array([gamma ** t for t in range(mod_rewards.shape[0])])
    return disc[:, None, None] * mod_rewards


def ppo_standardised_returns(mod_rewards, gamma, n_active=None):
    """PPOAgent.step (agent.py:276-281; PPOAgent extends ACAgent, so compute_returns is the
    reward-to-go of :200-206): x_t = G_t over the episode's T' steps, then
    (x - mean) / (std_unbiased + 1e-7); steps past T' are 0.  The reference does this in torch
    float32; here float64.  mod_rewards [T, E, A] (zero past T'), n_active [E] or None -> [T, E, A]."""
    x = reward_to_go(mod_rewards, gamma)
    T, E, A = x.shape
    n = np.full(E, T) if n_active is None else np.asarray(n_active)
    live = (np.arange(T)[:, None] < n[None, :])[:, :, None]
    xs = np.where(live, x, 0.0)
    cnt = n[None, :, None].astype(np.float64)
    mean = xs.sum(0, keepdims=True) / cnt
    with np.errstate(invalid="ignore", divide="ignore"):
        var = np.where(live, (x - mean) ** 2, 0.0).sum(0, keepdims=True) / (cnt - 1)
        out = (x - mean) / (np.sqrt(var) + 1e-7)
    return np.where(live, out, 0.0)


def lambda_update(lambdas, mean_cost, thresholds, lr):
    """meta_agent.py:32-36 (leq=True): lambda <- max(0, lambda + lr * (mean C - thr))."""
    lam = np.asarray(lambdas, dtype=np.float64) + lr * (
        np.asarray(mean_cost, dtype=np.float64) - np.asarray(thresholds, dtype=np.float64))
    return np.maximum(lam, 0.0)


def rollout(step_fn, T, gamma, lambdas):
    """Generic driver following main.py:28-57 for a batched ``step_fn(t)`` that
    returns (reward [E,A], cost [E,K]); returns per-step and per-episode products."""
    rs, cs = [], []
    for t in range(T):
        r, c = step_fn(t)
        rs.append(np.asarray(r, dtype=np.float64))
        cs.append(np.asarray(c, dtype=np.float64))
    rs = np.stack(rs)
    cs = np.stack(cs)
    mod = modified_reward(rs, cs, lambdas)
    return dict(reward=rs, cost=cs, mod_reward=mod,
                R=episode_returns(rs, gamma), modR=episode_returns(mod, gamma),
                C=episode_cost_sums(cs), G=reward_to_go(mod, gamma))
