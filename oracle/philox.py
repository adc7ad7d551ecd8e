"""Philox4x32-10 (Salmon et al., SC'11) in numpy.  TEST INFRASTRUCTURE ONLY.

The reference draws Congestion's action noise from Python's unseeded ``random``
(envs/congestion.py:4,64,67), so it has no reproducible stream to match.  The B200
path uses a counter-based generator keyed by the *global* env id, so results do not
depend on how envs are sharded over GPUs:

    out = philox4x32_10(counter=(env_lo, env_hi, t, agent >> 2 | episode << 3), key=(seed_lo, seed_hi))
    agent a uses the word w = out[a & 3]
    u1 = w * 2**-32,  u2 = (w mod 5 + 0.5) / 5     (both exact enough in f64: int(u2 * 5) == w mod 5)
    move = action if u1 < 1 - noise else int(u2 * 5)      (congestion.py:64-67)

In integers:  keep <=> w < ceil((1-noise) * 2**32);  replacement = w mod 5.  Conditional on the move
being replaced (w >= threshold), w mod 5 is uniform on 0..4 up to one count in 2**32 per outcome, the
same resolution a second 32-bit word would give, so one generator call serves four agents.
This file restates the generator so the test-suite can replay the very same
uniforms through the reference's ``random()`` hook and through the oracle.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays holding uint32 values; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def congestion_words(seed, env_ids, t, n_agents, episode=0):
    """uint32 word w, [E, A], for global env ids ``env_ids`` at step ``t`` of episode ``episode``."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    E = env_ids.shape[0]
    quads = (n_agents + 3) // 4
    c0 = (env_ids & MASK)[:, None]
    c1 = (env_ids >> np.uint64(32))[:, None]
    c2 = np.full((1, 1), t, dtype=np.uint64)
    c3 = np.arange(quads, dtype=np.uint64)[None, :] | np.uint64((int(episode) << 3) & 0xFFFFFFFF)
    o = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)   # 4 x [E, quads]
    w = np.empty((E, 4 * quads), dtype=np.uint32)
    for q in range(4):
        w[:, q::4] = o[q]
    return w[:, :n_agents]


def congestion_uniforms(seed, env_ids, t, n_agents, episode=0):
    """The f64 uniforms (u1, u2) in [0,1) equivalent to the device's integer tests: feeding them to the
    reference's two ``random()`` calls (congestion.py:64,67) reproduces the device's moves."""
    w = congestion_words(seed, env_ids, t, n_agents, episode)
    return w.astype(np.float64) * 2.0 ** -32, ((w % np.uint32(5)).astype(np.float64) + 0.5) / 5.0


def keep_threshold(noise):
    """Smallest integer thr with  (w * 2**-32 < 1 - noise)  <=>  (w < thr)  for uint32 w."""
    import math
    from fractions import Fraction
    lim = Fraction(1 - noise) * (1 << 32)        # exact value of the f64 ``1 - noise`` times 2**32
    return min(1 << 32, max(0, math.ceil(lim)))


def start_uniforms(seed, env_ids, episode, rows):
    """53-bit uniforms (ux, uy), each [E, len(rows)], of the shuffle=True start draws: counter
    (env id, episode, row), key (seed lo, seed hi ^ "RSET"), numpy's random_sample bit recipe."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    rows = np.asarray(rows, dtype=np.uint64)
    o = philox4x32_10((env_ids & MASK)[:, None], (env_ids >> np.uint64(32))[:, None],
                      np.full((1, 1), episode, dtype=np.uint64), rows[None, :],
                      seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ 0x52534554)

    def u53(a, b):
        return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) \
            * (1.0 / 9007199254740992.0)
    return u53(o[0], o[1]), u53(o[2], o[3])


def random_starts(kind, size, seed, env_ids, episode, n_rows, zoom=None, row_offset=0):
    """[E, n_rows, 2] starts: kind 0 floor(u*size); 1 same with row 0 at (0,0); 2 u*size;
    3 floor((u*size)*zoom)/zoom  (Agent.reset in the reference envs)."""
    ux, uy = start_uniforms(seed, env_ids, episode, np.arange(n_rows) + row_offset)
    xy = np.stack([ux * size, uy * size], axis=-1)
    if kind in (0, 1):
        xy = np.floor(xy)
        if kind == 1:
            xy[:, 0] = 0
    elif kind == 3:
        xy = np.floor(xy * zoom) / zoom
    return xy


def policy_uniforms(seed, env_ids, t, n_agents, episode=0):
    """u [E, A] of smarl_policy_act_discrete: Philox counter (env id lo, hi, t | episode << 16, agent >> 2), key
    (seed lo, seed hi ^ "PLCY"); agent a takes word a & 3 of its block; u = ((w >> 8) + 0.5) * 2**-24 (the kernel forms
    it in float32, where k + 0.5 rounds to even for k >= 2**23: off by 2**-25 there, below the margin the tests allow)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    tw = (int(t) | (int(episode) << 16)) & 0xFFFFFFFF
    o = philox4x32_10((env_ids & MASK)[:, None], (env_ids >> np.uint64(32))[:, None], np.full((1, 1), tw, dtype=np.uint64),
                      (np.arange(n_agents, dtype=np.uint64) >> np.uint64(2))[None, :], seed & 0xFFFFFFFF,
                      ((seed >> 32) & 0xFFFFFFFF) ^ 0x504C4359)
    w = np.stack(o, axis=0)[np.arange(n_agents) & 3, :, np.arange(n_agents)].T          # [E, A]: word a & 3
    return ((w >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0


def policy_sample(logits, u):
    """Inverse-CDF Categorical sample as the kernel draws it: logits [..., C] (float64), u [...]:
    a = #{c < C-1 : cumsum(e)[c] <= u * sum(e)}, e = exp(logits - max).  Returns (actions, log_prob of them,
    margin) where margin is the relative distance of u * sum(e) to the nearest CDF step (ties are measure-zero; the
    float32 kernel may legitimately differ where margin is below its rounding)."""
    logits = np.asarray(logits, dtype=np.float64)
    m = logits.max(axis=-1, keepdims=True)
    e = np.exp(logits - m)
    s = e.sum(axis=-1)
    cum = np.cumsum(e, axis=-1)[..., :-1]
    target = (np.asarray(u, dtype=np.float64) * s)[..., None]
    a = (cum <= target).sum(axis=-1)
    logp = np.take_along_axis(logits - m - np.log(s)[..., None], a[..., None], axis=-1)[..., 0]
    margin = np.abs(cum - target).min(axis=-1) / s
    return a, logp, margin


def gaussian_normals(seed, env_ids, t, n_agents, episode=0):
    """z [E, A, 2] of smarl_policy_act_gaussian: Philox counter (env id lo, hi, t | episode << 16, agent >> 1), key
    (seed lo, seed hi ^ "GAUS"); agent a takes words 2 (a & 1), 2 (a & 1) + 1; u = ((w >> 9) + 0.5) * 2**-23 (23 bits:
    exactly representable in float32, never 0 or 1);
    Box-Muller r = sqrt(-2 ln u0), z = (r cos 2 pi u1, r sin 2 pi u1) (float64 here)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    tw = (int(t) | (int(episode) << 16)) & 0xFFFFFFFF
    ag = np.arange(n_agents)
    o = philox4x32_10((env_ids & MASK)[:, None], (env_ids >> np.uint64(32))[:, None], np.full((1, 1), tw, dtype=np.uint64),
                      (ag.astype(np.uint64) >> np.uint64(1))[None, :], seed & 0xFFFFFFFF,
                      ((seed >> 32) & 0xFFFFFFFF) ^ 0x47415553)
    w = np.stack(o, axis=0)                                                              # [4, E, A]
    wa = w[2 * (ag & 1), :, ag].T                                                        # [E, A]
    wb = w[2 * (ag & 1) + 1, :, ag].T
    u0 = ((wa >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    u1 = ((wb >> np.uint32(9)).astype(np.float64) + 0.5) / 8388608.0
    r = np.sqrt(-2.0 * np.log(u0))
    return np.stack([r * np.cos(2.0 * np.pi * u1), r * np.sin(2.0 * np.pi * u1)], axis=-1)


def gaussian_sample(mu, var, z):
    """actions and log-probabilities as the kernel forms them: a = mu + sqrt(var) z, log N(a; mu, diag(var))."""
    mu, var, z = (np.asarray(x, dtype=np.float64) for x in (mu, var, z))
    a = mu + np.sqrt(var) * z
    logp = -0.5 * (((a - mu) ** 2 / var) + np.log(var)).sum(axis=-1) - np.log(2.0 * np.pi)
    return a, logp
