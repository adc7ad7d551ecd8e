"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only):

    python tests/golden/make_golden.py

Each fixture holds seeded inputs plus the reference's own outputs (envs/*.py stepped through
oracle/reference_harness.py, accounting through the reference's MetaAgent / Buffer /
compute_returns).  The GPU box has no /root/reference: the -m gpu tests and the oracle tests
read these files instead.  Stored dtypes are compact but lossless (ints as int16, floats f64).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import philox  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402


def accounting(rew, cost, lam, gamma, thr, lr, active=None):
    """rew [T,E,A], cost [T,E,K] -> reference accounting per env (main.py:36-57,66)."""
    T, E, A = rew.shape
    out = dict(R=[], modR=[], C=[], G=[], disc=[], lam_after=[])
    for e in range(E):
        n = T if active is None else int(active[:, e].sum())
        o = rh.run_accounting(rew[:, e], cost[:, e], lam, gamma, thr, lr, n_steps=n)
        G = np.zeros((T, A)); G[:n] = o["G"]
        D = np.zeros((T, A)); D[:n] = o["disc"]
        out["R"].append(o["R"]); out["modR"].append(o["modR"]); out["C"].append(o["C"])
        out["G"].append(G); out["disc"].append(D); out["lam_after"].append(o["lambdas_after"])
    res = {k: np.stack(v) for k, v in out.items()}
    res["G"] = res["G"].transpose(1, 0, 2)
    res["disc"] = res["disc"].transpose(1, 0, 2)
    return res


def coverage(name, size, A, E, T, weights, lam, gamma, thr, seed, fv=None):
    rng = np.random.default_rng(seed)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int16)
    actions = rng.integers(0, 5, size=(T, E, A)).astype(np.int16)
    pos = np.zeros((T, E, A, 2), np.int16); rew = np.zeros((T, E, A)); cost = np.zeros((T, E, A), np.int16)
    for e in range(E):
        tr = rh.run_coverage_discrete(size, A, starts[e], actions[:, e], weights=weights, fieldview_size=fv)
        pos[:, e], rew[:, e], cost[:, e] = tr["pos"], tr["reward"], tr["cost"]
    acc = accounting(rew, cost.astype(np.float64), lam, gamma, thr, 0.05)
    np.savez_compressed(os.path.join(HERE, name), env="CoverageDiscrete", size=size, n_agents=A, weights=weights,
                        fieldview=tr["fieldview"], lambdas=lam, gamma=gamma, thresholds=thr, meta_lr=0.05,
                        starts=starts, actions=actions, pos=pos, reward=rew, cost=cost, **acc)


def coverage_float(name, kind, size, A, E, T, coarse, weights, lam, gamma, thr, seed):
    rng = np.random.default_rng(seed)
    if kind == "continuous":
        starts = rng.random((E, A, 2)) * size
        actions = rng.normal(0, 0.8, size=(T, E, A, 2)).astype(np.float32)
    else:
        zoom = coarse / size
        starts = np.floor(rng.random((E, A, 2)) * size * zoom) / zoom
        actions = rng.integers(0, 9, size=(T, E, A)).astype(np.int16)
    pos = np.zeros((T, E, A, 2)); rew = np.zeros((T, E, A)); cost = np.zeros((T, E, A))
    for e in range(E):
        tr = rh.run_coverage_float(kind, size, A, starts[e], actions[:, e].astype(np.float64) if kind == "continuous"
                                   else actions[:, e], weights=weights, coarseness=coarse)
        pos[:, e], rew[:, e], cost[:, e] = tr["pos"], tr["reward"], tr["cost"]
    acc = accounting(rew, cost, lam, gamma, thr, 0.05)
    np.savez_compressed(os.path.join(HERE, name), env=kind, size=size, n_agents=A, coarseness=coarse, weights=weights,
                        fieldview=tr["fieldview"], lambdas=lam, gamma=gamma, thresholds=thr, meta_lr=0.05,
                        starts=starts, actions=actions, pos=pos, reward=rew, cost=cost, **acc)


def congestion(name, size, A, E, T, noise, lam, gamma, thr, seed, philox_seed, env_offset):
    rng = np.random.default_rng(seed)
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    if size == 3:
        demand = np.array([[2, 2, 4, 4], [3, 6, 10, 5], [3, 8, 3, 4], [4, 6, 7, 8]], dtype=np.float64)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int16)
    starts[:, 0] = 0
    actions = rng.integers(0, 5, size=(T, E, A)).astype(np.int16)
    ids = np.arange(env_offset, env_offset + E)
    u = np.zeros((T, E, A, 2))
    for t in range(T):
        u[t, :, :, 0], u[t, :, :, 1] = philox.congestion_uniforms(philox_seed, ids, t, A)
    pos = np.zeros((T, E, A, 2), np.int16); rew = np.zeros((T, E, A)); cost = np.zeros((T, E, 1), np.int16)
    con = np.zeros((T, E, A), np.int16); moves = np.zeros((T, E, A), np.int16)
    for e in range(E):
        tr = rh.run_congestion(size, A, starts[e], actions[:, e], demand, noise=noise, uniforms=u[:, e])
        pos[:, e], rew[:, e], cost[:, e], con[:, e] = tr["pos"], tr["reward"], tr["cost"], tr["congestions"]
        # effective move recovered from the reference's own edges (old -> new), only where unambiguous
    acc = accounting(rew, cost.astype(np.float64), lam, gamma, thr, 0.05)
    np.savez_compressed(os.path.join(HERE, name), env="Congestion", size=size, n_agents=A, noise=noise,
                        demand=demand, philox_seed=philox_seed, env_offset=env_offset, lambdas=lam, gamma=gamma,
                        thresholds=thr, meta_lr=0.05, starts=starts, actions=actions, pos=pos, reward=rew,
                        cost=cost, congestions=con, **acc)


def collision(name, size, A, L, E, T, lam, gamma, thr, seed):
    rng = np.random.default_rng(seed)
    starts = rng.random((E, A, 2)) * size
    landmarks = rng.random((E, L, 2)) * size
    actions = rng.normal(0, 0.5, size=(T, E, A, 2)).astype(np.float32)
    q = E // 4
    actions[:, :q] = ((landmarks[:q, :1] - starts[:q]) / 6).astype(np.float32)[None]
    starts[q:2 * q] = np.clip(landmarks[q:2 * q, :1] + rng.normal(0, 0.4, size=(q, A, 2)), 0, size)
    pos = np.zeros((T, E, A, 2)); rew = np.zeros((T, E, A)); cost = np.zeros((T, E, 1))
    done = np.zeros((T, E, A), bool); active = np.zeros((T, E), bool)
    for e in range(E):
        tr = rh.run_collision(size, A, starts[e], landmarks[e], actions[:, e].astype(np.float64), n_landmarks=L)
        pos[:, e], rew[:, e], cost[:, e], done[:, e], active[:, e] = \
            tr["pos"], tr["reward"], tr["cost"], tr["done"], tr["active"]
    acc = accounting(rew, cost, lam, gamma, thr, 0.05, active=active)
    np.savez_compressed(os.path.join(HERE, name), env="Collision", size=size, n_agents=A, n_landmarks=L,
                        lambdas=lam, gamma=gamma, thresholds=thr, meta_lr=0.05, starts=starts,
                        landmarks=landmarks, actions=actions, pos=pos, reward=rew, cost=cost, done=done,
                        active=active, **acc)


if __name__ == "__main__":
    assert rh.available(), "needs the reference tree"
    # BASELINE config 1 (params.json / README values)
    coverage("coverage_c1.npz", 5, 3, 50, 50, [1.0, 2.0, 3.0], [0.1, 0.2, 0.3], 0.999, [25.0, 25.0, 25.0], 0)
    # config 4 shape, few envs
    coverage("coverage_c4.npz", 32, 16, 24, 50, (1.0 + np.arange(16) % 3).tolist(),
             np.linspace(0.05, 0.4, 16).tolist(), 0.999, [25.0] * 16, 1)
    # paper Congestion config + config 3 shape (seeded demand table; the reference's own is 4x4)
    congestion("congestion_paper.npz", 3, 3, 48, 10, 0.1, [0.5], 0.9, [1.5], 2, 777, 0)
    congestion("congestion_c3.npz", 10, 8, 24, 100, 0.1, [0.35], 0.9, [1.5], 3, 4242, 1_000_000)
    # config 2 shape and the paper's Collision config
    collision("collision_c2.npz", 5, 3, 1, 64, 50, [0.5], 0.99, [1.0], 4)
    collision("collision_paper.npz", 2, 5, 1, 32, 20, [0.8], 0.99, [1.0], 5)
    # the paper's "ExploreContinuous" launcher config (coarseness 6, CLI defaults size 3 / 3 agents / max_t 8)
    coverage_float("coverage_continuous_paper.npz", "continuous", 3, 3, 32, 8, 6, [1.0, 1.0, 1.0], [0.2, 0.2, 0.2], 0.99,
                   [15.0] * 3, 6)
    coverage_float("coverage_discretized.npz", "discretized", 5, 3, 16, 30, 20, [1.0, 2.0, 3.0], [0.1, 0.2, 0.3], 0.999,
                   [5.0] * 3, 7)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
