"""The one-env-at-a-time CPU port (the bench's cpu_baseline on the GPU box) against the numpy
oracle, bit-exactly, plus its driver loop against the oracle's accounting."""
import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import scalar_port as sp


@pytest.mark.parametrize("size,A,T,seed", [(5, 3, 50, 0), (32, 16, 10, 1), (8, 32, 4, 2)])
def test_coverage_port(size, A, T, seed):
    rng = np.random.default_rng(seed)
    starts = np.floor(rng.random((A, 2)) * size).astype(np.int64)
    actions = rng.integers(0, 5, size=(T, A))
    w = (1.0 + np.arange(A) % 3).tolist()
    env = sp.CoveragePort(size, A, starts, weights=w)
    lut = no.coverage_penalty_lut(size, no.coverage_fieldview(size, A))
    pos = starts[None]
    env.reset()
    for t in range(T):
        st, r, c, d = env.step(actions[t].tolist())
        pos, r_o, c_o, _ = no.coverage_discrete_step(pos, actions[t][None], size, lut, w)
        assert np.array_equal(np.array(st), pos[0]) and np.array_equal(np.array(r), r_o[0])
        assert np.array_equal(np.array(c), c_o[0])


@pytest.mark.parametrize("size,A,T,seed", [(3, 3, 10, 0), (10, 8, 60, 1), (2, 7, 80, 2)])
def test_congestion_port(size, A, T, seed):
    rng = np.random.default_rng(seed)
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    starts = np.floor(rng.random((A, 2)) * size).astype(np.int64)
    actions = rng.integers(0, 5, size=(T, A))
    moves = np.where(rng.random((T, A)) < 0.8, actions, rng.integers(0, 5, size=(T, A)))
    env = sp.CongestionPort(size, A, starts, demand)
    pos = starts[None]
    env.reset()
    for t in range(T):
        st, r, c, d = env.step(actions[t].tolist(), moves[t].tolist())
        pos, r_o, c_o, _, _ = no.congestion_step(pos, actions[t][None], moves[t][None], size, demand)
        assert np.array_equal(np.array(st), pos[0]) and np.array_equal(np.array(r), r_o[0])
        assert np.array_equal(np.array(c), c_o[0])


@pytest.mark.parametrize("size,A,L,T,seed", [(5, 3, 1, 50, 0), (2, 5, 1, 20, 1), (5, 9, 2, 20, 2)])
def test_collision_port(size, A, L, T, seed):
    rng = np.random.default_rng(seed)
    lm = rng.random((L, 2)) * size
    starts = np.clip(lm[0][None] + rng.normal(0, 0.6, size=(A, 2)), 0, size)
    actions = rng.normal(0, 0.5, size=(T, A, 2)).astype(np.float32).astype(np.float64)
    env = sp.CollisionPort(size, A, starts, lm)
    pos, done = starts[None].copy(), np.zeros((1, A), bool)
    env.reset()
    for t in range(T):
        if done.all():
            break
        st, r, c, d = env.step(actions[t][:, None, :].tolist())
        pos, r_o, c_o, done, _ = no.collision_step(pos, done, actions[t][None], lm[None], size)
        assert np.array_equal(np.array(st), pos[0]) and np.array_equal(np.array(r), r_o[0])
        assert np.array_equal(np.array(c), c_o[0]) and np.array_equal(np.array(d), done[0])


def test_driver_loop_accounting():
    size, A, T, gamma = 5, 3, 30, 0.999
    env, actions, K = sp.make_workload("coverage", size, A, T, seed=3)
    meta = sp.MetaAgentPort([0.1, 0.2, 0.3], [25.0] * 3, 0.05)
    buf = sp.BufferPort(gamma)
    G, n = sp.run_episode(env, meta, buf, actions, gamma)
    assert n == T
    lut = no.coverage_penalty_lut(size, no.coverage_fieldview(size, A))
    pos = np.array(env.starts, dtype=np.int64)[None]
    acts = np.array(actions)

    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, acts[t][None], size, lut, env.weights)
        return r, c
    want = no.rollout(step_fn, T, gamma, [0.1, 0.2, 0.3])
    assert np.array_equal(np.array(buf.scores[-1]), want["R"][0])
    np.testing.assert_allclose(np.array(buf.modified_scores[-1]), want["modR"][0], rtol=1e-13)
    assert np.array_equal(np.array(buf.constraints[-1]), want["C"][0])
    np.testing.assert_allclose(np.array(G).T, want["G"][:, 0], rtol=1e-13)
    meta.update()
    np.testing.assert_allclose(meta.lambdas, no.lambda_update([0.1, 0.2, 0.3], want["C"][0], [25.0] * 3, 0.05), rtol=1e-14)


def test_timing_helper_runs():
    rate, procs = sp.time_all_cores("coverage", 5, 3, 10, 0.99, 0.2, 2)
    assert rate > 0 and procs == 2
