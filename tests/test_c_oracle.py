"""The plain-C oracle against the numpy oracle (which is pinned to the reference), bit-exact
wherever both use the same summation order."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import numpy_oracle as no
from oracle import philox


def am(a):
    """[E, rows] -> agent-major [rows, ld=E]; [T, E, rows] -> [T, rows, E]."""
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


@pytest.mark.parametrize("size,A,E,T,seed", [(5, 3, 50, 50, 0), (32, 16, 64, 20, 1), (64, 32, 16, 5, 2), (7, 1, 9, 4, 3)])
def test_coverage_c_vs_numpy(size, A, E, T, seed):
    rng = np.random.default_rng(seed)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    actions = rng.integers(0, 5, size=(T, E, A))
    w = (1.0 + np.arange(A) % 3)
    lam = np.linspace(0.1, 0.5, A)
    lut = no.coverage_penalty_lut(size, no.coverage_fieldview(size, A))
    pos = starts.copy()
    last = {}

    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, actions[t], size, lut, w)
        last["r"] = r
        return r, c
    want = no.rollout(step_fn, T, 0.999, lam)
    got = co.coverage_rollout(size, am(starts[:, :, 0]), am(starts[:, :, 1]), am(actions), lut, w, lam, 0.999, E,
                              want_G=True)
    assert np.array_equal(got["final_x"].T, pos[:, :, 0]) and np.array_equal(got["final_y"].T, pos[:, :, 1])
    assert np.array_equal(got["C"].T, want["C"])
    assert np.array_equal(got["reward_last"].T, last["r"])                 # bit-exact f64 rewards
    assert np.array_equal(got["R"].T, want["R"])                          # same left-to-right sum
    np.testing.assert_allclose(got["modR"].T, want["modR"], rtol=1e-13)
    np.testing.assert_allclose(np.swapaxes(got["G"], 1, 2), want["G"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("size,A,E,T,noise,seed", [(3, 3, 40, 10, 0.0, 0), (10, 8, 33, 60, 0.1, 1), (2, 7, 25, 30, 0.5, 2),
                                                    (5, 32, 9, 6, 0.3, 3)])
def test_congestion_c_vs_numpy(size, A, E, T, noise, seed):
    rng = np.random.default_rng(seed)
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    starts[:, 0] = 0
    actions = rng.integers(0, 5, size=(T, E, A))
    ids = np.arange(123, 123 + E)
    pos = starts.copy()

    def step_fn(t):
        nonlocal pos
        u1, u2 = philox.congestion_uniforms(77, ids, t, A)
        mv = no.congestion_noise_moves(actions[t], u1, u2, noise)
        pos, r, c, _, _ = no.congestion_step(pos, actions[t], mv, size, demand)
        return r, c
    want = no.rollout(step_fn, T, 0.9, [0.35])
    got = co.congestion_rollout(size, am(starts[:, :, 0]), am(starts[:, :, 1]), am(actions), demand, [0.35], 0.9, E,
                                noise_mode=2 if noise > 0 else 0, keep_threshold=philox.keep_threshold(noise), seed=77,
                                env_offset=123, round_f32=False, want_G=True)
    assert np.array_equal(got["final_x"].T, pos[:, :, 0]) and np.array_equal(got["final_y"].T, pos[:, :, 1])
    assert np.array_equal(got["C"].T, want["C"])
    assert np.array_equal(got["R"].T, want["R"])
    np.testing.assert_allclose(got["modR"].T, want["modR"], rtol=1e-13)
    np.testing.assert_allclose(np.swapaxes(got["G"], 1, 2), want["G"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("size,A,L,E,T,seed", [(5, 3, 1, 80, 50, 0), (2, 5, 1, 40, 20, 1), (5, 9, 3, 30, 20, 2), (4, 32, 1, 8, 5, 3)])
def test_collision_c_vs_numpy(size, A, L, E, T, seed):
    rng = np.random.default_rng(seed)
    starts = rng.random((E, A, 2)) * size
    lm = rng.random((E, L, 2)) * size
    actions = rng.normal(0, 0.5, size=(T, E, A, 2)).astype(np.float32)
    q = E // 4
    actions[:, :q] = ((lm[:q, :1] - starts[:q]) / 6).astype(np.float32)[None]
    starts[q:2 * q] = np.clip(lm[q:2 * q, :1] + rng.normal(0, 0.4, size=(q, A, 2)), 0, size)
    pos, done = starts.copy(), np.zeros((E, A), bool)
    n_active = np.zeros(E, np.int64)

    def step_fn(t):
        nonlocal pos, done, n_active
        pos, r, c, done, active = no.collision_step(pos, done, actions[t].astype(np.float64), lm, size)
        n_active += active
        return r, c
    want = no.rollout(step_fn, T, 0.99, [0.5])
    got = co.collision_rollout(size, am(starts[:, :, 0]), am(starts[:, :, 1]), am(lm.reshape(E, 2 * L)),
                               am(actions.reshape(T, E, 2 * A)), [0.5], 0.99, E, round_f32=False, want_G=True)
    assert np.array_equal(got["final_x"].T, pos[:, :, 0]) and np.array_equal(got["final_y"].T, pos[:, :, 1])
    assert np.array_equal(got["final_done"].T.astype(bool), done) and np.array_equal(got["n_active"], n_active)
    assert np.array_equal(got["C"].T, want["C"].astype(np.int64))
    assert np.array_equal(got["R"].T, want["R"])
    np.testing.assert_allclose(got["modR"].T, want["modR"], rtol=1e-13)
    np.testing.assert_allclose(np.swapaxes(got["G"], 1, 2), want["G"], rtol=1e-12, atol=1e-12)
    assert want["C"].sum() > 0 and (A > 12 or (n_active < T).any())
