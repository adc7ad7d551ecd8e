"""shuffle=True: per-episode re-randomised starts / landmarks drawn on the device (Philox, keyed by
the global env id and the episode index) against oracle/philox.random_starts, for every env, for
reset() and for the fused rollout, and independent of sharding."""
import numpy as np
import pytest
import torch

from oracle import philox

pytestmark = pytest.mark.gpu


def make(kind, E, seed, offset=0):
    import safe_multiagent_rl_b200 as s
    rng = np.random.default_rng(0)
    if kind == "coverage":
        return s.BatchedCoverageDiscrete(7, 5, n_envs=E, shuffle=True, seed=seed, env_offset=offset), 0, None
    if kind == "congestion":
        return s.BatchedCongestion(6, 4, n_envs=E, shuffle=True, seed=seed, env_offset=offset, noise=0.0,
                                   demand_rate=rng.random((7, 7)) * 8 + 2), 1, None
    if kind == "continuous":
        return s.BatchedCoverageContinuous(5, 3, n_envs=E, shuffle=True, seed=seed, env_offset=offset), 2, None
    if kind == "discretized":
        return s.BatchedCoverageDiscretized(5, 3, n_envs=E, coarseness=6, shuffle=True, seed=seed, env_offset=offset), 3, 6 / 5
    return s.BatchedCollisionAvoidance(5, 3, n_envs=E, n_landmarks=2, shuffle=True, seed=seed, env_offset=offset), 2, None


@pytest.mark.parametrize("kind", ["coverage", "congestion", "continuous", "discretized", "collision"])
def test_reset_draws_match_oracle_and_sharding(kind):
    E, seed = 333, 99
    env, code, zoom = make(kind, E, seed, offset=1000)
    ids = np.arange(1000, 1000 + E)
    for episode in range(3):
        obs = env.reset()
        want = philox.random_starts(code, env.size, seed, ids, episode, env.n_agents, zoom=zoom)
        assert np.array_equal(env.state().cpu().numpy().astype(np.float64), want), episode
        got_obs = obs.cpu().numpy()[:, : 2 * env.n_agents].reshape(E, env.n_agents, 2)
        assert np.array_equal(got_obs, want.astype(np.float32))
        if kind == "collision":
            lm = philox.random_starts(2, env.size, seed, ids, episode, 2, row_offset=env.n_agents)
            assert np.array_equal(env.landmarks[:, :E].t().cpu().numpy().reshape(E, 2, 2), lm)
            assert obs.shape[1] == env.state_space == 2 * 3 + 2 * 2               # landmarks are part of the state
            assert np.array_equal(obs.cpu().numpy()[:, 6:].reshape(E, 2, 2), lm.astype(np.float32))
        if kind in ("coverage", "congestion"):
            assert want.max() <= env.size - 1 and want.min() >= 0
    # a shard of the same global range draws the same starts
    part, _, _ = make(kind, 100, seed, offset=1100)
    part.reset(); part.reset(); part.reset()
    assert np.array_equal(part.state().cpu().numpy(), env.state().cpu().numpy()[100:200])


def test_fused_rollout_uses_fresh_starts():
    import safe_multiagent_rl_b200 as s
    E, A, T, seed = 200, 5, 12, 5
    rng = np.random.default_rng(1)
    env = s.BatchedCoverageDiscrete(7, A, n_envs=E, shuffle=True, seed=seed)
    actions = torch.as_tensor(rng.integers(0, 5, (T, A, env.ld)).astype(np.uint8), device="cuda")
    for episode in range(2):
        out = env.rollout(actions, gamma=0.99)
        starts = philox.random_starts(0, 7, seed, np.arange(E), episode, A)
        fixed = s.BatchedCoverageDiscrete(7, A, n_envs=E, starts=starts)
        ref = fixed.rollout(actions, gamma=0.99)
        assert torch.equal(out["R"], ref["R"]) and torch.equal(out["C"], ref["C"])
        assert torch.equal(env.state(), fixed.state())
