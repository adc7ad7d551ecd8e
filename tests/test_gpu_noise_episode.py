"""Congestion action noise must differ from episode to episode (the reference draws fresh random() values every
step, envs/congestion.py:64-67): the Philox counter carries an episode index that every reset() / rollout()
after the first advances on the device, also inside a replayed CUDA graph."""
import numpy as np
import pytest
import torch

from oracle import numpy_oracle as no
from oracle import philox

pytestmark = pytest.mark.gpu


def make(E=200, A=8, size=6, noise=0.3, seed=11, env_offset=1000):
    import safe_multiagent_rl_b200 as s
    rng = np.random.default_rng(0)
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    starts[:, 0] = 0
    env = s.BatchedCongestion(size, A, n_envs=E, noise=noise, starts=starts, demand_rate=demand, seed=seed,
                              env_offset=env_offset)
    return s, env, starts, demand, rng


def expected_moves(env, actions, episode, t):
    E, A = actions.shape
    w = philox.congestion_words(env.seed, env.env_offset + np.arange(E), t, A, episode)
    keep = w.astype(np.uint64) < philox.keep_threshold(env.noise)
    return np.where(keep, actions, w % np.uint32(5)).astype(np.int64)


def test_consecutive_episodes_draw_fresh_noise_and_match_the_oracle_stream():
    s, env, starts, demand, rng = make()
    E, A, T = env.n_envs, env.n_agents, 5
    actions = rng.integers(0, 5, size=(T, E, A))
    per_episode = []
    for episode in range(3):
        env.reset()
        moves = []
        for t in range(T):
            env.step(actions[t].astype(np.uint8))
            got = env.moves[:, :E].t().cpu().numpy().astype(np.int64)
            assert np.array_equal(got, expected_moves(env, actions[t], episode, t)), (episode, t)
            moves.append(got)
        per_episode.append(np.stack(moves))
    assert not np.array_equal(per_episode[0], per_episode[1])
    assert not np.array_equal(per_episode[1], per_episode[2])
    env.set_noise_episode(1)                         # pinning the index replays that episode's realisation
    env.reset()
    env.step(actions[0].astype(np.uint8))
    assert np.array_equal(env.moves[:, :E].t().cpu().numpy(), per_episode[1][0])


def test_fused_rollout_and_step_api_share_the_episode_stream():
    s, env, starts, demand, rng = make(E=96, A=5)
    E, A, T = env.n_envs, env.n_agents, 7
    actions = rng.integers(0, 5, size=(T, E, A))
    act_k = torch.zeros(T, A, env.ld, dtype=torch.uint8, device="cuda")
    act_k[:, :, :E] = torch.as_tensor(actions.astype(np.uint8), device="cuda").permute(0, 2, 1)
    finals = []
    for episode in range(2):                         # episodes 0, 1 through the fused kernel
        env.rollout(act_k, gamma=0.9)
        finals.append(env.state().cpu().numpy())
        pos = starts.copy()
        for t in range(T):
            pos, _, _, _, _ = no.congestion_step(pos, actions[t], expected_moves(env, actions[t], episode, t), env.size, demand)
        assert np.array_equal(finals[-1], pos), episode
    assert not np.array_equal(finals[0], finals[1])
    env.reset()                                      # episode 2 through the step API
    pos = starts.copy()
    for t in range(T):
        env.step(actions[t].astype(np.uint8))
        pos, _, _, _, _ = no.congestion_step(pos, actions[t], expected_moves(env, actions[t], 2, t), env.size, demand)
    assert np.array_equal(env.state().cpu().numpy(), pos)


def test_graph_replays_advance_the_episode():
    s, env, starts, demand, rng = make(E=64, A=4)
    E, A, T = env.n_envs, env.n_agents, 6
    actions = torch.as_tensor(rng.integers(0, 5, size=(T, A, env.ld)).astype(np.uint8), device="cuda")

    def policy(obs, t):
        env.action_buffer.copy_(actions[t])
    lam = torch.zeros(1, dtype=torch.float64, device="cuda")
    loop = s.GraphedClosedLoop(env, T, policy, lam, 0.9, thresholds=[1.5])
    seen = []
    for _ in range(3):
        loop.replay()
        torch.cuda.synchronize()
        seen.append(env.state().cpu().numpy().copy())
    assert not np.array_equal(seen[0], seen[1]) and not np.array_equal(seen[1], seen[2])
