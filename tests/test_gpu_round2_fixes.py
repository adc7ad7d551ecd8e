"""Round-2 regression tests for the advisor's findings: the lean (shared-reward) accounting honours early
episode ends under PPO standardisation; fused-rollout output dicts are re-validated on reuse; thresholds
must match the constraint count; out-of-range discrete actions are rejected on the checked ingest path;
the penalty-table limit holds for every kernel; MetaAgent.act records constraints under the gate."""
import numpy as np
import pytest
import torch

import test_gpu_parity as tp
from oracle import numpy_oracle as no

pytestmark = pytest.mark.gpu


def test_collision_lean_buffer_ppo_standardised_respects_episode_ends():
    s = tp.smarl()
    size, A, L, E, T, seed = tp.COLLISION_CASES[0]
    starts, landmarks, actions = tp.collision_setup(size, A, L, E, T, seed)
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    n_active = np.zeros(E, dtype=np.int64)

    def step_fn(t):
        nonlocal pos, done, n_active
        pos, r, c, done, active = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        n_active += active
        return r.astype(np.float32).astype(np.float64), c
    want = no.rollout(step_fn, T, 0.99, [0.5])
    act = torch.as_tensor(actions, device="cuda")
    lean = env.new_rollout_buffer(T, g_mode=s.G_PPO_STANDARDISED, lean=True)
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, torch.as_tensor([0.5], dtype=torch.float64, device="cuda"),
                                  0.99, buffer=lean, g_mode=s.G_PPO_STANDARDISED)
    assert (n_active < T).any()
    ok = n_active >= 2
    got = out["G"].cpu().numpy()
    ref = no.ppo_standardised_returns(want["mod_reward"], 0.99, n_active)
    tp.close(got[:, ok], ref[:, ok], 1.0, rtol=2e-5)
    short = np.flatnonzero((n_active < T) & ok)
    for e in short[:20]:                                     # steps past the episode end carry no return
        assert (got[n_active[e]:, e] == 0).all()
    assert np.isnan(got[0, n_active == 1]).all()


def test_rollout_output_dict_is_revalidated_on_reuse():
    s = tp.smarl()
    size, A, E = 6, 4, 100
    rng = np.random.default_rng(0)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, starts=starts)

    def acts(T):
        a = torch.zeros(T, A, env.ld, dtype=torch.uint8, device="cuda")
        a[:, :, :E] = torch.as_tensor(rng.integers(0, 5, size=(T, A, E)).astype(np.uint8), device="cuda")
        return a
    out = {}
    env.rollout(acts(5), gamma=0.9, g_mode=s.G_REWARD_TO_GO, out=out)
    small = out["G_buf"]
    assert tuple(small.shape) == (5, A, env.ld)
    env.rollout(acts(40), gamma=0.9, g_mode=s.G_REWARD_TO_GO, out=out)     # longer horizon: must not reuse the 5-step slabs
    assert tuple(out["G_buf"].shape) == (40, A, env.ld) and out["G_buf"].data_ptr() != small.data_ptr()
    assert tuple(out["g_scratch"].shape) == (80, 1, env.ld)
    big = out["G_buf"]
    env.rollout(acts(40), gamma=0.9, g_mode=s.G_REWARD_TO_GO, out=out)     # same shape: reused
    assert out["G_buf"].data_ptr() == big.data_ptr()


def test_thresholds_must_match_the_constraint_count():
    s = tp.smarl()
    env = s.BatchedCoverageDiscrete(5, 3, n_envs=20, starts=np.zeros((20, 3, 2), dtype=np.int64))
    act = torch.zeros(4, 3, env.ld, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        env.rollout(act, gamma=0.9, thresholds=[2])          # the reference's CLI default, K = 3 here
    buf = env.new_rollout_buffer(4)
    env.reset()
    for t in range(4):
        env.step(act[t], out=(buf, t), agent_major=True)
    with pytest.raises(ValueError):
        buf.finish(0.9, thresholds=[2.0, 2.0])
    buf.finish(0.9, thresholds=[2.0, 2.0, 2.0])


def test_out_of_range_discrete_actions_are_rejected():
    s = tp.smarl()
    env = s.BatchedCoverageDiscrete(5, 3, n_envs=8, starts=np.zeros((8, 3, 2), dtype=np.int64))
    env.reset()
    bad = np.zeros((8, 3), dtype=np.int64)
    bad[3, 1] = 5                                            # the reference's direction table has 5 entries: IndexError
    with pytest.raises(IndexError):
        env.step(bad)
    cenv = s.BatchedCongestion(3, 3, n_envs=8, noise=0.0, starts=np.zeros((8, 3, 2), dtype=np.int64))
    cenv.reset()
    with pytest.raises(IndexError):
        cenv.step(bad)
    env.step(np.full((8, 3), 4))                             # in range: fine


def test_penalty_table_limit_is_the_same_for_step_and_rollout():
    """A field of view whose table fills shared memory to the limit works in BOTH kernels; one entry more raises."""
    s = tp.smarl()
    from safe_multiagent_rl_b200.envs.coverage import MAX_LUT, penalty_table
    size, A, E, T = 127, 2, 40, 3
    fv_ok = np.sqrt(MAX_LUT - 0.5)                           # table length MAX_LUT
    assert len(penalty_table(size, A, fv_ok)[1]) == MAX_LUT
    rng = np.random.default_rng(0)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, fieldview_size=fv_ok, starts=starts)
    actions = rng.integers(0, 5, size=(T, E, A))
    lut = no.coverage_penalty_lut(size, fv_ok)
    env.reset()
    pos = starts.copy()
    for t in range(T):
        _, r, _, _ = env.step(actions[t].astype(np.uint8))
        pos, r_o, _, _ = no.coverage_discrete_step(pos, actions[t], size, lut, None)
        tp.close(r.cpu().numpy(), r_o)
    act_k = tp.kernel_layout(actions.astype(np.uint8), env.ld)
    out = env.rollout(act_k, gamma=0.9)                      # the fused kernel adds a static reduction buffer
    torch.cuda.synchronize()
    assert np.array_equal(env.state().cpu().numpy(), pos)
    with pytest.raises(NotImplementedError):
        s.BatchedCoverageDiscrete(size, A, n_envs=E, fieldview_size=np.sqrt(MAX_LUT + 0.5), starts=starts)


def test_meta_agent_act_records_under_the_gate_like_the_reference():
    s = tp.smarl()
    K, A, E, T = 2, 2, 6, 3
    meta = s.BatchedMetaAgent([1, 1], 0.9, 0.5, [1.0, 1.0], start_learning_cycle=1, lambda_0=0.5, n_agents=A)
    rng = np.random.default_rng(0)
    costs = rng.integers(0, 2, size=(T, E, K)).astype(np.float64)
    rewards = -rng.random((T, E, A))

    def episode():
        for t in range(T):
            m = meta.act(torch.as_tensor(costs[t], device="cuda"), torch.as_tensor(rewards[t], device="cuda"))
            want = rewards[t] - (costs[t] @ np.array([0.5, 0.5]))[:, None]
            np.testing.assert_allclose(m.cpu().numpy(), want, rtol=1e-12)
        meta.step()
    episode()                                                # gate closed (learning_cycle 0 < 1): nothing recorded
    meta.update()
    assert np.array_equal(meta.lambdas.cpu().numpy(), [0.5, 0.5])
    meta.increment_learning_cycle()
    episode()                                                # gate open: recorded, meta_agent.py:18-20,25-30
    lam0 = meta.lambdas.cpu().numpy().copy()
    meta.update()
    mean_c = costs.sum(0).mean(0)                            # mean over the E recorded episodes of sum_t c
    np.testing.assert_allclose(meta.lambdas.cpu().numpy(), np.maximum(lam0 + 0.5 * (mean_c - 1.0), 0), rtol=1e-12)


@pytest.mark.parametrize("size,coarse", [(5, 6), (3, 2), (33, 6), (7, 10), (10, 3), (64, 25), (100, 7), (127, 50), (2, 1), (13, 12)])
def test_discretized_positions_stay_bit_exact_with_the_invariant_divisor(size, coarse):
    """CoverageDiscretized divides by the kernel-invariant zoom = coarseness / size with a multiply and two exact-residual
    FMA corrections instead of a float64 division (coverage_float.cu, div_by_zoom).  Positions must stay bit-identical to
    numpy's float64 x / zoom over long trajectories, for zooms with terminating and non-terminating binary expansions
    (3/2, 6/33, 25/64, ...), many envs, all nine actions incl. walks into both walls."""
    import numpy as np
    import torch
    import safe_multiagent_rl_b200 as s
    from oracle import numpy_oracle as no
    rng = np.random.default_rng(size * 1000 + coarse)
    A, E, T = 3, 5000, 60
    zoom = coarse / size
    starts = np.floor(rng.random((E, A, 2)) * size * zoom) / zoom
    env = s.BatchedCoverageDiscretized(size, A, n_envs=E, coarseness=coarse, starts=starts)
    fv = no.coverage_fieldview(size, A)
    env.reset()
    pos = starts.copy()
    for t in range(T):
        # biased walks so that agents run into the walls and bounce along them
        act = np.where(rng.random((E, A)) < 0.7, rng.integers(0, 9, (E, A)), (t // 15) % 8)
        env.step(act.astype(np.uint8))
        pos, _, _, _ = no.coverage_discretized_step(pos, act, size, coarse, fv, None, exact_pow=False)
        got = np.ascontiguousarray(env.state().cpu().numpy())
        assert np.array_equal(got.view(np.int64), np.ascontiguousarray(pos).view(np.int64)), (t, np.abs(got - pos).max())
