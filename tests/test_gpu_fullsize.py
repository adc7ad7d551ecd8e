"""BASELINE.json's full-size configs on the GPU, checked (a) against the plain-C oracle on a large
contiguous slice of the envs and (b) through size-independent properties on all of them:
closed-loop == fused rollout, additive stats == sums of the per-env products, integer
dynamics identical in both modes."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import numpy_oracle as no
from oracle import philox

pytestmark = pytest.mark.gpu


def close(got, want, scale, rtol=1e-5):
    np.testing.assert_allclose(np.asarray(got, np.float64), want, rtol=rtol, atol=rtol * scale)


def test_config4_coverage_32x32_16_agents_4M_envs():
    """configs[3]: CoverageDiscrete 32x32, 16 agents, 2^22 envs, T=50."""
    import safe_multiagent_rl_b200 as s
    S, A, E, T, gamma = 32, 16, 1 << 22, 50, 0.999
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    weights = [1.0 + (i % 3) for i in range(A)]
    env = s.BatchedCoverageDiscrete(S, A, n_envs=E, weights=weights, starts=np.zeros((E, A, 2), np.uint8))
    env.start_x[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
    env.start_y[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
    actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
    lam_np, thr = np.linspace(0.05, 0.4, A), np.full(A, 39.0)
    lam = torch.as_tensor(lam_np, device="cuda")

    out = env.rollout_closed_loop(lambda obs, t: actions[t], T, lam, gamma, thresholds=thr)
    pos_closed = (env.pos_x.clone(), env.pos_y.clone())
    fused = env.rollout(actions, lambdas=lam, gamma=gamma, thresholds=thr, g_mode=1)
    # (b) properties over all 4M envs
    assert torch.equal(env.pos_x, pos_closed[0]) and torch.equal(env.pos_y, pos_closed[1])
    assert torch.equal(fused["C"], out["C"])
    scale = float(out["modR"].abs().max())
    assert float((fused["R"] - out["R"]).abs().max()) <= 2e-6 * scale
    assert float((fused["modR"] - out["modR"]).abs().max()) <= 2e-6 * scale
    assert float((fused["G"] - out["G"]).abs().max()) <= 2e-6 * scale
    for o in (out, fused):
        st = o["stats"]
        assert torch.equal(st.cost_sum, o["C"].sum(0).double())                       # exact integers
        assert torch.equal(st.violations, (o["C"].double() > torch.as_tensor(thr, device="cuda")).sum(0).double())
        assert float(st.count) == E
        rs = o["R"].double().sum(0)
        assert float(((st.return_sum - rs) / rs).abs().max()) < 1e-6
    # cost sums are bounded by T and the stay probability is 1/5
    assert 0 <= int(out["C"].min()) and int(out["C"].max()) <= T
    assert abs(float(out["C"].double().mean()) - 0.8 * T) < 0.05
    # (a) C oracle on the first 2^18 envs
    n = 1 << 18
    lut = no.coverage_penalty_lut(S, no.coverage_fieldview(S, A))
    ref = co.coverage_rollout(S, env.start_x[:, :n].cpu().numpy(), env.start_y[:, :n].cpu().numpy(),
                              actions[:, :, :n].cpu().numpy(), lut, np.asarray(weights), lam_np, gamma, n, want_G=False)
    assert np.array_equal(env.pos_x[:, :n].cpu().numpy(), ref["final_x"])             # bit-exact
    assert np.array_equal(env.pos_y[:, :n].cpu().numpy(), ref["final_y"])
    for o in (out, fused):
        assert np.array_equal(o["C"][:n].t().cpu().numpy(), ref["C"])
        close(o["R"][:n].t().cpu().numpy(), ref["R"], scale)
        close(o["modR"][:n].t().cpu().numpy(), ref["modR"], scale)
    # last step's per-agent rewards from the rollout buffer
    close(out["buffer"].reward[T - 1][:, :n].cpu().numpy(), ref["reward_last"], np.abs(ref["reward_last"]).max())


def test_config3_congestion_10x10_8_agents_1M_envs():
    """configs[2]: Congestion size 10, 8 agents, 2^20 envs, T=100, noise 0.1 (on-device Philox)."""
    import safe_multiagent_rl_b200 as s
    S, A, E, T, gamma, noise, seed, off = 10, 8, 1 << 20, 100, 0.9, 0.1, 2024, 3 << 20
    rng = np.random.default_rng(3)
    demand = rng.random((S + 1, S + 1)) * 8 + 2
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    env = s.BatchedCongestion(S, A, n_envs=E, noise=noise, starts=np.zeros((E, A, 2), np.uint8), demand_rate=demand,
                              seed=seed, env_offset=off)
    env.start_x[1:, :E] = torch.randint(0, S, (A - 1, E), generator=g, device="cuda", dtype=torch.uint8)
    env.start_y[1:, :E] = torch.randint(0, S, (A - 1, E), generator=g, device="cuda", dtype=torch.uint8)
    actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
    lam = torch.as_tensor([0.35], dtype=torch.float64, device="cuda")
    thr = [150.0]
    out = env.rollout_closed_loop(lambda obs, t: actions[t], T, lam, gamma, thresholds=thr)
    pos_closed = (env.pos_x.clone(), env.pos_y.clone())
    env.set_noise_episode(0)            # replay episode 0's noise: a second episode would draw a fresh realisation
    fused = env.rollout(actions, lambdas=lam, gamma=gamma, thresholds=thr, g_mode=1)
    assert torch.equal(env.pos_x, pos_closed[0]) and torch.equal(env.pos_y, pos_closed[1])   # same Philox stream
    assert torch.equal(fused["C"], out["C"])
    scale = float(out["modR"].abs().max())
    assert float((fused["R"] - out["R"]).abs().max()) <= 2e-6 * scale
    assert float((fused["G"] - out["G"]).abs().max()) <= 2e-6 * scale
    for o in (out, fused):
        assert torch.equal(o["stats"].cost_sum, o["C"].sum(0).double()) and float(o["stats"].count) == E
    n = 1 << 18
    ref = co.congestion_rollout(S, env.start_x[:, :n].cpu().numpy(), env.start_y[:, :n].cpu().numpy(),
                                actions[:, :, :n].cpu().numpy(), demand, [0.35], gamma, n, noise_mode=2,
                                keep_threshold=philox.keep_threshold(noise), seed=seed, env_offset=off, round_f32=True)
    assert np.array_equal(env.pos_x[:, :n].cpu().numpy(), ref["final_x"])
    assert np.array_equal(env.pos_y[:, :n].cpu().numpy(), ref["final_y"])
    for o in (out, fused):
        assert np.array_equal(o["C"][:n].t().cpu().numpy(), ref["C"])
        close(o["R"][:n].t().cpu().numpy(), ref["R"], scale)
        close(o["modR"][:n].t().cpu().numpy(), ref["modR"], scale)


@pytest.mark.parametrize("A,E,T", [(3, 65536, 50),          # configs[1]
                                   (3, 1 << 18, 20),        # large batch: the register-capped A <= 4 kernels
                                   (4, (1 << 18) + 40, 12)])
def test_config2_collision_5x5_3_agents_65536_envs(A, E, T):
    """configs[1]: CollisionAvoidance size 5, 3 agents, 65,536 envs, T=50 (and the kernels large batches select)."""
    import safe_multiagent_rl_b200 as s
    S, L, gamma = 5, 1, 0.99
    rng = np.random.default_rng(5)
    starts = rng.random((E, A, 2)) * S
    lm = rng.random((E, L, 2)) * S
    actions = rng.normal(0, 0.5, size=(T, E, A, 2)).astype(np.float32)
    q = E // 4
    actions[:, :q] = ((lm[:q, :1] - starts[:q]) / 6).astype(np.float32)[None]
    env = s.BatchedCollisionAvoidance(S, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=lm)
    lam = torch.as_tensor([0.5], dtype=torch.float64, device="cuda")
    act = torch.as_tensor(actions, device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, lam, gamma, thresholds=[1.0])
    pos_closed = env.state().clone()
    act_k = torch.zeros(T, 2 * A, env.ld, device="cuda")
    act_k[:, :, :E] = act.reshape(T, E, 2 * A).permute(0, 2, 1)
    fused = env.rollout(act_k, lambdas=lam, gamma=gamma, thresholds=[1.0], g_mode=1)
    assert torch.equal(env.state(), pos_closed)
    assert torch.equal(fused["C"], out["C"])
    scale = float(out["modR"].abs().max())
    assert float((fused["modR"] - out["modR"]).abs().max()) <= 2e-6 * scale
    assert float((fused["G"] - out["G"]).abs().max()) <= 2e-6 * scale
    ref = co.collision_rollout(S, env.start_x[:, :E].cpu().numpy(), env.start_y[:, :E].cpu().numpy(),
                               env.landmarks[:, :E].cpu().numpy(), act_k[:, :, :E].cpu().numpy(), [0.5], gamma, E)
    assert np.array_equal(env.pos_x[:, :E].cpu().numpy(), ref["final_x"])               # bit-exact float64
    assert np.array_equal(env.pos_y[:, :E].cpu().numpy(), ref["final_y"])
    assert np.array_equal(env.agent_done[:, :E].cpu().numpy(), ref["final_done"])
    assert np.array_equal(fused["n_active"].cpu().numpy(), ref["n_active"][:E])
    assert (ref["n_active"][:E] < T).sum() > 1000                                        # many episodes end early
    for o in (out, fused):
        assert np.array_equal(o["C"].t().cpu().numpy(), ref["C"][:, :E])
        close(o["R"].t().cpu().numpy(), ref["R"][:, :E], scale)
        close(o["modR"].t().cpu().numpy(), ref["modR"][:, :E], scale)
