"""smarl_policy_act_gaussian (the reference's per-agent ContinuousPolicy, agent.py:48-76, fused for all envs and agents)
against the PyTorch policy glue (mu / sigma^2 through the log-probability of the returned action) and against the numpy
restatement of its Philox Box-Muller sampling in oracle/philox.py."""
import numpy as np
import pytest
import torch

from oracle import philox

pytestmark = pytest.mark.gpu


def make(kind, A, E, seed=0, env_offset=0, shuffle=False, var_bias=1.0):
    import safe_multiagent_rl_b200 as s
    from safe_multiagent_rl_b200.policy import FusedGaussianPolicy
    rng = np.random.default_rng(seed)
    if kind == "collision":
        env = s.BatchedCollisionAvoidance(6, A, n_envs=E, n_landmarks=2, starts=rng.random((E, A, 2)) * 6,
                                          landmarks=rng.random((E, 2, 2)) * 6, shuffle=shuffle, env_offset=env_offset)
    else:
        env = s.BatchedCoverageContinuous(8, A, n_envs=E, starts=rng.random((E, A, 2)) * 8, env_offset=env_offset)
    torch.manual_seed(seed)
    pol = FusedGaussianPolicy(env, seed=91)
    with torch.no_grad():
        pol.fc2_.bias.add_(var_bias)          # sigma^2 away from the 1e-4 floor (the floor case has its own test)
    return env, pol


def heads(pol, obs):
    with torch.no_grad():
        h = torch.relu(pol.fc1(obs))
        return pol.fc2(h).double().cpu().numpy(), (torch.relu(pol.fc2_(h)) + 1e-4).double().cpu().numpy()   # [A, E, 2]


@pytest.mark.parametrize("kind,A,E,shuffle", [("collision", 3, 50, False), ("collision", 16, 1000, False),
                                              ("collision", 5, 4099, True), ("collision", 32, 70, False),
                                              ("coverage", 3, 333, False), ("coverage", 8, 2000, False),
                                              ("collision", 1, 20, False),
                                              # several tiles per persistent CTA
                                              ("collision", 16, 20000, False), ("coverage", 3, 160001, False),
                                              ("collision", 8, 40001, True)])
def test_fused_gaussian_policy_matches_torch_and_the_philox_oracle(kind, A, E, shuffle):
    env, pol = make(kind, A, E, env_offset=7_000_000_000, shuffle=shuffle)
    obs = env.reset()
    for t in (0, 1, 9):
        _, act, logp = pol.act(t=t)
        act_np, logp_np = act.cpu().numpy().astype(np.float64), logp.cpu().numpy().astype(np.float64)
        mu, var = heads(pol, obs)
        z_back = (act_np - mu) / np.sqrt(var)
        # float32 evaluations of mu and sigma^2 differ by ~1e-6 (1 + |.|) between the kernel and PyTorch (summation
        # order); in log N the mu difference is divided by sigma and multiplied by |z|, the sigma^2 difference enters
        # relative to sigma^2 (which can sit near its 1e-4 floor) times (z^2 + 1) / 2: the bound is that propagation
        tol = 2e-5 * (1.0 + np.abs(logp_np)) + (4e-6 * (1.0 + np.abs(mu)) * np.abs(z_back) / np.sqrt(var) +
                                                4e-6 * (1.0 + var) / var * 0.5 * (z_back ** 2 + 1.0)).sum(axis=-1)
        # (i) log N(action; mu, sigma^2) of the returned action, against the PyTorch glue on the same observation
        want_lp = pol.log_prob(obs, act).detach().cpu().numpy()
        assert np.all(np.abs(logp_np - want_lp) <= tol), np.abs(logp_np - want_lp).max()
        # (ii) the samples, against the float64 restatement driven by the same Philox normals
        z = philox.gaussian_normals(91, env.env_offset + np.arange(E), t, A).transpose(1, 0, 2)   # [A, E, 2]
        a_o, lp_o = philox.gaussian_sample(mu, var, z)
        tol_a = 2e-5 * (1.0 + np.abs(a_o)) + 4e-6 * (1.0 + var) / (2.0 * np.sqrt(var)) * np.abs(z)    # d sigma = d var / (2 sigma)
        bad = np.unravel_index(np.argmax(np.abs(act_np - a_o) / tol_a), tol_a.shape)
        assert np.all(np.abs(act_np - a_o) <= tol_a), (act_np[bad], a_o[bad], mu[bad], var[bad], z[bad], tol_a[bad])
        assert np.all(np.abs(logp_np - lp_o) <= 2 * tol), np.abs(logp_np - lp_o).max()
        # standard normals: the stream is not degenerate
        if A * E >= 1000:
            assert abs(z_back.mean()) < 0.1 and 0.9 < z_back.std() < 1.1
        obs, _, _, _ = env.step(env.action_buffer, agent_major=True)


def test_variance_floor_and_streams():
    """sigma^2 = relu(.) + 1e-4 at its floor (agent.py:67): finite log-probabilities, sigma = 1e-2; a shard draws what
    the full batch draws for its envs; another episode, other samples."""
    env, pol = make("collision", 4, 600, env_offset=1000, var_bias=-50.0)
    obs = env.reset()
    _, act, logp = pol.act(t=2)
    mu, var = heads(pol, obs)
    assert np.allclose(var, 1e-4)
    z = philox.gaussian_normals(91, env.env_offset + np.arange(600), 2, 4).transpose(1, 0, 2)
    np.testing.assert_allclose(act.cpu().numpy(), mu + 1e-2 * z, rtol=1e-5, atol=2e-6)
    assert torch.isfinite(logp).all()
    a_full = act.clone()
    sub_env, sub_pol = make("collision", 4, 200, env_offset=1000 + 150, var_bias=-50.0)
    sub_pol.load_state_dict(pol.state_dict())
    sub_env.obs[:, :200] = env.obs[:, 150:350]
    _, a_sub, _ = sub_pol.act(t=2)
    assert torch.equal(a_sub, a_full[:, 150:350])
    pol.next_episode()
    _, a_next, _ = pol.act(t=2)
    assert not torch.equal(a_next, a_full)


def test_closed_loop_in_a_cuda_graph():
    env, pol = make("collision", 3, 500)
    lam = torch.zeros(1, dtype=torch.float64, device="cuda")
    T = 10
    buf = env.new_rollout_buffer(T)

    def loop():
        env.reset()
        for t in range(T):
            pol.act(t=t)
            env.step(env.action_buffer, lambdas=lam, out=(buf, t), agent_major=True)
        return buf.finish(0.9, [1.5])
    eager = {k: v.clone() for k, v in loop().items() if isinstance(v, torch.Tensor)}
    pos = env.state().clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loop()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.state(), pos)
    assert torch.equal(buf.R[:, :500].t(), eager["R"])
