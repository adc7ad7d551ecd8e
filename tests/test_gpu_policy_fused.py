"""smarl_policy_act_discrete (the reference's per-agent DiscretePolicy, agent.py:23-47, fused for all envs and agents)
against the PyTorch policy glue (log-probabilities <= 1e-5) and against the numpy restatement of its Philox
inverse-CDF sampling in oracle/philox.py (actions identical wherever the uniform is not within float32 rounding of a
CDF step)."""
import numpy as np
import pytest
import torch

from oracle import philox

pytestmark = pytest.mark.gpu


def make(A, E, size=9, seed=0, env_offset=0, kind="coverage"):
    import safe_multiagent_rl_b200 as s
    from safe_multiagent_rl_b200.policy import FusedDiscretePolicy
    rng = np.random.default_rng(seed)
    starts = rng.integers(0, size, (E, A, 2))
    if kind == "coverage":
        env = s.BatchedCoverageDiscrete(size, A, n_envs=E, starts=starts, env_offset=env_offset)
    else:
        starts[:, 0] = 0
        env = s.BatchedCongestion(size, A, n_envs=E, noise=0.0, starts=starts, env_offset=env_offset,
                                  demand_rate=rng.random((size + 1, size + 1)) * 8 + 2)
    torch.manual_seed(seed)
    pol = FusedDiscretePolicy(env, seed=77)
    with torch.no_grad():                                  # larger weights than the default init: peaked softmaxes too
        pol.fc2.weight.mul_(4.0)
    return env, pol


def reference_logits(pol, obs):
    with torch.no_grad():
        return pol.logits(obs.double().float()).double().cpu().numpy()          # [A, E, 5]


# builds of the kernel (smarl_set_kernel_variant): 0 = FP32 pipes (FFMA2), 1 / 2 = fc1 on the tensor cores
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("A,E", [(3, 50), (16, 1000), (8, 333), (32, 70), (1, 20), (5, 4099), (12, 260), (20, 129),
                                 (17, 128), (24, 500)])
def test_fused_policy_matches_torch_logprobs_and_the_philox_oracle(A, E, variant):
    from safe_multiagent_rl_b200 import _lib
    env, pol = make(A, E, env_offset=5_000_000_000)
    obs = env.reset()
    for t in (0, 1, 7):
        with _lib.kernel_variant(_lib.KERNEL_POLICY, variant):
            _, act, logp = pol.act(t=t)
        act_np, logp_np = act.cpu().numpy().astype(np.int64), logp.cpu().numpy()
        assert act_np.min() >= 0 and act_np.max() <= 4
        # (i) the log-probability of the sampled action, against the PyTorch glue on the float observation
        want_lp = pol.log_prob(obs, act.long()).detach().cpu().numpy()
        np.testing.assert_allclose(logp_np, want_lp, rtol=1e-5, atol=1e-5)
        # (ii) the samples, against the float64 restatement driven by the same Philox uniforms
        u = philox.policy_uniforms(77, env.env_offset + np.arange(E), t, A).T               # [A, E]
        a_o, lp_o, margin = philox.policy_sample(reference_logits(pol, obs), u)
        clear = margin > 1e-5
        assert clear.mean() > 0.99
        assert np.array_equal(act_np[clear], a_o[clear])
        np.testing.assert_allclose(logp_np[clear], lp_o[clear], rtol=1e-5, atol=1e-5)
        # every action is drawn with roughly its probability (the stream is not degenerate)
        assert len(np.unique(act_np)) >= (3 if A * E > 200 else 1)
        obs, _, _, _ = env.step(env.action_buffer, agent_major=True)


@pytest.mark.parametrize("A,E,size,kind", [(16, 70_001, 32, "coverage"), (8, 33_000, 200, "congestion"),
                                           (32, 9_000, 64, "coverage"), (3, 50_000, 5, "coverage")])
def test_tensor_core_and_fp32_builds_agree(A, E, size, kind):
    """The two builds draw from the same Philox stream and evaluate the same network: identical actions except where
    the uniform sits within float32 rounding of a CDF step, log-probabilities within 2e-5 x the scale of the
    observations (fc1 on the tensor cores is f32-grade: exact bf16 positions x three bf16 pieces per weight, f32
    accumulation; only the order of the f32 additions differs)."""
    from safe_multiagent_rl_b200 import _lib
    env, pol = make(A, E, size=size, env_offset=123, kind=kind)
    env.reset()
    out = {}
    for variant in (0, 1, 2):
        with _lib.kernel_variant(_lib.KERNEL_POLICY, variant):
            _, act, logp = pol.act(t=5)
            out[variant] = (act.clone(), logp.clone())
    for variant in (1, 2):
        same = (out[variant][0] == out[0][0])
        assert same.float().mean().item() > 0.9995
        d = (out[variant][1] - out[0][1]).abs()[same]
        assert d.max().item() < 2e-5 * max(1.0, size / 16), d.max().item()
    assert torch.equal(out[1][0], out[2][0]) or (out[1][0] == out[2][0]).float().mean().item() > 0.9999


def test_streams_follow_global_env_ids_and_episodes():
    env, pol = make(4, 600, env_offset=1000)
    env.reset()
    _, a_full, _ = pol.act(t=3)
    a_full = a_full.clone()
    env2, pol2 = make(4, 600, env_offset=1000)            # same weights (same torch seed), same envs
    pol2.load_state_dict(pol.state_dict())
    env2.reset()
    sub_env, sub_pol = make(4, 200, env_offset=1000 + 150)
    sub_pol.load_state_dict(pol.state_dict())
    sub_env.start_x.copy_(torch.zeros_like(sub_env.start_x)); sub_env.start_y.copy_(torch.zeros_like(sub_env.start_y))
    sub_env.start_x[:, :200] = env.start_x[:, 150:350]; sub_env.start_y[:, :200] = env.start_y[:, 150:350]
    sub_env.reset()
    _, a_sub, _ = sub_pol.act(t=3)
    assert torch.equal(a_sub, a_full[:, 150:350])         # a shard draws what the full batch draws for its envs
    pol.next_episode()
    _, a_next, _ = pol.act(t=3)
    assert not torch.equal(a_next, a_full)                # another episode, other samples


def test_closed_loop_without_float_observations_and_in_a_cuda_graph():
    import safe_multiagent_rl_b200 as s
    env, pol = make(8, 500, kind="congestion")
    env.emit_obs = False                                  # the policy reads the u8 position rows
    lam = torch.zeros(1, dtype=torch.float64, device="cuda")
    T = 12
    buf = env.new_rollout_buffer(T)

    def loop():
        env.reset()
        for t in range(T):
            pol.act(t=t)
            obs, _, _, _ = env.step(env.action_buffer, lambdas=lam, out=(buf, t), agent_major=True)
            assert obs is None
        return buf.finish(0.9, [1.5])
    eager = {k: v.clone() for k, v in loop().items() if isinstance(v, torch.Tensor)}
    pos = env.state().clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loop()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.state(), pos)                  # same episode index, same samples, same trajectory
    assert torch.equal(buf.R[:, :500].t(), eager["R"])
