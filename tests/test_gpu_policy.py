"""Batched policy glue: actions sampled on the device land in env.action_buffer (kernel layout) and
drive env.step zero-copy; a short constrained training run lowers the constraint cost."""
import subprocess
import sys
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_discrete_policy_writes_kernel_layout():
    import safe_multiagent_rl_b200 as s
    from safe_multiagent_rl_b200.policy import BatchedDiscretePolicy
    torch.manual_seed(0)
    env = s.BatchedCoverageDiscrete(5, 3, n_envs=100, starts=np.random.default_rng(0).integers(0, 5, (100, 3, 2)))
    pol = BatchedDiscretePolicy(env)
    obs = env.reset().clone()            # reset()/step() return views of env.obs, which the next step overwrites
    buf, a, lp = pol.act(obs)
    assert buf.data_ptr() == env.action_buffer.data_ptr() and a.shape == (3, 100) and lp.shape == (3, 100)
    assert int(a.max()) <= 4 and torch.equal(env.action_buffer[:, :100].long(), a)
    before = env.state().clone()
    env.step(buf, agent_major=True)
    moved = (env.state() != before).any(dim=2)                       # [E, A]
    assert not (moved & (a.t() == 4)).any()                           # "stay" never moves an agent
    cost = env.cost[:, :100]
    assert torch.equal(cost.long(), (a != 4).long())
    np.testing.assert_allclose(pol.log_prob(obs, a).detach().cpu().numpy(), lp.cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_gaussian_policy_drives_collision():
    import safe_multiagent_rl_b200 as s
    from safe_multiagent_rl_b200.policy import BatchedGaussianPolicy
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    env = s.BatchedCollisionAvoidance(5, 3, n_envs=64, starts=rng.random((64, 3, 2)) * 5, landmarks=rng.random((64, 1, 2)) * 5)
    pol = BatchedGaussianPolicy(env)
    obs = env.reset()
    buf, a, lp = pol.act(obs)
    assert a.shape == (3, 64, 2) and lp.shape == (3, 64)
    assert torch.equal(env.action_buffer[0, :64], a[0, :, 0]) and torch.equal(env.action_buffer[5, :64], a[2, :, 1])
    env.step(buf, agent_major=True)
    assert torch.isfinite(env.reward[:, :64]).all()


def test_training_example_reduces_constraint_cost():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "train_coverage.py"), "--n_envs", "2048",
                          "--meta_cycles", "4", "--agent_cycles", "8", "--max_t", "30", "--thresholds", "10", "10", "10"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("meta")]
    assert len(lines) == 4
    first = eval(lines[0].split("cost - thr ")[1].split("  lambda")[0])
    last = eval(lines[-1].split("cost - thr ")[1].split("  lambda")[0])
    assert np.mean(last) < np.mean(first) - 1.0, (first, last)      # lambda pressure lowers the move count
