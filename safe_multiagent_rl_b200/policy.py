"""Batched policy glue: the n_agents independent policies of the reference
(one DiscretePolicy / ContinuousPolicy per agent, safe_multi_agent_RL/agent.py:23-76, each fed the
joint state ``np.array(state).flatten()``, main.py:32-35) evaluated for ALL envs at once.

This is plain PyTorch on purpose -- 16-unit MLPs are not the hot path -- but it closes the loop
without a host round trip: ``obs`` is the env's own ``[E, S]`` view (no copy), the per-agent
weights are stacked so one batched matmul serves all agents, and sampled actions are written
straight into ``env.action_buffer`` (the kernel layout ``[rows, ld]``), which ``env.step`` takes
zero-copy with ``agent_major=True``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Stacked(nn.Module):
    """A independent ``nn.Linear(in, out)`` layers as one [A, in, out] parameter (default Linear init)."""

    def __init__(self, n, fan_in, fan_out):
        super().__init__()
        bound = 1.0 / math.sqrt(fan_in)
        self.weight = nn.Parameter(torch.empty(n, fan_in, fan_out).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(n, 1, fan_out).uniform_(-bound, bound))

    def forward(self, x):            # x [A, E, in] or [E, in] (shared input)
        return torch.matmul(x, self.weight) + self.bias


class BatchedDiscretePolicy(nn.Module):
    """``DiscretePolicy`` (agent.py:23-47) x n_agents: fc1 -> relu -> fc2 -> softmax -> Categorical."""

    def __init__(self, env, hidden_size=16):
        super().__init__()
        self.env = env
        self.fc1 = _Stacked(env.n_agents, env.state_space, hidden_size)
        self.fc2 = _Stacked(env.n_agents, hidden_size, env.action_space)
        self.to(env.device)

    def logits(self, obs):           # obs [E, S] -> [A, E, n_actions]
        return self.fc2(F.relu(self.fc1(obs)))

    def dist(self, obs):
        return torch.distributions.Categorical(logits=self.logits(obs), validate_args=False)   # no host sync: graph-capturable

    @torch.no_grad()
    def act(self, obs):
        """Sample one action per (agent, env); writes ``env.action_buffer`` in place and returns
        (action_buffer, actions [A, E] int64, log_prob [A, E])."""
        d = self.dist(obs)
        a = d.sample()
        self.env.action_buffer[:, : self.env.n_envs].copy_(a)
        return self.env.action_buffer, a, d.log_prob(a)

    def log_prob(self, obs, actions):
        return self.dist(obs).log_prob(actions)


class BatchedGaussianPolicy(nn.Module):
    """``ContinuousPolicy`` (agent.py:50-76) x n_agents: diagonal Gaussian with variance relu(.) + 1e-4."""

    def __init__(self, env, hidden_size=16):
        super().__init__()
        self.env = env
        self.fc1 = _Stacked(env.n_agents, env.state_space, hidden_size)
        self.fc2 = _Stacked(env.n_agents, hidden_size, env.action_space)
        self.fc2_ = _Stacked(env.n_agents, hidden_size, env.action_space)
        self.to(env.device)

    def dist(self, obs):
        h = F.relu(self.fc1(obs))
        mu, var = self.fc2(h), F.relu(self.fc2_(h)) + 1e-4
        return torch.distributions.Independent(torch.distributions.Normal(mu, var.sqrt(), validate_args=False), 1,
                                              validate_args=False)

    @torch.no_grad()
    def act(self, obs):
        """actions [A, E, 2] float32 -> ``env.action_buffer`` rows dx0, dy0, dx1, ... in place."""
        d = self.dist(obs)
        a = d.sample()
        A, E = self.env.n_agents, self.env.n_envs
        self.env.action_buffer[:, :E].view(A, 2, E).copy_(a.permute(0, 2, 1))
        return self.env.action_buffer, a, d.log_prob(a)

    def log_prob(self, obs, actions):
        return self.dist(obs).log_prob(actions)
