"""CUDA-graph replay of a whole closed-loop batch of episodes.

For small batches (BASELINE config 1: 50 envs) a step kernel runs for ~3 us while the Python call
that launches it costs ~15 us, so the loop is launch-bound.  All libsmarl calls are capture-safe
(no synchronisation, no allocation), so reset + T x (policy, step) + accounting can be recorded once
into ONE graph and replayed: ~2.5 us per step on a B200.

The policy must be capturable too: it reads ``obs`` (a fixed view of ``env.obs``) and writes
``env.action_buffer`` in place -- e.g. ``BatchedDiscretePolicy.act`` or a copy from a recorded
action tensor whose CONTENTS may change between replays.
"""
from __future__ import annotations

import torch

from .rollout import G_REWARD_TO_GO


class GraphedClosedLoop:
    def __init__(self, env, n_steps, policy, lambdas, gamma, thresholds=None, g_mode=G_REWARD_TO_GO, lean=False):
        """policy(obs, t) must write env.action_buffer (kernel layout) in place."""
        self.env, self.T = env, int(n_steps)
        self.buffer = env.new_rollout_buffer(n_steps, g_mode if g_mode else G_REWARD_TO_GO, lean=lean)
        self._args = (policy, lambdas, gamma, thresholds, g_mode)
        self.out = self._run()                       # warm-up outside capture: allocations, caches, module load
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        episode = env._episode
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        env._episode = episode + 1 if env.shuffle else episode
        self.out["buffer"] = self.buffer

    def _run(self):
        policy, lambdas, gamma, thresholds, g_mode = self._args
        env, buf = self.env, self.buffer
        obs = env.reset()
        for t in range(self.T):
            policy(obs, t)
            obs, _, _, _ = env.step(env.action_buffer, lambdas=lambdas, out=(buf, t), agent_major=True)
        return buf.finish(gamma, thresholds, g_mode, n_active=getattr(env, "episode_len", None))

    def replay(self):
        """Run one batch of episodes; returns the same dict of views every time (contents updated).
        With shuffle=True the captured graph replays the SAME start draw: re-capture per episode
        index if fresh starts are needed."""
        self.graph.replay()
        return self.out
