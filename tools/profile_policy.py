#!/usr/bin/env python
"""A few steps of the closed loop with the fused policy kernel (for ncu captures).

    python tools/profile_policy.py <n_agents> <n_envs>
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s  # noqa: E402
from safe_multiagent_rl_b200.policy import FusedDiscretePolicy  # noqa: E402

A, E = int(sys.argv[1]), int(sys.argv[2])
S = 2 * A
env = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=np.random.default_rng(0).integers(0, S, (E, A, 2)))
env.emit_obs = False
pol = FusedDiscretePolicy(env, seed=1)
env.reset()
for t in range(6):
    pol.act(t=t)
    env.step(env.action_buffer, agent_major=True)
torch.cuda.synchronize()
print("ok")
