#!/usr/bin/env python
"""CUDA-event time per launch of smarl_policy_act_gaussian against the PyTorch glue (BatchedGaussianPolicy.act).

    python tools/time_policy_gauss.py <n_agents> <n_envs>
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s  # noqa: E402
from safe_multiagent_rl_b200.policy import BatchedGaussianPolicy, FusedGaussianPolicy  # noqa: E402

A, E = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
env = s.BatchedCollisionAvoidance(5, A, n_envs=E, n_landmarks=1, starts=rng.random((E, A, 2)) * 5, landmarks=rng.random((E, 1, 2)) * 5)
obs = env.reset()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


fused = FusedGaussianPolicy(env, seed=1)
glue = BatchedGaussianPolicy(env)
t_f = timed(lambda: fused.act(t=3))
t_g = timed(lambda: glue.act(obs))
print(f"A={A} E={E}: fused {t_f * 1e3:.1f} us per launch ({A * E / t_f / 1e6:.2f} G agent-steps/s), "
      f"PyTorch glue {t_g * 1e3:.1f} us ({t_g / t_f:.1f}x)")
