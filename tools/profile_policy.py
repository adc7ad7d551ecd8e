#!/usr/bin/env python
"""A few steps of the closed loop with the fused policy kernel (for ncu captures).

    python tools/profile_policy.py <n_agents> <n_envs> [variant] [--time]

variant: smarl_set_kernel_variant(SMARL_KERNEL_POLICY, .): -1 automatic, 0 FP32 pipes, 1 / 2 tensor cores (groups of
<= 16 / <= 8 agents).  --time prints the CUDA-event time per policy launch for every variant instead.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s  # noqa: E402
from safe_multiagent_rl_b200.policy import FusedDiscretePolicy  # noqa: E402

from safe_multiagent_rl_b200 import _lib  # noqa: E402

A, E = int(sys.argv[1]), int(sys.argv[2])
VARIANT = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else -1
S = 2 * A
env = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=np.random.default_rng(0).integers(0, S, (E, A, 2)))
env.emit_obs = False
pol = FusedDiscretePolicy(env, seed=1)
env.reset()
if "--time" in sys.argv:
    macs = 2 * A * 16 + 16 * 5
    for v in (0, 1, 2):
        with _lib.kernel_variant(_lib.KERNEL_POLICY, v):
            for t in range(3):
                pol.act(t=t)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            n = 20
            ev[0].record()
            for t in range(n):
                pol.act(t=t)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / n
            print(f"A={A} E={E} variant {v}: {ms * 1e3:.1f} us per launch, {A * E / ms / 1e6:.2f} G agent-steps/s, "
                  f"{2 * macs * A * E / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)")
    sys.exit(0)
with _lib.kernel_variant(_lib.KERNEL_POLICY, VARIANT):
    for t in range(6):
        pol.act(t=t)
        env.step(env.action_buffer, agent_major=True)
torch.cuda.synchronize()
print("ok")
