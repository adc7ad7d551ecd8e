#!/usr/bin/env python
"""Run a few closed-loop steps + one fused rollout of one env config (for ncu captures).

    python tools/profile_env.py congestion 10 8 1048576 100
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sweep import make  # noqa: E402

env_name, S, A, E, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
g = torch.Generator(device="cuda"); g.manual_seed(0)
env, K, _ = make(env_name, S, A, E, g)
if env_name in ("collision", "coverage_cont"):
    actions = torch.randn((T, 2 * A, env.ld), generator=g, device="cuda") * 0.5
elif env_name == "coverage_disc":
    actions = torch.randint(0, 9, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
else:
    actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
lam = torch.full((K,), 0.1, dtype=torch.float64, device="cuda")
buf = env.new_rollout_buffer(T)
for rep in range(2):
    env.reset()
    for t in range(T):
        env.step(actions[t], lambdas=lam, out=(buf, t), agent_major=True)
    buf.finish(0.99, [25.0] * K)
    env.rollout(actions, lambdas=lam, gamma=0.99, thresholds=[25.0] * K)
torch.cuda.synchronize()
print("ok")
