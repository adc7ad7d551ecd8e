"""Time the accounting kernels alone on the bench shape (CoverageDiscrete A=16, T=50, E envs).

    python tools/time_returns.py [E] [full|lean]

Prints the CUDA-event time per launch and the algorithmic GB/s (bench.py's returns bytes:
read reward 4 + cost 1 + penalty 4/A, write G 4 + (R, modR, C) 12/T per agent-step).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s
from safe_multiagent_rl_b200.rollout import RolloutBuffer

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
A, T, K = 16, 50, 16
dev = "cuda"
w = torch.tensor([1.0 + (i % 3) for i in range(A)], dtype=torch.float32, device=dev)
buf = RolloutBuffer(T, A, K, E, device=dev, g_mode=s.G_REWARD_TO_GO, shared_reward=(mode == "lean"), weights=w,
                    store_done=False)
buf.reward.normal_()
buf.cost.random_(0, 2)
buf.penalty.uniform_()
thr = torch.full((K,), 25.0, dtype=torch.float64, device=dev)
for _ in range(3):
    buf.finish(0.999, thr)
torch.cuda.synchronize()
n = 10
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    buf.finish(0.999, thr)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
if mode == "lean":
    per = 4.0 / A + 1 + 4.0 / A + 4 + 12.0 / T
else:
    per = 4 + 1 + 4.0 / A + 4 + 12.0 / T
gbs = per * A * E * T / ms / 1e6
print("%s E=%d batch=%s threads=%s: %.3f ms per launch, %.0f GB/s algorithmic (%.2f B per agent-step)"
      % (mode, E, os.environ.get("SMARL_EXP_BATCH", "default"), os.environ.get("SMARL_EXP_THREADS", "default"), ms, gbs, per))
