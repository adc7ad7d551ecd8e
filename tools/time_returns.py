"""Time smarl_rollout_returns on the headline shape for the CTA size in SMARL_RETURNS_THREADS."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s
from safe_multiagent_rl_b200 import _lib
A, E, T = 16, 1 << 22, 50
buf = s.RolloutBuffer(T, A, A, E, torch.uint8, "cuda", g_mode=1)
buf.reward.normal_(); buf.penalty.uniform_(); buf.cost.random_(0, 2)
for _ in range(3):
    buf.finish(0.999, [25.0] * A)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    buf.finish(0.999, [25.0] * A)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(os.environ.get("SMARL_RETURNS_THREADS", "default"), "threads: %.3f ms  %.0f GB/s" % (ms, 31.8e9 / ms / 1e6))
