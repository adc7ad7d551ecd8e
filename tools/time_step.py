"""Time the closed loop (T step launches) for one Coverage shape; SMARL_KEEP_POS=0/1 toggles the L2 policy."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s
E = int(sys.argv[1]); A, S, T = 16, 32, 50
g = torch.Generator(device="cuda"); g.manual_seed(0)
env = s.BatchedCoverageDiscrete(S, A, n_envs=E, weights=[1.0] * A, starts=np.zeros((E, A, 2), np.uint8))
env.start_x[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
env.start_y[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
lam = torch.full((A,), 0.1, dtype=torch.float64, device="cuda")
buf = env.new_rollout_buffer(T)
def loop():
    env.reset()
    for t in range(T):
        env.step(actions[t], lambdas=lam, out=(buf, t), agent_major=True)
for _ in range(3): loop()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): loop()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10 / T
print("E=%d keep=%s: %.1f us per step, %.0f GB/s algorithmic" % (E, os.environ.get("SMARL_KEEP_POS", "auto"), ms * 1e3, 19.25 * A * E / ms / 1e6))
