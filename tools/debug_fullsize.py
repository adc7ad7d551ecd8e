import sys, numpy as np, torch
sys.path.insert(0, ".")
import safe_multiagent_rl_b200 as s
from oracle import c_oracle as co, numpy_oracle as no
S, A, T, gamma = 32, 16, 4, 0.999
lut = no.coverage_penalty_lut(S, 8.0)
for E in [1 << 16, 1 << 20, 1 << 22]:
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    env = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=np.zeros((E, A, 2), np.uint8))
    env.start_x[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
    env.start_y[:, :E] = torch.randint(0, S, (A, E), generator=g, device="cuda", dtype=torch.uint8)
    actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
    ref = co.coverage_rollout(S, env.start_x.cpu().numpy(), env.start_y.cpu().numpy(), actions.cpu().numpy(), lut,
                              np.ones(A), np.zeros(A), gamma, E)
    out = env.rollout_closed_loop(lambda obs, t: actions[t], T, None, gamma)
    px = env.pos_x.cpu().numpy()
    bad = np.argwhere(px != ref["final_x"])
    print(E, "closed-loop mismatches", len(bad), bad[:3].tolist(), bad[-3:].tolist())
    fused = env.rollout(actions, gamma=gamma)
    px = env.pos_x.cpu().numpy()
    bad = np.argwhere(px != ref["final_x"])
    print(E, "fused mismatches", len(bad), bad[:3].tolist(), bad[-3:].tolist())
    # single step via step()
    env.reset()
    env.step(actions[0])
    r1 = co.coverage_rollout(S, env.start_x.cpu().numpy(), env.start_y.cpu().numpy(), actions[:1].cpu().numpy(), lut,
                             np.ones(A), np.zeros(A), gamma, E)
    bad = np.argwhere(env.pos_x.cpu().numpy() != r1["final_x"])
    print(E, "one step mismatches", len(bad), bad[:3].tolist())
