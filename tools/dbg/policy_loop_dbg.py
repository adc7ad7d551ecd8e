import os, sys, traceback, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import safe_multiagent_rl_b200 as s
import bench
a = argparse.Namespace(n_agents=16, max_t=50, size=32, n_envs=1 << 20, gamma=0.999)
try:
    out = bench.bench_policy_loop(s, a, torch.device("cuda:0"))
    import json
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "policy"}) for k, v in out.items() if k != "policy"}, indent=1))
except Exception:
    traceback.print_exc()
