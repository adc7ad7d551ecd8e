#!/usr/bin/env python
"""Crossover table of the thread mappings (smarl_set_kernel_variant): closed-loop (CUDA graph) and fused ms per
batch for every (env, agents) at lanes = 0 (one thread per env / per 4 envs), 2 and 4.

    python tools/time_coop.py [--envs collision,congestion,coverage] [--agents 12,16,24,32] [--n_envs 1048576] [--T 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from safe_multiagent_rl_b200 import _lib  # noqa: E402
from tools import sweep  # noqa: E402

KIND = {"coverage": _lib.ENV_COVERAGE, "congestion": _lib.ENV_CONGESTION, "collision": _lib.ENV_COLLISION}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="collision")
    ap.add_argument("--agents", default="12,16,20,24,28,32")
    ap.add_argument("--lanes", default="0,2,4")
    ap.add_argument("--n_envs", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=20)
    a = ap.parse_args()
    print("| env | size | agents | envs | T | lanes | closed ms | closed GB/s (frac of 6551) | fused ms |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for env_name in a.envs.split(","):
        for A in map(int, a.agents.split(",")):
            S = 2 * A
            for lanes in map(int, a.lanes.split(",")):
                with _lib.kernel_variant(KIND[env_name], lanes):
                    try:
                        r = sweep.run(env_name, min(S, 127), A, a.n_envs, a.T)
                    except Exception as ex:                      # an unsupported combination: say so and go on
                        print(f"| {env_name} | {S} | {A} | {a.n_envs} | {a.T} | {lanes} | error: {str(ex)[:80]} |", flush=True)
                        continue
                print(f"| {env_name} | {r['S']} | {A} | {a.n_envs} | {a.T} | {lanes} | {r['ms_closed']:.3f} | "
                      f"{r['closed_gbs']:.0f} ({r['closed_gbs'] / sweep.PEAK:.2f}) | {r['ms_fused']:.3f} |", flush=True)
                torch.cuda.empty_cache()
