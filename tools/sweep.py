#!/usr/bin/env python
"""Throughput sweep over BASELINE.json configs[0..4] shapes (all three envs, grid 5-64, agents 3-32,
envs 1K-16M) on one GPU: closed-loop (T step launches + returns kernel, captured in ONE CUDA graph so
small batches are not launch-bound) and fused open-loop rollout.  Prints a markdown table.

    python tools/sweep.py [--quick] > profiles/rNN/sweep.md

Under torchrun (WORLD_SIZE > 1) the same table is produced for N GPUs of one box, STRONG-scaled: `envs` is the total
batch, sharded over the ranks by contiguous global env ids (dist.shard_range); every closed-loop batch ends with the
NCCL all-reduce of the stats vector (smarl_stats_allreduce) and the lambda update; times are the max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py --multi
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safe_multiagent_rl_b200 as s  # noqa: E402
from safe_multiagent_rl_b200 import dist as sd  # noqa: E402

PEAK = 6551.0
RANK, WORLD, COMM = 0, 1, None


def make(env_name, S, A, E, g):
    dev = "cuda"
    if env_name == "coverage":
        env = s.BatchedCoverageDiscrete(S, A, n_envs=E, weights=[1.0 + (i % 3) for i in range(A)],
                                        starts=np.zeros((E, A, 2), np.uint8))
        env.start_x[:, :E] = torch.randint(0, S, (A, E), generator=g, device=dev, dtype=torch.uint8)
        env.start_y[:, :E] = torch.randint(0, S, (A, E), generator=g, device=dev, dtype=torch.uint8)
        K, bytes_step = A, 18.0 + 4.0 / A            # pos r/w 4, action 1, obs 8, reward 4, cost 1 (+ penalty 4/A); done not stored
    elif env_name == "congestion":
        rng = np.random.default_rng(0)
        env = s.BatchedCongestion(S, A, n_envs=E, noise=0.1, starts=np.zeros((E, A, 2), np.uint8),
                                  demand_rate=rng.random((S + 1, S + 1)) * 8 + 2, seed=1)
        env.start_x[1:, :E] = torch.randint(0, S, (A - 1, E), generator=g, device=dev, dtype=torch.uint8)
        env.start_y[1:, :E] = torch.randint(0, S, (A - 1, E), generator=g, device=dev, dtype=torch.uint8)
        K, bytes_step = 1, 18.0 + 8.0 / A            # pos r/w 4, action 1, effective move 1, obs 8, reward 4 (+ cost, penalty 8/A)
    elif env_name in ("coverage_cont", "coverage_disc"):       # the paper's Explore variants (f64 positions, coarseness 6)
        starts = np.zeros((E, A, 2))
        if env_name == "coverage_cont":
            env = s.BatchedCoverageContinuous(S, A, n_envs=E, weights=[1.0 + (i % 3) for i in range(A)], coarseness=6,
                                              starts=starts)
            pos = torch.rand((2, A, E), generator=g, device=dev, dtype=torch.float64) * S
            bytes_step = 32.0 + 8.0 + 8.0 + 4.0 + 4.0 + 4.0 / A            # pos r/w, action, obs, reward, cost, penalty
        else:
            env = s.BatchedCoverageDiscretized(S, A, n_envs=E, weights=[1.0 + (i % 3) for i in range(A)], coarseness=6,
                                               starts=starts)
            zoom = 6.0 / S
            pos = torch.floor(torch.rand((2, A, E), generator=g, device=dev, dtype=torch.float64) * S * zoom) / zoom
            bytes_step = 32.0 + 1.0 + 8.0 + 4.0 + 4.0 + 4.0 / A
        env.start_x[:, :E], env.start_y[:, :E] = pos[0], pos[1]
        K = A
    else:
        rng = np.random.default_rng(0)
        env = s.BatchedCollisionAvoidance(S, A, n_envs=E, starts=np.zeros((E, A, 2)), landmarks=np.zeros((E, 1, 2)))
        env.start_x[:, :E] = torch.rand((A, E), generator=g, device=dev, dtype=torch.float64) * S
        env.start_y[:, :E] = torch.rand((A, E), generator=g, device=dev, dtype=torch.float64) * S
        env.landmarks[:, :E] = torch.rand((2, E), generator=g, device=dev, dtype=torch.float64) * S
        K, bytes_step = 1, 55.0 + 24.0 / A
    return env, K, bytes_step


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    if WORLD > 1:                                # the slowest rank is the job's time
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run(env_name, S, A, E_total, T):
    g = torch.Generator(device="cuda"); g.manual_seed(RANK)
    _, E = sd.shard_range(E_total, RANK, WORLD)   # this rank's contiguous shard (the whole batch on one GPU)
    env, K, bytes_step = make(env_name, S, A, E, g)
    if env_name in ("collision", "coverage_cont"):
        actions = torch.randn((T, 2 * A, env.ld), generator=g, device="cuda") * 0.5
    elif env_name == "coverage_disc":
        actions = torch.randint(0, 9, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
    else:
        actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
    lam = torch.full((K,), 0.1, dtype=torch.float64, device="cuda")
    buf = env.new_rollout_buffer(T)
    gamma, thr = 0.99, [25.0] * K

    def closed():
        env.reset()
        for t in range(T):
            env.step(actions[t], lambdas=lam, out=(buf, t), agent_major=True)
        buf.finish(gamma, thr)
    closed()                                     # warm-up outside capture (allocations, module load)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        closed()
    iters = max(3, min(50, int(2e9 / (E * A * T)) + 3))
    thr_dev = torch.tensor(thr, dtype=torch.float64, device="cuda")

    def batch():
        graph.replay()
        if WORLD > 1:                            # MetaAgent.update's sum over ranks + the lambda update (meta_agent.py:32-39)
            COMM.allreduce(buf.stats_vec)
            s._lib.check(s._lib.load().smarl_lambda_update(s._lib.ptr(lam), s._lib.ptr(buf.stats_vec), s._lib.ptr(thr_dev), 0.0,
                                                         A, K, s._lib.stream_ptr()))
    ms_closed = timed(batch, iters)
    out = {}
    ms_fused = timed(lambda: env.rollout(actions, lambdas=lam, gamma=gamma, thresholds=thr, out=out), iters)
    E = E_total
    n = float(E) * A * T
    return dict(env=env_name, S=S, A=A, E=E, T=T, closed=n / ms_closed * 1e3, fused=n / ms_fused * 1e3,
                ms_closed=ms_closed, ms_fused=ms_fused,
                closed_gbs=(bytes_step + 9.3) * n / ms_closed / 1e6)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--multi", action="store_true", help="the reduced grid of the multi-GPU (torchrun) sweep")
    ap.add_argument("--cfg", action="append", default=[], help="env,size,agents,envs,T (repeatable): run only these")
    a = ap.parse_args()
    RANK, WORLD, local = sd.init_from_env()
    torch.cuda.set_device(local)
    if WORLD > 1:
        COMM = sd.StatsComm.from_process_group()
    cfgs = [("coverage", 5, 3, 50, 50),            # configs[0]
            ("collision", 5, 3, 65536, 50),        # configs[1]
            ("congestion", 10, 8, 1 << 20, 100),   # configs[2]
            ("coverage", 32, 16, 1 << 22, 50),     # configs[3] per GPU
            ("coverage_cont", 5, 3, 1 << 20, 50),  # ExploreContinuous / ExploreDiscretized, the paper's runs (coarseness 6)
            ("coverage_disc", 5, 3, 1 << 20, 50)]
    if a.cfg:
        cfgs = [(c.split(",")[0], *map(int, c.split(",")[1:])) for c in a.cfg]
    elif a.multi:                                    # configs[4] on N GPUs: the corners and the middle of the grid
        cfgs = [c for c in cfgs if c[3] >= 1 << 16]
        for env_name in ("coverage", "congestion", "collision"):
            for S, A in [(5, 3), (16, 8), (64, 32)]:
                for E in [1 << 16, 1 << 20, 1 << 22, 1 << 24]:
                    if env_name == "collision" and E * A > (1 << 27):
                        continue
                    if E * A * 50 > 8e9 * min(WORLD, 4):
                        continue
                    cfgs.append((env_name, S, A, E, 50))
    elif not a.quick:                                # configs[4]: sweep
        for env_name in ("coverage", "congestion", "collision"):
            for S, A in [(5, 3), (8, 4), (16, 8), (32, 16), (64, 32)]:
                for E in [1 << 10, 1 << 13, 1 << 16, 1 << 20, 1 << 22, 1 << 24]:
                    if env_name == "collision" and E * A > (1 << 26):
                        continue                     # f64 state + f32 actions: keep the buffers under ~60 GB
                    if E * A * 50 > 4e9:
                        continue
                    cfgs.append((env_name, S, A, E, 50))
    if RANK != 0:
        sys.stdout = open(os.devnull, "w")
    if WORLD > 1:
        print(f"## {WORLD} GPUs of one box, strong scaling: `envs` is the TOTAL batch, sharded over the ranks; every closed-loop batch "
              f"includes the NCCL stats all-reduce + lambda update; max over ranks; fractions are of {WORLD} x 6551 GB/s\n")
    print("Envelope of the kernels behind this table: n_agents <= 32; grid size <= 127 (Coverage: doubled u8 coordinates index the "
          "penalty table, fieldview^2 <= 12279 table entries) / <= 254 (Congestion); fused Coverage rollout T <= 255 (byte cost "
          "counters); per call (2 * n_agents + 1) * ld < 2^32 (32-bit element offsets: n_agents = 32 caps a call at ~6.6e7 envs; "
          "larger batches are split by the caller); Collision n_landmarks <= 64.  Outside it the entry points return "
          "SMARL_EUNSUPPORTED / the ctors raise.  Thread mapping per row: automatic (one thread per env / per four envs up to the "
          "measured crossover, lane-cooperative kernels above it; DESIGN.md section 3).\n")
    print("| env | size | agents | envs | T | closed-loop agent-steps/s (CUDA graph) | ms | closed-loop algorithmic GB/s "
          "(fraction of the 6551 GB/s HBM peak) | fused agent-steps/s | ms |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for c in cfgs:
        r = run(*c)
        print(f"| {r['env']} | {r['S']} | {r['A']} | {r['E']} | {r['T']} | {r['closed']:.3g} | {r['ms_closed']:.3f} | "
              f"{r['closed_gbs']:.0f} ({r['closed_gbs'] / (PEAK * WORLD):.2f}) | {r['fused']:.3g} | {r['ms_fused']:.3f} |", flush=True)
        torch.cuda.empty_cache()
