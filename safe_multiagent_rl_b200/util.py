"""Factory mirroring safe_multi_agent_RL/util.py:6-22 (make_env) for the batched envs."""
from __future__ import annotations

from .envs.collision_avoidance import BatchedCollisionAvoidance
from .envs.congestion import BatchedCongestion
from .envs.coverage import BatchedCoverageDiscrete
from .envs.coverage_float import BatchedCoverageContinuous, BatchedCoverageDiscretized


def make_env(params, n_envs=1, device="cuda", **kw):
    """``params`` carries the reference's CLI fields (cli_parse.py:6-34).  Returns (env, continuous)."""
    name = params.environment
    if name == "CoverageDiscrete":
        return BatchedCoverageDiscrete(params.size, params.n_agents, n_envs=n_envs, shuffle=params.shuffle,
                                       weights=params.weights, device=device, **kw), False
    if name == "Collision":
        return BatchedCollisionAvoidance(params.size, params.n_agents, n_envs=n_envs,
                                         n_landmarks=params.n_landmarks, shuffle=params.shuffle,
                                         device=device, **kw), True
    if name == "Congestion":
        return BatchedCongestion(params.size, params.n_agents, n_envs=n_envs, noise=params.noise,
                                 shuffle=params.shuffle, device=device, **kw), False
    if name == "CoverageDiscretized":
        return BatchedCoverageDiscretized(params.size, params.n_agents, n_envs=n_envs, coarseness=params.coarseness,
                                          shuffle=params.shuffle, weights=params.weights, device=device, **kw), False
    if name == "CoverageContinuous":
        return BatchedCoverageContinuous(params.size, params.n_agents, n_envs=n_envs, shuffle=params.shuffle,
                                         weights=params.weights, coarseness=params.coarseness, device=device, **kw), True
    raise ValueError("params.environment must be CoverageDiscrete, CoverageDiscretized, CoverageContinuous, "
                     "Collision or Congestion")
