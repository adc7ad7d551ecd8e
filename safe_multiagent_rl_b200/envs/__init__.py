from .base import BatchedEnv, SingleEnvAdapter
from .collision_avoidance import BatchedCollisionAvoidance
from .congestion import BatchedCongestion
from .coverage import BatchedCoverageDiscrete
from .coverage_float import BatchedCoverageContinuous, BatchedCoverageDiscretized

__all__ = ["BatchedEnv", "SingleEnvAdapter", "BatchedCoverageDiscrete", "BatchedCongestion",
           "BatchedCollisionAvoidance", "BatchedCoverageContinuous", "BatchedCoverageDiscretized"]
