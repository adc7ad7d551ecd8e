from .base import BatchedEnv, SingleEnvAdapter
from .collision_avoidance import BatchedCollisionAvoidance
from .congestion import BatchedCongestion
from .coverage import BatchedCoverageDiscrete

__all__ = ["BatchedEnv", "SingleEnvAdapter", "BatchedCoverageDiscrete", "BatchedCongestion",
           "BatchedCollisionAvoidance"]
