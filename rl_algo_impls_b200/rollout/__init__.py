from .rollout import Batch, Rollout, RolloutGenerator
from .vec_rollout import VecRollout

__all__ = ["Batch", "Rollout", "RolloutGenerator", "VecRollout"]
