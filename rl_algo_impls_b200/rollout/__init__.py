from .rollout import Batch, Rollout, RolloutGenerator
from .sync_step_rollout import SyncStepRolloutGenerator
from .vec_rollout import VecRollout

__all__ = ["Batch", "Rollout", "RolloutGenerator", "SyncStepRolloutGenerator", "VecRollout"]
