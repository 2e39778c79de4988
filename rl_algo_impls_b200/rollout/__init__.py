from .guided_learner_rollout import GuidedLearnerRolloutGenerator, RandomGuidedLearnerRolloutGenerator
from .reference_ai_rollout import ReferenceAIRolloutGenerator
from .rollout import Batch, Rollout, RolloutGenerator
from .sync_step_rollout import SyncStepRolloutGenerator
from .trajectory import DiscreteSkipsTrajectoryBuilder, StepStore, Trajectory, TrajectoryBuilder
from .trajectory_rollout import TrajectoryRollout
from .vec_rollout import VecRollout

__all__ = ["Batch", "Rollout", "RolloutGenerator", "SyncStepRolloutGenerator", "VecRollout", "Trajectory",
           "TrajectoryBuilder", "DiscreteSkipsTrajectoryBuilder", "StepStore", "TrajectoryRollout",
           "GuidedLearnerRolloutGenerator", "RandomGuidedLearnerRolloutGenerator", "ReferenceAIRolloutGenerator"]
