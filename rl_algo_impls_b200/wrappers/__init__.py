from .info_rewards_wrapper import InfoRewardsWrapper
from .normalize import (ExponentialMovingMeanVar, HybridMovingMeanVar, NormalizeObservation, NormalizeReward,
                        RunningMeanStd)

__all__ = ["InfoRewardsWrapper", "NormalizeObservation", "NormalizeReward", "RunningMeanStd", "ExponentialMovingMeanVar",
           "HybridMovingMeanVar"]
