from .info_rewards_wrapper import InfoRewardsWrapper
from .normalize import NormalizeObservation, NormalizeReward, RunningMeanStd

__all__ = ["InfoRewardsWrapper", "NormalizeObservation", "NormalizeReward", "RunningMeanStd"]
