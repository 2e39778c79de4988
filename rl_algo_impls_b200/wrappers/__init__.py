from .normalize import NormalizeObservation, NormalizeReward, RunningMeanStd

__all__ = ["NormalizeObservation", "NormalizeReward", "RunningMeanStd"]
