from .a2c import A2C, TrainStats, TrainStepStats

__all__ = ["A2C", "TrainStats", "TrainStepStats"]
