from .ppo import PPO, TrainStats, TrainStepStats

__all__ = ["PPO", "TrainStats", "TrainStepStats"]
