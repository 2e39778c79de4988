"""Oracle: running-moment observation / reward normalisers.  TEST INFRASTRUCTURE.

numpy restatement of ``rl_algo_impls/utils/running_mean_std.py:10-33`` (RunningMeanStd.update: batch
mean / var over axis 0, Chan merge) and ``rl_algo_impls/wrappers/normalize.py:39-48,84-110``
(NormalizeObservation.normalize, NormalizeReward.step).  Checked against the live reference classes by
tests/golden/make_golden.py.
"""
from typing import Tuple

import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon: float = 1e-4, shape: Tuple[int, ...] = ()) -> None:
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, x: np.ndarray) -> None:
        batch_mean, batch_var, batch_count = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
        delta = batch_mean - self.mean
        total_count = self.count + batch_count
        self.mean += delta * batch_count / total_count
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / total_count
        self.var = m2 / total_count
        self.count = total_count


class ObsNormalizer:
    def __init__(self, shape, epsilon: float = 1e-8, clip: float = 10.0, training: bool = True):
        self.rms, self.epsilon, self.clip, self.training = RunningMeanStd(shape=shape), epsilon, clip, training

    def normalize(self, obs: np.ndarray) -> np.ndarray:
        if self.training:
            self.rms.update(obs)
        return np.clip((obs - self.rms.mean) / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)


class RewardNormalizer:
    def __init__(self, num_envs: int, shape=(), gamma: float = 0.99, epsilon: float = 1e-8, clip: float = 10.0,
                 training: bool = True):
        self.rms = RunningMeanStd(shape=shape)
        self.gamma, self.epsilon, self.clip, self.training = gamma, epsilon, clip, training
        self.returns = np.zeros((num_envs,) + tuple(shape))

    def step(self, rewards: np.ndarray, dones: np.ndarray) -> np.ndarray:
        if self.training:
            self.returns = self.returns * self.gamma + rewards
            self.rms.update(self.returns)
        out = np.clip(rewards / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)
        self.returns[dones] = 0
        return out
