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


class ExponentialMovingMeanVar:
    """utils/running_mean_std.py:56-96"""

    def __init__(self, alpha=None, window_size=None, shape: Tuple[int, ...] = ()) -> None:
        if window_size is not None:
            alpha = 2 / (window_size + 1)
        self.alpha = alpha
        self.window_size = window_size if window_size is not None else (2 / alpha - 1)
        self.mean, self.squared_mean = np.zeros(shape, np.float64), np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.initialized = False

    def update(self, x: np.ndarray) -> None:
        if not self.initialized:
            self.mean = np.mean(x, axis=0, dtype=np.float64)
            self.squared_mean = np.mean(x**2, axis=0, dtype=np.float64)
            self.var = np.var(x, axis=0, dtype=np.float64)
            self.initialized = True
            return
        weights = (self.alpha * ((1 - self.alpha) ** np.arange(x.shape[0] - 1, -1, -1)))[:, None]
        self.mean = np.sum(weights * x, axis=0) + (1 - np.sum(weights)) * self.mean
        self.squared_mean = np.sum(weights * (x**2), axis=0) + (1 - np.sum(weights)) * self.squared_mean
        self.var = self.squared_mean - self.mean**2


class HybridMovingMeanVar:
    """utils/running_mean_std.py:120-153"""

    def __init__(self, alpha=None, window_size=None, shape: Tuple[int, ...] = ()) -> None:
        self.rms = RunningMeanStd(shape=shape)
        self.emmv = ExponentialMovingMeanVar(alpha=alpha, window_size=window_size, shape=shape)

    @property
    def var(self):
        frac = self.rms.count / self.emmv.window_size
        return self.emmv.var if frac >= 1 else self.rms.var * (1 - frac) + self.emmv.var * frac

    @property
    def mean(self):
        frac = self.rms.count / self.emmv.window_size
        return self.emmv.mean if frac >= 1 else self.rms.mean * (1 - frac) + self.emmv.mean * frac

    def update(self, x: np.ndarray) -> None:
        self.rms.update(x)
        self.emmv.update(x)


class ObsNormalizer:
    def __init__(self, shape, epsilon: float = 1e-8, clip: float = 10.0, training: bool = True):
        self.rms, self.epsilon, self.clip, self.training = RunningMeanStd(shape=shape), epsilon, clip, training

    def normalize(self, obs: np.ndarray) -> np.ndarray:
        if self.training:
            self.rms.update(obs)
        return np.clip((obs - self.rms.mean) / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)


class RewardNormalizer:
    def __init__(self, num_envs: int, shape=(), gamma: float = 0.99, epsilon: float = 1e-8, clip: float = 10.0,
                 training: bool = True, exponential_moving_mean_var: bool = False, emv_window_size=None):
        self.rms = (HybridMovingMeanVar(window_size=emv_window_size, shape=shape) if exponential_moving_mean_var
                    else RunningMeanStd(shape=shape))
        self.gamma, self.epsilon, self.clip, self.training = gamma, epsilon, clip, training
        self.returns = np.zeros((num_envs,) + tuple(shape))

    def step(self, rewards: np.ndarray, dones: np.ndarray) -> np.ndarray:
        if self.training:
            self.returns = self.returns * self.gamma + rewards
            self.rms.update(self.returns)
        out = np.clip(rewards / np.sqrt(self.rms.var + self.epsilon), -self.clip, self.clip)
        self.returns[dones] = 0
        return out
