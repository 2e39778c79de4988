"""NormalizeObservation / NormalizeReward for device envs (CUDA tensors in, CUDA tensors out).

Mirror ``rl_algo_impls/wrappers/normalize.py:18-122`` (constructor keywords, ``step`` / ``reset`` /
``masked_reset`` / ``save`` / ``load`` / ``load_from``) over ``utils/running_mean_std.py:10-48``.
The running moments are float64 tensors in HBM and one K6 launch per env step does the Chan merge
and the normalisation; nothing synchronises, so the wrappers sit inside the rollout step's CUDA graph.
A host (numpy) env keeps using the reference's own numpy wrappers: that side is not on this path.
"""
from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops


class RunningMeanStd:
    """Device-resident running (mean, var, count), float64 (running_mean_std.py:10-48)."""

    def __init__(self, device, epsilon: float = 1e-4, shape: Tuple[int, ...] = ()) -> None:
        self.shape = tuple(shape)
        n = int(np.prod(self.shape)) if self.shape else 1
        self.mean = torch.zeros(n, dtype=torch.float64, device=device)
        self.var = torch.ones(n, dtype=torch.float64, device=device)
        self._count = torch.full((n,), float(epsilon), dtype=torch.float64, device=device)  # scalar in the reference

    @property
    def count(self) -> float:
        return float(self._count[0].item())

    def save(self, path: str) -> None:
        np.savez_compressed(path, mean=self.mean.cpu().numpy().reshape(self.shape),
                            var=self.var.cpu().numpy().reshape(self.shape), count=self.count)

    def load(self, path: str, count_override: Optional[int] = None) -> None:
        data = np.load(path)
        self.mean.copy_(torch.from_numpy(np.asarray(data["mean"], np.float64).reshape(-1)))
        self.var.copy_(torch.from_numpy(np.asarray(data["var"], np.float64).reshape(-1)))
        self._count.fill_(float(data["count"]) if count_override is None else float(count_override))

    def load_from(self, existing: "RunningMeanStd") -> None:
        self.mean.copy_(existing.mean), self.var.copy_(existing.var), self._count.copy_(existing._count)


class ExponentialMovingMeanVar:
    """Device-resident exponential-moving (mean, mean of squares, var), float64 (running_mean_std.py:56-118)."""

    def __init__(self, device, alpha: Optional[float] = None, window_size=None, shape: Tuple[int, ...] = (),
                 per_env: int = 0) -> None:
        assert alpha is None or window_size is None, \
            f"Only one of alpha ({alpha}) or window_size ({window_size}) can be specified"
        if window_size is not None:
            alpha = 2 / (window_size + 1)
        assert alpha is not None, "Either alpha or window_size must be specified"
        assert 0 < alpha < 1, f"alpha ({alpha}) must be between 0 and 1 (exclusive)"
        self.alpha = alpha
        self.window_size = window_size if window_size is not None else (2 / alpha - 1)
        self.shape = tuple(shape)
        # per_env = N: the reference's update broadcasts against a 1-D batch (scalar rewards), which leaves one
        # set of moving moments PER ENV after the first update (running_mean_std.py:88-96); reproduced as is
        self.per_env = int(per_env)
        n = self.per_env if self.per_env else (int(np.prod(self.shape)) if self.shape else 1)
        self.mean = torch.zeros(n, dtype=torch.float64, device=device)
        self.squared_mean = torch.zeros(n, dtype=torch.float64, device=device)
        self.var = torch.ones(n, dtype=torch.float64, device=device)
        self._initialized = torch.zeros(n, dtype=torch.int32, device=device)

    @property
    def initialized(self) -> bool:
        return bool(self._initialized[0].item())

    def _state_shape(self) -> Tuple[int, ...]:
        """Shape of the arrays the reference would hold right now: `shape` before the first update and for vector
        rewards; (N,) once scalar rewards went through its broadcasting update (running_mean_std.py:79-96)."""
        return (self.per_env,) if self.per_env and self.initialized else self.shape

    def save(self, path: str) -> None:
        shape = self._state_shape()
        take = (lambda t: t.cpu().numpy().reshape(shape)) if shape != () or not self.per_env else (
            lambda t: t[:1].cpu().numpy().reshape(()))
        np.savez_compressed(path, mean=take(self.mean), var=take(self.var), initialized=self.initialized)

    def load(self, path: str, count_override: Optional[int] = None) -> None:
        """Accepts what either implementation wrote: arrays of the state's own size, or a single value that every
        entry takes (a reference file saved before its first update, shape ())."""
        data = np.load(path)
        for name, dst in (("mean", self.mean), ("var", self.var)):
            src = torch.from_numpy(np.asarray(data[name], np.float64).reshape(-1))
            if src.numel() not in (1, dst.numel()):
                raise ValueError(f"{path}: {name} has {src.numel()} entries, this normaliser holds {dst.numel()}")
            dst.copy_(src.expand(dst.numel()) if src.numel() == 1 else src)
        self.squared_mean.copy_(self.var + self.mean ** 2)
        self._initialized.fill_(int(bool(data["initialized"].item())))

    def load_from(self, existing: "ExponentialMovingMeanVar") -> None:
        self.mean.copy_(existing.mean), self.var.copy_(existing.var)
        self.squared_mean.copy_(existing.squared_mean), self._initialized.copy_(existing._initialized)


class HybridMovingMeanVar:
    """running_mean_std.py:120-170: running moments until `window_size` samples were seen, moving ones after."""

    def __init__(self, device, alpha: Optional[float] = None, window_size=None, shape: Tuple[int, ...] = (),
                 per_env: int = 0) -> None:
        self.rms = RunningMeanStd(device, shape=shape)
        self.emmv = ExponentialMovingMeanVar(device, alpha=alpha, window_size=window_size, shape=shape, per_env=per_env)

    def _blend(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        frac = self.rms.count / self.emmv.window_size
        return b.clone() if frac >= 1 else a * (1 - frac) + b * frac

    @property
    def mean(self) -> torch.Tensor:
        return self._blend(self.rms.mean, self.emmv.mean)

    @property
    def var(self) -> torch.Tensor:
        return self._blend(self.rms.var, self.emmv.var)

    def save(self, path: str) -> None:
        self.rms.save(path + "-rms.npz"), self.emmv.save(path + "-emmv.npz")

    def load(self, path: str, count_override: Optional[int] = None) -> None:
        self.rms.load(path + "-rms.npz", count_override=count_override)
        self.emmv.load(path + "-emmv.npz", count_override=count_override)

    def load_from(self, existing: "HybridMovingMeanVar") -> None:
        self.rms.load_from(existing.rms), self.emmv.load_from(existing.emmv)


class _Wrapper:
    def __init__(self, env) -> None:
        self.env = env
        if getattr(env, "device", None) is None:
            raise RuntimeError("the device normalisers wrap a device env (CUDA tensors); host envs keep numpy wrappers")
        self.device = env.device

    def __getattr__(self, name):  # num_envs, spaces, get_action_mask, ...
        return getattr(self.env, name)


class NormalizeObservation(_Wrapper):
    def __init__(self, env, training: bool = True, epsilon: float = 1e-8, clip: float = 10.0) -> None:
        super().__init__(env)
        self.rms = RunningMeanStd(self.device, shape=tuple(env.single_observation_space.shape))
        self.training, self.epsilon, self.clip = training, epsilon, clip
        self._out: Optional[torch.Tensor] = None

    def normalize(self, obs: torch.Tensor) -> torch.Tensor:
        x = obs.float().contiguous()
        if self._out is None or self._out.shape != x.shape:
            self._out = torch.empty_like(x)  # fixed output address: graph-capturable
        return ops.running_norm_obs(x, self.rms.mean, self.rms.var, self.rms._count, self.training, self.epsilon,
                                    self.clip, self._out)

    def step(self, action):
        obs, reward, terminations, truncations, info = self.env.step(action)
        return self.normalize(obs), reward, terminations, truncations, info

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.normalize(obs), info

    def save(self, path: str) -> None:
        self.rms.save(path)

    def load(self, path: str) -> None:
        self.rms.load(path)

    def load_from(self, existing: "NormalizeObservation") -> None:
        self.rms.load_from(existing.rms)


class NormalizeReward(_Wrapper):
    def __init__(self, env, training: bool = True, gamma: float = 0.99, epsilon: float = 1e-8, clip: float = 10.0,
                 shape: Tuple[int, ...] = (), exponential_moving_mean_var: bool = False, emv_window_size=None) -> None:
        super().__init__(env)
        self.rms = (HybridMovingMeanVar(self.device, window_size=emv_window_size, shape=tuple(shape),
                                        per_env=env.num_envs if tuple(shape) == () else 0)
                    if exponential_moving_mean_var else RunningMeanStd(self.device, shape=tuple(shape)))
        self.training, self.gamma, self.epsilon, self.clip = training, gamma, epsilon, clip
        self.returns = torch.zeros((env.num_envs,) + tuple(shape), dtype=torch.float64, device=self.device)
        self._out: Optional[torch.Tensor] = None

    def step(self, action):
        obs, reward, terminations, truncations, info = self.env.step(action)
        r = reward.float().contiguous()
        if self._out is None or self._out.shape != r.shape:
            self._out = torch.empty_like(r)
        done = torch.logical_or(terminations, truncations)
        if isinstance(self.rms, HybridMovingMeanVar):
            rms, ema = self.rms.rms, self.rms.emmv
            reward = ops.running_norm_reward_ema(r, done, self.returns, rms.mean, rms.var, rms._count, ema.mean,
                                                 ema.squared_mean, ema.var, ema._initialized, ema.alpha, self.gamma,
                                                 self.training, self.epsilon, self.clip, self._out,
                                                 per_env=bool(ema.per_env))
        else:
            reward = ops.running_norm_reward(r, done, self.returns, self.rms.mean, self.rms.var, self.rms._count,
                                             self.gamma, self.training, self.epsilon, self.clip, self._out)
        return obs, reward, terminations, truncations, info

    def reset(self, **kwargs):
        self.returns.zero_()
        return self.env.reset(**kwargs)

    def masked_reset(self, env_mask):
        self.returns[torch.as_tensor(env_mask, device=self.device)] = 0
        return self.env.masked_reset(env_mask)

    def save(self, path: str) -> None:
        self.rms.save(path)

    def load(self, path: str) -> None:
        self.rms.load(path)

    def load_from(self, existing: "NormalizeReward") -> None:
        self.rms.load_from(existing.rms)
