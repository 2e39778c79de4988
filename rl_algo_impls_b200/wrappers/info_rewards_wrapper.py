"""InfoRewardsWrapper for device envs: multi-head rewards assembled in HBM.

Mirrors ``rl_algo_impls/wrappers/info_rewards_wrapper.py:13-64`` (constructor keywords ``info_paths``,
``episode_end``, ``multiplier``; ``get_by_path``).  The env reports extra per-env reward series inside ``infos``
(CUDA float32 tensors of shape [N]); ``step`` appends them behind the env's own reward as the ``[N, 1 + K]``
reward row a multi-head critic trains on -- one K7 launch, no host synchronisation, output at a fixed address
(capturable in the rollout step's CUDA graph).  A host (numpy) env keeps the reference's numpy wrapper.
"""
import collections.abc
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from .. import ops
from .normalize import _Wrapper


def get_by_path(info: Dict, path: List[str]):
    """info_rewards_wrapper.py:60-64"""
    for key in path:
        info = info[key]
    return info


class InfoRewardsWrapper(_Wrapper):
    def __init__(self, env, info_paths: List[List[str]], episode_end: Union[bool, List[bool]] = True,
                 multiplier: Union[None, float, List[float]] = None) -> None:
        super().__init__(env)
        self.info_paths = info_paths
        K = len(info_paths)
        if isinstance(episode_end, collections.abc.Sequence):
            self.episode_end = np.array(episode_end, dtype=np.bool_)
        else:
            self.episode_end = np.full((K,), episode_end, dtype=np.bool_)
        if isinstance(multiplier, collections.abc.Sequence):
            self.multiplier: Optional[np.ndarray] = np.array(multiplier, dtype=np.float32)
        elif multiplier is not None and multiplier != 1.0:
            self.multiplier = np.full((K,), multiplier, dtype=np.float32)
        else:
            self.multiplier = None
        assert len(self.episode_end) == K and (self.multiplier is None or len(self.multiplier) == K)
        self._out: Optional[torch.Tensor] = None

    def step(self, action):
        o, r, terminations, truncations, infos = self.env.step(action)
        series = [get_by_path(infos, path).float().contiguous() for path in self.info_paths]
        base = r.float().contiguous()
        width = (base.shape[1] if base.dim() > 1 else 1) + len(series)
        if self._out is None or self._out.shape != (base.shape[0], width):
            self._out = torch.empty((base.shape[0], width), dtype=torch.float32, device=base.device)
        rewards = ops.reward_assemble(base, series, terminations, truncations, self.episode_end.tolist(),
                                      None if self.multiplier is None else self.multiplier.tolist(), self._out)
        return o, rewards, terminations, truncations, infos
