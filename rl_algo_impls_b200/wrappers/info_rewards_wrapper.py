"""InfoRewardsWrapper for device envs: multi-head rewards assembled in HBM.

Mirrors ``rl_algo_impls/wrappers/info_rewards_wrapper.py:13-64`` (constructor keywords ``info_paths``,
``episode_end``, ``multiplier``; ``get_by_path``).  The env reports extra per-env reward series inside ``infos``
(CUDA float32 tensors of shape [N]); ``step`` appends them behind the env's own reward as the ``[N, 1 + K]``
reward row a multi-head critic trains on -- one K7 launch, no host synchronisation, output at a fixed address
(capturable in the rollout step's CUDA graph).  A host (numpy) env keeps the reference's numpy wrapper.
"""
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
        per_series = lambda x, dtype: np.array(np.broadcast_to(np.asarray(x, dtype=dtype), (K,)))
        # one flag / factor per series; a scalar applies to all of them (info_rewards_wrapper.py:23-37)
        self.episode_end = per_series(episode_end, np.bool_)
        no_scaling = multiplier is None or (np.isscalar(multiplier) and multiplier == 1.0)
        self.multiplier: Optional[np.ndarray] = None if no_scaling else per_series(multiplier, np.float32)
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
