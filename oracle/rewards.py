"""Oracle: multi-head reward assembly.  TEST INFRASTRUCTURE.

numpy restatement of ``rl_algo_impls/wrappers/info_rewards_wrapper.py:39-57`` (InfoRewardsWrapper.step):
stack the info series, zero episode-end series on steps that do not end an episode, scale, append behind the
env's own reward.  Checked against the live reference class by tests/golden/make_golden_info_rewards.py.
"""
from typing import Optional, Sequence

import numpy as np


def assemble_rewards(r: np.ndarray, series: Sequence[np.ndarray], terminations: np.ndarray, truncations: np.ndarray,
                     episode_end: np.ndarray, multiplier: Optional[np.ndarray]) -> np.ndarray:
    r_to_add = np.stack(list(series), axis=-1)
    if episode_end.any():
        done = np.logical_or(terminations, truncations)[:, None]
        r_to_add = np.where(np.logical_or(done, ~episode_end[None, :]), r_to_add, 0)
    if multiplier is not None:
        r_to_add *= multiplier[None, :]
    if len(r.shape) == 1:
        r = np.expand_dims(r, axis=-1)
    return np.concatenate([r, r_to_add], axis=-1)
