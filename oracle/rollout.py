"""Oracle: rollout flattening, minibatch index streams, row gather, num_actions.  TEST INFRASTRUCTURE.

Restates
  * ``rl_algo_impls/rollout/rollout.py:120-127``  flatten [T, N, ...] -> [T*N, ...]  (flat = t*N + n)
  * ``rl_algo_impls/rollout/vec_rollout.py:166-175``  torch.randperm / arange index stream, batch_size slices,
    short last minibatch kept
  * ``rl_algo_impls/rollout/rollout.py:56-69`` + ``shared/tensor_utils.py:66-72``  row gather of every Batch field
  * ``rl_algo_impls/rollout/rollout.py:130-180``  num_actions / per_position_num_actions
"""
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np
import torch

NumpyOrDict = Union[np.ndarray, Dict[str, np.ndarray]]


def flatten_time_major(a: np.ndarray) -> np.ndarray:
    return a.reshape((-1,) + a.shape[2:])


def minibatch_index_stream(total_steps: int, batch_size: int, shuffle: bool = True) -> List[torch.Tensor]:
    """One epoch of minibatch indices.  Consumes exactly one ``torch.randperm`` draw from the
    CPU default generator when ``shuffle`` (vec_rollout.py:168-172)."""
    order = torch.randperm(total_steps) if shuffle else torch.arange(total_steps)
    return [order[i : i + batch_size] for i in range(0, total_steps, batch_size)]


def gather_rows(fields: Dict[str, Union[torch.Tensor, Dict[str, torch.Tensor], None]], idx: torch.Tensor):
    out = {}
    for k, v in fields.items():
        if v is None:
            out[k] = None
        elif isinstance(v, dict):
            out[k] = {kk: vv[idx] for kk, vv in v.items()}
        else:
            out[k] = v[idx]
    return out


def per_position_num_actions(
    actions: np.ndarray,  # [..., HW, A]
    action_masks: np.ndarray,  # [..., HW, S] bool
    gates: Optional[Dict[int, Tuple[int, int]]],
    nvec: Optional[np.ndarray],
) -> np.ndarray:
    if not gates:
        return np.sum(np.any(action_masks, axis=-1), axis=-1)
    assert nvec is not None
    count = np.zeros(actions.shape[:-2], dtype=np.int32)
    start = 0
    for h, n in enumerate(nvec):
        m = action_masks[..., start : start + n]
        if h in gates:
            ref, required = gates[h]
            m = np.where(np.expand_dims(actions[..., ref] == required, axis=-1), m, False)
        count += np.sum(np.any(m, axis=-1), axis=-1)
        start += n
    return count


def num_actions(
    actions: NumpyOrDict,
    action_masks: Optional[NumpyOrDict],
    gates: Optional[Dict[int, Tuple[int, int]]],
    nvec: Optional[np.ndarray],
) -> Optional[np.ndarray]:
    if action_masks is None:
        return None
    if isinstance(action_masks, dict):
        cells = per_position_num_actions(actions["per_position"], action_masks["per_position"], gates, nvec)
        picks = action_masks["pick_position"].any(axis=-2).sum(axis=-1)
        with np.errstate(divide="ignore"):
            return (cells + np.where(picks > 0, np.log(picks), 0)).astype(np.float32)
    return per_position_num_actions(actions, action_masks, gates, nvec)
