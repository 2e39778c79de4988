"""VecRollout: a fixed-shape [T, N, ...] rollout that lives in HBM.

Mirrors ``rl_algo_impls/rollout/vec_rollout.py:21-175``.  Differences are in *where* things
run, not in what they compute:

* the constructor accepts numpy arrays (the reference's contract; uploaded once) or CUDA
  tensors (the device-resident generator hands over its buffers without a copy);
* GAE + returns are one launch of the K1 scan (vec_rollout.py:78-88 -> ops.gae_scan),
  bit-exact with ``compute_advantages``;
* flattening [T, N, ...] -> [T*N, ...] is a view (flat = t*N + n, rollout.py:120-121);
* ``minibatches`` draws ``torch.randperm`` from the CPU default generator exactly where the
  reference does (vec_rollout.py:168-172), so the index stream is bit-identical, then each
  minibatch is one fused gather (K3).
"""
from collections import defaultdict
from typing import DefaultDict, Dict, Iterator, List, Optional, Union

import numpy as np
import torch

from .. import ops
from .rollout import Batch, BatchMapFn, Rollout

ArrayOrDict = Union[np.ndarray, torch.Tensor, Dict[str, Union[np.ndarray, torch.Tensor]]]


def _to_device(a, device: torch.device):
    if a is None:
        return None
    if isinstance(a, dict):
        return {k: _to_device(v, device) for k, v in a.items()}
    if isinstance(a, torch.Tensor):
        return a if a.device == device else a.to(device, non_blocking=True)
    return torch.as_tensor(np.ascontiguousarray(a)).to(device, non_blocking=True)


def _flatten(t):
    if t is None:
        return None
    if isinstance(t, dict):
        return {k: _flatten(v) for k, v in t.items()}
    return t.reshape((-1,) + tuple(t.shape[2:]))


class VecRollout(Rollout):
    def __init__(
        self,
        device: torch.device,
        next_episode_starts: ArrayOrDict,
        next_values: ArrayOrDict,
        obs: ArrayOrDict,
        actions: ArrayOrDict,
        rewards: ArrayOrDict,
        episode_starts: ArrayOrDict,
        values: ArrayOrDict,
        logprobs: Optional[ArrayOrDict],
        action_masks: Optional[ArrayOrDict],
        gamma,
        gae_lambda,
        scale_advantage_by_values_accuracy: bool = False,
        full_batch_off_accelerator: bool = False,
        subaction_mask: Optional[Dict[int, Dict[int, int]]] = None,
        action_plane_space=None,
        include_num_actions: bool = False,
        out_advantages: Optional[torch.Tensor] = None,
        out_returns: Optional[torch.Tensor] = None,
    ) -> None:
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VecRollout is device resident: it needs a CUDA device (no CPU path)")
        # full_batch_off_accelerator exists in the reference to spare a 24-40 GB GPU the whole
        # rollout (vec_rollout.py:114); with 180 GB of HBM the rollout stays resident.
        self.full_batch_off_accelerator = full_batch_off_accelerator
        self.obs = _to_device(obs, self.device)
        self.actions = _to_device(actions, self.device)
        self.rewards = _to_device(rewards, self.device).float()
        self.episode_starts = _to_device(episode_starts, self.device)
        self.values = _to_device(values, self.device).float()
        self.logprobs = _to_device(logprobs, self.device)
        self.action_masks = _to_device(action_masks, self.device)
        next_episode_starts = _to_device(next_episode_starts, self.device)
        next_values = _to_device(next_values, self.device).float()
        self.subaction_mask = subaction_mask
        self.action_plane_space = action_plane_space
        self._include_num_actions = include_num_actions

        self.advantages, self.returns = ops.gae_scan(
            self.rewards.contiguous(),
            self.values.contiguous(),
            self.episode_starts.contiguous(),
            next_episode_starts.contiguous(),
            next_values.contiguous(),
            gamma,
            gae_lambda,
            out_advantages,
            out_returns,
        )
        if scale_advantage_by_values_accuracy:  # vec_rollout.py:91-94
            spread = self.returns.max() - self.returns.min()
            self.advantages *= torch.exp(-torch.abs(self.values - self.returns) / spread)
        self._batch: Optional[Batch] = None
        self._y_true: Optional[np.ndarray] = None
        self._y_pred: Optional[np.ndarray] = None

    # -- Rollout surface ------------------------------------------------------------------------
    @property
    def y_true(self) -> np.ndarray:
        if self._y_true is None:
            self._y_true = _flatten(self.returns).cpu().numpy()
        return self._y_true

    @property
    def y_pred(self) -> np.ndarray:
        if self._y_pred is None:
            self._y_pred = _flatten(self.values).cpu().numpy()
        return self._y_pred

    def explained_variance(self) -> torch.Tensor:
        """1 - Var[y_true - y_pred] / Var[y_true] as a device scalar (ppo.py:415-418 without the D2H)."""
        y_true, y_pred = _flatten(self.returns).double(), _flatten(self.values).double()
        var_y = y_true.var(unbiased=False)
        return torch.where(var_y == 0, torch.full_like(var_y, float("nan")), 1 - (y_true - y_pred).var(unbiased=False) / var_y)

    @property
    def total_steps(self) -> int:
        return int(self.rewards.shape[0] * self.rewards.shape[1])

    @property
    def value_heads(self) -> int:
        return int(np.prod(self.values.shape[2:]))

    def num_minibatches(self, batch_size: int) -> int:
        return self.total_steps // batch_size + (1 if self.total_steps % batch_size else 0)

    def _num_actions(self) -> Optional[torch.Tensor]:
        """rollout.py:130-180; PPO ignores the field (ppo.py:300) so it is only built on request."""
        if not self._include_num_actions or self.action_masks is None:
            return None
        from ..actor.gridnet import num_actions_device

        return num_actions_device(self.actions, self.action_masks, self.subaction_mask, self.action_plane_space)

    def batch(self) -> Batch:
        if self._batch is None:
            na = self._num_actions()
            self._batch = Batch(
                _flatten(self.obs),
                _flatten(self.logprobs),
                _flatten(self.actions),
                _flatten(self.action_masks),
                _flatten(na) if na is not None else None,
                _flatten(self.values),
                _flatten(self.advantages),
                _flatten(self.returns),
            )
        return self._batch

    def add_to_batch(self, map_fn: BatchMapFn, batch_size: int) -> None:
        batch = self.batch()
        to_add: DefaultDict[str, List[torch.Tensor]] = defaultdict(list)
        for i in range(0, self.total_steps, batch_size):
            rows = torch.arange(i, min(i + batch_size, self.total_steps), device=self.device)
            for k, v in map_fn(batch[rows]).items():
                to_add[k].append(v)
        batch.additional.update({k: torch.cat(v).to(self.device) for k, v in to_add.items()})

    def minibatch_indices(self, batch_size: int, shuffle: bool = True) -> List[torch.Tensor]:
        """One epoch of index slices, on the device.  Same RNG call, same place, same dtype as
        the reference (vec_rollout.py:168-172): torch.randperm on the CPU default generator."""
        order = torch.randperm(self.total_steps) if shuffle else torch.arange(self.total_steps)
        order = order.to(self.device, non_blocking=True)
        return [order[i : i + batch_size] for i in range(0, self.total_steps, batch_size)]

    def minibatches(self, batch_size: int, shuffle: bool = True) -> Iterator[Batch]:
        batch = self.batch()
        for mb_idxs in self.minibatch_indices(batch_size, shuffle):
            yield batch[mb_idxs]
