"""TrajectoryRollout: a learner-facing Rollout over ragged trajectories, resident in HBM.

Mirrors ``rl_algo_impls/rollout/trajectory_rollout.py:15-121``: same constructor keywords, ``y_true`` /
``y_pred`` / ``total_steps`` / ``num_minibatches`` (floor division: the ragged tail is dropped, :83-84) /
``minibatches`` (``torch.randperm`` on the CPU default generator, bit-identical index stream, :104-110) /
``add_to_batch``.  Row order is the reference's: trajectories concatenated in list order.

What changes: the reference concatenates per-trajectory numpy arrays on the host, keeps the batch on the CPU
and ships every minibatch H2D; here the rows of ALL trajectories are gathered from the generator's StepStore
with one K3 launch, GAE over all segments is ONE K1b launch, ``num_actions`` is counted on the device and every
minibatch is one fused K3 gather.
"""
from collections import defaultdict
from typing import DefaultDict, Dict, Iterator, List, Optional

import numpy as np
import torch

from .rollout import Batch, BatchMapFn, Rollout
from .trajectory import Trajectory, _split_fields, segmented_gae


def concatenate_fields(trajectories: List[Trajectory]) -> Dict[str, torch.Tensor]:
    """Every field of the trajectories as one ``[total, ...]`` tensor: one K3 gather when all of them reference
    the same StepStore, a per-field ``torch.cat`` of materialised trajectories otherwise."""
    first = trajectories[0]
    if first.store is not None and all(t.store is first.store and t._fields is None for t in trajectories):
        rows = np.concatenate([t.rows for t in trajectories]).astype(np.int64)
        return first.store.gather(torch.from_numpy(rows).to(first.store.device))
    parts = [t.fields() for t in trajectories]
    return {k: torch.cat([p[k] for p in parts]) for k in parts[0]}


class TrajectoryRollout(Rollout):
    def __init__(
        self,
        device: torch.device,
        trajectories: List[Trajectory],
        scale_advantage_by_values_accuracy: bool = False,
        full_batch_off_accelerator: bool = True,  # accepted and ignored: 180 GB of HBM keeps the batch resident
        subaction_mask: Optional[Dict[int, Dict[int, int]]] = None,
        action_plane_space=None,
    ) -> None:
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TrajectoryRollout is device resident: it needs a CUDA device (no CPU path)")
        assert not scale_advantage_by_values_accuracy, \
            f"{self.__class__.__name__} doesn't implement scale_advantage_by_values_accuracy"  # trajectory_rollout.py:50-52
        assert len(trajectories) > 0, "TrajectoryRollout needs at least one trajectory"
        fields = concatenate_fields(trajectories)
        self.values = fields["values"].float()
        self.advantages, self.returns = segmented_gae(trajectories, self.values)
        actions, action_masks = _split_fields(fields, "actions"), _split_fields(fields, "action_masks")
        from ..actor.gridnet import num_actions_device

        num_actions = num_actions_device(actions, action_masks, subaction_mask, action_plane_space)
        self.batch = Batch(obs=fields["obs"], logprobs=fields.get("logprobs"), actions=actions,
                           action_masks=action_masks, num_actions=num_actions, values=self.values,
                           advantages=self.advantages, returns=self.returns)
        self._y_true: Optional[np.ndarray] = None
        self._y_pred: Optional[np.ndarray] = None

    @property
    def y_true(self) -> np.ndarray:
        if self._y_true is None:
            self._y_true = self.returns.cpu().numpy()
        return self._y_true

    @property
    def y_pred(self) -> np.ndarray:
        if self._y_pred is None:
            self._y_pred = self.values.cpu().numpy()
        return self._y_pred

    @property
    def total_steps(self) -> int:
        return len(self.batch)

    @property
    def value_heads(self) -> int:
        return int(np.prod(self.values.shape[1:]))

    def num_minibatches(self, batch_size: int) -> int:
        return self.total_steps // batch_size

    def explained_variance(self) -> torch.Tensor:
        """1 - Var[y_true - y_pred] / Var[y_true] as a device scalar (ppo.py:415-418 without the D2H)."""
        y_true, y_pred = self.returns.double(), self.values.double()
        var_y = y_true.var(unbiased=False)
        return torch.where(var_y == 0, torch.full_like(var_y, float("nan")), 1 - (y_true - y_pred).var(unbiased=False) / var_y)

    def add_to_batch(self, map_fn: BatchMapFn, batch_size: int) -> None:
        to_add: DefaultDict[str, List[torch.Tensor]] = defaultdict(list)
        for i in range(0, self.total_steps, batch_size):
            rows = torch.arange(i, min(i + batch_size, self.total_steps), device=self.device)
            for k, v in map_fn(self.batch[rows]).items():
                to_add[k].append(v)
        self.batch.additional.update({k: torch.cat(v).to(self.device) for k, v in to_add.items()})

    def minibatch_indices(self, batch_size: int, shuffle: bool = True) -> List[torch.Tensor]:
        order = torch.randperm(self.total_steps) if shuffle else torch.arange(self.total_steps)
        order = order.to(self.device, non_blocking=True)
        return [order[i * batch_size: (i + 1) * batch_size] for i in range(self.num_minibatches(batch_size))]

    def minibatches(self, batch_size: int, shuffle: bool = True) -> Iterator[Batch]:
        for mb_idxs in self.minibatch_indices(batch_size, shuffle):
            yield self.batch[mb_idxs]
