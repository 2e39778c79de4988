"""Batch / Rollout / RolloutGenerator: the reference's rollout contract, device resident.

Mirrors ``rl_algo_impls/rollout/rollout.py:23-117``: ``Batch`` keeps the reference's field
order (consumers unpack it with ``dataclasses.astuple``, ppo/ppo.py:295-305), ``Rollout``
and ``RolloutGenerator`` keep its abstract surface.  ``Batch.__getitem__`` (rollout.py:56-69)
is one fused row-gather launch (K3) instead of one fancy-index kernel per field.
"""
import dataclasses
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Tuple, TypeVar, Union

import numpy as np
import torch

from .. import ops

TensorOrDict = Union[torch.Tensor, Dict[str, torch.Tensor]]
BatchSelf = TypeVar("BatchSelf", bound="Batch")


@dataclass
class Batch:
    obs: torch.Tensor
    logprobs: Optional[torch.Tensor]

    actions: TensorOrDict
    action_masks: Optional[TensorOrDict]
    num_actions: Optional[torch.Tensor]

    values: torch.Tensor

    advantages: torch.Tensor
    returns: torch.Tensor
    additional: Dict[str, torch.Tensor] = dataclasses.field(default_factory=dict)

    @property
    def device(self) -> torch.device:
        return self.obs.device

    def __len__(self) -> int:
        return self.obs.shape[0]

    def _flat(self) -> Tuple[List[Tuple[str, Optional[str]]], List[torch.Tensor]]:
        """(field, dict key) slots and their tensors, in field order, skipping None."""
        slots, tensors = [], []
        for f in dataclasses.fields(self):
            value = getattr(self, f.name)
            if value is None:
                continue
            if isinstance(value, dict):
                for k, t in value.items():
                    slots.append((f.name, k))
                    tensors.append(t)
            else:
                slots.append((f.name, None))
                tensors.append(value)
        return slots, tensors

    def _rebuild(self: BatchSelf, slots, tensors) -> BatchSelf:
        values: Dict[str, object] = {
            f.name: ({} if isinstance(getattr(self, f.name), dict) else None) for f in dataclasses.fields(self)
        }
        for (name, key), t in zip(slots, tensors):
            if key is None:
                values[name] = t
            else:
                values[name][key] = t  # type: ignore[index]
        return self.__class__(**values)  # type: ignore[arg-type]

    def to(self: BatchSelf, device: torch.device) -> BatchSelf:
        if self.device == torch.device(device):
            return self
        slots, tensors = self._flat()
        return self._rebuild(slots, [t.to(device) for t in tensors])

    def __getitem__(self: BatchSelf, indices: torch.Tensor) -> BatchSelf:
        slots, tensors = self._flat()
        idx = indices.to(device=self.device, dtype=torch.int64)
        if idx.dim() != 1:
            raise IndexError("Batch is indexed by a 1-D tensor of row numbers")
        return self._rebuild(slots, ops.gather_rows(tensors, idx))


BatchMapFn = Callable[[Batch], Dict[str, torch.Tensor]]


class Rollout(ABC):
    """rollout.py:78-103"""

    @property
    @abstractmethod
    def y_true(self) -> np.ndarray: ...

    @property
    @abstractmethod
    def y_pred(self) -> np.ndarray: ...

    @property
    @abstractmethod
    def total_steps(self) -> int: ...

    @abstractmethod
    def num_minibatches(self, batch_size: int) -> int: ...

    @abstractmethod
    def minibatches(self, batch_size: int, shuffle: bool = True) -> Iterator[Batch]: ...

    def add_to_batch(self, map_fn: BatchMapFn, batch_size: int) -> None: ...


class RolloutGenerator(ABC):
    """rollout.py:106-117"""

    def __init__(self, policy, vec_env, **kwargs) -> None:
        super().__init__()
        self.policy = policy
        self.vec_env = vec_env

    def prepare(self) -> None:
        pass

    @abstractmethod
    def rollout(self, **kwargs) -> Rollout: ...
