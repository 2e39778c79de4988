"""Minimal observation / action space descriptors (gymnasium is not a dependency of the data path).

Duck-type compatible with the attributes the reference reads from ``gymnasium.spaces``:
``shape``, ``dtype``, ``n``, ``nvec``, ``low``/``high``, ``Dict.spaces`` / ``__getitem__``.
"""
from typing import Dict as TDict
from typing import Optional, Sequence, Tuple

import numpy as np


class Space:
    def __init__(self, shape: Optional[Tuple[int, ...]] = None, dtype=None):
        self.shape = shape
        self.dtype = np.dtype(dtype) if dtype is not None else None


class Box(Space):
    def __init__(self, low, high, shape: Optional[Sequence[int]] = None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype), self.shape)


class Discrete(Space):
    def __init__(self, n: int):
        super().__init__((), np.int64)
        self.n = int(n)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def __len__(self) -> int:
        return len(self.nvec)


class Dict(Space):
    def __init__(self, spaces: TDict[str, Space]):
        super().__init__()
        self.spaces = dict(spaces)

    def __getitem__(self, key: str) -> Space:
        return self.spaces[key]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()


def is_discrete(space) -> bool:
    return hasattr(space, "n") and not hasattr(space, "nvec")


def is_multi_discrete(space) -> bool:
    return hasattr(space, "nvec")


def is_box(space) -> bool:
    return hasattr(space, "low") and hasattr(space, "high")


def is_dict(space) -> bool:
    return hasattr(space, "spaces")
