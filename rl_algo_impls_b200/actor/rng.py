"""Seed / offset bookkeeping for the counter-based sampling kernels (K5).

The kernels are stateless: every launch gets (seed, offset) and derives its random bits from
Philox4x32-10 counters.  The seed is drawn once from torch's CPU default generator (so
``torch.manual_seed`` / the reference's ``set_seeds`` makes runs reproducible) and the offset
advances by one per launch.
"""
from typing import Optional, Tuple

import torch

_seed: Optional[int] = None
_offset = 0


def reseed(seed: Optional[int] = None) -> None:
    global _seed, _offset
    _seed = int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed)
    _offset = 0


def next_sample_stream() -> Tuple[int, int]:
    global _offset
    if _seed is None:
        reseed()
    _offset += 1
    return _seed, _offset


def current_seed() -> int:
    if _seed is None:
        reseed()
    return _seed
