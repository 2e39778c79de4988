"""Seeded synthetic inputs for the PPO data path.  Shared by the parity
tests, smoke() and bench.py so that every leg sees the same tensors."""
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

MICRORTS_NVEC = (6, 4, 4, 4, 4, 7, 49)
MICRORTS_GATES = {0: {1: 1, 2: 2, 3: 3, 4: 4, 5: 4, 6: 5}}  # ppo-Microrts.yml subaction_mask
LUX_NVEC = (4, 6, 4, 4, 5, 5)
LUX_GATES = {1: {2: 0, 3: 1, 4: 1, 5: 2}}  # ppo-LuxAI_S2.yml subaction_mask


def gae_inputs(seed: int, T: int, N: int, V: int = 1, p_start: float = 1 / 200):
    rng = np.random.default_rng(seed)
    shape = (T, N) if V == 1 else (T, N, V)
    nshape = (N,) if V == 1 else (N, V)
    return dict(
        rewards=rng.standard_normal(shape, dtype=np.float32),
        values=rng.standard_normal(shape, dtype=np.float32),
        episode_starts=rng.random((T, N)) < p_start,
        next_episode_starts=rng.random((N,)) < p_start,
        next_values=rng.standard_normal(nshape, dtype=np.float32),
    )


def gridnet_masks(rng, B: int, HW: int, nvec: Sequence[int], unit_p: float) -> np.ndarray:
    """[B, HW, S] bool: a cell without a unit is all False; a cell with one has >= 1 valid entry per head."""
    S = int(sum(nvec))
    mask = rng.random((B, HW, S)) < 0.5
    start = 0
    for n in nvec:
        forced = rng.integers(0, n, size=(B, HW))
        np.put_along_axis(mask[..., start : start + n], forced[..., None], True, axis=-1)
        start += n
    has_unit = rng.random((B, HW)) < unit_p
    return mask & has_unit[..., None]


def gridnet_inputs(
    seed: int,
    B: int,
    HW: int,
    nvec: Sequence[int] = MICRORTS_NVEC,
    n_pick: int = 0,
    unit_p: float = 0.06,
    logit_scale: float = 1.0,
    masked_action_p: float = 0.0,
):
    """logits [B, HW, S + n_pick] f32, mask [B, HW, S] bool, pick_mask [B, n_pick, HW] bool or None,
    actions [B, HW, A] int64 (valid wherever the head has a valid entry), pick_actions [B, n_pick] or None.
    ``masked_action_p`` > 0: that fraction of the heads (and pick heads) that have both valid and masked entries get a
    MASKED entry as their action -- the reference then keeps the finfo.min logit for it (categorical.py:25-36)."""
    rng = np.random.default_rng(seed)
    S, A = int(sum(nvec)), len(nvec)
    logits = (rng.standard_normal((B, HW, S + n_pick)) * logit_scale).astype(np.float32)
    mask = gridnet_masks(rng, B, HW, nvec, unit_p)
    actions = np.zeros((B, HW, A), dtype=np.int64)
    start = 0
    for h, n in enumerate(nvec):
        m = mask[..., start : start + n]
        score = np.where(m, logits[..., start : start + n] + rng.gumbel(size=m.shape), -np.inf)
        any_valid = m.any(-1)
        actions[..., h] = np.where(any_valid, np.argmax(np.where(any_valid[..., None], score, 0.0), -1),
                                   rng.integers(0, n, size=(B, HW)))
        if masked_action_p > 0:
            can = any_valid & (~m).any(-1) & (rng.random((B, HW)) < masked_action_p)
            actions[..., h] = np.where(can, np.argmax(np.where(~m, rng.random(m.shape), -1.0), -1), actions[..., h])
        start += n
    pick_mask = pick_actions = None
    if n_pick:
        pick_mask = rng.random((B, n_pick, HW)) < 0.05
        forced = rng.integers(0, HW, size=(B, n_pick))
        np.put_along_axis(pick_mask, forced[..., None], True, axis=-1)
        empty = rng.random((B, n_pick)) < 0.25  # late game: nothing to pick
        pick_mask &= ~empty[..., None]
        pl = np.transpose(logits[..., S:], (0, 2, 1))
        score = np.where(pick_mask, pl + rng.gumbel(size=pl.shape), -np.inf)
        anyp = pick_mask.any(-1)
        pick_actions = np.where(anyp, np.argmax(np.where(anyp[..., None], score, 0.0), -1),
                                rng.integers(0, HW, size=(B, n_pick))).astype(np.int64)
        if masked_action_p > 0:
            can = anyp & (rng.random((B, n_pick)) < max(masked_action_p, 0.3))
            pick_actions = np.where(can, np.argmax(np.where(~pick_mask, rng.random(pick_mask.shape), -1.0), -1),
                                    pick_actions).astype(np.int64)
    return dict(logits=logits, mask=mask, pick_mask=pick_mask, actions=actions, pick_actions=pick_actions)


def ppo_inputs(seed: int, B: int, V: int = 1, adv_v: Optional[int] = None, logp_scale: float = 1.0):
    rng = np.random.default_rng(seed + 77)
    adv_v = V if adv_v is None else adv_v
    vs = (B,) if V == 1 else (B, V)
    as_ = (B,) if adv_v == 1 else (B, adv_v)
    old_values = rng.standard_normal(vs, dtype=np.float32)
    return dict(
        old_logp_noise=(rng.standard_normal((B,)) * 0.15 * logp_scale).astype(np.float32),
        adv=rng.standard_normal(as_, dtype=np.float32),
        old_values=old_values,
        returns=(old_values + rng.standard_normal(vs, dtype=np.float32) * 0.5).astype(np.float32),
        new_values=(old_values + rng.standard_normal(vs, dtype=np.float32) * 0.15).astype(np.float32),
    )


def to_torch(d: Dict[str, Optional[np.ndarray]], device=None) -> Dict[str, Optional[torch.Tensor]]:
    out = {}
    for k, v in d.items():
        out[k] = None if v is None else torch.from_numpy(np.ascontiguousarray(v)).to(device or "cpu")
    return out
