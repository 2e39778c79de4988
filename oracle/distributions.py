"""Oracle: masked-categorical / GridNet / Gaussian log-prob and entropy.  TEST INFRASTRUCTURE.

torch-CPU restatement (the reference's own arithmetic library) of

* ``rl_algo_impls/shared/actor/categorical.py:12-54``  (MaskedCategorical)
* ``rl_algo_impls/shared/actor/gridnet.py:38-193``     (GridnetDistribution)
* ``rl_algo_impls/shared/actor/gaussian.py:11-16,42-45`` (GaussianDistribution)

Gradients come from torch autograd over these same ops, as in the reference.
Checked against the live reference classes by tests/golden/make_golden.py.
"""
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

TensorOrDict = Union[torch.Tensor, Dict[str, torch.Tensor]]
Gate = Dict[int, Tuple[int, int]]  # head index -> (reference head index, required value)


def gates_from_subaction_mask(subaction_mask: Optional[Dict[int, Dict[int, int]]]) -> Gate:
    """gridnet.py:26-35 -- {ref: {head: value}} -> {head: (ref, value)}."""
    gates: Gate = {}
    for ref, per_head in (subaction_mask or {}).items():
        for head, value in per_head.items():
            gates[int(head)] = (int(ref), int(value))
    return gates


class MaskedLogits:
    """One categorical over the last dim: normalised log-probs after masking.

    categorical.py:22-31: ``where(mask, logits, finfo(dtype).min)`` then torch
    ``Categorical.__init__`` which stores ``logits - logsumexp(logits)``.
    """

    def __init__(self, logits: torch.Tensor, mask: Optional[torch.Tensor]):
        self.mask = mask
        if mask is not None:
            logits = torch.where(mask, logits, torch.finfo(logits.dtype).min)
        self.logp_all = logits - logits.logsumexp(dim=-1, keepdim=True)
        self._probs = None

    @property
    def probs(self) -> torch.Tensor:
        # torch.distributions.utils.logits_to_probs == softmax of the *normalised* logits
        if self._probs is None:
            self._probs = torch.softmax(self.logp_all, dim=-1)
        return self._probs

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        # torch Categorical.log_prob: gather on the normalised logits (categorical.py:33-34)
        value = value.long().unsqueeze(-1)
        value, logp = torch.broadcast_tensors(value, self.logp_all)
        return logp.gather(-1, value[..., :1]).squeeze(-1)

    def entropy(self) -> torch.Tensor:
        if self.mask is None:
            # torch Categorical.entropy: clamp at finfo.min, -sum(p * logp)
            lp = torch.clamp(self.logp_all, min=torch.finfo(self.logp_all.dtype).min)
            return -(lp * self.probs).sum(-1)
        # categorical.py:44-54
        plogp = self.logp_all * self.probs
        zero = torch.tensor(0, dtype=plogp.dtype, device=plogp.device)
        return -torch.where(self.mask, plogp, zero).sum(-1)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        p2 = self.probs.reshape(-1, self.probs.shape[-1])
        s = torch.multinomial(p2, 1, True, generator=generator)
        return s.reshape(self.probs.shape[:-1])


class Gridnet:
    """Per-cell MultiDiscrete distribution (+ optional pick_position categoricals over cells).

    logits ``[B, H, W, S']`` (S' = sum(nvec) + n_pick), masks ``[B, HW, S]`` bool or a dict
    with ``per_position`` and ``pick_position [B, n_pick, HW]``; actions ``[B, HW, A]`` or a
    dict with ``per_position`` and ``pick_position [B, n_pick]``.  gridnet.py:39-102.
    """

    def __init__(
        self,
        map_size: int,
        nvec: Sequence[int],
        logits: torch.Tensor,
        masks: TensorOrDict,
        gates: Optional[Gate] = None,
    ):
        self.map_size = int(map_size)
        self.nvec = [int(n) for n in nvec]
        self.gates = gates or {}
        S = sum(self.nvec)
        cell_masks = masks["per_position"] if isinstance(masks, dict) else masks
        cell_masks = cell_masks.reshape(-1, cell_masks.shape[-1])
        rows = logits.reshape(-1, logits.shape[-1])
        self.heads: List[MaskedLogits] = [
            MaskedLogits(lg, m)
            for lg, m in zip(
                torch.split(rows[:, :S], self.nvec, dim=1), torch.split(cell_masks, self.nvec, dim=1)
            )
        ]
        self.picks: Optional[List[MaskedLogits]] = None
        if isinstance(masks, dict) and "pick_position" in masks:
            pm = masks["pick_position"]  # [B, n_pick, HW]
            n_pick, hw = pm.shape[-2], pm.shape[-1]
            self.n_pick = n_pick
            # gridnet.py:79-91: [B,H,W,S:] -> [B, HW, n_pick] -> [B, n_pick, HW] -> [B*n_pick, HW]
            pl = logits[..., S:].reshape(logits.shape[0], -1, logits.shape[-1] - S).transpose(-1, -2)
            pl = pl.reshape(-1, pl.shape[-1])
            pm = pm.reshape(-1, hw)
            # gridnet.py:72-78,92-99: columns split by [HW]*n_pick (a single chunk for n_pick==1)
            self.picks = [
                MaskedLogits(lg, m)
                for lg, m in zip(torch.split(pl, [hw] * n_pick, dim=1), torch.split(pm, [hw] * n_pick, dim=1))
            ]

    def log_prob(self, action: TensorOrDict) -> torch.Tensor:
        cells = action["per_position"] if isinstance(action, dict) else action
        per_head_actions = cells.reshape(-1, cells.shape[-1]).T  # [A, B*HW]
        terms = []
        for h, (a, dist) in enumerate(zip(per_head_actions, self.heads)):
            lp = dist.log_prob(a)
            if h in self.gates:  # gridnet.py:119-127
                ref, required = self.gates[h]
                lp = torch.where(per_head_actions[ref] == required, lp, 0)
            terms.append(lp)
        total = torch.stack(terms, dim=-1).view(-1, self.map_size, len(self.nvec)).sum(dim=(1, 2))
        if isinstance(action, dict) and "pick_position" in action:
            assert self.picks is not None
            pa = action["pick_position"].view(-1, self.n_pick).T
            total = total + torch.stack([d.log_prob(a) for a, d in zip(pa, self.picks)], dim=-1).sum(dim=-1)
        return total

    def entropy(self) -> torch.Tensor:
        ent = (
            torch.stack([d.entropy() for d in self.heads], dim=-1)
            .view(-1, self.map_size, len(self.nvec))
            .sum(dim=(1, 2))
        )
        if self.picks:
            ent = ent + torch.stack([d.entropy() for d in self.picks], dim=-1).view(-1, self.n_pick).sum(dim=1)
        return ent

    def sample(self, generator: Optional[torch.Generator] = None) -> TensorOrDict:
        cells = torch.stack([d.sample(generator) for d in self.heads], dim=-1).view(
            -1, self.map_size, len(self.nvec)
        )
        if self.picks:
            return {
                "per_position": cells,
                "pick_position": torch.stack([d.sample(generator) for d in self.picks], dim=-1),
            }
        return cells


def gaussian_logp_entropy(
    mu: torch.Tensor, log_std: torch.Tensor, action: torch.Tensor
) -> Tuple[torch.Tensor, torch.Tensor]:
    """gaussian.py:11-16,42-45 + torch Normal: logp summed over act_dim, entropy NOT summed ([B, act_dim])."""
    std = torch.exp(log_std)
    mu_b, std_b = torch.broadcast_tensors(mu, std)
    var = std_b**2
    logp = -((action - mu_b) ** 2) / (2 * var) - std_b.log() - np.log(np.sqrt(2 * np.pi))
    entropy = 0.5 + 0.5 * np.log(2 * np.pi) + torch.log(std_b)
    return logp.sum(-1), entropy
