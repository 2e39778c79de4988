"""GridnetDistribution backed by the K4c / K5 kernels.

Mirrors ``rl_algo_impls/shared/actor/gridnet.py:22-222``: same constructor
``(map_size, action_vec, logits, masks, validate_args=None, subaction_mask=None)``, same
``log_prob`` / ``entropy`` / ``sample`` / ``mode`` surface, masks and actions either plain
tensors or dicts with ``per_position`` / ``pick_position``.  Where the reference builds one
``MaskedCategorical`` per action plane (7 for MicroRTS, 6 + pick for Lux) and stacks their
results, this class issues one fused forward launch for (log_prob, entropy) together and one
fused backward launch.
"""
from typing import Dict, NamedTuple, Optional, Type, TypeVar, Union

import numpy as np
import torch

from .. import ops
from .rng import next_sample_stream

TensorOrDict = Union[torch.Tensor, Dict[str, torch.Tensor]]
ValueDependentMaskSelf = TypeVar("ValueDependentMaskSelf", bound="ValueDependentMask")


class ValueDependentMask(NamedTuple):
    """gridnet.py:22-35"""

    reference_index: int
    value: int

    @classmethod
    def from_reference_index_to_index_to_value(
        cls: Type[ValueDependentMaskSelf], ref_idx_to_idx_to_value: Dict[int, Dict[int, int]]
    ) -> Dict[int, ValueDependentMaskSelf]:
        return {
            idx: cls(ref_idx, value)
            for ref_idx, idx_to_value in ref_idx_to_idx_to_value.items()
            for idx, value in idx_to_value.items()
        }


def _spec(action_vec, subaction_mask, n_pick: int) -> ops.GridnetSpec:
    gates = []
    for head, vdm in (subaction_mask or {}).items():
        ref, value = vdm  # ValueDependentMask or a (reference_index, value) pair
        gates.append((int(head), int(ref), int(value)))
    return ops.GridnetSpec(tuple(int(n) for n in np.asarray(action_vec).reshape(-1)), tuple(sorted(gates)), n_pick)


class GridnetDistribution:
    def __init__(
        self,
        map_size: int,
        action_vec,
        logits: torch.Tensor,
        masks: TensorOrDict,
        validate_args: Optional[bool] = None,
        subaction_mask: Optional[Dict[int, ValueDependentMask]] = None,
    ) -> None:
        self.map_size = int(map_size)
        self.action_vec = np.asarray(action_vec).reshape(-1)
        self.subaction_mask = subaction_mask
        per_position = masks["per_position"] if isinstance(masks, dict) else masks
        self.pick_mask = masks.get("pick_position") if isinstance(masks, dict) else None
        n_pick = int(self.pick_mask.shape[-2]) if self.pick_mask is not None else 0
        self.spec = _spec(self.action_vec, subaction_mask, n_pick)
        S = int(self.action_vec.sum())
        B = logits.shape[0]
        self.logits = logits.reshape(B, self.map_size, logits.shape[-1])
        if not self.logits.is_contiguous():
            self.logits = self.logits.contiguous()
        self.mask = per_position.reshape(B, self.map_size, S)
        self.batch_shape = logits.shape[:-1]
        self._cached = None

    # -- helpers ---------------------------------------------------------------------------------
    def _split_action(self, action: TensorOrDict):
        if isinstance(action, dict):
            return action["per_position"], action.get("pick_position")
        return action, None

    def _forward(self, action: TensorOrDict):
        cells, pick = self._split_action(action)
        B = self.logits.shape[0]
        cells = cells.reshape(B, self.map_size, len(self.action_vec)).contiguous()
        if self.spec.n_pick:
            assert pick is not None, "pick_position actions required"
            pick = pick.reshape(B, self.spec.n_pick).contiguous()
        logp, ent = ops.gridnet_logp_entropy(self.spec, self.logits, self.mask, self.pick_mask, cells, pick)
        self._cached = (logp, ent)
        return self._cached

    # -- Distribution surface -------------------------------------------------------------------
    def log_prob(self, action: TensorOrDict) -> torch.Tensor:
        return self._forward(action)[0]

    def entropy(self) -> torch.Tensor:
        if self._cached is None:  # entropy does not depend on the action: use any
            B = self.logits.shape[0]
            dev = self.logits.device
            cells = torch.zeros((B, self.map_size, len(self.action_vec)), dtype=torch.uint8, device=dev)
            action: TensorOrDict = cells
            if self.spec.n_pick:
                action = {
                    "per_position": cells,
                    "pick_position": torch.zeros((B, self.spec.n_pick), dtype=torch.int64, device=dev),
                }
            self._forward(action)
        return self._cached[1]

    def sample_with_log_prob(self, act_dtype: torch.dtype = torch.int64):
        """One launch: actions (+ pick) and their joint log-prob (rollout-time fast path)."""
        seed, offset = next_sample_stream()
        cells, pick, logp = ops.gridnet_sample(
            self.spec, self.logits.detach(), self.mask, self.pick_mask, seed, offset, act_dtype
        )
        action: TensorOrDict = cells
        if self.spec.n_pick:
            action = {"per_position": cells, "pick_position": pick}
        return action, logp

    def sample(self, sample_shape: torch.Size = torch.Size()) -> TensorOrDict:
        if len(sample_shape):
            raise NotImplementedError("only one draw per cell is supported")
        return self.sample_with_log_prob()[0]

    @property
    def mode(self) -> TensorOrDict:
        # evaluation-time arg-max (gridnet.py:208-218); not on the training path
        S = int(self.action_vec.sum())
        masked = torch.where(self.mask.bool(), self.logits[..., :S].float(), torch.finfo(torch.float32).min)
        cells = torch.stack(
            [chunk.argmax(dim=-1) for chunk in torch.split(masked, self.action_vec.tolist(), dim=-1)], dim=-1
        )
        if self.spec.n_pick:
            pick_logits = self.logits[..., S:S + self.spec.n_pick].float().transpose(-1, -2)
            pick_logits = torch.where(self.pick_mask.bool(), pick_logits, torch.finfo(torch.float32).min)
            return {"per_position": cells, "pick_position": pick_logits.argmax(dim=-1)}
        return cells

    @property
    def arg_constraints(self):
        return {}


def num_actions_device(actions, action_masks, subaction_mask, action_plane_space) -> Optional[torch.Tensor]:
    """rollout.py:130-180 over the whole device-resident rollout in ONE launch (ops.gridnet_num_actions; only built on
    request: PPO ignores the field, ppo.py:300).  Same result types as the reference: an integer count per step, or
    float32 count + log(#pick cells) for a dict action space."""
    if action_masks is None:
        return None
    cells_mask = action_masks["per_position"] if isinstance(action_masks, dict) else action_masks
    pick_mask = action_masks.get("pick_position") if isinstance(action_masks, dict) else None
    cells_act = actions["per_position"] if isinstance(actions, dict) else actions
    lead = tuple(cells_mask.shape[:-2])  # [T, N] or [M]
    HW, S = cells_mask.shape[-2], cells_mask.shape[-1]
    if subaction_mask:
        assert action_plane_space is not None
        nvec = [int(n) for n in action_plane_space.nvec]
        gates = subaction_mask if not isinstance(next(iter(subaction_mask.values())), dict) else \
            ValueDependentMask.from_reference_index_to_index_to_value(subaction_mask)
    else:
        nvec, gates = [int(S)], None  # no gating: "any valid entry in the cell" is one plane of S entries
    n_pick = int(pick_mask.shape[-2]) if pick_mask is not None else 0
    spec = _spec(nvec, gates, n_pick)
    R = int(np.prod(lead)) if lead else 1
    act = cells_act.reshape(R, HW, -1).contiguous() if gates else None
    cells, picks = ops.gridnet_num_actions(spec, cells_mask.reshape(R, HW, S).contiguous(),
                                           pick_mask.reshape(R, n_pick, HW).contiguous() if n_pick else None, act)
    if isinstance(action_masks, dict):
        p = picks.double()
        out = (cells.double() + torch.where(p > 0, torch.log(p.clamp_min(1.0)), torch.zeros_like(p))).float()
        return out.reshape(lead)
    return (cells if gates else cells.long()).reshape(lead)
