"""MaskedCategorical backed by the K4a kernels.

Mirrors ``rl_algo_impls/shared/actor/categorical.py:12-54`` (constructor keywords, ``log_prob``,
``entropy``, ``sample``, ``mode``): masked logits behave as ``finfo.min`` logits, a fully masked
row has log-prob 0 / entropy 0 / zero gradient.  ``log_prob`` and ``entropy`` come out of one
fused forward launch and share one fused backward launch.
"""
from typing import Optional

import torch

from .. import ops
from .rng import next_sample_stream


class MaskedCategorical:
    def __init__(
        self,
        probs=None,
        logits: Optional[torch.Tensor] = None,
        validate_args=None,
        mask: Optional[torch.Tensor] = None,
        verify: bool = False,
        neg_inf: Optional[float] = None,
    ):
        if probs is not None:
            raise ValueError("MaskedCategorical takes logits, not probs (as the reference with a mask)")
        assert logits is not None, "logits required"
        if neg_inf is not None and neg_inf != torch.finfo(logits.dtype).min:
            raise NotImplementedError("only the default neg_inf (finfo.min) is supported")
        self.batch_shape = logits.shape[:-1]
        self.n = logits.shape[-1]
        self.logits_raw = logits
        self.mask = mask
        self._cached = None  # (actions, logp, entropy) of the last fused forward

    def _rows(self):
        logits = self.logits_raw.reshape(-1, self.n)
        mask = self.mask.reshape(-1, self.n) if self.mask is not None else None
        return logits.float().contiguous(), mask

    def _forward(self, value: torch.Tensor):
        logits, mask = self._rows()
        actions = value.reshape(-1).contiguous()
        logp, ent = ops.categorical_logp_entropy(logits, mask, actions)
        self._cached = (value, logp.reshape(self.batch_shape), ent.reshape(self.batch_shape))
        return self._cached

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        return self._forward(value)[1]

    def entropy(self) -> torch.Tensor:
        if self._cached is None:
            zeros = torch.zeros(self.batch_shape, dtype=torch.int64, device=self.logits_raw.device)
            self._forward(zeros)
        return self._cached[2]

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        if len(sample_shape):
            raise NotImplementedError("only one draw per row is supported")
        logits, mask = self._rows()
        seed, offset = next_sample_stream()
        actions, _ = ops.categorical_sample(logits.detach(), mask, seed, offset)
        return actions.reshape(self.batch_shape)

    @property
    def mode(self) -> torch.Tensor:
        logits = self.logits_raw
        if self.mask is not None:
            logits = torch.where(self.mask, logits, torch.finfo(logits.dtype).min)
        return logits.argmax(dim=-1)
