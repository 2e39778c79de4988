"""TeacherKLLoss: the optional teacher-KL term of the PPO loss.

Mirrors ``rl_algo_impls/loss/teacher_kl_loss.py:10-50``: ``add_to_batch(batch)`` evaluates the
latest teacher checkpoint's log-prob of the rollout's actions (stored as
``Batch.additional["teacher_logprobs"]``), and the loss is ``mean(w * f(teacher_logp - logp))`` with
``f(d) = (e^d - 1) - d`` (``unbiased``) or ``d^2 / 2``.  With the fused kernels the term itself is
computed inside the loss launch (``b200rl_ppo_args.teacher_logp``); ``forward`` keeps the reference's
eager form for callers that hold the log-prob tensors themselves.
"""
from typing import Dict, Optional

import torch


def teacher_kl_term(delta: torch.Tensor, unbiased: bool) -> torch.Tensor:
    """Per-sample penalty of delta = teacher_logp - logp: the k3 estimator (e^delta - 1) - delta, which is
    non-negative and unbiased for KL(policy || teacher), or the second-order delta^2 / 2 (teacher_kl_loss.py:41-46)."""
    return (torch.exp(delta) - 1) - delta if unbiased else 0.5 * delta**2  # exp - 1, not expm1: the reference's rounding


class TeacherKLLoss(torch.nn.Module):
    """Constructor keywords and the two entry points of the reference class; ``reduction`` only accepts "mean"."""

    FIELD = "teacher_logprobs"  # key under Batch.additional

    def __init__(self, ckpts_manager, unbiased: bool = True, reduction: str = "mean") -> None:
        super().__init__()
        if reduction != "mean":
            raise AssertionError(f"reduction must be 'mean', got {reduction}")
        self.ckpts_manager, self.unbiased, self.reduction = ckpts_manager, bool(unbiased), reduction

    @torch.no_grad()
    def add_to_batch(self, batch) -> Dict[str, torch.Tensor]:
        """Log-prob of the rollout's actions under the latest teacher checkpoint, one row per sample."""
        checkpoint = self.ckpts_manager.latest_checkpoint
        assert checkpoint is not None, "No checkpoints available"
        result = checkpoint(batch.obs, batch.actions, action_masks=batch.action_masks)
        logp = result.logp_a if hasattr(result, "logp_a") else result[0]
        return {self.FIELD: logp.float().contiguous()}

    def forward(self, training_logprobs: torch.Tensor, mb_additional: Dict[str, torch.Tensor],
                weights: Optional[torch.Tensor]) -> torch.Tensor:
        term = teacher_kl_term(mb_additional[self.FIELD] - training_logprobs, self.unbiased)
        return (term if weights is None else term * weights).mean()
