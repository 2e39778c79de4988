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


class TeacherKLLoss(torch.nn.Module):
    def __init__(self, ckpts_manager, unbiased: bool = True, reduction: str = "mean") -> None:
        super().__init__()
        assert reduction == "mean", f"reduction must be 'mean', got {reduction}"
        self.ckpts_manager = ckpts_manager
        self.unbiased = unbiased
        self.reduction = reduction

    def add_to_batch(self, batch) -> Dict[str, torch.Tensor]:
        teacher = self.ckpts_manager.latest_checkpoint
        assert teacher is not None, "No checkpoints available"
        with torch.no_grad():
            out = teacher(batch.obs, batch.actions, action_masks=batch.action_masks)
        return {"teacher_logprobs": (out.logp_a if hasattr(out, "logp_a") else out[0]).float().contiguous()}

    def forward(self, training_logprobs: torch.Tensor, mb_additional: Dict[str, torch.Tensor],
                weights: Optional[torch.Tensor]) -> torch.Tensor:
        logratio = mb_additional["teacher_logprobs"] - training_logprobs
        loss = (torch.exp(logratio) - 1) - logratio if self.unbiased else 0.5 * logratio**2
        if weights is not None:
            loss = loss * weights
        return loss.mean()
