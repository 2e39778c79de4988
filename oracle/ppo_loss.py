"""Oracle: per-minibatch advantage normalisation and the PPO loss.  TEST INFRASTRUCTURE.

torch-CPU restatement of ``rl_algo_impls/ppo/ppo.py``:
  * :307-318  advantage normalise / standardise / reward-weight contraction
  * :326-361  ratio, clipped surrogate, (clipped) value loss, entropy loss, approx-KL, KL cut-off, total
  * :373-374  division by the number of minibatches under gradient accumulation
  * :379-409  clipped fractions and per-step stats
Checked against the live reference ``PPO.learn_epoch`` by tests/golden/make_golden.py.
"""
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def normalize_advantages(
    adv: torch.Tensor,  # [B] or [B, V]
    *,
    normalize_advantage: bool = True,
    standardize_advantage: bool = False,
    normalize_advantages_after_scaling: bool = False,
    multi_reward_weights: Optional[torch.Tensor] = None,  # [V]
) -> torch.Tensor:
    if normalize_advantages_after_scaling:  # ppo.py:307-311
        if multi_reward_weights is not None:
            adv = adv @ multi_reward_weights
        return (adv - adv.mean()) / (adv.std() + 1e-8)
    if normalize_advantage:  # ppo.py:313-314 (torch.std is the unbiased estimator)
        adv = (adv - adv.mean(0)) / (adv.std(0) + 1e-8)
    elif standardize_advantage:  # ppo.py:315-316
        adv = adv / (adv.std(0) + 1e-8)
    if multi_reward_weights is not None:  # ppo.py:317-318
        adv = adv @ multi_reward_weights
    return adv


@dataclass
class LossParts:
    loss: torch.Tensor
    pi_loss: torch.Tensor
    v_loss: torch.Tensor  # [V] or scalar
    entropy_loss: torch.Tensor
    approx_kl: float
    clipped_frac: float
    val_clipped_frac: np.ndarray
    pi_coef: float
    ratio: torch.Tensor = field(repr=False, default=None)
    teacher_kl_loss: Optional[torch.Tensor] = None


def ppo_loss(
    new_logprobs: torch.Tensor,  # [B]
    entropy: torch.Tensor,  # [B] (or [B, act_dim] for Gaussian)
    new_values: torch.Tensor,  # [B] or [B, V]
    old_logprobs: torch.Tensor,
    adv: torch.Tensor,  # [B], already normalised / contracted
    old_values: torch.Tensor,
    returns: torch.Tensor,
    *,
    clip_range: float,
    clip_range_vf: Optional[float],
    ent_coef: float,
    vf_coef: torch.Tensor,  # scalar tensor or [V]
    ppo2_vf_coef_halving: bool = False,
    vf_weights: Optional[torch.Tensor] = None,
    pi_coef: float = 1,
    kl_cutoff: Optional[float] = None,
    loss_divisor: Optional[int] = None,  # num_minibatches under gradient accumulation
    vf_loss_fn=F.mse_loss,
    teacher_logprobs: Optional[torch.Tensor] = None,  # loss/teacher_kl_loss.py:35-50 + ppo.py:363-371
    teacher_kl_loss_coef: Optional[float] = None,
    teacher_unbiased: bool = True,
    teacher_loss_importance_sampling: bool = True,
) -> LossParts:
    logratio = new_logprobs - old_logprobs
    ratio = torch.exp(logratio)
    clipped_ratio = torch.clamp(ratio, min=1 - clip_range, max=1 + clip_range)
    pi_loss = -torch.min(ratio * adv, clipped_ratio * adv).mean()

    v_unclipped = vf_loss_fn(new_values, returns, reduction="none")
    if clip_range_vf is not None:
        v_clipped = vf_loss_fn(
            old_values + torch.clamp(new_values - old_values, -clip_range_vf, clip_range_vf),
            returns,
            reduction="none",
        )
        v_loss = torch.max(v_unclipped, v_clipped)
    else:
        v_loss = v_unclipped
    if vf_weights is not None:
        v_loss = v_loss @ vf_weights
    v_loss = v_loss.mean(0)
    if ppo2_vf_coef_halving:
        v_loss = v_loss * 0.5

    entropy_loss = -entropy.mean()
    with torch.no_grad():
        approx_kl = ((ratio - 1) - logratio).mean().cpu().numpy().item()
    if kl_cutoff is not None and approx_kl > kl_cutoff:  # sticky within a learn_epoch (ppo.py:279,354-355)
        pi_coef = 0

    loss = pi_coef * pi_loss + ent_coef * entropy_loss + (vf_coef * v_loss).sum()
    teacher_kl_loss = None
    if teacher_kl_loss_coef:
        t_logratio = teacher_logprobs - new_logprobs
        t_loss = (torch.exp(t_logratio) - 1) - t_logratio if teacher_unbiased else 0.5 * t_logratio**2
        if teacher_loss_importance_sampling:
            t_loss = t_loss * ratio  # not detached, as the reference
        teacher_kl_loss = t_loss.mean()
        loss = loss + teacher_kl_loss_coef * teacher_kl_loss
    if loss_divisor is not None:
        loss = loss / loss_divisor

    with torch.no_grad():
        clipped_frac = ((ratio - 1).abs() > clip_range).float().mean().item()
        if clip_range_vf is not None:
            val_clipped_frac = ((new_values - old_values).abs() > clip_range_vf).float().mean(0).cpu().numpy()
        else:
            val_clipped_frac = np.zeros(v_loss.shape)
    parts = LossParts(
        loss, pi_loss, v_loss, entropy_loss, approx_kl, clipped_frac, val_clipped_frac, pi_coef, ratio.detach()
    )
    parts.teacher_kl_loss = teacher_kl_loss
    return parts
