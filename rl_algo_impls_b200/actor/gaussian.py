"""Diagonal Gaussian action distribution (K4b forward kernel).

Mirrors ``rl_algo_impls/shared/actor/gaussian.py:11-16``: ``log_prob`` is summed over the
action dimension, ``entropy`` is per dimension ([B, act_dim], not summed).  The differentiable
path used by the learner is the fused ``ops.ppo_gaussian_loss``; this class serves rollouts
(sample + log-prob) and the distribution surface.
"""
import torch

from .. import ops


class GaussianDistribution:
    def __init__(self, loc: torch.Tensor, scale: torch.Tensor):
        self.loc = loc
        self.scale = scale

    def sample(self) -> torch.Tensor:
        # the reference samples with rsample (gaussian.py:15-16): loc + scale * eps
        return self.loc + self.scale * torch.randn_like(self.loc)

    @property
    def mode(self) -> torch.Tensor:
        return self.loc

    def _fwd(self, a: torch.Tensor):
        mu = self.loc.detach().float().contiguous()
        log_std = torch.log(self.scale.detach().float()).reshape(-1).contiguous()
        return ops.gaussian_logp_entropy(mu.reshape(-1, mu.shape[-1]), log_std, a.float().reshape(-1, mu.shape[-1]).contiguous())

    def log_prob(self, a: torch.Tensor) -> torch.Tensor:
        return self._fwd(a)[0].reshape(self.loc.shape[:-1])

    def entropy(self) -> torch.Tensor:
        return self._fwd(torch.zeros_like(self.loc))[1].reshape(self.loc.shape)
