"""Tiny deterministic trunks shared by the golden generator (live reference), the oracle learner
tests (CPU) and the learner parity tests (GPU).  Plain torch modules; outputs carry ``pi``,
``values`` and ``log_std`` like rl_algo_impls_b200.policy.networks.HeadOutputs."""
from types import SimpleNamespace

import torch
import torch.nn as nn


class TinyMlp(nn.Module):
    def __init__(self, obs_dim: int, pi_out: int, n_values: int = 1, gaussian: bool = False, hidden: int = 16):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh())
        self.pi, self.v = nn.Linear(hidden, pi_out), nn.Linear(hidden, n_values)
        self.log_std = nn.Parameter(torch.full((pi_out,), -1.0)) if gaussian else None
        self.n_values = n_values

    def forward(self, obs):
        x = self.body(obs.to(self.pi.weight.dtype).reshape(obs.shape[0], -1))  # f32; f64 in the generators' float64 runs
        v = self.v(x)
        return SimpleNamespace(pi=self.pi(x), values=v.squeeze(-1) if self.n_values == 1 else v, log_std=self.log_std)


class TinyGrid(nn.Module):
    """[B, C, H, W] -> logits [B, H, W, S'] with one 3x3 conv, values from a pooled linear head."""

    def __init__(self, in_channels: int, n_logits: int, n_values: int = 1, hidden: int = 8):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, hidden, 3, padding=1)
        self.actor = nn.Conv2d(hidden, n_logits, 3, padding=1)
        self.v = nn.Linear(hidden, n_values)
        self.n_values = n_values

    def forward(self, obs):
        x = torch.tanh(self.conv(obs.to(self.conv.weight.dtype)))
        v = self.v(x.mean(dim=(2, 3)))
        return SimpleNamespace(pi=self.actor(x).permute(0, 2, 3, 1), values=v.squeeze(-1) if self.n_values == 1 else v,
                               log_std=None)
