"""ACBC: actor-critic behaviour cloning on the device path.

Mirrors ``rl_algo_impls/acbc/acbc.py:28-166`` (constructor keywords, ``learn``): maximise the log-prob of
the rollout's (teacher / reference-bot) actions and regress the values onto the returns.  Rollout
buffer, K1 GAE, K3 gather and the heads' fused log-prob forward / backward are the PPO path's; the loss
itself (acbc.py:111-121) is a mean and a squared error.
"""
from time import perf_counter
from typing import Dict, List, Optional, TypeVar, Union

import numpy as np
import torch
import torch.nn as nn
from torch.optim import Adam

from ..algorithm import Algorithm, update_learning_rate
from ..ppo.ppo import _world, num_or_array

ACBCSelf = TypeVar("ACBCSelf", bound="ACBC")


class ACBC(Algorithm):
    def __init__(self, policy, device: torch.device, tb_writer=None, learning_rate: float = 3e-4, batch_size: int = 64,
                 n_epochs: int = 10, gamma=0.99, gae_lambda=0.95, vf_coef=0.25, max_grad_norm: float = 0.5,
                 gradient_accumulation: bool = False, scale_loss_by_num_actions: bool = False) -> None:
        super().__init__(policy, device, tb_writer, learning_rate, Adam(policy.parameters(), lr=learning_rate))
        self.batch_size, self.n_epochs = batch_size, n_epochs
        self.gamma, self.gae_lambda, self.vf_coef = num_or_array(gamma), num_or_array(gae_lambda), num_or_array(vf_coef)
        self.max_grad_norm = max_grad_norm
        self.gradient_accumulation = gradient_accumulation
        self.scale_loss_by_num_actions = scale_loss_by_num_actions
        self.last_stats: Optional[Dict[str, Union[float, np.ndarray]]] = None

    # -- one minibatch: imitation log-likelihood + value regression (acbc.py:111-121) -----------------------------
    def _minibatch(self, mb, value_weights: torch.Tensor, n_accumulate: int) -> torch.Tensor:
        """Backward of one minibatch; returns the detached [loss, pi_loss, v_loss...] row on the device."""
        self.policy.reset_noise(self.batch_size)
        logp, _, values = self.policy(mb.obs, mb.actions, action_masks=mb.action_masks)
        if self.scale_loss_by_num_actions:  # per-sample mean over the cells that had a legal action
            logp = torch.where(mb.num_actions > 0, logp / mb.num_actions, 0)
        imitation = -logp.mean()
        regression = (values - mb.returns).square().mean(0)
        objective = imitation + (value_weights * regression).sum()
        scaled = objective / n_accumulate if n_accumulate > 1 else objective  # acbc.py:123-124
        scaled.backward()
        return torch.cat([t.detach().reshape(-1).float() for t in (scaled, imitation, regression)])

    def _epochs(self, r) -> np.ndarray:
        """n_epochs passes over the rollout; mean of the LAST epoch's minibatch rows (one device -> host read)."""
        value_weights = torch.as_tensor(np.array(self.vf_coef), dtype=torch.float32, device=self.device)
        accumulate = r.num_minibatches(self.batch_size) if self.gradient_accumulation else 1
        rows: List[torch.Tensor] = []
        for _ in range(self.n_epochs):
            rows = []  # only the last epoch's rows are reported
            for mb in r.minibatches(self.batch_size, shuffle=not self.gradient_accumulation):
                rows.append(self._minibatch(mb, value_weights, accumulate))
                if accumulate == 1:
                    self.optimizer_step()
            if accumulate > 1:
                self.optimizer_step()
        return torch.stack(rows).double().mean(0).cpu().numpy()

    def _report(self, r, host: np.ndarray, seconds: float) -> None:
        spread = np.var(r.y_true).item()
        self.last_stats = {"loss": float(host[0]), "pi_loss": float(host[1]),
                           "v_loss": host[2:] if host.size > 3 else float(host[2]),
                           "explained_var": np.nan if spread == 0 else 1 - np.var(r.y_true - r.y_pred).item() / spread}
        if self.tb_writer is None:
            return
        for name, value in self.last_stats.items():
            for i, x in enumerate(np.atleast_1d(value)):
                self.tb_writer.add_scalar(f"losses/{name}" + (f"_{i}" if np.ndim(value) else ""), float(x))
        self.tb_writer.add_scalar("train/steps_per_second", r.total_steps / seconds)
        if hasattr(self.tb_writer, "on_steps"):
            self.tb_writer.on_steps(r.total_steps)

    def learn(self: ACBCSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> ACBCSelf:
        budget_end = start_timesteps + train_timesteps
        assert budget_end <= (train_timesteps if total_timesteps is None else total_timesteps)
        if self.scale_loss_by_num_actions and hasattr(rollout_generator, "include_num_actions"):
            rollout_generator.include_num_actions = True
        done = start_timesteps
        while done < budget_end:
            began = perf_counter()
            update_learning_rate(self.optimizer, self.learning_rate)
            r = rollout_generator.rollout(self.gamma, self.gae_lambda)
            done += r.total_steps
            self._report(r, self._epochs(r), perf_counter() - began)
            if callbacks and not all(c.on_step(timesteps_elapsed=r.total_steps) for c in callbacks):
                break
        return self

    def optimizer_step(self) -> None:
        params = [p for p in self.policy.parameters() if p.grad is not None]
        world = _world()
        if world > 1:  # envs sharded across ranks: mean gradient before the clip
            flat = torch._utils._flatten_dense_tensors([p.grad for p in params])
            torch.distributed.all_reduce(flat)
            flat.div_(world)
            for p, f in zip(params, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in params])):
                p.grad.copy_(f)
        nn.utils.clip_grad_norm_(params, self.max_grad_norm)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
