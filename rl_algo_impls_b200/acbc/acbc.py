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

    def learn(self: ACBCSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> ACBCSelf:
        total_timesteps = train_timesteps if total_timesteps is None else total_timesteps
        assert start_timesteps + train_timesteps <= total_timesteps
        elapsed = start_timesteps
        while elapsed < start_timesteps + train_timesteps:
            t0 = perf_counter()
            update_learning_rate(self.optimizer, self.learning_rate)
            if self.scale_loss_by_num_actions and hasattr(rollout_generator, "include_num_actions"):
                rollout_generator.include_num_actions = True
            r = rollout_generator.rollout(self.gamma, self.gae_lambda)
            elapsed += r.total_steps
            vf_coef = torch.as_tensor(np.array(self.vf_coef), dtype=torch.float32, device=self.device)
            n_mb = r.num_minibatches(self.batch_size)
            rows: List[torch.Tensor] = []
            for _ in range(self.n_epochs):
                rows.clear()  # the last epoch's losses are the ones reported
                for mb in r.minibatches(self.batch_size, shuffle=not self.gradient_accumulation):
                    self.policy.reset_noise(self.batch_size)
                    logp, _, values = self.policy(mb.obs, mb.actions, action_masks=mb.action_masks)
                    if self.scale_loss_by_num_actions:
                        logp = torch.where(mb.num_actions > 0, logp / mb.num_actions, 0)
                    pi_loss = -logp.mean()
                    v_loss = (values - mb.returns).square().mean(0)
                    loss = pi_loss + (vf_coef * v_loss).sum()
                    if self.gradient_accumulation:
                        loss = loss / n_mb
                    loss.backward()
                    if not self.gradient_accumulation:
                        self.optimizer_step()
                    rows.append(torch.cat([t.detach().reshape(-1).float() for t in (loss, pi_loss, v_loss)]))
                if self.gradient_accumulation:
                    self.optimizer_step()
            host = torch.stack(rows).double().mean(0).cpu().numpy()  # one device -> host read per iteration
            var_y = np.var(r.y_true).item()
            self.last_stats = {"loss": float(host[0]), "pi_loss": float(host[1]),
                               "v_loss": host[2:] if host.size > 3 else float(host[2]),
                               "explained_var": np.nan if var_y == 0 else 1 - np.var(r.y_true - r.y_pred).item() / var_y}
            if self.tb_writer is not None:
                for k, v in self.last_stats.items():
                    for i, x in enumerate(np.atleast_1d(v)):
                        self.tb_writer.add_scalar(f"losses/{k}" + (f"_{i}" if np.ndim(v) else ""), float(x))
                self.tb_writer.add_scalar("train/steps_per_second", r.total_steps / (perf_counter() - t0))
                if hasattr(self.tb_writer, "on_steps"):
                    self.tb_writer.on_steps(r.total_steps)
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
