"""A2C over the same device-resident rollout and distribution kernels as PPO.

Mirrors ``rl_algo_impls/a2c/a2c.py:23-229`` (constructor keywords, ``learn``, ``optimizer_step``) and
``a2c/train_stats.py``.  The rollout (HBM buffers, graph-replayed steps, K1 GAE), the minibatch gather
(K3), the advantage normalisation (K2) and the heads' log-prob / entropy forward + backward (K4 fwd / bwd
through ``ActorCritic.forward``) are the PPO path's; what is left of the A2C loss (a2c.py:147-161) is four
elementwise / reduction ops on ``[B]`` tensors, kept in torch.
"""
import logging
from dataclasses import asdict, dataclass
from time import perf_counter
from typing import Dict, List, Optional, TypeVar, Union

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..algorithm import Algorithm, update_learning_rate
from ..ppo.ppo import _world, num_or_array

A2CSelf = TypeVar("A2CSelf", bound="A2C")


@dataclass
class TrainStepStats:
    loss: float
    pi_loss: float
    v_loss: np.ndarray
    entropy_loss: float


class TrainStats:
    data: Dict[str, Union[float, np.ndarray]]

    def __init__(self, step_stats: List[TrainStepStats], explain_var: float) -> None:
        self.data = {"explained_var": explain_var}
        for k in asdict(step_stats[0]).keys():
            if isinstance(getattr(step_stats[0], k), np.ndarray):
                self.data[k] = np.mean([getattr(s, k) for s in step_stats], axis=0)
            else:
                self.data[k] = np.mean([getattr(s, k) for s in step_stats]).item()

    def write_to_tensorboard(self, tb_writer) -> None:
        for name, value in self.data.items():
            if isinstance(value, np.ndarray):
                for idx, v in enumerate(value.flatten()):
                    tb_writer.add_scalar(f"losses/{name}_{idx}", v)
            else:
                tb_writer.add_scalar(f"losses/{name}", value)


class A2C(Algorithm):
    def __init__(
        self,
        policy,
        device: torch.device,
        tb_writer=None,
        learning_rate: float = 7e-4,
        gamma=0.99,
        gae_lambda=1.0,
        ent_coef: float = 0.0,
        vf_coef=0.5,
        max_grad_norm: float = 0.5,
        rms_prop_eps: float = 1e-5,
        use_rms_prop: bool = True,
        normalize_advantage: bool = False,
        multi_reward_weights: Optional[List[float]] = None,
        scale_loss_by_num_actions: bool = False,
        gradient_accumulation: bool = False,
        autocast_loss: bool = False,
        num_minibatches: Optional[int] = None,
    ) -> None:
        if use_rms_prop:
            optimizer = torch.optim.RMSprop(policy.parameters(), lr=learning_rate, eps=rms_prop_eps)
        else:
            optimizer = torch.optim.Adam(policy.parameters(), lr=learning_rate)
        super().__init__(policy, device, tb_writer, learning_rate, optimizer)
        self.policy = policy
        self.gamma = num_or_array(gamma)
        self.gae_lambda = num_or_array(gae_lambda)
        self.vf_coef = num_or_array(vf_coef)
        self.ent_coef = ent_coef
        self.max_grad_norm = max_grad_norm
        self.normalize_advantage = normalize_advantage
        self.multi_reward_weights = np.array(multi_reward_weights) if multi_reward_weights else None
        self.scale_loss_by_num_actions = scale_loss_by_num_actions
        self.gradient_accumulation = gradient_accumulation
        self.autocast_loss = autocast_loss
        self.num_minibatches = num_minibatches or 1
        assert self.num_minibatches == 1 or self.gradient_accumulation, (
            "A2C only supports single step batches. Therefore, non-1 minibatches must be gradient accumulated")
        self.last_train_stats: Optional[TrainStats] = None

    def learn(self: A2CSelf, train_timesteps: int, rollout_generator, callbacks: Optional[List] = None,
              total_timesteps: Optional[int] = None, start_timesteps: int = 0) -> A2CSelf:
        if total_timesteps is None:
            total_timesteps = train_timesteps
        assert start_timesteps + train_timesteps <= total_timesteps
        timesteps_elapsed = start_timesteps
        while timesteps_elapsed < start_timesteps + train_timesteps:
            start_time = perf_counter()
            update_learning_rate(self.optimizer, self.learning_rate)
            if self.scale_loss_by_num_actions and hasattr(rollout_generator, "include_num_actions"):
                rollout_generator.include_num_actions = True
            r = rollout_generator.rollout(gamma=self.gamma, gae_lambda=self.gae_lambda)
            timesteps_elapsed += r.total_steps

            vf_coef = torch.as_tensor(np.array(self.vf_coef), dtype=torch.float32, device=self.device)
            step_stats = [self._minibatch(mb, vf_coef)
                          for mb in r.minibatches(r.total_steps // self.num_minibatches,
                                                  shuffle=not self.gradient_accumulation)]
            if self.gradient_accumulation:
                self.optimizer_step()

            host = torch.stack(step_stats).double().cpu().numpy()  # one device -> host read per iteration
            V = host.shape[1] - 3
            steps = [TrainStepStats(float(x[0]), float(x[1]),
                                    np.asarray(x[3:], np.float64) if V > 1 else np.asarray(x[3], np.float64), float(x[2]))
                     for x in host]
            var_y = np.var(r.y_true).item()
            explained_var = np.nan if var_y == 0 else 1 - np.var(r.y_true - r.y_pred).item() / var_y
            stats = TrainStats(steps, explained_var)
            self.last_train_stats = stats
            rollout_steps = r.total_steps
            if self.tb_writer is not None:
                self.tb_writer.add_scalar("train/steps_per_second", rollout_steps / (perf_counter() - start_time))
                stats.write_to_tensorboard(self.tb_writer)
                if hasattr(self.tb_writer, "on_steps"):
                    self.tb_writer.on_steps(rollout_steps)
            if callbacks:
                if not all(c.on_step(timesteps_elapsed=rollout_steps) for c in callbacks):
                    logging.info(f"Callback terminated training at {timesteps_elapsed} timesteps")
                    break
        return self

    def _advantages(self, adv: torch.Tensor) -> torch.Tensor:
        """a2c.py:133-138 on the device: K2 moments + normalise, then the reward-weight contraction."""
        weights = None if self.multi_reward_weights is None else [float(x) for x in self.multi_reward_weights]
        if not self.normalize_advantage and weights is None:
            return adv
        B = adv.shape[0]
        mode = ops.ADV_NORMALIZE if self.normalize_advantage else ops.ADV_NONE
        moments = None
        if self.normalize_advantage:
            moments = ops.adv_moments(adv.reshape(B, -1), None, mode, weights)
            if _world() > 1:
                torch.distributed.all_reduce(moments)
        return ops.adv_normalize(adv.reshape(B, -1), None, mode, weights, moments).reshape(B)

    def _minibatch(self, mb, vf_coef: torch.Tensor) -> torch.Tensor:
        """Forward, A2C loss (a2c.py:140-161), backward (and the optimizer step unless gradients
        accumulate).  Returns [loss, pi_loss, entropy_loss, v_loss...] on the device."""
        adv = self._advantages(mb.advantages)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(self.autocast_loss)):
            logp_a, entropy, v = self.policy(mb.obs, mb.actions, action_masks=mb.action_masks)
            if self.scale_loss_by_num_actions:
                logp_a = torch.where(mb.num_actions > 0, logp_a / mb.num_actions, 0)
            pi_loss = -(adv * logp_a).mean()
            value_loss = (v - mb.returns).square().mean(0)
            entropy_loss = -entropy.mean()
            loss = pi_loss + (vf_coef * value_loss).sum() + self.ent_coef * entropy_loss
            if self.gradient_accumulation:
                loss = loss / self.num_minibatches
        loss.backward()
        if not self.gradient_accumulation:
            self.optimizer_step()
        return torch.cat([t.detach().reshape(-1).float() for t in (loss, pi_loss, entropy_loss, value_loss)])

    def optimizer_step(self) -> None:
        params = [p for p in self.policy.parameters() if p.grad is not None]
        world = _world()
        if world > 1:  # envs sharded across ranks: average the gradients before the clip
            flat = torch._utils._flatten_dense_tensors([p.grad for p in params])
            torch.distributed.all_reduce(flat)
            flat.div_(world)
            for p, f in zip(params, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in params])):
                p.grad.copy_(f)
        nn.utils.clip_grad_norm_(params, self.max_grad_norm)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
