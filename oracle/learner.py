"""Oracle: one PPO ``learn_epoch`` on the CPU -- rollout loop, GAE, minibatch loop, optimizer step.
TEST INFRASTRUCTURE (see oracle/__init__.py); also the timed CPU baseline of bench.py.

Restates, with torch-CPU / numpy (the reference's own arithmetic libraries):
  * ``rl_algo_impls/rollout/sync_step_rollout.py:181-216``  the host rollout loop into numpy [T, N, ...] buffers
  * ``rl_algo_impls/shared/policy/actor_critic.py:306-318``  policy.step: sample + log_prob, no grad
  * ``rl_algo_impls/rollout/vec_rollout.py:38-175``          GAE, returns, flatten, randperm minibatches, gather
  * ``rl_algo_impls/ppo/ppo.py:214-447``                     the minibatch loop, stats and optimizer step
The policy trunk is whatever ``nn.Module`` the caller passes (it returns an object with ``pi``,
``values`` and optionally ``log_std``); heads are the oracle distributions of oracle/distributions.py.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .distributions import Gridnet, MaskedLogits, gates_from_subaction_mask, gaussian_logp_entropy
from .gae import gae_advantages, gae_returns
from .ppo_loss import normalize_advantages, ppo_loss
from .rollout import flatten_time_major, minibatch_index_stream


@dataclass
class Hyper:
    """The reference's PPO keyword arguments (ppo/ppo.py:107-140), same names and defaults."""

    batch_size: int = 64
    n_epochs: int = 10
    gamma: Union[float, np.ndarray] = 0.99
    gae_lambda: Union[float, np.ndarray] = 0.95
    clip_range: float = 0.2
    clip_range_vf: Optional[float] = None
    normalize_advantage: bool = True
    standardize_advantage: bool = False
    ent_coef: float = 0.0
    vf_coef: Union[float, Sequence[float]] = 0.5
    ppo2_vf_coef_halving: bool = False
    max_grad_norm: float = 0.5
    multi_reward_weights: Optional[Sequence[float]] = None
    gradient_accumulation: bool = False
    kl_cutoff: Optional[float] = None
    normalize_advantages_after_scaling: bool = False
    learning_rate: float = 3e-4
    teacher_kl_loss_coef: Optional[float] = None
    teacher_unbiased: bool = True
    teacher_loss_importance_sampling: bool = True
    vf_loss_fn: str = "mse_loss"  # ppo.py:135,186: getattr(torch.nn.functional, vf_loss_fn)
    vf_weights: Optional[Sequence[float]] = None  # ppo.py:344-345: v_loss @ vf_weights before the batch mean
    # ppo.py:321 + shared/autocast.py:8-12: bf16 autocast around forward + loss -- ONLY on a CUDA device with bf16
    # support; on the CPU (where this oracle and the live reference run) the flag changes nothing
    autocast_loss: bool = False
    logp_shift: float = 0.0  # fixture knob, not a reference keyword: added to the stored behaviour log-probs


class OraclePolicy:
    """forward / step of shared/policy/actor_critic.py over a caller-supplied trunk."""

    def __init__(self, network: nn.Module, kind: str, nvec: Sequence[int] = (), map_size: int = 0,
                 subaction_mask: Optional[Dict[int, Dict[int, int]]] = None):
        self.network, self.kind, self.nvec, self.map_size = network, kind, tuple(nvec), map_size
        self.gates = gates_from_subaction_mask(subaction_mask)

    def parameters(self):
        return self.network.parameters()

    def _dist(self, out, masks):
        if self.kind == "gridnet":
            return Gridnet(self.map_size, self.nvec, out.pi, masks, self.gates)
        if self.kind == "categorical":
            return MaskedLogits(out.pi, masks)
        return None

    def forward(self, obs, actions, masks):
        out = self.network(obs)
        if self.kind == "gaussian":
            logp, ent = gaussian_logp_entropy(out.pi, out.log_std, actions)
            return logp, ent, out.values
        d = self._dist(out, masks)
        return d.log_prob(actions), d.entropy(), out.values

    @torch.no_grad()
    def step(self, obs, masks):
        out = self.network(obs)
        if self.kind == "gaussian":
            a = out.pi + torch.exp(out.log_std) * torch.randn_like(out.pi)
            return a, out.values, gaussian_logp_entropy(out.pi, out.log_std, a)[0]
        d = self._dist(out, masks)
        a = d.sample()
        return a, out.values, d.log_prob(a)

    @torch.no_grad()
    def value(self, obs):
        return self.network(obs).values


def _t(a):
    if a is None:
        return None
    if isinstance(a, dict):
        return {k: _t(v) for k, v in a.items()}
    return torch.as_tensor(a)


def collect_rollout(policy: OraclePolicy, env, n_steps: int, state: dict) -> dict:
    """sync_step_rollout.py:181-216 into freshly allocated numpy buffers.  ``state`` carries
    next_obs / next_action_masks / next_episode_starts between rollouts."""
    N = env.num_envs
    if not state:
        state["next_obs"], _ = env.reset()
        state["next_masks"] = env.get_action_mask()
        state["next_starts"] = np.ones(N, dtype=np.bool_)
    bufs: Dict[str, list] = {k: [] for k in ("obs", "episode_starts", "masks", "actions", "values", "logprobs", "rewards")}
    for _ in range(n_steps):
        obs, masks = state["next_obs"], state["next_masks"]
        bufs["obs"].append(np.asarray(obs)), bufs["episode_starts"].append(np.asarray(state["next_starts"]))
        bufs["masks"].append(masks)
        a, v, logp = policy.step(_t(np.asarray(obs)), _t(masks))
        a_np = {k: x.numpy() for k, x in a.items()} if isinstance(a, dict) else a.numpy()
        bufs["actions"].append(a_np), bufs["values"].append(v.numpy()), bufs["logprobs"].append(logp.numpy())
        nobs, rew, term, trunc, _ = env.step(a_np)
        bufs["rewards"].append(np.asarray(rew))
        state["next_obs"], state["next_starts"] = nobs, np.logical_or(term, trunc)
        state["next_masks"] = env.get_action_mask()
    stack = lambda xs: ({k: np.stack([x[k] for x in xs]) for k in xs[0]} if isinstance(xs[0], dict)
                        else (None if xs[0] is None else np.stack(xs)))
    out = {k: stack(v) for k, v in bufs.items()}
    out["next_values"] = policy.value(_t(np.asarray(state["next_obs"]))).numpy()
    out["next_episode_starts"] = np.asarray(state["next_starts"])
    return out


def _flat(a):
    if a is None:
        return None
    if isinstance(a, dict):
        return {k: _flat(v) for k, v in a.items()}
    return torch.as_tensor(flatten_time_major(np.asarray(a)))


def _take(a, idx):
    if a is None:
        return None
    if isinstance(a, dict):
        return {k: v[idx] for k, v in a.items()}
    return a[idx]


def learn_epoch(policy: OraclePolicy, optimizer: torch.optim.Optimizer, ro: dict, hp: Hyper,
                teacher: Optional[OraclePolicy] = None) -> dict:
    """vec_rollout.py:38-175 + ppo.py:258-447 on a collected rollout ``ro`` (numpy [T, N, ...] arrays)."""
    adv = gae_advantages(ro["rewards"], ro["values"], ro["episode_starts"], ro["next_episode_starts"],
                         ro["next_values"], hp.gamma, hp.gae_lambda)
    ret = gae_returns(adv, ro["values"])
    b = dict(obs=_flat(ro["obs"]), logprobs=_flat(ro["logprobs"]), actions=_flat(ro["actions"]),
             masks=_flat(ro["masks"]), values=_flat(ro["values"]), adv=_flat(adv), returns=_flat(ret))
    total = b["logprobs"].shape[0]
    teacher_logp = None
    if teacher is not None:  # TeacherKLLoss.add_to_batch over num_envs-sized slices (ppo.py:259-262, vec_rollout.py:155-164)
        n_envs = np.asarray(ro["rewards"]).shape[1]
        with torch.no_grad():
            teacher_logp = torch.cat([
                teacher.forward(b["obs"][i:i + n_envs], _take(b["actions"], slice(i, i + n_envs)),
                                _take(b["masks"], slice(i, i + n_envs)))[0]
                for i in range(0, total, n_envs)])
    n_mb = total // hp.batch_size + (1 if total % hp.batch_size else 0)
    dt = b["adv"].dtype  # float32 (the reference's torch.Tensor(...) constants); float64 in the generators' f64 runs
    w = torch.tensor(np.asarray(hp.multi_reward_weights), dtype=dt) if hp.multi_reward_weights is not None else None
    vf_coef = torch.tensor(np.asarray(hp.vf_coef), dtype=dt)
    vf_weights = torch.tensor(np.asarray(hp.vf_weights), dtype=dt) if hp.vf_weights is not None else None
    params = list(policy.parameters())
    pi_coef = 1
    step_stats: List[dict] = []
    grad_norms: List[float] = []

    def optimizer_step() -> float:
        gn = nn.utils.clip_grad_norm_(params, hp.max_grad_norm).item()
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
        return gn

    for _ in range(hp.n_epochs):
        step_stats.clear(), grad_norms.clear()
        for idx in minibatch_index_stream(total, hp.batch_size, shuffle=not hp.gradient_accumulation):
            mb_adv = normalize_advantages(
                b["adv"][idx], normalize_advantage=hp.normalize_advantage,
                standardize_advantage=hp.standardize_advantage,
                normalize_advantages_after_scaling=hp.normalize_advantages_after_scaling, multi_reward_weights=w)
            logp, ent, v = policy.forward(b["obs"][idx], _take(b["actions"], idx), _take(b["masks"], idx))
            parts = ppo_loss(logp, ent, v, b["logprobs"][idx], mb_adv, b["values"][idx], b["returns"][idx],
                             clip_range=hp.clip_range, clip_range_vf=hp.clip_range_vf, ent_coef=hp.ent_coef,
                             vf_coef=vf_coef, ppo2_vf_coef_halving=hp.ppo2_vf_coef_halving, vf_weights=vf_weights,
                             pi_coef=pi_coef, kl_cutoff=hp.kl_cutoff, loss_divisor=n_mb if hp.gradient_accumulation else None,
                             teacher_logprobs=teacher_logp[idx] if teacher_logp is not None else None,
                             teacher_kl_loss_coef=hp.teacher_kl_loss_coef, teacher_unbiased=hp.teacher_unbiased,
                             teacher_loss_importance_sampling=hp.teacher_loss_importance_sampling,
                             vf_loss_fn=getattr(torch.nn.functional, hp.vf_loss_fn))
            pi_coef = parts.pi_coef
            parts.loss.backward()
            if not hp.gradient_accumulation:
                grad_norms.append(optimizer_step())
            step_stats.append(dict(loss=parts.loss.item(), pi_loss=parts.pi_loss.item(),
                                   v_loss=parts.v_loss.detach().numpy().copy(), entropy_loss=parts.entropy_loss.item(),
                                   approx_kl=parts.approx_kl, clipped_frac=parts.clipped_frac,
                                   val_clipped_frac=np.asarray(parts.val_clipped_frac),
                                   **({"teacher_kl_loss": parts.teacher_kl_loss.item()} if parts.teacher_kl_loss is not None else {})))
        if hp.gradient_accumulation:
            grad_norms.append(optimizer_step())
    y_true, y_pred = flatten_time_major(ret), flatten_time_major(ro["values"])
    var_y = np.var(y_true).item()
    stats = {k: np.mean([s[k] for s in step_stats], axis=0) for k in step_stats[0]}
    stats["explained_var"] = np.nan if var_y == 0 else 1 - np.var(y_true - y_pred).item() / var_y
    stats["grad_norm"] = float(np.mean(grad_norms))
    stats["total_steps"] = total
    return stats


@dataclass
class A2CHyper:
    """The reference's A2C keyword arguments (a2c/a2c.py:24-43)."""

    learning_rate: float = 7e-4
    gamma: Union[float, np.ndarray] = 0.99
    gae_lambda: Union[float, np.ndarray] = 1.0
    ent_coef: float = 0.0
    vf_coef: Union[float, Sequence[float]] = 0.5
    max_grad_norm: float = 0.5
    rms_prop_eps: float = 1e-5
    use_rms_prop: bool = True
    normalize_advantage: bool = False
    multi_reward_weights: Optional[Sequence[float]] = None
    gradient_accumulation: bool = False
    num_minibatches: int = 1


def a2c_learn_iteration(policy: OraclePolicy, optimizer: torch.optim.Optimizer, ro: dict, hp: A2CHyper) -> dict:
    """One iteration of A2C.learn (a2c/a2c.py:104-173) on a collected rollout."""
    adv = gae_advantages(ro["rewards"], ro["values"], ro["episode_starts"], ro["next_episode_starts"],
                         ro["next_values"], hp.gamma, hp.gae_lambda)
    ret = gae_returns(adv, ro["values"])
    b = dict(obs=_flat(ro["obs"]), actions=_flat(ro["actions"]), masks=_flat(ro["masks"]), adv=_flat(adv),
             returns=_flat(ret))
    total = b["adv"].shape[0]
    w = torch.tensor(np.asarray(hp.multi_reward_weights), dtype=torch.float32) if hp.multi_reward_weights is not None else None
    vf_coef = torch.tensor(np.asarray(hp.vf_coef), dtype=torch.float32)
    params = list(policy.parameters())
    step_stats = []

    def optimizer_step():
        nn.utils.clip_grad_norm_(params, hp.max_grad_norm)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)

    for idx in minibatch_index_stream(total, total // hp.num_minibatches, shuffle=not hp.gradient_accumulation):
        mb_adv = b["adv"][idx]
        if hp.normalize_advantage:
            mb_adv = (mb_adv - mb_adv.mean(0)) / (mb_adv.std(0) + 1e-8)
        if w is not None:
            mb_adv = mb_adv @ w
        logp, ent, v = policy.forward(b["obs"][idx], _take(b["actions"], idx), _take(b["masks"], idx))
        pi_loss = -(mb_adv * logp).mean()
        value_loss = ((v - b["returns"][idx]) ** 2).mean(0)
        entropy_loss = -ent.mean()
        loss = pi_loss + (vf_coef * value_loss).sum() + hp.ent_coef * entropy_loss
        if hp.gradient_accumulation:
            loss = loss / hp.num_minibatches
        loss.backward()
        if not hp.gradient_accumulation:
            optimizer_step()
        step_stats.append(dict(loss=loss.item(), pi_loss=pi_loss.item(), v_loss=value_loss.detach().numpy().copy(),
                               entropy_loss=entropy_loss.item()))
    if hp.gradient_accumulation:
        optimizer_step()
    stats = {k: np.mean([s[k] for s in step_stats], axis=0) for k in step_stats[0]}
    y_true, y_pred = flatten_time_major(ret), flatten_time_major(ro["values"])
    var_y = np.var(y_true).item()
    stats["explained_var"] = np.nan if var_y == 0 else 1 - np.var(y_true - y_pred).item() / var_y
    return stats


def acbc_learn_iteration(policy: OraclePolicy, optimizer: torch.optim.Optimizer, ro: dict, batch_size: int, n_epochs: int,
                         gamma, gae_lambda, vf_coef, max_grad_norm: float = 0.5, gradient_accumulation: bool = False) -> dict:
    """One iteration of ACBC.learn (acbc/acbc.py:75-141) on a collected rollout."""
    adv = gae_advantages(ro["rewards"], ro["values"], ro["episode_starts"], ro["next_episode_starts"],
                         ro["next_values"], gamma, gae_lambda)
    ret = gae_returns(adv, ro["values"])
    b = dict(obs=_flat(ro["obs"]), actions=_flat(ro["actions"]), masks=_flat(ro["masks"]), returns=_flat(ret))
    total = b["returns"].shape[0]
    n_mb = total // batch_size + (1 if total % batch_size else 0)
    vf = torch.tensor(np.asarray(vf_coef), dtype=torch.float32)
    params = list(policy.parameters())
    stats: List[dict] = []

    def optimizer_step():
        nn.utils.clip_grad_norm_(params, max_grad_norm)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)

    for _ in range(n_epochs):
        stats.clear()
        for idx in minibatch_index_stream(total, batch_size, shuffle=not gradient_accumulation):
            logp, _, v = policy.forward(b["obs"][idx], _take(b["actions"], idx), _take(b["masks"], idx))
            pi_loss = -logp.mean()
            v_loss = ((v - b["returns"][idx]) ** 2).mean(0)
            loss = pi_loss + (vf * v_loss).sum()
            if gradient_accumulation:
                loss = loss / n_mb
            loss.backward()
            if not gradient_accumulation:
                optimizer_step()
            stats.append(dict(loss=loss.item(), pi_loss=pi_loss.item(), v_loss=v_loss.detach().numpy().copy()))
        if gradient_accumulation:
            optimizer_step()
    return {k: np.mean([s[k] for s in stats], axis=0) for k in stats[0]}
