"""ActorCritic: the policy object the PPO learner and the rollout generator talk to.

Mirrors the surface of ``rl_algo_impls/shared/policy/actor_critic.py:109-395`` that the hot path
uses -- ``forward(obs, action, action_masks) -> ACForward(logp_a, entropy, v)`` (:286-296),
``step(obs, action_masks) -> Step(a, v, logp_a, clamped_a)`` on numpy (:306-318), ``value(obs)``
(:298-304), ``action_shape`` / ``value_shape``, ``reset_noise``, ``freeze`` / ``unfreeze`` -- and adds
the device-resident entry points the B200 path is built on:

* ``head_outputs(obs)``: raw logits / Gaussian mean + values of one trunk pass, which the fused
  loss kernels differentiate directly (no per-head distribution objects);
* ``step_device(obs, action_masks)``: sample + log-prob in one launch, everything stays in HBM.
"""
from typing import Dict, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import ops, spaces
from ..actor.categorical import MaskedCategorical
from ..actor.gridnet import GridnetDistribution, ValueDependentMask
from ..actor.rng import current_seed, next_sample_stream
from .networks import (GridEncoderDecoderActorCritic, HeadOutputs, MlpActorCritic, NatureCnnActorCritic,
                       SqueezeUnetActorCritic)

TensorOrDict = Union[torch.Tensor, Dict[str, torch.Tensor]]
NumpyOrDict = Union[np.ndarray, Dict[str, np.ndarray]]


class ACForward(NamedTuple):
    logp_a: torch.Tensor
    entropy: torch.Tensor
    v: torch.Tensor


class Step(NamedTuple):
    a: NumpyOrDict
    v: np.ndarray
    logp_a: np.ndarray
    clamped_a: NumpyOrDict


def clamp_actions(actions: NumpyOrDict, action_space, squash_output: bool) -> NumpyOrDict:
    """actor_critic.py:62-76 (the reference's one unit-tested function): clip Box actions to the
    bounds, or rescale squashed [-1, 1] actions into them."""
    if spaces.is_box(action_space):
        low, high = action_space.low, action_space.high
        if squash_output:
            return low + 0.5 * (actions + 1) * (high - low)
        return np.clip(actions, low, high)
    return actions


class ActorCritic(nn.Module):
    supports_wide_actions = True  # step_device(wide_out=...): GridNet per-cell actions also as int64 (host envs)

    def __init__(self, env, network: Optional[nn.Module] = None, subaction_mask=None, squash_output: bool = False,
                 **hyperparams) -> None:
        super().__init__()
        self.env = env
        self.action_space = env.single_action_space
        self.observation_space = env.single_observation_space
        self.action_plane_space = getattr(env, "action_plane_space", None)
        self.squash_output = squash_output
        self.subaction_mask = subaction_mask
        self.n_pick = 0
        if self.action_plane_space is not None:
            self.kind = "gridnet"
            self.nvec = tuple(int(n) for n in self.action_plane_space.nvec)
            per_pos = self.action_space["per_position"] if spaces.is_dict(self.action_space) else self.action_space
            self.map_size = len(per_pos.nvec) // len(self.nvec)
            if spaces.is_dict(self.action_space) and "pick_position" in self.action_space.keys():
                self.n_pick = len(self.action_space["pick_position"].nvec)
            self.spec = ops.GridnetSpec.from_subaction_mask(self.nvec, subaction_mask, self.n_pick)
        elif spaces.is_discrete(self.action_space):
            self.kind = "categorical"
        elif spaces.is_box(self.action_space):
            self.kind = "gaussian"
        else:
            raise NotImplementedError(f"unsupported action space {self.action_space!r}")
        self.network = network if network is not None else default_network(env, self, **hyperparams)
        self._n_values: Optional[int] = getattr(self.network, "n_values", 1)

    # -- shape contract (actor_critic.py:372-378) ------------------------------------------------
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def action_shape(self):
        if self.kind == "gridnet":
            cells = (self.map_size, len(self.nvec))
            return {"per_position": cells, "pick_position": (self.n_pick,)} if self.n_pick else cells
        if self.kind == "categorical":
            return ()
        return tuple(self.action_space.shape)

    @property
    def value_shape(self) -> Tuple[int, ...]:
        return () if self._n_values == 1 else (self._n_values,)

    def reset_noise(self, batch_size: Optional[int] = None) -> None:  # gSDE only in the reference
        pass

    def freeze(self, freeze_policy_head: bool, freeze_value_head: bool, freeze_backbone: bool = True) -> None:
        """actor_critic.py:384-395 over the networks' freeze (backbone_actor_critic.py:254-265, unet.py:221-245):
        requires_grad of the policy head, the value head(s) and everything else (the backbone).  A trunk names its
        heads in ``policy_head_modules`` / ``value_head_modules`` (attribute names); or implements ``freeze`` itself."""
        net = self.network
        if hasattr(net, "freeze"):
            net.freeze(freeze_policy_head, freeze_value_head, freeze_backbone=freeze_backbone)
            return
        if not hasattr(net, "policy_head_modules") or not hasattr(net, "value_head_modules"):
            raise NotImplementedError(f"{type(net).__name__} does not say which of its modules are the policy / value heads")

        def params_of(names):
            out = []
            for n in names:
                try:
                    out += list(net.get_submodule(n).parameters())
                except AttributeError:
                    out.append(net.get_parameter(n))
            return out

        policy, value = params_of(net.policy_head_modules), params_of(net.value_head_modules)
        head_ids = {id(p) for p in policy + value}
        for p in policy:
            p.requires_grad = not freeze_policy_head
        for p in value:
            p.requires_grad = not freeze_value_head
        for p in net.parameters():
            if id(p) not in head_ids:
                p.requires_grad = not freeze_backbone

    def unfreeze(self) -> None:
        self.freeze(False, False, freeze_backbone=False)

    # -- device-resident path ---------------------------------------------------------------------
    def head_outputs(self, obs: torch.Tensor) -> HeadOutputs:
        return self.network(obs)

    @property
    def packed_obs_shape(self):
        """Shape of one observation in the layout the trunk consumes without a copy ((H, W, Cp): channels last, planes
        padded to a multiple of 8), or None when the trunk takes the env's layout as is.  A rollout generator that
        stores observations packed hands the learner trunk-ready minibatch rows (networks._PaddedEnds)."""
        fn = getattr(self.network, "packed_obs_shape", None)
        return fn() if fn is not None else None

    def pack_observations(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.network.pack_observations(obs, out)

    def _grid_logits(self, out: HeadOutputs) -> torch.Tensor:
        lg = out.pi
        return lg.reshape(lg.shape[0], self.map_size, lg.shape[-1])

    def forward(self, obs: torch.Tensor, action: TensorOrDict, action_masks: Optional[TensorOrDict] = None) -> ACForward:
        """Distribution-level path: differentiable (logp_a, entropy, v) through the fused fwd/bwd
        kernels.  The PPO learner uses the fully fused loss kernels instead (ppo/ppo.py)."""
        out = self.head_outputs(obs)
        if self.kind == "gridnet":
            assert action_masks is not None, "GridNet heads need action masks"
            pi = GridnetDistribution(self.map_size, np.asarray(self.nvec), self._grid_logits(out), action_masks,
                                     subaction_mask=_gates(self.subaction_mask))
            return ACForward(pi.log_prob(action), pi.entropy(), out.values)
        if self.kind == "categorical":
            pi = MaskedCategorical(logits=out.pi, mask=action_masks)
            return ACForward(pi.log_prob(action), pi.entropy(), out.values)
        logp, ent = _GaussianFn.apply(out.pi, out.log_std, action.float())
        return ACForward(logp, ent, out.values)

    @torch.no_grad()
    def step_device(self, obs: torch.Tensor, action_masks: Optional[TensorOrDict] = None,
                    offset_dev: Optional[torch.Tensor] = None, wide_out: Optional[torch.Tensor] = None):
        """(actions, values, logp) on the device; per-cell actions are uint8, pick / discrete int64.
        ``offset_dev`` (int64 device scalar) is added to the RNG offset on the device, so a launch
        captured in a CUDA graph draws fresh numbers on every replay."""
        out = self.head_outputs(obs)
        if offset_dev is None:
            seed, offset = next_sample_stream()
        else:  # the device counter alone advances the stream: eager calls and graph replays draw alike
            seed, offset = current_seed(), 0
        if self.kind == "gridnet":
            cells_mask = action_masks["per_position"] if isinstance(action_masks, dict) else action_masks
            pick_mask = action_masks.get("pick_position") if isinstance(action_masks, dict) else None
            logits = self._grid_logits(out).contiguous()
            # wide_out ([N, HW, A] int64, GridNet only): the per-cell actions once more in the dtype a host env takes
            cells, pick, logp = ops.gridnet_sample(self.spec, logits, cells_mask, pick_mask, seed, offset, torch.uint8,
                                                   offset_dev, wide_out)
            a: TensorOrDict = {"per_position": cells, "pick_position": pick} if self.n_pick else cells
            return a, out.values, logp
        if self.kind == "categorical":
            a, logp = ops.categorical_sample(out.pi.float().contiguous(), action_masks, seed, offset, offset_dev)
            return a, out.values, logp
        std = torch.exp(out.log_std)
        a = out.pi + std * torch.randn_like(out.pi)
        logp, _ = ops.gaussian_logp_entropy(out.pi.contiguous(), out.log_std.contiguous(), a.contiguous())
        return a, out.values, logp

    @torch.no_grad()
    def value_device(self, obs: torch.Tensor) -> torch.Tensor:
        return self.head_outputs(obs).values

    # -- the reference's numpy-facing contract ----------------------------------------------------
    def _as_tensor(self, a):
        if a is None:
            return None
        if isinstance(a, dict):
            return {k: self._as_tensor(v) for k, v in a.items()}
        if isinstance(a, torch.Tensor):
            return a.to(self.device)
        return torch.as_tensor(a).to(self.device)

    def step(self, obs: np.ndarray, action_masks: Optional[NumpyOrDict] = None) -> Step:
        a, v, logp = self.step_device(self._as_tensor(obs), self._as_tensor(action_masks))
        if isinstance(a, dict):
            a_np: NumpyOrDict = {k: t.cpu().numpy().astype(np.int64) for k, t in a.items()}
        else:
            a_np = a.cpu().numpy()
            if a_np.dtype == np.uint8:
                a_np = a_np.astype(np.int64)
        return Step(a_np, v.cpu().numpy(), logp.cpu().numpy(), clamp_actions(a_np, self.action_space, self.squash_output))

    def value(self, obs: np.ndarray) -> np.ndarray:
        return self.value_device(self._as_tensor(obs)).cpu().numpy()

    def act(self, obs: np.ndarray, deterministic: bool = True, action_masks: Optional[NumpyOrDict] = None):
        if not deterministic:
            return self.step(obs, action_masks=action_masks).clamped_a
        with torch.no_grad():
            out = self.head_outputs(self._as_tensor(obs))
            masks = self._as_tensor(action_masks)
            if self.kind == "gridnet":
                mode = GridnetDistribution(self.map_size, np.asarray(self.nvec), self._grid_logits(out), masks,
                                           subaction_mask=_gates(self.subaction_mask)).mode
                return {k: v.cpu().numpy() for k, v in mode.items()} if isinstance(mode, dict) else mode.cpu().numpy()
            if self.kind == "categorical":
                return MaskedCategorical(logits=out.pi, mask=masks).mode.cpu().numpy()
            return clamp_actions(out.pi.cpu().numpy(), self.action_space, self.squash_output)


def _gates(subaction_mask):
    if not subaction_mask:
        return None
    return ValueDependentMask.from_reference_index_to_index_to_value(subaction_mask)


class _GaussianFn(torch.autograd.Function):
    """Differentiable Gaussian (logp [B], entropy [B, D]) for the distribution-level path."""

    @staticmethod
    def forward(ctx, mu, log_std, actions):
        mu, log_std, actions = mu.contiguous(), log_std.contiguous(), actions.contiguous()
        logp, ent = ops.gaussian_logp_entropy(mu, log_std, actions)
        ctx.save_for_backward(mu, log_std, actions)
        return logp, ent

    @staticmethod
    def backward(ctx, dlogp, dent):
        mu, log_std, actions = ctx.saved_tensors
        var = torch.exp(2 * log_std)
        diff = actions - mu
        dmu = dlogp.unsqueeze(-1) * diff / var
        dls = (dlogp.unsqueeze(-1) * (diff * diff / var - 1) + dent).sum(0)
        return dmu, dls, None


def default_network(env, policy: ActorCritic, pi_hidden_sizes=None, v_hidden_sizes=None, activation_fn: str = "tanh",
                    log_std_init: float = -0.5, cnn_flatten_dim: int = 512, actor_head_style: str = "single",
                    channels_per_level=None, strides_per_level=None, deconv_strides_per_level=None,
                    encoder_residual_blocks_per_level=None, decoder_residual_blocks_per_level=None,
                    critic_channels: int = 64, num_additional_critics: int = 0,
                    additional_critic_activation_functions=None, output_activation_fn: str = "identity",
                    shared_critic_head: bool = False, increment_kernel_size_on_down_conv: bool = False,
                    normalization=None, critic_shares_backbone: bool = True, **_ignored) -> nn.Module:
    """Trunk chosen from the spaces and the reference's policy hyperparameter names
    (runner/running_utils.py:187-210 -> ActorCritic(env, **policy_hyperparams)).  Trunks are plain PyTorch modules
    and out of this repo's scope: the families of the five BASELINE configs are built here; for any other
    (``unet``, ``double_cone``, ``sacus``, normalised variants) pass the module itself as ``network=``."""
    obs_shape = tuple(env.single_observation_space.shape)
    if policy.kind == "gridnet":
        n_logits = sum(policy.nvec) + policy.n_pick
        side = int(round(np.sqrt(policy.map_size)))
        if actor_head_style == "squeeze_unet" or (policy.n_pick and actor_head_style == "single"):
            if normalization is not None or not critic_shares_backbone:
                raise NotImplementedError("squeeze_unet with normalization / a separate critic backbone: pass network=")
            extra = list(additional_critic_activation_functions or ["identity"] * int(num_additional_critics))
            space = env.single_observation_space
            obs_range = float(np.max(space.high) - np.min(space.low)) if spaces.is_box(space) else 1.0
            return SqueezeUnetActorCritic(
                obs_shape[0], n_logits, channels_per_level=tuple(channels_per_level or (64, 128, 256)),
                strides_per_level=strides_per_level, deconv_strides_per_level=deconv_strides_per_level,
                encoder_residual_blocks_per_level=encoder_residual_blocks_per_level,
                decoder_residual_blocks_per_level=decoder_residual_blocks_per_level, critic_channels=critic_channels,
                critic_activations=[output_activation_fn] + extra, shared_critic_head=shared_critic_head,
                increment_kernel_size_on_down_conv=increment_kernel_size_on_down_conv,
                obs_range=obs_range if np.isfinite(obs_range) and obs_range > 0 else 1.0, obs_hw=tuple(obs_shape[1:]))
        if actor_head_style in ("unet", "double_cone", "sacus"):
            raise NotImplementedError(f"actor_head_style={actor_head_style!r}: pass the trunk module as network=")
        return GridEncoderDecoderActorCritic(obs_shape[0], (side, side), n_logits, tuple(v_hidden_sizes or (128,)),
                                             1 + int(num_additional_critics))
    if len(obs_shape) == 3:
        return NatureCnnActorCritic(obs_shape[0], policy.action_space.n, obs_shape[1:], cnn_flatten_dim)
    obs_dim = int(np.prod(obs_shape))
    if policy.kind == "categorical":
        return MlpActorCritic(obs_dim, policy.action_space.n, tuple(pi_hidden_sizes or (64, 64)),
                              tuple(v_hidden_sizes or (64, 64)), activation_fn)
    return MlpActorCritic(obs_dim, int(np.prod(policy.action_space.shape)), tuple(pi_hidden_sizes or (64, 64)),
                          tuple(v_hidden_sizes or (64, 64)), activation_fn, gaussian=True, log_std_init=log_std_init)
