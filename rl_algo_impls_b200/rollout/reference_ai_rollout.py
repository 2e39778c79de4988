"""ReferenceAIRolloutGenerator: behaviour-cloning rollouts where a scripted AI inside the env chooses the actions.

Mirrors ``rl_algo_impls/rollout/reference_ai_rollout.py:20-102``: the policy only supplies values (and, with
``include_logp``, log-probs); the actions stored in the rollout are the ones the env reports back in
``vec_env.last_action``.  Buffers are the HBM-resident ``[T, N, ...]`` tensors of ``SyncStepRolloutGenerator``;
each step's observation / start flags / masks / values go to row ``s`` with one K0 launch, the env's
``last_action`` and rewards are uploaded into row ``s`` afterwards, and GAE + returns are the K1 launch inside
``VecRollout``.

Two deliberate differences from the reference revision, both of which are defects there: ``vec_env.step`` is
unpacked as the 5-tuple every other generator uses (reference_ai_rollout.py:62-67 still expects the
pre-gymnasium 4-tuple), and with ``include_logp`` the step's log-probs go to row ``s`` (the reference rebinds
the whole ``self.logprobs`` array to one step's vector, :49-56).
"""
import numpy as np
import torch

from .. import ops
from .sync_step_rollout import SyncStepRolloutGenerator
from .trajectory import _map
from .vec_rollout import VecRollout


class ReferenceAIRolloutGenerator(SyncStepRolloutGenerator):
    def __init__(self, policy, vec_env, **kwargs) -> None:
        kwargs.setdefault("cuda_graph", False)  # the env (a host AI) sits between the policy and the buffer write
        super().__init__(policy, vec_env, **kwargs)
        if self._pack_in_step:  # this loop writes next_obs itself: no deferred packing
            self._pack_raw()
            self._pack_in_step = False
        self._row = torch.zeros(1, dtype=torch.int64, device=self.device)
        if not self.include_logp:
            self.zero_action = _map(lambda buf: torch.zeros(tuple(buf.shape[1:]), dtype=buf.dtype, device=self.device),
                                    self.actions)

    def _store_actions(self, s: int, last_action) -> None:
        """vec_env.last_action (numpy, tensors, a dict of either, or an object array of per-env dicts) -> row s."""
        if isinstance(last_action, np.ndarray) and last_action.dtype == object:  # tensor_utils.batch_dict_keys
            last_action = {k: np.array([a[k] for a in last_action]) for k in last_action[0]}
        if isinstance(self.actions, dict):
            for k, buf in self.actions.items():
                self._upload_cast("last_action_" + k, last_action[k], buf[s])
        else:
            self._upload_cast("last_action", last_action, self.actions[s])

    def _upload_cast(self, name: str, src, dst: torch.Tensor) -> None:
        if isinstance(src, torch.Tensor):
            dst.copy_(src.reshape(dst.shape))
        else:
            self._upload(name, np.asarray(src).reshape(tuple(dst.shape)), dst)

    def rollout(self, gamma, gae_lambda) -> VecRollout:
        self.policy.eval()
        self.policy.reset_noise()
        for s in range(self.n_steps):
            if self.sde_sample_freq > 0 and s > 0 and s % self.sde_sample_freq == 0:
                self.policy.reset_noise()
            self._row.fill_(s)
            with torch.no_grad():
                if self.include_logp:
                    step_actions, values, logp = self.policy.step_device(self.next_obs, self.next_action_masks)
                else:
                    values, logp, step_actions = self.policy.value_device(self.next_obs), None, self.zero_action
            # this step's pre-env fields -> row s of the buffers (one launch), before next_obs / masks move on
            src, dst = [self.next_obs, self.next_episode_starts], [self.obs, self.episode_starts]
            if self.action_masks is not None:
                if isinstance(self.action_masks, dict):
                    for k, buf in self.action_masks.items():
                        src.append(self.next_action_masks[k]), dst.append(buf)
                else:
                    src.append(self.next_action_masks), dst.append(self.action_masks)
            src.append(values.float().reshape(self.values.shape[1:]).contiguous()), dst.append(self.values)
            if logp is not None and self.logprobs is not None:
                src.append(logp.float().contiguous()), dst.append(self.logprobs)
            ops.rollout_store_step(src, dst, self._row)

            next_obs, rewards, terminations, truncations, _ = self.vec_env.step(self._env_actions(step_actions))
            self._store_actions(s, getattr(self.vec_env, "last_action"))
            self._upload("obs", next_obs, self.next_obs)
            self._upload_cast("rewards", rewards, self.rewards[s])
            if isinstance(terminations, torch.Tensor):
                torch.logical_or(terminations, truncations, out=self.next_episode_starts)
            else:
                self._upload("starts", np.logical_or(terminations, truncations), self.next_episode_starts)
            if self.get_action_mask is not None and self.next_action_masks is not None:
                self._upload_masks(self.get_action_mask())

        with torch.no_grad():
            next_values = self.policy.value_device(self.next_obs)
        self.policy.train()
        return VecRollout(
            device=self.device,
            next_episode_starts=self.next_episode_starts,
            next_values=next_values,
            obs=self.obs,
            actions=self.actions,
            rewards=self.rewards,
            episode_starts=self.episode_starts,
            values=self.values,
            logprobs=self.logprobs,
            action_masks=self.action_masks,
            gamma=gamma,
            gae_lambda=gae_lambda,
            scale_advantage_by_values_accuracy=self.scale_advantage_by_values_accuracy,
            full_batch_off_accelerator=self.full_batch_off_accelerator,
            subaction_mask=self.subaction_mask,
            action_plane_space=getattr(self.vec_env, "action_plane_space", None),
            out_advantages=self.advantages,
            out_returns=self.returns,
            include_num_actions=True,
        )
