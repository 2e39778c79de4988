"""Guided rollouts: a guide policy plays part of every episode, only the learner's steps are kept.

Mirrors ``rl_algo_impls/rollout/guided_learner_rollout.py:16-208`` (``GuidedLearnerRolloutGenerator``: the guide
plays the first ``switch_step ~ U[0, switch_range)`` steps of an episode) and
``rollout/random_guided_learner_rollout.py:21-256`` (``RandomGuidedLearnerRolloutGenerator``: every step is the
guide's with probability ``guide_probability``; skipped steps fold their discounted reward into the learner's
last kept step) -- same constructor keywords, same ``np.random`` calls in the same order (so the same seed
yields the same guide/learner schedule), same trajectory order in the resulting ``TrajectoryRollout``.

What changes is where the data lives.  The per-env control flow (who acts, which builder a step belongs to,
reward folding) needs the env's rewards / dones on the host and stays there; everything bulky -- observations,
masks, sampled actions, values, log-probs -- is produced on the device by ``policy.step_device`` and written
once per vec-env step into a ``StepStore``; builders record row numbers only.  A host env's observations and
masks are uploaded through pinned staging buffers, a device env's tensors are used in place.
"""
import logging
from typing import Dict, List, Optional

import numpy as np
import torch

from .rollout import RolloutGenerator
from .sync_step_rollout import _Uploader, _torch_dtype
from .trajectory import DiscreteSkipsTrajectoryBuilder, StepStore, TrajectoryBuilder, _map, step_fields
from .trajectory_rollout import TrajectoryRollout


def rearrange(lst: List, indices: List[int]) -> List:
    """guided_learner_rollout.py:190-191"""
    return [itm for _, itm in sorted(zip(indices, lst), key=lambda p: p[0])]


def has_actions(action_mask) -> np.ndarray:
    """random_guided_learner_rollout.py:253-256 for numpy masks, CUDA tensors or dicts of either: [N] bool on the host."""
    if isinstance(action_mask, dict):
        parts = [has_actions(m) for m in action_mask.values()]
        return np.logical_or.reduce(parts)
    if isinstance(action_mask, torch.Tensor):
        return action_mask.flatten(1).any(dim=1).cpu().numpy()
    return np.asarray(action_mask).reshape(len(action_mask), -1).any(axis=1)


class _DeviceStepper(RolloutGenerator):
    """What both guided generators share: device copies of the env's current observation / masks, subset policy
    steps scattered into full ``[N, ...]`` step tensors, and the host <-> device traffic of one env step."""

    def __init__(self, learning_policy, vec_env, guide_policy, n_steps: int, sde_sample_freq: int,
                 scale_advantage_by_values_accuracy: bool, full_batch_off_accelerator: bool, include_logp: bool,
                 subaction_mask: Optional[Dict[int, Dict[int, int]]]) -> None:
        super().__init__(learning_policy, vec_env)
        self.learning_policy, self.guide_policy = learning_policy, guide_policy
        self.n_steps, self.sde_sample_freq = int(n_steps), sde_sample_freq
        self.scale_advantage_by_values_accuracy = scale_advantage_by_values_accuracy
        if not full_batch_off_accelerator:
            logging.warning(f"{self.__class__.__name__}: full_batch_off_accelerator is ignored (the batch lives in HBM)")
        if not include_logp:
            logging.warning(f"{self.__class__.__name__} doesn't implement include_logp=False")
        self.subaction_mask = subaction_mask
        self.device = torch.device(learning_policy.device)
        if self.device.type != "cuda":
            raise RuntimeError(f"{self.__class__.__name__} keeps rollout steps in HBM: the policy must be on a CUDA device")
        self.get_action_mask = getattr(vec_env, "get_action_mask", None)
        self._upload = _Uploader(self.device)
        self._device_env = getattr(vec_env, "device", None) is not None
        self.d2h_bytes = 0
        self.store = StepStore(self.device, vec_env.num_envs, capacity=max(2 * self.n_steps, 16))
        self._full: Dict[str, torch.Tensor] = {}
        first_obs, _ = vec_env.reset()
        self.next_obs = self._obs_to_device(first_obs)
        self.next_action_masks_host = self.get_action_mask() if self.get_action_mask else None
        self.next_action_masks = self._masks_to_device(self.next_action_masks_host)

    @property
    def num_envs(self) -> int:
        return self.vec_env.num_envs

    # -- host <-> device --------------------------------------------------------------------------------
    def _obs_to_device(self, obs) -> torch.Tensor:
        if isinstance(obs, torch.Tensor):
            return obs
        if "obs" not in self._full:
            self._full["obs"] = torch.zeros(obs.shape, dtype=_torch_dtype(obs.dtype), device=self.device)
        self._upload("obs", obs, self._full["obs"])
        return self._full["obs"]

    def _masks_to_device(self, masks):
        if masks is None:
            return None
        if isinstance(masks, dict):
            return {k: self._mask_to_device("mask_" + k, v) for k, v in masks.items()}
        return self._mask_to_device("mask", masks)

    def _mask_to_device(self, name: str, m) -> torch.Tensor:
        if isinstance(m, torch.Tensor):
            return m
        if name not in self._full:
            self._full[name] = torch.zeros(m.shape, dtype=torch.bool, device=self.device)
        self._upload(name, m, self._full[name])
        return self._full[name]

    def _host(self, x) -> np.ndarray:
        return x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)

    def _env_actions(self, a):
        """What vec_env.step receives: CUDA tensors for a device env, int64 numpy for a host env."""
        if self._device_env:
            return a
        self.d2h_bytes += sum(t.numel() * t.element_size() for t in (a.values() if isinstance(a, dict) else [a]))
        to_np = lambda t: t.cpu().numpy().astype(np.int64) if t.dtype == torch.uint8 else t.cpu().numpy()
        return _map(to_np, a)

    # -- subset policy steps ----------------------------------------------------------------------------
    def _scatter(self, name: str, part: torch.Tensor, rows: Optional[torch.Tensor]) -> torch.Tensor:
        """Rows of this step's full ``[N, ...]`` tensor `name` <- the subset result `part`."""
        if rows is None:
            return part
        key = "step_" + name
        full = self._full.get(key)
        if full is None or full.shape[1:] != part.shape[1:] or full.dtype != part.dtype:
            full = self._full[key] = torch.zeros((self.num_envs,) + tuple(part.shape[1:]), dtype=part.dtype, device=self.device)
        full.index_copy_(0, rows, part)
        return full

    def _step_policy(self, policy, select: np.ndarray, obs, masks, out: dict) -> None:
        """policy.step_device on the envs `select` ([N] bool, host); results land in out['a' / 'v' / 'logp']."""
        rows = None if select.all() else torch.from_numpy(np.nonzero(select)[0]).to(self.device)
        sub = (lambda t: t) if rows is None else (lambda t: t.index_select(0, rows))
        a, v, logp = policy.step_device(sub(obs), _map(sub, masks))
        if isinstance(a, dict):
            out["a"] = {k: self._scatter("a_" + k, t, rows) for k, t in a.items()}
        else:
            out["a"] = self._scatter("a", a, rows)
        out["v"], out["logp"] = self._scatter("v", v.float(), rows), self._scatter("logp", logp.float(), rows)

    def _advance(self, actions):
        """vec_env.step + the uploads of what the next policy step reads; (rewards, dones) come back on the host."""
        next_obs, rewards, terminations, truncations, _ = self.vec_env.step(self._env_actions(actions))
        self.next_obs = self._obs_to_device(next_obs)
        self.next_action_masks_host = self.get_action_mask() if self.get_action_mask else None
        self.next_action_masks = self._masks_to_device(self.next_action_masks_host)
        return self._host(rewards), self._host(terminations) | self._host(truncations)

    def _finish(self, trajectories) -> TrajectoryRollout:
        rollout = TrajectoryRollout(
            self.device,
            trajectories,
            scale_advantage_by_values_accuracy=self.scale_advantage_by_values_accuracy,
            subaction_mask=self.subaction_mask,
            action_plane_space=getattr(self.vec_env, "action_plane_space", None),
        )
        self.store.reset()  # the rollout owns gathered copies; the step log is reused by the next rollout
        return rollout


class GuidedLearnerRolloutGenerator(_DeviceStepper):
    def __init__(
        self,
        learning_policy,
        vec_env,
        guide_policy,
        switch_range: int,
        n_steps: int = 2048,
        sde_sample_freq: int = -1,
        scale_advantage_by_values_accuracy: bool = False,
        full_batch_off_accelerator: bool = True,
        include_logp: bool = True,
        subaction_mask: Optional[Dict[int, Dict[int, int]]] = None,
    ) -> None:
        self.switch_range = switch_range
        guide_policy.eval()
        N = vec_env.num_envs
        # same draws, same order as guided_learner_rollout.py:52-60 (the env reset below draws nothing from np.random)
        self.traj_step_by_index = np.zeros(N, dtype=np.int32)
        self.switch_step_by_index = np.random.randint(0, self.switch_range, N, dtype=np.int32)
        super().__init__(learning_policy, vec_env, guide_policy, n_steps, sde_sample_freq,
                         scale_advantage_by_values_accuracy, full_batch_off_accelerator, include_logp, subaction_mask)
        self.policies_by_index = [self.guide_policy if s > 0 else self.learning_policy for s in self.switch_step_by_index]

    def rollout(self, gamma, gae_lambda) -> TrajectoryRollout:
        self.learning_policy.eval()
        self.learning_policy.reset_noise()
        self.guide_policy.reset_noise()
        N = self.num_envs
        builders = [TrajectoryBuilder(self.store, n) for n in range(N)]
        completed = []
        goal_steps, steps, s = self.n_steps * N, 0, 0
        while steps < goal_steps:
            if self.sde_sample_freq > 0 and s > 0 and s % self.sde_sample_freq == 0:
                self.learning_policy.reset_noise()
                self.guide_policy.reset_noise()
            s += 1
            obs, masks = self.next_obs, self.next_action_masks
            out: dict = {}
            by_guide = np.array([p is self.guide_policy for p in self.policies_by_index])
            if by_guide.any():
                self._step_policy(self.guide_policy, by_guide, obs, masks, out)
            if not by_guide.all():
                self._step_policy(self.learning_policy, ~by_guide, obs, masks, out)
            self.store.append(step_fields(obs, out["v"], out["logp"], out["a"], masks))
            rewards, dones = self._advance(out["a"])

            self.traj_step_by_index += 1
            for idx in range(N):  # guided_learner_rollout.py:133-171
                traj_step, switch_step, done = self.traj_step_by_index[idx], self.switch_step_by_index[idx], bool(dones[idx])
                if traj_step <= switch_step:
                    if done:
                        self.traj_step_by_index[idx] = 0
                        self.switch_step_by_index[idx] = np.random.randint(self.switch_range)
                    elif traj_step == switch_step:
                        self.policies_by_index[idx] = self.learning_policy
                    continue
                builders[idx].add(None, rewards[idx], done, None, None, None, None)
                steps += 1
                if done:
                    self.traj_step_by_index[idx] = 0
                    switch_step = np.random.randint(self.switch_range)
                    self.switch_step_by_index[idx] = switch_step
                    self.policies_by_index[idx] = self.guide_policy if switch_step > 0 else self.learning_policy
                    completed.append(builders[idx].trajectory(gamma, gae_lambda))
                    builders[idx].reset()

        next_values = self.learning_policy.value_device(self.next_obs).float()
        self.learning_policy.train()
        trajectories = completed + [b.trajectory(gamma, gae_lambda, next_values=next_values[n])
                                    for n, b in enumerate(builders) if len(b) > 0]
        return self._finish(trajectories)


class RandomGuidedLearnerRolloutGenerator(_DeviceStepper):
    def __init__(
        self,
        learning_policy,
        vec_env,
        guide_policy,
        guide_probability: float,
        n_steps: int = 2048,
        sde_sample_freq: int = -1,
        scale_advantage_by_values_accuracy: bool = False,
        full_batch_off_accelerator: bool = True,
        include_logp: bool = True,
        subaction_mask: Optional[Dict[int, Dict[int, int]]] = None,
        skip_no_action_steps: bool = False,
        num_envs_reset_every_rollout: int = 0,
    ) -> None:
        super().__init__(learning_policy, vec_env, guide_policy, n_steps, sde_sample_freq,
                         scale_advantage_by_values_accuracy, full_batch_off_accelerator, include_logp, subaction_mask)
        self.guide_probability = guide_probability
        self.skip_no_action_steps = skip_no_action_steps
        self.num_envs_reset_every_rollout = num_envs_reset_every_rollout
        if skip_no_action_steps:
            assert self.get_action_mask is not None, \
                f"skip_no_action_steps requires get_action_mask to be implemented on {vec_env}"

    def _merge_actions(self, step_actions, part, select: np.ndarray):
        """Rows `select` of the [N, ...] action tensor(s) sent to the env <- `part` (already full size)."""
        if step_actions is None or select.all():
            return _map(lambda t: t.clone(), part)
        rows = torch.from_numpy(np.nonzero(select)[0]).to(self.device)
        if isinstance(part, dict):
            for k, t in part.items():
                step_actions[k].index_copy_(0, rows, t.index_select(0, rows))
        else:
            step_actions.index_copy_(0, rows, part.index_select(0, rows))
        return step_actions

    def rollout(self, gamma, gae_lambda) -> TrajectoryRollout:
        self.learning_policy.eval()
        self.learning_policy.reset_noise()
        self.guide_policy.eval()
        self.guide_policy.reset_noise()
        N = self.num_envs
        builders = [DiscreteSkipsTrajectoryBuilder(self.store, n) for n in range(N)]
        completed = []
        zero_actions = None
        goal_steps, steps, s = self.n_steps * N, 0, 0
        while steps < goal_steps:
            if self.sde_sample_freq > 0 and s > 0 and s % self.sde_sample_freq == 0:
                self.learning_policy.reset_noise()
                self.guide_policy.reset_noise()
            s += 1
            obs, masks = self.next_obs, self.next_action_masks
            use_zero = (~has_actions(self.next_action_masks_host) if self.skip_no_action_steps and masks is not None
                        else np.full(N, False))
            use_learner = ~use_zero & (np.random.rand(N) >= self.guide_probability)
            use_guide = ~use_zero & ~use_learner
            step_actions = None if zero_actions is None else _map(torch.zeros_like, zero_actions)
            out: dict = {}
            if use_guide.any():
                self._step_policy(self.guide_policy, use_guide, obs, masks, out)
                step_actions = self._merge_actions(step_actions, out["a"], use_guide)
            if use_learner.any():
                self._step_policy(self.learning_policy, use_learner, obs, masks, out)
                step_actions = self._merge_actions(step_actions, out["a"], use_learner)
                self.store.append(step_fields(obs, out["v"], out["logp"], out["a"], masks))
            if step_actions is None:  # nobody acted yet in this generator's life: ask the learner for the action layout
                probe: dict = {}
                self._step_policy(self.learning_policy, np.full(N, True), obs, masks, probe)
                step_actions = _map(torch.zeros_like, probe["a"])
            if zero_actions is None:
                zero_actions = _map(torch.zeros_like, step_actions)
            if use_zero.any():
                rows = torch.from_numpy(np.nonzero(use_zero)[0]).to(self.device)
                _map(lambda t: t.index_fill_(0, rows, 0), step_actions)
            rewards, dones = self._advance(step_actions)

            for idx in np.where(use_guide | use_zero)[0]:
                builders[idx].step_no_add(rewards[idx], dones[idx], gamma)
            steps += int(use_learner.sum())
            for idx in np.where(use_learner)[0]:
                builders[idx].step_add(None, rewards[idx], dones[idx], None, None, None, None, gamma)
            for b in builders:
                if b.done:
                    if len(b) > 0:
                        completed.append(b.trajectory(gamma, gae_lambda))
                    b.reset()

        next_values = self.learning_policy.value_device(self.next_obs).float()
        self.learning_policy.train()
        self.guide_policy.train()
        trajectories = completed + [b.trajectory(gamma, gae_lambda, next_values=next_values[n])
                                    for n, b in enumerate(builders) if len(b) > 0]

        if self.num_envs_reset_every_rollout > 0:  # random_guided_learner_rollout.py:232-241
            k = self.num_envs_reset_every_rollout
            reset = np.zeros(N, dtype=np.bool_)
            reset[-k:] = True
            next_obs, action_mask, _ = self.vec_env.masked_reset(reset)
            self.next_obs[-k:] = torch.as_tensor(next_obs).to(self.device, dtype=self.next_obs.dtype)
            if self.next_action_masks is not None:
                if isinstance(self.next_action_masks, dict):
                    for key, dst in self.next_action_masks.items():
                        dst[-k:] = torch.as_tensor(action_mask[key]).to(self.device, dtype=torch.bool)
                        if isinstance(self.next_action_masks_host[key], np.ndarray):
                            self.next_action_masks_host[key][-k:] = action_mask[key]
                else:
                    self.next_action_masks[-k:] = torch.as_tensor(action_mask).to(self.device, dtype=torch.bool)
                    if isinstance(self.next_action_masks_host, np.ndarray):
                        self.next_action_masks_host[-k:] = action_mask
        return self._finish(trajectories)
