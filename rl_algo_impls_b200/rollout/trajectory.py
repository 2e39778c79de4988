"""Trajectories (ragged, per-env segments of experience) whose bulk data stays in HBM.

Mirrors ``rl_algo_impls/rollout/trajectory.py:11-103`` (``Trajectory``, ``TrajectoryBuilder``, ``batch_actions``)
and ``rollout/discrete_skips_trajectory_builder.py:12-109`` (``DiscreteSkipsTrajectoryBuilder``): same method
names, argument meaning and assertions.  What changes is where the rows live and when GAE runs:

* the reference appends one numpy row per env per step to Python lists and, at ``trajectory()``, stacks them
  and runs the T-iteration numpy GAE loop once per trajectory (hundreds of trajectories per rollout);
* here a guided rollout writes every vec-env step ONCE into a ``StepStore`` -- ``[capacity, N, ...]`` CUDA
  tensors, one K0 launch per step -- and a builder only records row numbers (``s * N + env``) next to the
  scalar host bookkeeping the control flow needs anyway (rewards, dones, steps_elapsed: the reference's exact
  numpy arithmetic, including the in-place ``reward * gamma ** steps_elapsed`` accumulation).  ``trajectory()``
  returns a ``Trajectory`` that carries the GAE *recipe*; ``TrajectoryRollout`` gathers the rows of all
  trajectories with one K3 launch and scans all of them with ONE K1b launch (``ops.gae_segments``).
  ``Trajectory.advantages`` read on its own runs the one-segment scan on demand.

The builders also accept explicit rows (numpy or tensors; the reference's ``add`` / ``step_add`` signature):
those are stacked onto the device at ``trajectory()``.  There is no CPU path: GAE always runs in K1b.
"""
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import ops

TensorOrDict = Union[torch.Tensor, Dict[str, torch.Tensor]]


def _map(fn, t):
    if t is None:
        return None
    if isinstance(t, dict):
        return {k: fn(v) for k, v in t.items()}
    return fn(t)


class StepStore:
    """Append-only log of vec-env steps on the device: row ``s * N + n`` of every field is env n at step s."""

    def __init__(self, device: torch.device, num_envs: int, capacity: int = 64) -> None:
        self.device, self.num_envs = torch.device(device), int(num_envs)
        if self.device.type != "cuda":
            raise RuntimeError("StepStore keeps rollout steps in HBM: it needs a CUDA device (no CPU path)")
        self.capacity = int(capacity)
        self.steps = 0
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._names: List[str] = []
        self._buffers: List[torch.Tensor] = []

    def _allocate(self, fields: Dict[str, torch.Tensor]) -> None:
        self._names = list(fields)
        self._buffers = [torch.zeros((self.capacity,) + tuple(t.shape), dtype=t.dtype, device=self.device)
                         for t in fields.values()]

    def _grow(self) -> None:
        self.capacity *= 2
        grown = []
        for b in self._buffers:
            g = torch.zeros((self.capacity,) + tuple(b.shape[1:]), dtype=b.dtype, device=self.device)
            g[: b.shape[0]].copy_(b)
            grown.append(g)
        self._buffers = grown

    def append(self, fields: Dict[str, torch.Tensor]) -> int:
        """Write one vec-env step (every tensor is ``[N, ...]``); returns its step number s."""
        if not self._buffers:
            self._allocate(fields)
        assert list(fields) == self._names, f"StepStore fields changed: {list(fields)} vs {self._names}"
        if self.steps == self.capacity:
            self._grow()
        self._step_dev.fill_(self.steps)
        ops.rollout_store_step([t.contiguous() for t in fields.values()], self._buffers, self._step_dev)
        self.steps += 1
        return self.steps - 1

    def reset(self) -> None:
        self.steps = 0

    def flat(self) -> Dict[str, torch.Tensor]:
        """Every field as ``[capacity * N, ...]`` (a view)."""
        return {n: b.reshape((-1,) + tuple(b.shape[2:])) for n, b in zip(self._names, self._buffers)}

    def gather(self, rows: torch.Tensor) -> Dict[str, torch.Tensor]:
        flat = self.flat()
        return dict(zip(flat, ops.gather_rows(list(flat.values()), rows)))


def _split_fields(fields: Dict[str, torch.Tensor], prefix: str) -> Optional[TensorOrDict]:
    """'actions' -> tensor, 'actions.per_position' / 'actions.pick_position' -> dict, absent -> None."""
    if prefix in fields:
        return fields[prefix]
    sub = {k[len(prefix) + 1:]: v for k, v in fields.items() if k.startswith(prefix + ".")}
    return sub or None


def step_fields(obs, values, logprobs, actions, action_masks) -> Dict[str, torch.Tensor]:
    """The flat field dict a StepStore holds, from the per-step tensors of a generator."""
    out = {"obs": obs, "values": values, "logprobs": logprobs}
    for name, t in (("actions", actions), ("action_masks", action_masks)):
        if isinstance(t, dict):
            out.update({f"{name}.{k}": v for k, v in t.items()})
        elif t is not None:
            out[name] = t
    return out


class Trajectory:
    """rollout/trajectory.py:11-21.  ``obs / values / logprobs / actions / action_masks`` are device tensors
    (gathered from the store on first use), ``advantages`` is computed by K1b on first use."""

    def __init__(self, *, length: int, rewards: np.ndarray, gamma, gae_lambda, next_done: bool,
                 next_values: Optional[torch.Tensor], episode_starts: Optional[np.ndarray] = None,
                 steps_elapsed: Optional[np.ndarray] = None, store: Optional[StepStore] = None,
                 rows: Optional[np.ndarray] = None, fields: Optional[Dict[str, torch.Tensor]] = None) -> None:
        assert (store is None) != (fields is None), "a Trajectory is either store-backed or materialised"
        assert (episode_starts is None) != (steps_elapsed is None)
        self.length = int(length)
        self.rewards = np.asarray(rewards, dtype=np.float32)
        self.gamma, self.gae_lambda = gamma, gae_lambda
        self.next_done, self.next_values = bool(next_done), next_values
        self.episode_starts, self.steps_elapsed = episode_starts, steps_elapsed
        self.store, self.rows, self._fields = store, rows, fields
        self._advantages: Optional[torch.Tensor] = None

    def __len__(self) -> int:
        return self.length

    @property
    def device(self) -> torch.device:
        return self.store.device if self.store is not None else self._fields["values"].device

    def fields(self) -> Dict[str, torch.Tensor]:
        if self._fields is None:
            rows = torch.from_numpy(np.asarray(self.rows, dtype=np.int64)).to(self.store.device)
            self._fields = self.store.gather(rows)
        return self._fields

    obs = property(lambda self: self.fields()["obs"])
    values = property(lambda self: self.fields()["values"])
    logprobs = property(lambda self: self.fields()["logprobs"])
    actions = property(lambda self: _split_fields(self.fields(), "actions"))
    action_masks = property(lambda self: _split_fields(self.fields(), "action_masks"))

    @property
    def advantages(self) -> torch.Tensor:
        if self._advantages is None:
            self._advantages, _ = segmented_gae([self], self.values.float())
        return self._advantages


def segmented_gae(trajectories: Sequence[Trajectory], values: torch.Tensor):
    """(advantages, returns) of the concatenated trajectories in ONE K1b launch.  ``values`` is the float32
    concatenation ``[total(, V)]`` of their value rows, already on the device."""
    dev = values.device
    first = trajectories[0]
    skips = first.steps_elapsed is not None
    lengths = np.array([len(t) for t in trajectories], dtype=np.int64)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)).to(dev)
    rewards = torch.from_numpy(np.concatenate([t.rewards for t in trajectories])).to(dev)
    head_shape = tuple(values.shape[1:])
    next_values = torch.zeros((len(trajectories),) + head_shape, dtype=torch.float32, device=dev)
    have = [i for i, t in enumerate(trajectories) if t.next_values is not None]
    if have:
        stacked = torch.stack([torch.as_tensor(trajectories[i].next_values, device=dev).float().reshape(head_shape)
                               for i in have])
        next_values[torch.tensor(have, device=dev)] = stacked
    next_done = torch.from_numpy(np.array([t.next_done for t in trajectories], dtype=np.bool_)).to(dev)
    if skips:
        assert all(t.steps_elapsed is not None for t in trajectories), "mixed trajectory kinds"
        steps = torch.from_numpy(np.concatenate([t.steps_elapsed for t in trajectories]).astype(np.int32)).to(dev)
        return ops.gae_segments(rewards, values.contiguous(), offsets, next_done, next_values, first.gamma,
                                first.gae_lambda, steps_elapsed=steps)
    assert all(t.episode_starts is not None for t in trajectories), "mixed trajectory kinds"
    starts = torch.from_numpy(np.concatenate([t.episode_starts for t in trajectories]).astype(np.bool_)).to(dev)
    return ops.gae_segments(rewards, values.contiguous(), offsets, next_done, next_values, first.gamma,
                            first.gae_lambda, episode_starts=starts)


def _row_to_tensor(x, device: torch.device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device)
    return torch.as_tensor(np.asarray(x)).to(device)


def batch_actions(rows: List, device: torch.device):
    """rollout/trajectory.py:98-103 on the device: a list of per-step rows -> one ``[L, ...]`` tensor (or dict)."""
    if rows[0] is None:
        return None
    if isinstance(rows[0], dict):
        return {k: torch.stack([_row_to_tensor(r[k], device) for r in rows]) for k in rows[0]}
    return torch.stack([_row_to_tensor(r, device) for r in rows])


class _BuilderBase:
    """Rows are either references into a StepStore (``store`` given: ``obs`` etc. are ignored and may be None)
    or explicit per-step rows (the reference's call signature)."""

    def __init__(self, store: Optional[StepStore] = None, env_index: int = 0,
                 device: Optional[torch.device] = None) -> None:
        self.store, self.env_index = store, int(env_index)
        self.device = store.device if store is not None else torch.device(device if device is not None else "cuda")
        self.reset()

    def __len__(self) -> int:
        return len(self.rows) if self.store is not None else len(self.obs)

    def _reset_rows(self) -> None:
        self.rows: List[int] = []
        self.obs: List = []
        self.values: List = []
        self.logprobs: List = []
        self.actions: List = []
        self.action_masks: List = []

    def _add_row(self, obs, value, logprob, action, action_mask, step: Optional[int]) -> None:
        if self.store is not None:
            s = self.store.steps - 1 if step is None else int(step)
            self.rows.append(s * self.store.num_envs + self.env_index)
            return
        self.obs.append(obs), self.values.append(value), self.logprobs.append(logprob)
        self.actions.append(action), self.action_masks.append(action_mask)

    def _payload(self) -> dict:
        if self.store is not None:
            return dict(store=self.store, rows=np.asarray(self.rows, dtype=np.int64))
        dev = self.device
        fields = step_fields(batch_actions(self.obs, dev), batch_actions(self.values, dev).float(),
                             batch_actions(self.logprobs, dev).float(), batch_actions(self.actions, dev),
                             batch_actions(self.action_masks, dev))
        return dict(fields=fields)


class TrajectoryBuilder(_BuilderBase):
    """rollout/trajectory.py:24-92."""

    def add(self, obs, reward, done: bool, value, logprob, action, action_mask, step: Optional[int] = None) -> None:
        self._add_row(obs, value, logprob, action, action_mask, step)
        self.rewards.append(reward)
        self.dones.append(bool(done))

    def reset(self) -> None:
        self._reset_rows()
        self.rewards: List = []
        self.dones: List[bool] = []

    def trajectory(self, gamma, gae_lambda, next_values: Optional[torch.Tensor] = None) -> Trajectory:
        dones = np.array(self.dones, dtype=np.bool_)
        # trajectory.py:70-73: the first flag is never read by the scan
        episode_starts = np.concatenate([[True], dones[:-1]])
        return Trajectory(length=len(self), rewards=np.array(self.rewards, dtype=np.float32), gamma=gamma,
                          gae_lambda=gae_lambda, next_done=bool(dones[-1]), next_values=next_values,
                          episode_starts=episode_starts, **self._payload())


class DiscreteSkipsTrajectoryBuilder(_BuilderBase):
    """rollout/discrete_skips_trajectory_builder.py:12-109: steps taken by another policy (or skipped because no
    action was legal) fold their reward into the last kept step, discounted by gamma ** steps_elapsed."""

    def reset(self) -> None:
        self._reset_rows()
        self.rewards: List = []
        self.done = False
        self.steps_elapsed: List[int] = []

    def step_no_add(self, reward, done: bool, gamma) -> None:
        assert not self.done, "Shouldn't be stepping a done trajectory"
        if self.rewards:
            self.rewards[-1] += reward * gamma ** self.steps_elapsed[-1]  # numpy float32 arithmetic, as the reference
        if self.steps_elapsed:
            self.steps_elapsed[-1] += 1
        self.done = bool(done)

    def step_add(self, obs, reward, done: bool, value, logprob, action, action_mask, gamma,
                 step: Optional[int] = None) -> None:
        assert not self.done, "Shouldn't be adding to a done trajectory"
        self._add_row(obs, value, logprob, action, action_mask, step)
        self.rewards.append(np.zeros_like(reward))
        self.steps_elapsed.append(0)
        self.step_no_add(reward, done, gamma)

    def trajectory(self, gamma, gae_lambda, next_values: Optional[torch.Tensor] = None) -> Trajectory:
        assert self.done or next_values is not None, "Need next_values if trajectory isn't done"
        return Trajectory(length=len(self), rewards=np.array(self.rewards, dtype=np.float32), gamma=gamma,
                          gae_lambda=gae_lambda, next_done=self.done, next_values=None if self.done else next_values,
                          steps_elapsed=np.array(self.steps_elapsed, dtype=np.int32), **self._payload())
