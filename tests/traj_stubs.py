"""Scripted VectorEnv + closed-form stub policies for the trajectory-rollout fixtures.

TEST INFRASTRUCTURE.  The same script drives the live reference's rollout generators (numpy side,
``tests/golden/make_golden_trajectory_rollouts.py``, build container only) and this repo's device
generators (``step_device`` / ``value_device``, GPU tests), so both see identical observations,
rewards, dones and masks and the stub policies return bit-identical float32 values on either side
(every formula is a chain of single float32 operations).
"""
from typing import NamedTuple

import numpy as np

HW, NVEC = 4, (3, 2)
S, A, D = sum(NVEC), len(NVEC), 3
GATES = {1: {0: 1}}  # head 1 only counts / contributes when head 0 chose value 1


class _Box:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, np.int64)
        self.shape, self.dtype = self.nvec.shape, np.dtype(np.int64)

    def __len__(self):
        return len(self.nvec)


class ScriptedVecEnv:
    """Cycles through a seeded pool; ``step`` ignores the actions it is given (but records a scripted
    ``last_action``, the field ReferenceAIRolloutGenerator reads back)."""

    def __init__(self, num_envs: int, seed: int, pool: int = 64, p_done: float = 0.12, p_empty: float = 0.3,
                 space_factory=None):
        rng = np.random.default_rng(seed)
        N = self.num_envs = num_envs
        self.pool = pool
        self.obs = rng.standard_normal((pool, N, D)).astype(np.float32)
        self.rewards = rng.standard_normal((pool, N)).astype(np.float32)
        self.dones = rng.random((pool, N)) < p_done
        cell = rng.random((pool, N, HW, 1)) < 0.6
        self.masks = (rng.random((pool, N, HW, S)) < 0.6) & cell
        self.masks[rng.random((pool, N)) < p_empty] = False  # whole env without a legal action on that step
        self.ai_actions = np.stack([rng.integers(0, n, (pool, N, HW)) for n in NVEC], -1).astype(np.int64)
        box, md = (space_factory or (_Box, _MultiDiscrete))
        self.single_observation_space = box((D,), np.float32)
        self.single_action_space = md(np.tile(np.asarray(NVEC), HW))
        self.action_plane_space = md(NVEC)
        self.t = 0
        self.last_action = None

    @property
    def unwrapped(self):
        return self

    def reset(self, **_):
        self.t = 0
        return self.obs[0].copy(), {}

    def step(self, actions):
        k = self.t % self.pool
        self.t += 1
        self.last_action = self.ai_actions[k]
        return (self.obs[self.t % self.pool].copy(), self.rewards[k].copy(), self.dones[k].copy(),
                np.zeros(self.num_envs, np.bool_), {})

    def get_action_mask(self):
        return self.masks[self.t % self.pool].copy()

    def masked_reset(self, mask):
        """Fresh episodes for the envs in `mask`: observations / masks that differ from what step() just returned
        (another pool slot, negated), so that a generator which forgets to fold them in is caught."""
        k = (self.t + 7) % self.pool
        return (-self.obs[k][mask]).copy(), self.masks[k][mask].copy(), {}


class Step(NamedTuple):
    a: object
    v: object
    logp_a: object
    clamped_a: object


class StubPolicy:
    """value = obs0 * c + obs1, logp = -|obs2| - c, action of head h = first legal entry (0 if none).
    ``step`` / ``value`` are the reference's numpy contract; ``step_device`` / ``value_device`` this repo's."""

    def __init__(self, c: float, device="cpu"):
        import torch

        self.c = np.float32(c)
        self.device = torch.device(device)
        self.action_shape = (HW, A)
        self.value_shape = ()
        self.kind = "gridnet"  # per-cell actions are uint8 on the device (rollout buffers are typed from this)
        self.training = True

    # control-plane no-ops the generators call
    def eval(self):
        self.training = False

    def train(self, mode: bool = True):
        self.training = mode

    def reset_noise(self, *a, **k):
        pass

    # numpy contract (rl_algo_impls/shared/policy/actor_critic.py:306-318)
    def value(self, obs):
        obs = np.asarray(obs, np.float32)
        return obs[:, 0] * self.c + obs[:, 1]

    def step(self, obs, action_masks=None):
        obs = np.asarray(obs, np.float32)
        acts, start = [], 0
        for n in NVEC:
            acts.append(np.argmax(action_masks[..., start:start + n], axis=-1))
            start += n
        a = np.stack(acts, -1).astype(np.int64)
        return Step(a, self.value(obs), -np.abs(obs[:, 2]) - self.c, a)

    # device contract
    def value_device(self, obs):
        return obs[:, 0] * float(self.c) + obs[:, 1]

    def step_device(self, obs, action_masks=None, offset_dev=None):
        import torch

        acts, start = [], 0
        for n in NVEC:
            acts.append(torch.argmax(action_masks[..., start:start + n].to(torch.uint8), dim=-1))
            start += n
        a = torch.stack(acts, -1).to(torch.uint8)
        return a, self.value_device(obs), -torch.abs(obs[:, 2]) - float(self.c)
