"""Synthetic VectorEnvs of the five BASELINE.json configs (SURVEY.md section 8 table, 8d inputs).

The environments of the reference (gym / JVM MicroRTS / Lux) are out of scope; the data path only
needs something that honours the VectorEnv duck type the rollout generator consumes
(rollout/sync_step_rollout.py:80-97,202-212): ``num_envs``, ``single_observation_space``,
``single_action_space``, optional ``action_plane_space``, ``reset()``, ``step(actions)``,
optional ``get_action_mask()``.  Observations / masks / rewards come from a seeded pool that is
generated once and cycled, so the env costs nothing per step and every run sees the same data.

``device=None``  -> a host env: numpy in / numpy out (the reference's contract; used for the
                    end-to-end measurement where inputs cross PCIe every step);
``device=cuda``  -> a device env: CUDA tensors in / out (Jux-style GPU env; nothing crosses PCIe).
"""
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import spaces


@dataclass(frozen=True)
class EnvSpec:
    name: str
    obs_shape: Tuple[int, ...]
    obs_dtype: str
    kind: str  # categorical | gaussian | gridnet
    n_actions: int = 0  # categorical: n; gaussian: act dim
    nvec: Tuple[int, ...] = ()
    map_hw: Tuple[int, int] = (0, 0)
    n_pick: int = 0
    n_values: int = 1
    unit_p: float = 0.06
    p_done: float = 1 / 200
    subaction_mask: Optional[Dict[int, Dict[int, int]]] = None


_GEN_ENVS = 128

MICRORTS_NVEC = (6, 4, 4, 4, 4, 7, 49)
LUX_NVEC = (4, 6, 4, 4, 5, 5)

SPECS: Dict[str, EnvSpec] = {
    "CartPole-v1": EnvSpec("CartPole-v1", (4,), "float32", "categorical", n_actions=2),
    "BreakoutNoFrameskip-v4": EnvSpec("BreakoutNoFrameskip-v4", (4, 84, 84), "uint8", "categorical", n_actions=4),
    "HalfCheetah-v4": EnvSpec("HalfCheetah-v4", (17,), "float32", "gaussian", n_actions=6),
    "Microrts-16x16": EnvSpec("Microrts-16x16", (74, 16, 16), "float32", "gridnet", nvec=MICRORTS_NVEC,
                              map_hw=(16, 16), unit_p=0.06, p_done=1 / 2000,
                              subaction_mask={0: {1: 1, 2: 2, 3: 3, 4: 4, 5: 4, 6: 5}}),
    "LuxAI_S2-64x64": EnvSpec("LuxAI_S2-64x64", (75, 64, 64), "float32", "gridnet", nvec=LUX_NVEC, map_hw=(64, 64),
                              n_pick=1, n_values=13, unit_p=0.02, p_done=1 / 1000,
                              subaction_mask={1: {2: 0, 3: 1, 4: 1, 5: 2}}),
}


def _cell_masks(rng: np.random.Generator, n: int, hw: int, nvec, unit_p: float) -> np.ndarray:
    S = int(sum(nvec))
    mask = rng.random((n, hw, S)) < 0.5
    start = 0
    for k in nvec:
        forced = rng.integers(0, k, size=(n, hw))
        np.put_along_axis(mask[..., start : start + k], forced[..., None], True, axis=-1)
        start += k
    return mask & (rng.random((n, hw, 1)) < unit_p)


class SyntheticVecEnv:
    def __init__(self, spec: EnvSpec, num_envs: int, seed: int = 0, device: Optional[torch.device] = None,
                 pool: int = 8) -> None:
        self.spec, self.num_envs, self.device, self.pool = spec, int(num_envs), device, int(pool)
        rng = np.random.default_rng(seed)
        N_all, V = self.num_envs, spec.n_values
        # Large env counts (the 1024-env Lux config) draw the pool for a block of `_GEN_ENVS` envs and repeat it over
        # the env axis (on the device for a device env): the data path moves the same bytes either way, and
        # drawing 5 GB of Bernoulli planes on the host would take longer than the benchmark.
        N = N_all if N_all <= _GEN_ENVS or N_all % _GEN_ENVS else _GEN_ENVS
        reps = N_all // N
        if spec.obs_dtype == "uint8":
            obs = rng.integers(0, 256, size=(pool, N) + spec.obs_shape, dtype=np.uint8)
            self.single_observation_space = spaces.Box(0, 255, spec.obs_shape, np.uint8)
        elif spec.kind == "gridnet":
            obs = (rng.random((pool, N) + spec.obs_shape) < 0.1).astype(np.float32)
            self.single_observation_space = spaces.Box(0, 1, spec.obs_shape, np.float32)
        else:
            obs = rng.standard_normal((pool, N) + spec.obs_shape, dtype=np.float32)
            self.single_observation_space = spaces.Box(-np.inf, np.inf, spec.obs_shape, np.float32)
        rewards = rng.standard_normal((pool, N) if V == 1 else (pool, N, V), dtype=np.float32)
        dones = rng.random((pool, N)) < spec.p_done
        masks = None
        self.action_plane_space = None
        if spec.kind == "categorical":
            self.single_action_space = spaces.Discrete(spec.n_actions)
        elif spec.kind == "gaussian":
            self.single_action_space = spaces.Box(-1.0, 1.0, (spec.n_actions,), np.float32)
        else:
            hw = spec.map_hw[0] * spec.map_hw[1]
            self.action_plane_space = spaces.MultiDiscrete(spec.nvec)
            per_pos = spaces.MultiDiscrete(np.tile(np.asarray(spec.nvec), hw))
            cells = np.stack([_cell_masks(rng, N, hw, spec.nvec, spec.unit_p) for _ in range(pool)])
            if spec.n_pick:
                self.single_action_space = spaces.Dict(
                    {"per_position": per_pos, "pick_position": spaces.MultiDiscrete([hw] * spec.n_pick)})
                pick = rng.random((pool, N, spec.n_pick, hw)) < 0.05
                pick &= ~(rng.random((pool, N, spec.n_pick, 1)) < 0.25)
                masks = {"per_position": cells, "pick_position": pick}
            else:
                self.single_action_space = per_pos
                masks = cells
        N = N_all
        if reps > 1 and device is None:
            tile = lambda a: np.tile(a, (1, reps) + (1,) * (a.ndim - 2))
            obs, rewards, dones = tile(obs), tile(rewards), tile(dones)
            if masks is not None:
                masks = {k: tile(v) for k, v in masks.items()} if isinstance(masks, dict) else tile(masks)
        self._obs, self._rewards, self._dones, self._masks = obs, rewards, dones, masks
        if device is not None:
            # Device env: the pool lives in HBM, the slot counter is a device scalar and every output is
            # a fixed tensor refreshed in place (index_select with out=), so a step is a short, sync-free
            # kernel sequence with static addresses -- capturable in a CUDA graph like a Jux-style env.
            up = lambda a: torch.from_numpy(a).to(device)
            to = up if reps == 1 else (lambda a: up(a).repeat((1, reps) + (1,) * (a.ndim - 2)))
            self._obs, self._rewards, self._dones = to(obs), to(rewards), to(dones)
            if masks is not None:
                self._masks = {k: to(v) for k, v in masks.items()} if isinstance(masks, dict) else to(masks)
            self._no_trunc = torch.zeros(N, dtype=torch.bool, device=device)
            # [this step's slot, the next one]: rewards / dones are read at the first, the observation and masks the
            # step returns at the second; both advance together
            self._slots = torch.tensor([0, 1 % self.pool], dtype=torch.int64, device=device)
            self._out_obs = self._out_rew = self._out_done = self._out_masks = None
        else:
            self._no_trunc = np.zeros(N, dtype=np.bool_)
        self._t = 0

    @property
    def unwrapped(self):
        return self

    def _slot(self) -> int:
        return self._t % self.pool

    def _gather_device_outputs(self, scalars: bool) -> None:
        """The pool rows of the current slot pair -> this step's outputs: one gather launch for the observation and the
        masks (slot 1), one for rewards and dones (slot 0) -- the minibatch gather kernel with a single index."""
        from .. import ops

        wide = [self._obs]
        if self._masks is not None:
            wide += list(self._masks.values()) if isinstance(self._masks, dict) else [self._masks]
        got = ops.gather_rows(wide, self._slots[1:2])
        self._out_obs = got[0]
        if self._masks is not None:
            self._out_masks = dict(zip(self._masks, got[1:])) if isinstance(self._masks, dict) else got[1]
        if scalars:
            self._out_rew, self._out_done = ops.gather_rows([self._rewards, self._dones], self._slots[0:1])

    def reset(self, **_kwargs):
        self._t = 0
        if self.device is not None:
            self._slots.copy_(torch.tensor([self.pool - 1, 0], dtype=torch.int64))  # the observation of slot 0 comes first
            self._gather_device_outputs(scalars=False)
            self._slots.add_(1).remainder_(self.pool)
            return self._out_obs[0], {}
        return self._obs[0], {}

    def step(self, actions):
        """-> (next_obs, rewards, terminations, truncations, infos); actions are accepted and ignored."""
        if self.device is not None:
            self._gather_device_outputs(scalars=True)
            self._slots.add_(1).remainder_(self.pool)
            return self._out_obs[0], self._out_rew[0], self._out_done[0], self._no_trunc, {}
        k = self._slot()
        self._t += 1
        return self._obs[self._slot()], self._rewards[k], self._dones[k], self._no_trunc, {}

    def get_action_mask(self):
        if self._masks is None:
            return None
        if self.device is not None:
            return ({n: m[0] for n, m in self._out_masks.items()} if isinstance(self._out_masks, dict)
                    else self._out_masks[0])
        k = self._slot()
        return {n: m[k] for n, m in self._masks.items()} if isinstance(self._masks, dict) else self._masks[k]

    def close(self) -> None:
        pass


def make_synthetic_env(name: str, num_envs: int, seed: int = 0, device=None, pool: int = 8) -> SyntheticVecEnv:
    return SyntheticVecEnv(SPECS[name], num_envs, seed, device, pool)
