"""Drive the UNMODIFIED reference learner on the CPU.  TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).

``build(cfg, ...)`` wires the reference's own classes exactly like ``rl_algo_impls/runner/train.py:104-217`` does
from the YAML -- ``ActorCritic(env, **policy_hyperparams)`` (shared/policy/actor_critic.py:109), ``PPO(policy, device,
tb_writer, **algo_hyperparams)`` (ppo/ppo.py:107), ``SyncStepRolloutGenerator(policy, env, **rollout_hyperparams)``
(rollout/sync_step_rollout.py:14) -- over this repo's host (numpy) synthetic VectorEnv, whose spaces are re-expressed
as the shim's ``gymnasium.spaces`` so that the reference's ``isinstance`` checks hold.  Nothing of this repo's
kernels, policy classes or learner is on that path: ``bench.py --impl reference`` times ``PPO.learn_epoch`` of
the reference itself; the reference-backed tests use it as a second oracle.

The reference is imported through oracle/ref_shim.py from /root/reference (build container) or oracle/_ref (the
byte-identical copy made by oracle/make_ref.sh, which travels to the GPU box).
"""
import time
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import ref_shim


class _Writer:
    """What PPO.learn_epoch needs of the tensorboard SummaryWrapper (shared/callbacks/summary_wrapper.py:8-25)."""

    def __init__(self):
        self.scalars: Dict[str, Any] = {}

    def add_scalar(self, name, value, *a, **k):
        self.scalars[name] = value

    def on_steps(self, *a, **k):
        pass


class RefEnvAdapter:
    """A host SyntheticVecEnv seen through gymnasium-style spaces (the shim's classes).  Dict action masks are
    handed over the way the reference's Lux envs do -- an object array of per-env dicts that
    ``batch_dict_keys`` (shared/tensor_utils.py:75-80) re-batches."""

    def __init__(self, env):
        self._env = env
        self.num_envs = env.num_envs
        sp = env.single_observation_space
        self.single_observation_space = ref_shim.Box(sp.low, sp.high, sp.shape, sp.dtype)
        self.single_action_space = self._convert(env.single_action_space)
        aps = getattr(env, "action_plane_space", None)
        if aps is not None:
            self.action_plane_space = ref_shim.MultiDiscrete(aps.nvec)

    @staticmethod
    def _convert(space):
        if hasattr(space, "spaces"):
            return ref_shim.DictSpace({k: RefEnvAdapter._convert(v) for k, v in space.spaces.items()})
        if hasattr(space, "nvec"):
            return ref_shim.MultiDiscrete(space.nvec)
        if hasattr(space, "n"):
            return ref_shim.Discrete(space.n)
        return ref_shim.Box(space.low, space.high, space.shape, space.dtype)

    @property
    def unwrapped(self):
        return self

    def reset(self, **kw):
        return self._env.reset(**kw)

    def step(self, actions):
        return self._env.step(actions)

    def get_action_mask(self):
        m = self._env.get_action_mask()
        if isinstance(m, dict):
            out = np.empty(self.num_envs, dtype=object)
            for i in range(self.num_envs):
                out[i] = {k: v[i] for k, v in m.items()}
            return out
        return m

    def close(self):
        pass


def build(cfg, n_envs: Optional[int] = None, n_steps: Optional[int] = None, seed: int = 0, pool: int = 2,
          algo_overrides: Optional[dict] = None):
    """(env, policy, rollout_generator, algo) of the reference for a RunConfig (rl_algo_impls_b200/configs.py)."""
    ref_shim.install()
    from rl_algo_impls.ppo.ppo import PPO
    from rl_algo_impls.rollout.sync_step_rollout import SyncStepRolloutGenerator
    from rl_algo_impls.shared.policy.actor_critic import ActorCritic

    from rl_algo_impls_b200.envs import make_synthetic_env  # the env only: numpy in, numpy out

    torch.manual_seed(seed)
    np.random.seed(seed)
    env = RefEnvAdapter(make_synthetic_env(cfg.env, n_envs or cfg.n_envs, seed=seed, device=None, pool=pool))
    device = torch.device("cpu")
    policy = ActorCritic(env, **cfg.policy).to(device)
    rollout_kw = dict(cfg.rollout)
    if n_steps:
        rollout_kw["n_steps"] = n_steps
    if not env._env.spec.kind == "gridnet":
        rollout_kw.pop("subaction_mask", None)
    algo = PPO(policy, device, _Writer(), **{**cfg.algo, **(algo_overrides or {})})
    gen = SyncStepRolloutGenerator(policy, env, **rollout_kw)
    return env, policy, gen, algo


def time_learn_epochs(cfg, steps: int, warmup: int, threads: int, n_envs: Optional[int] = None,
                      n_steps: Optional[int] = None, algo_overrides: Optional[dict] = None, budget_s: Optional[float] = None):
    """env-steps/s of the reference's PPO.learn_epoch (ppo/ppo.py:214-439) on the host cores.
    Returns (value, ms_per_step, steps actually timed, warm-ups actually run, sample description).  ``budget_s``
    bounds the run: timing stops after the first step that ends beyond it (at least one step is always timed)."""
    torch.set_num_threads(threads)
    env, policy, gen, algo = build(cfg, n_envs, n_steps, algo_overrides=algo_overrides)
    total = gen.n_steps * env.num_envs
    t_begin = time.perf_counter()
    done_warm = 0
    for _ in range(warmup):
        algo.learn_epoch(0, 1 << 40, gen, None)
        done_warm += 1
        if budget_s is not None and time.perf_counter() - t_begin > 0.4 * budget_s:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(1, steps)):
        algo.learn_epoch(0, 1 << 40, gen, None)
        done += 1
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    dt = time.perf_counter() - t0
    sample = (f"unmodified reference (PPO.learn_epoch + SyncStepRolloutGenerator + ActorCritic) on {threads} threads: "
              f"{env.num_envs} envs x {gen.n_steps} steps, batch {algo.batch_size}, {algo.n_epochs} epochs; "
              f"{done} learn_epochs timed after {done_warm} warm-up")
    return total * done / dt, dt / done * 1e3, done, done_warm, sample
