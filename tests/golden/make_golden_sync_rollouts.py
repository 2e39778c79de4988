"""Golden fixtures for SyncStepRolloutGenerator's masked resets (rollout/sync_step_rollout.py:152-278) from the LIVE,
UNMODIFIED reference:

    python tests/golden/make_golden_sync_rollouts.py        (build container only)

Runs the reference generator over the scripted env / stub policy of ``tests/traj_stubs.py`` with
``num_envs_reset_every_rollout``, ``rolling_num_envs_reset_every_rollout`` (crossing a re-draw of its permutation) and
``random_num_envs_reset_every_rollout``, three rollouts each, and stores the flat Batch every rollout hands to the
learner plus the ``next_obs`` / next masks the generator carries into the following rollout (what the resets rewrite).
The GPU test replays the same script through this repo's device-resident generator.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _ref_shim  # noqa: E402

_ref_shim.install()

from rl_algo_impls.rollout.sync_step_rollout import SyncStepRolloutGenerator as RefSync  # noqa: E402

from tests.traj_stubs import GATES, ScriptedVecEnv, StubPolicy  # noqa: E402

CASES = {
    "fixed": dict(N=6, n_steps=5, env_seed=21, np_seed=31, gamma=0.97, lam=0.9, kw=dict(num_envs_reset_every_rollout=2)),
    "rolling": dict(N=6, n_steps=4, env_seed=22, np_seed=32, gamma=0.99, lam=0.95,
                    kw=dict(rolling_num_envs_reset_every_rollout=4)),
    "random": dict(N=8, n_steps=4, env_seed=23, np_seed=33, gamma=0.99, lam=0.95,
                   kw=dict(random_num_envs_reset_every_rollout=2)),
    "none": dict(N=4, n_steps=6, env_seed=24, np_seed=34, gamma=0.98, lam=0.8, kw=dict()),
}
ROLLOUTS = 3


def spaces():
    return (lambda shape, dtype: _ref_shim.Box(-np.inf, np.inf, shape, dtype)), _ref_shim.MultiDiscrete


def run_case(name, c):
    env = ScriptedVecEnv(c["N"], c["env_seed"], space_factory=spaces())
    np.random.seed(c["np_seed"])  # the rolling permutation is drawn in the constructor
    gen = RefSync(StubPolicy(0.5), env, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    out = {}
    for r in range(ROLLOUTS):
        ro = gen.rollout(gamma=c["gamma"], gae_lambda=c["lam"])
        b = ro.batch()
        out.update({
            f"r{r}.obs": b.obs.numpy(), f"r{r}.actions": b.actions.numpy(), f"r{r}.action_masks": b.action_masks.numpy(),
            f"r{r}.values": b.values.numpy(), f"r{r}.logprobs": b.logprobs.numpy(), f"r{r}.advantages": b.advantages.numpy(),
            f"r{r}.returns": b.returns.numpy(), f"r{r}.next_obs": np.asarray(gen.next_obs),
            f"r{r}.next_action_masks": np.asarray(gen.next_action_masks),
            f"r{r}.next_episode_starts": np.asarray(gen.next_episode_starts),
        })
        out = {k: np.array(v) for k, v in out.items()}  # the batch aliases the generator's reused buffers
        print(f"{name} rollout {r}: {ro.total_steps} rows")
    return out


if __name__ == "__main__":
    out = {}
    for name, c in CASES.items():
        for k, v in run_case(name, c).items():
            out[f"{name}.{k}"] = v
    path = os.path.join(HERE, "sync_rollouts.npz")
    np.savez_compressed(path, **out)
    print(f"wrote sync_rollouts.npz ({os.path.getsize(path) / 1024:.1f} KB)")
