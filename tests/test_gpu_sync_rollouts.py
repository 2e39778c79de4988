"""SyncStepRolloutGenerator's masked resets on the device vs the live reference (fixture:
tests/golden/sync_rollouts.npz, made by tests/golden/make_golden_sync_rollouts.py from the unmodified
rollout/sync_step_rollout.py:152-278 over tests/traj_stubs.py): ``num_envs_reset_every_rollout``,
``rolling_num_envs_reset_every_rollout`` (its permutation comes from numpy's global generator, seeded alike) and
``random_num_envs_reset_every_rollout``.  Three consecutive rollouts each: rows, GAE and the carried-over next
observation / masks are bit-identical."""
import os

import numpy as np
import pytest
import torch

from tests.traj_stubs import GATES, ScriptedVecEnv, StubPolicy

pytestmark = pytest.mark.gpu

CASES = {
    "fixed": dict(N=6, n_steps=5, env_seed=21, np_seed=31, gamma=0.97, lam=0.9, kw=dict(num_envs_reset_every_rollout=2)),
    "rolling": dict(N=6, n_steps=4, env_seed=22, np_seed=32, gamma=0.99, lam=0.95,
                    kw=dict(rolling_num_envs_reset_every_rollout=4)),
    "random": dict(N=8, n_steps=4, env_seed=23, np_seed=33, gamma=0.99, lam=0.95,
                   kw=dict(random_num_envs_reset_every_rollout=2)),
    "none": dict(N=4, n_steps=6, env_seed=24, np_seed=34, gamma=0.98, lam=0.8, kw=dict()),
}


@pytest.mark.parametrize("cuda_graph", [False, True], ids=["eager", "graphed"])
@pytest.mark.parametrize("name", list(CASES))
def test_sync_generator_masked_resets_match_the_reference(cuda, name, cuda_graph):
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sync_rollouts.npz"))
    c = CASES[name]
    env = ScriptedVecEnv(c["N"], c["env_seed"])
    np.random.seed(c["np_seed"])
    gen = SyncStepRolloutGenerator(StubPolicy(0.5, cuda), env, n_steps=c["n_steps"], subaction_mask=GATES,
                                   cuda_graph=cuda_graph, **c["kw"])
    for r in range(3):
        ro = gen.rollout(gamma=c["gamma"], gae_lambda=c["lam"])
        g = lambda k: z[f"{name}.r{r}.{k}"]
        b = ro.batch()
        np.testing.assert_array_equal(b.obs.cpu().numpy(), g("obs"))
        np.testing.assert_array_equal(b.action_masks.cpu().numpy(), g("action_masks"))
        np.testing.assert_array_equal(b.actions.cpu().numpy().astype(np.int64), g("actions"))
        np.testing.assert_array_equal(b.values.cpu().numpy(), g("values"))
        np.testing.assert_array_equal(b.logprobs.cpu().numpy(), g("logprobs"))
        np.testing.assert_array_equal(b.advantages.cpu().numpy(), g("advantages"))
        np.testing.assert_array_equal(b.returns.cpu().numpy(), g("returns"))
        # what the masked reset rewrote for the next rollout
        np.testing.assert_array_equal(gen.next_obs.cpu().numpy(), g("next_obs"))
        np.testing.assert_array_equal(gen.next_action_masks.cpu().numpy(), g("next_action_masks"))
        np.testing.assert_array_equal(gen.next_episode_starts.cpu().numpy(), g("next_episode_starts"))
