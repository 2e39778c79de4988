"""Golden fixtures for the trajectory rollouts from the LIVE, UNMODIFIED reference:

    python tests/golden/make_golden_trajectory_rollouts.py        (build container only)

Runs the reference's GuidedLearnerRolloutGenerator, RandomGuidedLearnerRolloutGenerator (with and without
skip_no_action_steps) and ReferenceAIRolloutGenerator over the scripted env / stub policies of
``tests/traj_stubs.py`` and stores the flat Batch each one hands to the learner (row order included) plus the
minibatch index stream.  The GPU tests replay the same script through this repo's device generators.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _ref_shim  # noqa: E402

_ref_shim.install()

from rl_algo_impls.rollout.guided_learner_rollout import GuidedLearnerRolloutGenerator as RefGuided  # noqa: E402
from rl_algo_impls.rollout.random_guided_learner_rollout import (  # noqa: E402
    RandomGuidedLearnerRolloutGenerator as RefRandomGuided,
)
from rl_algo_impls.rollout.reference_ai_rollout import ReferenceAIRolloutGenerator as RefReferenceAI  # noqa: E402

from tests.traj_stubs import GATES, ScriptedVecEnv, StubPolicy  # noqa: E402

CASES = {
    "guided": dict(cls="guided", N=6, n_steps=8, env_seed=5, np_seed=11, gamma=0.97, lam=0.9, kw=dict(switch_range=4)),
    "random_guided": dict(cls="random", N=6, n_steps=8, env_seed=6, np_seed=12, gamma=0.97, lam=0.9,
                          kw=dict(guide_probability=0.4)),
    "random_guided_skip": dict(cls="random", N=6, n_steps=10, env_seed=7, np_seed=13, gamma=0.95, lam=0.8,
                               kw=dict(guide_probability=0.3, skip_no_action_steps=True, num_envs_reset_every_rollout=2)),
    "reference_ai": dict(cls="ai", N=4, n_steps=7, env_seed=8, np_seed=14, gamma=0.99, lam=0.95, kw=dict(include_logp=False)),
}


class _FourTuple:
    """reference_ai_rollout.py:62-67 still unpacks the pre-gymnasium 4-tuple from vec_env.step."""

    def __init__(self, env):
        self._env = env

    def __getattr__(self, name):
        return getattr(self._env, name)

    def step(self, actions):
        obs, rew, term, trunc, info = self._env.step(actions)
        return obs, rew, term | trunc, info


def spaces():
    return (lambda shape, dtype: _ref_shim.Box(-np.inf, np.inf, shape, dtype)), _ref_shim.MultiDiscrete


def run_case(name, c):
    env = ScriptedVecEnv(c["N"], c["env_seed"], space_factory=spaces())
    learner, guide = StubPolicy(0.5), StubPolicy(-0.25)
    np.random.seed(c["np_seed"])
    if c["cls"] == "guided":
        gen = RefGuided(learner, env, guide, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    elif c["cls"] == "random":
        gen = RefRandomGuided(learner, env, guide, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    else:
        gen = RefReferenceAI(learner, _FourTuple(env), n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    out = {}
    for r in range(2):  # two rollouts: carries the builder / switch state across the rollout boundary
        ro = gen.rollout(gamma=c["gamma"], gae_lambda=c["lam"])
        b = ro.batch if not callable(ro.batch) else ro.batch()
        torch.manual_seed(100 + r)
        idx = [mb_idx for mb_idx in _index_stream(ro, 5)]
        out.update({
            f"r{r}.obs": b.obs.numpy(), f"r{r}.actions": b.actions.numpy(), f"r{r}.action_masks": b.action_masks.numpy(),
            f"r{r}.num_actions": b.num_actions.numpy(), f"r{r}.values": b.values.numpy(),
            f"r{r}.advantages": b.advantages.numpy(), f"r{r}.returns": b.returns.numpy(),
            f"r{r}.y_true": np.asarray(ro.y_true), f"r{r}.y_pred": np.asarray(ro.y_pred),
            f"r{r}.total_steps": np.asarray(ro.total_steps), f"r{r}.num_minibatches": np.asarray(ro.num_minibatches(5)),
            f"r{r}.index_stream": np.concatenate(idx) if idx else np.zeros((0,), np.int64),  # minibatches of 5, concatenated
        })
        if b.logprobs is not None:
            out[f"r{r}.logprobs"] = b.logprobs.numpy()
        out = {k: np.array(v) for k, v in out.items()}  # VecRollout's batch aliases the generator's reused buffers
        print(f"{name} rollout {r}: {ro.total_steps} rows")
    return out


def _index_stream(ro, batch_size):
    """the indices minibatches() draws: same torch.randperm call on the CPU default generator"""
    state = torch.get_rng_state()
    n = ro.total_steps
    perm = torch.randperm(n)
    torch.set_rng_state(state)
    got = [mb.obs.numpy() for mb in ro.minibatches(batch_size)]
    b = ro.batch if not callable(ro.batch) else ro.batch()
    for i, g in enumerate(got):
        assert np.array_equal(g, b.obs.numpy()[perm[i * batch_size:(i + 1) * batch_size].numpy()])
    return [perm[i * batch_size:(i + 1) * batch_size].numpy() for i in range(len(got))]


if __name__ == "__main__":
    out = {}
    for name, c in CASES.items():
        for k, v in run_case(name, c).items():
            out[f"{name}.{k}"] = v
    path = os.path.join(HERE, "trajectory_rollouts.npz")
    np.savez_compressed(path, **out)
    print(f"wrote trajectory_rollouts.npz ({os.path.getsize(path) / 1024:.1f} KB)")
