"""Trajectory rollouts on the device vs the live reference's generators (fixtures: tests/golden/trajectory_rollouts.npz,
made by tests/golden/make_golden_trajectory_rollouts.py from the unmodified reference over tests/traj_stubs.py)."""
import os

import numpy as np
import pytest
import torch

from tests.traj_stubs import GATES, ScriptedVecEnv, StubPolicy

pytestmark = pytest.mark.gpu

CASES = {
    "guided": dict(cls="guided", N=6, n_steps=8, env_seed=5, np_seed=11, gamma=0.97, lam=0.9, kw=dict(switch_range=4)),
    "random_guided": dict(cls="random", N=6, n_steps=8, env_seed=6, np_seed=12, gamma=0.97, lam=0.9,
                          kw=dict(guide_probability=0.4)),
    "random_guided_skip": dict(cls="random", N=6, n_steps=10, env_seed=7, np_seed=13, gamma=0.95, lam=0.8,
                               kw=dict(guide_probability=0.3, skip_no_action_steps=True, num_envs_reset_every_rollout=2)),
    "reference_ai": dict(cls="ai", N=4, n_steps=7, env_seed=8, np_seed=14, gamma=0.99, lam=0.95, kw=dict(include_logp=False)),
}


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectory_rollouts.npz"))


def make_generator(c, cuda):
    from rl_algo_impls_b200.rollout import (GuidedLearnerRolloutGenerator, RandomGuidedLearnerRolloutGenerator,
                                            ReferenceAIRolloutGenerator)

    env = ScriptedVecEnv(c["N"], c["env_seed"])
    learner, guide = StubPolicy(0.5, cuda), StubPolicy(-0.25, cuda)
    np.random.seed(c["np_seed"])
    if c["cls"] == "guided":
        return GuidedLearnerRolloutGenerator(learner, env, guide, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    if c["cls"] == "random":
        return RandomGuidedLearnerRolloutGenerator(learner, env, guide, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])
    return ReferenceAIRolloutGenerator(learner, env, n_steps=c["n_steps"], subaction_mask=GATES, **c["kw"])


@pytest.mark.parametrize("name", list(CASES))
def test_rollout_matches_the_reference_generator(cuda, golden, name):
    """Same rows in the same order, bit-exact observations / actions / masks / values / num_actions and minibatch
    index stream; advantages / returns bit-exact for the standard GAE, 1e-6 for the gamma ** steps_elapsed one."""
    c = CASES[name]
    gen = make_generator(c, cuda)
    for r in range(2):
        ro = gen.rollout(gamma=c["gamma"], gae_lambda=c["lam"])
        g = lambda k: golden[f"{name}.r{r}.{k}"]
        b = ro.batch() if callable(ro.batch) else ro.batch
        assert ro.total_steps == int(g("total_steps")) and ro.num_minibatches(5) == int(g("num_minibatches"))
        np.testing.assert_array_equal(b.obs.cpu().numpy(), g("obs"))
        np.testing.assert_array_equal(b.action_masks.cpu().numpy(), g("action_masks"))
        np.testing.assert_array_equal(b.values.cpu().numpy(), g("values"))
        np.testing.assert_array_equal(b.actions.cpu().numpy().astype(np.int64), g("actions"))
        np.testing.assert_array_equal(b.num_actions.cpu().numpy(), g("num_actions"))
        if b.logprobs is not None and f"{name}.r{r}.logprobs" in golden:
            np.testing.assert_array_equal(b.logprobs.cpu().numpy(), g("logprobs"))
        if c["cls"] == "random":  # float64 pow on the device vs numpy: 1e-6
            np.testing.assert_allclose(b.advantages.cpu().numpy(), g("advantages"), rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(ro.y_true, g("y_true"), rtol=1e-6, atol=1e-6)
        else:
            np.testing.assert_array_equal(b.advantages.cpu().numpy(), g("advantages"))
            np.testing.assert_array_equal(b.returns.cpu().numpy(), g("returns"))
            np.testing.assert_array_equal(ro.y_true, g("y_true"))
        np.testing.assert_array_equal(ro.y_pred, g("y_pred"))
        torch.manual_seed(100 + r)
        idx = torch.cat([i.cpu() for i in ro.minibatch_indices(5)]).numpy()
        np.testing.assert_array_equal(idx, g("index_stream"))
        torch.manual_seed(100 + r)
        got = torch.cat([mb.obs for mb in ro.minibatches(5)]).cpu().numpy()
        np.testing.assert_array_equal(got, g("obs")[g("index_stream")])


def test_builders_accept_explicit_rows(cuda):
    """The reference's add() / step_add() call signature with explicit rows (numpy or tensors): stacked on the
    device, one-segment K1b scan on demand, same numbers as a store-backed builder fed the same steps."""
    from rl_algo_impls_b200.rollout import DiscreteSkipsTrajectoryBuilder, TrajectoryBuilder, TrajectoryRollout
    from tests.test_oracle_golden import load

    z = load("trajectories")
    off = z["scalar.offsets"]
    trajs = []
    for i in range(len(off) - 1):
        tb = TrajectoryBuilder(device=cuda)
        lo, hi = int(off[i]), int(off[i + 1])
        starts = z["scalar.starts"][lo:hi]
        dones = np.concatenate([starts[1:], [z["scalar.next_starts"][i]]])
        for t in range(lo, hi):
            tb.add(np.full(2, t, np.float32), z["scalar.rewards"][t], bool(dones[t - lo]), z["scalar.values"][t], 0.0,
                   np.zeros((1, 1), np.int64), None)
        traj = tb.trajectory(float(z["scalar.gamma"]), float(z["scalar.gae_lambda"]),
                             next_values=torch.tensor(z["scalar.next_values"][i], device=cuda))
        np.testing.assert_array_equal(traj.advantages.cpu().numpy(), z["scalar.adv"][lo:hi])
        trajs.append(traj)
    ro = TrajectoryRollout(cuda, trajs)
    np.testing.assert_array_equal(ro.batch.advantages.cpu().numpy(), z["scalar.adv"])
    np.testing.assert_array_equal(ro.batch.obs.cpu().numpy()[:, 0], np.arange(off[-1], dtype=np.float32))
    assert ro.batch.num_actions is None and ro.batch.action_masks is None

    sb = DiscreteSkipsTrajectoryBuilder(device=cuda)
    sb.step_add(np.zeros(2, np.float32), np.float32(1.0), False, np.float32(0.5), 0.0, np.zeros(1, np.int64), None, 0.9)
    sb.step_no_add(np.float32(2.0), False, 0.9)
    sb.step_no_add(np.float32(4.0), True, 0.9)
    assert sb.steps_elapsed == [3] and sb.done
    want = np.float32(0) + np.float32(1.0) * 0.9 ** 0
    want += np.float32(2.0) * 0.9 ** 1
    want += np.float32(4.0) * 0.9 ** 2
    assert np.float32(sb.rewards[0]) == np.float32(want)
    traj = sb.trajectory(0.9, 0.8)
    np.testing.assert_allclose(traj.advantages.cpu().numpy(), [want - 0.5], rtol=1e-6)


def test_no_cpu_path():
    from rl_algo_impls_b200.rollout import StepStore, TrajectoryRollout

    with pytest.raises(RuntimeError):
        StepStore(torch.device("cpu"), 4)
    with pytest.raises(RuntimeError):
        TrajectoryRollout(torch.device("cpu"), [])
