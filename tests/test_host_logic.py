"""Host-side logic that needs no GPU: the Batch contract, spaces, synthetic envs, configs, the lane
plan of the GridNet kernel's host mirror, TrainStats, and the loud failure without a CUDA device."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from rl_algo_impls_b200 import ops, spaces
from rl_algo_impls_b200.configs import CONFIGS
from rl_algo_impls_b200.envs import SPECS, make_synthetic_env
from rl_algo_impls_b200.policy import ActorCritic, clamp_actions
from rl_algo_impls_b200.ppo.ppo import PPO, TrainStats, TrainStepStats, num_or_array
from rl_algo_impls_b200.rollout import Batch, VecRollout
from tests.test_oracle_golden import load


def test_batch_field_order_is_the_references():
    # consumers unpack with dataclasses.astuple (ppo/ppo.py:295-305): the order is part of the contract
    names = [f.name for f in dataclasses.fields(Batch)]
    assert names == ["obs", "logprobs", "actions", "action_masks", "num_actions", "values", "advantages", "returns",
                     "additional"]


def test_clamp_actions_is_the_references():
    """The reference's one unit test (tests/shared/policy/test_actor_critic.py:8-17), against the
    values the live reference produced (tests/golden/index_stream.npz)."""
    z = load("index_stream")
    got = clamp_actions(np.array([-1.5, 0, 1.5]), spaces.Box(-1, 1, (1,)), squash_output=False)
    np.testing.assert_array_equal(got, z["clamp_noscale"])
    np.testing.assert_array_equal(got, np.array([-1, 0, 1]))
    got = clamp_actions(np.array([-1, 0, 1]), spaces.Box(-3, 2, (1,)), squash_output=True)
    np.testing.assert_array_equal(got, z["clamp_squash"])
    np.testing.assert_array_equal(got, np.array([-3, -0.5, 2]))


def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        VecRollout(torch.device("cpu"), np.zeros(2, bool), np.zeros(2, np.float32), np.zeros((3, 2, 4), np.float32),
                   np.zeros((3, 2), np.int64), np.zeros((3, 2), np.float32), np.zeros((3, 2), bool),
                   np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32), None, 0.99, 0.95)
    with pytest.raises(Exception, match="CUDA"):
        ops.gae_scan(torch.zeros(3, 2), torch.zeros(3, 2), torch.zeros(3, 2, dtype=torch.bool),
                     torch.zeros(2, dtype=torch.bool), torch.zeros(2), 0.99, 0.95)


def test_gridnet_spec_from_subaction_mask():
    spec = ops.GridnetSpec.from_subaction_mask((6, 4, 4, 4, 4, 7, 49), {0: {1: 1, 2: 2, 3: 3, 4: 4, 5: 4, 6: 5}})
    assert spec.nvec == (6, 4, 4, 4, 4, 7, 49) and spec.n_pick == 0
    assert spec.gates == ((1, 0, 1), (2, 0, 2), (3, 0, 3), (4, 0, 4), (5, 0, 4), (6, 0, 5))
    lux = ops.GridnetSpec.from_subaction_mask((4, 6, 4, 4, 5, 5), {1: {2: 0, 3: 1, 4: 1, 5: 2}}, n_pick=1)
    assert lux.gates == ((2, 1, 0), (3, 1, 1), (4, 1, 1), (5, 1, 2)) and lux.n_pick == 1


@pytest.mark.parametrize("name", list(SPECS))
def test_synthetic_env_contract(name):
    env = make_synthetic_env(name, 4, seed=3, pool=2)
    obs, info = env.reset()
    assert obs.shape == (4,) + SPECS[name].obs_shape and isinstance(info, dict)
    nobs, rew, term, trunc, _ = env.step(None)
    V = SPECS[name].n_values
    assert rew.shape == ((4,) if V == 1 else (4, V)) and rew.dtype == np.float32
    assert term.dtype == np.bool_ and trunc.dtype == np.bool_ and term.shape == (4,)
    mask = env.get_action_mask()
    if SPECS[name].kind == "gridnet":
        cells = mask["per_position"] if isinstance(mask, dict) else mask
        hw = SPECS[name].map_hw[0] * SPECS[name].map_hw[1]
        assert cells.shape == (4, hw, sum(SPECS[name].nvec)) and cells.dtype == np.bool_
        # a cell is either empty or has at least one valid entry in every head
        start = 0
        any_cell = cells.any(-1)
        for n in SPECS[name].nvec:
            assert (cells[..., start:start + n].any(-1) == any_cell).all()
            start += n
    else:
        assert mask is None
    # same seed -> same data
    env2 = make_synthetic_env(name, 4, seed=3, pool=2)
    np.testing.assert_array_equal(env2.reset()[0], obs)


@pytest.mark.parametrize("key", list(CONFIGS))
def test_policies_match_the_reference_parameter_counts(key):
    """Trunks are ours, but built to the reference architectures: the parameter counts of the
    instantiated reference models are the check."""
    cfg = CONFIGS[key]
    env = make_synthetic_env(cfg.env, 2, pool=1)
    policy = ActorCritic(env, **cfg.policy)
    n = sum(p.numel() for p in policy.parameters())
    # C5: the reference's squeeze_unet at the 64x64 Lux YAML entry (squeeze_unet.py:20/:198)
    want = {"C1": 9155, "C2": 1686693, "C3": 142605, "C4": 851727, "C5": 4719274}[key]
    assert n == want, (key, n)
    assert policy.value_shape == (() if SPECS[cfg.env].n_values == 1 else (SPECS[cfg.env].n_values,))
    if key == "C5":
        assert policy.action_shape == {"per_position": (4096, 6), "pick_position": (1,)}
    if key == "C4":
        assert policy.action_shape == (256, 7)


def test_train_stats_aggregation_matches_the_reference_rules():
    steps = [TrainStepStats(1.0, 0.5, np.array([1.0, 2.0]), -0.1, 0.01, 0.2, np.array([0.0, 0.5]), {}),
             TrainStepStats(3.0, 1.5, np.array([3.0, 4.0]), -0.3, 0.03, 0.4, np.array([1.0, 0.5]), {})]
    s = TrainStats(steps, explained_var=0.25, grad_norms=[1.0, 3.0])
    assert s.loss == 2.0 and s.pi_loss == 1.0 and s.approx_kl == pytest.approx(0.02) and s.grad_norm == 2.0
    np.testing.assert_allclose(s.v_loss, [2.0, 3.0])
    np.testing.assert_allclose(s.val_clipped_frac, [0.5, 0.5])


def test_ppo_constructor_keeps_the_reference_attribute_names():
    env = make_synthetic_env("CartPole-v1", 2, pool=1)
    algo = PPO(ActorCritic(env), torch.device("cpu"), None, gamma=[1.0, 0.99], gae_lambda=0.9, vf_coef=[0.5, 0.25])
    # callbacks setattr these between epochs (hyperparam_transitions.py:19-43)
    for name in ("learning_rate", "clip_range", "clip_range_vf", "ent_coef", "gamma", "gae_lambda", "vf_coef",
                 "multi_reward_weights", "switch_range", "guide_probability", "teacher_kl_loss_coef",
                 "freeze_policy_head", "freeze_value_head", "freeze_backbone", "batch_size", "n_epochs", "max_grad_norm"):
        assert hasattr(algo, name), name
    assert isinstance(algo.gamma, np.ndarray) and algo.gamma.dtype == np.float64 and algo.gae_lambda == 0.9
    assert num_or_array(0.5) == 0.5
    assert algo.optimizer.defaults["eps"] == 1e-7  # ppo.py:146
    with pytest.raises(AssertionError):
        PPO(ActorCritic(env), torch.device("cpu"), None, normalize_advantage=True, standardize_advantage=True)


def test_discrete_skips_builder_host_bookkeeping():
    """rollout/discrete_skips_trajectory_builder.py:30-62: skipped steps fold reward * gamma ** steps_elapsed into
    the last kept step in float32 (numpy's scalar arithmetic, as the reference) -- host logic, no device needed."""
    from rl_algo_impls_b200.rollout.trajectory import DiscreteSkipsTrajectoryBuilder

    rng = np.random.default_rng(3)
    b = DiscreteSkipsTrajectoryBuilder(device="cpu")  # explicit-row mode never touches the device before trajectory()
    gamma = 0.97
    want_rewards, want_steps = [], []
    for t in range(40):
        r = np.float32(rng.standard_normal())
        if t == 0 or rng.random() < 0.4:
            b.step_add(np.zeros(2, np.float32), r, False, np.float32(0), 0.0, np.zeros(1, np.int64), None, gamma)
            want_rewards.append(np.zeros_like(r)), want_steps.append(0)
        else:
            b.step_no_add(r, False, gamma)
        want_rewards[-1] += r * gamma ** want_steps[-1]
        want_steps[-1] += 1
    assert b.steps_elapsed == want_steps and len(b) == len(want_steps)
    assert all(np.float32(x) == np.float32(y) and np.asarray(x).dtype == np.float32 for x, y in zip(b.rewards, want_rewards))
    b.step_no_add(np.float32(1), True, gamma)
    with pytest.raises(AssertionError):
        b.step_no_add(np.float32(1), False, gamma)


def test_guided_rollout_host_helpers():
    from rl_algo_impls_b200.rollout.guided_learner_rollout import has_actions, rearrange

    assert rearrange(["c", "a", "b"], [2, 0, 1]) == ["a", "b", "c"]
    m = np.zeros((3, 4, 5), np.bool_)
    m[1, 2, 3] = True
    assert has_actions(m).tolist() == [False, True, False]
    assert has_actions({"per_position": m, "pick_position": np.zeros((3, 1, 4), np.bool_)}).tolist() == [False, True, False]


def test_trajectory_rollout_fixture_is_present():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectory_rollouts.npz"))
    for case in ("guided", "random_guided", "random_guided_skip", "reference_ai"):
        assert int(z[f"{case}.r0.total_steps"]) == z[f"{case}.r0.obs"].shape[0] > 0


@pytest.mark.parametrize("key", ["C1", "C2", "C3", "C4", "C5"])
def test_freeze_and_unfreeze_parameter_groups(key):
    """ActorCritic.freeze (actor_critic.py:384-395): heads and backbone toggle independently; unfreeze restores all."""
    cfg = CONFIGS[key]
    env = make_synthetic_env(cfg.env, 2, seed=0)
    policy = ActorCritic(env, **cfg.policy)
    total = sum(p.numel() for p in policy.parameters())
    count = lambda: sum(p.numel() for p in policy.parameters() if p.requires_grad)
    policy.freeze(True, False, freeze_backbone=False)
    without_policy_head = count()
    policy.freeze(False, True, freeze_backbone=False)
    without_value_head = count()
    policy.freeze(False, False, freeze_backbone=True)
    heads_only = count()
    assert 0 < heads_only < total and without_policy_head < total and without_value_head < total
    assert (total - without_policy_head) + (total - without_value_head) == heads_only
    policy.freeze(True, True, freeze_backbone=True)
    assert count() == 0
    policy.unfreeze()
    assert count() == total


def test_uploader_page_lock_policy(monkeypatch):
    """_Uploader._page_locked: a host buffer is page-locked in place on its SECOND sighting, by its owning array
    (views of one pool share a registration), a refusal is remembered, and close() releases what was locked."""
    from rl_algo_impls_b200 import _lib
    from rl_algo_impls_b200.rollout.sync_step_rollout import _Uploader

    calls = {"register": [], "unregister": []}

    class FakeLib:
        refuse = set()

        def b200rl_host_register(self, ptr, nbytes):
            calls["register"].append((ptr, nbytes))
            return -3 if ptr in self.refuse else 0

        def b200rl_host_unregister(self, ptr):
            calls["unregister"].append(ptr)
            return 0

    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    up = _Uploader(torch.device("cpu"))
    pool = np.zeros((4, 1 << 16), np.float32)
    assert not up._page_locked(pool[0])          # first sighting: not yet
    assert up._page_locked(pool[1])              # second sighting of the same owning array (another slot)
    assert up._page_locked(pool[2]) and len(calls["register"]) == 1
    assert calls["register"][0] == (pool.ctypes.data, pool.nbytes)
    other = np.ones(1 << 18, np.float32)
    fake.refuse.add(other.ctypes.data)
    assert not up._page_locked(other) and not up._page_locked(other) and not up._page_locked(other)
    assert len(calls["register"]) == 2           # the refusal is remembered, not retried
    borrowed = np.frombuffer(bytearray(1 << 20), dtype=np.uint8)  # memory owned by something that is not an ndarray
    assert not up._page_locked(borrowed) and not up._page_locked(borrowed)
    up.close()
    assert calls["unregister"] == [pool.ctypes.data]
    up.close()
    assert calls["unregister"] == [pool.ctypes.data]


def test_moving_normaliser_checkpoints_round_trip(tmp_path):
    """ExponentialMovingMeanVar / HybridMovingMeanVar save -> load, per-env mode included: the reference writes the
    arrays it holds (utils/running_mean_std.py:98-112) -- shape () before the first update and for vector rewards,
    (N,) once scalar rewards went through its broadcasting update -- and load must take either."""
    from rl_algo_impls_b200.wrappers.normalize import ExponentialMovingMeanVar, HybridMovingMeanVar, RunningMeanStd

    cpu = torch.device("cpu")
    h = HybridMovingMeanVar(cpu, window_size=5.0, shape=(), per_env=4)
    h.emmv.mean.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64))
    h.emmv.var.copy_(torch.tensor([0.5, 0.25, 2.0, 1.0], dtype=torch.float64))
    h.emmv._initialized.fill_(1)
    h.rms.mean.fill_(0.75), h.rms.var.fill_(1.5), h.rms._count.fill_(12.0)
    h.save(str(tmp_path / "norm_reward.npz"))
    z = np.load(str(tmp_path / "norm_reward.npz-emmv.npz"))
    assert z["mean"].shape == (4,) and bool(z["initialized"])  # the (N,) arrays the reference holds after an update
    h2 = HybridMovingMeanVar(cpu, window_size=5.0, shape=(), per_env=4)
    h2.load(str(tmp_path / "norm_reward.npz"))
    assert torch.equal(h2.emmv.mean, h.emmv.mean) and torch.equal(h2.emmv.var, h.emmv.var) and h2.emmv.initialized
    assert torch.equal(h2.emmv.squared_mean, h.emmv.var + h.emmv.mean ** 2)
    assert h2.rms.count == 12.0 and torch.equal(h2.rms.mean, h.rms.mean)
    # a file the reference wrote BEFORE its first update: shape-() arrays, not initialised
    np.savez_compressed(str(tmp_path / "fresh.npz"), mean=np.zeros(()), var=np.ones(()), initialized=False)
    e = ExponentialMovingMeanVar(cpu, window_size=5.0, shape=(), per_env=4)
    e.load(str(tmp_path / "fresh.npz"))
    assert not e.initialized and (e.var == 1).all() and e.mean.numel() == 4
    e.save(str(tmp_path / "fresh_out.npz"))
    assert np.load(str(tmp_path / "fresh_out.npz"))["mean"].shape == ()
    # vector rewards: plain [V] state
    v = ExponentialMovingMeanVar(cpu, alpha=0.1, shape=(3,))
    v.mean.copy_(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)), v._initialized.fill_(1)
    v.save(str(tmp_path / "vec.npz"))
    v2 = ExponentialMovingMeanVar(cpu, alpha=0.1, shape=(3,))
    v2.load(str(tmp_path / "vec.npz"))
    assert torch.equal(v2.mean, v.mean) and np.load(str(tmp_path / "vec.npz"))["mean"].shape == (3,)
    with pytest.raises(ValueError):
        ExponentialMovingMeanVar(cpu, alpha=0.1, shape=(2,)).load(str(tmp_path / "vec.npz"))
    r = RunningMeanStd(cpu, shape=(3,))
    r.mean.copy_(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64))
    r.save(str(tmp_path / "rms.npz"))
    r2 = RunningMeanStd(cpu, shape=(3,))
    r2.load(str(tmp_path / "rms.npz"))
    assert torch.equal(r2.mean, r.mean)


def test_optimizer_state_written_by_the_reference_loads(tmp_path):
    """Algorithm.load (shared/algorithm.py:48-60): a reference checkpoint carries capturable=False, a Python-float
    learning rate and CPU step counters; loading it must not rebind the learning-rate object this optimizer was built
    with (captured update graphs read it) nor change its capturable flag, and must drop captured update graphs."""
    env = make_synthetic_env("CartPole-v1", 2, pool=1)
    policy = ActorCritic(env)
    algo = PPO(policy, torch.device("cpu"), None, learning_rate=3e-4)
    ref_opt = torch.optim.Adam(policy.parameters(), lr=7e-4, eps=1e-7)  # ppo.py:146
    for p in policy.parameters():
        p.grad = torch.ones_like(p)
    ref_opt.step()
    torch.save(ref_opt.state_dict(), tmp_path / "optimizer.pt")
    group = algo.optimizer.param_groups[0]
    group["lr"] = torch.tensor(3e-4)  # what PPO builds on a CUDA device: a tensor learning rate, capturable
    group["capturable"] = True
    lr_obj = group["lr"]
    algo._update_graphs["stale"] = object()
    algo.load(str(tmp_path))
    group = algo.optimizer.param_groups[0]
    assert group["lr"] is lr_obj and abs(float(lr_obj) - 7e-4) < 1e-9 and group["capturable"] is True
    assert not algo._update_graphs
    for p in policy.parameters():
        st = algo.optimizer.state[p]
        assert isinstance(st["step"], torch.Tensor) and st["step"].dtype == torch.float32 and float(st["step"]) == 1.0
        assert torch.equal(st["exp_avg"], ref_opt.state[p]["exp_avg"])


def test_reference_copy_matches_its_manifest():
    """oracle/_ref (made by oracle/make_ref.sh, git-ignored, shipped to the GPU box) is byte-identical to the
    reference sources the tracked manifest pins."""
    import hashlib

    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    root = os.path.join(here, "_ref")
    if not os.path.isdir(root):
        pytest.skip("oracle/_ref has not been made here (oracle/make_ref.sh)")
    lines = open(os.path.join(here, "ref_manifest.sha256")).read().split("\n")
    entries = [line.split(None, 1) for line in lines if line.strip()]
    assert len(entries) > 100
    for digest, rel in entries:
        with open(os.path.join(root, rel.strip()), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel


def test_no_gc_during_capture_restores_the_collector():
    """The capture guard collects first, keeps the cyclic collector off inside, and restores its previous state --
    also when it was off to begin with, and when the body raises."""
    import gc

    from rl_algo_impls_b200 import ops

    assert gc.isenabled()
    with ops.no_gc_during_capture():
        assert not gc.isenabled()
    assert gc.isenabled()
    gc.disable()
    try:
        with ops.no_gc_during_capture():
            assert not gc.isenabled()
        assert not gc.isenabled()
    finally:
        gc.enable()
    try:
        with ops.no_gc_during_capture():
            raise RuntimeError("boom")
    except RuntimeError:
        pass
    assert gc.isenabled()
