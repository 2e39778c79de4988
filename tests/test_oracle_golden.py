"""The oracle against the golden fixtures generated from the live reference
(tests/golden/make_golden.py).  CPU only: this is what pins the checker."""
import os

import numpy as np
import pytest
import torch

from oracle import learner as olearn
from oracle.distributions import Gridnet, MaskedLogits, gaussian_logp_entropy
from oracle.gae import gae_advantages, gae_returns
from oracle.rollout import minibatch_index_stream
from tests.golden.stub_nets import TinyGrid, TinyMlp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def cases(z):
    return sorted({k.split(".")[0] for k in z})


def gates_of(arr):
    return {int(h): (int(r), int(v)) for h, r, v in arr}


@pytest.mark.parametrize("name", cases(load("gae")))
def test_gae_matches_reference(name):
    z = load("gae")
    g = lambda k: z[f"{name}.{k}"]
    gamma = float(g("gamma")) if bool(g("gamma_is_scalar")) else g("gamma")
    lam = float(g("gae_lambda")) if bool(g("gamma_is_scalar")) else g("gae_lambda")
    adv = gae_advantages(g("rewards"), g("values"), g("episode_starts"), g("next_episode_starts"), g("next_values"),
                         gamma, lam)
    np.testing.assert_array_equal(adv, g("advantages"))
    np.testing.assert_array_equal(gae_returns(adv, g("values")), g("returns"))


@pytest.mark.parametrize("name", cases(load("gridnet")))
def test_gridnet_matches_reference(name):
    z = load("gridnet")
    g = lambda k: z.get(f"{name}.{k}")
    t = lambda k: None if g(k) is None else torch.from_numpy(g(k))
    masks = t("mask") if g("pick_mask") is None else {"per_position": t("mask"), "pick_position": t("pick_mask")}
    action = t("actions") if g("pick_actions") is None else {"per_position": t("actions"), "pick_position": t("pick_actions")}
    logits = t("logits").requires_grad_(True)
    d = Gridnet(logits.shape[1], g("nvec").tolist(), logits, masks, gates_of(g("gates")))
    logp, ent = d.log_prob(action), d.entropy()
    (logp * t("dlogp") + ent * t("dentropy")).sum().backward()
    np.testing.assert_array_equal(logp.detach().numpy(), g("logp"))
    np.testing.assert_array_equal(ent.detach().numpy(), g("entropy"))
    np.testing.assert_array_equal(logits.grad.numpy(), g("dlogits"))


def test_fully_masked_rows_are_exact_zeros():
    z = load("gridnet")
    assert (z["all_masked.dlogits"][..., :-1] == 0).all()
    # only the pick_position head (last logit) can carry anything when no cell has a unit
    assert np.isfinite(z["all_masked.logp"]).all()


@pytest.mark.parametrize("name", ["cartpole", "atari", "masked"])
def test_categorical_matches_reference(name):
    z = load("heads")
    g = lambda k: z.get(f"{name}.{k}")
    logits = torch.from_numpy(g("logits")).requires_grad_(True)
    mask = torch.from_numpy(g("mask")) if g("mask") is not None else None
    d = MaskedLogits(logits, mask)
    a = torch.from_numpy(g("actions"))
    logp, ent = d.log_prob(a), d.entropy()
    (logp * torch.from_numpy(g("dlogp")) + ent * torch.from_numpy(g("dentropy"))).sum().backward()
    np.testing.assert_array_equal(logp.detach().numpy(), g("logp"))
    np.testing.assert_array_equal(ent.detach().numpy(), g("entropy"))
    np.testing.assert_allclose(logits.grad.numpy(), g("dlogits"), rtol=1e-6, atol=1e-7)


def test_gaussian_matches_reference():
    z = load("heads")
    g = lambda k: torch.from_numpy(z[f"gaussian.{k}"])
    mu, ls = g("mu").requires_grad_(True), g("log_std").requires_grad_(True)
    logp, ent = gaussian_logp_entropy(mu, ls, g("actions"))
    np.testing.assert_array_equal(logp.detach().numpy(), z["gaussian.logp"])
    np.testing.assert_array_equal(ent.detach().numpy(), z["gaussian.entropy"])
    ((logp * g("dlogp")).sum() + (ent * g("dentropy")).sum()).backward()
    np.testing.assert_allclose(mu.grad.numpy(), z["gaussian.dmu"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ls.grad.numpy(), z["gaussian.dlog_std"], rtol=1e-5, atol=1e-6)


def test_minibatch_index_stream_is_the_references():
    z = load("index_stream")
    torch.manual_seed(int(z["seed"]))
    stream = minibatch_index_stream(int(z["total"]), int(z["batch_size"]), shuffle=True)
    np.testing.assert_array_equal(np.concatenate([i.numpy() for i in stream]), z["shuffled"])
    np.testing.assert_array_equal([len(i) for i in stream], z["sizes"])
    seq = minibatch_index_stream(int(z["total"]), int(z["batch_size"]), shuffle=False)
    np.testing.assert_array_equal(np.concatenate([i.numpy() for i in seq]), z["sequential"])


def rollout_from(z):
    ro = {}
    for k, v in z.items():
        if not k.startswith("ro."):
            continue
        parts = k.split(".")[1:]
        if len(parts) == 2:
            ro.setdefault(parts[0], {})[parts[1]] = v
        else:
            ro[parts[0]] = v
    ro.setdefault("masks", None)
    return ro


def learner_setup(name):
    from tests.golden.make_golden_cases import LEARNER_CASES, make_net_for

    case = LEARNER_CASES[name]
    z = load("learner_" + name)
    net = make_net_for(case)()
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("init.")})
    return case, z, net


LEARNER_CASES_ALL = ["cartpole", "gaussian", "microrts", "lux", "microrts_teacher", "cartpole_teacher_biased",
                     "cartpole_huber", "gaussian_l1", "microrts_kl_cutoff", "cartpole_standardize", "lux_after_scaling",
                     "lux_vf_weights", "microrts_autocast"]


class FirstGradients:
    """Keeps a copy of every parameter's gradient at the first clip (the reference clips first in optimizer_step,
    ppo.py:441-443): the gradient of the first minibatch -- of the first epoch under gradient accumulation -- before
    clipping and before Adam.  `fn` is called as the wrapped callable's replacement."""

    def __init__(self, named_parameters):
        self.named, self.grads = list(named_parameters), {}

    def capture(self):
        if not self.grads:
            self.grads = {n: p.grad.detach().clone() for n, p in self.named if p.grad is not None}

    def __enter__(self):
        self._orig = torch.nn.utils.clip_grad_norm_

        def wrapped(parameters, *a, **k):
            self.capture()
            return self._orig(parameters, *a, **k)

        torch.nn.utils.clip_grad_norm_ = wrapped
        return self

    def __exit__(self, *exc):
        torch.nn.utils.clip_grad_norm_ = self._orig
        return False


def teacher_net(case, z):
    """The teacher checkpoint of a teacher-KL case (None otherwise)."""
    from tests.golden.make_golden_cases import make_net_for

    keys = [k for k in z if k.startswith("teacher.")]
    if not keys:
        return None
    net = make_net_for(case)()
    net.load_state_dict({k[8:]: torch.from_numpy(z[k]) for k in keys})
    return net


@pytest.mark.parametrize("name", LEARNER_CASES_ALL)
def test_oracle_learn_epoch_reproduces_the_reference(name):
    """Same initial weights + rollout + randperm seed -> the reference PPO.learn_epoch's final
    parameters and TrainStats, bit for bit."""
    case, z, net = learner_setup(name)
    hp = case["hp"]
    pol = olearn.OraclePolicy(net, case["kind"], case["nvec"], case.get("side", 0) ** 2, case.get("gates"))
    opt = torch.optim.Adam(net.parameters(), lr=hp.learning_rate, eps=1e-7)
    tnet = teacher_net(case, z)  # before seeding: building a module draws from the generator
    teacher = olearn.OraclePolicy(tnet, case["kind"], case["nvec"], case.get("side", 0) ** 2, case.get("gates")) if tnet else None
    torch.manual_seed(int(z["seed"]) + 100)
    with FirstGradients(net.named_parameters()) as first:
        stats = olearn.learn_epoch(pol, opt, rollout_from(z), hp, teacher=teacher)
    if teacher is not None:
        assert np.float64(stats["teacher_kl_loss"]) == z["stats.teacher_kl_loss"]
    for k, g in first.grads.items():  # the pre-Adam gradient of the first minibatch, bit for bit
        np.testing.assert_array_equal(g.numpy(), z[f"grad0.{k}"], err_msg=f"first gradient {k}")
    for k, v in net.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), z[f"final.{k}"], err_msg=k)
    for k in ("loss", "pi_loss", "entropy_loss", "approx_kl", "clipped_frac", "grad_norm", "explained_var"):
        assert np.float64(stats[k]) == z[f"stats.{k}"], k
    np.testing.assert_array_equal(np.asarray(stats["v_loss"], np.float64), z["stats.v_loss"])


def test_normalizers_match_the_reference():
    """oracle/normalize.py against NormalizeObservation / NormalizeReward of the live reference."""
    from oracle.normalize import ObsNormalizer, RewardNormalizer

    z = load("normalizers")
    o = ObsNormalizer(z["obs"].shape[2:])
    for t in range(z["obs"].shape[0]):
        np.testing.assert_array_equal(o.normalize(z["obs"][t]), z["obs_out"][t])
    np.testing.assert_array_equal(o.rms.var, z["obs_var"])
    for tag in ("scalar", "multi"):
        rew, dones = z[f"rew_{tag}"], z[f"dones_{tag}"]
        r = RewardNormalizer(rew.shape[1], rew.shape[2:], gamma=0.98)
        for t in range(rew.shape[0]):
            np.testing.assert_array_equal(r.step(rew[t], dones[t]), z[f"rew_{tag}_out"][t])
        np.testing.assert_array_equal(r.returns, z[f"rew_{tag}_returns"])


def a2c_setup(name):
    from tests.golden.make_golden_cases import A2C_CASES, make_net_for

    case = A2C_CASES[name]
    z = load("a2c_" + name)
    net = make_net_for(case)()
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("init.")})
    return case, z, net


@pytest.mark.parametrize("name", ["cartpole", "microrts"])
def test_oracle_a2c_iteration_reproduces_the_reference(name):
    case, z, net = a2c_setup(name)
    hp = case["hp"]
    pol = olearn.OraclePolicy(net, case["kind"], case["nvec"], case.get("side", 0) ** 2, case.get("gates"))
    opt = (torch.optim.RMSprop(net.parameters(), lr=hp.learning_rate, eps=hp.rms_prop_eps) if hp.use_rms_prop
           else torch.optim.Adam(net.parameters(), lr=hp.learning_rate))
    torch.manual_seed(int(z["seed"]) + 100)
    stats = olearn.a2c_learn_iteration(pol, opt, rollout_from(z), hp)
    for k, v in net.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), z[f"final.{k}"], err_msg=k)
    for k in ("loss", "pi_loss", "entropy_loss", "explained_var"):
        assert np.float64(stats[k]) == z[f"stats.{k}"], k


def test_oracle_acbc_iteration_reproduces_the_reference():
    """behaviour cloning (acbc/acbc.py:75-141): final parameters of the live reference, bit for bit."""
    from tests.golden.make_golden_cases import A2C_CASES, make_net_for

    case, z = A2C_CASES["microrts"], load("acbc_microrts")
    net = make_net_for(case)()
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("init.")})
    pol = olearn.OraclePolicy(net, "gridnet", case["nvec"], case["side"] ** 2, case["gates"])
    opt = torch.optim.Adam(net.parameters(), lr=float(z["hp.learning_rate"]))
    torch.manual_seed(int(z["seed"]) + 100)
    stats = olearn.acbc_learn_iteration(pol, opt, rollout_from(z), int(z["hp.batch_size"]), int(z["hp.n_epochs"]),
                                        float(z["hp.gamma"]), float(z["hp.gae_lambda"]), float(z["hp.vf_coef"]))
    for k, v in net.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), z[f"final.{k}"], err_msg=k)
    for k in ("loss", "pi_loss"):
        assert np.float64(stats[k]) == z[f"stats.{k}"], k


@pytest.mark.parametrize("tag", ["scalar", "heads"])
def test_trajectory_gae_matches_the_reference(tag):
    """oracle GAE over ragged trajectories (TrajectoryBuilder / DiscreteSkipsTrajectoryBuilder of the reference)."""
    from oracle.gae import discrete_skips_advantages

    z = load("trajectories")
    g = lambda k: z[f"{tag}.{k}"]
    gamma, lam = (float(g("gamma")), float(g("gae_lambda"))) if g("gamma").ndim == 0 else (g("gamma"), g("gae_lambda"))
    off = g("offsets")
    for s in range(len(off) - 1):
        sl = slice(off[s], off[s + 1])
        adv = gae_advantages(g("rewards")[sl], g("values")[sl], g("starts")[sl], np.array(g("next_starts")[s]),
                             g("next_values")[s], gamma, lam)
        np.testing.assert_array_equal(adv, g("adv")[sl])
        sk = discrete_skips_advantages(g("rewards")[sl], g("values")[sl], g("steps")[sl], bool(g("skip_done")[s]),
                                       g("next_values")[s], gamma, lam)
        np.testing.assert_array_equal(sk, g("skip_adv")[sl])


@pytest.mark.parametrize("case", ["lux_like", "all_episode_end", "scalar_multiplier", "multi_base"])
def test_reward_assembly_matches_the_reference(case):
    """oracle/rewards.py vs the live reference's InfoRewardsWrapper.step (wrappers/info_rewards_wrapper.py:39-57)."""
    from oracle.rewards import assemble_rewards

    z = load("info_rewards")
    mult = z[f"{case}.multiplier"] if f"{case}.multiplier" in z else None
    for t in range(z[f"{case}.base"].shape[0]):
        got = assemble_rewards(z[f"{case}.base"][t], [s.copy() for s in z[f"{case}.series"][t]], z[f"{case}.terminations"][t],
                               z[f"{case}.truncations"][t], z[f"{case}.episode_end"], mult)
        np.testing.assert_array_equal(got, z[f"{case}.rewards"][t])


@pytest.mark.parametrize("tag", ["scalar", "multi"])
def test_ema_reward_normalizer_matches_the_reference(tag):
    """oracle HybridMovingMeanVar / ExponentialMovingMeanVar vs NormalizeReward(exponential_moving_mean_var=True)."""
    from oracle.normalize import RewardNormalizer

    z = load("normalizers_ema")
    rew, dones = z[f"{tag}.rewards"], z[f"{tag}.dones"]
    r = RewardNormalizer(rew.shape[1], rew.shape[2:], gamma=0.98, exponential_moving_mean_var=True,
                         emv_window_size=float(z[f"{tag}.window"]))
    for t in range(rew.shape[0]):
        np.testing.assert_array_equal(r.step(rew[t], dones[t]), z[f"{tag}.out"][t])
    np.testing.assert_array_equal(np.asarray(r.rms.var), z[f"{tag}.var"])
