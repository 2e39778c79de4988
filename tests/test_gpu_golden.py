"""The CUDA path against the golden fixtures generated from the live reference
(tests/golden/*.npz): GAE bit-exact; GridNet / categorical / Gaussian log-prob, entropy and
gradients within 1e-5; and a whole PPO.learn_epoch (same initial weights, same rollout arrays,
same randperm seed) landing on the reference's final parameters and TrainStats."""
import numpy as np
import pytest
import torch

from tests.parity import close, rel_err
from tests.test_oracle_golden import (LEARNER_CASES_ALL, FirstGradients, cases, gates_of, learner_setup, load, rollout_from,
                                      teacher_net)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", cases(load("gae")))
def test_gae_vs_reference_fixture(cuda, name):
    from rl_algo_impls_b200 import ops

    z = load("gae")
    g = lambda k: z[f"{name}.{k}"]
    t = lambda k: torch.from_numpy(g(k)).to(cuda)
    scalar = bool(g("gamma_is_scalar"))
    gamma = float(g("gamma")) if scalar else g("gamma")
    lam = float(g("gae_lambda")) if scalar else g("gae_lambda")
    adv, ret = ops.gae_scan(t("rewards"), t("values"), t("episode_starts"), t("next_episode_starts"), t("next_values"),
                            gamma, lam)
    np.testing.assert_array_equal(adv.cpu().numpy(), g("advantages"))
    np.testing.assert_array_equal(ret.cpu().numpy(), g("returns"))


@pytest.mark.parametrize("name", cases(load("gridnet")))
def test_gridnet_vs_reference_fixture(cuda, name):
    from rl_algo_impls_b200 import ops

    z = load("gridnet")
    g = lambda k: z.get(f"{name}.{k}")
    t = lambda k: None if g(k) is None else torch.from_numpy(g(k)).to(cuda)
    gates = {}
    for h, (r, v) in gates_of(g("gates")).items():
        gates.setdefault(r, {})[h] = v
    n_pick = 0 if g("pick_mask") is None else g("pick_mask").shape[1]
    spec = ops.GridnetSpec.from_subaction_mask(g("nvec").tolist(), gates, n_pick)
    logits = t("logits").requires_grad_(True)
    logp, ent = ops.gridnet_logp_entropy(spec, logits, t("mask"), t("pick_mask"), t("actions"), t("pick_actions"))
    (logp * t("dlogp") + ent * t("dentropy")).sum().backward()
    atol = 4e-7 * float(np.abs(g("logits")).max())
    # a chosen action on a masked entry gives finfo.min (one per sample) or -inf (several): those must match exactly
    want_logp, got_logp = g("logp"), logp.detach().cpu().numpy()
    huge = ~np.isfinite(want_logp) | (np.abs(want_logp) > 1e30)
    np.testing.assert_array_equal(got_logp[huge], want_logp[huge])
    if "masked_actions" in name:
        assert huge.any(), "the fixture is meant to exercise the masked-chosen-action branch"
    close(got_logp[~huge], want_logp[~huge], atol=atol, what="logp")
    close(ent, g("entropy"), atol=atol, what="entropy")
    close(logits.grad, g("dlogits"), atol=atol * 0.1, what="dlogits")
    S = int(g("nvec").sum())
    assert (logits.grad.cpu().numpy()[..., :S][~g("mask")] == 0).all()


@pytest.mark.parametrize("name", ["cartpole", "atari", "masked"])
def test_categorical_vs_reference_fixture(cuda, name):
    from rl_algo_impls_b200 import ops

    z = load("heads")
    g = lambda k: z.get(f"{name}.{k}")
    t = lambda k: None if g(k) is None else torch.from_numpy(g(k)).to(cuda)
    logits = t("logits").requires_grad_(True)
    logp, ent = ops.categorical_logp_entropy(logits, t("mask"), t("actions"))
    (logp * t("dlogp") + ent * t("dentropy")).sum().backward()
    close(logp, g("logp"), atol=1e-6, what="logp")
    close(ent, g("entropy"), atol=1e-6, what="entropy")
    close(logits.grad, g("dlogits"), atol=1e-6, what="dlogits")


def test_gaussian_vs_reference_fixture(cuda):
    from rl_algo_impls_b200 import ops

    z = load("heads")
    t = lambda k: torch.from_numpy(z[f"gaussian.{k}"]).to(cuda)
    logp, ent = ops.gaussian_logp_entropy(t("mu"), t("log_std"), t("actions"))
    close(logp, z["gaussian.logp"], what="logp")
    close(ent, z["gaussian.entropy"], what="entropy")


def _device_policy(case, net, cuda):
    """An ActorCritic over the stub trunk, on the GPU."""
    from rl_algo_impls_b200 import spaces
    from rl_algo_impls_b200.policy import ActorCritic

    class Env:
        num_envs = case["N"]
        single_observation_space = spaces.Box(-np.inf, np.inf, case["obs_shape"], np.float32)

    env = Env()
    if case["kind"] == "categorical":
        env.single_action_space = spaces.Discrete(case["nvec"][0])
    elif case["kind"] == "gaussian":
        env.single_action_space = spaces.Box(-1, 1, (case["nvec"][0],), np.float32)
    else:
        hw = case["side"] ** 2
        env.action_plane_space = spaces.MultiDiscrete(case["nvec"])
        per_pos = spaces.MultiDiscrete(np.tile(np.asarray(case["nvec"]), hw))
        env.single_action_space = per_pos if not case.get("n_pick") else spaces.Dict(
            {"per_position": per_pos, "pick_position": spaces.MultiDiscrete([hw] * case["n_pick"])})
    net.n_values = case["V"]
    return ActorCritic(env, network=net, subaction_mask=case.get("gates")).to(cuda)


@pytest.mark.parametrize("graphed", [False, True], ids=["eager", "graphed"])
@pytest.mark.parametrize("name", LEARNER_CASES_ALL)
def test_learn_epoch_vs_reference_fixture(cuda, name, graphed):
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import VecRollout

    case, z, net = learner_setup(name)
    hp = case["hp"]
    torch.backends.cudnn.allow_tf32 = False  # the trunk is PyTorch's; keep its convs in fp32 like the CPU reference
    torch.backends.cuda.matmul.allow_tf32 = False
    policy = _device_policy(case, net, cuda)
    ro = rollout_from(z)

    class Gen:
        n_steps = case["T"]
        vec_env = type("E", (), {"num_envs": case["N"]})()

        def rollout(self, gamma, gae_lambda):
            return VecRollout(cuda, ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"], ro["rewards"],
                              ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma, gae_lambda,
                              subaction_mask=case.get("gates"))

    kw = {k: getattr(hp, k) for k in ("batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range", "clip_range_vf",
                                      "normalize_advantage", "standardize_advantage", "ent_coef", "vf_coef",
                                      "ppo2_vf_coef_halving", "max_grad_norm", "multi_reward_weights",
                                      "gradient_accumulation", "kl_cutoff", "normalize_advantages_after_scaling",
                                      "learning_rate", "vf_loss_fn", "vf_weights", "autocast_loss")}
    tnet = teacher_net(case, z)
    if tnet is not None:  # teacher-KL term: the teacher checkpoint is an ActorCritic over the stored teacher weights
        from rl_algo_impls_b200.loss import TeacherKLLoss

        teacher = _device_policy(case, tnet, cuda)
        mgr = type("Mgr", (), {"latest_checkpoint": teacher})()
        kw.update(teacher_kl_loss_coef=hp.teacher_kl_loss_coef,
                  teacher_kl_loss_fn=TeacherKLLoss(mgr, unbiased=hp.teacher_unbiased),
                  teacher_loss_importance_sampling=hp.teacher_loss_importance_sampling)
    algo = PPO(policy, cuda, None, **kw)
    algo.cuda_graph_update = graphed
    box = {}

    class CB:
        def on_step(self, timesteps_elapsed, train_stats):
            box["s"] = train_stats
            return True

    # the gradient handed to the first clip + Adam (flat-buffer path: PPO._clip_and_step; per-parameter path:
    # torch.nn.utils.clip_grad_norm_), copied out before anything scales it
    first = FirstGradients(policy.network.named_parameters())
    orig_step = algo._clip_and_step

    def spy(flat, world):
        first.capture()
        return orig_step(flat, world)

    algo._clip_and_step = spy
    torch.manual_seed(int(z["seed"]) + 100)  # same randperm stream as the reference run
    with first:
        steps, cont = algo.learn_epoch(0, case["T"] * case["N"], Gen(), [CB()])
    assert steps == case["T"] * case["N"] and cont
    s = box["s"]
    # bf16 autocast (CUDA only in the reference: shared/autocast.py:8-12) runs the trunk in bf16; the CPU reference that
    # made the fixture ran f32, so that case is held to a bf16 bar: 2^-8 relative steps through the trunk
    bf16 = bool(hp.autocast_loss)
    # ---- pre-Adam gradient of the first minibatch: 1e-5 of the tensor's largest entry, conditioned (tests/parity.py) ----
    if not graphed:  # a captured update takes its first step inside the warm-up runs of the capture
        assert first.grads, "no gradient was captured"
        for k, g in first.grads.items():
            want32, want64 = z[f"grad0.{k}"], z[f"grad0_f64.{k}"]
            e32 = rel_err(g, torch.from_numpy(want32))
            if bf16:
                assert e32 <= 5e-2, f"{name} first gradient {k}: {e32:.2e} (bf16 trunk)"
                continue
            if e32 <= 1e-5:
                continue
            e64 = rel_err(g, torch.from_numpy(want64))
            ref = rel_err(torch.from_numpy(want32), torch.from_numpy(want64))
            assert e64 <= max(1e-5, 3 * ref), (f"{name} first gradient {k}: {e32:.2e} from the f32 reference, {e64:.2e} from "
                                               f"the f64 oracle (the f32 reference itself: {ref:.2e})")
    # parameters after n_epochs x minibatches of Adam steps: each step is lr * a unit-scale update, so
    # compare against the parameter *change* (final - init), not the parameter magnitude
    # Adam divides by sqrt(v): an element whose gradient is ~0 takes a +-lr step on rounding noise
    # alone, so the bound is two-sided -- RMS error small against the RMS update, and no element
    # further off than the steps Adam could have taken.
    n_updates = hp.n_epochs * (1 if hp.gradient_accumulation else -(-case["T"] * case["N"] // hp.batch_size))
    for k, v in policy.network.state_dict().items():
        want, init = z[f"final.{k}"].astype(np.float64), z[f"init.{k}"].astype(np.float64)
        got = v.cpu().numpy().astype(np.float64)
        rms_update = np.sqrt(np.mean((want - init) ** 2))
        rms_err = np.sqrt(np.mean((got - want) ** 2))
        bar = 2e-1 if bf16 else 1e-2
        assert rms_err <= bar * rms_update + 1e-8, f"{name} param {k}: rms err {rms_err:.3e} vs rms update {rms_update:.3e}"
        assert np.abs(got - want).max() <= 2 * hp.learning_rate * n_updates
    for k, tol in (("loss", 1e-4), ("pi_loss", 2e-3), ("entropy_loss", 1e-5), ("approx_kl", 2e-3), ("grad_norm", 1e-3),
                   ("explained_var", 1e-5)):
        got, want = getattr(s, k), float(z[f"stats.{k}"])
        if bf16 and k != "explained_var":  # explained_var comes from the rollout alone
            tol = 5e-2
        assert abs(got - want) <= tol * max(abs(want), 1e-2), f"{name} {k}: {got} vs {want}"
    np.testing.assert_allclose(np.asarray(s.v_loss, np.float64), z["stats.v_loss"], rtol=2e-2 if bf16 else 1e-4)
    assert np.asarray(s.v_loss).shape == z["stats.v_loss"].shape
    assert abs(s.clipped_frac - float(z["stats.clipped_frac"])) <= (4.0 if bf16 else 2.0) / hp.batch_size
    if tnet is not None:
        want = float(z["stats.teacher_kl_loss"])
        assert abs(s.additional_losses["teacher_kl_loss"] - want) <= 2e-3 * max(abs(want), 1e-2)


@pytest.mark.parametrize("name", ["cartpole", "microrts"])
def test_a2c_iteration_vs_reference_fixture(cuda, name):
    """A2C.learn (one iteration) on the device path vs the live reference's final parameters / losses."""
    from rl_algo_impls_b200.a2c import A2C
    from rl_algo_impls_b200.rollout import VecRollout
    from tests.test_oracle_golden import a2c_setup

    case, z, net = a2c_setup(name)
    hp = case["hp"]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    policy = _device_policy(case, net, cuda)
    ro = rollout_from(z)

    class Gen:
        vec_env = type("E", (), {"num_envs": case["N"]})()

        def rollout(self, gamma, gae_lambda):
            return VecRollout(cuda, ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"], ro["rewards"],
                              ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma, gae_lambda,
                              subaction_mask=case.get("gates"))

    algo = A2C(policy, cuda, None, learning_rate=hp.learning_rate, gamma=hp.gamma, gae_lambda=hp.gae_lambda,
               ent_coef=hp.ent_coef, vf_coef=hp.vf_coef, max_grad_norm=hp.max_grad_norm, rms_prop_eps=hp.rms_prop_eps,
               use_rms_prop=hp.use_rms_prop, normalize_advantage=hp.normalize_advantage,
               gradient_accumulation=hp.gradient_accumulation, num_minibatches=hp.num_minibatches)
    torch.manual_seed(int(z["seed"]) + 100)
    algo.learn(case["T"] * case["N"], Gen())
    stats = algo.last_train_stats.data
    for k, tol in (("loss", 1e-4), ("pi_loss", 1e-3), ("entropy_loss", 1e-5), ("explained_var", 1e-5)):
        want = float(z[f"stats.{k}"])
        assert abs(stats[k] - want) <= tol * max(abs(want), 1e-2), (k, stats[k], want)
    for k, v in policy.network.state_dict().items():
        want, init = z[f"final.{k}"].astype(np.float64), z[f"init.{k}"].astype(np.float64)
        got = v.cpu().numpy().astype(np.float64)
        rms_update = np.sqrt(np.mean((want - init) ** 2))
        rms_err = np.sqrt(np.mean((got - want) ** 2))
        assert rms_err <= 1e-2 * rms_update + 1e-8, f"{name} param {k}: rms err {rms_err:.3e} vs rms update {rms_update:.3e}"


def test_acbc_iteration_vs_reference_fixture(cuda):
    """ACBC.learn (one iteration, two epochs) on the device path vs the live reference."""
    from rl_algo_impls_b200.acbc import ACBC
    from rl_algo_impls_b200.rollout import VecRollout
    from tests.golden.make_golden_cases import A2C_CASES, make_net_for
    from tests.test_oracle_golden import load

    case, z = A2C_CASES["microrts"], load("acbc_microrts")
    net = make_net_for(case)()
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("init.")})
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    policy = _device_policy(case, net, cuda)
    ro = rollout_from(z)

    class Gen:
        vec_env = type("E", (), {"num_envs": case["N"]})()

        def rollout(self, gamma, gae_lambda):
            return VecRollout(cuda, ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"], ro["rewards"],
                              ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma, gae_lambda,
                              subaction_mask=case["gates"])

    algo = ACBC(policy, cuda, None, learning_rate=float(z["hp.learning_rate"]), batch_size=int(z["hp.batch_size"]),
                n_epochs=int(z["hp.n_epochs"]), gamma=float(z["hp.gamma"]), gae_lambda=float(z["hp.gae_lambda"]),
                vf_coef=float(z["hp.vf_coef"]))
    torch.manual_seed(int(z["seed"]) + 100)
    algo.learn(case["T"] * case["N"], Gen())
    for k, tol in (("loss", 1e-4), ("pi_loss", 1e-4)):
        want = float(z[f"stats.{k}"])
        assert abs(algo.last_stats[k] - want) <= tol * max(abs(want), 1e-2), (k, algo.last_stats[k], want)
    for k, v in policy.network.state_dict().items():
        want, init = z[f"final.{k}"].astype(np.float64), z[f"init.{k}"].astype(np.float64)
        got = v.cpu().numpy().astype(np.float64)
        rms_update = np.sqrt(np.mean((want - init) ** 2))
        rms_err = np.sqrt(np.mean((got - want) ** 2))
        assert rms_err <= 1e-2 * rms_update + 1e-8, f"acbc param {k}: rms err {rms_err:.3e} vs rms update {rms_update:.3e}"


def test_learn_epoch_with_kl_cutoff_matches_the_oracle(cuda):
    """kl_cutoff (ppo.py:352-355): the cut-off decision is taken on the device between forward and
    backward (no host sync) and stays sticky for the rest of the learn_epoch; same final parameters as
    the oracle learner, whose cut-off logic is the reference's."""
    from oracle import learner as olearn
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import VecRollout
    from tests.golden.make_golden_cases import make_net_for

    case, z, net = learner_setup("microrts")
    import copy
    import dataclasses

    hp = dataclasses.replace(case["hp"], kl_cutoff=2e-3, batch_size=24, n_epochs=3)
    ro = rollout_from(z)
    ro["logprobs"] = ro["logprobs"] + 0.05  # push approx_kl over the cut-off after the first updates
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cpu_net = copy.deepcopy(net)
    opol = olearn.OraclePolicy(cpu_net, case["kind"], case["nvec"], case["side"] ** 2, case["gates"])
    opt = torch.optim.Adam(cpu_net.parameters(), lr=hp.learning_rate, eps=1e-7)
    torch.manual_seed(7)
    ostats = olearn.learn_epoch(opol, opt, ro, hp)

    policy = _device_policy(case, net, cuda)

    class Gen:
        n_steps = case["T"]
        vec_env = type("E", (), {"num_envs": case["N"]})()

        def rollout(self, gamma, gae_lambda):
            return VecRollout(cuda, ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"], ro["rewards"],
                              ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma, gae_lambda,
                              subaction_mask=case.get("gates"))

    kw = {k: getattr(hp, k) for k in ("batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range", "clip_range_vf",
                                      "ent_coef", "vf_coef", "ppo2_vf_coef_halving", "max_grad_norm", "kl_cutoff",
                                      "learning_rate")}
    algo = PPO(policy, cuda, None, **kw)
    torch.manual_seed(7)
    algo.learn_epoch(0, case["T"] * case["N"], Gen(), None)
    s = algo.last_train_stats
    assert abs(s.approx_kl - ostats["approx_kl"]) <= 2e-3 * max(abs(ostats["approx_kl"]), 1e-3)
    assert abs(s.loss - ostats["loss"]) <= 1e-3 * max(abs(ostats["loss"]), 1e-2)
    for (k, v), w in zip(policy.network.state_dict().items(), cpu_net.state_dict().values()):
        err = (v.cpu().double() - w.double()).pow(2).mean().sqrt().item()
        scale = (w.double() - torch.from_numpy(z[f"init.{k}"]).double()).pow(2).mean().sqrt().item()
        assert err <= 1e-2 * scale + 1e-8, (k, err, scale)
