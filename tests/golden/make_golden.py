"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (rl-algo-impls at /root/reference)
and pin the oracle restatement against it.  Run in the build container only:

    python tests/golden/make_golden.py

For every fixture the script (1) runs the reference's own function / class on seeded inputs,
(2) asserts that the oracle (oracle/*.py) reproduces the reference output (bit-exact on CPU, since
both are the same torch / numpy arithmetic), and (3) stores inputs + reference outputs.  The
fixtures are what travels: neither the tests nor bench.py import the reference at run time.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _ref_shim  # noqa: E402

_ref_shim.install()

from rl_algo_impls.a2c.a2c import A2C as RefA2C  # noqa: E402
from rl_algo_impls.acbc.acbc import ACBC as RefACBC  # noqa: E402
from rl_algo_impls.loss.teacher_kl_loss import TeacherKLLoss as RefTeacherKLLoss  # noqa: E402
from rl_algo_impls.ppo.ppo import PPO as RefPPO  # noqa: E402
from rl_algo_impls.rollout.vec_rollout import VecRollout as RefVecRollout  # noqa: E402
from rl_algo_impls.shared.actor.categorical import MaskedCategorical as RefMaskedCategorical  # noqa: E402
from rl_algo_impls.shared.actor.gaussian import GaussianDistribution as RefGaussian  # noqa: E402
from rl_algo_impls.shared.actor.gridnet import GridnetDistribution as RefGridnet  # noqa: E402
from rl_algo_impls.shared.actor.gridnet import ValueDependentMask as RefVDM  # noqa: E402
from rl_algo_impls.shared.gae import compute_advantages as ref_compute_advantages  # noqa: E402
from rl_algo_impls.shared.policy.actor_critic import clamp_actions as ref_clamp_actions  # noqa: E402

from oracle import learner as olearn  # noqa: E402
from oracle.distributions import Gridnet, MaskedLogits, gates_from_subaction_mask, gaussian_logp_entropy  # noqa: E402
from oracle.gae import gae_advantages, gae_returns  # noqa: E402
from oracle.normalize import ObsNormalizer, RewardNormalizer  # noqa: E402
from oracle.rollout import minibatch_index_stream, num_actions  # noqa: E402
from tests.golden.stub_nets import TinyGrid, TinyMlp  # noqa: E402
from tests.synth import (LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gae_inputs, gridnet_inputs,  # noqa: E402
                         to_torch)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: v for k, v in arrays.items() if v is not None})
    print(f"wrote {name}.npz ({os.path.getsize(path) / 1024:.1f} KB)")


def exact(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), f"oracle != reference: {what}"


# ---------------------------------------------------------------------------------------------
def golden_gae():
    cases = {
        "c1_cartpole": (32, 8, 1, 1 / 20, 0.98, 0.8),
        "c4_microrts": (512, 24, 1, 1 / 200, 0.999, 0.99),
        "ragged": (7, 5, 1, 0.3, 0.9, 0.5),
        "multihead_scalar_gamma": (16, 6, 3, 0.1, 0.99, 0.95),
        "c5_lux_per_head": (32, 16, 13, 1 / 10, "per_head", "per_head"),
    }
    out = {}
    for name, (T, N, V, p, gamma, lam) in cases.items():
        inp = gae_inputs(zlib.crc32(name.encode()) % 1000, T, N, V, p)
        if gamma == "per_head":
            gamma, lam = np.linspace(1.0, 0.95, V), np.full(V, 0.95)
        ref = ref_compute_advantages(inp["rewards"], inp["values"], inp["episode_starts"], inp["next_episode_starts"],
                                     inp["next_values"], gamma, lam)
        exact(gae_advantages(gamma=gamma, gae_lambda=lam, **inp), ref, f"gae {name}")
        for k, v in inp.items():
            out[f"{name}.{k}"] = v
        out[f"{name}.gamma"], out[f"{name}.gae_lambda"] = np.asarray(gamma), np.asarray(lam)
        out[f"{name}.gamma_is_scalar"] = np.asarray(not isinstance(gamma, np.ndarray))
        out[f"{name}.advantages"] = ref
        out[f"{name}.returns"] = ref + inp["values"]  # vec_rollout.py:88
    save("gae", **out)


def _ref_gridnet(inp, nvec, gates, HW, side):
    masks = inp["mask"] if inp["pick_mask"] is None else {"per_position": inp["mask"], "pick_position": inp["pick_mask"]}
    action = inp["actions"] if inp["pick_actions"] is None else {
        "per_position": inp["actions"], "pick_position": inp["pick_actions"]}
    logits = inp["logits"].reshape(inp["logits"].shape[0], side, side, -1).clone().requires_grad_(True)
    sub = RefVDM.from_reference_index_to_index_to_value(gates) if gates else None
    return logits, RefGridnet(HW, np.asarray(nvec), logits, masks, subaction_mask=sub), action, masks


def golden_gridnet():
    cases = {
        "microrts_8x8": (5, 8, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.15),
        "microrts_8x8_ungated_dense": (3, 8, MICRORTS_NVEC, None, 0, 0.6),
        "lux_8x8_pick": (4, 8, LUX_NVEC, LUX_GATES, 1, 0.1),
        "all_masked": (2, 4, LUX_NVEC, LUX_GATES, 1, 0.0),
        # round 2: the chosen action lands on a MASKED entry of a head that has valid entries (categorical.py:25-36 keeps
        # the finfo.min logit: log-prob finfo.min, -inf once two of them meet in one sample's sum)
        "microrts_8x8_masked_actions": (6, 8, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.2, 0.05),
        "lux_8x8_pick_masked_actions": (6, 8, LUX_NVEC, LUX_GATES, 1, 0.15, 0.05),
    }
    out = {}
    for name, (B, side, nvec, gates, n_pick, unit_p, *rest) in cases.items():
        HW = side * side
        inp = to_torch(gridnet_inputs(zlib.crc32(name.encode()) % 1000, B, HW, nvec, n_pick, unit_p, logit_scale=2.0,
                                      masked_action_p=rest[0] if rest else 0.0))
        logits, dist, action, masks = _ref_gridnet(inp, nvec, gates, HW, side)
        logp, ent = dist.log_prob(action), dist.entropy()
        g = torch.Generator().manual_seed(1)
        dlogp, dent = torch.randn(B, generator=g), torch.randn(B, generator=g)
        (logp * dlogp + ent * dent).sum().backward()
        # the oracle restatement must agree exactly
        ol = inp["logits"].clone().requires_grad_(True)
        od = Gridnet(HW, nvec, ol, masks, gates_from_subaction_mask(gates))
        olp, oent = od.log_prob(action), od.entropy()
        (olp * dlogp + oent * dent).sum().backward()
        exact(olp.detach(), logp.detach(), f"gridnet logp {name}")
        exact(oent.detach(), ent.detach(), f"gridnet entropy {name}")
        exact(ol.grad, logits.grad.reshape(ol.shape), f"gridnet dlogits {name}")
        for k, v in inp.items():
            if v is not None:
                out[f"{name}.{k}"] = v.numpy()
        out[f"{name}.nvec"] = np.asarray(nvec)
        out[f"{name}.gates"] = np.asarray([[h, r, v] for r, d in (gates or {}).items() for h, v in d.items()]).reshape(-1, 3)
        out[f"{name}.logp"], out[f"{name}.entropy"] = logp.detach().numpy(), ent.detach().numpy()
        out[f"{name}.dlogp"], out[f"{name}.dentropy"] = dlogp.numpy(), dent.numpy()
        out[f"{name}.dlogits"] = logits.grad.reshape(ol.shape).numpy()
    save("gridnet", **out)


def golden_categorical_gaussian():
    g = torch.Generator().manual_seed(7)
    out = {}
    for name, (B, n, masked) in {"cartpole": (64, 2, False), "atari": (64, 4, False), "masked": (70, 6, True)}.items():
        logits = (torch.randn(B, n, generator=g) * 2).requires_grad_(True)
        mask = None
        if masked:
            mask = torch.rand(B, n, generator=g) < 0.6
            mask[::7] = False
            mask[1::7, 0] = True
        dist = RefMaskedCategorical(logits=logits, mask=mask)
        actions = dist.sample()
        logp, ent = dist.log_prob(actions), dist.entropy()
        dl, de = torch.randn(B, generator=g), torch.randn(B, generator=g)
        (logp * dl + ent * de).sum().backward()
        ol = logits.detach().clone().requires_grad_(True)
        od = MaskedLogits(ol, mask)
        olp, oent = od.log_prob(actions), od.entropy()
        (olp * dl + oent * de).sum().backward()
        exact(olp.detach(), logp.detach(), f"categorical logp {name}")
        exact(oent.detach(), ent.detach(), f"categorical entropy {name}")
        # the two gradient contributions meet in a different autograd accumulation order: 1-ulp noise
        assert torch.allclose(ol.grad, logits.grad, rtol=1e-6, atol=1e-7), f"categorical dlogits {name}"
        out.update({f"{name}.logits": logits.detach().numpy(), f"{name}.actions": actions.numpy(),
                    f"{name}.logp": logp.detach().numpy(), f"{name}.entropy": ent.detach().numpy(),
                    f"{name}.dlogp": dl.numpy(), f"{name}.dentropy": de.numpy(), f"{name}.dlogits": logits.grad.numpy()})
        if mask is not None:
            out[f"{name}.mask"] = mask.numpy()
    B, D = 48, 6
    mu = torch.randn(B, D, generator=g).requires_grad_(True)
    log_std = (torch.full((D,), -2.0) + 0.1 * torch.randn(D, generator=g)).requires_grad_(True)
    dist = RefGaussian(mu, torch.exp(log_std))
    actions = (mu + torch.exp(log_std) * torch.randn(B, D, generator=g)).detach()
    logp, ent = dist.log_prob(actions), dist.entropy()
    dl, de = torch.randn(B, generator=g), torch.randn(B, D, generator=g)
    (logp * dl).sum().backward(retain_graph=True)
    (ent * de).sum().backward()
    olp, oent = gaussian_logp_entropy(mu.detach(), log_std.detach(), actions)
    exact(olp, logp.detach(), "gaussian logp")
    exact(oent, ent.detach(), "gaussian entropy")
    out.update({"gaussian.mu": mu.detach().numpy(), "gaussian.log_std": log_std.detach().numpy(),
                "gaussian.actions": actions.numpy(), "gaussian.logp": logp.detach().numpy(),
                "gaussian.entropy": ent.detach().numpy(), "gaussian.dlogp": dl.numpy(), "gaussian.dentropy": de.numpy(),
                "gaussian.dmu": mu.grad.numpy(), "gaussian.dlog_std": log_std.grad.numpy()})
    save("heads", **out)


# ---------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, name, value, *a, **k):
        self.scalars[name] = value

    def on_steps(self, *a, **k):
        pass


import collections  # noqa: E402

_ACForward = collections.namedtuple("_ACForward", "logp_a entropy v")


class _RefPolicy(torch.nn.Module):
    """forward(obs, actions, action_masks) -> (logp_a, entropy, v) over the REFERENCE distributions."""

    def __init__(self, network, kind, nvec=(), map_size=0, gates=None):
        super().__init__()
        self.network, self.kind, self.nvec, self.map_size = network, kind, np.asarray(nvec), map_size
        self.sub = RefVDM.from_reference_index_to_index_to_value(gates) if gates else None

    def reset_noise(self, *_a):
        pass

    def forward(self, obs, actions, action_masks=None):
        out = self.network(obs)
        if self.kind == "gridnet":
            pi = RefGridnet(self.map_size, self.nvec, out.pi, action_masks, subaction_mask=self.sub)
        elif self.kind == "categorical":
            pi = RefMaskedCategorical(logits=out.pi, mask=action_masks)
        else:
            pi = RefGaussian(out.pi, torch.exp(out.log_std))
        return _ACForward(pi.log_prob(actions), pi.entropy(), out.values)


def _f64(a):
    if a is None:
        return None
    a = np.asarray(a)
    return a.astype(np.float64) if a.dtype == np.float32 else a


class _StopAfterFirstGradient(Exception):
    pass


class _FirstGradients:
    """Instruments the run from OUTSIDE: wraps torch.nn.utils.clip_grad_norm_ (what optimizer_step calls first,
    ppo.py:441-443) and keeps a copy of every parameter's gradient at its first call -- the gradient of the first
    minibatch (of the first whole epoch under gradient accumulation), before clipping and before Adam."""

    def __init__(self, net, stop: bool = False):
        self.net, self.stop, self.grads = net, stop, {}

    def __enter__(self):
        self._orig = torch.nn.utils.clip_grad_norm_

        def wrapped(parameters, *a, **k):
            if not self.grads:
                self.grads = {n: p.grad.detach().clone() for n, p in self.net.named_parameters() if p.grad is not None}
                if self.stop:
                    raise _StopAfterFirstGradient()
            return self._orig(parameters, *a, **k)

        torch.nn.utils.clip_grad_norm_ = wrapped
        return self

    def __exit__(self, exc_type, exc, tb):
        torch.nn.utils.clip_grad_norm_ = self._orig
        return exc_type is _StopAfterFirstGradient


class _Gen:
    def __init__(self, rollout, num_envs):
        self._r = rollout
        self.vec_env = type("E", (), {"num_envs": num_envs})()

    def rollout(self, gamma, gae_lambda):
        return self._r(gamma, gae_lambda)


def _learner_case(name, kind, make_net, T, N, obs_shape, hp: olearn.Hyper, nvec=(), side=0, gates=None, n_pick=0, V=1,
                  seed=0):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    net = make_net()
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    HW = side * side
    ro = gae_inputs(seed + 1, T, N, V, 0.1)
    ro["obs"] = rng.standard_normal((T, N) + obs_shape, dtype=np.float32)
    ro["masks"] = None
    if kind == "categorical":
        ro["actions"] = rng.integers(0, nvec[0], size=(T, N))
    elif kind == "gaussian":
        ro["actions"] = rng.standard_normal((T, N, nvec[0]), dtype=np.float32)
    else:
        g = gridnet_inputs(seed + 2, T * N, HW, nvec, n_pick, 0.2)
        cells, mask = g["actions"].reshape(T, N, HW, len(nvec)), g["mask"].reshape(T, N, HW, -1)
        if n_pick:
            ro["actions"] = {"per_position": cells, "pick_position": g["pick_actions"].reshape(T, N, n_pick)}
            ro["masks"] = {"per_position": mask, "pick_position": g["pick_mask"].reshape(T, N, n_pick, HW)}
        else:
            ro["actions"], ro["masks"] = cells, mask
    # behaviour log-probs: the initial policy's own log-prob of the stored actions, plus noise
    pol = _RefPolicy(net, kind, nvec, HW, gates)
    flat = lambda a: ({k: torch.as_tensor(v.reshape((-1,) + v.shape[2:])) for k, v in a.items()} if isinstance(a, dict)
                      else (None if a is None else torch.as_tensor(a.reshape((-1,) + a.shape[2:]))))
    with torch.no_grad():
        lp = pol(flat(ro["obs"]), flat(ro["actions"]), flat(ro["masks"]))[0]
    ro["logprobs"] = (lp.numpy().reshape(T, N) + rng.standard_normal((T, N)).astype(np.float32) * 0.1).astype(np.float32)
    if hp.logp_shift:
        ro["logprobs"] = (ro["logprobs"] + np.float32(hp.logp_shift)).astype(np.float32)

    def ref_rollout(gamma, gae_lambda):
        return RefVecRollout(torch.device("cpu"), ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"],
                             ro["rewards"], ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma,
                             gae_lambda, subaction_mask=gates,
                             action_plane_space=_ref_shim.MultiDiscrete(nvec) if kind == "gridnet" else None)

    kw = {k: getattr(hp, k) for k in ("batch_size", "n_epochs", "clip_range", "clip_range_vf", "normalize_advantage",
                                      "standardize_advantage", "ent_coef", "ppo2_vf_coef_halving", "max_grad_norm",
                                      "gradient_accumulation", "kl_cutoff", "normalize_advantages_after_scaling",
                                      "learning_rate", "vf_loss_fn", "autocast_loss")}
    if hp.vf_weights is not None:
        kw["vf_weights"] = list(hp.vf_weights)
    as_list = lambda x: x.tolist() if isinstance(x, np.ndarray) else x
    teacher_net = teacher_pol = None
    if hp.teacher_kl_loss_coef:
        torch.manual_seed(seed + 50)
        teacher_net = make_net()
        teacher_pol = _RefPolicy(teacher_net, kind, nvec, HW, gates)
        mgr = type("Mgr", (), {"latest_checkpoint": teacher_pol})()
        kw.update(teacher_kl_loss_coef=hp.teacher_kl_loss_coef,
                  teacher_kl_loss_fn=RefTeacherKLLoss(mgr, unbiased=hp.teacher_unbiased),
                  teacher_loss_importance_sampling=hp.teacher_loss_importance_sampling)
    algo = RefPPO(pol, torch.device("cpu"), _Writer(), gamma=as_list(hp.gamma), gae_lambda=as_list(hp.gae_lambda),
                  vf_coef=as_list(hp.vf_coef) if not np.isscalar(hp.vf_coef) else hp.vf_coef,
                  multi_reward_weights=list(hp.multi_reward_weights) if hp.multi_reward_weights is not None else None, **kw)
    stats_box = {}

    class _CB:
        def on_step(self, timesteps_elapsed, train_stats):
            stats_box["s"] = train_stats
            return True

    torch.manual_seed(seed + 100)  # the randperm stream of the minibatch loop
    with _FirstGradients(net) as ref_grads:  # the gradient the reference hands to its FIRST clip_grad_norm_ (ppo.py:441-443)
        algo.learn_epoch(0, T * N, _Gen(ref_rollout, N), [_CB()])
    s = stats_box["s"]
    final = {k: v.detach().clone() for k, v in net.state_dict().items()}

    # the oracle learner from the same start must land on the same parameters
    net2 = make_net()
    net2.load_state_dict(init)
    opol = olearn.OraclePolicy(net2, kind, nvec, HW, gates)
    opt = torch.optim.Adam(net2.parameters(), lr=hp.learning_rate, eps=1e-7)
    torch.manual_seed(seed + 100)
    oteacher = olearn.OraclePolicy(teacher_net, kind, nvec, HW, gates) if teacher_net is not None else None
    with _FirstGradients(net2) as oracle_grads:
        ostats = olearn.learn_epoch(opol, opt, ro, hp, teacher=oteacher)
    for k, g in ref_grads.grads.items():
        exact(oracle_grads.grads[k], g, f"learner {name} first gradient {k}")
    # the same first gradient evaluated in float64 (the conditioned bar of tests/parity.py needs it)
    net64 = make_net()
    net64.load_state_dict(init)
    net64.double()
    ro64 = {k: ({kk: _f64(vv) for kk, vv in v.items()} if isinstance(v, dict) else _f64(v)) for k, v in ro.items()}
    torch.manual_seed(seed + 100)
    teacher64 = None
    if teacher_net is not None:
        import copy
        teacher64 = olearn.OraclePolicy(copy.deepcopy(teacher_net).double(), kind, nvec, HW, gates)
    with _FirstGradients(net64, stop=True) as grads64:
        olearn.learn_epoch(olearn.OraclePolicy(net64, kind, nvec, HW, gates),
                           torch.optim.Adam(net64.parameters(), lr=hp.learning_rate, eps=1e-7), ro64, hp, teacher=teacher64)
    if teacher_net is not None:
        exact(np.float64(ostats["teacher_kl_loss"]), np.float64(s.additional_losses["teacher_kl_loss"]), f"learner {name} teacher_kl_loss")
    for k, v in net2.state_dict().items():
        exact(v, final[k], f"learner {name} param {k}")
    for k in ("loss", "pi_loss", "entropy_loss", "approx_kl", "clipped_frac", "grad_norm", "explained_var"):
        exact(np.float64(ostats[k]), np.float64(getattr(s, k)), f"learner {name} stat {k}")
    exact(np.asarray(ostats["v_loss"], np.float64), np.asarray(s.v_loss, np.float64), f"learner {name} v_loss")

    out = {f"init.{k}": v.numpy() for k, v in init.items()}
    out.update({f"final.{k}": v.numpy() for k, v in final.items()})
    out.update({f"grad0.{k}": v.numpy() for k, v in ref_grads.grads.items()})
    out.update({f"grad0_f64.{k}": v.numpy() for k, v in grads64.grads.items()})
    for k, v in ro.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[f"ro.{k}.{kk}"] = vv
        elif v is not None:
            out[f"ro.{k}"] = v
    for k in ("loss", "pi_loss", "entropy_loss", "approx_kl", "clipped_frac", "grad_norm", "explained_var"):
        out[f"stats.{k}"] = np.float64(getattr(s, k))
    out["stats.v_loss"] = np.asarray(s.v_loss, np.float64)
    out["stats.val_clipped_frac"] = np.asarray(s.val_clipped_frac, np.float64)
    out["seed"] = np.asarray(seed)
    if teacher_net is not None:
        out.update({f"teacher.{k}": v.detach().numpy() for k, v in teacher_net.state_dict().items()})
        out["stats.teacher_kl_loss"] = np.float64(s.additional_losses["teacher_kl_loss"])
    save("learner_" + name, **out)


from tests.golden.make_golden_cases import LEARNER_CASES, make_net_for  # noqa: E402


def golden_learner(only=None):
    for i, (name, case) in enumerate(LEARNER_CASES.items()):
        if only is not None and name not in only:
            continue
        _learner_case(name, case["kind"], make_net_for(case), case["T"], case["N"], case["obs_shape"], case["hp"],
                      nvec=case["nvec"], side=case.get("side", 0), gates=case.get("gates"), n_pick=case.get("n_pick", 0),
                      V=case["V"], seed=10 + i)


def golden_index_stream_and_misc():
    # VecRollout.minibatches: obs = arange(T*N) so each minibatch's obs ARE its indices
    T, N, bs = 6, 7, 10
    inp = gae_inputs(3, T, N, 1, 0.1)
    obs = np.arange(T * N, dtype=np.int64).reshape(T, N)
    r = RefVecRollout(torch.device("cpu"), inp["next_episode_starts"], inp["next_values"], obs,
                      np.zeros((T, N), np.int64), inp["rewards"], inp["episode_starts"], inp["values"],
                      np.zeros((T, N), np.float32), None, 0.99, 0.95)
    torch.manual_seed(123)
    ref_idx = [mb.obs.numpy() for mb in r.minibatches(bs, shuffle=True)]
    torch.manual_seed(123)
    ours = [i.numpy() for i in minibatch_index_stream(T * N, bs, shuffle=True)]
    assert len(ref_idx) == len(ours) and all(np.array_equal(a, b) for a, b in zip(ref_idx, ours))
    ref_seq = [mb.obs.numpy() for mb in r.minibatches(bs, shuffle=False)]
    # clamp_actions: the reference's one unit test (tests/shared/policy/test_actor_critic.py:8-17)
    box1, box2 = _ref_shim.Box(-1, 1, (1,)), _ref_shim.Box(-3, 2, (1,))
    c1 = ref_clamp_actions(np.array([-1.5, 0, 1.5]), box1, squash_output=False)
    c2 = ref_clamp_actions(np.array([-1, 0, 1]), box2, squash_output=True)
    save("index_stream", seed=np.asarray(123), total=np.asarray(T * N), batch_size=np.asarray(bs),
         shuffled=np.concatenate(ref_idx), sizes=np.asarray([len(x) for x in ref_idx]),
         sequential=np.concatenate(ref_seq), clamp_noscale=c1, clamp_squash=c2)


def golden_a2c():
    """One A2C.learn iteration of the live reference (a2c/a2c.py:104-173) on fixed rollouts."""
    from tests.golden.make_golden_cases import A2C_CASES, make_net_for

    for ci, (name, case) in enumerate(A2C_CASES.items()):
        seed = 40 + ci
        rng = np.random.default_rng(seed)
        torch.manual_seed(seed)
        make_net = make_net_for(case)
        net = make_net()
        init = {k: v.detach().clone() for k, v in net.state_dict().items()}
        T, N, V, kind, nvec = case["T"], case["N"], case["V"], case["kind"], case["nvec"]
        side = case.get("side", 0)
        HW = side * side
        gates = case.get("gates")
        ro = gae_inputs(seed + 1, T, N, V, 0.1)
        ro["obs"] = rng.standard_normal((T, N) + case["obs_shape"], dtype=np.float32)
        ro["masks"] = None
        if kind == "categorical":
            ro["actions"] = rng.integers(0, nvec[0], size=(T, N))
        else:
            g = gridnet_inputs(seed + 2, T * N, HW, nvec, 0, 0.2)
            ro["actions"], ro["masks"] = g["actions"].reshape(T, N, HW, len(nvec)), g["mask"].reshape(T, N, HW, -1)
        ro["logprobs"] = np.zeros((T, N), np.float32)
        hp = case["hp"]
        pol = _RefPolicy(net, kind, nvec, HW, gates)

        def ref_rollout(gamma, gae_lambda):
            return RefVecRollout(torch.device("cpu"), ro["next_episode_starts"], ro["next_values"], ro["obs"],
                                 ro["actions"], ro["rewards"], ro["episode_starts"], ro["values"], ro["logprobs"],
                                 ro["masks"], gamma, gae_lambda, subaction_mask=gates,
                                 action_plane_space=_ref_shim.MultiDiscrete(nvec) if kind == "gridnet" else None)

        as_list = lambda x: x.tolist() if isinstance(x, np.ndarray) else x
        writer = _Writer()
        algo = RefA2C(pol, torch.device("cpu"), writer, learning_rate=hp.learning_rate, gamma=as_list(hp.gamma),
                      gae_lambda=as_list(hp.gae_lambda), ent_coef=hp.ent_coef, vf_coef=as_list(hp.vf_coef) if not np.isscalar(hp.vf_coef) else hp.vf_coef,
                      max_grad_norm=hp.max_grad_norm, rms_prop_eps=hp.rms_prop_eps, use_rms_prop=hp.use_rms_prop,
                      normalize_advantage=hp.normalize_advantage,
                      multi_reward_weights=list(hp.multi_reward_weights) if hp.multi_reward_weights is not None else None,
                      gradient_accumulation=hp.gradient_accumulation, num_minibatches=hp.num_minibatches)
        torch.manual_seed(seed + 100)
        algo.learn(T * N, _Gen(ref_rollout, N))
        final = {k: v.detach().clone() for k, v in net.state_dict().items()}

        net2 = make_net()
        net2.load_state_dict(init)
        opol = olearn.OraclePolicy(net2, kind, nvec, HW, gates)
        opt = (torch.optim.RMSprop(net2.parameters(), lr=hp.learning_rate, eps=hp.rms_prop_eps) if hp.use_rms_prop
               else torch.optim.Adam(net2.parameters(), lr=hp.learning_rate))
        torch.manual_seed(seed + 100)
        ostats = olearn.a2c_learn_iteration(opol, opt, ro, hp)
        for k, v in net2.state_dict().items():
            exact(v, final[k], f"a2c {name} param {k}")
        for k in ("loss", "pi_loss", "entropy_loss", "explained_var"):
            exact(np.float64(ostats[k]), np.float64(writer.scalars[f"losses/{k}"]), f"a2c {name} stat {k}")
        out = {f"init.{k}": v.numpy() for k, v in init.items()}
        out.update({f"final.{k}": v.numpy() for k, v in final.items()})
        for k, v in ro.items():
            if v is not None:
                out[f"ro.{k}"] = v
        for k in ("loss", "pi_loss", "entropy_loss", "explained_var"):
            out[f"stats.{k}"] = np.float64(writer.scalars[f"losses/{k}"])
        out["seed"] = np.asarray(seed)
        save("a2c_" + name, **out)


def golden_acbc():
    """One ACBC.learn iteration of the live reference (acbc/acbc.py:75-141) on the microrts A2C rollout shape."""
    from tests.golden.make_golden_cases import A2C_CASES, make_net_for

    case = A2C_CASES["microrts"]
    seed = 60
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    make_net = make_net_for(case)
    net = make_net()
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    T, N, nvec, side, gates = case["T"], case["N"], case["nvec"], case["side"], case["gates"]
    HW = side * side
    ro = gae_inputs(seed + 1, T, N, 1, 0.1)
    ro["obs"] = rng.standard_normal((T, N) + case["obs_shape"], dtype=np.float32)
    g = gridnet_inputs(seed + 2, T * N, HW, nvec, 0, 0.2)
    ro["actions"], ro["masks"] = g["actions"].reshape(T, N, HW, len(nvec)), g["mask"].reshape(T, N, HW, -1)
    ro["logprobs"] = np.zeros((T, N), np.float32)
    pol = _RefPolicy(net, "gridnet", nvec, HW, gates)

    def ref_rollout(gamma, gae_lambda):
        return RefVecRollout(torch.device("cpu"), ro["next_episode_starts"], ro["next_values"], ro["obs"], ro["actions"],
                             ro["rewards"], ro["episode_starts"], ro["values"], ro["logprobs"], ro["masks"], gamma,
                             gae_lambda, subaction_mask=gates, action_plane_space=_ref_shim.MultiDiscrete(nvec))

    kw = dict(learning_rate=1e-3, batch_size=8, n_epochs=2, gamma=0.99, gae_lambda=0.95, vf_coef=0.25)
    writer = _Writer()
    algo = RefACBC(pol, torch.device("cpu"), writer, **kw)
    torch.manual_seed(seed + 100)
    algo.learn(T * N, _Gen(ref_rollout, N))
    final = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net2 = make_net()
    net2.load_state_dict(init)
    opol = olearn.OraclePolicy(net2, "gridnet", nvec, HW, gates)
    opt = torch.optim.Adam(net2.parameters(), lr=kw["learning_rate"])
    torch.manual_seed(seed + 100)
    ostats = olearn.acbc_learn_iteration(opol, opt, ro, kw["batch_size"], kw["n_epochs"], kw["gamma"], kw["gae_lambda"],
                                         kw["vf_coef"])
    for k, v in net2.state_dict().items():
        exact(v, final[k], f"acbc param {k}")
    for k in ("loss", "pi_loss"):
        exact(np.float64(ostats[k]), np.float64(writer.scalars[f"losses/{k}"]), f"acbc stat {k}")
    out = {f"init.{k}": v.numpy() for k, v in init.items()}
    out.update({f"final.{k}": v.numpy() for k, v in final.items()})
    out.update({f"ro.{k}": v for k, v in ro.items() if v is not None})
    out.update({f"stats.{k}": np.float64(writer.scalars[f"losses/{k}"]) for k in ("loss", "pi_loss")})
    out.update({f"hp.{k}": np.asarray(v) for k, v in kw.items()})
    out["seed"] = np.asarray(seed)
    save("acbc_microrts", **out)


def golden_trajectories():
    """Ragged trajectories through the live reference's TrajectoryBuilder.trajectory (rollout/trajectory.py)
    and DiscreteSkipsTrajectoryBuilder.trajectory (rollout/discrete_skips_trajectory_builder.py)."""
    from rl_algo_impls.rollout.discrete_skips_trajectory_builder import DiscreteSkipsTrajectoryBuilder
    from rl_algo_impls.rollout.trajectory import TrajectoryBuilder

    from oracle.gae import discrete_skips_advantages

    rng = np.random.default_rng(33)
    out = {}
    for tag, V, gamma, lam in (("scalar", 1, 0.99, 0.95), ("heads", 3, np.array([1.0, 0.99, 0.9]), np.array([0.95, 0.9, 0.8]))):
        lengths = [1, 7, 12, 3, 30]
        shape = () if V == 1 else (V,)
        cat = {k: [] for k in ("rewards", "values", "starts", "steps", "adv", "skip_adv", "skip_rewards")}
        nxt_v, nxt_d, skip_done = [], [], []
        for L in lengths:
            rew = rng.standard_normal((L,) + shape).astype(np.float32)
            val = rng.standard_normal((L,) + shape).astype(np.float32)
            dones = rng.random(L) < 0.15
            next_values = rng.standard_normal(shape).astype(np.float32)
            tb = TrajectoryBuilder()
            for t in range(L):
                tb.add(np.zeros(2, np.float32), rew[t], bool(dones[t]), val[t], 0.0, np.zeros(1, np.int64), None)
            traj = tb.trajectory(gamma, lam, next_values=next_values)
            starts = np.concatenate([[True], dones[:-1]])
            exact(gae_advantages(rew, val, starts, np.array(dones[-1]), next_values, gamma, lam), traj.advantages,
                  f"trajectory GAE {tag} L={L}")
            # discrete skips: every kept step absorbed k - 1 skipped env steps
            steps = rng.integers(1, 5, size=L).astype(np.int32)
            done = bool(rng.random() < 0.5)
            sb = DiscreteSkipsTrajectoryBuilder()
            sb.obs = [np.zeros(2, np.float32)] * L
            sb.rewards, sb.values = list(rew), list(val)
            sb.logprobs, sb.actions, sb.action_masks = [0.0] * L, [np.zeros(1, np.int64)] * L, [None] * L
            sb.steps_elapsed, sb.done = list(steps), done
            straj = sb.trajectory(gamma, lam, next_values=None if done else next_values)
            exact(discrete_skips_advantages(rew, val, steps, done, next_values, gamma, lam), straj.advantages,
                  f"discrete skips {tag} L={L}")
            cat["rewards"].append(rew), cat["values"].append(val), cat["starts"].append(starts), cat["steps"].append(steps)
            cat["adv"].append(traj.advantages), cat["skip_adv"].append(straj.advantages)
            nxt_v.append(next_values), nxt_d.append(dones[-1]), skip_done.append(done)
        out.update({f"{tag}.{k}": np.concatenate(v) for k, v in cat.items() if v})
        out[f"{tag}.offsets"] = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        out[f"{tag}.next_values"] = np.stack(nxt_v)
        out[f"{tag}.next_starts"], out[f"{tag}.skip_done"] = np.asarray(nxt_d), np.asarray(skip_done)
        out[f"{tag}.gamma"], out[f"{tag}.gae_lambda"] = np.asarray(gamma), np.asarray(lam)
    save("trajectories", **out)


def golden_normalizers():
    """NormalizeObservation / NormalizeReward of the live reference over a few env steps."""
    from rl_algo_impls.wrappers.normalize import NormalizeObservation as RefNormObs
    from rl_algo_impls.wrappers.normalize import NormalizeReward as RefNormRew

    rng = np.random.default_rng(21)
    steps, N, D, V = 6, 64, 17, 3

    class Env:
        num_envs = N
        single_observation_space = _ref_shim.Box(-np.inf, np.inf, (D,))
        single_action_space = _ref_shim.Discrete(2)

    obs = (rng.standard_normal((steps, N, D)) * np.linspace(0.1, 30.0, D) + np.linspace(-5, 5, D)).astype(np.float32)
    ref = RefNormObs.__new__(RefNormObs)
    RefNormObs.__init__(ref, Env())
    ours = ObsNormalizer((D,))
    outs = []
    for t in range(steps):
        want = ref.normalize(obs[t])
        exact(ours.normalize(obs[t]), want, f"obs normalizer step {t}")
        outs.append(want)
    out = {"obs": obs, "obs_out": np.stack(outs), "obs_mean": ref.rms.mean, "obs_var": ref.rms.var,
           "obs_count": np.float64(ref.rms.count)}
    for tag, shape in (("scalar", ()), ("multi", (V,))):
        rew = rng.standard_normal((steps, N) + shape).astype(np.float32) * 3
        dones = rng.random((steps, N)) < 0.1
        refr = RefNormRew.__new__(RefNormRew)
        refr.num_envs = N  # the stubbed VectorWrapper base does not forward attributes
        RefNormRew.__init__(refr, Env(), gamma=0.98, shape=shape)
        oursr = RewardNormalizer(N, shape, gamma=0.98)
        routs = []
        for t in range(steps):
            want = refr.normalize(rew[t])
            refr.returns[dones[t]] = 0  # NormalizeReward.step, wrappers/normalize.py:91
            exact(oursr.step(rew[t], dones[t]), want, f"reward normalizer {tag} step {t}")
            routs.append(want)
        out.update({f"rew_{tag}": rew, f"dones_{tag}": dones, f"rew_{tag}_out": np.stack(routs),
                    f"rew_{tag}_var": np.asarray(refr.rms.var), f"rew_{tag}_mean": np.asarray(refr.rms.mean),
                    f"rew_{tag}_returns": refr.returns, f"rew_{tag}_count": np.float64(refr.rms.count)})
    save("normalizers", **out)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "learner":  # python make_golden.py learner <case> [<case> ...]
        golden_learner(set(sys.argv[2:]))
        sys.exit(0)
    golden_acbc()
    golden_trajectories()
    golden_a2c()
    golden_normalizers()
    golden_gae()
    golden_gridnet()
    golden_categorical_gaussian()
    golden_index_stream_and_misc()
    golden_learner()
    print("all fixtures written; the oracle reproduces the live reference on every one")
