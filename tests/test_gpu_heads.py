"""K4a / K4b parity: categorical and Gaussian heads, forward / backward / fused PPO loss, and the
distribution-level scalar stage, vs the torch-CPU oracle."""
import numpy as np
import pytest
import torch

from oracle.distributions import MaskedLogits, gaussian_logp_entropy
from oracle.ppo_loss import normalize_advantages, ppo_loss
from tests.synth import ppo_inputs, to_torch
from tests.parity import close, close_conditioned

pytestmark = pytest.mark.gpu


def _cat_inputs(seed, B, n, masked):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, n, generator=g) * 2
    mask = None
    if masked:
        mask = torch.rand(B, n, generator=g) < 0.6
        mask[::7] = False  # fully masked rows
        mask[1::7, 0] = True
    dist = MaskedLogits(logits, mask)
    actions = dist.sample(g)
    return logits, mask, actions


@pytest.mark.parametrize("B,n,masked", [(256, 2, False), (1024, 4, False), (333, 6, True), (64, 49, True)])
def test_categorical_fwd_bwd(cuda, B, n, masked):
    from rl_algo_impls_b200 import ops

    logits, mask, actions = _cat_inputs(B + n, B, n, masked)
    lg = logits.clone().requires_grad_(True)
    dist = MaskedLogits(lg, mask)
    logp, ent = dist.log_prob(actions), dist.entropy()
    g = torch.Generator().manual_seed(1)
    dl, de = torch.randn(B, generator=g), torch.randn(B, generator=g)
    (logp * dl + ent * de).sum().backward()
    lgc = logits.to(cuda).requires_grad_(True)
    logp_g, ent_g = ops.categorical_logp_entropy(lgc, mask.to(cuda) if masked else None, actions.to(cuda))
    (logp_g * dl.to(cuda) + ent_g * de.to(cuda)).sum().backward()
    close(logp_g, logp, what="logp")
    close(ent_g, ent, what="entropy")
    close(lgc.grad, lg.grad, what="dlogits")
    if masked:
        assert (lgc.grad.cpu()[~mask] == 0).all()


@pytest.mark.parametrize("B,n,clip_vf,norm", [(256, 2, None, True), (256, 4, None, True), (1000, 4, 0.2, False)])
def test_fused_categorical(cuda, B, n, clip_vf, norm):
    from rl_algo_impls_b200 import ops

    logits, mask, actions = _cat_inputs(5 + B, B, n, False)
    pp = to_torch(ppo_inputs(B, B, 1))
    lg = logits.clone().requires_grad_(True)
    dist = MaskedLogits(lg, None)
    with torch.no_grad():
        old_logp = dist.log_prob(actions) + pp["old_logp_noise"]
    nv = pp["new_values"].clone().requires_grad_(True)
    adv = normalize_advantages(pp["adv"], normalize_advantage=norm)
    parts = ppo_loss(dist.log_prob(actions), dist.entropy(), nv, old_logp, adv, pp["old_values"], pp["returns"],
                     clip_range=0.2, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=torch.tensor(0.5))
    parts.loss.backward()
    h = ops.PpoHyper(clip_range=0.2, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=[0.5],
                     adv_mode=ops.ADV_NORMALIZE if norm else ops.ADV_NONE)
    out = ops.ppo_categorical_loss(h, logits.to(cuda), None, actions.to(cuda), old_logp.to(cuda), pp["adv"].to(cuda),
                                   pp["old_values"].to(cuda), pp["returns"].to(cuda), pp["new_values"].to(cuda))
    stats = out.stats.cpu()
    close(stats[0], parts.loss, what="loss")
    close(stats[1], parts.pi_loss, what="pi_loss")
    close(stats[2], parts.entropy_loss, what="entropy_loss")
    assert abs(stats[3].item() - parts.approx_kl) <= 1e-5 * max(abs(parts.approx_kl), 1e-3)
    assert abs(stats[4].item() - parts.clipped_frac) < 0.5 / B
    close(stats[5:6], parts.v_loss.reshape(-1), what="v_loss")
    close(out.grads[0], lg.grad, what="dlogits")
    close(out.dvalues, nv.grad, what="dvalues")


def _gauss_oracle(mu, log_std, actions, pp, old_logp, dtype):
    f = lambda t: t.to(dtype)
    mu_r, ls_r = f(mu).clone().requires_grad_(True), f(log_std).clone().requires_grad_(True)
    nv = f(pp["new_values"]).clone().requires_grad_(True)
    logp, ent = gaussian_logp_entropy(mu_r, ls_r, f(actions))
    if old_logp is None:
        old_logp = (logp.detach() + pp["old_logp_noise"]).float()
    adv = normalize_advantages(f(pp["adv"]), normalize_advantage=True)
    parts = ppo_loss(logp, ent, nv, f(old_logp), adv, f(pp["old_values"]), f(pp["returns"]), clip_range=0.1,
                     clip_range_vf=None, ent_coef=4e-4, vf_coef=torch.tensor(0.581, dtype=dtype))
    parts.loss.backward()
    return dict(parts=parts, logp=logp.detach(), ent=ent.detach(), dmu=mu_r.grad, dls=ls_r.grad, dv=nv.grad,
                old_logp=old_logp)


@pytest.mark.parametrize("B,D", [(64, 6), (16384, 6), (100, 17)])
def test_fused_gaussian(cuda, B, D):
    from rl_algo_impls_b200 import ops

    g = torch.Generator().manual_seed(B + D)
    mu = torch.randn(B, D, generator=g)
    log_std = torch.full((D,), -2.0) + 0.1 * torch.randn(D, generator=g)
    actions = mu + torch.exp(log_std) * torch.randn(B, D, generator=g)
    pp = to_torch(ppo_inputs(B + 1, B, 1))
    o32 = _gauss_oracle(mu, log_std, actions, pp, None, torch.float32)
    o64 = _gauss_oracle(mu, log_std, actions, pp, o32["old_logp"], torch.float64)
    old_logp, parts, parts64 = o32["old_logp"], o32["parts"], o64["parts"]
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=None, ent_coef=4e-4, vf_coef=[0.581])
    out = ops.ppo_gaussian_loss(h, mu.to(cuda), log_std.to(cuda), actions.to(cuda), old_logp.to(cuda),
                                pp["adv"].to(cuda), pp["old_values"].to(cuda), pp["returns"].to(cuda),
                                pp["new_values"].to(cuda))
    stats = out.stats.cpu()
    close_conditioned(stats[0], parts.loss, parts64.loss, what="loss")
    close_conditioned(stats[1], parts.pi_loss, parts64.pi_loss, what="pi_loss")
    close(stats[2], parts.entropy_loss, what="entropy_loss")
    close_conditioned(out.grads[0], o32["dmu"], o64["dmu"], what="dmu")
    close_conditioned(out.grads[1], o32["dls"], o64["dls"], what="dlog_std")
    close(out.dvalues, o32["dv"], what="dvalues")
    lp_g, ent_g = ops.gaussian_logp_entropy(mu.to(cuda), log_std.to(cuda), actions.to(cuda))
    close(lp_g, o32["logp"], what="logp")
    close(ent_g, o32["ent"], what="entropy")


def test_scalar_stage_and_kl_cutoff(cuda):
    """Distribution-level path: (new_logp, entropy) in, (dlogp, dentropy, dvalues, stats) out;
    the KL cut-off zeroes pi_coef on the device and stays sticky (ppo.py:279,354-355)."""
    from rl_algo_impls_b200 import ops

    B, V = 512, 3
    pp = to_torch(ppo_inputs(8, B, V))
    g = torch.Generator().manual_seed(2)
    new_logp = torch.randn(B, generator=g).requires_grad_(True)
    entropy = torch.rand(B, generator=g).requires_grad_(True)
    old_logp = new_logp.detach() + pp["old_logp_noise"] * 3
    nv = pp["new_values"].clone().requires_grad_(True)
    w = torch.tensor([0.5, 0.3, 0.2])
    adv = normalize_advantages(pp["adv"], multi_reward_weights=w)
    vf = torch.tensor([0.5, 0.25, 0.25])
    for kl_cutoff in (None, 1e-4):
        for t in (new_logp, entropy, nv):
            t.grad = None
        parts = ppo_loss(new_logp, entropy, nv, old_logp, adv, pp["old_values"], pp["returns"], clip_range=0.1,
                         clip_range_vf=0.1, ent_coef=0.01, vf_coef=vf, kl_cutoff=kl_cutoff)
        parts.loss.backward()
        h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=vf.tolist(),
                         adv_weights=w.tolist())
        state = torch.ones(1, device=cuda)
        out = ops.ppo_scalar_loss(h, new_logp.detach().to(cuda), entropy.detach().to(cuda), old_logp.to(cuda),
                                  pp["adv"].to(cuda), pp["old_values"].to(cuda), pp["returns"].to(cuda),
                                  pp["new_values"].to(cuda), kl_cutoff=kl_cutoff,
                                  pi_coef_state=state if kl_cutoff is not None else None)
        stats = out.stats.cpu()
        close(stats[0], parts.loss, what="loss")
        close(out.grads[0], new_logp.grad if new_logp.grad is not None else torch.zeros(B), what="dlogp")
        close(out.grads[1], entropy.grad, what="dentropy")
        close(out.dvalues, nv.grad, what="dvalues")
        if kl_cutoff is not None:
            assert parts.pi_coef == 0 and state.item() == 0.0


@pytest.mark.parametrize("name", ["mse_loss", "huber_loss", "smooth_l1_loss", "l1_loss"])
@pytest.mark.parametrize("clip_vf", [None, 0.3])
def test_value_loss_functions(cuda, name, clip_vf):
    """vf_loss_fn = getattr(F, name) (ppo.py:186,331-343): value loss, its max with the clipped variant and
    dvalues of the fused scalar stage vs the oracle evaluated with torch.nn.functional."""
    import torch.nn.functional as F

    from rl_algo_impls_b200 import ops

    B, V = 1024, 2
    pp = to_torch(ppo_inputs(11, B, V))
    g = torch.Generator().manual_seed(5)
    new_logp, entropy = torch.randn(B, generator=g), torch.rand(B, generator=g)
    old_logp = new_logp + pp["old_logp_noise"]
    nv = (pp["new_values"] * 2).clone().requires_grad_(True)  # |new - returns| on both sides of delta = 1
    w, vf = torch.tensor([0.7, 0.3]), torch.tensor([0.5, 0.25])
    adv = normalize_advantages(pp["adv"], multi_reward_weights=w)
    parts = ppo_loss(new_logp, entropy, nv, old_logp, adv, pp["old_values"], pp["returns"], clip_range=0.1,
                     clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf, vf_loss_fn=getattr(F, name))
    parts.loss.backward()
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf.tolist(), adv_weights=w.tolist(),
                     vf_loss=ops.VF_LOSSES[name])
    out = ops.ppo_scalar_loss(h, new_logp.to(cuda), entropy.to(cuda), old_logp.to(cuda), pp["adv"].to(cuda),
                              pp["old_values"].to(cuda), pp["returns"].to(cuda), nv.detach().to(cuda))
    stats = out.stats.cpu()
    close(stats[0], parts.loss, what="loss")
    close(stats[5:5 + V], parts.v_loss, what="v_loss")
    close(out.dvalues, nv.grad, what="dvalues")
