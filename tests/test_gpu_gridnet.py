"""K4c parity: GridNet log-prob / entropy forward, backward and the fused PPO loss vs the
torch-CPU oracle (gridnet.py:38-193 over categorical.py:12-54, ppo.py:307-361)."""
import numpy as np
import pytest
import torch

from oracle.distributions import Gridnet, gates_from_subaction_mask
from oracle.ppo_loss import normalize_advantages, ppo_loss
from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gridnet_inputs, ppo_inputs, to_torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: within 1e-5 relative fp32


def close(got: torch.Tensor, want: torch.Tensor, rtol=RTOL, what=""):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    scale = want.abs().max().item() if want.numel() else 0.0
    err = (got - want).abs().max().item() if want.numel() else 0.0
    assert err <= rtol * max(scale, 1e-30) + 1e-30, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


def oracle_dist(inp, nvec, gates, HW):
    masks = inp["mask"] if inp["pick_mask"] is None else {"per_position": inp["mask"], "pick_position": inp["pick_mask"]}
    action = inp["actions"] if inp["pick_actions"] is None else {
        "per_position": inp["actions"], "pick_position": inp["pick_actions"]}
    logits = inp["logits"].clone().requires_grad_(True)
    dist = Gridnet(HW, nvec, logits, masks, gates_from_subaction_mask(gates))
    return logits, dist, action


def spec_of(nvec, gates, n_pick):
    from rl_algo_impls_b200 import ops

    return ops.GridnetSpec.from_subaction_mask(nvec, gates, n_pick)


SHAPES = [
    # B, HW, nvec, gates, n_pick, unit_p
    (6, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06),  # C4 16x16
    (3, 4096, LUX_NVEC, LUX_GATES, 1, 0.02),  # C5 64x64 + pick_position (cluster path)
    (5, 64, MICRORTS_NVEC, None, 0, 0.5),  # 8x8, ungated, dense units
    (4, 100, (3, 5), None, 1, 0.3),  # odd sizes, pick head, unaligned tiles
    (7, 81, (5, 3, 2), {0: {1: 1, 2: 4}}, 0, 0.4),  # 9x9, odd S, gated
    (2, 1024, LUX_NVEC, LUX_GATES, 1, 0.0),  # no unit anywhere: every row fully masked
    (1, 1, (2,), None, 0, 1.0),  # a single cell
]


@pytest.mark.parametrize("B,HW,nvec,gates,n_pick,unit_p", SHAPES)
def test_gridnet_fwd_bwd(cuda, B, HW, nvec, gates, n_pick, unit_p):
    from rl_algo_impls_b200 import ops

    inp = to_torch(gridnet_inputs(11 + B + HW, B, HW, nvec, n_pick, unit_p, logit_scale=2.0))
    logits, dist, action = oracle_dist(inp, nvec, gates, HW)
    logp, ent = dist.log_prob(action), dist.entropy()
    g = torch.Generator().manual_seed(3)
    dlogp, dent = torch.randn(B, generator=g), torch.randn(B, generator=g)
    (logp * dlogp + ent * dent).sum().backward()

    dv = {k: (v.to(cuda) if v is not None else None) for k, v in inp.items()}
    lg = dv["logits"].clone().requires_grad_(True)
    spec = spec_of(nvec, gates, n_pick)
    logp_g, ent_g = ops.gridnet_logp_entropy(spec, lg, dv["mask"], dv["pick_mask"], dv["actions"], dv["pick_actions"])
    (logp_g * dlogp.to(cuda) + ent_g * dent.to(cuda)).sum().backward()
    close(logp_g, logp, what="logp")
    close(ent_g, ent, what="entropy")
    close(lg.grad, logits.grad, what="dlogits")
    # bit-exact mask handling: masked entries and cells without a unit get exactly zero gradient
    S = sum(nvec)
    masked = ~inp["mask"]
    assert (lg.grad.cpu()[..., :S][masked] == 0).all()
    assert (logits.grad[..., :S][masked] == 0).all()


@pytest.mark.parametrize("act_dtype", [torch.uint8, torch.int32, torch.int64])
def test_gridnet_action_dtypes(cuda, act_dtype):
    from rl_algo_impls_b200 import ops

    B, HW = 4, 256
    inp = to_torch(gridnet_inputs(21, B, HW, MICRORTS_NVEC, 0, 0.1))
    _, dist, action = oracle_dist(inp, MICRORTS_NVEC, MICRORTS_GATES, HW)
    spec = spec_of(MICRORTS_NVEC, MICRORTS_GATES, 0)
    logp_g, ent_g = ops.gridnet_fwd(spec, inp["logits"].to(cuda), inp["mask"].to(cuda), None,
                                    inp["actions"].to(cuda).to(act_dtype), None)
    close(logp_g, dist.log_prob(action), what="logp")
    close(ent_g, dist.entropy(), what="entropy")


def _fused_case(cuda, B, HW, nvec, gates, n_pick, unit_p, V, adv_mode_kw, clip_vf, halving, weights, seed=0):
    from rl_algo_impls_b200 import ops

    inp = to_torch(gridnet_inputs(31 + seed + B + HW, B, HW, nvec, n_pick, unit_p))
    pp = to_torch(ppo_inputs(seed, B, V))
    logits, dist, action = oracle_dist(inp, nvec, gates, HW)
    with torch.no_grad():
        old_logp = dist.log_prob(action) + pp["old_logp_noise"]
    new_values = pp["new_values"].clone().requires_grad_(True)
    w = torch.tensor(weights, dtype=torch.float32) if weights is not None else None
    adv = normalize_advantages(pp["adv"], multi_reward_weights=w, **adv_mode_kw)
    vf_coef = torch.linspace(0.5, 1.0, V) if V > 1 else torch.tensor(0.5)
    parts = ppo_loss(dist.log_prob(action), dist.entropy(), new_values, old_logp, adv, pp["old_values"], pp["returns"],
                     clip_range=0.1, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf_coef,
                     ppo2_vf_coef_halving=halving)
    parts.loss.backward()

    if adv_mode_kw.get("normalize_advantages_after_scaling"):
        mode = ops.ADV_AFTER_SCALING
    elif adv_mode_kw.get("normalize_advantage", True):
        mode = ops.ADV_NORMALIZE
    elif adv_mode_kw.get("standardize_advantage"):
        mode = ops.ADV_STANDARDIZE
    else:
        mode = ops.ADV_NONE
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf_coef.reshape(-1).tolist(),
                     vf_halving=halving, adv_mode=mode, adv_weights=weights)
    dv = {k: (v.to(cuda) if v is not None else None) for k, v in inp.items()}
    out = ops.ppo_gridnet_loss(h, spec_of(nvec, gates, n_pick), dv["logits"], dv["mask"], dv["pick_mask"],
                               dv["actions"], dv["pick_actions"], old_logp.to(cuda), pp["adv"].to(cuda),
                               pp["old_values"].to(cuda), pp["returns"].to(cuda), pp["new_values"].to(cuda),
                               want_logp=True)
    torch.cuda.synchronize()
    stats = out.stats.cpu()
    close(out.logp, dist.log_prob(action), what="logp")
    close(stats[0], parts.loss, what="loss")
    close(stats[1], parts.pi_loss, what="pi_loss")
    close(stats[2], parts.entropy_loss, what="entropy_loss")
    assert abs(stats[3].item() - parts.approx_kl) <= 1e-5 * max(abs(parts.approx_kl), 1e-3)
    assert abs(stats[4].item() - parts.clipped_frac) < 0.5 / B  # exact count
    close(stats[5 : 5 + V], parts.v_loss.reshape(-1), what="v_loss")
    np.testing.assert_allclose(stats[5 + V : 5 + 2 * V].numpy(), np.asarray(parts.val_clipped_frac).reshape(-1),
                               atol=0.5 / B)
    close(out.grads[0], logits.grad, what="dlogits")
    close(out.dvalues, new_values.grad, what="dvalues")
    S = sum(nvec)
    assert (out.grads[0].cpu()[..., :S][~inp["mask"]] == 0).all()


def test_fused_microrts(cuda):
    _fused_case(cuda, 48, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06, 1, dict(normalize_advantage=True), 0.1, True, None)


def test_fused_microrts_no_vclip_no_norm(cuda):
    _fused_case(cuda, 16, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06, 1,
                dict(normalize_advantage=False), None, False, None, seed=1)


def test_fused_lux_multi_head(cuda):
    V = 13
    w = np.linspace(0.2, 1.0, V).tolist()
    _fused_case(cuda, 12, 4096, LUX_NVEC, LUX_GATES, 1, 0.02, V, dict(normalize_advantage=True), None, False, w, seed=2)


def test_fused_lux_after_scaling(cuda):
    V = 3
    _fused_case(cuda, 10, 1024, LUX_NVEC, LUX_GATES, 1, 0.05, V,
                dict(normalize_advantages_after_scaling=True), 0.2, False, [0.5, 0.3, 0.2], seed=3)


def test_fused_standardize(cuda):
    _fused_case(cuda, 9, 64, MICRORTS_NVEC, None, 0, 0.3, 1,
                dict(normalize_advantage=False, standardize_advantage=True), 0.1, False, None, seed=4)


def test_gridnet_bf16_logits(cuda):
    """autocast_loss configs hand bf16 logits; the kernel computes in f32 from the bf16 values."""
    from rl_algo_impls_b200 import ops

    B, HW = 4, 256
    inp = to_torch(gridnet_inputs(41, B, HW, MICRORTS_NVEC, 0, 0.1))
    inp["logits"] = inp["logits"].bfloat16().float()  # same values both sides
    _, dist, action = oracle_dist(inp, MICRORTS_NVEC, MICRORTS_GATES, HW)
    spec = spec_of(MICRORTS_NVEC, MICRORTS_GATES, 0)
    logp_g, ent_g = ops.gridnet_fwd(spec, inp["logits"].to(cuda).bfloat16(), inp["mask"].to(cuda), None,
                                    inp["actions"].to(cuda), None)
    close(logp_g, dist.log_prob(action), what="logp")
    close(ent_g, dist.entropy(), what="entropy")
