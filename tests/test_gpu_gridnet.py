"""K4c parity: GridNet log-prob / entropy forward, backward and the fused PPO loss vs the
torch-CPU oracle (gridnet.py:38-193 over categorical.py:12-54, ppo.py:307-361)."""
import numpy as np
import pytest
import torch

from oracle.distributions import Gridnet, gates_from_subaction_mask
from oracle.ppo_loss import normalize_advantages, ppo_loss
from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gridnet_inputs, ppo_inputs, to_torch

pytestmark = pytest.mark.gpu

from tests.parity import close, close_conditioned  # noqa: E402


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
    atol = 4e-7 * inp["logits"].abs().max().item()  # logp = x_a - lse cancels to ~0 on near-certain rows
    close(logp_g, logp, atol=atol, what="logp")
    close(ent_g, ent, atol=atol, what="entropy")
    close(lg.grad, logits.grad, atol=atol * 0.1, what="dlogits")
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


def _oracle_fused(inp, pp, nvec, gates, HW, dtype, old_logp, V, adv_mode_kw, clip_vf, halving, weights):
    f = lambda t: t.to(dtype) if t is not None and t.is_floating_point() else t
    inp = dict(inp, logits=f(inp["logits"]))
    logits, dist, action = oracle_dist(inp, nvec, gates, HW)
    logp, ent = dist.log_prob(action), dist.entropy()
    if old_logp is None:
        old_logp = (logp.detach() + pp["old_logp_noise"]).float()
    new_values = f(pp["new_values"]).clone().requires_grad_(True)
    w = torch.tensor(weights, dtype=dtype) if weights is not None else None
    adv = normalize_advantages(f(pp["adv"]), multi_reward_weights=w, **adv_mode_kw)
    vf_coef = torch.linspace(0.5, 1.0, V, dtype=dtype) if V > 1 else torch.tensor(0.5, dtype=dtype)
    parts = ppo_loss(logp, ent, new_values, f(old_logp), adv, f(pp["old_values"]), f(pp["returns"]),
                     clip_range=0.1, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf_coef,
                     ppo2_vf_coef_halving=halving)
    parts.loss.backward()
    return dict(logp=logp.detach(), entropy=ent.detach(), parts=parts, dlogits=logits.grad, dvalues=new_values.grad,
                old_logp=old_logp, vf_coef=vf_coef)


def _fused_case(cuda, B, HW, nvec, gates, n_pick, unit_p, V, adv_mode_kw, clip_vf, halving, weights, seed=0):
    from rl_algo_impls_b200 import ops

    inp = to_torch(gridnet_inputs(31 + seed + B + HW, B, HW, nvec, n_pick, unit_p))
    pp = to_torch(ppo_inputs(seed, B, V))
    o32 = _oracle_fused(inp, pp, nvec, gates, HW, torch.float32, None, V, adv_mode_kw, clip_vf, halving, weights)
    o64 = _oracle_fused(inp, pp, nvec, gates, HW, torch.float64, o32["old_logp"], V, adv_mode_kw, clip_vf, halving,
                        weights)
    parts, parts64, old_logp = o32["parts"], o64["parts"], o32["old_logp"]

    if adv_mode_kw.get("normalize_advantages_after_scaling"):
        mode = ops.ADV_AFTER_SCALING
    elif adv_mode_kw.get("normalize_advantage", True):
        mode = ops.ADV_NORMALIZE
    elif adv_mode_kw.get("standardize_advantage"):
        mode = ops.ADV_STANDARDIZE
    else:
        mode = ops.ADV_NONE
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=o32["vf_coef"].reshape(-1).tolist(),
                     vf_halving=halving, adv_mode=mode, adv_weights=weights)
    dv = {k: (v.to(cuda) if v is not None else None) for k, v in inp.items()}
    out = ops.ppo_gridnet_loss(h, spec_of(nvec, gates, n_pick), dv["logits"], dv["mask"], dv["pick_mask"],
                               dv["actions"], dv["pick_actions"], old_logp.to(cuda), pp["adv"].to(cuda),
                               pp["old_values"].to(cuda), pp["returns"].to(cuda), pp["new_values"].to(cuda),
                               want_logp=True)
    torch.cuda.synchronize()
    stats = out.stats.cpu()
    close(out.logp, o32["logp"], what="logp")
    close(out.entropy, o32["entropy"], what="entropy")
    close_conditioned(stats[0], parts.loss, parts64.loss, what="loss")
    close_conditioned(stats[1], parts.pi_loss, parts64.pi_loss, what="pi_loss")
    close(stats[2], parts.entropy_loss, what="entropy_loss")
    close_conditioned(stats[3], torch.tensor(parts.approx_kl), torch.tensor(parts64.approx_kl), what="approx_kl")
    assert abs(stats[4].item() - parts.clipped_frac) <= 1.5 / B  # counts; a ratio on the clip edge may flip
    close(stats[5 : 5 + V], parts.v_loss.reshape(-1), what="v_loss")
    np.testing.assert_allclose(stats[5 + V : 5 + 2 * V].numpy(), np.asarray(parts.val_clipped_frac).reshape(-1),
                               atol=0.5 / B)
    close_conditioned(out.grads[0], o32["dlogits"], o64["dlogits"], what="dlogits")
    close(out.dvalues, o32["dvalues"], what="dvalues")
    S = sum(nvec)
    assert (out.grads[0].cpu()[..., :S][~inp["mask"]] == 0).all()  # bit-exact mask handling


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


def test_fused_bf16_logits_and_gradients(cuda):
    """autocast_loss: bf16 logits in, bf16 dlogits out, f32 math in between.  Against the f32 oracle on
    the same (bf16-representable) logits: scalars to 1e-5 (conditioned), dlogits to bf16 rounding."""
    from rl_algo_impls_b200 import ops

    B, HW, nvec, gates = 24, 256, MICRORTS_NVEC, MICRORTS_GATES
    inp = to_torch(gridnet_inputs(77, B, HW, nvec, 0, 0.08))
    inp["logits"] = inp["logits"].bfloat16().float()
    pp = to_torch(ppo_inputs(9, B, 1))
    kw = dict(normalize_advantage=True)
    o32 = _oracle_fused(inp, pp, nvec, gates, HW, torch.float32, None, 1, kw, 0.1, True, None)
    o64 = _oracle_fused(inp, pp, nvec, gates, HW, torch.float64, o32["old_logp"], 1, kw, 0.1, True, None)
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=[0.5], vf_halving=True)
    out = ops.ppo_gridnet_loss(h, spec_of(nvec, gates, 0), inp["logits"].to(cuda).bfloat16(), inp["mask"].to(cuda), None,
                               inp["actions"].to(cuda), None, o32["old_logp"].to(cuda), pp["adv"].to(cuda),
                               pp["old_values"].to(cuda), pp["returns"].to(cuda), pp["new_values"].to(cuda), want_logp=True)
    assert out.grads[0].dtype == torch.bfloat16
    close(out.logp, o32["logp"], what="logp")
    close_conditioned(out.stats[0].cpu(), o32["parts"].loss, o64["parts"].loss, what="loss")
    got, want = out.grads[0].float().cpu(), o32["dlogits"]
    assert ((got - want).abs() <= 2 ** -8 * want.abs() + 1e-3 * want.abs().max() * 2 ** -8).all()  # bf16: 8 bits of mantissa
    assert (got[~inp["mask"]] == 0).all()


@pytest.mark.parametrize("B,HW,nvec,gates,n_pick,V,unit_p", [
    (3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06),   # the C4 minibatch (558 MB of logits + gradients)
    (512, 4096, LUX_NVEC, LUX_GATES, 1, 13, 0.02),            # the C5 roofline shape, pick head, 13 value heads
])
def test_fused_loss_at_full_size_properties(cuda, B, HW, nvec, gates, n_pick, V, unit_p):
    """BASELINE-size launches, checked through properties that do not need a full-size oracle run:
    a slice of samples against the oracle; per-sample results independent of the batch they sit in (bit-exact);
    gradients of masked entries exactly zero; gradients linear in loss_scale (exact for a power of two);
    the total loss equal to the combination of its own reported parts; determinism across launches."""
    from benchmarks.kernels import gridnet_tensors
    from rl_algo_impls_b200 import ops

    logits, mask, pick_mask, actions, pick = gridnet_tensors(B, HW, nvec, n_pick, unit_p)
    spec = spec_of(nvec, gates, n_pick)
    g = torch.Generator(device=cuda).manual_seed(3)
    # legal actions (a masked action drives the log-prob to -inf in the reference and here alike)
    start, chosen = 0, []
    for n in nvec:
        score = torch.rand((B, HW, n), device=cuda, generator=g).masked_fill(~mask[..., start:start + n], -1.0)
        chosen.append(score.argmax(-1))
        start += n
    actions = torch.stack(chosen, -1).to(actions.dtype)
    if n_pick:
        pick = torch.rand((B, n_pick, HW), device=cuda, generator=g).masked_fill(~pick_mask, -1.0).argmax(-1)
    vs = (B,) if V == 1 else (B, V)
    rnd = lambda *s: torch.randn(*s, device=cuda, generator=g)
    adv, ov, rt, nv = rnd(*vs), rnd(*vs), rnd(*vs), rnd(*vs)
    w = np.linspace(0.2, 1.0, V).tolist() if V > 1 else None
    vf = np.linspace(0.5, 1.0, V).tolist()
    fwd_logp, fwd_ent = ops.gridnet_fwd(spec, logits, mask, pick_mask, actions, pick)
    old_logp = fwd_logp + rnd(B) * 0.05

    def run(scale, rows=slice(None)):
        h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=vf, adv_weights=w, loss_scale=scale,
                         adv_mode=ops.ADV_NONE if V == 1 else ops.ADV_NORMALIZE)
        sub = lambda t: None if t is None else t[rows].contiguous()
        return ops.ppo_gridnet_loss(h, spec, sub(logits), sub(mask), sub(pick_mask), sub(actions), sub(pick), sub(old_logp),
                                    sub(adv), sub(ov), sub(rt), sub(nv), want_logp=True)

    full = run(1.0)
    S = sum(nvec)
    # (1) the forward-only kernel and the fused kernel agree bit for bit on log-prob / entropy
    assert torch.equal(full.logp, fwd_logp) and torch.equal(full.entropy, fwd_ent)
    # (2) a slice against the oracle
    rows = slice(B // 3, B // 3 + 8)
    cpu = lambda t: None if t is None else t[rows].cpu()
    inp = dict(logits=cpu(logits), mask=cpu(mask), pick_mask=cpu(pick_mask), actions=cpu(actions).long(),
               pick_actions=cpu(pick))
    _, dist, action = oracle_dist(inp, nvec, gates, HW)
    close(full.logp[rows], dist.log_prob(action).detach(), what="logp slice")
    close(full.entropy[rows], dist.entropy().detach(), what="entropy slice")
    # (3) a sample's log-prob / entropy do not depend on the batch around it
    part = run(1.0, rows)
    assert torch.equal(part.logp, full.logp[rows]) and torch.equal(part.entropy, full.entropy[rows])
    # (4) masked entries: exactly zero gradient, everywhere
    assert int((full.grads[0][..., :S][~mask] != 0).sum().item()) == 0
    empty = ~mask.any(-1).any(-1)
    if n_pick == 0 and bool(empty.any()):
        assert float(full.logp[empty].abs().max()) == 0.0 and float(full.entropy[empty].abs().max()) == 0.0
    # (5) linear in loss_scale, deterministic
    half, again = run(0.5), run(1.0)
    assert torch.equal(half.grads[0] * 2, full.grads[0]) and torch.equal(half.dvalues * 2, full.dvalues)
    assert torch.equal(again.grads[0], full.grads[0]) and torch.equal(again.stats, full.stats)
    # (6) the reported total is the combination of the reported parts (ppo.py:357-361)
    st = full.stats.cpu().double()
    total = st[1] + 0.01 * st[2] + sum(vf[v] * st[5 + v] for v in range(V))
    assert abs(float(total - st[0])) <= 1e-5 * max(1.0, abs(float(st[0])))


@pytest.mark.parametrize("nvec,gates,n_pick,HW,p", [(MICRORTS_NVEC, MICRORTS_GATES, 0, 256, 0.02),
                                                    (LUX_NVEC, LUX_GATES, 1, 1024, 0.002)])
def test_fused_with_masked_chosen_actions(cuda, nvec, gates, n_pick, HW, p):
    """The chosen action lands on a MASKED entry of a head that has valid entries (gridnet.cu: `da == kF32Lowest`;
    categorical.py:25-36 keeps the finfo.min logit).  The new log-prob is finfo.min (-inf once two meet in a sample's
    sum), the ratio against a finite behaviour log-prob is exactly 0, the policy gradient of that sample vanishes
    and its entropy / value terms are untouched: every output equals the oracle's, non-finite values included."""
    from rl_algo_impls_b200 import ops

    B, V = 24, 1
    inp = to_torch(gridnet_inputs(91 + HW, B, HW, nvec, n_pick, 0.05, masked_action_p=p))
    pp = to_torch(ppo_inputs(9, B, V))
    logits, dist, action = oracle_dist(inp, nvec, gates, HW)
    logp, ent = dist.log_prob(action), dist.entropy()
    hit = ~torch.isfinite(logp.detach()) | (logp.detach() < -1e30)
    assert hit.any() and not hit.all(), "the inputs are meant to mix regular and masked-chosen samples"
    # a finite behaviour log-prob everywhere (the stored one of a masked-chosen sample cannot be finfo.min in practice)
    old_logp = (torch.where(hit, torch.full_like(logp, -60.0), logp.detach()) + pp["old_logp_noise"]).float()
    nv = pp["new_values"].clone().requires_grad_(True)
    parts = ppo_loss(logp, ent, nv, old_logp, normalize_advantages(pp["adv"]), pp["old_values"], pp["returns"],
                     clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=torch.tensor(0.5), ppo2_vf_coef_halving=True)
    parts.loss.backward()
    assert torch.isfinite(logits.grad).all() and torch.isfinite(parts.loss)
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=[0.5], vf_halving=True)
    dv = {k: (v.to(cuda) if v is not None else None) for k, v in inp.items()}
    out = ops.ppo_gridnet_loss(h, spec_of(nvec, gates, n_pick), dv["logits"], dv["mask"], dv["pick_mask"], dv["actions"],
                               dv["pick_actions"], old_logp.to(cuda), pp["adv"].to(cuda), pp["old_values"].to(cuda),
                               pp["returns"].to(cuda), pp["new_values"].to(cuda), want_logp=True)
    torch.cuda.synchronize()
    got_logp = out.logp.cpu()
    assert torch.equal(got_logp[hit], logp.detach()[hit])  # finfo.min / -inf, exactly
    close(got_logp[~hit], logp.detach()[~hit], what="logp of the regular samples")
    close(out.entropy, ent.detach(), what="entropy")
    stats = out.stats.cpu()
    close(stats[0], parts.loss, rtol=3e-5, what="loss")
    close(stats[1], parts.pi_loss, rtol=3e-5, what="pi_loss")
    close(stats[2], parts.entropy_loss, what="entropy_loss")
    if np.isfinite(parts.approx_kl):  # one masked-chosen sample: (0 - 1) + 3.4e38, averaged
        close(stats[3], torch.tensor(parts.approx_kl), what="approx_kl")
    else:  # several: the reference's f32 batch sum overflows to inf; the kernel sums in f64 and stays (hugely) finite
        assert stats[3].item() > 1e36
    assert abs(stats[4].item() - parts.clipped_frac) <= 1.5 / B
    close(out.grads[0], logits.grad, rtol=3e-5, what="dlogits")
    close(out.dvalues, nv.grad, what="dvalues")
    # the masked-chosen samples' policy gradient is exactly the entropy term alone: same as the oracle's rows
    S = sum(nvec)
    assert (out.grads[0].cpu()[..., :S][~inp["mask"]] == 0).all()


def _pad_rows(logits: torch.Tensor, ld: int) -> torch.Tensor:
    """[B, HW, Sp] -> [B, HW, ld] with NaN in the pad columns: a kernel that read them would poison its sums."""
    out = torch.full(logits.shape[:-1] + (ld,), float("nan"), dtype=logits.dtype, device=logits.device)
    out[..., :logits.shape[-1]] = logits
    return out


@pytest.mark.parametrize("nvec,gates,n_pick,HW,ld", [(MICRORTS_NVEC, MICRORTS_GATES, 0, 256, 80), (LUX_NVEC, LUX_GATES, 1, 4096, 32),
                                                     ((3, 5), None, 1, 100, 16)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_channel_padded_logit_rows(cuda, nvec, gates, n_pick, HW, ld, dtype):
    """logits_ld (b200rl.h): a head that emits channel-padded NHWC rows (80 wide for S = 78, 32 for S' = 29) hands
    them over as they are.  Forward, fused loss, backward and sampling give exactly the dense-row results; the pad
    columns (NaN on input here) are never read and come back as zero gradient."""
    from rl_algo_impls_b200 import ops

    B, V = 9, 1
    inp = to_torch(gridnet_inputs(3 + HW, B, HW, nvec, n_pick, 0.06), cuda)
    pp = to_torch(ppo_inputs(4, B, V), cuda)
    spec = spec_of(nvec, gates, n_pick)
    dense = inp["logits"].to(dtype)
    padded = _pad_rows(dense, ld)
    Sp = dense.shape[-1]
    lp0, en0 = ops.gridnet_fwd(spec, dense, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"])
    lp1, en1 = ops.gridnet_fwd(spec, padded, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"])
    assert torch.equal(lp0, lp1) and torch.equal(en0, en1)
    old_logp = lp0 + pp["old_logp_noise"]
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=[0.5], vf_halving=True)
    args = (inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"], old_logp, pp["adv"], pp["old_values"],
            pp["returns"], pp["new_values"])
    o0 = ops.ppo_gridnet_loss(h, spec, dense, *args)
    o1 = ops.ppo_gridnet_loss(h, spec, padded, *args)
    assert torch.equal(o0.stats, o1.stats) and torch.equal(o0.dvalues, o1.dvalues)
    assert torch.equal(o1.grads[0][..., :Sp], o0.grads[0]) and (o1.grads[0][..., Sp:] == 0).all()
    dl, de = torch.randn(B, device=cuda), torch.randn(B, device=cuda)
    g0 = ops.gridnet_bwd(spec, dense, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"], dl, de)
    g1 = ops.gridnet_bwd(spec, padded, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"], dl, de)
    assert torch.equal(g1[..., :Sp], g0) and (g1[..., Sp:] == 0).all()
    a0, p0, s0 = ops.gridnet_sample(spec, dense, inp["mask"], inp["pick_mask"], 77, 5)
    a1, p1, s1 = ops.gridnet_sample(spec, padded, inp["mask"], inp["pick_mask"], 77, 5)
    assert torch.equal(a0, a1) and torch.equal(s0, s1) and (p0 is None or torch.equal(p0, p1))


@pytest.mark.parametrize("nvec,gates,n_pick,HW,ld,B", [(MICRORTS_NVEC, MICRORTS_GATES, 0, 256, 0, 64),
                                                       (MICRORTS_NVEC, MICRORTS_GATES, 0, 256, 80, 300),
                                                       (LUX_NVEC, LUX_GATES, 1, 4096, 0, 20),
                                                       (LUX_NVEC, LUX_GATES, 1, 4096, 32, 160),
                                                       ((3, 5), None, 1, 100, 0, 33)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_persistent_gradient_buffer_equals_fresh_buffers(cuda, nvec, gates, n_pick, HW, ld, B, dtype):
    """b200rl_ppo_gridnet_loss_inplace: over a sequence of minibatches of one shape whose unit cells move around
    (dense, sparse, empty, dense again), the buffer that only has its previously written rows cleared holds, after
    every call, bit for bit what a freshly zero-filled one holds -- inside a captured graph too."""
    from rl_algo_impls_b200 import ops

    ops.clear_caches()
    spec = spec_of(nvec, gates, n_pick)
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=None, ent_coef=0.01, vf_coef=[0.5])
    static = None
    graph = None
    for step, unit_p in enumerate([0.3, 0.02, 0.0, 0.5, 0.06, 0.06, 0.9, 0.01]):
        inp = to_torch(gridnet_inputs(100 + step, B, HW, nvec, n_pick, unit_p), cuda)
        pp = to_torch(ppo_inputs(step, B, 1), cuda)
        logits = inp["logits"].to(dtype)
        if ld:
            logits = _pad_rows(logits, ld)
        old_logp = torch.full((B,), -30.0, device=cuda) + pp["old_logp_noise"]
        tensors = [logits, inp["mask"], inp["pick_mask"], inp["actions"], inp["pick_actions"], old_logp, pp["adv"],
                   pp["old_values"], pp["returns"], pp["new_values"]]
        fresh = ops.ppo_gridnet_loss(h, spec, *tensors)
        if step < 4:  # eager calls
            got = ops.ppo_gridnet_loss(h, spec, *tensors, inplace=True)
        else:  # the same buffer, driven by replays of ONE captured launch over static inputs
            if static is None:
                static = [t.clone() if t is not None else None for t in tensors]
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    ops.ppo_gridnet_loss(h, spec, *static, inplace=True)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    got = ops.ppo_gridnet_loss(h, spec, *static, inplace=True)
            for s_t, t in zip(static, tensors):
                if s_t is not None:
                    s_t.copy_(t)
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(got.grads[0], fresh.grads[0]), f"call {step} (unit density {unit_p})"
        assert torch.equal(got.stats, fresh.stats) and torch.equal(got.dvalues, fresh.dvalues)
