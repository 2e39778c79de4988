"""Numerical-conditioning report for the fused GridNet PPO loss: for each quantity, the distance
of (a) the CUDA kernel and (b) the f32 oracle from the same oracle evaluated in float64.
`ours64` <= a small multiple of `ref64` means the kernel is as accurate as the reference itself."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.distributions import Gridnet, gates_from_subaction_mask  # noqa: E402
from oracle.ppo_loss import normalize_advantages, ppo_loss  # noqa: E402
from rl_algo_impls_b200 import ops  # noqa: E402
from tests.synth import (LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC, gridnet_inputs, ppo_inputs,  # noqa: E402
                         to_torch)


def oracle(inp, pp, nvec, gates, HW, dtype, old_logp=None, V=1, weights=None, clip_vf=0.1, halving=True):
    f = lambda t: t.to(dtype) if t is not None and t.is_floating_point() else t
    masks = inp["mask"] if inp["pick_mask"] is None else {"per_position": inp["mask"], "pick_position": inp["pick_mask"]}
    action = inp["actions"] if inp["pick_actions"] is None else {
        "per_position": inp["actions"], "pick_position": inp["pick_actions"]}
    logits = f(inp["logits"]).clone().requires_grad_(True)
    dist = Gridnet(HW, nvec, logits, masks, gates_from_subaction_mask(gates))
    logp, ent = dist.log_prob(action), dist.entropy()
    if old_logp is None:
        old_logp = (logp.detach() + pp["old_logp_noise"]).float()
    nv = f(pp["new_values"]).clone().requires_grad_(True)
    w = torch.tensor(weights, dtype=dtype) if weights is not None else None
    adv = normalize_advantages(f(pp["adv"]), multi_reward_weights=w)
    vf = torch.linspace(0.5, 1.0, V, dtype=dtype) if V > 1 else torch.tensor(0.5, dtype=dtype)
    parts = ppo_loss(logp, ent, nv, f(old_logp), adv, f(pp["old_values"]), f(pp["returns"]), clip_range=0.1,
                     clip_range_vf=clip_vf, ent_coef=0.01, vf_coef=vf, ppo2_vf_coef_halving=halving)
    parts.loss.backward()
    return dict(logp=logp, entropy=ent, loss=parts.loss, pi_loss=parts.pi_loss, entropy_loss=parts.entropy_loss,
                v_loss=parts.v_loss.reshape(-1), approx_kl=torch.tensor(parts.approx_kl), dlogits=logits.grad,
                dvalues=nv.grad), old_logp, vf


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def case(name, B, HW, nvec, gates, n_pick, unit_p, V=1, weights=None, seed=0):
    dev = torch.device("cuda")
    inp = to_torch(gridnet_inputs(31 + seed + B + HW, B, HW, nvec, n_pick, unit_p))
    pp = to_torch(ppo_inputs(seed, B, V))
    o32, old_logp, vf = oracle(inp, pp, nvec, gates, HW, torch.float32, V=V, weights=weights)
    o64, _, _ = oracle(inp, pp, nvec, gates, HW, torch.float64, old_logp=old_logp, V=V, weights=weights)
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=vf.reshape(-1).tolist(),
                     vf_halving=True, adv_weights=weights)
    dv = {k: (v.to(dev) if v is not None else None) for k, v in inp.items()}
    spec = ops.GridnetSpec.from_subaction_mask(nvec, gates, n_pick)
    out = ops.ppo_gridnet_loss(h, spec, dv["logits"], dv["mask"], dv["pick_mask"], dv["actions"], dv["pick_actions"],
                               old_logp.to(dev), pp["adv"].to(dev), pp["old_values"].to(dev), pp["returns"].to(dev),
                               pp["new_values"].to(dev), want_logp=True)
    st = out.stats.cpu()
    ours = dict(logp=out.logp, entropy=out.entropy, loss=st[0], pi_loss=st[1], entropy_loss=st[2],
                v_loss=st[5:5 + V], approx_kl=st[3], dlogits=out.grads[0], dvalues=out.dvalues)
    rows = {}
    for k in ours:
        rows[k] = dict(ours_vs_ref32=rel(ours[k], o32[k]), ours_vs_f64=rel(ours[k], o64[k]),
                       ref32_vs_f64=rel(o32[k], o64[k]))
    print(name, f"max|logp|={o64['logp'].abs().max().item():.1f}")
    for k, r in rows.items():
        print(f"   {k:13s} ours-ref32 {r['ours_vs_ref32']:.2e}  ours-f64 {r['ours_vs_f64']:.2e}  ref32-f64 {r['ref32_vs_f64']:.2e}")
    return {name: rows}


if __name__ == "__main__":
    res = {}
    res.update(case("microrts B=48 16x16", 48, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06))
    res.update(case("microrts B=3072 16x16", 3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06, seed=1))
    res.update(case("lux B=12 64x64 V=13", 12, 4096, LUX_NVEC, LUX_GATES, 1, 0.02, 13, np.linspace(0.2, 1, 13).tolist(), 2))
    res.update(case("dense 8x8 B=9", 9, 64, MICRORTS_NVEC, None, 0, 0.3, seed=4))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
