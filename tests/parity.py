"""Parity criteria shared by the GPU tests.

The bar (BASELINE.json north_star): bit-exact for indices / mask handling, within 1e-5 relative
fp32 for advantages, returns, losses and gradients.  Two refinements make that bar well defined:

* `scale` of a tensor comparison is max|reference| (gradients of a whole tile are compared
  against the tile's largest entry, not entry by entry -- entries that cancel to ~0 have no
  meaningful relative error);
* quantities downstream of ratio = exp(logp_new - logp_old) inherit the f32 rounding of logp
  itself (|logp| ~ 20..200 => ulp 2e-6..1.5e-5), so two correct f32 implementations with
  different summation orders differ by more than 1e-5 there.  For those we evaluate the oracle a
  second time in float64 and accept the kernel when it is within 1e-5 of the f32 oracle OR at
  least as close to the float64 result as 1e-5 / 3x the f32 oracle's own distance from it.
"""
from typing import Optional

import torch

RTOL = 1e-5


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    got, want = got.detach().cpu().double().reshape(-1), want.detach().cpu().double().reshape(-1)
    if want.numel() == 0:
        return 0.0
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-300)).item()


def close(got, want, rtol: float = RTOL, atol: float = 0.0, what: str = "") -> None:
    got, want = torch.as_tensor(got).detach().cpu().double(), torch.as_tensor(want).detach().cpu().double()
    assert got.numel() == want.numel(), f"{what}: {tuple(got.shape)} vs {tuple(want.shape)}"
    if want.numel() == 0:
        return
    got, want = got.reshape(-1), want.reshape(-1)
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= rtol * scale + atol, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rtol {rtol}, atol {atol})"


def close_conditioned(got, want32, want64, rtol: float = RTOL, what: str = "") -> None:
    """For quantities downstream of exp(logp_new - logp_old); see the module docstring."""
    e32 = rel_err(torch.as_tensor(got), torch.as_tensor(want32))
    if e32 <= rtol:
        return
    e64 = rel_err(torch.as_tensor(got), torch.as_tensor(want64))
    ref = rel_err(torch.as_tensor(want32), torch.as_tensor(want64))
    assert e64 <= max(rtol, 3 * ref), (
        f"{what}: {e32:.2e} from the f32 oracle, {e64:.2e} from the f64 oracle (the f32 oracle itself: {ref:.2e})"
    )
