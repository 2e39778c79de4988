"""K1 parity: the CUDA GAE scan vs the oracle restatement of shared/gae.py:97-124.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle.gae import gae_advantages, gae_returns
from tests.synth import gae_inputs, to_torch

pytestmark = pytest.mark.gpu


def _run(cuda, inp, gamma, lam):
    from rl_algo_impls_b200 import ops

    d = to_torch(inp, cuda)
    adv, ret = ops.gae_scan(d["rewards"], d["values"], d["episode_starts"], d["next_episode_starts"],
                            d["next_values"], gamma, lam)
    torch.cuda.synchronize()
    return adv.cpu().numpy(), ret.cpu().numpy()


CASES = [
    # T, N, V, p_start, gamma, lambda
    (32, 8, 1, 1 / 20, 0.98, 0.8),  # C1 CartPole
    (128, 8, 1, 1 / 200, 0.99, 0.95),  # C2 Atari
    (64, 4096, 1, 1 / 200, 0.98, 0.92),  # C3 HalfCheetah scaled
    (512, 24, 1, 1 / 2000, 0.999, 0.99),  # C4 MicroRTS
    (1, 5, 1, 0.5, 0.99, 0.95),  # single step, ragged lane count (scalar path)
    (7, 3, 1, 0.3, 1.0, 1.0),  # undiscounted
    (33, 1027, 1, 0.05, 0.9, 0.5),  # N % 4 != 0
]


@pytest.mark.parametrize("T,N,V,p,gamma,lam", CASES)
def test_gae_scalar_gamma_bit_exact(cuda, T, N, V, p, gamma, lam):
    inp = gae_inputs(1234 + T + N, T, N, V, p)
    want = gae_advantages(gamma=gamma, gae_lambda=lam, **inp)
    adv, ret = _run(cuda, inp, gamma, lam)
    np.testing.assert_array_equal(adv, want)
    np.testing.assert_array_equal(ret, gae_returns(want, inp["values"]))


@pytest.mark.parametrize("T,N,V", [(32, 64, 13), (32, 16, 3), (5, 7, 2)])
def test_gae_per_head_gamma_bit_exact(cuda, T, N, V):
    inp = gae_inputs(99 + V, T, N, V, 1 / 10)
    gamma = np.linspace(1.0, 0.95, V)  # Lux: gamma = 1 for every head (float64 ndarray)
    lam = np.full(V, 0.95)
    want = gae_advantages(gamma=gamma, gae_lambda=lam, **inp)
    adv, ret = _run(cuda, inp, gamma, lam)
    np.testing.assert_array_equal(adv, want)
    np.testing.assert_array_equal(ret, gae_returns(want, inp["values"]))


def test_gae_scalar_gamma_multi_head(cuda):
    inp = gae_inputs(5, 16, 12, 3, 0.1)
    want = gae_advantages(gamma=0.99, gae_lambda=0.95, **inp)
    adv, _ = _run(cuda, inp, 0.99, 0.95)
    np.testing.assert_array_equal(adv, want)


def test_gae_every_step_is_an_episode_start(cuda):
    inp = gae_inputs(6, 9, 8, 1, 1.1)  # all True: advantages = r - v
    adv, _ = _run(cuda, inp, 0.99, 0.95)
    np.testing.assert_array_equal(adv, inp["rewards"] - inp["values"])


def test_gae_large_shape_properties(cuda):
    """Roofline-size shape: check against the oracle on a slice of lanes (lanes are independent)."""
    T, N = 128, 1 << 18
    inp = gae_inputs(7, T, N, 1, 1 / 200)
    adv, ret = _run(cuda, inp, 0.99, 0.95)
    sl = slice(1000, 1512)
    sub = {k: (v[:, sl] if v.ndim == 2 else v[sl]) for k, v in inp.items()}
    np.testing.assert_array_equal(adv[:, sl], gae_advantages(gamma=0.99, gae_lambda=0.95, **sub))
    np.testing.assert_array_equal(ret, adv + inp["values"])


def test_gae_empty(cuda):
    from rl_algo_impls_b200 import ops

    z = torch.empty((0, 4), device=cuda)
    adv, ret = ops.gae_scan(z, z, torch.empty((0, 4), dtype=torch.bool, device=cuda),
                            torch.zeros(4, dtype=torch.bool, device=cuda), torch.zeros(4, device=cuda), 0.99, 0.95)
    assert adv.shape == (0, 4)


@pytest.mark.parametrize("tag", ["scalar", "heads"])
def test_segmented_gae_vs_reference_fixture(cuda, tag):
    """K1b: ragged trajectories in one launch vs the live reference's TrajectoryBuilder.trajectory
    (bit-exact) and DiscreteSkipsTrajectoryBuilder.trajectory (gamma ** steps_elapsed; float64 pow, 1e-6)."""
    from rl_algo_impls_b200 import ops
    from tests.test_oracle_golden import load

    z = load("trajectories")
    g = lambda k: z[f"{tag}.{k}"]
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g(k))).to(cuda)
    scalar = g("gamma").ndim == 0
    gamma, lam = (float(g("gamma")), float(g("gae_lambda"))) if scalar else (g("gamma"), g("gae_lambda"))
    adv, ret = ops.gae_segments(t("rewards"), t("values"), t("offsets"), t("next_starts"), t("next_values"), gamma, lam,
                                episode_starts=t("starts"))
    np.testing.assert_array_equal(adv.cpu().numpy(), g("adv"))
    np.testing.assert_array_equal(ret.cpu().numpy(), g("adv") + g("values"))
    sk, _ = ops.gae_segments(t("rewards"), t("values"), t("offsets"), t("skip_done"), t("next_values"), gamma, lam,
                             steps_elapsed=t("steps"))
    np.testing.assert_allclose(sk.cpu().numpy(), g("skip_adv"), rtol=1e-6, atol=1e-6)
