"""K6 parity: device running-moment normalisers vs the live reference's NormalizeObservation /
NormalizeReward (tests/golden/normalizers.npz).  The reference reduces float32 observations with numpy's
float32 pairwise mean / two-pass variance; the kernel accumulates in float64, so the bar is 1e-5 relative
on outputs and moments (returns, which the reference keeps in float64 too, to 1e-12)."""
import numpy as np
import pytest
import torch

from tests.parity import close
from tests.test_oracle_golden import load

pytestmark = pytest.mark.gpu


def test_obs_normalizer_vs_reference(cuda):
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.wrappers import NormalizeObservation

    z = load("normalizers")
    steps, N, D = z["obs"].shape
    env = make_synthetic_env("HalfCheetah-v4", N, device=cuda, pool=1)  # supplies spaces / device only
    norm = NormalizeObservation(env)
    for t in range(steps):
        out = norm.normalize(torch.from_numpy(z["obs"][t]).to(cuda))
        close(out, z["obs_out"][t], rtol=1e-5, atol=1e-6, what=f"normalised obs, step {t}")
    close(norm.rms.mean, z["obs_mean"], rtol=1e-6, atol=1e-9, what="running mean")
    close(norm.rms.var, z["obs_var"], rtol=1e-5, what="running var")
    assert abs(norm.rms.count - float(z["obs_count"])) < 1e-9
    # evaluation mode: moments frozen
    norm.training = False
    before = norm.rms.var.clone()
    norm.normalize(torch.from_numpy(z["obs"][0]).to(cuda))
    assert torch.equal(before, norm.rms.var)


@pytest.mark.parametrize("tag", ["scalar", "multi"])
def test_reward_normalizer_vs_reference(cuda, tag):
    from rl_algo_impls_b200 import ops

    z = load("normalizers")
    rew, dones = z[f"rew_{tag}"], z[f"dones_{tag}"]
    steps, N = rew.shape[:2]
    V = int(np.prod(rew.shape[2:])) if rew.ndim > 2 else 1
    returns = torch.zeros((N, V), dtype=torch.float64, device=cuda)
    mean = torch.zeros(V, dtype=torch.float64, device=cuda)
    var = torch.ones(V, dtype=torch.float64, device=cuda)
    count = torch.full((V,), 1e-4, dtype=torch.float64, device=cuda)
    for t in range(steps):
        out = ops.running_norm_reward(torch.from_numpy(rew[t]).to(cuda), torch.from_numpy(dones[t]).to(cuda), returns,
                                      mean, var, count, 0.98, True, 1e-8, 10.0)
        close(out, z[f"rew_{tag}_out"][t], rtol=1e-6, atol=1e-7, what=f"normalised reward, step {t}")
    close(returns.reshape(z[f"rew_{tag}_returns"].shape), z[f"rew_{tag}_returns"], rtol=1e-12, atol=1e-12, what="returns")
    close(var, np.asarray(z[f"rew_{tag}_var"]).reshape(-1), rtol=1e-9, what="running var")


def test_normalizers_inside_the_graphed_rollout(cuda):
    """C3 as the reference configures it (`normalize: true`, ppo.yml:339-340): the device wrappers sit
    in the captured env step and the moments keep moving across graph replays."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator
    from rl_algo_impls_b200.wrappers import NormalizeObservation, NormalizeReward

    env = NormalizeReward(NormalizeObservation(make_synthetic_env("HalfCheetah-v4", 32, seed=3, device=cuda, pool=4)),
                          gamma=0.98)
    policy = ActorCritic(env).to(cuda)
    gen = SyncStepRolloutGenerator(policy, env, n_steps=8)
    r1 = gen.rollout(gamma=0.98, gae_lambda=0.92)
    c1 = env.rms.count
    obs_count1 = env.env.rms.count
    r2 = gen.rollout(gamma=0.98, gae_lambda=0.92)
    assert env.rms.count > c1 and env.env.rms.count > obs_count1  # replays update the device state
    assert r2.obs.abs().max().item() <= 10.0 + 1e-6 and torch.isfinite(r2.rewards).all()
