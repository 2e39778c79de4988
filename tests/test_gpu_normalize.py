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


class _ScriptedDeviceEnv:
    """Replays (reward, terminations, truncations, infos) tuples as CUDA tensors."""

    def __init__(self, script, device):
        self.script, self.t, self.device = script, 0, device
        self.num_envs = script[0][0].shape[0]

    def step(self, action):
        out = self.script[self.t]
        self.t += 1
        return (None,) + out


@pytest.mark.parametrize("case", ["lux_like", "all_episode_end", "scalar_multiplier", "multi_base"])
def test_info_rewards_wrapper_vs_reference_fixture(cuda, case):
    """K7 behind InfoRewardsWrapper (device env) vs the live reference's wrapper: bit-exact [N, V0 + K] rewards."""
    from rl_algo_impls_b200.wrappers import InfoRewardsWrapper
    from tests.test_oracle_golden import load

    z = load("info_rewards")
    g = lambda k: z[f"{case}.{k}"]
    K = g("series").shape[1]
    paths = [["stats", f"s{k}"] if k % 2 else [f"s{k}"] for k in range(K)]
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    script = []
    for t in range(g("base").shape[0]):
        infos = {"stats": {}}
        for k, path in enumerate(paths):
            (infos["stats"] if len(path) == 2 else infos)[path[-1]] = to(g("series")[t, k])
        script.append((to(g("base")[t]), to(g("terminations")[t]), to(g("truncations")[t]), infos))
    mult = g("multiplier").tolist() if f"{case}.multiplier" in z else None
    env = InfoRewardsWrapper(_ScriptedDeviceEnv(script, cuda), paths, episode_end=g("episode_end").tolist(), multiplier=mult)
    for t in range(len(script)):
        _, rewards, _, _, _ = env.step(None)
        np.testing.assert_array_equal(rewards.cpu().numpy(), g("rewards")[t])


def test_reward_assemble_rejects_bad_arguments(cuda):
    from rl_algo_impls_b200 import ops

    base = torch.zeros(4, device=cuda)
    with pytest.raises(ValueError):
        ops.reward_assemble(base, [torch.zeros(3, device=cuda)], None, None, [False])
    with pytest.raises(Exception):  # episode-end gating without the terminations flags
        ops.reward_assemble(base, [torch.zeros(4, device=cuda)], None, None, [True])
    out = ops.reward_assemble(base, [], None, None, [])
    assert out.shape == (4, 1)


@pytest.mark.parametrize("tag", ["scalar", "multi"])
def test_ema_reward_normalizer_vs_reference_fixture(cuda, tag):
    """NormalizeReward(exponential_moving_mean_var=True) on a device env vs the live reference: normalised rewards
    1e-6, float64 moments 1e-10 (the kernel's pow / summation order differ from numpy's in the last bits)."""
    from rl_algo_impls_b200.wrappers import NormalizeReward

    z = load("normalizers_ema")
    rew, dones = z[f"{tag}.rewards"], z[f"{tag}.dones"]
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    script = [(to(rew[t]), to(dones[t]), torch.zeros(rew.shape[1], dtype=torch.bool, device=cuda), {}) for t in range(rew.shape[0])]
    env = NormalizeReward(_ScriptedDeviceEnv(script, cuda), gamma=0.98, shape=tuple(rew.shape[2:]),
                          exponential_moving_mean_var=True, emv_window_size=float(z[f"{tag}.window"]))
    for t in range(rew.shape[0]):
        _, out, _, _, _ = env.step(None)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"{tag}.out"][t], rtol=1e-6, atol=1e-6)
    # scalar rewards: the reference's moving moments are per-env vectors [N] (its update broadcasts against the 1-D batch)
    flat = lambda a: np.asarray(a, np.float64).reshape(-1)
    np.testing.assert_allclose(flat(env.rms.var.cpu().numpy()), flat(z[f"{tag}.var"]), rtol=1e-10)
    np.testing.assert_allclose(flat(env.rms.emmv.mean.cpu().numpy()), flat(z[f"{tag}.ema_mean"]), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(flat(env.rms.emmv.var.cpu().numpy()), flat(z[f"{tag}.ema_var"]), rtol=1e-10)
    np.testing.assert_allclose(flat(env.rms.rms.var.cpu().numpy()), flat(z[f"{tag}.rms_var"]), rtol=1e-12)
    np.testing.assert_allclose(env.returns.cpu().numpy(), z[f"{tag}.returns"], rtol=1e-12, atol=1e-12)
    assert env.rms.rms.count == float(z[f"{tag}.count"]) and env.rms.emmv.initialized
