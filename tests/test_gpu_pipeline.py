"""End-to-end: synthetic env -> SyncStepRolloutGenerator (HBM-resident buffers) -> VecRollout (GAE)
-> PPO.learn (gather, fused loss, optimizer) for all five BASELINE configs at reduced sizes, with a
host env (numpy contract, PCIe every step) and a device env."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (env, n_envs, n_steps, batch, policy kwargs, ppo kwargs)
    "C1": ("CartPole-v1", 8, 32, 256, {}, dict(n_epochs=3, gamma=0.98, gae_lambda=0.8, learning_rate=1e-3)),
    "C2": ("BreakoutNoFrameskip-v4", 8, 16, 64, {}, dict(n_epochs=2, clip_range=0.1, vf_coef=0.5, ent_coef=0.01)),
    "C3": ("HalfCheetah-v4", 64, 16, 256, dict(pi_hidden_sizes=[256, 256], v_hidden_sizes=[256, 256], log_std_init=-2),
           dict(n_epochs=2, clip_range=0.1, ent_coef=4e-4, vf_coef=0.581, max_grad_norm=0.8)),
    "C4": ("Microrts-16x16", 6, 16, 32, {}, dict(n_epochs=2, clip_range=0.1, clip_range_vf=0.1, ppo2_vf_coef_halving=True,
                                                 ent_coef=0.01, vf_coef=0.5)),
    "C5": ("LuxAI_S2-64x64", 4, 4, 8, dict(num_additional_critics=12),
           dict(n_epochs=2, gamma=[1.0] * 13, gae_lambda=[0.95] * 13, clip_range=0.1, ent_coef=0.01,
                vf_coef=[0.5] + [0.1] * 12, multi_reward_weights=[0.9] + [0.1 / 12] * 12, gradient_accumulation=True,
                autocast_loss=True)),
}


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("where", ["device_env", "host_env"])
def test_learn_runs(cuda, cfg, where):
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    name, n_envs, n_steps, batch, pkw, akw = CONFIGS[cfg]
    torch.manual_seed(0)
    env = make_synthetic_env(name, n_envs, seed=1, device=cuda if where == "device_env" else None, pool=3)
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask, **pkw).to(cuda)
    gen = SyncStepRolloutGenerator(policy, env, n_steps=n_steps, subaction_mask=env.spec.subaction_mask)
    algo = PPO(policy, cuda, None, batch_size=batch, **akw)
    seen = []

    class CB:
        def on_step(self, timesteps_elapsed, train_stats):
            seen.append(train_stats)
            return True

    before = [p.detach().clone() for p in policy.parameters()]
    algo.learn(2 * n_envs * n_steps, gen, callbacks=[CB()])
    assert len(seen) == 2
    for s in seen:
        assert np.isfinite(s.loss) and np.isfinite(s.pi_loss) and np.isfinite(s.entropy_loss) and np.isfinite(s.grad_norm)
        assert np.all(np.isfinite(np.asarray(s.v_loss))) and 0 <= s.clipped_frac <= 1
    assert any((a != b.detach()).any().item() for a, b in zip(before, policy.parameters())), "parameters did not move"
    assert algo.launches_last_epoch > 0
    # the rollout buffers live in HBM and were filled
    assert gen.obs.is_cuda and gen.logprobs.abs().sum().item() > 0


def test_step_and_value_numpy_contract(cuda):
    """policy.step / policy.value keep the reference's numpy contract (actor_critic.py:298-318)."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic

    env = make_synthetic_env("Microrts-16x16", 3, seed=2)
    policy = ActorCritic(env, subaction_mask=env.spec.subaction_mask).to(cuda)
    obs, _ = env.reset()
    step = policy.step(obs, action_masks=env.get_action_mask())
    assert step.a.shape == (3, 256, 7) and step.a.dtype == np.int64
    assert step.v.shape == (3,) and step.logp_a.shape == (3,) and np.all(step.logp_a <= 0)
    assert policy.value(obs).shape == (3,)
    # sampled actions are valid wherever the head has a valid entry
    mask = env.get_action_mask()
    start = 0
    for h, n in enumerate(env.spec.nvec):
        m = mask[..., start:start + n]
        has = m.any(-1)
        chosen = np.take_along_axis(m, step.a[..., h][..., None], axis=-1)[..., 0]
        assert chosen[has].all(), f"head {h} sampled a masked action"
        start += n
