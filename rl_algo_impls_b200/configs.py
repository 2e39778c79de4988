"""The five BASELINE.json configs as (env spec, policy / rollout / algo hyperparameters).

Values are the reference's YAML entries (cited per config), restated as Python dicts whose keys are
the reference's own keyword names, so they splat into ``ActorCritic`` / ``SyncStepRolloutGenerator``
/ ``PPO`` exactly like ``runner/train.py:104-191`` does with the YAML.  Deviations forced by the
BASELINE config text (synthetic envs, scaled env counts) are stated in ``notes``.
"""
from dataclasses import dataclass
from typing import Any, Dict

_LUX_V = 13


@dataclass(frozen=True)
class RunConfig:
    key: str
    title: str
    env: str
    n_envs: int
    policy: Dict[str, Any]
    rollout: Dict[str, Any]
    algo: Dict[str, Any]
    notes: str = ""

    @property
    def n_steps(self) -> int:
        return self.rollout["n_steps"]

    @property
    def rollout_steps(self) -> int:
        return self.n_envs * self.n_steps


CONFIGS: Dict[str, RunConfig] = {
    # hyperparams/ppo.yml:1-23
    "C1": RunConfig("C1", "PPO CartPole-v1 MLP, 8 envs x 32 steps", "CartPole-v1", 8, {},
                    dict(n_steps=32),
                    dict(batch_size=256, n_epochs=20, gae_lambda=0.8, gamma=0.98, ent_coef=0.0, learning_rate=1e-3,
                         clip_range=0.2)),
    # hyperparams/ppo.yml:225-253 (_atari)
    "C2": RunConfig("C2", "PPO BreakoutNoFrameskip-v4 NatureCNN, 8 envs x 128 steps, synthetic 84x84x4 frames",
                    "BreakoutNoFrameskip-v4", 8, dict(activation_fn="relu"), dict(n_steps=128),
                    dict(batch_size=256, n_epochs=4, learning_rate=2.5e-4, clip_range=0.1, vf_coef=0.5, ent_coef=0.01)),
    # hyperparams/ppo.yml:337-359 (HalfCheetah-v4), scaled per BASELINE.json to 4096 envs x 64 steps
    "C3": RunConfig("C3", "PPO HalfCheetah-v4 Gaussian policy, 4096 synthetic envs x 64 steps", "HalfCheetah-v4", 4096,
                    dict(pi_hidden_sizes=[256, 256], v_hidden_sizes=[256, 256], activation_fn="relu", log_std_init=-2),
                    dict(n_steps=64),
                    dict(batch_size=16384, n_epochs=4, gamma=0.98, gae_lambda=0.92, ent_coef=0.000401762,
                         max_grad_norm=0.8, vf_coef=0.58096, learning_rate=2.0633e-05, clip_range=0.1),
                    notes="reference batch 64 / 20 epochs is for 1 env x 512 steps; the scaled rollout uses 16,384 / 4"),
    # hyperparams/ppo-Microrts.yml:1-27,48-99,119-125 (enc-dec variant; the `gridnet` head is broken, SURVEY note G)
    "C4": RunConfig("C4", "PPO GridNet MicroRTS 16x16, 24 envs x 512 steps, masked per-cell MultiDiscrete heads",
                    "Microrts-16x16", 24,
                    dict(activation_fn="relu", cnn_style="gridnet_encoder", actor_head_style="gridnet_decoder",
                         v_hidden_sizes=[128], subaction_mask={0: {1: 1, 2: 2, 3: 3, 4: 4, 5: 4, 6: 5}}),
                    dict(n_steps=512, subaction_mask={0: {1: 1, 2: 2, 3: 3, 4: 4, 5: 4, 6: 5}}),
                    dict(batch_size=3072, n_epochs=4, learning_rate=2.5e-4, clip_range=0.1, vf_coef=0.5, ent_coef=0.01,
                         clip_range_vf=0.1, ppo2_vf_coef_halving=True, max_grad_norm=0.5)),
    # hyperparams/ppo-LuxAI_S2.yml: the 64x64 squeeze-U-net entry (LuxAI_S2-v0-sSqnet-j512env64-80m-close-ore-ice2,
    # :2468-2517 over its anchors: policy 4,719,274 parameters, batch 128, gradient accumulation, bf16 autocast) at the
    # rollout shape BASELINE.json names (1024 envs x 32 steps, :1299-1322).  n_envs is the GLOBAL env count: R ranks
    # take 1024 / R envs each (bench.py); the reference's 256 contiguous minibatches of 128 are then the union of the
    # ranks' own minibatches, so the sharded epoch accumulates exactly the reference's gradient.
    "C5": RunConfig("C5", "PPO Lux AI S2 squeeze U-net 64x64, 1024 envs x 32 steps, envs sharded across the GPUs",
                    "LuxAI_S2-64x64", 1024,
                    dict(actor_head_style="squeeze_unet", channels_per_level=[128, 128, 128], strides_per_level=[4, 4],
                         deconv_strides_per_level=[[2, 2], [2, 2]], encoder_residual_blocks_per_level=[3, 2, 2],
                         decoder_residual_blocks_per_level=[2, 3], output_activation_fn="tanh",
                         additional_critic_activation_functions=["identity"] * (_LUX_V - 1), critic_shares_backbone=True,
                         shared_critic_head=True, critic_channels=128, subaction_mask={1: {2: 0, 3: 1, 4: 1, 5: 2}}),
                    dict(n_steps=32, subaction_mask={1: {2: 0, 3: 1, 4: 1, 5: 2}}, full_batch_off_accelerator=True),
                    dict(batch_size=128, n_epochs=2, gamma=[1.0] * _LUX_V, gae_lambda=[0.95] * _LUX_V, clip_range=0.1,
                         clip_range_vf=None, ppo2_vf_coef_halving=True, ent_coef=0.01, vf_coef=[0.2] + [0.1] * (_LUX_V - 1),
                         multi_reward_weights=[1, 0.1, 0.1, 0.1, 0.1, 0, 0.1, 0.1, 0.1, 0.1, 0.1, 0, 0], max_grad_norm=0.5,
                         learning_rate=1e-4, gradient_accumulation=True, autocast_loss=True,
                         normalize_advantages_after_scaling=False),
                    notes="global env count; bench.py --gpus R runs 1024 / R envs per GPU (strong scaling)"),
}


def build(cfg: RunConfig, device, env_device=None, seed: int = 0, n_envs=None, n_steps=None, pool: int = 4, **algo_overrides):
    """(env, policy, rollout_generator, algo) wired like runner/train.py:61-217 does from the YAML."""
    import torch

    from .envs import make_synthetic_env
    from .policy import ActorCritic
    from .ppo import PPO
    from .rollout import SyncStepRolloutGenerator

    torch.manual_seed(seed)
    env = make_synthetic_env(cfg.env, n_envs or cfg.n_envs, seed=seed, device=env_device, pool=pool)
    policy = ActorCritic(env, **cfg.policy).to(device)
    rollout_kw = dict(cfg.rollout)
    if n_steps:
        rollout_kw["n_steps"] = n_steps
    gen = SyncStepRolloutGenerator(policy, env, **rollout_kw)
    algo = PPO(policy, device, None, **{**cfg.algo, **algo_overrides})
    return env, policy, gen, algo
