"""The reference's own hyperparameter files splat into this repo's classes (runner/train.py:104-191 does
``ALGOS[algo](policy, device, tb_writer, **algo_hyperparams)``, ``ActorCritic(env, **policy_hyperparams)`` and
``rollout_generator_cls(policy, env, **rollout_hyperparams)``).  The YAMLs are read from the byte-identical reference copy
under oracle/_ref (oracle/make_ref.sh); nothing here imports reference code."""
import inspect
import os

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HYPER = os.path.join(ROOT, "oracle", "_ref", "rl_algo_impls", "hyperparams")

pytestmark = pytest.mark.skipif(not os.path.isdir(HYPER), reason="oracle/_ref has not been made (oracle/make_ref.sh)")

# (file, entry, synthetic env of this repo with the same spaces)
ENTRIES = [
    ("ppo.yml", "CartPole-v1", "CartPole-v1"),
    ("ppo.yml", "HalfCheetah-v4", "HalfCheetah-v4"),
    ("ppo-Microrts.yml", "Microrts-DefeatRandomEnemySparseReward-v3-enc-dec", "Microrts-16x16"),
    ("ppo-LuxAI_S2.yml", "LuxAI_S2-v0-sSqnet-j512env64-80m-close-ore-ice2", "LuxAI_S2-64x64"),
]


def _entries(name):
    return yaml.safe_load(open(os.path.join(HYPER, name)))


def _keywords(cls):
    """Named constructor keywords of `cls`, following **kwargs up the MRO (ReferenceAIRolloutGenerator forwards its
    keywords to SyncStepRolloutGenerator, like the reference class does)."""
    names = set()
    for klass in cls.__mro__:
        if "__init__" not in vars(klass):
            continue
        params = inspect.signature(klass.__init__).parameters
        names |= {n for n, p in params.items() if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)}
        if not any(p.kind == p.VAR_KEYWORD for p in params.values()):
            break
    return names


@pytest.mark.parametrize("name", ["ppo.yml", "ppo-Microrts.yml", "ppo-LuxAI_S2.yml"])
def test_every_ppo_yaml_entry_is_accepted_by_the_constructors(name):
    """Every key of every algo_hyperparams / rollout_hyperparams block is a keyword of PPO / a rollout generator."""
    from rl_algo_impls_b200 import rollout
    from rl_algo_impls_b200.ppo import PPO

    algo_kw = _keywords(PPO)
    gens = {"sync": rollout.SyncStepRolloutGenerator, "reference": rollout.ReferenceAIRolloutGenerator,
            "guided": rollout.GuidedLearnerRolloutGenerator, "guided_random": rollout.RandomGuidedLearnerRolloutGenerator}
    n = 0
    for key, entry in _entries(name).items():
        if not isinstance(entry, dict) or "algo_hyperparams" not in entry:
            continue
        n += 1
        unknown = set(entry["algo_hyperparams"]) - algo_kw
        assert not unknown, (key, unknown)
        kind = entry.get("rollout_type", "sync")
        if kind in gens and entry.get("rollout_hyperparams"):
            gen_kw = _keywords(gens[kind])
            unknown = set(entry["rollout_hyperparams"]) - gen_kw
            assert not unknown, (key, kind, unknown)
    assert n > 0


@pytest.mark.parametrize("fname,key,env_name", ENTRIES)
def test_policy_and_algo_build_from_the_yaml_entry(fname, key, env_name):
    """ActorCritic(env, **policy_hyperparams) and PPO(policy, device, None, **algo_hyperparams) from the entry as is."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO

    entry = _entries(fname)[key]
    ph = {k: v for k, v in (entry.get("policy_hyperparams") or {}).items() if k not in ("load_run_path", "load_path")}
    env = make_synthetic_env(env_name, 2, seed=0, pool=1)
    policy = ActorCritic(env, **ph)
    algo = PPO(policy, torch.device("cpu"), None, **entry["algo_hyperparams"])
    assert algo.batch_size == entry["algo_hyperparams"]["batch_size"]
    if "Lux" in key:
        assert sum(p.numel() for p in policy.parameters()) == 4719274 and policy.value_shape == (13,)
        assert np.allclose(algo.multi_reward_weights, entry["algo_hyperparams"]["multi_reward_weights"])
    if "Microrts" in key:
        assert sum(p.numel() for p in policy.parameters()) == 851727


@pytest.mark.gpu
@pytest.mark.parametrize("fname,key,env_name", ENTRIES[2:])
def test_learn_epoch_from_the_yaml_entry(cuda, fname, key, env_name):
    """The three objects wired like runner/train.py:104-217 from the YAML entry (env count, rollout length and batch
    reduced to test size) run a learn_epoch on the device path."""
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic
    from rl_algo_impls_b200.ppo import PPO
    from rl_algo_impls_b200.rollout import SyncStepRolloutGenerator

    entry = _entries(fname)[key]
    ph = {k: v for k, v in entry["policy_hyperparams"].items() if k not in ("load_run_path", "load_path")}
    torch.manual_seed(0)
    env = make_synthetic_env(env_name, 8, seed=0, device=cuda, pool=2)
    policy = ActorCritic(env, **ph).to(cuda)
    rollout_kw = dict(entry["rollout_hyperparams"], n_steps=4)
    rollout_kw.pop("num_envs_reset_every_rollout", None)  # the synthetic env has no masked_reset
    gen = SyncStepRolloutGenerator(policy, env, **rollout_kw)
    algo = PPO(policy, cuda, None, **dict(entry["algo_hyperparams"], batch_size=16))
    steps, cont = algo.learn_epoch(0, 1 << 30, gen, None)
    assert steps == 32 and cont and np.isfinite(algo.last_train_stats.loss)
