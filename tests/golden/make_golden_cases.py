"""Learner-level golden cases (hyperparameters + stub trunks), shared by the fixture generator and
the tests.  Importable without the reference."""
import numpy as np

from oracle import learner as olearn
from tests.golden.stub_nets import TinyGrid, TinyMlp
from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC

LEARNER_CASES = {
    "cartpole": dict(kind="categorical", T=32, N=8, obs_shape=(4,), nvec=(2,), V=1,
                     hp=olearn.Hyper(batch_size=64, n_epochs=3, gamma=0.98, gae_lambda=0.8, clip_range=0.2,
                                     ent_coef=0.0, learning_rate=1e-3)),
    "gaussian": dict(kind="gaussian", T=16, N=8, obs_shape=(17,), nvec=(6,), V=1,
                     hp=olearn.Hyper(batch_size=48, n_epochs=2, gamma=0.98, gae_lambda=0.92, clip_range=0.1,
                                     ent_coef=4e-4, vf_coef=0.581, max_grad_norm=0.8, learning_rate=1e-3)),
    "microrts": dict(kind="gridnet", T=8, N=6, obs_shape=(5, 8, 8), nvec=MICRORTS_NVEC, side=8, gates=MICRORTS_GATES, V=1,
                     hp=olearn.Hyper(batch_size=20, n_epochs=2, gamma=0.999, gae_lambda=0.99, clip_range=0.1,
                                     clip_range_vf=0.1, ppo2_vf_coef_halving=True, ent_coef=0.01, vf_coef=0.5,
                                     learning_rate=1e-3)),
    "microrts_teacher": dict(kind="gridnet", T=8, N=6, obs_shape=(5, 8, 8), nvec=MICRORTS_NVEC, side=8, gates=MICRORTS_GATES,
                             V=1,
                             hp=olearn.Hyper(batch_size=24, n_epochs=2, gamma=0.999, gae_lambda=0.99, clip_range=0.1,
                                             clip_range_vf=0.1, ppo2_vf_coef_halving=True, ent_coef=0.01, vf_coef=0.5,
                                             learning_rate=1e-3, teacher_kl_loss_coef=0.5)),
    "cartpole_teacher_biased": dict(kind="categorical", T=16, N=8, obs_shape=(4,), nvec=(2,), V=1,
                                    hp=olearn.Hyper(batch_size=32, n_epochs=2, gamma=0.98, gae_lambda=0.8, clip_range=0.2,
                                                    learning_rate=1e-3, teacher_kl_loss_coef=1.0, teacher_unbiased=False,
                                                    teacher_loss_importance_sampling=False)),
    "lux": dict(kind="gridnet", T=6, N=4, obs_shape=(5, 8, 8), nvec=LUX_NVEC, side=8, gates=LUX_GATES, n_pick=1, V=3,
                hp=olearn.Hyper(batch_size=8, n_epochs=2, gamma=np.array([1.0, 1.0, 0.99]),
                                gae_lambda=np.array([0.95, 0.95, 0.9]), clip_range=0.1, ent_coef=0.01,
                                vf_coef=[0.5, 0.25, 0.25], multi_reward_weights=[0.6, 0.3, 0.1],
                                gradient_accumulation=True, learning_rate=1e-3)),
    "cartpole_huber": dict(kind="categorical", T=16, N=8, obs_shape=(4,), nvec=(2,), V=1,
                           hp=olearn.Hyper(batch_size=32, n_epochs=2, gamma=0.98, gae_lambda=0.8, clip_range=0.2,
                                           clip_range_vf=0.3, vf_coef=0.7, learning_rate=1e-3, vf_loss_fn="huber_loss")),
    "gaussian_l1": dict(kind="gaussian", T=8, N=8, obs_shape=(17,), nvec=(6,), V=1,
                        hp=olearn.Hyper(batch_size=32, n_epochs=2, gamma=0.98, gae_lambda=0.92, clip_range=0.1,
                                        ent_coef=4e-4, vf_coef=0.581, max_grad_norm=0.8, learning_rate=1e-3,
                                        vf_loss_fn="l1_loss")),
    # ---- round 2: the branches of ppo.py:307-355 that had no live-reference fixture ------------------------------
    # kl_cutoff (ppo.py:352-355): the stored behaviour log-probs are shifted so that approx_kl crosses the cut-off
    "microrts_kl_cutoff": dict(kind="gridnet", T=8, N=6, obs_shape=(5, 8, 8), nvec=MICRORTS_NVEC, side=8,
                               gates=MICRORTS_GATES, V=1,
                               hp=olearn.Hyper(batch_size=24, n_epochs=3, gamma=0.999, gae_lambda=0.99, clip_range=0.1,
                                               clip_range_vf=0.1, ppo2_vf_coef_halving=True, ent_coef=0.01, vf_coef=0.5,
                                               learning_rate=1e-3, kl_cutoff=2e-3, logp_shift=0.05)),
    # standardize_advantage (ppo.py:315-316)
    "cartpole_standardize": dict(kind="categorical", T=16, N=8, obs_shape=(4,), nvec=(2,), V=1,
                                 hp=olearn.Hyper(batch_size=32, n_epochs=2, gamma=0.98, gae_lambda=0.8, clip_range=0.2,
                                                 normalize_advantage=False, standardize_advantage=True,
                                                 learning_rate=1e-3)),
    # normalize_advantages_after_scaling with multi-head rewards (ppo.py:307-311)
    "lux_after_scaling": dict(kind="gridnet", T=6, N=4, obs_shape=(5, 8, 8), nvec=LUX_NVEC, side=8, gates=LUX_GATES,
                              n_pick=1, V=3,
                              hp=olearn.Hyper(batch_size=8, n_epochs=2, gamma=np.array([1.0, 1.0, 0.99]),
                                              gae_lambda=np.array([0.95, 0.95, 0.9]), clip_range=0.1, ent_coef=0.01,
                                              vf_coef=[0.5, 0.25, 0.25], multi_reward_weights=[0.6, 0.3, 0.1],
                                              normalize_advantages_after_scaling=True, learning_rate=1e-3)),
    # vf_weights (ppo.py:344-345): the per-head value losses are contracted before the batch mean; scalar vf_coef
    "lux_vf_weights": dict(kind="gridnet", T=6, N=4, obs_shape=(5, 8, 8), nvec=LUX_NVEC, side=8, gates=LUX_GATES,
                           n_pick=1, V=3,
                           hp=olearn.Hyper(batch_size=12, n_epochs=2, gamma=np.array([1.0, 1.0, 0.99]),
                                           gae_lambda=np.array([0.95, 0.95, 0.9]), clip_range=0.1, clip_range_vf=0.2,
                                           ent_coef=0.01, vf_coef=0.5, vf_weights=[1.0, 0.5, 0.25],
                                           multi_reward_weights=[0.6, 0.3, 0.1], learning_rate=1e-3)),
    # autocast_loss (ppo.py:321): a no-op on the CPU reference (shared/autocast.py:8-12 enables it on CUDA only), so
    # this fixture pins the f32 result; the GPU path under the flag runs the trunk in bf16 and is held to a bf16 bar
    "microrts_autocast": dict(kind="gridnet", T=8, N=6, obs_shape=(5, 8, 8), nvec=MICRORTS_NVEC, side=8,
                              gates=MICRORTS_GATES, V=1,
                              hp=olearn.Hyper(batch_size=24, n_epochs=2, gamma=0.999, gae_lambda=0.99, clip_range=0.1,
                                              ent_coef=0.01, vf_coef=0.5, learning_rate=1e-3, autocast_loss=True)),
}


A2C_CASES = {
    "cartpole": dict(kind="categorical", T=5, N=8, obs_shape=(4,), nvec=(2,), V=1,
                     hp=olearn.A2CHyper(learning_rate=7e-4, gamma=0.99, gae_lambda=1.0, ent_coef=0.01, vf_coef=0.5)),
    "microrts": dict(kind="gridnet", T=6, N=4, obs_shape=(5, 8, 8), nvec=MICRORTS_NVEC, side=8, gates=MICRORTS_GATES, V=1,
                     hp=olearn.A2CHyper(learning_rate=1e-3, gamma=0.99, gae_lambda=0.95, ent_coef=0.01, vf_coef=0.5,
                                        normalize_advantage=True, use_rms_prop=False, gradient_accumulation=True,
                                        num_minibatches=3)),
}


def make_net_for(case):
    if case["kind"] == "gridnet":
        n_logits = sum(case["nvec"]) + case.get("n_pick", 0)
        return lambda: TinyGrid(case["obs_shape"][0], n_logits, case["V"])
    return lambda: TinyMlp(case["obs_shape"][0], case["nvec"][0], case["V"], gaussian=case["kind"] == "gaussian")


