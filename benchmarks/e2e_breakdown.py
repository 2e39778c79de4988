"""Where one host-env rollout step goes (C4 by default): wall-clock per phase with a synchronize after each, so
GPU time is attributed to the phase that launched it.  Usage: python benchmarks/e2e_breakdown.py [C4]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rl_algo_impls_b200.configs import CONFIGS, build  # noqa: E402


def main(key="C4"):
    dev = torch.device("cuda", 0)
    cfg = CONFIGS[key]
    env, policy, gen, algo = build(cfg, dev, env_device=None, seed=1)
    for _ in range(2):
        gen.rollout(0.99, 0.95)
    T = gen.n_steps
    acc = {k: 0.0 for k in ("replay_enqueue", "gpu_step+actions_d2h", "env.step", "uploads_enqueue", "uploads_dma")}
    policy.eval()
    gen.step_count.fill_(0)
    sync = torch.cuda.synchronize
    sync()
    t_all = time.perf_counter()
    for s in range(T):
        t0 = time.perf_counter()
        gen._graph.replay()
        a = gen._graph_outputs
        t1 = time.perf_counter()
        acts = gen._env_actions(a, landed=True)
        t2 = time.perf_counter()
        next_obs, rewards, term, trunc, _ = env.step(acts)
        t3 = time.perf_counter()
        gen._set_next_obs(next_obs)
        gen._upload_masks(gen.get_action_mask())
        t4 = time.perf_counter()
        sync()
        t5 = time.perf_counter()
        for k, d in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[k] += d
    total = time.perf_counter() - t_all
    print({k: round(v / T * 1e6, 1) for k, v in acc.items()}, "us per env step; total", round(total / T * 1e6, 1))


if __name__ == "__main__":
    main(*sys.argv[1:2])
