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


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] == "events"):
    main(*sys.argv[1:2])


def events(key="C4"):
    """The same loop with no extra synchronisation: CUDA events around the upload block and the graph replay of every
    env step -> device-side time of each and the idle gap between a replay's end and the next upload's start."""
    dev = torch.device("cuda", 0)
    cfg = CONFIGS[key]
    env, policy, gen, algo = build(cfg, dev, env_device=None, seed=1)
    for _ in range(2):
        gen.rollout(0.99, 0.95)
    T = gen.n_steps
    policy.eval()
    gen.step_count.fill_(0)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(T)]
    cpu = {"replay": 0.0, "wait_actions": 0.0, "env.step": 0.0, "uploads_enqueue": 0.0}
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for s in range(T):
        t0 = time.perf_counter()
        ev[s][1].record()
        gen._graph.replay()
        ev[s][2].record()
        if gen._wide_actions is not None:
            gen._download_wide(gen._graph_outputs)
        t1 = time.perf_counter()
        acts = gen._env_actions(gen._graph_outputs, landed=True)
        t2 = time.perf_counter()
        next_obs, rewards, term, trunc, _ = env.step(acts)
        t3 = time.perf_counter()
        if s + 1 < T:
            ev[s + 1][0].record()
        gen._upload_env_outputs(next_obs, gen.get_action_mask())
        t4 = time.perf_counter()
        for k, d in zip(cpu, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            cpu[k] += d
    torch.cuda.synchronize()
    total = time.perf_counter() - t_all
    up = sum(ev[s][0].elapsed_time(ev[s][1]) for s in range(1, T)) / (T - 1) * 1e3
    graph = sum(ev[s][1].elapsed_time(ev[s][2]) for s in range(T)) / T * 1e3
    gap = sum(ev[s][2].elapsed_time(ev[s + 1][0]) for s in range(T - 1)) / (T - 1) * 1e3
    print({"gpu_uploads+pack": round(up, 1), "gpu_graph": round(graph, 1), "gpu_idle_gap": round(gap, 1)},
          {k: round(v / T * 1e6, 1) for k, v in cpu.items()}, "us per env step; total", round(total / T * 1e6, 1))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "events":
    events(sys.argv[1])
