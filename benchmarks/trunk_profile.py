"""Kernel-time table of one trunk forward + backward of a config's policy at its minibatch size (torch.profiler, CUPTI).
Usage: python benchmarks/trunk_profile.py [C5] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rl_algo_impls_b200.configs import CONFIGS, build  # noqa: E402


def main(key="C5", batch=None):
    dev = torch.device("cuda", 0)
    cfg = CONFIGS[key]
    env, policy, gen, algo = build(cfg, dev, env_device=dev, seed=1, n_envs=8)
    B = int(batch) if batch else cfg.algo["batch_size"]
    net = policy.network if hasattr(policy, "network") else policy
    shape = tuple(gen.next_obs.shape[1:])
    obs = (torch.rand((B,) + shape, device=dev) < 0.1).float()
    autocast = bool(cfg.algo.get("autocast_loss"))

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = policy.head_outputs(obs)
            loss = out.pi.float().square().mean() + out.values.float().square().mean()
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))


if __name__ == "__main__":
    main(*sys.argv[1:3])
