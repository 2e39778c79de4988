"""Kernel-time table of a few device-env rollout steps of a config (torch.profiler, CUPTI), eager (no graph replay).
Usage: python benchmarks/rollout_profile.py [C5] [n_envs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rl_algo_impls_b200.configs import CONFIGS, build  # noqa: E402


def main(key="C5", n_envs=None):
    dev = torch.device("cuda", 0)
    cfg = CONFIGS[key]
    env, policy, gen, algo = build(cfg, dev, env_device=dev, seed=1, n_envs=int(n_envs) if n_envs else None)
    gen.cuda_graph = False
    policy.eval()
    autocast = bool(cfg.algo.get("autocast_loss"))
    with torch.no_grad():
        for _ in range(3):
            gen._device_env_step()
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                gen._device_env_step()
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))


if __name__ == "__main__":
    main(*sys.argv[1:3])
