#!/usr/bin/env python
"""bench.py -- PPO env-steps/sec (rollout + GAE + update) on B200, with the roofline of the
dominant kernel and the CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl b200|reference]

One "step" is one PPO ``learn_epoch`` (one rollout of T x N env-steps through
the HBM-resident buffer, one GAE scan, n_epochs x minibatches of gather + trunk + fused loss +
optimizer step) on a synthetic environment of the named BASELINE.json config.  Reported on ONE
JSON line (rank 0):

  value      env-steps/s, whole job, device-resident env (nothing crosses PCIe in the timed region)
  e2e        the same metric through the reference-facing contract: a HOST (numpy) VectorEnv, so
             every env step uploads obs / masks / rewards and downloads the sampled actions
  roofline   the fused GridNet PPO-loss kernel (K4), CUDA events around every launch inside the
             timed region, algorithmic bytes / mean duration against MEASURED_PEAKS.json
  cpu_baseline  the oracle restatement of the reference learner timed on the host cores (bounded sample)

N > 1 (torchrun, one process per GPU): envs shard across ranks (weak scaling: every rank runs the
config's per-GPU env slice), gradients and advantage moments are all-reduced over NCCL.
``--impl reference`` times the reference's own CPU path (the oracle port; the reference is pure
Python + torch-CPU and cannot travel to the GPU box) on a bounded sample of the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "PPO env-steps/sec (rollout+GAE+update)"
UNIT = "env-steps/s"

# bounded CPU samples per config: (n_envs, n_steps, batch_size, n_epochs) -- same maps / heads / trunk,
# fewer env-steps per learn_epoch so that the CPU leg ends in tens of seconds
CPU_SAMPLE = {
    "C1": (8, 32, 256, 20),
    "C2": (8, 32, 64, 4),
    "C3": (512, 16, 2048, 4),
    "C4": (24, 32, 192, 4),
    "C5": (2, 4, 4, 2),
}


def peak_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line):
    NVML every 20 ms; falls back to one nvidia-smi query per 200 ms when NVML is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.sm_max = index, [], set(), None
        self._stop_evt = threading.Event()
        self.source = "nvml"
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml, self.source = None, "nvidia-smi"

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self._h))
        for bit, name in self.NVML_REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6 and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
            self.sm_max = float(r[1])
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v == "Active":
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self._nvml is not None else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self._nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def measured_traffic(cfg_key: str, batch: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r01/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r01", "traffic.json")
    if not os.path.exists(p):
        return None
    entry = json.load(open(p)).get(cfg_key)
    return int(entry["traffic_bytes_per_sample"] * batch) if entry else None


def loss_kernel_bytes(policy, batch: int, V: int, logits_bytes: int) -> int:
    """Algorithmic bytes of one fused-loss launch: logits read + dlogits written,
    1 B per mask element, 1 B per per-cell action, 2 B per pick index, 4*(2+5V) per-sample scalars."""
    if policy.kind != "gridnet":
        return 0
    HW, S, A, n_pick = policy.map_size, sum(policy.nvec), len(policy.nvec), policy.n_pick
    per_sample = 2 * logits_bytes * HW * (S + n_pick) + HW * S + n_pick * HW + HW * A + 2 * n_pick + 4 * (2 + 5 * V)
    return batch * per_sample


def timed_epochs(algo, gen, steps: int, world: int):
    """K learn_epochs bracketed by barrier + synchronize, CUDA events on the launching stream;
    returns the max-over-ranks elapsed ms."""
    total = gen.n_steps * gen.vec_env.num_envs
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    elapsed_steps = 0
    for _ in range(steps):
        elapsed_steps, _ = algo.learn_epoch(elapsed_steps, steps * total, gen, None)
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([start.elapsed_time(end)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def run_cpu_learner(cfg_key: str, steps: int, warmup: int, threads: int):
    """The oracle learner (CPU restatement of the reference, oracle/learner.py) on a bounded sample."""
    from oracle import learner as olearn
    from rl_algo_impls_b200.configs import CONFIGS
    from rl_algo_impls_b200.envs import make_synthetic_env
    from rl_algo_impls_b200.policy import ActorCritic

    cfg = CONFIGS[cfg_key]
    n_envs, n_steps, batch, n_epochs = CPU_SAMPLE[cfg_key]
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    env = make_synthetic_env(cfg.env, n_envs, seed=0, device=None, pool=2)
    ac = ActorCritic(env, **cfg.policy)  # the trunk (plain torch modules) on the CPU
    pol = olearn.OraclePolicy(ac.network, ac.kind, getattr(ac, "nvec", ()), getattr(ac, "map_size", 0),
                              cfg.policy.get("subaction_mask"))
    a = cfg.algo
    as_np = lambda x: np.asarray(x, dtype=np.float64) if isinstance(x, (list, tuple)) else x
    hp = olearn.Hyper(batch_size=batch, n_epochs=n_epochs, gamma=as_np(a.get("gamma", 0.99)),
                      gae_lambda=as_np(a.get("gae_lambda", 0.95)), clip_range=a.get("clip_range", 0.2),
                      clip_range_vf=a.get("clip_range_vf"), ent_coef=a.get("ent_coef", 0.0),
                      vf_coef=a.get("vf_coef", 0.5), ppo2_vf_coef_halving=a.get("ppo2_vf_coef_halving", False),
                      max_grad_norm=a.get("max_grad_norm", 0.5), multi_reward_weights=a.get("multi_reward_weights"),
                      gradient_accumulation=a.get("gradient_accumulation", False),
                      learning_rate=a.get("learning_rate", 3e-4))
    opt = torch.optim.Adam(ac.parameters(), lr=hp.learning_rate, eps=1e-7)
    state = {}

    def epoch():
        ro = olearn.collect_rollout(pol, env, n_steps, state)
        olearn.learn_epoch(pol, opt, ro, hp)

    for _ in range(warmup):
        epoch()
    t0 = time.perf_counter()
    for _ in range(steps):
        epoch()
    dt = time.perf_counter() - t0
    sample = (f"{cfg.key} shapes, {n_envs} envs x {n_steps} steps, batch {batch}, {n_epochs} epochs per learn_epoch; "
              f"{steps} learn_epochs after {warmup} warm-up")
    return n_envs * n_steps * steps / dt, dt / steps * 1e3, sample


def emit(line: dict) -> None:
    """The ONE JSON line, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries chat on stdout (NCCL prints its version there): keep fd 1 for the JSON line alone.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    from rl_algo_impls_b200.configs import CONFIGS, build

    cfg = CONFIGS[args.config]
    workload = {"workload": f"{cfg.key}: {cfg.title}", "env": cfg.env, "envs_per_gpu": cfg.n_envs,
                "n_steps": cfg.n_steps, "batch_size": cfg.algo["batch_size"], "n_epochs": cfg.algo["n_epochs"],
                "parallelism": f"env-sharded dp{args.gpus}",
                "l2": ("not flushed: every minibatch gathers fresh rows and the minibatch working set (logits + dlogits "
                       "490 MB per launch at C4, 119 MB at C5; rollout buffer 0.4-5.5 GB) exceeds the 126 MB L2"
                       if cfg.key in ("C4", "C5") else
                       "not flushed and the working set fits L2: this config is launch-bound, reported for completeness")}

    if args.impl == "reference":
        if rank != 0:
            return
        value, ms, sample = run_cpu_learner(args.config, max(1, args.steps), max(1, min(args.warmup, 2)), threads)
        emit({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rl_algo_impls_b200 import ops

    # ---- device-resident leg -------------------------------------------------------------------
    env, policy, gen, algo = build(cfg, dev, env_device=dev, seed=1234 + rank)
    # The roofline leg brackets every fused-loss launch with CUDA events, which a graph replay would
    # hide.  On the GridNet configs the update is GPU-bound on the trunk (measured: graphed 181.2 vs
    # eager 181.9 ms per C4 step), so their update runs eagerly here; the rollout stays graph-replayed.
    if policy.kind == "gridnet":
        algo.cuda_graph_update = False
    if world > 1:  # replicas start from rank 0's weights
        for p in policy.parameters():
            dist.broadcast(p.data, 0)
    for _ in range(max(3, args.warmup)):
        algo.learn_epoch(0, 1 << 40, gen, None)
    algo.profile_stages = True
    algo.learn_epoch(0, 1 << 40, gen, None)
    algo.profile_stages = False
    stages = algo.stage_ms
    timer = ops.KernelTimer(["b200rl_ppo_gridnet_loss", "b200rl_gae_scan_f32", "b200rl_gather_rows",
                             "b200rl_ppo_categorical_loss_f32", "b200rl_ppo_gaussian_loss_f32"])
    ops.set_kernel_timer(timer)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms = timed_epochs(algo, gen, args.steps, world)
    clock_info = clocks.stop()
    ops.set_kernel_timer(None)
    launches = algo.launches_last_epoch * args.steps
    kernel_ms = timer.summary()
    steps_total = cfg.rollout_steps * world * args.steps
    value = steps_total / (ms / 1e3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = peak_hbm_gbs()
    V = max(1, int(np.prod(policy.value_shape)))
    roofline = None
    if "b200rl_ppo_gridnet_loss" in kernel_ms:
        n, mean_ms = kernel_ms["b200rl_ppo_gridnet_loss"]
        nbytes = loss_kernel_bytes(policy, cfg.algo["batch_size"], V, 2 if cfg.algo.get("autocast_loss") else 4)
        achieved = nbytes / (mean_ms * 1e-3) / 1e9
        roofline = {"kernel": "gridnet_kernel<kPpo> via b200rl_ppo_gridnet_loss (+ its 1-block stats finaliser)",
                    "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": measured_traffic(cfg.key, cfg.algo["batch_size"]), "bytes_per_launch": nbytes,
                    "ms_per_launch": mean_ms, "launches_timed": n, "peak_source": peak_src,
                    "note": "mask-driven kernel: logits are read only for cells with a valid action, so DRAM "
                            "traffic is below the algorithmic bytes (DESIGN.md 4.1)"}
    elif "b200rl_gae_scan_f32" in kernel_ms:
        n, mean_ms = kernel_ms["b200rl_gae_scan_f32"]
        nbytes = cfg.rollout_steps * (16 * V + 1)
        achieved = nbytes / (mean_ms * 1e-3) / 1e9
        roofline = {"kernel": "gae_scan_kernel via b200rl_gae_scan_f32", "bound": "hbm", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "bytes_per_launch": nbytes,
                    "ms_per_launch": mean_ms, "launches_timed": n, "peak_source": peak_src}
    kernels = {k: {"launches": n, "ms_mean": m} for k, (n, m) in kernel_ms.items()}

    # ---- end-to-end leg: host env, PCIe inside the timed region ------------------------------------
    e2e = None
    if not args.no_e2e:
        del gen, env
        henv, _, hgen, halgo = build(cfg, dev, env_device=None, seed=1234 + rank)
        halgo.cuda_graph_update = algo.cuda_graph_update
        halgo.policy = policy
        hgen.policy = policy
        halgo.optimizer = algo.optimizer
        for _ in range(3):
            halgo.learn_epoch(0, 1 << 40, hgen, None)
        h2d0, d2h0 = hgen._upload.bytes, hgen.d2h_bytes
        e2e_steps = max(1, min(args.steps, 3))
        ms_e2e = timed_epochs(halgo, hgen, e2e_steps, world)
        e2e = {"value": cfg.rollout_steps * world * e2e_steps / (ms_e2e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": (hgen._upload.bytes - h2d0) // e2e_steps,
               "d2h_bytes_per_step": (hgen.d2h_bytes - d2h0) // e2e_steps + halgo.d2h_bytes_last_epoch,
               "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
               "path": "host numpy VectorEnv -> DMA out of the env buffers (page-locked in place; small fields via pinned staging) -> HBM rollout buffer; sampled actions -> host"}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, _, sample = run_cpu_learner(args.config, 2, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if not cfg.algo.get("autocast_loss") else "bf16 trunk / f32 loss math",
            "data": "synthetic", "config": workload, "clocks": clock_info, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels, "stages_ms": stages,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
