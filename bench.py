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
  roofline   the fused GridNet PPO-loss kernel (K4), CUDA events around its launches inside the
             timed region, algorithmic bytes / mean duration against MEASURED_PEAKS.json.  The minibatch
             updates replay from CUDA graphs like the learner's default does; every 10th timed step (the
             first included) runs them eagerly so that the events have host calls to bracket
             (``update_mode``, ``roofline.launches_timed``)
  cpu_baseline  the unmodified reference learner (oracle/_ref) timed on the host cores (bounded sample)

Workloads.  N = 1 runs **C4** (GridNet MicroRTS 16x16, 24 envs x 512 steps: the largest config the reference's own CPU
path can also run at full size) and adds ``scale_base``: the C5 single-GPU point of the scaling curve.  N > 1
(torchrun, one process per GPU) runs **C5** (Lux 64x64, 1024 envs x 32 steps -- the config BASELINE.json names for the
8-GPU split) with the 1024 envs sharded 1024 / N per GPU: ``"scaling": "strong"``.  Gradients are all-reduced over
NCCL once per epoch (gradient accumulation); nothing else crosses ranks.  ``--config`` overrides the workload.

The trunk's convolutions stay cuDNN calls whose algorithms the library's autotuner picks during the warm-up steps
(``torch.backends.cudnn.benchmark``; said in the line's ``dtype``; ``B200RL_CUDNN_BENCHMARK=0`` keeps the heuristic
choice).

``--impl reference`` times the UNMODIFIED reference (``rl_algo_impls.ppo.ppo.PPO`` + ``SyncStepRolloutGenerator`` +
``ActorCritic`` from oracle/_ref, imported through oracle/ref_shim.py) on the host cores with all threads, on the same
config: C1-C4 at full size with the number of timed steps bounded by a time budget (the line's ``steps`` / ``warmup``
are what actually ran); C5 on a bounded env count stated in ``config``.
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

# The reference arm runs the full config; only C5 (50 GB of host buffers, minutes per minibatch on a CPU) is
# sampled: the same map, heads, trunk, T and batch on fewer envs (env-steps/s is per-env linear on the CPU path).
REF_ENVS = {"C5": 4}
REF_BUDGET_S = 240.0  # the reference arm stops timing new learn_epochs after this much wall time
# cpu_baseline inside the GPU arm (rank 0, N = 1): a bounded sample of the same reference learner,
# (n_envs, n_steps, batch_size) -- tens of seconds of CPU work; the full-size number is the reference arm's
CPU_SAMPLE = {"C1": (8, 32, 256), "C2": (8, 32, 64), "C3": (512, 16, 2048), "C4": (24, 64, 384), "C5": (2, 8, 16)}


CUDNN_AUTOTUNE = os.environ.get("B200RL_CUDNN_BENCHMARK", "1") != "0"
EAGER_EVERY = 10  # GridNet configs: every 10th timed step runs its update eagerly (event-timed fused-loss launches)


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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r0N/traffic.json)."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "traffic.json")
        if os.path.exists(p):
            entry = json.load(open(p)).get(cfg_key)
            if entry:
                return int(entry["traffic_bytes_per_sample"] * batch)
    return None


def loss_kernel_bytes(policy, batch: int, V: int, logits_bytes: int) -> int:
    """Algorithmic bytes of one fused-loss launch: logits read + dlogits written,
    1 B per mask element, 1 B per per-cell action, 2 B per pick index, 4*(2+5V) per-sample scalars."""
    if policy.kind != "gridnet":
        return 0
    HW, S, A, n_pick = policy.map_size, sum(policy.nvec), len(policy.nvec), policy.n_pick
    per_sample = 2 * logits_bytes * HW * (S + n_pick) + HW * S + n_pick * HW + HW * A + 2 * n_pick + 4 * (2 + 5 * V)
    return batch * per_sample


def timed_epochs(algo, gen, steps: int, world: int, eager_every: int = 0):
    """K learn_epochs bracketed by barrier + synchronize, CUDA events on the launching stream;
    returns the max-over-ranks elapsed ms.  eager_every > 0: the minibatch updates replay from CUDA graphs (the learner's
    default) except in every eager_every-th step (the first included), whose update runs eagerly so that the kernel
    timer's CUDA events bracket its launches -- a replay has no host call to bracket."""
    total = gen.n_steps * gen.vec_env.num_envs
    graphed = algo.cuda_graph_update
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    elapsed_steps = 0
    for i in range(steps):
        if eager_every:
            algo.cuda_graph_update = graphed and i % eager_every != 0
        elapsed_steps, _ = algo.learn_epoch(elapsed_steps, steps * total, gen, None)
    end.record()
    torch.cuda.synchronize()
    algo.cuda_graph_update = graphed
    if world > 1:
        dist.barrier()
    ms = torch.tensor([start.elapsed_time(end)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def workload_of(cfg, world: int, envs_per_gpu: int, n_steps: int, batch: int, n_epochs: int) -> dict:
    big = cfg.key in ("C4", "C5")
    return {"workload": f"{cfg.key}: {cfg.title}", "env": cfg.env, "n_envs_total": envs_per_gpu * world,
            "envs_per_gpu": envs_per_gpu, "n_steps": n_steps, "batch_size": batch, "n_epochs": n_epochs,
            "parallelism": f"env-sharded dp{world}",
            "l2": ("not flushed: every minibatch gathers fresh rows and the minibatch working set (logits + dlogits "
                   "490 MB per launch at C4, 119 MB at C5; rollout buffer 0.4-45 GB) exceeds the 126 MB L2" if big else
                   "not flushed and the working set fits L2: this config is launch-bound, reported for completeness")}


def run_reference_arm(args, cfg, threads: int):
    """The unmodified reference on the host cores (oracle/_ref through oracle/ref_shim.py)."""
    from oracle import ref_runner, ref_shim

    if not ref_shim.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref is missing: run oracle/make_ref.sh where /root/reference exists"})
        return
    n_envs = REF_ENVS.get(cfg.key, cfg.n_envs)
    value, ms, done, warm, sample = ref_runner.time_learn_epochs(cfg, max(1, args.steps), max(1, args.warmup), threads,
                                                                 n_envs=n_envs, budget_s=REF_BUDGET_S)
    w = workload_of(cfg, 1, n_envs, cfg.n_steps, cfg.algo["batch_size"], cfg.algo["n_epochs"])
    w["parallelism"] = f"one CPU process, {threads} torch threads"
    w["same_config_as_gpu_arm"] = n_envs == cfg.n_envs
    if n_envs != cfg.n_envs:
        w["sample"] = f"{n_envs} of the config's {cfg.n_envs} envs (same map, heads, trunk, n_steps and batch size)"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if cfg.key == "C5" else "weak", "vs_baseline": None,
        "dtype": "f32" if not cfg.algo.get("autocast_loss") else "bf16 autocast (CPU) / f32",
        "data": "synthetic", "config": w,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_cpu_baseline(cfg, threads: int):
    """cpu_baseline of the GPU arm: the same unmodified reference learner on a bounded sample (tens of seconds)."""
    from oracle import ref_runner, ref_shim

    if not ref_shim.available():
        return None
    n_envs, n_steps, batch = CPU_SAMPLE[cfg.key]
    v, _, _, _, sample = ref_runner.time_learn_epochs(cfg, 2, 1, threads, n_envs=n_envs, n_steps=n_steps,
                                                      algo_overrides={"batch_size": batch}, budget_s=45.0)
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample}


def gpu_leg(cfg, dev, rank: int, world: int, envs_per_gpu: int, steps: int, warmup: int, e2e_steps: int, local_rank: int):
    """Device-resident leg (value, roofline, kernels, stages) + end-to-end leg of one config on this rank."""
    from rl_algo_impls_b200 import ops
    from rl_algo_impls_b200.configs import build

    # The trunk's convolutions stay cuDNN calls; with fixed shapes (one rollout batch, one minibatch size) the library's
    # own autotuner picks their algorithms during the warm-up steps (measured on C4: 99.6 -> 91.2 ms per step; the
    # reference's set_seeds() turns it off for run-to-run determinism, runner/running_utils.py:181 -- it changes which
    # convolution kernel runs, not the arithmetic type).  B200RL_CUDNN_BENCHMARK=0 keeps cuDNN's heuristic choice.
    torch.backends.cudnn.benchmark = CUDNN_AUTOTUNE

    env, policy, gen, algo = build(cfg, dev, env_device=dev, seed=1234 + rank, n_envs=envs_per_gpu)
    # The roofline leg brackets the fused-loss launches with CUDA events, which a graph replay would hide: on the
    # GridNet configs one step in EAGER_EVERY runs its update eagerly (timed launches), the others replay the captured
    # update like the learner does by default (measured on C4: 41.7 ms graphed vs 46.6 ms eager update stage).
    eager_every = EAGER_EVERY if policy.kind == "gridnet" and algo.cuda_graph_update else 0
    algo_graphed = bool(algo.cuda_graph_update)
    for _ in range(max(3, warmup)):
        algo.learn_epoch(0, 1 << 40, gen, None)
    if eager_every:  # the eager path warm as well
        algo.cuda_graph_update = False
        algo.learn_epoch(0, 1 << 40, gen, None)
        algo.cuda_graph_update = True
    algo.profile_stages = True
    algo.learn_epoch(0, 1 << 40, gen, None)
    algo.profile_stages = False
    stages = algo.stage_ms
    timer = ops.KernelTimer(["b200rl_ppo_gridnet_loss", "b200rl_gae_scan_f32", "b200rl_gather_rows",
                             "b200rl_ppo_categorical_loss_f32", "b200rl_ppo_gaussian_loss_f32"])
    ops.set_kernel_timer(timer)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms = timed_epochs(algo, gen, steps, world, eager_every)
    clock_info = clocks.stop()
    ops.set_kernel_timer(None)
    launches = algo.launches_last_epoch * steps
    kernel_ms = timer.summary()
    rollout_steps = envs_per_gpu * cfg.n_steps
    value = rollout_steps * world * steps / (ms / 1e3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = peak_hbm_gbs()
    V = max(1, int(np.prod(policy.value_shape)))
    roofline = None
    if "b200rl_ppo_gridnet_loss" in kernel_ms:
        n, mean_ms = kernel_ms["b200rl_ppo_gridnet_loss"]
        nbytes = loss_kernel_bytes(policy, cfg.algo["batch_size"], V, 2 if cfg.algo.get("autocast_loss") else 4)
        achieved = nbytes / (mean_ms * 1e-3) / 1e9
        roofline = {"kernel": "gridnet_kernel<kPpo> via b200rl_ppo_gridnet_loss (+ its 1-block stats finaliser"
                              + ("" if policy.map_size <= 256 else " and the streaming pre-pass launch") + ")",
                    "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": measured_traffic(cfg.key, cfg.algo["batch_size"]), "bytes_per_launch": nbytes,
                    "ms_per_launch": mean_ms, "launches_timed": n, "peak_source": peak_src,
                    "note": "mask-driven kernel: logits are read only for cells with a valid action, so DRAM "
                            "traffic is below the algorithmic bytes (DESIGN.md 4.1)"}
    elif "b200rl_gae_scan_f32" in kernel_ms:
        n, mean_ms = kernel_ms["b200rl_gae_scan_f32"]
        nbytes = rollout_steps * (16 * V + 1)
        achieved = nbytes / (mean_ms * 1e-3) / 1e9
        roofline = {"kernel": "gae_scan_kernel via b200rl_gae_scan_f32", "bound": "hbm", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "bytes_per_launch": nbytes,
                    "ms_per_launch": mean_ms, "launches_timed": n, "peak_source": peak_src}
    kernels = {k: {"launches": n, "ms_mean": m} for k, (n, m) in kernel_ms.items()}

    # ---- end-to-end leg: host env, PCIe inside the timed region ------------------------------------
    e2e = None
    if e2e_steps > 0:
        del gen, env
        henv, _, hgen, halgo = build(cfg, dev, env_device=None, seed=1234 + rank, n_envs=envs_per_gpu)
        halgo.cuda_graph_update = algo.cuda_graph_update
        halgo.policy = policy
        hgen.policy = policy
        halgo.optimizer = algo.optimizer
        halgo._params_broadcast = True
        for _ in range(3):
            halgo.learn_epoch(0, 1 << 40, hgen, None)
        h2d0, d2h0 = hgen._upload.bytes, hgen.d2h_bytes
        ms_e2e = timed_epochs(halgo, hgen, e2e_steps, world)
        e2e = {"value": rollout_steps * world * e2e_steps / (ms_e2e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": (hgen._upload.bytes - h2d0) // e2e_steps,
               "d2h_bytes_per_step": (hgen.d2h_bytes - d2h0) // e2e_steps + halgo.d2h_bytes_last_epoch,
               "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "bytes_are": "per rank",
               "path": "host numpy VectorEnv -> DMA out of the env buffers (page-locked in place; small fields via pinned staging) -> HBM rollout buffer; sampled actions -> host"}
        del hgen, henv, halgo
    del algo, policy
    torch.cuda.empty_cache()
    update_mode = ("minibatch updates replay from CUDA graphs (the learner's default); every %dth timed step (the first "
                   "included) runs them eagerly so that CUDA events bracket its fused-loss launches: roofline / kernels "
                   "are those launches" % eager_every) if eager_every else (
                   "CUDA-graph replays" if algo_graphed else "eager")
    return dict(value=value, ms=ms, clocks=clock_info, launches=launches, roofline=roofline, kernels=kernels,
                stages=stages, e2e=e2e, update_mode=update_mode)


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
    ap.add_argument("--config", default=None, choices=["C1", "C2", "C3", "C4", "C5"],
                    help="default: C4 on one GPU, C5 (1024 envs sharded) on several")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=None, help="override the env count per GPU (weak-scaling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scale-base", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    from rl_algo_impls_b200.configs import CONFIGS

    key = args.config or ("C4" if max(args.gpus, world) == 1 else "C5")
    cfg = CONFIGS[key]
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, cfg, threads)
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sharded = key == "C5" and args.envs_per_gpu is None  # the config's GLOBAL env count split over the ranks
    envs_per_gpu = args.envs_per_gpu or (cfg.n_envs // world if key == "C5" else cfg.n_envs)
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5 if key == "C5" else 10))
    res = gpu_leg(cfg, dev, rank, world, envs_per_gpu, args.steps, args.warmup, e2e_steps, local_rank)

    # ---- the single-GPU point of the multi-GPU (C5, strong-scaling) curve ----------------------------
    scale_base = None
    if world == 1 and key == "C4" and args.config is None and not args.no_scale_base:
        c5 = CONFIGS["C5"]
        b = gpu_leg(c5, dev, rank, 1, c5.n_envs, 3, 3, 0, local_rank)
        scale_base = {"workload": workload_of(c5, 1, c5.n_envs, c5.n_steps, c5.algo["batch_size"], c5.algo["n_epochs"]),
                      "value": b["value"], "unit": UNIT, "steps": 3, "warmup": 3, "ms_per_step": b["ms"] / 3,
                      "roofline": b["roofline"], "kernels": b["kernels"], "stages_ms": b["stages"],
                      "note": "what `bench.py --gpus N` (N > 1) runs, on one GPU: divide the N-GPU values by this one"}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(cfg, threads)

    if rank == 0:
        bf16 = bool(cfg.algo.get("autocast_loss"))
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": res["ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": (("bf16 autocast trunk / f32 loss math on bf16 logits" if bf16 else
                       "f32 (loss kernels, GAE f64 carry); trunk convolutions run cuDNN's default TF32 tensor-op kernels")
                      + ("; convolution algorithms autotuned (torch.backends.cudnn.benchmark)" if CUDNN_AUTOTUNE else "")),
            "data": "synthetic",
            "config": workload_of(cfg, world, envs_per_gpu, cfg.n_steps, cfg.algo["batch_size"], cfg.algo["n_epochs"]),
            "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["launches"], "roofline": res["roofline"],
            "cpu_baseline": cpu, "kernels": res["kernels"], "stages_ms": res["stages"],
            "update_mode": res["update_mode"],
        }
        if scale_base is not None:
            line["scale_base"] = scale_base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
