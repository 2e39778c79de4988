"""Kernel-level roofline measurements (CUDA events on torch's current stream, inputs > L2 or L2
flushed between iterations).  Usage: python benchmarks/kernels.py [gae|loss|gather|all] [--json out]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rl_algo_impls_b200 import ops  # noqa: E402
from tests.synth import LUX_GATES, LUX_NVEC, MICRORTS_GATES, MICRORTS_NVEC  # noqa: E402


def peak_gbs() -> float:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"])
    return 6650.0


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


GRAPH = False  # --graph: time replays of ONE captured call (what the learner's graphed update replays)


def time_graph(fn, iters=20, warmup=3, flush=True):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    return float(np.median(times)), float(np.min(times))


def time_kernel(fn, name, iters=20, warmup=3, flush=True):
    """CUDA events recorded immediately around the C-ABI call `name` (ops.KernelTimer), so that the
    Python-side argument marshalling of the wrapper is outside the bracket."""
    if GRAPH:
        return time_graph(fn, iters, warmup, flush)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    timer = ops.KernelTimer([name])
    ops.set_kernel_timer(timer)
    for _ in range(iters):
        if flush:
            flush_l2()
        fn()
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    times = [s.elapsed_time(e) for s, e in timer.events[name]]
    return float(np.median(times)), float(np.min(times))


def bench_gae(T, N, V):
    dev = "cuda"
    shape = (T, N) if V == 1 else (T, N, V)
    r, v = torch.randn(shape, device=dev), torch.randn(shape, device=dev)
    es = torch.rand((T, N), device=dev) < 0.005
    nes = torch.rand((N,), device=dev) < 0.005
    nv = torch.randn(shape[1:], device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    gamma = 0.99 if V == 1 else np.full(V, 1.0)
    lam = 0.95 if V == 1 else np.full(V, 0.95)
    fn = lambda: ops.gae_scan(r, v, es, nes, nv, gamma, lam, adv, ret)
    med, best = time_kernel(fn, 'b200rl_gae_scan_f32')
    nbytes = T * N * (16 * V + 1) + N * (4 * V + 1)
    return dict(kernel="gae_scan", T=T, N=N, V=V, ms_median=med, ms_best=best, bytes=nbytes,
                gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / peak_gbs())


def bench_gae_segments(n_seg, mean_len, V, skips=False):
    """K1b: ragged trajectories concatenated along axis 0 (rollout/trajectory.py, discrete_skips_trajectory_builder.py)."""
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(max(1, mean_len // 2), mean_len * 3 // 2 + 1, (n_seg,), generator=g)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)]).to(dev)
    total = int(offsets[-1].item())
    shape = (total,) if V == 1 else (total, V)
    r, v = torch.randn(shape, device=dev), torch.randn(shape, device=dev)
    nes = torch.rand(n_seg, device=dev) < 0.5
    nv = torch.randn((n_seg,) if V == 1 else (n_seg, V), device=dev)
    gamma = 0.99 if V == 1 else np.full(V, 0.99)
    lam = 0.95 if V == 1 else np.full(V, 0.95)
    kw = (dict(steps_elapsed=torch.randint(1, 5, (total,), device=dev, dtype=torch.int32)) if skips
          else dict(episode_starts=torch.rand(total, device=dev) < 0.02))
    fn = lambda: ops.gae_segments(r, v, offsets, nes, nv, gamma, lam, **kw)
    med, best = time_kernel(fn, 'b200rl_gae_segments_f32', flush=False)
    nbytes = total * (16 * V + (4 if skips else 1))
    return dict(kernel="gae_segments" + ("_skips" if skips else ""), segments=n_seg, steps=total, V=V, ms_median=med,
                ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / peak_gbs())


def bench_small(name):
    """K6 / K7: per-env-step kernels (launch-bound by construction: one env step's [N, D] at a time)."""
    dev = "cuda"
    if name == "norm_obs":
        N, D = 4096, 17
        x = torch.randn(N, D, device=dev)
        st = [torch.zeros(D, dtype=torch.float64, device=dev), torch.ones(D, dtype=torch.float64, device=dev),
              torch.full((D,), 1e-4, dtype=torch.float64, device=dev)]
        out = torch.empty_like(x)
        fn = lambda: ops.running_norm_obs(x, *st, True, 1e-8, 10.0, out)
        nbytes, cname = 2 * x.numel() * 4, "b200rl_running_norm_obs_f32"
    elif name == "norm_reward":
        N, V = 1024, 13
        r = torch.randn(N, V, device=dev)
        d = torch.rand(N, device=dev) < 0.01
        ret = torch.zeros(N, V, dtype=torch.float64, device=dev)
        st = [torch.zeros(V, dtype=torch.float64, device=dev), torch.ones(V, dtype=torch.float64, device=dev),
              torch.full((V,), 1e-4, dtype=torch.float64, device=dev)]
        out = torch.empty_like(r)
        fn = lambda: ops.running_norm_reward(r, d, ret, *st, 0.99, True, 1e-8, 10.0, out)
        nbytes, cname = r.numel() * (4 + 4 + 16) + N, "b200rl_running_norm_reward_f32"
    else:
        N, K = 1024, 12
        base = torch.randn(N, device=dev)
        series = [torch.randn(N, device=dev) for _ in range(K)]
        term, trunc = torch.rand(N, device=dev) < 0.01, torch.zeros(N, dtype=torch.bool, device=dev)
        out = torch.empty(N, 1 + K, device=dev)
        fn = lambda: ops.reward_assemble(base, series, term, trunc, [False] * (K - 1) + [True], [1.0] * (K - 1) + [0.001], out)
        nbytes, cname = N * (1 + K) * 8 + 2 * N, "b200rl_reward_assemble_f32"
    med, best = time_kernel(fn, cname, flush=False)
    return dict(kernel=name, ms_median=med, ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6, note="launch-bound by construction")


def gridnet_tensors(B, HW, nvec, n_pick, unit_p, dtype=torch.float32, act_dtype=torch.uint8):
    dev = "cuda"
    S, A = sum(nvec), len(nvec)
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn((B, HW, S + n_pick), device=dev, generator=g).to(dtype)
    has_unit = torch.rand((B, HW, 1), device=dev, generator=g) < unit_p
    mask = (torch.rand((B, HW, S), device=dev, generator=g) < 0.5) & has_unit
    actions = torch.stack([torch.randint(0, n, (B, HW), device=dev, generator=g) for n in nvec], -1).to(act_dtype)
    pick_mask = pick = None
    if n_pick:
        pick_mask = torch.rand((B, n_pick, HW), device=dev, generator=g) < 0.05
        pick = torch.randint(0, HW, (B, n_pick), device=dev, generator=g)
    return logits, mask, pick_mask, actions, pick


def bench_loss(B, HW, nvec, gates, n_pick, V, unit_p, dtype=torch.float32, inplace=False, ld=0):
    """inplace: the learner's call (persistent dlogits, previous rows cleared); ld: channel-padded logit rows."""
    dev = "cuda"
    logits, mask, pick_mask, actions, pick = gridnet_tensors(B, HW, nvec, n_pick, unit_p, dtype)
    if ld:
        padded = torch.zeros(logits.shape[:-1] + (ld,), dtype=dtype, device=dev)
        padded[..., :logits.shape[-1]] = logits
        logits = padded
    spec = ops.GridnetSpec.from_subaction_mask(nvec, gates, n_pick)
    vs = (B,) if V == 1 else (B, V)
    old_logp = torch.randn(B, device=dev) * 0.1 - 20
    adv = torch.randn(vs, device=dev)
    ov, rt, nv = torch.randn(vs, device=dev), torch.randn(vs, device=dev), torch.randn(vs, device=dev)
    w = [1.0 / V] * V if V > 1 else None
    h = ops.PpoHyper(clip_range=0.1, clip_range_vf=0.1, ent_coef=0.01, vf_coef=[0.5] * V, adv_weights=w)
    moments = ops.adv_moments(adv.view(B, V), None, ops.ADV_NORMALIZE, w)
    fn = lambda: ops.ppo_gridnet_loss(h, spec, logits, mask, pick_mask, actions, pick, old_logp, adv, ov, rt, nv,
                                      moments=moments, inplace=inplace)
    med, best = time_kernel(fn, 'b200rl_ppo_gridnet_loss')
    S, A, es = sum(nvec), len(nvec), logits.element_size()
    nbytes = B * (2 * es * HW * (S + n_pick) + HW * S + n_pick * HW + HW * A + 2 * n_pick + 4 * (2 + 5 * V))
    return dict(kernel="ppo_gridnet_loss" + ("_inplace" if inplace else ""), B=B, HW=HW, S=S + n_pick, ld=ld or S + n_pick,
                V=V, unit_p=unit_p, dtype=str(dtype), ms_median=med, ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6,
                frac=nbytes / med / 1e6 / peak_gbs())


def bench_gather(M, B, row_shapes):
    dev = "cuda"
    srcs = [torch.empty((M,) + tuple(s), dtype=dt, device=dev) for s, dt in row_shapes]
    for s in srcs:
        s.view(torch.uint8).random_(0, 255) if s.dtype != torch.bool else None
    idx = torch.randperm(M, device=dev)[:B]
    fn = lambda: ops.gather_rows(srcs, idx)
    med, best = time_kernel(fn, 'b200rl_gather_rows')
    nbytes = 2 * B * sum(int(np.prod(s.shape[1:])) * s.element_size() for s in srcs)
    return dict(kernel="gather_rows", M=M, B=B, ms_median=med, ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6,
                frac=nbytes / med / 1e6 / peak_gbs())


def bench_glue(N, C, H, W):
    """K8: bias + max-pool(3, 2, 1) + ReLU forward / backward and bias + ReLU forward / backward on a channels-last map."""
    dev = "cuda"
    y = torch.randn((N, C, H, W), device=dev).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, device=dev)
    rows = []
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    yy, bb = y.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    out = ops.bias_pool_relu(yy, bb)
    dout = torch.randn_like(out)
    n_in, n_out = N * C * H * W, N * C * Ho * Wo
    for name, fn, call, nbytes in (
        ("bias_pool_relu_fwd", lambda: ops.bias_pool_relu(yy, bb), "b200rl_nhwc_bias_pool_relu_fwd", 4 * n_in + 5 * n_out),
        ("bias_pool_relu_bwd", lambda: out.backward(dout, retain_graph=True), "b200rl_nhwc_bias_pool_relu_bwd",
         4 * n_in + 5 * n_out + 5 * n_out),  # dx written; dout + codes read by the gather and again by the bias sum
    ):
        med, best = time_kernel(fn, call)
        rows.append(dict(kernel=name, N=N, C=C, H=H, W=W, ms_median=med, ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6,
                         frac=nbytes / med / 1e6 / peak_gbs()))
    y2 = (y * 1.0).requires_grad_(True)
    out2 = ops.bias_relu(y2 * 1.0, bb)
    d2 = torch.randn_like(out2)
    for name, fn, call, nbytes in (
        ("bias_relu_fwd", lambda: ops.bias_relu(y2 * 1.0, bb), "b200rl_nhwc_bias_relu_fwd", 8 * n_in),
        ("bias_relu_bwd", lambda: out2.backward(d2, retain_graph=True), "b200rl_nhwc_bias_relu_bwd", 12 * n_in + 8 * n_in),
    ):
        med, best = time_kernel(fn, call)
        rows.append(dict(kernel=name, N=N, C=C, H=H, W=W, ms_median=med, ms_best=best, bytes=nbytes, gbs=nbytes / med / 1e6,
                         frac=nbytes / med / 1e6 / peak_gbs()))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--json", default=None)
    ap.add_argument("--graph", action="store_true", help="time graph replays of one captured call instead of eager calls")
    a = ap.parse_args()
    global GRAPH
    GRAPH = a.graph
    rows = []
    if a.what in ("gae", "all"):
        for T, N, V in [(32, 8, 1), (128, 8, 1), (64, 4096, 1), (512, 24, 1), (32, 1024, 13), (128, 1 << 20, 1),
                        (32, 131072, 13)]:
            rows.append(bench_gae(T, N, V))
            print(json.dumps(rows[-1]), flush=True)
    if a.what in ("gae", "gae_small", "all"):
        rows.append(bench_gae_segments(24, 512, 1))
        print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_gae_segments(2048, 32, 13))
        print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_gae_segments(2048, 32, 1, skips=True))
        print(json.dumps(rows[-1]), flush=True)
        for name in ("norm_obs", "norm_reward", "reward_assemble"):
            rows.append(bench_small(name))
            print(json.dumps(rows[-1]), flush=True)
    if a.what == "gae_small":
        for T, N, V in [(32, 8, 1), (128, 8, 1), (64, 4096, 1), (512, 24, 1), (32, 1024, 13), (2048, 24, 1), (512, 4096, 1)]:
            rows.append(bench_gae(T, N, V))
            print(json.dumps(rows[-1]), flush=True)
    if a.what in ("loss", "all"):
        for B in (256, 3072):
            rows.append(bench_loss(B, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06))
            print(json.dumps(rows[-1]), flush=True)
        for B in (128, 512):
            rows.append(bench_loss(B, 4096, LUX_NVEC, LUX_GATES, 1, 13, 0.02))
            print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06, torch.bfloat16))
        print(json.dumps(rows[-1]), flush=True)
    if a.what in ("loss", "all"):  # dense masks: every cell has a unit (worst case for the mask-driven kernel)
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 1.0))
        print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_loss(512, 4096, LUX_NVEC, LUX_GATES, 1, 13, 1.0))
        print(json.dumps(rows[-1]), flush=True)
    if a.what in ("loss_inplace", "all"):  # what the learner launches: persistent dlogits, padded rows
        for B, up in ((3072, 0.06), (3072, 0.12), (3072, 0.25), (3072, 1.0), (256, 0.06)):
            rows.append(bench_loss(B, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, up, inplace=True, ld=80))
            print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06, torch.bfloat16, inplace=True, ld=80))
        print(json.dumps(rows[-1]), flush=True)
        for B, dt in ((128, torch.bfloat16), (128, torch.float32), (512, torch.bfloat16), (512, torch.float32)):
            rows.append(bench_loss(B, 4096, LUX_NVEC, LUX_GATES, 1, 13, 0.02, dt, inplace=True, ld=32))
            print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_loss(512, 4096, LUX_NVEC, LUX_GATES, 1, 13, 1.0, torch.bfloat16, inplace=True, ld=32))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_c4_inplace":  # the ncu target of round 2
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06, inplace=True, ld=80))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_c5_inplace":
        rows.append(bench_loss(128, 4096, LUX_NVEC, LUX_GATES, 1, 13, 0.02, torch.bfloat16, inplace=True, ld=32))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_c4":  # one shape, few launches: the ncu target
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_dense":
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 1.0))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_bf16":
        rows.append(bench_loss(3072, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 1, 0.06, torch.bfloat16))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "loss_c5":
        rows.append(bench_loss(512, 4096, LUX_NVEC, LUX_GATES, 1, 13, 0.02))
        print(json.dumps(rows[-1]), flush=True)
    if a.what == "gae_big":
        rows.append(bench_gae(128, 1 << 20, 1))
        print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_gae(32, 131072, 13))
        print(json.dumps(rows[-1]), flush=True)
    if a.what in ("glue", "all"):  # K8 at the C4 minibatch's encoder levels 1 and 2, and the rollout step's level 1
        for shape in ((3072, 32, 16, 16), (3072, 64, 8, 8), (24, 32, 16, 16)):
            for r in bench_glue(*shape):
                rows.append(r)
                print(json.dumps(r), flush=True)
    if a.what in ("gather", "all"):
        rows.append(bench_gather(12288, 3072, [((74, 16, 16), torch.float32), ((256, 78), torch.uint8),
                                               ((256, 7), torch.uint8), ((), torch.float32), ((), torch.float32),
                                               ((), torch.float32), ((), torch.float32)]))
        print(json.dumps(rows[-1]), flush=True)
        rows.append(bench_gather(8192, 2048, [((4, 84, 84), torch.uint8), ((), torch.float32), ((), torch.int64)]))
        print(json.dumps(rows[-1]), flush=True)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "sample"):
    main()


def bench_sample(N, HW, nvec, gates, n_pick, unit_p):
    logits, mask, pick_mask, _, _ = gridnet_tensors(N, HW, nvec, n_pick, unit_p)
    spec = ops.GridnetSpec.from_subaction_mask(nvec, gates, n_pick)
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    fn = lambda: ops.gridnet_sample(spec, logits, mask, pick_mask, 1, 0, torch.uint8, counter)
    med, best = time_kernel(fn, "b200rl_gridnet_sample", flush=False)
    return dict(kernel="gridnet_sample", N=N, HW=HW, unit_p=unit_p, ms_median=med, ms_best=best)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sample":
    print(json.dumps(bench_sample(24, 256, MICRORTS_NVEC, MICRORTS_GATES, 0, 0.06)))
    print(json.dumps(bench_sample(128, 4096, LUX_NVEC, LUX_GATES, 1, 0.02)))
