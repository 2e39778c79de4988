"""Tensor-level entry points over the C ABI (include/b200rl.h).

Every function takes CUDA tensors, checks dtype / contiguity, and launches on torch's current
stream.  Nothing here computes on the host and nothing falls back to eager PyTorch: a missing
library or a CPU tensor raises.
"""
import contextlib
import ctypes as C
import gc
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import GridnetDesc, PpoArgs, check

NumOrArray = Union[float, int, np.ndarray, Sequence[float]]

ADV_NONE, ADV_NORMALIZE, ADV_STANDARDIZE, ADV_AFTER_SCALING = 0, 1, 2, 3

_INDEX_DTYPES = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.int64: _lib.I64}
_LOGIT_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """CUDA-event timing of individual C-ABI calls on torch's current stream (bench.py's roofline
    leg).  Events are recorded on the stream the kernel is launched on; read with ``summary()``."""

    def __init__(self, names: Optional[Sequence[str]] = None):
        self.names = set(names) if names else None
        self.events: Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = {}

    def summary(self) -> Dict[str, Tuple[int, float]]:
        torch.cuda.synchronize()
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v) / len(v)) for k, v in self.events.items() if v}


LAUNCHES = 0  # kernels of libb200rl.so launched since import (bench.py reports the delta)
_timer: Optional[KernelTimer] = None


def set_kernel_timer(timer: Optional[KernelTimer]) -> None:
    global _timer
    _timer = timer


def _call(name: str, n_kernels: int, fn, *args) -> int:
    global LAUNCHES
    LAUNCHES += n_kernels
    t = _timer
    if t is None or (t.names is not None and name not in t.names):
        return fn(*args)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    t.events.setdefault(name, []).append((s, e))
    return rc


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _cuda(t: torch.Tensor, dtype: Optional[torch.dtype], name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B200RLError(f"{name}: expected a CUDA tensor (rl_algo_impls_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return t


def _as_u8(mask: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if not mask.is_cuda:
        raise _lib.B200RLError(f"{name}: expected a CUDA tensor")
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)  # same bytes, no copy
    if mask.dtype != torch.uint8:
        raise TypeError(f"{name}: expected bool or uint8, got {mask.dtype}")
    return mask.contiguous()


def _f32_array(values: Sequence[float]):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr


_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


@contextlib.contextmanager
def no_gc_during_capture():
    """Python's cyclic collector must not run finalizers in the middle of a stream capture: an object from an earlier
    run (a rollout generator's page-locked host buffers, an old graph's tensors) released then can issue a CUDA call that
    is illegal under global-mode capture and invalidates it.  Collect first, keep the collector off for the capture."""
    was_enabled = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        yield
    finally:
        if was_enabled:
            gc.enable()


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per (device, stream) scratch buffer.  Calls on one stream are ordered, so reuse is safe.  While a CUDA graph is
    being captured the scratch is a fresh allocation out of that graph's own pool instead: a cached buffer may belong
    to the pool of an earlier graph captured on a recycled stream handle, and dropping / regrowing it mid-capture would
    touch that other pool from inside this capture."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# K1
def gae_scan(
    rewards: torch.Tensor,
    values: torch.Tensor,
    episode_starts: torch.Tensor,
    next_episode_starts: torch.Tensor,
    next_values: torch.Tensor,
    gamma: NumOrArray,
    gae_lambda: NumOrArray,
    out_advantages: Optional[torch.Tensor] = None,
    out_returns: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """advantages, returns over a time-major [T, N(, V)] rollout (shared/gae.py:97-124, vec_rollout.py:88)."""
    _cuda(rewards, torch.float32, "rewards")
    _cuda(values, torch.float32, "values")
    _cuda(next_values, torch.float32, "next_values")
    es, nes = _as_u8(episode_starts, "episode_starts"), _as_u8(next_episode_starts, "next_episode_starts")
    if rewards.shape != values.shape or rewards.dim() not in (2, 3):
        raise ValueError(f"rewards {tuple(rewards.shape)} / values {tuple(values.shape)} must be [T, N] or [T, N, V]")
    T, N = rewards.shape[:2]
    V = rewards.shape[2] if rewards.dim() == 3 else 1
    if es.shape != (T, N) or nes.shape != (N,) or next_values.numel() != N * V:
        raise ValueError("episode_starts / next_episode_starts / next_values shapes do not match rewards")
    gamma_is_scalar = not isinstance(gamma, (np.ndarray, list, tuple))

    def per_head(x: NumOrArray, name: str):
        a = np.asarray(x, dtype=np.float64).reshape(-1)
        if a.size == 1:
            a = np.repeat(a, V)
        if a.size != V:
            raise ValueError(f"{name} has {a.size} entries for {V} value heads")
        return (C.c_double * V)(*a.tolist())

    g, lam = per_head(gamma, "gamma"), per_head(gae_lambda, "gae_lambda")
    adv = out_advantages if out_advantages is not None else torch.empty_like(rewards)
    ret = out_returns if out_returns is not None else torch.empty_like(rewards)
    _cuda(adv, torch.float32, "out_advantages"), _cuda(ret, torch.float32, "out_returns")
    if T == 0 or N == 0:  # nothing to scan (and empty tensors have null data pointers)
        return adv, ret
    rc = _call("b200rl_gae_scan_f32", 1, _lib.lib().b200rl_gae_scan_f32,
        rewards.data_ptr(), values.data_ptr(), es.data_ptr(), nes.data_ptr(), next_values.data_ptr(),
        g, lam, int(gamma_is_scalar), adv.data_ptr(), ret.data_ptr(), T, N, V, _stream(),
    )  # fmt: skip
    check(rc, "b200rl_gae_scan_f32")
    return adv, ret


def gae_segments(
    rewards: torch.Tensor,
    values: torch.Tensor,
    seg_offsets: torch.Tensor,
    next_episode_starts: torch.Tensor,
    next_values: torch.Tensor,
    gamma: NumOrArray,
    gae_lambda: NumOrArray,
    episode_starts: Optional[torch.Tensor] = None,
    steps_elapsed: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """advantages, returns over ragged trajectories concatenated along axis 0 ([total] or [total, V]):
    `episode_starts` -> the reference's per-Trajectory compute_advantages (rollout/trajectory.py:56-95),
    `steps_elapsed`  -> the discrete-skips recurrence (discrete_skips_trajectory_builder.py:84-100)."""
    _cuda(rewards, torch.float32, "rewards"), _cuda(values, torch.float32, "values")
    _cuda(next_values, torch.float32, "next_values"), _cuda(seg_offsets, torch.int64, "seg_offsets")
    if (episode_starts is None) == (steps_elapsed is None):
        raise ValueError("pass exactly one of episode_starts / steps_elapsed")
    total = rewards.shape[0]
    V = rewards.numel() // max(total, 1) if total else (next_values.numel() // max(seg_offsets.numel() - 1, 1))
    n_seg = seg_offsets.numel() - 1
    nes = _as_u8(next_episode_starts, "next_episode_starts")
    es = _as_u8(episode_starts, "episode_starts") if episode_starts is not None else None
    if steps_elapsed is not None:
        _cuda(steps_elapsed, torch.int32, "steps_elapsed")
    gamma_is_scalar = not isinstance(gamma, (np.ndarray, list, tuple))

    def per_head(x):
        a = np.asarray(x, dtype=np.float64).reshape(-1)
        a = np.repeat(a, V) if a.size == 1 else a
        return (C.c_double * V)(*a.tolist())

    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    if total == 0 or n_seg <= 0:
        return adv, ret
    rc = _call("b200rl_gae_segments_f32", 1, _lib.lib().b200rl_gae_segments_f32, rewards.data_ptr(), values.data_ptr(),
               _ptr(es), _ptr(steps_elapsed), seg_offsets.data_ptr(), nes.data_ptr(), next_values.data_ptr(),
               per_head(gamma), per_head(gae_lambda), int(gamma_is_scalar), adv.data_ptr(), ret.data_ptr(), n_seg, V,
               _stream())
    check(rc, "b200rl_gae_segments_f32")
    return adv, ret


# ------------------------------------------------------------------------------------------------
# K2
def adv_moments(
    adv: torch.Tensor, idx: Optional[torch.Tensor], mode: int, weights: Optional[Sequence[float]] = None
) -> torch.Tensor:
    """(sum[Vm], sumsq[Vm], count) in f64 over rows ``idx`` of adv [M, V] (ppo.py:307-318 statistics)."""
    _cuda(adv, torch.float32, "adv")
    V = adv.shape[1] if adv.dim() == 2 else 1
    B = adv.shape[0] if idx is None else idx.numel()
    if idx is not None:
        _cuda(idx, torch.int64, "idx")
    vm = 1 if mode == ADV_AFTER_SCALING else V
    moments = torch.empty(2 * vm + 1, dtype=torch.float64, device=adv.device)
    L = _lib.lib()
    nbytes = L.b200rl_adv_moments_workspace_bytes(B, V)
    ws = _workspace(nbytes, adv.device)
    w = _f32_array(weights) if weights is not None else None
    rc = _call("b200rl_adv_moments_f64", 2, L.b200rl_adv_moments_f64,
        adv.data_ptr(), _ptr(idx), B, V, mode, w, moments.data_ptr(), ws.data_ptr(), ws.numel(), _stream()
    )
    check(rc, "b200rl_adv_moments_f64")
    return moments


def adv_normalize(
    adv: torch.Tensor,
    idx: Optional[torch.Tensor],
    mode: int,
    weights: Optional[Sequence[float]] = None,
    moments: Optional[torch.Tensor] = None,
    contract: bool = True,
) -> torch.Tensor:
    """Normalised (and, with ``contract``, reward-weight contracted) advantages of a minibatch."""
    _cuda(adv, torch.float32, "adv")
    V = adv.shape[1] if adv.dim() == 2 else 1
    B = adv.shape[0] if idx is None else idx.numel()
    if mode != ADV_NONE and moments is None:
        moments = adv_moments(adv, idx, mode, weights)
    out_v = 1 if (contract and (weights is not None or V == 1)) or mode == ADV_AFTER_SCALING else V
    out = torch.empty((B,) if out_v == 1 else (B, V), dtype=torch.float32, device=adv.device)
    w = _f32_array(weights) if weights is not None else None
    rc = _call("b200rl_adv_normalize_f32", 1, _lib.lib().b200rl_adv_normalize_f32,
        adv.data_ptr(), _ptr(idx), B, V, mode, w, _ptr(moments), out.data_ptr(), out_v, _stream()
    )
    check(rc, "b200rl_adv_normalize_f32")
    return out


# ------------------------------------------------------------------------------------------------
# K3
def gather_rows(sources: Sequence[torch.Tensor], idx: torch.Tensor) -> List[torch.Tensor]:
    """[src[idx] for src in sources] in (at most) two launches (rollout.py:56-69)."""
    _cuda(idx, torch.int64, "idx")
    B = idx.numel()
    if not sources:
        return []
    n_rows = sources[0].shape[0]
    outs: List[torch.Tensor] = []
    for start in range(0, len(sources), _lib.MAX_GATHER):
        group = sources[start : start + _lib.MAX_GATHER]
        n = len(group)
        src_arr, dst_arr, rb_arr = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int64 * n)()
        for k, s in enumerate(group):
            _cuda(s, None, f"sources[{start + k}]")
            if s.shape[0] != n_rows:
                raise ValueError("every gathered tensor must have the same number of rows")
            d = torch.empty((B,) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
            outs.append(d)
            src_arr[k], dst_arr[k] = s.data_ptr(), d.data_ptr()
            rb_arr[k] = (s.numel() // max(n_rows, 1)) * s.element_size()
        rc = _call("b200rl_gather_rows", 2, _lib.lib().b200rl_gather_rows,src_arr, dst_arr, rb_arr, n, idx.data_ptr(), B, n_rows, _stream())
        check(rc, "b200rl_gather_rows")
    return outs


# ------------------------------------------------------------------------------------------------
# K0
def rollout_store_step(step_tensors: Sequence[torch.Tensor], buffers: Sequence[torch.Tensor], step_dev: torch.Tensor,
                       carry: Optional[Sequence[Optional[torch.Tensor]]] = None,
                       carry_or: Optional[Sequence[Optional[torch.Tensor]]] = None,
                       pack: Optional[Tuple[int, int, int, int, int]] = None,
                       advance_ticket: Optional[torch.Tensor] = None) -> None:
    """buffers[t][*step_dev % T] = step_tensors[t] for every rollout field in one launch
    (sync_step_rollout.py:188-201); the step index is read on the device (CUDA-graph friendly).
    ``carry[t]`` (same shape / dtype as step_tensors[t], or None) is then copied INTO step_tensors[t] by the same
    launch: the env's output for the next step replaces the slice that was just stored (:202-212).
    ``carry_or[t]``: step_tensors[t] <- carry[t] | carry_or[t] (bool / uint8 fields: terminations | truncations).
    ``pack = (field, N, C, HW, Cp)``: step_tensors[field] is a packed [N, HW, Cp] float32 observation and carry[field]
    the env's raw [N, C, HW] float32 one, transposed into that layout by the launch.
    ``advance_ticket`` (zeroed int32 device scalar, kept by the caller): ``step_dev += 1`` once every CTA has read it."""
    _cuda(step_dev, torch.int64, "step_dev")
    n = len(step_tensors)
    if n == 0 and advance_ticket is None:
        return
    if n > _lib.MAX_GATHER:
        raise ValueError(f"at most {_lib.MAX_GATHER} fields per call")
    T = buffers[0].shape[0] if n else 1
    src_arr, dst_arr, sb_arr = (C.c_void_p * max(n, 1))(), (C.c_void_p * max(n, 1))(), (C.c_int64 * max(n, 1))()
    carry_arr, or_arr = (C.c_void_p * max(n, 1))(), (C.c_void_p * max(n, 1))()
    for k, (s, d) in enumerate(zip(step_tensors, buffers)):
        _cuda(s, None, f"step_tensors[{k}]"), _cuda(d, None, f"buffers[{k}]")
        if d.shape[0] != T or s.dtype != d.dtype or s.numel() * T != d.numel():
            raise ValueError(f"field {k}: step slice {tuple(s.shape)} {s.dtype} does not fit buffer {tuple(d.shape)} {d.dtype}")
        src_arr[k], dst_arr[k], sb_arr[k] = s.data_ptr(), d.data_ptr(), s.numel() * s.element_size()
        c = carry[k] if carry is not None else None
        packed = pack is not None and pack[0] == k
        if c is not None:
            _cuda(c, None, f"carry[{k}]")
            if packed:
                _, N, Cc, HW, Cp = pack
                if c.dtype != torch.float32 or s.dtype != torch.float32 or c.numel() != N * Cc * HW or s.numel() != N * HW * Cp:
                    raise ValueError(f"pack: raw {tuple(c.shape)} {c.dtype} / packed {tuple(s.shape)} {s.dtype} do not match {pack}")
            elif c.dtype != s.dtype or c.numel() != s.numel():
                raise ValueError(f"carry[{k}] {tuple(c.shape)} {c.dtype} does not match the step slice {tuple(s.shape)} {s.dtype}")
            carry_arr[k] = c.data_ptr()
        elif packed:
            raise ValueError("pack: the packed field needs the raw observation as its carry")
        o = carry_or[k] if carry_or is not None else None
        if o is not None:
            _cuda(o, None, f"carry_or[{k}]")
            if c is None or o.numel() != s.numel() or o.element_size() != 1 or s.element_size() != 1:
                raise ValueError(f"carry_or[{k}]: needs carry[{k}] and one-byte elements of the step slice's size")
            or_arr[k] = o.data_ptr()
    L = _lib.lib()
    if carry_or is None and pack is None and advance_ticket is None:
        rc = _call("b200rl_rollout_store_step_carry", 1, L.b200rl_rollout_store_step_carry, src_arr, dst_arr, sb_arr,
                   carry_arr if carry is not None else None, n, step_dev.data_ptr(), T, _stream())
        check(rc, "b200rl_rollout_store_step_carry")
        return
    pk = None
    if pack is not None:
        pk = _lib.StorePack()
        pk.field, pk.N, pk.C, pk.HW, pk.Cp = pack
    if advance_ticket is not None:
        _cuda(advance_ticket, torch.int32, "advance_ticket")
    rc = _call("b200rl_rollout_store_step_fused", 1, L.b200rl_rollout_store_step_fused, src_arr, dst_arr, sb_arr,
               carry_arr if carry is not None else None, or_arr if carry_or is not None else None, n,
               C.byref(pk) if pk is not None else None, step_dev.data_ptr(), T, int(advance_ticket is not None),
               _ptr(advance_ticket), _stream())
    check(rc, "b200rl_rollout_store_step_fused")


# ------------------------------------------------------------------------------------------------
# K6
def running_norm_obs(x: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, count: torch.Tensor, training: bool,
                     epsilon: float, clip: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NormalizeObservation.normalize (wrappers/normalize.py:39-48): update the running moments with
    this batch (when training), then clip((x - mean) / sqrt(var + eps)).  x: [N, ...] float32."""
    _cuda(x, torch.float32, "x")
    N = x.shape[0]
    D = x.numel() // N
    for name, t in (("mean", mean), ("var", var), ("count", count)):
        _cuda(t, torch.float64, name)
        if t.numel() != D:
            raise ValueError(f"{name} must have {D} entries")
    out = torch.empty_like(x) if out is None else _cuda(out, torch.float32, "out")
    rc = _call("b200rl_running_norm_obs_f32", 1, _lib.lib().b200rl_running_norm_obs_f32, x.data_ptr(), N, D,
               mean.data_ptr(), var.data_ptr(), count.data_ptr(), int(training), float(epsilon), float(clip),
               out.data_ptr(), _stream())
    check(rc, "b200rl_running_norm_obs_f32")
    return out


def running_norm_reward(rewards: torch.Tensor, dones: torch.Tensor, returns: torch.Tensor, mean: torch.Tensor,
                        var: torch.Tensor, count: torch.Tensor, gamma: float, training: bool, epsilon: float,
                        clip: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NormalizeReward.step (wrappers/normalize.py:84-110): discounted-return accumulator, running
    variance of the returns, clip(r / sqrt(var + eps)), returns[done] = 0.  rewards: [N] or [N, V]."""
    _cuda(rewards, torch.float32, "rewards")
    N = rewards.shape[0]
    V = rewards.numel() // N
    d = _as_u8(dones, "dones")
    _cuda(returns, torch.float64, "returns")
    if returns.numel() != N * V or d.numel() != N:
        raise ValueError("returns must be [N, V] and dones [N]")
    for name, t in (("mean", mean), ("var", var), ("count", count)):
        _cuda(t, torch.float64, name)
        if t.numel() != V:
            raise ValueError(f"{name} must have {V} entries")
    out = torch.empty_like(rewards) if out is None else _cuda(out, torch.float32, "out")
    rc = _call("b200rl_running_norm_reward_f32", 1, _lib.lib().b200rl_running_norm_reward_f32, rewards.data_ptr(),
               d.data_ptr(), N, V, float(gamma), returns.data_ptr(), mean.data_ptr(), var.data_ptr(), count.data_ptr(),
               int(training), float(epsilon), float(clip), out.data_ptr(), _stream())
    check(rc, "b200rl_running_norm_reward_f32")
    return out


def running_norm_reward_ema(rewards: torch.Tensor, dones: torch.Tensor, returns: torch.Tensor, mean: torch.Tensor,
                            var: torch.Tensor, count: torch.Tensor, ema_mean: torch.Tensor, ema_sq: torch.Tensor,
                            ema_var: torch.Tensor, ema_init: torch.Tensor, alpha: float, gamma: float, training: bool,
                            epsilon: float, clip: float, out: Optional[torch.Tensor] = None,
                            per_env: bool = False) -> torch.Tensor:
    """NormalizeReward.step with exponential_moving_mean_var=True (wrappers/normalize.py:74-110 over
    HybridMovingMeanVar, utils/running_mean_std.py:120-170): running and exponential-moving moments of the
    discounted returns updated together, the reward divided by the blended standard deviation."""
    _cuda(rewards, torch.float32, "rewards")
    N = rewards.shape[0]
    V = rewards.numel() // N
    d = _as_u8(dones, "dones")
    _cuda(returns, torch.float64, "returns")
    if returns.numel() != N * V or d.numel() != N:
        raise ValueError("returns must be [N, V] and dones [N]")
    for name, t in (("mean", mean), ("var", var), ("count", count), ("ema_mean", ema_mean), ("ema_sq", ema_sq),
                    ("ema_var", ema_var)):
        _cuda(t, torch.float64, name)
        want = N if per_env and name.startswith("ema_") else V
        if t.numel() != want:
            raise ValueError(f"{name} must have {want} entries")
    _cuda(ema_init, torch.int32, "ema_init")
    out = torch.empty_like(rewards) if out is None else _cuda(out, torch.float32, "out")
    rc = _call("b200rl_running_norm_reward_ema_f32", 1, _lib.lib().b200rl_running_norm_reward_ema_f32,
               rewards.data_ptr(), d.data_ptr(), N, V, float(gamma), returns.data_ptr(), mean.data_ptr(), var.data_ptr(),
               count.data_ptr(), ema_mean.data_ptr(), ema_sq.data_ptr(), ema_var.data_ptr(), ema_init.data_ptr(),
               float(alpha), int(bool(per_env)), int(training), float(epsilon), float(clip), out.data_ptr(), _stream())
    check(rc, "b200rl_running_norm_reward_ema_f32")
    return out


# ------------------------------------------------------------------------------------------------
# K7
def reward_assemble(base: torch.Tensor, series: Sequence[torch.Tensor], terminations: Optional[torch.Tensor],
                    truncations: Optional[torch.Tensor], episode_end: Sequence[bool],
                    multiplier: Optional[Sequence[float]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """InfoRewardsWrapper.step (wrappers/info_rewards_wrapper.py:39-57): [N, V0 + K] rewards = the env's own
    reward head(s) followed by K per-env info series, episode-end gated and scaled.  base: [N] or [N, V0]."""
    _cuda(base, torch.float32, "base")
    N = base.shape[0]
    V0 = base.numel() // max(N, 1) if base.dim() > 1 else 1
    K = len(series)
    if len(episode_end) != K or (multiplier is not None and len(multiplier) != K):
        raise ValueError("episode_end / multiplier need one entry per series")
    ptrs = (C.c_void_p * max(K, 1))()
    for k, t in enumerate(series):
        _cuda(t, torch.float32, f"series[{k}]")
        if t.numel() != N:
            raise ValueError(f"series[{k}] must have one value per env")
        ptrs[k] = t.data_ptr()
    term = _as_u8(terminations, "terminations") if terminations is not None else None
    trunc = _as_u8(truncations, "truncations") if truncations is not None else None
    out = torch.empty((N, V0 + K), dtype=torch.float32, device=base.device) if out is None else _cuda(out, torch.float32, "out")
    if out.numel() != N * (V0 + K):
        raise ValueError("out must be [N, V0 + K]")
    ends = (C.c_uint8 * max(K, 1))(*[int(bool(e)) for e in episode_end])
    mult = _f32_array(multiplier) if multiplier is not None else None
    rc = _call("b200rl_reward_assemble_f32", 1, _lib.lib().b200rl_reward_assemble_f32, base.data_ptr(), V0, ptrs, K,
               _ptr(term), _ptr(trunc), ends, mult, out.data_ptr(), N, _stream())
    check(rc, "b200rl_reward_assemble_f32")
    return out


# ------------------------------------------------------------------------------------------------
# PPO arguments shared by the fused-loss entry points
@dataclass
class PpoHyper:
    """Host-side scalars of ppo/ppo.py:307-375 for one minibatch."""

    clip_range: float
    clip_range_vf: Optional[float]
    ent_coef: float
    vf_coef: Sequence[float]  # one per value head
    vf_halving: bool = False
    pi_coef: float = 1.0
    loss_scale: float = 1.0
    adv_mode: int = ADV_NORMALIZE
    adv_weights: Optional[Sequence[float]] = None
    teacher_kl_coef: float = 0.0  # > 0 with teacher_logp: adds the teacher-KL term (loss/teacher_kl_loss.py)
    teacher_unbiased: bool = True
    teacher_importance: bool = True
    vf_loss: int = 0  # VF_LOSSES[vf_loss_fn]


# ppo.py:186 vf_loss_fn = getattr(F, name) -> b200rl.h B200RL_VF_*
VF_LOSSES = {"mse_loss": 0, "huber_loss": 1, "smooth_l1_loss": 2, "l1_loss": 3}


class PpoCall:
    """Builds a b200rl_ppo_args and keeps every buffer it points at alive until the call returns."""

    def __init__(
        self,
        h: PpoHyper,
        old_logp: torch.Tensor,
        adv: torch.Tensor,
        old_values: torch.Tensor,
        returns: torch.Tensor,
        new_values: torch.Tensor,
        moments: Optional[torch.Tensor] = None,
        need_dvalues: bool = True,
        workspace_bytes: Optional[int] = None,
        teacher_logp: Optional[torch.Tensor] = None,
    ):
        B = old_logp.numel()
        dev = old_logp.device
        for name, t in (("old_logp", old_logp), ("adv", adv), ("old_values", old_values), ("returns", returns),
                        ("new_values", new_values)):  # fmt: skip
            _cuda(t, torch.float32, name)
        V = new_values.numel() // B
        adv_v = adv.numel() // B
        if old_values.numel() != B * V or returns.numel() != B * V:
            raise ValueError("old_values / returns / new_values shapes differ")
        if len(h.vf_coef) != V:
            raise ValueError(f"vf_coef has {len(h.vf_coef)} entries for {V} value heads")
        if adv_v > 1 and h.adv_weights is None:
            raise ValueError("multi-head advantages need multi_reward_weights")
        if h.adv_mode != ADV_NONE and moments is None:
            moments = adv_moments(adv.view(B, adv_v), None, h.adv_mode, h.adv_weights)
        self.B, self.V = B, V
        self.moments = moments
        self.dvalues = torch.empty_like(new_values) if need_dvalues else None
        self.stats = torch.empty(6 + 2 * V, dtype=torch.float32, device=dev)
        self._w = _f32_array(h.adv_weights) if h.adv_weights is not None else None
        self._vf = _f32_array(h.vf_coef)
        L = _lib.lib()
        self.workspace = _workspace(workspace_bytes or L.b200rl_ppo_workspace_bytes(B, V), dev)
        a = PpoArgs()
        a.old_logp, a.adv, a.moments = old_logp.data_ptr(), adv.data_ptr(), _ptr(moments)
        a.adv_weights_host = self._w
        a.adv_v, a.adv_mode = adv_v, h.adv_mode
        a.old_values, a.returns, a.new_values = old_values.data_ptr(), returns.data_ptr(), new_values.data_ptr()
        a.dvalues = _ptr(self.dvalues)
        a.V = V
        a.clip_range = float(h.clip_range)
        a.clip_range_vf = -1.0 if h.clip_range_vf is None else float(h.clip_range_vf)
        a.vf_coef_host = self._vf
        a.ent_coef, a.pi_coef = float(h.ent_coef), float(h.pi_coef)
        a.vf_halving, a.loss_scale = int(bool(h.vf_halving)), float(h.loss_scale)
        a.vf_loss = int(h.vf_loss)
        a.stats_out = self.stats.data_ptr()
        if teacher_logp is not None and h.teacher_kl_coef:
            _cuda(teacher_logp, torch.float32, "teacher_logp")
            if teacher_logp.numel() != B:
                raise ValueError("teacher_logp must be [B]")
            a.teacher_logp = teacher_logp.data_ptr()
            a.teacher_kl_coef = float(h.teacher_kl_coef)
            a.teacher_unbiased, a.teacher_importance = int(h.teacher_unbiased), int(h.teacher_importance)
        self.args = a
        self._keep = (old_logp, adv, old_values, returns, new_values, moments, teacher_logp)


@dataclass
class LossOut:
    """Device-side results of one fused loss launch.  ``stats`` layout: see include/b200rl.h."""

    stats: torch.Tensor
    dvalues: torch.Tensor
    grads: Tuple[torch.Tensor, ...]
    logp: Optional[torch.Tensor] = None
    entropy: Optional[torch.Tensor] = None


def ppo_scalar_loss(
    h: PpoHyper,
    new_logp: torch.Tensor,
    entropy: torch.Tensor,
    old_logp: torch.Tensor,
    adv: torch.Tensor,
    old_values: torch.Tensor,
    returns: torch.Tensor,
    new_values: torch.Tensor,
    moments: Optional[torch.Tensor] = None,
    kl_cutoff: Optional[float] = None,
    pi_coef_state: Optional[torch.Tensor] = None,
    teacher_logp: Optional[torch.Tensor] = None,
) -> LossOut:
    """Per-sample stage for heads that produced (new_logp [B], entropy [B] or [B, D]) elsewhere."""
    _cuda(new_logp, torch.float32, "new_logp"), _cuda(entropy, torch.float32, "entropy")
    B = new_logp.numel()
    ent_d = entropy.numel() // B
    call = PpoCall(h, old_logp, adv, old_values, returns, new_values, moments, teacher_logp=teacher_logp)
    dlogp, dent = torch.empty_like(new_logp), torch.empty_like(entropy)
    if kl_cutoff is not None and pi_coef_state is None:
        raise ValueError("kl_cutoff needs a device pi_coef_state tensor")
    rc = _call("b200rl_ppo_scalar_loss_f32", 2, _lib.lib().b200rl_ppo_scalar_loss_f32,
        new_logp.data_ptr(), entropy.data_ptr(), ent_d, B, C.byref(call.args),
        -1.0 if kl_cutoff is None else float(kl_cutoff), _ptr(pi_coef_state),
        dlogp.data_ptr(), dent.data_ptr(), call.workspace.data_ptr(), call.workspace.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_ppo_scalar_loss_f32")
    return LossOut(call.stats, call.dvalues, (dlogp, dent))


def ppo_categorical_loss(
    h: PpoHyper,
    logits: torch.Tensor,
    mask: Optional[torch.Tensor],
    actions: torch.Tensor,
    old_logp: torch.Tensor,
    adv: torch.Tensor,
    old_values: torch.Tensor,
    returns: torch.Tensor,
    new_values: torch.Tensor,
    moments: Optional[torch.Tensor] = None,
    teacher_logp: Optional[torch.Tensor] = None,
) -> LossOut:
    _cuda(logits, torch.float32, "logits")
    B, n = logits.shape
    m = _as_u8(mask, "mask")
    _cuda(actions, None, "actions")
    call = PpoCall(h, old_logp, adv, old_values, returns, new_values, moments, teacher_logp=teacher_logp)
    dlogits = torch.empty_like(logits)
    rc = _call("b200rl_ppo_categorical_loss_f32", 2, _lib.lib().b200rl_ppo_categorical_loss_f32,
        logits.data_ptr(), _ptr(m), actions.data_ptr(), _INDEX_DTYPES[actions.dtype], B, n, C.byref(call.args),
        dlogits.data_ptr(), call.workspace.data_ptr(), call.workspace.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_ppo_categorical_loss_f32")
    return LossOut(call.stats, call.dvalues, (dlogits,))


def ppo_gaussian_loss(
    h: PpoHyper,
    mu: torch.Tensor,
    log_std: torch.Tensor,
    actions: torch.Tensor,
    old_logp: torch.Tensor,
    adv: torch.Tensor,
    old_values: torch.Tensor,
    returns: torch.Tensor,
    new_values: torch.Tensor,
    moments: Optional[torch.Tensor] = None,
    teacher_logp: Optional[torch.Tensor] = None,
) -> LossOut:
    for name, t in (("mu", mu), ("log_std", log_std), ("actions", actions)):
        _cuda(t, torch.float32, name)
    B, D = mu.shape
    call = PpoCall(h, old_logp, adv, old_values, returns, new_values, moments, teacher_logp=teacher_logp)
    dmu, dls = torch.empty_like(mu), torch.empty_like(log_std)
    rc = _call("b200rl_ppo_gaussian_loss_f32", 3, _lib.lib().b200rl_ppo_gaussian_loss_f32,
        mu.data_ptr(), log_std.data_ptr(), actions.data_ptr(), B, D, C.byref(call.args), dmu.data_ptr(),
        dls.data_ptr(), call.workspace.data_ptr(), call.workspace.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_ppo_gaussian_loss_f32")
    return LossOut(call.stats, call.dvalues, (dmu, dls))


def gaussian_logp_entropy(mu: torch.Tensor, log_std: torch.Tensor, actions: torch.Tensor):
    for name, t in (("mu", mu), ("log_std", log_std), ("actions", actions)):
        _cuda(t, torch.float32, name)
    B, D = mu.shape
    logp = torch.empty(B, dtype=torch.float32, device=mu.device)
    ent = torch.empty(B, D, dtype=torch.float32, device=mu.device)
    rc = _call("b200rl_gaussian_fwd_f32", 1, _lib.lib().b200rl_gaussian_fwd_f32,
        mu.data_ptr(), log_std.data_ptr(), actions.data_ptr(), B, D, logp.data_ptr(), ent.data_ptr(), _stream()
    )
    check(rc, "b200rl_gaussian_fwd_f32")
    return logp, ent


# ------------------------------------------------------------------------------------------------
# categorical distribution-level ops
def categorical_fwd(logits: torch.Tensor, mask: Optional[torch.Tensor], actions: torch.Tensor):
    _cuda(logits, torch.float32, "logits")
    R, n = logits.shape
    m = _as_u8(mask, "mask")
    _cuda(actions, None, "actions")
    logp = torch.empty(R, dtype=torch.float32, device=logits.device)
    ent = torch.empty(R, dtype=torch.float32, device=logits.device)
    rc = _call("b200rl_categorical_fwd_f32", 1, _lib.lib().b200rl_categorical_fwd_f32,
        logits.data_ptr(), _ptr(m), actions.data_ptr(), _INDEX_DTYPES[actions.dtype], R, n, logp.data_ptr(),
        ent.data_ptr(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_categorical_fwd_f32")
    return logp, ent


def categorical_bwd(logits, mask, actions, dlogp, dent):
    R, n = logits.shape
    m = _as_u8(mask, "mask")
    dlogits = torch.empty_like(logits)
    rc = _call("b200rl_categorical_bwd_f32", 1, _lib.lib().b200rl_categorical_bwd_f32,
        logits.data_ptr(), _ptr(m), actions.data_ptr(), _INDEX_DTYPES[actions.dtype], R, n,
        _cuda(dlogp.contiguous(), torch.float32, "dlogp").data_ptr(),
        _cuda(dent.contiguous(), torch.float32, "dentropy").data_ptr(), dlogits.data_ptr(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_categorical_bwd_f32")
    return dlogits


class _CategoricalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mask, actions):
        logits = logits.contiguous()
        logp, ent = categorical_fwd(logits, mask, actions)
        ctx.save_for_backward(logits, actions)
        ctx.mask = mask
        return logp, ent

    @staticmethod
    def backward(ctx, dlogp, dent):
        logits, actions = ctx.saved_tensors
        return categorical_bwd(logits, ctx.mask, actions, dlogp, dent), None, None


def categorical_logp_entropy(logits, mask, actions):
    """Differentiable (log_prob [R], entropy [R]) of a masked categorical (categorical.py:12-54)."""
    return _CategoricalFn.apply(logits, mask, actions)


def categorical_sample(logits: torch.Tensor, mask: Optional[torch.Tensor], seed: int, offset: int,
                       offset_dev: Optional[torch.Tensor] = None):
    _cuda(logits, torch.float32, "logits")
    R, n = logits.shape
    m = _as_u8(mask, "mask")
    actions = torch.empty(R, dtype=torch.int64, device=logits.device)
    logp = torch.empty(R, dtype=torch.float32, device=logits.device)
    rc = _call("b200rl_categorical_sample_f32", 1, _lib.lib().b200rl_categorical_sample_f32,
        logits.data_ptr(), _ptr(m), R, n, seed, offset, _ptr(offset_dev), actions.data_ptr(), logp.data_ptr(), _stream()
    )
    check(rc, "b200rl_categorical_sample_f32")
    return actions, logp


# ------------------------------------------------------------------------------------------------
# GridNet
@dataclass(frozen=True)
class GridnetSpec:
    """Static description of a per-cell MultiDiscrete head (gridnet.py:39-66)."""

    nvec: Tuple[int, ...]
    gates: Tuple[Tuple[int, int, int], ...] = ()  # (head, reference head, required value)
    n_pick: int = 0

    @staticmethod
    def from_subaction_mask(nvec, subaction_mask=None, n_pick: int = 0) -> "GridnetSpec":
        gates = []
        for ref, per_head in (subaction_mask or {}).items():
            if hasattr(per_head, "reference_index"):  # already {head: ValueDependentMask}
                gates.append((int(ref), int(per_head.reference_index), int(per_head.value)))
            else:
                for head, value in per_head.items():
                    gates.append((int(head), int(ref), int(value)))
        return GridnetSpec(tuple(int(n) for n in nvec), tuple(sorted(gates)), int(n_pick))


class _GridCall:
    def __init__(self, spec: GridnetSpec, logits, mask, pick_mask, actions, pick_actions):
        if logits.dtype not in _LOGIT_DTYPES:
            raise TypeError(f"logits: expected float32 or bfloat16, got {logits.dtype}")
        _cuda(logits, None, "logits")
        A, S = len(spec.nvec), sum(spec.nvec)
        Sp = S + spec.n_pick
        # the last dim may be WIDER than a row of logits: a head that emits channel-padded NHWC (80 channels for
        # S = 78) hands its tensor over as is; columns Sp.. are ignored and get zero gradient (b200rl.h logits_ld)
        ld = logits.shape[-1]
        if ld < Sp:
            raise ValueError(f"logits last dim {ld} < sum(nvec) + n_pick = {Sp}")
        B = logits.shape[0]
        HW = logits.numel() // (B * ld) if B else 0
        self.B, self.HW, self.A, self.S, self.Sp, self.ld = B, HW, A, S, Sp, ld
        self.mask = _as_u8(mask, "mask")
        if self.mask.numel() != B * HW * S:
            raise ValueError(f"mask has {self.mask.numel()} elements, expected {B * HW * S}")
        self.pick_mask = _as_u8(pick_mask, "pick_mask") if spec.n_pick else None
        if spec.n_pick and (self.pick_mask is None or self.pick_mask.numel() != B * spec.n_pick * HW):
            raise ValueError("pick_mask must be [B, n_pick, HW]")
        self.actions, self.pick_actions = actions, pick_actions
        if actions is not None:
            _cuda(actions, None, "actions")
            if actions.numel() != B * HW * A:
                raise ValueError(f"actions has {actions.numel()} elements, expected {B * HW * A}")
        if spec.n_pick and pick_actions is not None:
            _cuda(pick_actions, None, "pick_actions")
        self._nvec = (C.c_int32 * A)(*spec.nvec)
        gate_ref, gate_val = [-1] * A, [0] * A
        for head, ref, value in spec.gates:
            gate_ref[head], gate_val[head] = ref, value
        self._gref, self._gval = (C.c_int32 * A)(*gate_ref), (C.c_int32 * A)(*gate_val)
        d = GridnetDesc()
        d.B, d.HW, d.A, d.n_pick = B, HW, A, spec.n_pick
        d.logits_dtype = _LOGIT_DTYPES[logits.dtype]
        d.act_dtype = _INDEX_DTYPES[actions.dtype] if actions is not None else _lib.U8
        d.pick_dtype = _INDEX_DTYPES[pick_actions.dtype] if pick_actions is not None else _lib.I64
        d.nvec_host, d.gate_ref_host, d.gate_val_host = self._nvec, self._gref, self._gval
        d.logits_ld = 0 if ld == Sp else ld
        self.desc = d


def gridnet_fwd(spec, logits, mask, pick_mask, actions, pick_actions):
    g = _GridCall(spec, logits, mask, pick_mask, actions, pick_actions)
    logp = torch.empty(g.B, dtype=torch.float32, device=logits.device)
    ent = torch.empty(g.B, dtype=torch.float32, device=logits.device)
    ws = _workspace(_lib.lib().b200rl_gridnet_workspace_bytes(g.B, g.HW, spec.n_pick), logits.device)
    rc = _call("b200rl_gridnet_fwd", 1, _lib.lib().b200rl_gridnet_fwd,
        C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), _ptr(g.pick_mask), actions.data_ptr(),
        _ptr(pick_actions), logp.data_ptr(), ent.data_ptr(), ws.data_ptr(), ws.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_gridnet_fwd")
    return logp, ent


def gridnet_bwd(spec, logits, mask, pick_mask, actions, pick_actions, dlogp, dent):
    g = _GridCall(spec, logits, mask, pick_mask, actions, pick_actions)
    dlogits = torch.empty_like(logits)
    dlogp = _cuda(dlogp.contiguous(), torch.float32, "dlogp")
    dent = _cuda(dent.contiguous(), torch.float32, "dentropy")
    ws = _workspace(_lib.lib().b200rl_gridnet_workspace_bytes(g.B, g.HW, spec.n_pick), logits.device)
    rc = _call("b200rl_gridnet_bwd", 1, _lib.lib().b200rl_gridnet_bwd,
        C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), _ptr(g.pick_mask), actions.data_ptr(),
        _ptr(pick_actions), dlogp.data_ptr(), dent.data_ptr(), dlogits.data_ptr(), ws.data_ptr(), ws.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_gridnet_bwd")
    return dlogits


class _GridnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, spec, mask, pick_mask, actions, pick_actions):
        logits = logits.contiguous()
        logp, ent = gridnet_fwd(spec, logits, mask, pick_mask, actions, pick_actions)
        ctx.save_for_backward(logits)
        ctx.rest = (spec, mask, pick_mask, actions, pick_actions)
        return logp, ent

    @staticmethod
    def backward(ctx, dlogp, dent):
        (logits,) = ctx.saved_tensors
        spec, mask, pick_mask, actions, pick_actions = ctx.rest
        return gridnet_bwd(spec, logits, mask, pick_mask, actions, pick_actions, dlogp, dent), None, None, None, None, None


def gridnet_logp_entropy(spec, logits, mask, pick_mask, actions, pick_actions):
    """Differentiable (log_prob [B], entropy [B]) of the GridNet distribution (gridnet.py:104-193)."""
    return _GridnetFn.apply(logits, spec, mask, pick_mask, actions, pick_actions)


class _PersistentGrad:
    """A d loss / d logits buffer that outlives the call, plus the record of the rows the last call wrote into it
    (b200rl_ppo_gridnet_loss_inplace): the next call of the same shape clears those rows instead of zero-filling the
    tensor.  One per (device, shape, dtype, head layout); allocated on first use, outside any CUDA-graph pool when the
    first use is a warm-up run, so eager calls and captured replays share it and all keep its invariant."""

    def __init__(self, logits: torch.Tensor, rows_bytes: int):
        self.dlogits = torch.empty_like(logits)
        self.rows = torch.empty(rows_bytes, dtype=torch.uint8, device=logits.device)
        self.valid = False


_persistent_grads: Dict[tuple, _PersistentGrad] = {}


def clear_caches() -> None:
    """Drop the persistent gradient buffers and per-stream workspaces (they are re-created on demand)."""
    _persistent_grads.clear()
    _workspaces.clear()


def ppo_gridnet_loss(
    h: PpoHyper,
    spec: GridnetSpec,
    logits: torch.Tensor,
    mask: torch.Tensor,
    pick_mask: Optional[torch.Tensor],
    actions: torch.Tensor,
    pick_actions: Optional[torch.Tensor],
    old_logp: torch.Tensor,
    adv: torch.Tensor,
    old_values: torch.Tensor,
    returns: torch.Tensor,
    new_values: torch.Tensor,
    moments: Optional[torch.Tensor] = None,
    want_logp: bool = False,
    teacher_logp: Optional[torch.Tensor] = None,
    inplace: bool = False,
) -> LossOut:
    """One launch: masked log-prob/entropy forward, PPO loss, backward into dlogits / dvalues.

    ``inplace``: the returned ``grads[0]`` is a buffer this module keeps across calls of the same shape and is only
    valid until the next such call (the learner consumes it in the trunk's backward right away).  The gradient is zero
    outside the rows of cells with a valid action, so the kernel then clears just the rows the previous call wrote
    instead of zero-filling the whole tensor -- same bits, a fraction of the HBM traffic."""
    g = _GridCall(spec, logits, mask, pick_mask, actions, pick_actions)
    V = new_values.numel() // max(g.B, 1)
    L = _lib.lib()
    call = PpoCall(h, old_logp, adv, old_values, returns, new_values, moments,
                   workspace_bytes=L.b200rl_ppo_gridnet_workspace_bytes(g.B, g.HW, spec.n_pick, V),
                   teacher_logp=teacher_logp)
    logp = torch.empty(g.B, dtype=torch.float32, device=logits.device) if want_logp else None
    ent = torch.empty(g.B, dtype=torch.float32, device=logits.device) if want_logp else None
    if inplace and g.B > 0:
        dev = logits.device.index if logits.device.index is not None else torch.cuda.current_device()
        key = (dev, tuple(logits.shape), logits.dtype, spec.nvec, spec.n_pick)
        st = _persistent_grads.get(key)
        if st is None:
            st = _persistent_grads[key] = _PersistentGrad(logits, L.b200rl_gridnet_rows_bytes(g.B, g.HW, spec.n_pick))
        valid, st.valid = st.valid, False  # a failed launch leaves the buffer's state unknown
        rc = _call("b200rl_ppo_gridnet_loss", 4, L.b200rl_ppo_gridnet_loss_inplace,
            C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), _ptr(g.pick_mask), actions.data_ptr(),
            _ptr(pick_actions), C.byref(call.args), st.dlogits.data_ptr(), _ptr(logp), _ptr(ent),
            call.workspace.data_ptr(), call.workspace.numel(), st.rows.data_ptr(), st.rows.numel(), int(valid), _stream(),
        )  # fmt: skip
        check(rc, "b200rl_ppo_gridnet_loss_inplace")
        st.valid = True
        return LossOut(call.stats, call.dvalues, (st.dlogits,), logp, ent)
    dlogits = torch.empty_like(logits)
    rc = _call("b200rl_ppo_gridnet_loss", 4, L.b200rl_ppo_gridnet_loss,
        C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), _ptr(g.pick_mask), actions.data_ptr(),
        _ptr(pick_actions), C.byref(call.args), dlogits.data_ptr(), _ptr(logp), _ptr(ent),
        call.workspace.data_ptr(), call.workspace.numel(), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_ppo_gridnet_loss")
    return LossOut(call.stats, call.dvalues, (dlogits,), logp, ent)


def gridnet_num_actions(spec: GridnetSpec, mask: torch.Tensor, pick_mask: Optional[torch.Tensor],
                        actions: Optional[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(cells [R] int32, picks [R] int32 or None) over R steps (rollout.py:130-180): mask [R, HW, S], pick_mask
    [R, n_pick, HW], actions [R, HW, A] (needed when the spec has gates).  See b200rl.h for what the counts are."""
    m = _as_u8(mask, "mask")
    A, S = len(spec.nvec), sum(spec.nvec)
    R = m.shape[0]
    HW = m.numel() // max(R * S, 1)
    if m.numel() != R * HW * S:
        raise ValueError(f"mask has {m.numel()} elements, not a multiple of sum(nvec) = {S} per step")
    pm = _as_u8(pick_mask, "pick_mask") if spec.n_pick else None
    if spec.gates:
        _cuda(actions, None, "actions")
        if actions.numel() != R * HW * A:
            raise ValueError(f"actions has {actions.numel()} elements, expected {R * HW * A}")
    nvec = (C.c_int32 * A)(*spec.nvec)
    gate_ref, gate_val = [-1] * A, [0] * A
    for head, ref, value in spec.gates:
        gate_ref[head], gate_val[head] = ref, value
    gref, gval = (C.c_int32 * A)(*gate_ref), (C.c_int32 * A)(*gate_val)
    d = GridnetDesc()
    d.B, d.HW, d.A, d.n_pick = R, HW, A, spec.n_pick
    d.act_dtype = _INDEX_DTYPES[actions.dtype] if (spec.gates and actions is not None) else _lib.U8
    d.nvec_host, d.gate_ref_host, d.gate_val_host = nvec, gref, gval
    cells = torch.empty(R, dtype=torch.int32, device=m.device)
    picks = torch.empty(R, dtype=torch.int32, device=m.device) if spec.n_pick else None
    if R:
        rc = _call("b200rl_gridnet_num_actions", 1, _lib.lib().b200rl_gridnet_num_actions, C.byref(d), m.data_ptr(), _ptr(pm),
                   _ptr(actions) if spec.gates else None, cells.data_ptr(), _ptr(picks), _stream())
        check(rc, "b200rl_gridnet_num_actions")
    return cells, picks


def gridnet_sample(spec: GridnetSpec, logits, mask, pick_mask, seed: int, offset: int, act_dtype=torch.uint8,
                   offset_dev: Optional[torch.Tensor] = None, wide_out: Optional[torch.Tensor] = None):
    """Sample per-cell actions (+ pick) and their joint log-prob in one launch (gridnet.py:195-207).
    ``wide_out`` ([B, HW, A] int64, optional) receives the per-cell actions a second time as int64: what a host env
    is handed, without a cast on the host."""
    g = _GridCall(spec, logits, mask, pick_mask, None, None)
    if wide_out is not None:
        _cuda(wide_out, torch.int64, "wide_out")
        if wide_out.numel() != g.B * g.HW * g.A:
            raise ValueError(f"wide_out {tuple(wide_out.shape)} does not hold [{g.B}, {g.HW}, {g.A}] actions")
    actions = torch.empty((g.B, g.HW, g.A), dtype=act_dtype, device=logits.device)
    pick = torch.empty((g.B, spec.n_pick), dtype=torch.int64, device=logits.device) if spec.n_pick else None
    logp = torch.empty(g.B, dtype=torch.float32, device=logits.device)
    g.desc.act_dtype = _INDEX_DTYPES[act_dtype]
    g.desc.pick_dtype = _lib.I64
    rc = _call("b200rl_gridnet_sample", 1, _lib.lib().b200rl_gridnet_sample,
        C.byref(g.desc), logits.data_ptr(), g.mask.data_ptr(), _ptr(g.pick_mask), seed, offset, _ptr(offset_dev),
        actions.data_ptr(), _ptr(pick), logp.data_ptr(), _ptr(wide_out), _stream(),
    )  # fmt: skip
    check(rc, "b200rl_gridnet_sample")
    return actions, pick, logp


# ------------------------------------------------------------------------------------------------
# K8: channels-last glue between the convolutions of the GridNet encoder / decoder
def nhwc_glue_supported(y: torch.Tensor) -> bool:
    """The fused bias (+ max-pool) + ReLU kernels take float32 CUDA feature maps; anything else (CPU construction
    passes, autocast's bfloat16 maps) stays on the PyTorch modules the trunk was built from."""
    return y.is_cuda and y.dtype == torch.float32 and y.dim() == 4


def _check_map(y: torch.Tensor) -> None:
    if not isinstance(y, torch.Tensor) or not y.is_cuda:
        raise _lib.B200RLError("feature map: expected a CUDA tensor (rl_algo_impls_b200 has no CPU path)")
    if y.dtype != torch.float32 or y.dim() != 4:
        raise TypeError(f"feature map: expected a float32 [N, C, H, W] tensor, got {y.dtype} {tuple(y.shape)}")


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous(memory_format=torch.channels_last)


class _BiasPoolReluFn(torch.autograd.Function):
    """relu(max_pool2d(y + bias)) on a channels-last [N, C, H, W] map in one launch; one gather launch (+ the
    two-stage bias-gradient sum) backward.  Forward bit-identical to the PyTorch sequence."""

    @staticmethod
    def forward(ctx, y, bias, kernel, stride, padding, relu):
        y = _nhwc(y)
        N, Cc, H, W = y.shape
        Ho, Wo = (H + 2 * padding - kernel) // stride + 1, (W + 2 * padding - kernel) // stride + 1
        out = torch.empty((N, Cc, Ho, Wo), dtype=y.dtype, device=y.device, memory_format=torch.channels_last)
        need = ctx.needs_input_grad[0] or (bias is not None and ctx.needs_input_grad[1])
        argmax = torch.empty((N, Ho, Wo, Cc), dtype=torch.uint8, device=y.device) if need else None
        rc = _call("b200rl_nhwc_bias_pool_relu_fwd", 1, _lib.lib().b200rl_nhwc_bias_pool_relu_fwd, y.data_ptr(), _ptr(bias),
                   out.data_ptr(), _ptr(argmax), N, H, W, Cc, kernel, stride, padding, int(relu), _stream())
        check(rc, "b200rl_nhwc_bias_pool_relu_fwd")
        if need:
            ctx.save_for_backward(argmax)
            ctx.geometry = (N, Cc, H, W, kernel, stride, padding, bias is not None and ctx.needs_input_grad[1])
        return out

    @staticmethod
    def backward(ctx, dout):
        (argmax,) = ctx.saved_tensors
        N, Cc, H, W, kernel, stride, padding, want_bias = ctx.geometry
        dout = _nhwc(dout)
        dx = torch.empty((N, Cc, H, W), dtype=dout.dtype, device=dout.device, memory_format=torch.channels_last)
        dbias = torch.empty(Cc, dtype=torch.float32, device=dout.device) if want_bias else None
        L = _lib.lib()
        ws = _workspace(L.b200rl_nhwc_bias_grad_workspace_bytes(argmax.numel() // Cc, Cc), dout.device)
        rc = _call("b200rl_nhwc_bias_pool_relu_bwd", 3 if want_bias else 1, L.b200rl_nhwc_bias_pool_relu_bwd, dout.data_ptr(),
                   argmax.data_ptr(), dx.data_ptr(), _ptr(dbias), ws.data_ptr(), ws.numel(), N, H, W, Cc, kernel, stride,
                   padding, _stream())
        check(rc, "b200rl_nhwc_bias_pool_relu_bwd")
        return dx, dbias, None, None, None, None


class _BiasReluFn(torch.autograd.Function):
    """relu(y + bias) -- or y + bias alone -- on a channels-last map in one launch, written over y (a convolution's
    output, which its own backward never reads); backward = ReLU mask + two-stage bias-gradient sum."""

    @staticmethod
    def forward(ctx, y, bias, relu):
        yc = _nhwc(y)
        N, Cc, H, W = yc.shape
        # in place when y itself is the dense channels-last buffer nobody else needs (no-grad passes included)
        out = yc if (yc is y and not (y.requires_grad and y.is_leaf)) else torch.empty_like(yc, memory_format=torch.channels_last)
        rc = _call("b200rl_nhwc_bias_relu_fwd", 1, _lib.lib().b200rl_nhwc_bias_relu_fwd, yc.data_ptr(), bias.data_ptr(),
                   out.data_ptr(), N * H * W, Cc, int(relu), _stream())
        check(rc, "b200rl_nhwc_bias_relu_fwd")
        if out is y:
            ctx.mark_dirty(y)
        ctx.relu, ctx.want_bias = relu, ctx.needs_input_grad[1]
        if relu and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _nhwc(dout)
        N, Cc, H, W = dout.shape
        out = ctx.saved_tensors[0] if ctx.relu else None
        dx = torch.empty_like(dout, memory_format=torch.channels_last) if ctx.relu else dout
        dbias = torch.empty(Cc, dtype=torch.float32, device=dout.device) if ctx.want_bias else None
        L = _lib.lib()
        ws = _workspace(L.b200rl_nhwc_bias_grad_workspace_bytes(N * H * W, Cc), dout.device)
        rc = _call("b200rl_nhwc_bias_relu_bwd", (2 if ctx.want_bias else 0) + int(ctx.relu), L.b200rl_nhwc_bias_relu_bwd,
                   dout.data_ptr(), _ptr(out), dx.data_ptr(), _ptr(dbias), ws.data_ptr(), ws.numel(), N * H * W, Cc, _stream())
        check(rc, "b200rl_nhwc_bias_relu_bwd")
        return dx, dbias, None


def bias_pool_relu(y: torch.Tensor, bias: Optional[torch.Tensor], kernel: int = 3, stride: int = 2, padding: int = 1,
                   relu: bool = True) -> torch.Tensor:
    """``relu(max_pool2d(y + bias[None, :, None, None], kernel, stride, padding))`` for a float32 CUDA map, channels-last
    in memory (gridnet_encoder.py:26-51 between its convolutions)."""
    _check_map(y)
    if bias is not None:
        _cuda(bias.detach(), torch.float32, "bias")
    return _BiasPoolReluFn.apply(y, bias, int(kernel), int(stride), int(padding), bool(relu))


def bias_relu(y: torch.Tensor, bias: torch.Tensor, relu: bool = True) -> torch.Tensor:
    """``relu(y + bias[None, :, None, None])`` (``relu=False``: the sum alone) for a float32 CUDA map, channels-last in
    memory (gridnet_decoder.py:36-53).  ``y`` -- a convolution's output -- is overwritten when it can be."""
    _check_map(y)
    _cuda(bias.detach(), torch.float32, "bias")
    return _BiasReluFn.apply(y, bias, bool(relu))


# ------------------------------------------------------------------------------------------------
# K9: channels-last glue of the squeeze U-net (float32 / bfloat16 maps)
_ACTS = {"none": 0, "relu": 1, "gelu": 2}
_MAP_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def act_glue_supported(y: torch.Tensor) -> bool:
    return y.is_cuda and y.dtype in _MAP_DTYPES and y.dim() == 4


class _BiasActFn(torch.autograd.Function):
    """act(y + bias) on a channels-last map in one launch.  y -- a convolution's bias-free output -- is what the
    backward keeps: the pre-activation is recomputed from it, no second map is saved."""

    @staticmethod
    def forward(ctx, y, bias, act):
        y = _nhwc(y)
        N, Cc, H, W = y.shape
        out = torch.empty_like(y, memory_format=torch.channels_last)
        rc = _call("b200rl_nhwc_bias_act_fwd", 1, _lib.lib().b200rl_nhwc_bias_act_fwd, y.data_ptr(), bias.data_ptr(),
                   out.data_ptr(), N * H * W, Cc, act, _MAP_DTYPES[y.dtype], _stream())
        check(rc, "b200rl_nhwc_bias_act_fwd")
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(y, bias)
            ctx.act, ctx.want_bias = act, ctx.needs_input_grad[1]
        return out

    @staticmethod
    def backward(ctx, dout):
        y, bias = ctx.saved_tensors
        N, Cc, H, W = y.shape
        dout = _nhwc(dout if dout.dtype == y.dtype else dout.to(y.dtype))
        dx = torch.empty_like(y, memory_format=torch.channels_last)
        dbias = torch.empty(Cc, dtype=torch.float32, device=y.device) if ctx.want_bias else None
        L = _lib.lib()
        ws = _workspace(L.b200rl_nhwc_bias_act_workspace_bytes(N * H * W, Cc), y.device)
        rc = _call("b200rl_nhwc_bias_act_bwd", 3 if ctx.want_bias else 1, L.b200rl_nhwc_bias_act_bwd, dout.data_ptr(),
                   y.data_ptr(), bias.data_ptr(), dx.data_ptr(), _ptr(dbias), ws.data_ptr(), ws.numel(), N * H * W, Cc,
                   ctx.act, _MAP_DTYPES[y.dtype], _stream())
        check(rc, "b200rl_nhwc_bias_act_bwd")
        return dx, dbias, None


def bias_act(y: torch.Tensor, bias: torch.Tensor, act: str = "gelu") -> torch.Tensor:
    """``act(y + bias[None, :, None, None])`` for a float32 / bfloat16 CUDA map, channels-last in memory; ``bias`` float32
    (rounded to the map's dtype as autocast would).  squeeze_unet.py's convolution -> GELU pairs."""
    if not act_glue_supported(y):
        raise TypeError(f"feature map: expected a float32 / bfloat16 CUDA [N, C, H, W] tensor, got {y.dtype} {tuple(y.shape)}")
    _cuda(bias.detach(), torch.float32, "bias")
    return _BiasActFn.apply(y, bias, _ACTS[act])


class _SeTailFn(torch.autograd.Function):
    """gelu(x + (y2 + b2) * sigmoid(W2 gelu(W1 mean_hw(y2 + b2)))): the tail of an SE-residual block (double_cone.py:18-86)
    after its second convolution.  Forward: a per-(sample, channel) sum pass, the two small linears (library GEMMs on
    [N, C]), one elementwise pass.  Backward: one reduction pass (the gate's gradient), the linears' backward on [N, C],
    one elementwise pass (+ the bias column sums)."""

    @staticmethod
    def forward(ctx, x, y2, b2, w1, w2):
        x, y2 = _nhwc(x), _nhwc(y2)
        N, Cc, H, W = y2.shape
        T, dt, L = y2.dtype, _MAP_DTYPES[y2.dtype], _lib.lib()
        sums = torch.empty((N, Cc), dtype=torch.float32, device=y2.device)
        ws = _workspace(L.b200rl_se_workspace_bytes(N, H * W, Cc), y2.device)
        check(_call("b200rl_se_mean_sums", 2, L.b200rl_se_mean_sums, y2.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(),
                    N, H * W, Cc, dt, _stream()), "b200rl_se_mean_sums")
        m = (sums / (H * W) + b2.to(T).float()).to(T)                 # what autocast hands the first linear
        h1 = torch.nn.functional.linear(m, w1.to(T))
        g1 = torch.nn.functional.gelu(h1)
        gate = torch.sigmoid(torch.nn.functional.linear(g1, w2.to(T))).contiguous()
        out = torch.empty_like(y2, memory_format=torch.channels_last)
        check(_call("b200rl_se_tail_fwd", 1, L.b200rl_se_tail_fwd, x.data_ptr(), y2.data_ptr(), b2.data_ptr(), gate.data_ptr(),
                    out.data_ptr(), N, H * W, Cc, dt, _stream()), "b200rl_se_tail_fwd")
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, y2, b2, w1, w2, m, h1, g1, gate)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y2, b2, w1, w2, m, h1, g1, gate = ctx.saved_tensors
        N, Cc, H, W = y2.shape
        dt, L = _MAP_DTYPES[y2.dtype], _lib.lib()
        dout = _nhwc(dout if dout.dtype == y2.dtype else dout.to(y2.dtype))
        dgate = torch.empty((N, Cc), dtype=torch.float32, device=y2.device)
        dx = torch.empty_like(y2, memory_format=torch.channels_last)
        ws = _workspace(L.b200rl_se_workspace_bytes(N, H * W, Cc), y2.device)
        check(_call("b200rl_se_tail_gate_grad", 2, L.b200rl_se_tail_gate_grad, dout.data_ptr(), x.data_ptr(), y2.data_ptr(),
                    b2.data_ptr(), gate.data_ptr(), dgate.data_ptr(), dx.data_ptr(), ws.data_ptr(), ws.numel(), N, H * W, Cc, dt,
                    _stream()), "b200rl_se_tail_gate_grad")
        # the gate's two linears backward, float32 on [N, C] / [N, C / 16]
        sf, g1f, h1f, mf = gate.float(), g1.float(), h1.float(), m.float()
        dh2 = dgate * sf * (1.0 - sf)
        dw2 = dh2.t() @ g1f
        dh1 = torch.ops.aten.gelu_backward(dh2 @ w2.float(), h1f)
        dw1 = dh1.t() @ mf
        dmean = ((dh1 @ w1.float()) / (H * W)).contiguous()
        dy2 = torch.empty_like(y2, memory_format=torch.channels_last)
        db2 = torch.empty(Cc, dtype=torch.float32, device=y2.device) if ctx.needs_input_grad[2] else None
        ws = _workspace(L.b200rl_nhwc_bias_act_workspace_bytes(N * H * W, Cc), y2.device)
        check(_call("b200rl_se_tail_bwd", 3 if db2 is not None else 1, L.b200rl_se_tail_bwd, dx.data_ptr(), gate.data_ptr(),
                    dmean.data_ptr(), dy2.data_ptr(), _ptr(db2), ws.data_ptr(), ws.numel(), N, H * W, Cc, dt, _stream()),
              "b200rl_se_tail_bwd")
        return dx, dy2, db2, dw1.to(w1.dtype), dw2.to(w2.dtype)


def se_tail(x: torch.Tensor, y2: torch.Tensor, b2: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """The tail of an SE-residual block: ``gelu(x + SE(y2 + b2))`` with ``SE(u) = u * sigmoid(w2 gelu(w1 mean_hw(u)))``
    (double_cone.py:18-86).  x: the block's input; y2: its second convolution's bias-free output (same dtype / shape)."""
    if not act_glue_supported(y2) or x.dtype != y2.dtype or x.shape != y2.shape:
        raise TypeError(f"se_tail: maps {x.dtype} {tuple(x.shape)} / {y2.dtype} {tuple(y2.shape)}")
    _cuda(b2.detach(), torch.float32, "b2")
    return _SeTailFn.apply(x, y2, b2, w1, w2)
