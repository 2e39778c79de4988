"""ctypes binding of libb200rl.so (include/b200rl.h).

There is no CPU path: if the shared library is missing, importing succeeds (so that host-only
logic stays testable) but the first kernel call raises ``RuntimeError``; with the library
present and no CUDA device the kernels themselves return B200RL_ECUDA and we raise.
"""
import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rl.so")
REPO_ROOT = os.path.dirname(_HERE)

F32, BF16, U8, I32, I64 = 0, 1, 2, 3, 4
MAX_HEADS, MAX_VALUE_HEADS, MAX_GATHER = 16, 32, 16

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


class PpoArgs(C.Structure):
    """b200rl_ppo_args"""

    _fields_ = [
        ("old_logp", C.c_void_p),
        ("adv", C.c_void_p),
        ("moments", C.c_void_p),
        ("adv_weights_host", c_f32p),
        ("adv_v", C.c_int64),
        ("adv_mode", C.c_int),
        ("old_values", C.c_void_p),
        ("returns", C.c_void_p),
        ("new_values", C.c_void_p),
        ("dvalues", C.c_void_p),
        ("V", C.c_int64),
        ("clip_range", C.c_double),
        ("clip_range_vf", C.c_double),
        ("vf_coef_host", c_f32p),
        ("ent_coef", C.c_float),
        ("pi_coef", C.c_float),
        ("vf_halving", C.c_int),
        ("loss_scale", C.c_float),
        ("stats_out", C.c_void_p),
        ("teacher_logp", C.c_void_p),
        ("teacher_kl_coef", C.c_float),
        ("teacher_unbiased", C.c_int),
        ("teacher_importance", C.c_int),
        ("vf_loss", C.c_int),
    ]


class GridnetDesc(C.Structure):
    """b200rl_gridnet_desc"""

    _fields_ = [
        ("B", C.c_int64),
        ("HW", C.c_int64),
        ("A", C.c_int),
        ("n_pick", C.c_int),
        ("logits_dtype", C.c_int),
        ("act_dtype", C.c_int),
        ("pick_dtype", C.c_int),
        ("nvec_host", c_i32p),
        ("gate_ref_host", c_i32p),
        ("gate_val_host", c_i32p),
        ("logits_ld", C.c_int64),
    ]


class StorePack(C.Structure):
    """b200rl_store_pack"""

    _fields_ = [("field", C.c_int), ("N", C.c_int64), ("C", C.c_int64), ("HW", C.c_int64), ("Cp", C.c_int64)]


_vp, _i64, _int, _sz, _u64, _f = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_uint64, C.c_float

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py
PROTOTYPES = {
    "b200rl_version": (_int, []),
    "b200rl_last_error": (C.c_char_p, []),
    "b200rl_host_register": (_int, [_vp, _sz]),
    "b200rl_host_unregister": (_int, [_vp]),
    "b200rl_gae_scan_f32": (_int, [_vp, _vp, _vp, _vp, _vp, c_f64p, c_f64p, _int, _vp, _vp, _i64, _i64, _i64, _vp]),
    "b200rl_gae_segments_f32": (
        _int,
        [_vp, _vp, _vp, _vp, _vp, _vp, _vp, c_f64p, c_f64p, _int, _vp, _vp, _i64, _i64, _vp],
    ),
    "b200rl_adv_moments_workspace_bytes": (_sz, [_i64, _i64]),
    "b200rl_adv_moments_f64": (_int, [_vp, _vp, _i64, _i64, _int, c_f32p, _vp, _vp, _sz, _vp]),
    "b200rl_adv_normalize_f32": (_int, [_vp, _vp, _i64, _i64, _int, c_f32p, _vp, _vp, _i64, _vp]),
    "b200rl_gather_rows": (_int, [C.POINTER(_vp), C.POINTER(_vp), c_i64p, _int, _vp, _i64, _i64, _vp]),
    "b200rl_ppo_workspace_bytes": (_sz, [_i64, _i64]),
    "b200rl_ppo_scalar_loss_f32": (
        _int,
        [_vp, _vp, _i64, _i64, C.POINTER(PpoArgs), _f, _vp, _vp, _vp, _vp, _sz, _vp],
    ),
    "b200rl_categorical_fwd_f32": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _vp, _vp, _vp]),
    "b200rl_categorical_bwd_f32": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "b200rl_ppo_categorical_loss_f32": (
        _int,
        [_vp, _vp, _vp, _int, _i64, _i64, C.POINTER(PpoArgs), _vp, _vp, _sz, _vp],
    ),
    "b200rl_gaussian_fwd_f32": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "b200rl_ppo_gaussian_loss_f32": (_int, [_vp, _vp, _vp, _i64, _i64, C.POINTER(PpoArgs), _vp, _vp, _vp, _sz, _vp]),
    "b200rl_gridnet_workspace_bytes": (_sz, [_i64, _i64, _int]),
    "b200rl_ppo_gridnet_workspace_bytes": (_sz, [_i64, _i64, _int, _i64]),
    "b200rl_gridnet_fwd": (_int, [C.POINTER(GridnetDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200rl_gridnet_bwd": (_int, [C.POINTER(GridnetDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200rl_ppo_gridnet_loss": (
        _int,
        [C.POINTER(GridnetDesc), _vp, _vp, _vp, _vp, _vp, C.POINTER(PpoArgs), _vp, _vp, _vp, _vp, _sz, _vp],
    ),
    "b200rl_gridnet_rows_bytes": (_sz, [_i64, _i64, _int]),
    "b200rl_ppo_gridnet_loss_inplace": (
        _int,
        [C.POINTER(GridnetDesc), _vp, _vp, _vp, _vp, _vp, C.POINTER(PpoArgs), _vp, _vp, _vp, _vp, _sz, _vp, _sz, _int, _vp],
    ),
    "b200rl_gridnet_num_actions": (_int, [C.POINTER(GridnetDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200rl_gridnet_sample": (_int, [C.POINTER(GridnetDesc), _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200rl_categorical_sample_f32": (_int, [_vp, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _vp]),
    "b200rl_running_norm_obs_f32": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _int, C.c_double, C.c_double, _vp, _vp]),
    "b200rl_running_norm_reward_f32": (
        _int,
        [_vp, _vp, _i64, _i64, C.c_double, _vp, _vp, _vp, _vp, _int, C.c_double, C.c_double, _vp, _vp],
    ),
    "b200rl_running_norm_reward_ema_f32": (
        _int,
        [_vp, _vp, _i64, _i64, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, _int, _int, C.c_double,
         C.c_double, _vp, _vp],
    ),
    "b200rl_rollout_store_step": (_int, [C.POINTER(_vp), C.POINTER(_vp), c_i64p, _int, _vp, _i64, _vp]),
    "b200rl_rollout_store_step_carry": (
        _int,
        [C.POINTER(_vp), C.POINTER(_vp), c_i64p, C.POINTER(_vp), _int, _vp, _i64, _vp],
    ),
    "b200rl_rollout_store_step_fused": (
        _int,
        [C.POINTER(_vp), C.POINTER(_vp), c_i64p, C.POINTER(_vp), C.POINTER(_vp), _int, C.POINTER(StorePack), _vp, _i64,
         _int, _vp, _vp],
    ),
    "b200rl_reward_assemble_f32": (
        _int,
        [_vp, _i64, C.POINTER(_vp), _int, _vp, _vp, C.POINTER(C.c_uint8), c_f32p, _vp, _i64, _vp],
    ),
    "b200rl_nhwc_bias_act_workspace_bytes": (_sz, [_i64, _i64]),
    "b200rl_nhwc_bias_act_fwd": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _int, _vp]),
    "b200rl_nhwc_bias_act_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _int, _int, _vp]),
    "b200rl_se_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "b200rl_se_mean_sums": (_int, [_vp, _vp, _vp, _sz, _i64, _i64, _i64, _int, _vp]),
    "b200rl_se_tail_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "b200rl_se_tail_gate_grad": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _i64, _int, _vp]),
    "b200rl_se_tail_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _i64, _int, _vp]),
    "b200rl_h2d_batch": (_int, [_int, C.POINTER(_vp), C.POINTER(_vp), c_i64p, _vp]),
    "b200rl_nhwc_bias_grad_workspace_bytes": (_sz, [_i64, _i64]),
    "b200rl_nhwc_bias_pool_relu_fwd": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _int, _vp]),
    "b200rl_nhwc_bias_pool_relu_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _i64, _i64, _int, _int, _int, _vp]),
    "b200rl_nhwc_bias_relu_fwd": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "b200rl_nhwc_bias_relu_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _i64, _i64, _vp]),
}

_lib: Optional[C.CDLL] = None


class B200RLError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libb200rl.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", REPO_ROOT, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise B200RLError(f"building libb200rl.so failed (exit {res.returncode})")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200RLError(
                f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()). "
                "rl_algo_impls_b200 has no CPU or eager-PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200rl_last_error()
        raise B200RLError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
