"""Oracle: GAE(lambda) advantages and returns.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates ``rl_algo_impls/shared/gae.py:97-124`` (``compute_advantages``) and the
returns add of ``rl_algo_impls/rollout/vec_rollout.py:88``.

Arithmetic facts of the reference that the CUDA kernel reproduces bit for bit
(checked against the live reference by tests/golden/make_golden.py):

* ``1.0 - episode_starts[t + 1]`` has a bool operand, so the non-terminal factor is
  **float64** (gae.py:118) and ``delta`` / the running ``last_gae_lam`` are carried
  in float64; the result is rounded to float32 only by the store into
  ``advantages[t]`` (gae.py:123).
* a Python-float ``gamma`` is a weak scalar: ``gamma * next_value`` is rounded to
  **float32** before it meets the float64 factor (gae.py:121).  A per-value-head
  ``gamma`` (``np.ndarray`` float64, gae.py:109-110) makes that product float64.
* evaluation order: ``((r + (gamma*nv)*nt) - v)`` and
  ``delta + ((gamma*lambda)*nt)*last``.
"""
from typing import Tuple, Union

import numpy as np

NumOrArray = Union[float, np.ndarray]


def _head_broadcast(x: NumOrArray, trailing_shape: Tuple[int, ...]) -> NumOrArray:
    """gae.py:109-112 via shared/tensor_utils.py:25-32: a [V] array lines up with the last dims."""
    if isinstance(x, np.ndarray):
        assert x.shape == trailing_shape[len(trailing_shape) - x.ndim:], (x.shape, trailing_shape)
        return x.reshape((1,) * (len(trailing_shape) - x.ndim) + x.shape)
    return x


def gae_advantages(
    rewards: np.ndarray,  # [T, N] or [T, N, V] float32
    values: np.ndarray,  # same shape, float32
    episode_starts: np.ndarray,  # [T, N] bool
    next_episode_starts: np.ndarray,  # [N] bool
    next_values: np.ndarray,  # [N] or [N, V] float32
    gamma: NumOrArray,
    gae_lambda: NumOrArray,
) -> np.ndarray:
    T = rewards.shape[0]
    per_step_shape = values.shape[1:]
    gamma = _head_broadcast(gamma, per_step_shape)
    gae_lambda = _head_broadcast(gae_lambda, per_step_shape)
    out = np.zeros_like(rewards)
    carry = 0
    for t in range(T - 1, -1, -1):
        started = next_episode_starts if t == T - 1 else episode_starts[t + 1]
        v_next = next_values if t == T - 1 else values[t + 1]
        alive = 1.0 - started  # float64 (gae.py:115/118)
        alive = alive.reshape(alive.shape + (1,) * (v_next.ndim - alive.ndim))
        delta = rewards[t] + gamma * v_next * alive - values[t]
        carry = delta + gamma * gae_lambda * alive * carry
        out[t] = carry  # float32 store (gae.py:123)
    return out


def gae_returns(advantages: np.ndarray, values: np.ndarray) -> np.ndarray:
    """vec_rollout.py:88 -- float32 add."""
    return advantages + values


def discrete_skips_advantages(
    rewards: np.ndarray,  # [T] or [T, V] float32
    values: np.ndarray,
    steps_elapsed: np.ndarray,  # [T] int32
    done: bool,
    next_values,  # [V] / scalar, used when not done
    gamma: NumOrArray,
    gae_lambda: NumOrArray,
) -> np.ndarray:
    """rollout/discrete_skips_trajectory_builder.py:84-100 (the trajectory() recurrence)."""
    advantages = np.zeros_like(rewards)
    last_advantage = np.zeros_like(advantages[-1])
    n_steps = advantages.shape[0]
    gamma = _head_broadcast(gamma, values.shape[1:])
    gae_lambda = _head_broadcast(gae_lambda, values.shape[1:])
    for t in reversed(range(n_steps)):
        if t == n_steps - 1:
            next_value = np.zeros_like(values[t]) if done else next_values
        else:
            next_value = values[t + 1]
        delta = rewards[t] + gamma ** steps_elapsed[t] * next_value - values[t]
        last_advantage = delta + gamma ** steps_elapsed[t] * gae_lambda * last_advantage
        advantages[t] = last_advantage
    return advantages
