"""rl_algo_impls_b200: the PPO data path of rl-algo-impls as hand-written sm_100a CUDA kernels.

Host side mirrors the reference's Python interface for the path (``Batch`` / ``Rollout`` /
``RolloutGenerator`` / ``PPO`` / the action distributions); the kernels live in
``libb200rl.so`` behind the C ABI of ``include/b200rl.h``.  There is no CPU path.
"""
from . import _lib

__version__ = "0.1.0"
__all__ = ["_lib", "ops"]
