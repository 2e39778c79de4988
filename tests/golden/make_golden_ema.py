"""Golden fixture for NormalizeReward(exponential_moving_mean_var=True) from the LIVE, UNMODIFIED reference
(wrappers/normalize.py:62-110 over utils/running_mean_std.py HybridMovingMeanVar):

    python tests/golden/make_golden_ema.py        (build container only)

Asserts that oracle/normalize.py reproduces the reference bit for bit and stores inputs + reference outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _ref_shim  # noqa: E402

_ref_shim.install()

import gymnasium.experimental.vector.vector_env as _gym_vec  # noqa: E402  (fabricated by the shim)


class _MiniVectorWrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


_gym_vec.VectorWrapper = _MiniVectorWrapper

from rl_algo_impls.wrappers.normalize import NormalizeReward as RefNormRew  # noqa: E402

from oracle.normalize import RewardNormalizer  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(91)
    out = {}
    for tag, N, shape, window, steps in (("scalar", 16, (), 100, 12), ("multi", 8, (3,), 40.5, 10)):
        class Env:
            num_envs = N

        rew = rng.standard_normal((steps, N) + shape).astype(np.float32) * 3
        dones = rng.random((steps, N)) < 0.1
        ref = RefNormRew(Env(), gamma=0.98, shape=shape, exponential_moving_mean_var=True, emv_window_size=window)
        ours = RewardNormalizer(N, shape, gamma=0.98, exponential_moving_mean_var=True, emv_window_size=window)
        outs = []
        for t in range(steps):
            want = ref.normalize(rew[t])
            ref.returns[dones[t]] = 0
            got = ours.step(rew[t], dones[t])
            assert np.array_equal(got, want), f"oracle != reference: ema {tag} step {t}"
            outs.append(want)
        out.update({f"{tag}.rewards": rew, f"{tag}.dones": dones, f"{tag}.out": np.stack(outs),
                    f"{tag}.window": np.float64(window), f"{tag}.var": np.asarray(ref.rms.var),
                    f"{tag}.ema_mean": np.asarray(ref.rms.emmv.mean), f"{tag}.ema_var": np.asarray(ref.rms.emmv.var),
                    f"{tag}.rms_var": np.asarray(ref.rms.rms.var), f"{tag}.count": np.float64(ref.rms.rms.count),
                    f"{tag}.returns": ref.returns})
        print(tag, outs[0].dtype, float(ref.rms.rms.count) / window)
    path = os.path.join(HERE, "normalizers_ema.npz")
    np.savez_compressed(path, **out)
    print(f"wrote normalizers_ema.npz ({os.path.getsize(path) / 1024:.1f} KB)")
