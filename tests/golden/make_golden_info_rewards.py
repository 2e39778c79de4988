"""Golden fixture for the multi-head reward assembly from the LIVE, UNMODIFIED
reference's InfoRewardsWrapper (wrappers/info_rewards_wrapper.py):

    python tests/golden/make_golden_info_rewards.py        (build container only)

Asserts that oracle/rewards.py reproduces the reference bit for bit and stores inputs + reference outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import _ref_shim  # noqa: E402

_ref_shim.install()

import gymnasium.experimental.vector.vector_env as _gym_vec  # noqa: E402  (fabricated by the shim: gymnasium is absent)


class _MiniVectorWrapper:
    """The two behaviours of gymnasium's VectorWrapper the reference class relies on: keep `env`, delegate `step`."""

    def __init__(self, env):
        self.env = env

    def step(self, action):
        return self.env.step(action)


_gym_vec.VectorWrapper = _MiniVectorWrapper

from rl_algo_impls.wrappers.info_rewards_wrapper import InfoRewardsWrapper as RefInfoRewards  # noqa: E402

from oracle.rewards import assemble_rewards  # noqa: E402

CASES = {
    # name: (N, base heads, paths, episode_end, multiplier)
    "lux_like": (16, 1, [["stats", "ice"], ["stats", "ore"], ["stats", "power"], ["score"]], [False, False, False, True],
                 [0.01, 0.02, 0.001, 1.0]),
    "all_episode_end": (8, 1, [["a"], ["b"]], True, None),
    "scalar_multiplier": (8, 1, [["a"], ["b"], ["c"]], False, 0.5),
    "multi_base": (8, 3, [["a"]], [True], [2.0]),
}


class _Env:
    def __init__(self, script):
        self.script, self.t = script, 0
        self.num_envs = script[0][1].shape[0]

    @property
    def unwrapped(self):
        return self

    def step(self, action):
        out = self.script[self.t]
        self.t += 1
        return out


def nest(paths, arrays):
    infos = {}
    for path, a in zip(paths, arrays):
        d = infos
        for key in path[:-1]:
            d = d.setdefault(key, {})
        d[path[-1]] = a
    return infos


if __name__ == "__main__":
    rng = np.random.default_rng(77)
    out = {}
    for name, (N, V0, paths, episode_end, multiplier) in CASES.items():
        steps = 5
        script, series_all = [], []
        for _ in range(steps):
            r = rng.standard_normal((N,) if V0 == 1 else (N, V0)).astype(np.float32)
            series = [rng.standard_normal(N).astype(np.float32) * 10 for _ in paths]
            term, trunc = rng.random(N) < 0.2, rng.random(N) < 0.1
            script.append((np.zeros((N, 2), np.float32), r, term, trunc, nest(paths, series)))
            series_all.append(series)
        ref = RefInfoRewards(_Env(script), paths, episode_end=episode_end, multiplier=multiplier)
        got = []
        for t in range(steps):
            _, rewards, _, _, _ = ref.step(None)
            want = assemble_rewards(script[t][1], [s.copy() for s in series_all[t]], script[t][2], script[t][3],
                                    ref.episode_end, ref.multiplier)
            assert rewards.dtype == want.dtype and np.array_equal(rewards, want), f"oracle != reference: {name} step {t}"
            got.append(rewards)
        out[f"{name}.base"] = np.stack([s[1] for s in script])
        out[f"{name}.series"] = np.stack([np.stack(s) for s in series_all])  # [steps, K, N]
        out[f"{name}.terminations"] = np.stack([s[2] for s in script])
        out[f"{name}.truncations"] = np.stack([s[3] for s in script])
        out[f"{name}.episode_end"] = np.asarray(ref.episode_end)
        if ref.multiplier is not None:
            out[f"{name}.multiplier"] = np.asarray(ref.multiplier)
        out[f"{name}.rewards"] = np.stack(got)
        print(name, got[0].shape, got[0].dtype)
    path = os.path.join(HERE, "info_rewards.npz")
    np.savez_compressed(path, **out)
    print(f"wrote info_rewards.npz ({os.path.getsize(path) / 1024:.1f} KB)")
