"""Import shim for the *unmodified* reference (rl-algo-impls) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` to import the
reference from ``/root/reference`` and generate golden fixtures.  ``/root/reference``
does not exist on the GPU box, so nothing in tests/, smoke() or bench.py imports
this module at run time -- only the fixture generator does, here.

The reference imports ``gymnasium`` / ``stable_baselines3`` / ``matplotlib`` at module
scope; none is installed.  A ``sys.meta_path`` finder fabricates those packages with
just enough real behaviour (mini ``spaces`` classes) for the hot-path modules to run.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
import typing

import numpy as np

REFERENCE_ROOT = os.environ.get("RL_ALGO_IMPLS_REFERENCE", "/root/reference")


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = shape, dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.low = np.broadcast_to(np.asarray(low, dtype), shape)
        self.high = np.broadcast_to(np.asarray(high, dtype), shape)
        super().__init__(tuple(shape), np.dtype(dtype))

    def sample(self):
        return np.zeros(self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        super().__init__((), np.dtype(np.int64))


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, np.int64)
        super().__init__(self.nvec.shape, np.dtype(np.int64))

    def __len__(self):
        return len(self.nvec)


class DictSpace(Space):
    def __init__(self, d):
        self.spaces = dict(d)
        super().__init__()

    def __getitem__(self, k):
        return self.spaces[k]

    def items(self):
        return self.spaces.items()

    def keys(self):
        return self.spaces.keys()


class _Auto(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in ("ObsType", "ArrayType", "ActType"):
            value = typing.TypeVar(name)
        else:
            value = type(
                name,
                (),
                {
                    "__init__": lambda self, *a, **k: None,
                    "__class_getitem__": classmethod(lambda c, i: c),
                },
            )
        setattr(self, name, value)
        return value


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = ("gymnasium", "stable_baselines3", "matplotlib", "moviepy")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Auto(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        if m.__name__ == "gymnasium.spaces":
            m.Space, m.Box, m.Discrete = Space, Box, Discrete
            m.MultiDiscrete, m.Dict = MultiDiscrete, DictSpace
        if m.__name__ == "gymnasium":
            m.Space = Space
        if m.__name__ == "stable_baselines3.common.preprocessing":
            m.get_flattened_obs_dim = lambda sp: int(np.prod(sp.shape))


_installed = False


def install():
    """Make ``import rl_algo_impls`` resolve to the unmodified reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(
            f"reference not found at {REFERENCE_ROOT}; the golden generator only runs in the build container"
        )
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
