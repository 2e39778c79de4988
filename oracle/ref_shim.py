"""Import shim for the *unmodified* reference (rl-algo-impls).  TEST / BASELINE INFRASTRUCTURE ONLY.

Two users: the fixture generators (``tests/golden/make_golden*.py``), which import the reference from
``/root/reference`` in the build container, and ``bench.py --impl reference`` / the reference-backed GPU tests,
which import the byte-identical copy of its Python sources that ``oracle/make_ref.sh`` places under ``oracle/_ref/``
(git-ignored; it travels to the GPU box with the snapshot; ``oracle/ref_manifest.sha256`` pins its content).
Nothing under ``rl_algo_impls_b200/`` imports this module.

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

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIPPED_ROOT = os.path.join(_HERE, "_ref")


def reference_root() -> str:
    """$RL_ALGO_IMPLS_REFERENCE, else /root/reference (build container), else oracle/_ref (GPU box)."""
    env = os.environ.get("RL_ALGO_IMPLS_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/rl_algo_impls"):
        return "/root/reference"
    return SHIPPED_ROOT


REFERENCE_ROOT = reference_root()


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


def available(root: str = None) -> bool:
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, "rl_algo_impls"))


def install(root: str = None) -> str:
    """Make ``import rl_algo_impls`` resolve to the unmodified reference; returns the root used."""
    global _installed, REFERENCE_ROOT
    if _installed:
        return REFERENCE_ROOT
    root = root or REFERENCE_ROOT
    if not available(root):
        raise RuntimeError(f"reference not found at {root}: run oracle/make_ref.sh where /root/reference exists")
    REFERENCE_ROOT = root
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, root)
    _installed = True
    return root
