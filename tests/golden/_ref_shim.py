"""The import shim for the unmodified reference lives in oracle/ref_shim.py (bench.py's reference arm needs it on
the GPU box too); re-exported here for the fixture generators."""
from oracle.ref_shim import *  # noqa: F401,F403
from oracle.ref_shim import Box, DictSpace, Discrete, MultiDiscrete, Space, install  # noqa: F401
