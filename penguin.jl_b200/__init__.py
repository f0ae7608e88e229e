"""penguin.jl_b200 -- B200-native (sm_100a, fp64) implementation of Penguin.jl's unsteady cut-cell diffusion hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI ``libpenguin_b200.so``) and ``api.py``, the
host-side mirror of the reference's Julia API.  The directory name contains a dot, so import it as ``penguin_b200``
(the loader module at the repo root).
"""
from .api import *  # noqa: F401,F403
from . import api, _lib, slab, build as _build  # noqa: F401

build = _build.build
