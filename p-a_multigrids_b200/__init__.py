"""pamg-b200: B200-native (sm_100a CUDA) implementation of the P-A_multigrids DG smoother / multigrid hot path.

Import with ``importlib.import_module("p-a_multigrids_b200")`` (the directory name is not a Python identifier).
"""
from . import build as _build  # noqa: F401
from .pamg import *  # noqa: F401,F403
from .pamg import (GAUSS_SEIDEL, JACOBI, LIB_PATH, RES, RHS, RICHARDSON, TNEW, TNONLIN, TOLD, Mesh, Params,  # noqa: F401
                   PamgError, PinnedBuffer, SemiImplicitIterative, default_params, device_count, get_unique_id, halo_plan, lib)
