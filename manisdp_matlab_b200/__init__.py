"""manisdp_matlab_b200 -- B200-native engine for ManiSDP's inner hot path.

Public surface = the reference's primal drivers with unchanged signatures (solvers.py) on top of the C ABI in
include/manisdp_b200.h (libmanisdp_b200.so, built from csrc/ by build.py).
"""
from .solvers import (ManiDSDP_unitdiag, ManiSDP, ManiSDP_multiblock, ManiSDP_onlyunitdiag, ManiSDP_unitdiag,  # noqa: F401
                      ManiSDP_unittrace)
from ._lib import Handle, GroupHandle, EngineError, load, LIB_PATH  # noqa: F401

__all__ = ["ManiDSDP_unitdiag", "ManiSDP", "ManiSDP_multiblock", "ManiSDP_onlyunitdiag", "ManiSDP_unitdiag", "ManiSDP_unittrace", "Handle", "EngineError", "load"]
