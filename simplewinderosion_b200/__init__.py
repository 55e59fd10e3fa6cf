"""B200-native wind-erosion engine: the erode-cycle hot path of weigert/SimpleWindErosion
(Wind::fly, World::erode, World::cascade) as hand-written CUDA for sm_100a behind a C-ABI
(include/swe.h, libswe_b200.so). This package is the thin host-side mirror of the
reference's World/Wind interface over that ABI; importing it loads the CUDA library and
fails if it is missing -- there is no CPU fallback."""
from ._capi import LIB_PATH, lib  # noqa: F401
from .world import *  # noqa: F401,F403
from .world import __all__ as _world_all

__all__ = list(_world_all) + ["lib", "LIB_PATH"]
