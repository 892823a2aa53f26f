"""B200-native render hot path for OpenRayAi (leonZtiger/Ray-Tracer-engine).

Only what the path needs: `csrc/` (sm_100a kernels + the C ABI of include/ore_render.h),
`host/` (C++ shims that keep the reference's onStart()/update()/memManager/sprite symbols),
`scene.py` (harness-defined synthetic inputs), `capi.py` (ctypes over the C ABI) and
`multigpu.py` (row-band sharding across the GPUs of one box).
"""
from . import build, multigpu, scene  # noqa: F401
from . import capi  # noqa: F401,E402
from .capi import OreError, Renderer, load_library  # noqa: F401,E402
