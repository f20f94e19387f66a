"""yunikorn_k8shim_b200 -- B200-native engine for YuniKorn's pod->node allocation hot path.

Only what the path needs: csrc/ (sm_100a kernels, host engine, the C ABI of include/ykgpu.h),
engine.py / dictionary.py / podres.py (ctypes stubs of that ABI), synth.py (seeded synthetic snapshots of the BASELINE configs),
build.py (nvcc recipe).  Importing the package does not load the library; Engine() does, and fails
loudly if it is missing or no CUDA device is present.
"""
from .engine import Engine, YkError, load_library, LIB_PATH, EXPORTS  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["Engine", "YkError", "load_library", "LIB_PATH", "EXPORTS", "synth"]
