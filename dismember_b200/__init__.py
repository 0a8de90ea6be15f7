"""dismember_b200 -- B200-native (sm_100a) engine for dismember's tree-retrieval hot path.

Layout: csrc/ (CUDA kernels + C ABI, built in-tree into libdismember_gpu.so),
_capi.py (ctypes binding = what the JNI shim binds), tdm.py / otm.py / dr.py (host
mirrors of the reference's recommend API), formats/ (tree/mapping/model files).
"""
from ._capi import (DmgArgumentError, DmgError, DmgIndexError, Engine, declared_symbols,  # noqa: F401
                    load_library)

__all__ = ["Engine", "DmgError", "DmgIndexError", "DmgArgumentError", "load_library", "declared_symbols"]
