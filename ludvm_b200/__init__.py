"""ludvm_b200 -- B200-native (sm_100a) implementation of the vortex-velocity hot path of jcatalang/LUDVM.

Importing the package never touches the oracle and never falls back to the CPU: every numerical entry point
goes through libludvm_b200.so (hand-written CUDA behind the C ABI of include/ludvm_b200.h).
"""
from . import _lib, ops  # noqa: F401
from ._lib import Context, LudvmError  # noqa: F401

__all__ = ["Context", "LudvmError", "ops", "LUDVM"]


def __getattr__(name):
    if name == "LUDVM":
        from . import ludvm as _m
        return _m.LUDVM
    if name in ("generate_free_vortices", "generate_free_single_vortex", "generate_flowfield_vortices",
                "generate_flowfield_turbulence"):
        from . import freevort as _f
        return getattr(_f, name)
    if name in ("sweep", "sharded", "freevort"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
