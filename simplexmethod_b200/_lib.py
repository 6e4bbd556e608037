"""Loader for the in-tree CUDA library (simplexmethod_b200/libenumgpu.so).

There is no fallback of any kind: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
simplexmethod_b200/csrc``) importing the solver raises, and without a CUDA
device every solve call returns ENUMGPU_ERR_CUDA, which the wrappers turn into
``EnumGpuError``.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libenumgpu.so")
_LIB = None


class EnumGpuError(RuntimeError):
    """CUDA / library failure (no device, launch error, missing build)."""


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EnumGpuError(
                f"{LIB_PATH} is missing: build it with `make -C simplexmethod_b200/csrc` "
                "(libenumgpu has no CPU fallback)")
        _LIB = _abi.bind(C.CDLL(LIB_PATH))
    return _LIB


def last_error() -> str:
    return lib().enumgpu_last_error().decode("utf-8", "replace")
