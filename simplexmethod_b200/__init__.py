"""simplexmethod_b200 — B200-native extreme-point enumeration (libenumgpu).

Public surface: ``Canonical`` and ``EnumerationSolver`` (the reference's
interface for this path, see solver.py), ``Symmetrical`` / ``Common`` /
``SymmetricalParser`` (the forms and the text format that feed it, problem_types.py), ``lpgen`` (synthetic LPs) and the raw
C ABI via ``_lib.lib()`` / ``_abi``.
"""
from . import _abi, lpgen  # noqa: F401
from ._lib import EnumGpuError, lib, last_error  # noqa: F401
from .solver import Canonical, EnumerationSolver  # noqa: F401
from .problem_types import Common, ConstraintType, Symmetrical, SymmetricalParser, VariableType  # noqa: F401

__all__ = ["Canonical", "Symmetrical", "Common", "ConstraintType", "VariableType", "SymmetricalParser", "EnumerationSolver", "EnumGpuError", "lpgen", "lib", "last_error"]
