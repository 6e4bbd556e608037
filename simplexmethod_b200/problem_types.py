"""Host-side mirror of the reference's other problem types and its parser.

``Symmetrical`` (reference: src/ProblemTypes/Symmetrical.h:17-45), ``Common``
(src/ProblemTypes/Common.h:10-55) and ``SymmetricalParser``
(src/SymmetricalParser.h:13-55) keep the reference's method names and argument
meaning; they exist so that reference user code —

    ParseFromFile -> ToCanonical -> EnumerationSolver(problem).solve()
    Common(...).ToSymmetrical().GetDual() ...

— reads the same here.  They only reshape data (no numerics): everything they
produce ends in a ``Canonical`` that the GPU enumeration takes.  The C++
versions are under simplexmethod_b200/cpp.
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence

import numpy as np

from .solver import Canonical


def _checked(A, b, c):
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    c = np.asarray(c, dtype=np.float64).reshape(-1)
    if A.ndim != 2:
        raise ValueError("A must be a matrix")
    if A.shape[0] != b.size:
        raise ValueError("dimensions of A and b do not match")
    if A.shape[1] != c.size:
        raise ValueError("dimensions of A and c do not match")
    return np.asfortranarray(A), np.ascontiguousarray(b), np.ascontiguousarray(c)


def _terms(values, sep: str) -> str:
    """"c1*x1 + c2*x2-c3*x3": " + " only in front of a non-negative coefficient; numbers as operator<<(double) prints them."""
    out = []
    for j, v in enumerate(values):
        out.append((" + " if j > 0 and v >= 0 else "") + f"{v:g}{sep}x{j + 1}")
    return "".join(out)


def _lp_text(title, maximize, A, b, c, heading, sep, rel) -> str:
    """Header, objective and rows in the layout of the reference's Print() methods (Common.cpp:86-137,
    Symmetrical.cpp:70-97, Canonical.cpp:88-123); rel(i) is the text between row i and its right-hand side."""
    lines = [title, ("Максимизировать: " if maximize else "Минимизировать: ") + _terms(c, "*"), "", heading]
    lines += [_terms(A[i], sep) + rel(i) + f"{b[i]:g}" for i in range(A.shape[0])]
    return "\n".join(lines) + "\n"


class _Problem:
    def Print(self):
        print(self.PrintText(), end="")

    def GetConstraintsMatrix(self): return self._A
    def GetRightHandSide(self): return self._b
    def GetObjectiveCoefficients(self): return self._c
    def IsMaximization(self): return self._maximize

    def Evaluate(self, solution) -> float:
        x = np.asarray(solution, dtype=np.float64).reshape(-1)
        if x.size != self._c.size:
            raise ValueError("solution size differs from the number of variables")
        return float(self._c @ x)


class Symmetrical(_Problem):
    """max c'x, Ax <= b, x >= 0   or   min c'x, Ax >= b, x >= 0."""

    def __init__(self, A, b, c, maximize: bool):
        self._A, self._b, self._c = _checked(A, b, c)      # Symmetrical.cpp:17-29
        self._maximize = bool(maximize)

    def PrintText(self) -> str:
        rel = " <= " if self._maximize else " >= "
        return _lp_text("=== Симметричная форма задачи ЛП ===", self._maximize, self._A, self._b, self._c,
                        "При ограничениях:", "*", lambda i: rel) + "\nВсе переменные неотрицательны: x_i >= 0\n"

    def GetDual(self) -> "Symmetrical":
        """Transpose A, swap b and c, flip the sense (Symmetrical.cpp:119-140)."""
        return Symmetrical(self._A.T, self._c, self._b, not self._maximize)

    def ToCanonical(self) -> Canonical:
        """max: [A | I] with the slack basis; min: [A | -I | I] with surplus columns and a
        zero-cost artificial basis (Symmetrical.cpp:142-223).  n_orig = n."""
        m, n = self._A.shape
        eye = np.eye(m)
        if self._maximize:
            Ac = np.hstack([self._A, eye])
            basis = [n + i for i in range(m)]
        else:
            Ac = np.hstack([self._A, -eye, eye])
            basis = [n + m + i for i in range(m)]
        cc = np.concatenate([self._c, np.zeros(Ac.shape[1] - n)])
        out = Canonical(Ac, self._b, cc, basis, minimize=not self._maximize)
        out.SetOriginalVariablesCount(n)
        return out

    def ToCommon(self) -> "Common":
        """Same data; every row <= (max) or >= (min), every variable >= 0 (Symmetrical.cpp:225-273)."""
        m, n = self._A.shape
        row = ConstraintType.LessOrEqual if self._maximize else ConstraintType.GreaterOrEqual
        return Common(self._A, self._b, self._c, [row] * m, [VariableType.NonNegative] * n, self._maximize)


class ConstraintType(enum.Enum):
    LessOrEqual = 0
    GreaterOrEqual = 1
    Equal = 2


class VariableType(enum.Enum):
    Free = 0
    NonNegative = 1
    NonPositive = 2


class Common(_Problem):
    """General form: rows are <=, >= or =; variables free, >= 0 or <= 0."""

    ConstraintType = ConstraintType
    VariableType = VariableType

    def __init__(self, A, b, c, constraintTypes: Sequence[ConstraintType], variableTypes: Sequence[VariableType],
                 maximize: bool):
        self._A, self._b, self._c = _checked(A, b, c)      # Common.cpp:29-44
        if len(constraintTypes) != self._A.shape[0]:
            raise ValueError("number of constraint types differs from the number of rows of A")
        if len(variableTypes) != self._A.shape[1]:
            raise ValueError("number of variable types differs from the number of columns of A")
        self._rows = [ConstraintType(t) for t in constraintTypes]
        self._vars = [VariableType(t) for t in variableTypes]
        self._maximize = bool(maximize)

    def PrintText(self) -> str:
        rel = {ConstraintType.LessOrEqual: "<=", ConstraintType.GreaterOrEqual: ">=", ConstraintType.Equal: "="}
        dom = {VariableType.Free: "∈R", VariableType.NonNegative: " >= 0", VariableType.NonPositive: " <= 0"}
        return _lp_text("=== Общая форма задачи ЛП ===", self._maximize, self._A, self._b, self._c,
                        "При ограничениях:", "", lambda i: rel[self._rows[i]]) + "\nОграничения на переменные:\n" + \
            "".join(f"x{j + 1}: {dom[t]}\n" for j, t in enumerate(self._vars))

    def GetConstraintTypes(self) -> List[ConstraintType]: return self._rows
    def GetVariableTypes(self) -> List[VariableType]: return self._vars

    def ToSymmetrical(self) -> Symmetrical:
        """Always "max, <=" (Common.cpp:169-348): free x_j -> (x', x''), x_j <= 0 -> -x_j,
        a >= row is negated, an = row becomes (row, -row), min costs are negated."""
        col_blocks, cost = [], []
        for j, vt in enumerate(self._vars):
            col = self._A[:, j]
            if vt is VariableType.NonNegative:
                col_blocks.append(col); cost.append(self._c[j])
            elif vt is VariableType.NonPositive:
                col_blocks.append(-col); cost.append(-self._c[j])
            else:
                col_blocks += [col, -col]; cost += [self._c[j], -self._c[j]]
        wide = np.column_stack(col_blocks)
        rows, rhs = [], []
        for i, rt in enumerate(self._rows):
            if rt is ConstraintType.LessOrEqual:
                rows.append(wide[i]); rhs.append(self._b[i])
            elif rt is ConstraintType.GreaterOrEqual:
                rows.append(-wide[i]); rhs.append(-self._b[i])
            else:
                rows += [wide[i], -wide[i]]; rhs += [self._b[i], -self._b[i]]
        cs = np.array(cost)
        return Symmetrical(np.vstack(rows), np.array(rhs), cs if self._maximize else -cs, True)

    def ToCanonical(self) -> Canonical:
        return self.ToSymmetrical().ToCanonical()           # Common.cpp:351-362

    def GetDual(self) -> "Common":
        """Transpose, swap b and c, flip the sense; row kinds become variable kinds and
        vice versa (Common.cpp:366-448)."""
        mx = self._maximize
        dual_vars = [VariableType.Free if t is ConstraintType.Equal else
                     VariableType.NonNegative if (t is ConstraintType.LessOrEqual) == mx else VariableType.NonPositive
                     for t in self._rows]
        dual_rows = [ConstraintType.Equal if t is VariableType.Free else
                     ConstraintType.GreaterOrEqual if (t is VariableType.NonNegative) == mx else ConstraintType.LessOrEqual
                     for t in self._vars]
        return Common(self._A.T, self._c, self._b, dual_rows, dual_vars, not mx)


class SymmetricalParser:
    """Text format of the reference (SymmetricalParser.cpp:44-193): a sense line
    (maximize | max | minimize | min), ``objective:`` followed by coefficient rows,
    ``constraints:`` / ``subject to:`` followed by rows "a_1 ... a_n rhs"; '#' starts a
    comment.  Errors give ``None`` + ``GetLastError()``, never an exception."""

    def __init__(self):
        self._last_error = ""

    def GetLastError(self) -> str:
        return self._last_error

    def ParseFromFile(self, filename: str) -> Optional[Symmetrical]:
        try:
            with open(filename, "r", encoding="utf-8") as f:
                text = f.read()
        except OSError:
            self._last_error = "cannot open file: " + filename
            return None
        return self.ParseFromString(text)

    def ParseFromString(self, content: str) -> Optional[Symmetrical]:
        where, maximize = None, True
        obj: List[float] = []
        rows: List[List[float]] = []
        for raw in content.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line in ("maximize", "max"):
                maximize = True
            elif line in ("minimize", "min"):
                maximize = False
            elif line in ("objective:", "objective"):
                where = "objective"
            elif line in ("constraints:", "constraints", "subject to:", "subject to"):
                where = "constraints"
            else:
                vals = []
                for tok in line.split():          # like operator>>: stop at the first non-number
                    try:
                        vals.append(float(tok))
                    except ValueError:
                        break
                if where is None:
                    self._last_error = "data outside of a section: " + line
                    return None
                if where == "objective":
                    obj += vals
                elif len(vals) < 2:
                    self._last_error = "not enough numbers in constraint: " + line
                    return None
                else:
                    rows.append(vals)
        if not obj:
            self._last_error = "objective is missing"
            return None
        if not rows:
            self._last_error = "constraints are missing"
            return None
        if any(len(r) != len(obj) + 1 for r in rows):
            self._last_error = "constraint width differs from the objective"
            return None
        data = np.array(rows)
        return Symmetrical(data[:, :-1], data[:, -1], np.array(obj), maximize)
