"""Regenerates tests/golden/tiny_lps.json with the exact-rational enumerator.

    python tests/golden/make_goldens.py

Inputs are the reference's own fixtures restated in simplexmethod_b200/lpgen.py
(input_symmetric.txt -> ToCanonical; src/main.cpp:48-57; tests/test_canonical.cpp
:12-22) plus Beale's LP and a README-shaped lab LP.  Every number is produced by
oracle/exact.py (fractions.Fraction Gauss-Jordan) — independent of the
floating-point oracle and of the CUDA kernels — and stored as "p/q" strings.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import exact  # noqa: E402
from simplexmethod_b200 import lpgen  # noqa: E402

CASES = {
    "lab_symmetric": (lpgen.lab_symmetric_canonical, 3),
    "main_cpp": (lpgen.main_cpp_canonical, 3),
    "test_canonical": (lpgen.test_canonical_fixture, 4),
    "beale": (lpgen.beale_lp, 7),
    "readme_shaped": (lpgen.readme_shaped_lp, 5),
}


def main():
    out = {}
    for name, (fn, n_orig) in CASES.items():
        A, b, c, mx = fn()
        ex = exact.enumerate_exact(A.tolist(), b.tolist(), c.tolist(), mx)
        out[name] = dict(
            A=A.tolist(), b=b.tolist(), c=c.tolist(), maximize=mx, n_orig=n_orig,
            status=ex["status"],
            z=[None if z is None else str(z) for z in ex["z"]],
            x=[None if x is None else [str(v) for v in x] for x in ex["x"]],
            n_singular=ex["n_singular"], n_infeasible=ex["n_infeasible"], n_feasible=ex["n_feasible"],
            best_rank=ex["best_rank"], best_basis=ex["best_basis"],
            best_x=[str(v) for v in ex["best_x"]], best_z=str(ex["best_z"]),
            tied_ranks=ex["tied_ranks"],
        )
    with open(os.path.join(HERE, "tiny_lps.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
