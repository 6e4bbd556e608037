// dropin_demo.cpp — the drop-in, exercised with the REFERENCE'S OWN types.  TEST INFRASTRUCTURE ONLY.
//
// INTEGRATION.md §1 tells a reference maintainer to replace src/EnumerationSolver.h (an empty stub) by
// simplexmethod_b200/cpp/include/EnumerationSolver.h and to link libenumgpu.so.  This program is that build:
// oracle/Makefile (target "dropin") copies our header into a scratch include directory that shadows the stub,
// compiles it together with the reference's unmodified Canonical / Symmetrical / Common / SymmetricalParser /
// SimplexSolover (from /root/reference, against oracle/eigen_shim in place of the absent Eigen) and runs the
// reference user's flow — README.md:40-42, "compare the answers of SimplexSolver and EnumerationSolver":
//
//     SymmetricalParser::ParseFromFile -> Symmetrical::ToCanonical -> Solver(*canonical).solve()
//                                                                   -> EnumerationSolver(*canonical).solve()   [GPU]
//
// usage: dropin_demo <symmetric LP file>      prints one machine-readable line per fact
#include <cstdio>
#include <exception>

#include "EnumerationSolver.h"      // ours, in place of the reference's stub
#include "SimplexSolover.h"         // the reference's
#include "SymmetricalParser.h"      // the reference's

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <symmetric LP file>\n", argv[0]); return 2; }
    SymmetricalParser parser;
    auto sym = parser.ParseFromFile(argv[1]);
    if (!sym) { std::fprintf(stderr, "parse error: %s\n", parser.GetLastError().c_str()); return 3; }
    auto canonical = sym->ToCanonical();
    try {
        Solver simplex(*canonical);
        const Eigen::VectorXd xs = simplex.solve();
        std::printf("simplex_x");
        for (Eigen::Index j = 0; j < xs.size(); ++j) std::printf(" %.17g", xs[j]);
        std::printf("\n");
        EnumerationSolver enumerator(*canonical);
        const Eigen::VectorXd xe = enumerator.solve();
        std::printf("enumeration_x");
        for (Eigen::Index j = 0; j < xe.size(); ++j) std::printf(" %.17g", xe[j]);
        std::printf("\nobjective %.17g\nbasis", enumerator.objective());
        for (int j : enumerator.optimalBasis()) std::printf(" %d", j);
        std::printf("\ncounts %llu %llu %llu %llu\n", (unsigned long long)enumerator.basesEvaluated(),
                    (unsigned long long)enumerator.singularCount(), (unsigned long long)enumerator.infeasibleCount(),
                    (unsigned long long)enumerator.feasibleCount());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
