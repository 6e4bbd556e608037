// enum_demo.cpp — the reference user's flow on the GPU path:
//   SymmetricalParser::ParseFromFile -> Symmetrical::ToCanonical -> EnumerationSolver(problem).solve()
// usage: enum_demo <lp file>      prints one machine-readable line per fact.
#include <cstdio>
#include <exception>

#include "EnumerationSolver.h"
#include "SymmetricalParser.h"

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <symmetric LP file>\n", argv[0]); return 2; }
    SymmetricalParser parser;
    auto sym = parser.ParseFromFile(argv[1]);
    if (!sym) { std::fprintf(stderr, "parse error: %s\n", parser.GetLastError().c_str()); return 3; }
    auto can = sym->ToCanonical();
    try {
        std::printf("initial_basis_feasible %d\n", can->IsFeasibleBasis() ? 1 : 0);
        const Eigen::VectorXd x0 = can->GetBasicSolution();
        std::printf("initial_x");
        for (Eigen::Index j = 0; j < x0.size(); ++j) std::printf(" %.17g", x0[j]);
        std::printf("\n");
        EnumerationSolver solver(*can);
        const Eigen::VectorXd x = solver.solve();
        std::printf("x");
        for (Eigen::Index j = 0; j < x.size(); ++j) std::printf(" %.17g", x[j]);
        std::printf("\nobjective %.17g\nbasis", solver.objective());
        for (int j : solver.optimalBasis()) std::printf(" %d", j);
        std::printf("\nbest_rank %llu\ncounts %llu %llu %llu %llu\n", (unsigned long long)solver.bestRank(),
                    (unsigned long long)solver.basesEvaluated(), (unsigned long long)solver.singularCount(),
                    (unsigned long long)solver.infeasibleCount(), (unsigned long long)solver.feasibleCount());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
