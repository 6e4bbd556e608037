// enum_demo.cpp — the reference user's flow on the GPU path:
//   SymmetricalParser::ParseFromFile -> Symmetrical::ToCanonical -> EnumerationSolver(problem).solve()
// usage: enum_demo <lp file>      prints one machine-readable line per fact.
//        enum_demo --main         the flow of the reference's demo (src/main.cpp:48-113): the LP hard-coded
//                                 there as Common -> ToSymmetrical -> GetDual / ToCanonical -> GetBasicSolution,
//                                 Evaluate, IsFeasibleBasis -> solver (EnumerationSolver in place of Solver)
//        enum_demo --lab          a general-form LP of the lab's shape (README.md:5-8: 5 variables, 3 with a sign
//                                 restriction, 3 inequalities + 1 equality, no zero coefficient): the primal and
//                                 its Common::GetDual(), both through ToCanonical() and the GPU enumeration
#include <cstdio>
#include <cstring>
#include <exception>

#include "EnumerationSolver.h"
#include "ProblemTypes/Common.h"
#include "SymmetricalParser.h"

namespace {
Eigen::MatrixXd mat(int r, int c, std::initializer_list<double> rowmajor)
{
    Eigen::MatrixXd m(r, c);
    int k = 0;
    for (double v : rowmajor) { m(k / c, k % c) = v; ++k; }
    return m;
}
Eigen::VectorXd vec(std::initializer_list<double> il)
{
    Eigen::VectorXd v(static_cast<Eigen::Index>(il.size()));
    Eigen::Index k = 0;
    for (double x : il) v[k++] = x;
    return v;
}
void print_vec(const char* name, const Eigen::VectorXd& x)
{
    std::printf("%s", name);
    for (Eigen::Index j = 0; j < x.size(); ++j) std::printf(" %.17g", x[j]);
    std::printf("\n");
}
void solve_and_print(const char* tag, const Canonical& can)
{
    EnumerationSolver solver(can);
    const Eigen::VectorXd x = solver.solve();
    char name[64];
    std::snprintf(name, sizeof name, "%s_x", tag);
    print_vec(name, x);
    std::printf("%s_objective %.17g\n%s_basis", tag, solver.objective(), tag);
    for (int j : solver.optimalBasis()) std::printf(" %d", j);
    std::printf("\n%s_counts %llu %llu %llu %llu\n", tag, (unsigned long long)solver.basesEvaluated(),
                (unsigned long long)solver.singularCount(), (unsigned long long)solver.infeasibleCount(),
                (unsigned long long)solver.feasibleCount());
}

int demo_main_cpp()
{
    using CT = Common::ConstraintType;
    using VT = Common::VariableType;
    Common common(mat(2, 3, {1, 1, 1, 2, 1, 0}), vec({6, 8}), vec({3, 2, 4}), {CT::LessOrEqual, CT::LessOrEqual},
                  {VT::NonNegative, VT::NonNegative, VT::NonNegative}, true);
    auto symmetrical = common.ToSymmetrical();
    auto dual = symmetrical->GetDual();
    auto canonical = symmetrical->ToCanonical();
    std::printf("dual_shape %d %d %d\n", (int)dual->GetConstraintsMatrix().rows(), (int)dual->GetConstraintsMatrix().cols(),
                dual->IsMaximization() ? 1 : 0);
    const Eigen::VectorXd x0 = canonical->GetBasicSolution();
    print_vec("initial_x", x0);
    std::printf("initial_z %.17g\ninitial_basis_feasible %d\n", canonical->Evaluate(x0), canonical->IsFeasibleBasis() ? 1 : 0);
    solve_and_print("primal", *canonical);
    // The dual is solved from the general form: Common::ToSymmetrical turns "min b'y, A'y >= c" into
    // "max -b'y, -A'y <= -c", whose canonical form has slacks only.  (Symmetrical::ToCanonical of a min
    // problem adds zero-cost artificial columns, reference Symmetrical.cpp:191-222; enumerating THAT is a
    // relaxation whose optimum is 0 — the reference leaves Big-M to its simplex solver.)
    solve_and_print("dual", *common.GetDual()->ToCanonical());
    return 0;
}

int demo_lab()
{
    using CT = Common::ConstraintType;
    using VT = Common::VariableType;
    // max 3x1 + 5x2 + x3 + 2x4 - 6x5;  x1,x2,x3 >= 0, x4 free, x5 <= 0
    Common primal(mat(4, 5, {2, 1, 1, 1, -3,
                             1, 3, 2, 2, -1,
                             1, 1, 4, 1, -2,
                             1, 1, 1, 1, -1}),
                  vec({12, 15, 16, 7}), vec({3, 5, 1, 2, -6}),
                  {CT::LessOrEqual, CT::LessOrEqual, CT::GreaterOrEqual, CT::Equal},
                  {VT::NonNegative, VT::NonNegative, VT::NonNegative, VT::Free, VT::NonPositive}, true);
    solve_and_print("primal", *primal.ToCanonical());
    solve_and_print("dual", *primal.GetDual()->ToCanonical());
    return 0;
}
}  // namespace

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: %s <symmetric LP file> | --main | --lab\n", argv[0]); return 2; }
    if (!std::strcmp(argv[1], "--main") || !std::strcmp(argv[1], "--lab")) {
        try {
            return !std::strcmp(argv[1], "--main") ? demo_main_cpp() : demo_lab();
        } catch (const std::exception& e) {
            std::fprintf(stderr, "error: %s\n", e.what());
            return 1;
        }
    }
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
