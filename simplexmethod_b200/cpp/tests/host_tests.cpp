// host_tests.cpp — CPU-only checks of the host types, restating what the
// reference's gtest suite pins for them (reference: tests/test_canonical.cpp:
// 31-39,68-76; tests/test_symmetrical.cpp:27-97; tests/test_parser.cpp:4-81;
// tests/test_common.cpp:39-94; tests/test_transformations.cpp:6-61).
// Exit code 0 = all passed.  Needs no GPU (no numerical method is called).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include "EnumerationSolver.h"
#include "ProblemTypes/Canonical.h"
#include "ProblemTypes/Common.h"
#include "ProblemTypes/Symmetrical.h"
#include "SymmetricalParser.h"

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)
template <class E, class F> static bool throws(F&& f) { try { f(); } catch (const E&) { return true; } catch (...) { return false; } return false; }

static Eigen::MatrixXd mat(int r, int c, std::initializer_list<double> rowmajor)
{
    Eigen::MatrixXd m(r, c);
    int k = 0;
    for (double v : rowmajor) { m(k / c, k % c) = v; ++k; }
    return m;
}

static Eigen::VectorXd vec(std::initializer_list<double> il)
{
    Eigen::VectorXd v(static_cast<Eigen::Index>(il.size()));
    Eigen::Index k = 0;
    for (double x : il) v[k++] = x;
    return v;
}

// the three Print() layouts on a fixture with negative, zero and fractional coefficients (compared with the
// reference's own output by tests/test_reference_code.py)
static void print_fixture()
{
    using CT = Common::ConstraintType;
    using VT = Common::VariableType;
    const Eigen::MatrixXd A = mat(3, 3, {1, -2, 0.5, 0, 4, -6, 7, 8.25, 9});
    const Eigen::VectorXd b = vec({10, -11, 1e-7}), c = vec({1, -2, 0});
    Common(A, b, c, {CT::LessOrEqual, CT::GreaterOrEqual, CT::Equal}, {VT::Free, VT::NonNegative, VT::NonPositive}, false).Print();
    std::cout << "---\n";
    Symmetrical(A, b, c, true).Print();
    std::cout << "---\n";
    Symmetrical(A, b, c, false).Print();
    std::cout << "---\n";
    Canonical can(mat(2, 4, {1, -2, 1, 0, 3, 4.5, 0, 1}), vec({5, 6}), vec({7, -8, 0, 0}), {2, 3}, false);
    can.SetOriginalVariablesCount(2);
    can.Print();
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string(argv[1]) == "--print") { print_fixture(); return 0; }
    if (argc > 2 && std::string(argv[1]) == "--parse") {        // one line per file: "null" or "maximize m n A(row-major) b c"
        for (int k = 2; k < argc; ++k) {
            SymmetricalParser fp;
            auto sym = fp.ParseFromFile(argv[k]);
            if (!sym) { std::puts("null"); continue; }
            const Eigen::MatrixXd& M = sym->GetConstraintsMatrix();
            std::printf("%d %d %d", sym->IsMaximization() ? 1 : 0, (int)M.rows(), (int)M.cols());
            for (Eigen::Index i = 0; i < M.rows(); ++i) for (Eigen::Index j = 0; j < M.cols(); ++j) std::printf(" %.17g", M(i, j));
            for (Eigen::Index i = 0; i < M.rows(); ++i) std::printf(" %.17g", sym->GetRightHandSide()[i]);
            for (Eigen::Index j = 0; j < M.cols(); ++j) std::printf(" %.17g", sym->GetObjectiveCoefficients()[j]);
            std::puts("");
        }
        return 0;
    }
    // --- Canonical: fixture of tests/test_canonical.cpp:12-22
    const Eigen::MatrixXd A = mat(2, 4, {1, 2, 1, 0, 3, 4, 0, 1});
    Eigen::VectorXd b(2); b[0] = 5; b[1] = 6;
    Eigen::VectorXd c(4); c[0] = 7; c[1] = 8; c[2] = 0; c[3] = 0;
    Canonical can(A, b, c, {2, 3}, true);
    CHECK(can.GetConstraintsMatrix().rows() == 2 && can.GetConstraintsMatrix().cols() == 4);
    CHECK(can.GetRightHandSide().size() == 2 && can.GetObjectiveCoefficients().size() == 4);
    CHECK(!can.IsMaximization());
    CHECK(can.GetBasisIndices().size() == 2 && can.GetBasisIndices()[0] == 2);
    CHECK(can.GetOriginalVariablesCount() == 4);
    CHECK(can.GetConstraintsMatrix().data()[1] == 3.0);                 // column-major: A(1,0)
    CHECK(throws<std::invalid_argument>([&] { Canonical bad(A, b, c, {2, 10}, true); }));
    CHECK(throws<std::invalid_argument>([&] { Canonical bad(A, b, c, {2}, true); }));
    CHECK(throws<std::invalid_argument>([&] { can.SetOriginalVariablesCount(0); }));
    Eigen::VectorXd x(4); x[0] = 1; x[1] = 2; x[2] = 0; x[3] = 0;
    CHECK(can.Evaluate(x) == 23.0);                                     // 7*1 + 8*2, cf. test_common.cpp:57-58
    CHECK(throws<std::invalid_argument>([&] { can.Evaluate(b); }));
    Canonical copy(can);                                               // value semantics
    copy.SetOriginalVariablesCount(2);
    CHECK(can.GetOriginalVariablesCount() == 4 && copy.GetOriginalVariablesCount() == 2);

    {   // dual of the canonical fixture: 4 rows x (2*2 + 4) columns, max, slack basis (Canonical.cpp:305-364)
        auto d = can.GetDual();
        CHECK(d->GetConstraintsMatrix().rows() == 4 && d->GetConstraintsMatrix().cols() == 8 && d->IsMaximization());
        CHECK(d->GetConstraintsMatrix()(1, 0) == 2.0 && d->GetConstraintsMatrix()(1, 2) == -2.0 && d->GetConstraintsMatrix()(1, 5) == 1.0);
        CHECK(d->GetRightHandSide()[1] == 8.0 && d->GetObjectiveCoefficients()[0] == 5.0 && d->GetObjectiveCoefficients()[3] == -6.0);
        CHECK(d->GetBasisIndices()[0] == 4 && d->GetOriginalVariablesCount() == 4);
    }

    // --- Symmetrical -> Canonical (tests/test_symmetrical.cpp:64-71, 83-84) and the dual (:47-52)
    const Eigen::MatrixXd As = mat(2, 2, {1, 2, 3, 4});
    Eigen::VectorXd bs(2); bs[0] = 5; bs[1] = 6;
    Eigen::VectorXd cs(2); cs[0] = 7; cs[1] = 8;
    Symmetrical smax(As, bs, cs, true);
    auto cmax = smax.ToCanonical();
    CHECK(cmax->GetConstraintsMatrix().cols() == 4 && cmax->IsMaximization());
    CHECK(cmax->GetBasisIndices()[0] == 2 && cmax->GetBasisIndices()[1] == 3);
    CHECK(cmax->GetOriginalVariablesCount() == 2);
    CHECK(cmax->GetConstraintsMatrix()(0, 2) == 1.0 && cmax->GetConstraintsMatrix()(1, 2) == 0.0 && cmax->GetConstraintsMatrix()(1, 3) == 1.0);
    Symmetrical smin(As, bs, cs, false);
    auto cmin = smin.ToCanonical();
    CHECK(cmin->GetConstraintsMatrix().cols() == 6 && !cmin->IsMaximization());
    CHECK(cmin->GetConstraintsMatrix()(0, 2) == -1.0 && cmin->GetConstraintsMatrix()(0, 4) == 1.0);
    CHECK(cmin->GetBasisIndices()[0] == 4 && cmin->GetObjectiveCoefficients()[4] == 0.0);
    auto dual = smax.GetDual();
    CHECK(!dual->IsMaximization() && dual->GetConstraintsMatrix()(0, 1) == 3.0);
    CHECK(dual->GetRightHandSide()[1] == 8.0 && dual->GetObjectiveCoefficients()[0] == 5.0);
    CHECK(throws<std::invalid_argument>([&] { Symmetrical bad(As, cs, c, true); }));

    // --- Common (tests/test_common.cpp:39-94): fixture A=[1 2;3 4], b=(5,6), c=(7,8), rows (<=, >=), x >= 0, max
    using CT = Common::ConstraintType;
    using VT = Common::VariableType;
    {
        Common com(As, bs, cs, {CT::LessOrEqual, CT::GreaterOrEqual}, {VT::NonNegative, VT::NonNegative}, true);
        CHECK(com.IsMaximization() && com.GetConstraintsMatrix().rows() == 2 && com.GetConstraintsMatrix().cols() == 2);
        CHECK(com.GetRightHandSide().size() == 2 && com.GetObjectiveCoefficients().size() == 2);
        Eigen::VectorXd x12(2); x12[0] = 1; x12[1] = 2;
        CHECK(com.Evaluate(x12) == 23.0);
        CHECK(throws<std::invalid_argument>([&] { com.Evaluate(c); }));
        Eigen::VectorXd b3(3);
        CHECK(throws<std::invalid_argument>([&] { Common bad(As, b3, cs, {CT::LessOrEqual, CT::GreaterOrEqual}, {VT::NonNegative, VT::NonNegative}, true); }));
        CHECK(throws<std::invalid_argument>([&] { Common bad(As, bs, cs, {CT::LessOrEqual}, {VT::NonNegative, VT::NonNegative}, true); }));
        CHECK(throws<std::invalid_argument>([&] { Common bad(As, bs, cs, {CT::LessOrEqual, CT::Equal}, {VT::Free}, true); }));
        Common com2(com);
        CHECK(com2.IsMaximization() && com2.GetConstraintsMatrix().rows() == 2);
        auto dcom = com.GetDual();                       // max -> min; <= row -> y >= 0, >= row -> y <= 0; x >= 0 -> >= rows
        CHECK(dcom && !dcom->IsMaximization() && dcom->GetConstraintsMatrix().rows() == 2 && dcom->GetConstraintsMatrix().cols() == 2);
        CHECK(dcom->GetConstraintsMatrix()(0, 1) == 3.0 && dcom->GetRightHandSide()[0] == 7.0 && dcom->GetObjectiveCoefficients()[1] == 6.0);
        CHECK(dcom->GetVariableTypes()[0] == VT::NonNegative && dcom->GetVariableTypes()[1] == VT::NonPositive);
        CHECK(dcom->GetConstraintTypes()[0] == CT::GreaterOrEqual && dcom->GetConstraintTypes()[1] == CT::GreaterOrEqual);
        auto ddcom = dcom->GetDual();                    // the dual of the dual is the primal, types included
        CHECK(ddcom->IsMaximization() && ddcom->GetConstraintsMatrix() == As && ddcom->GetRightHandSide() == bs);
        CHECK(ddcom->GetConstraintTypes() == com.GetConstraintTypes() && ddcom->GetVariableTypes() == com.GetVariableTypes());
        // symmetric form of the fixture: the >= row is negated
        auto s1 = com.ToSymmetrical();
        CHECK(s1->IsMaximization() && s1->GetConstraintsMatrix().rows() == 2 && s1->GetConstraintsMatrix()(1, 0) == -3.0 && s1->GetRightHandSide()[1] == -6.0);
    }
    {   // every row and variable kind at once, minimisation (Common.cpp:169-348)
        Common gen(mat(3, 3, {1, 2, 3, 4, 5, 6, 7, 8, 9}), vec({10, 11, 12}), vec({1, -2, 3}),
                   {CT::LessOrEqual, CT::GreaterOrEqual, CT::Equal}, {VT::Free, VT::NonNegative, VT::NonPositive}, false);
        auto s2 = gen.ToSymmetrical();
        CHECK(s2->IsMaximization());                     // always max / <=; min costs are negated
        const Eigen::MatrixXd want = mat(4, 4, { 1, -1,  2, -3,
                                                -4,  4, -5,  6,
                                                 7, -7,  8, -9,
                                                -7,  7, -8,  9});
        CHECK(s2->GetConstraintsMatrix() == want);
        CHECK(s2->GetRightHandSide() == vec({10, -11, 12, -12}));
        CHECK(s2->GetObjectiveCoefficients() == vec({-1, 1, 2, 3}));
        auto c2 = gen.ToCanonical();                     // = ToSymmetrical()->ToCanonical(): [A | I], slack basis, 4 original variables
        CHECK(c2->GetConstraintsMatrix().rows() == 4 && c2->GetConstraintsMatrix().cols() == 8 && c2->IsMaximization());
        CHECK(c2->GetOriginalVariablesCount() == 4 && c2->GetBasisIndices()[3] == 7 && c2->GetConstraintsMatrix()(3, 7) == 1.0);
        auto dg = gen.GetDual();                         // min -> max: <= row -> y <= 0, >= row -> y >= 0, = row -> free
        CHECK(dg->IsMaximization() && dg->GetVariableTypes() == (std::vector<VT>{VT::NonPositive, VT::NonNegative, VT::Free}));
        CHECK(dg->GetConstraintTypes() == (std::vector<CT>{CT::Equal, CT::LessOrEqual, CT::GreaterOrEqual}));
    }
    {   // tests/test_transformations.cpp:6-40 Common -> Symmetrical -> Canonical, :42-61 dual of the dual
        Common com(As, bs, cs, {CT::LessOrEqual, CT::LessOrEqual}, {VT::NonNegative, VT::NonNegative}, true);
        auto sym = com.ToSymmetrical();
        CHECK(sym && sym->GetConstraintsMatrix() == As && sym->GetRightHandSide() == bs && sym->GetObjectiveCoefficients() == cs);
        auto can2 = sym->ToCanonical();
        CHECK(can2 && can2->GetConstraintsMatrix().rows() == 2);
        auto dd = smax.GetDual()->GetDual();
        CHECK(dd->IsMaximization() && dd->GetConstraintsMatrix() == As && dd->GetRightHandSide() == bs && dd->GetObjectiveCoefficients() == cs);
    }
    {   // tests/test_symmetrical.cpp:87-97 ToCommon; tests/test_canonical.cpp:78-89 ToCommon; Canonical.cpp:230-303 ToSymmetrical
        auto com = smax.ToCommon();
        CHECK(com && com->IsMaximization() && com->GetConstraintTypes().size() == 2 && com->GetVariableTypes().size() == 2);
        CHECK(com->GetConstraintTypes()[0] == CT::LessOrEqual && com->GetVariableTypes()[1] == VT::NonNegative);
        CHECK(smin.ToCommon()->GetConstraintTypes()[1] == CT::GreaterOrEqual && !smin.ToCommon()->IsMaximization());
        Canonical can2(can);
        can2.SetOriginalVariablesCount(2);
        auto cc = can2.ToCommon();
        CHECK(cc && cc->GetObjectiveCoefficients().size() == 2 && cc->GetConstraintsMatrix().cols() == 2 && !cc->IsMaximization());
        CHECK(cc->GetConstraintTypes()[0] == CT::Equal && cc->GetConstraintsMatrix()(1, 1) == 4.0 && cc->GetRightHandSide()[1] == 6.0);
        auto cs2 = can2.ToSymmetrical();
        CHECK(!cs2->IsMaximization() && cs2->GetConstraintsMatrix() == mat(4, 2, {1, 2, -1, -2, 3, 4, -3, -4}));
        CHECK(cs2->GetRightHandSide() == vec({5, -5, 6, -6}) && cs2->GetObjectiveCoefficients() == cs);
    }

    // --- parser (tests/test_parser.cpp)
    SymmetricalParser parser;
    auto p1 = parser.ParseFromString("\n  maximize\n\n objective:\n 3 5\n\n constraints:\n 1 2 10\n 3 4 20\n");
    CHECK(p1 && p1->IsMaximization() && p1->GetConstraintsMatrix().rows() == 2 && p1->GetConstraintsMatrix().cols() == 2);
    CHECK(p1->GetRightHandSide()[1] == 20.0 && p1->GetConstraintsMatrix()(1, 0) == 3.0);
    auto p2 = parser.ParseFromString("minimize\nobjective:\n7 8\nsubject to:\n1 1 5\n2 3 12\n");
    CHECK(p2 && !p2->IsMaximization());
    auto p3 = parser.ParseFromString("# c\nmax\n# f\nobjective:\n1 2 3  # more\nconstraints:\n1 0 0 5 # a\r\n0 1 0 6\r\n0 0 1 7\r\n");
    CHECK(p3 && p3->GetObjectiveCoefficients().size() == 3 && p3->GetConstraintsMatrix().rows() == 3);
    auto p4 = parser.ParseFromString("maximize\n# nothing else\n");
    CHECK(!p4 && !parser.GetLastError().empty());
    CHECK(!parser.ParseFromString("1 2 3\n"));                             // data outside a section
    CHECK(!parser.ParseFromString("max\nobjective:\n1 2\nconstraints:\n1 2 3 4\n"));   // width mismatch
    CHECK(!parser.ParseFromFile("/nonexistent/file.txt"));

    // --- EnumerationSolver: argument checks that need no device
    CHECK(throws<std::invalid_argument>([&] {
        Canonical tall(mat(3, 2, {1, 0, 0, 1, 1, 1}), Eigen::VectorXd(3), Eigen::VectorXd(2), {0, 1, 1}, true);
        EnumerationSolver s(tall);
    }));
    EnumerationSolver s(can);
    CHECK(throws<std::logic_error>([&] { s.objective(); }));
    std::puts("host_tests: all passed");
    return 0;
}
