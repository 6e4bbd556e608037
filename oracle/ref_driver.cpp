// ref_driver.cpp — C-ABI doorway to the REFERENCE'S OWN CODE.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile (target "ref") compiles this file together with the reference's sources,
// unmodified and from where they lie — /root/reference/src/ProblemTypes/{Canonical,Symmetrical,
// Common}.cpp, src/SymmetricalParser.cpp and the header-only src/SimplexSolover.h — against the
// Eigen API stand-in oracle/eigen_shim (NOT Eigen; see its header) into oracle/_ref/libsimplexref.so.
// Nothing under /root/reference is copied into the repository.
//
// Entry points (plain pointers and sizes; matrices column-major, lda = rows):
//   ref_convert          Common / Symmetrical / Canonical -> ToSymmetrical / ToCanonical / ToCommon / GetDual
//   ref_parse            SymmetricalParser::ParseFromString
//   ref_print            Print() of any of the three forms, captured from std::cout
//   ref_basic_solution   Canonical::GetBasicSolution + IsFeasibleBasis + Evaluate for one designated basis
//   ref_enumerate        the enumeration path AS THE REFERENCE WOULD RUN IT: the class
//                        EnumerationSolver is an empty stub (src/EnumerationSolver.h:3-10), so the loop over
//                        the C(n,m) sorted column tuples (lexicographic, strict improvement => lowest rank
//                        wins ties) is written here, and every per-basis step inside it is a call into the
//                        reference: singular bases rejected as Solver::computeBFS does (FullPivLU::isInvertible,
//                        SimplexSolover.h:124-126), Canonical(A,b,c,S).IsFeasibleBasis() (Canonical.cpp:165-177),
//                        GetBasicSolution() (:179-197), Evaluate() (:79-87), IsMaximization() (:141-144)
//   ref_simplex_solve    Solver(canonical).solve() (SimplexSolover.h:285-328)
#include <cstdint>
#include <cstring>
#include <exception>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "ProblemTypes/Canonical.h"
#include "ProblemTypes/Common.h"
#include "ProblemTypes/Symmetrical.h"
#include "SimplexSolover.h"
#include "SymmetricalParser.h"

namespace {
thread_local std::string g_err;

Eigen::MatrixXd mat(int m, int n, const double* A)
{
    Eigen::MatrixXd M(m, n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) M(i, j) = A[i + (size_t)j * m];
    return M;
}
Eigen::VectorXd vec(int n, const double* v)
{
    Eigen::VectorXd x(n);
    for (int i = 0; i < n; ++i) x(i) = v[i];
    return x;
}
}  // namespace

extern "C" {

enum { REF_COMMON = 0, REF_SYMMETRICAL = 1, REF_CANONICAL = 2 };
enum { REF_TO_SYMMETRICAL = 0, REF_TO_CANONICAL = 1, REF_TO_COMMON = 2, REF_GET_DUAL = 3 };

// One LP in any of the three forms.  Inputs: caller-owned arrays.  Outputs: caller-allocated arrays of
// capacity cap_m x cap_n (A), cap_m (b, row_types, basis), cap_n (c, var_types); dimensions written back.
struct ref_problem {
    int32_t kind;             // REF_*
    int32_t m, n;
    int32_t maximize;         // Canonical: !minimize
    int32_t n_orig;           // Canonical only
    double* A;                // column-major, lda = m
    double* b;
    double* c;
    int32_t* row_types;       // Common: Common::ConstraintType as int (LessOrEqual, GreaterOrEqual, Equal)
    int32_t* var_types;       // Common: Common::VariableType as int (Free, NonNegative, NonPositive)
    int32_t* basis;           // Canonical: m indices
    int32_t cap_m, cap_n;     // capacities (outputs only)
};

const char* ref_last_error(void) { return g_err.c_str(); }

static int store(const IProblem& p, int kind, ref_problem* out)
{
    const Eigen::MatrixXd& A = p.GetConstraintsMatrix();
    const int m = (int)A.rows(), n = (int)A.cols();
    if (m > out->cap_m || n > out->cap_n) { g_err = "output capacity too small"; return -2; }
    out->kind = kind; out->m = m; out->n = n; out->maximize = p.IsMaximization() ? 1 : 0; out->n_orig = n;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) out->A[i + (size_t)j * m] = A(i, j);
    for (int i = 0; i < m; ++i) out->b[i] = p.GetRightHandSide()(i);
    for (int j = 0; j < n; ++j) out->c[j] = p.GetObjectiveCoefficients()(j);
    if (kind == REF_COMMON) {
        const Common& c = static_cast<const Common&>(p);
        for (int i = 0; i < m; ++i) out->row_types[i] = (int)c.GetConstraintTypes()[(size_t)i];
        for (int j = 0; j < n; ++j) out->var_types[j] = (int)c.GetVariableTypes()[(size_t)j];
    }
    if (kind == REF_CANONICAL) {
        const Canonical& c = static_cast<const Canonical&>(p);
        out->n_orig = c.GetOriginalVariablesCount();
        for (int i = 0; i < m; ++i) out->basis[i] = c.GetBasisIndices()[(size_t)i];
    }
    return 0;
}

int ref_convert(const ref_problem* in, int32_t op, ref_problem* out)
{
    g_err.clear();
    try {
        const Eigen::MatrixXd A = mat(in->m, in->n, in->A);
        const Eigen::VectorXd b = vec(in->m, in->b), c = vec(in->n, in->c);
        if (in->kind == REF_COMMON) {
            std::vector<Common::ConstraintType> rt;
            std::vector<Common::VariableType> vt;
            for (int i = 0; i < in->m; ++i) rt.push_back((Common::ConstraintType)in->row_types[i]);
            for (int j = 0; j < in->n; ++j) vt.push_back((Common::VariableType)in->var_types[j]);
            Common p(A, b, c, rt, vt, in->maximize != 0);
            switch (op) {
                case REF_TO_SYMMETRICAL: return store(*p.ToSymmetrical(), REF_SYMMETRICAL, out);
                case REF_TO_CANONICAL:   return store(*p.ToCanonical(), REF_CANONICAL, out);
                case REF_GET_DUAL:       return store(*p.GetDual(), REF_COMMON, out);
            }
        } else if (in->kind == REF_SYMMETRICAL) {
            Symmetrical p(A, b, c, in->maximize != 0);
            switch (op) {
                case REF_TO_CANONICAL: return store(*p.ToCanonical(), REF_CANONICAL, out);
                case REF_TO_COMMON:    return store(*p.ToCommon(), REF_COMMON, out);
                case REF_GET_DUAL:     return store(*p.GetDual(), REF_SYMMETRICAL, out);
            }
        } else if (in->kind == REF_CANONICAL) {
            Canonical p(A, b, c, std::vector<int>(in->basis, in->basis + in->m), in->maximize == 0);
            p.SetOriginalVariablesCount(in->n_orig);
            switch (op) {
                case REF_TO_SYMMETRICAL: return store(*p.ToSymmetrical(), REF_SYMMETRICAL, out);
                case REF_TO_COMMON:      return store(*p.ToCommon(), REF_COMMON, out);
                case REF_GET_DUAL:       return store(*p.GetDual(), REF_CANONICAL, out);
            }
        }
        g_err = "unsupported (kind, op)";
        return -1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

// the text the reference's Print() writes to std::cout for this problem; returns its length (or -3 / -2: too long)
int ref_print(const ref_problem* in, char* buf, int32_t cap)
{
    g_err.clear();
    std::ostringstream text;
    std::streambuf* old = std::cout.rdbuf(text.rdbuf());
    int rc = 0;
    try {
        const Eigen::MatrixXd A = mat(in->m, in->n, in->A);
        const Eigen::VectorXd b = vec(in->m, in->b), c = vec(in->n, in->c);
        if (in->kind == REF_COMMON) {
            std::vector<Common::ConstraintType> rt;
            std::vector<Common::VariableType> vt;
            for (int i = 0; i < in->m; ++i) rt.push_back((Common::ConstraintType)in->row_types[i]);
            for (int j = 0; j < in->n; ++j) vt.push_back((Common::VariableType)in->var_types[j]);
            Common(A, b, c, rt, vt, in->maximize != 0).Print();
        } else if (in->kind == REF_SYMMETRICAL) {
            Symmetrical(A, b, c, in->maximize != 0).Print();
        } else {
            Canonical p(A, b, c, std::vector<int>(in->basis, in->basis + in->m), in->maximize == 0);
            p.SetOriginalVariablesCount(in->n_orig);
            p.Print();
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        rc = -3;
    }
    std::cout.rdbuf(old);
    if (rc) return rc;
    const std::string t = text.str();
    if ((int)t.size() + 1 > cap) { g_err = "output capacity too small"; return -2; }
    std::memcpy(buf, t.c_str(), t.size() + 1);
    return (int)t.size();
}

// returns 0 and fills *out (kind = SYMMETRICAL), or 1 = the parser returned nullptr (message in ref_last_error)
int ref_parse(const char* text, ref_problem* out)
{
    g_err.clear();
    try {
        SymmetricalParser parser;
        auto p = parser.ParseFromString(text);
        if (!p) { g_err = parser.GetLastError(); return 1; }
        return store(*p, REF_SYMMETRICAL, out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

// x: n values; returns 0, or -3 with the exception text
int ref_basic_solution(int32_t m, int32_t n, const double* A, const double* b, const double* c, const int32_t* basis,
                       int32_t minimize, double* x, int32_t* feasible, double* z)
{
    g_err.clear();
    try {
        Canonical p(mat(m, n, A), vec(m, b), vec(n, c), std::vector<int>(basis, basis + m), minimize != 0);
        const Eigen::VectorXd sol = p.GetBasicSolution();
        for (int j = 0; j < n; ++j) x[j] = sol(j);
        *feasible = p.IsFeasibleBasis() ? 1 : 0;
        *z = p.Evaluate(sol);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

struct ref_enum_result {
    int32_t  status;          // 0 optimum found, 1 no feasible basis
    int32_t  m;
    int32_t  basis[16];
    double   x_B[16];         // basic values in basis order
    double   objective;
    uint64_t best_rank, n_bases, n_singular, n_infeasible, n_feasible;
};

// ranks [rank_begin, rank_end) of the lexicographic order (0, 0 = all); status_out (may be NULL): one byte per
// rank of the range, 0 feasible / 1 infeasible / 2 singular
int ref_enumerate_range(int32_t m, int32_t n, const double* A, const double* b, const double* c, int32_t maximize,
                        uint64_t rank_begin, uint64_t rank_end, uint8_t* status_out, ref_enum_result* out)
{
    g_err.clear();
    if (m < 1 || m > 16 || n < m || n > 64) { g_err = "bad dimensions"; return -1; }
    try {
        const Eigen::MatrixXd Am = mat(m, n, A);
        const Eigen::VectorXd bv = vec(m, b), cv = vec(n, c);
        std::memset(out, 0, sizeof *out);
        out->m = m; out->status = 1; out->best_rank = UINT64_MAX;
        // Pascal's triangle for the starting tuple (combinatorial number system)
        std::vector<std::vector<uint64_t>> C((size_t)n + 1, std::vector<uint64_t>((size_t)m + 1, 0));
        for (int t = 0; t <= n; ++t) { C[(size_t)t][0] = 1; for (int k = 1; k <= m && k <= t; ++k) C[(size_t)t][(size_t)k] = C[(size_t)t - 1][(size_t)k - 1] + C[(size_t)t - 1][(size_t)k]; }
        const uint64_t total = C[(size_t)n][(size_t)m];
        if (rank_begin == 0 && rank_end == 0) rank_end = total;
        if (rank_begin > rank_end || rank_end > total) { g_err = "bad rank range"; return -1; }
        if (rank_begin == rank_end) return 0;
        std::vector<int> S((size_t)m);
        {
            uint64_t r = rank_begin;
            int v = 0;
            for (int i = 0; i < m; ++i) {
                for (;;) { const uint64_t cnt = C[(size_t)(n - 1 - v)][(size_t)(m - 1 - i)]; if (cnt <= r) { r -= cnt; ++v; } else break; }
                S[(size_t)i] = v++;
            }
        }
        bool have = false;
        double best = 0.0;
        for (uint64_t rank = rank_begin; rank < rank_end; ++rank) {
            ++out->n_bases;
            int cls;
            // singular bases are rejected the way the reference's solver does (SimplexSolover.h:110-126)
            Eigen::MatrixXd B(m, m);
            for (int t = 0; t < m; ++t) B.col(t) = Am.col(S[(size_t)t]);
            Eigen::FullPivLU<Eigen::MatrixXd> lu(B);
            if (!lu.isInvertible()) { cls = 2; ++out->n_singular; }
            else {
                Canonical p(Am, bv, cv, S, maximize == 0);
                if (!p.IsFeasibleBasis()) { cls = 1; ++out->n_infeasible; }
                else {
                    cls = 0; ++out->n_feasible;
                    const Eigen::VectorXd x = p.GetBasicSolution();
                    const double z = p.Evaluate(x);
                    const bool improves = !have || (p.IsMaximization() ? z > best : z < best);
                    if (improves) {
                        have = true; best = z;
                        out->status = 0; out->objective = z; out->best_rank = rank;
                        for (int t = 0; t < m; ++t) { out->basis[t] = S[(size_t)t]; out->x_B[t] = x(S[(size_t)t]); }
                    }
                }
            }
            if (status_out) status_out[rank - rank_begin] = (uint8_t)cls;
            int i = m - 1;                                   // next sorted tuple, lexicographic
            while (i >= 0 && S[(size_t)i] == n - m + i) --i;
            if (i < 0) break;
            ++S[(size_t)i];
            for (int j = i + 1; j < m; ++j) S[(size_t)j] = S[(size_t)j - 1] + 1;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

int ref_enumerate(int32_t m, int32_t n, const double* A, const double* b, const double* c, int32_t maximize,
                  uint8_t* status_out, ref_enum_result* out)
{
    return ref_enumerate_range(m, n, A, b, c, maximize, 0, 0, status_out, out);
}

// x: n_orig values.  returns 0, or -3 with the solver's exception text ("Целевая функция неограничена", ...)
int ref_simplex_solve(int32_t m, int32_t n, const double* A, const double* b, const double* c, const int32_t* basis,
                      int32_t minimize, int32_t n_orig, double* x)
{
    g_err.clear();
    try {
        Canonical p(mat(m, n, A), vec(m, b), vec(n, c), std::vector<int>(basis, basis + m), minimize != 0);
        p.SetOriginalVariablesCount(n_orig);
        Solver s(p);
        const Eigen::VectorXd sol = s.solve();
        for (int j = 0; j < n_orig; ++j) x[j] = sol(j);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -3;
    }
}

}  // extern "C"
