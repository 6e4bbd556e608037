// Canonical.cpp — host container for the canonical form; validation order and
// messages' meaning follow the reference constructor (reference:
// src/ProblemTypes/Canonical.cpp:27-46); numerical per-basis methods call
// libenumgpu.
#include "ProblemTypes/Canonical.h"

#include <iostream>
#include <stdexcept>
#include <string>

#include "PrintLP.h"
#include "ProblemTypes/Common.h"
#include "ProblemTypes/Symmetrical.h"
#include "enumgpu.h"

Canonical::Canonical(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c,
                     const std::vector<int>& basisIndices, bool minimize)
    : A_(A), b_(b), c_(c), basis_(basisIndices), minimize_(minimize), n_orig_(static_cast<int>(c.size()))
{
    if (A_.rows() != b_.size()) throw std::invalid_argument("Canonical: rows of A and size of b differ");
    if (A_.cols() != c_.size()) throw std::invalid_argument("Canonical: columns of A and size of c differ");
    if (static_cast<long>(basis_.size()) != static_cast<long>(A_.rows()))
        throw std::invalid_argument("Canonical: number of basis indices differs from the number of rows of A");
    for (int idx : basis_)
        if (idx < 0 || idx >= A_.cols()) throw std::invalid_argument("Canonical: basis index out of range");
}

double Canonical::Evaluate(const Eigen::VectorXd& solution) const
{
    if (solution.size() != c_.size()) throw std::invalid_argument("Canonical::Evaluate: solution size differs from the number of variables");
    return c_.dot(solution);
}

void Canonical::Print() const
{
    std::ostream& os = std::cout;
    lp_print::objective(os, "=== Каноническая форма задачи ЛП ===", !minimize_, c_);
    lp_print::rows(os, "При ограничениях (Ax = b):", A_, b_, "*", [](Eigen::Index) { return " = "; });
    os << "\nВсе переменные неотрицательны: x_i >= 0\n\nБазисные переменные: ";
    for (size_t i = 0; i < basis_.size(); ++i) os << (i ? ", " : "") << "x" << (basis_[i] + 1);
    os << "\nКоличество исходных переменных: " << n_orig_ << "\nДополнительных переменных: " << (c_.size() - n_orig_) << "\n";
}

void Canonical::SetOriginalVariablesCount(int count)
{
    if (count <= 0 || count > c_.size()) throw std::invalid_argument("Canonical: invalid number of original variables");
    n_orig_ = count;
}

namespace {
int eval_designated_basis(const Canonical& p, std::vector<double>& xB, double& z)
{
    const Eigen::MatrixXd& A = p.GetConstraintsMatrix();
    enumgpu_problem ep{};
    ep.m = static_cast<int32_t>(A.rows());
    ep.n = static_cast<int32_t>(A.cols());
    ep.lda = ep.m;
    ep.maximize = p.IsMaximization() ? 1 : 0;
    ep.A_colmajor = A.data();
    ep.b = p.GetRightHandSide().data();
    ep.c = p.GetObjectiveCoefficients().data();
    std::vector<int32_t> basis(p.GetBasisIndices().begin(), p.GetBasisIndices().end());
    xB.assign(static_cast<size_t>(ep.m), 0.0);
    int32_t cls = 0;
    const int rc = enumgpu_eval_basis(&ep, nullptr, basis.data(), xB.data(), &z, &cls);
    if (rc == ENUMGPU_ERR_CUDA) throw std::runtime_error(std::string("libenumgpu: ") + enumgpu_last_error());
    if (rc != ENUMGPU_OK) throw std::invalid_argument(std::string("libenumgpu: ") + enumgpu_last_error());
    return cls;
}
}  // namespace

Eigen::VectorXd Canonical::GetBasicSolution() const
{
    std::vector<double> xB;
    double z = 0.0;
    if (eval_designated_basis(*this, xB, z) == ENUMGPU_BASIS_SINGULAR) throw std::runtime_error("Singular basis matrix");
    Eigen::VectorXd x = Eigen::VectorXd::Zero(c_.size());
    for (size_t i = 0; i < basis_.size(); ++i) x[basis_[i]] = xB[i];
    return x;
}

bool Canonical::IsFeasibleBasis() const
{
    std::vector<double> xB;
    double z = 0.0;
    return eval_designated_basis(*this, xB, z) == ENUMGPU_BASIS_FEASIBLE;
}

std::unique_ptr<Canonical> Canonical::GetDual() const
{
    const auto m = A_.rows(), n = A_.cols();
    Eigen::MatrixXd Ad = Eigen::MatrixXd::Zero(n, 2 * m + n);     // real Eigen does not zero MatrixXd(r, c)
    Eigen::VectorXd cd = Eigen::VectorXd::Zero(2 * m + n);
    for (Eigen::Index j = 0; j < n; ++j) {
        for (Eigen::Index i = 0; i < m; ++i) {
            Ad(j, i) = A_(i, j);            // y'
            Ad(j, m + i) = -A_(i, j);       // y''
        }
        Ad(j, 2 * m + j) = 1.0;             // slack of dual row j
    }
    for (Eigen::Index i = 0; i < m; ++i) { cd[i] = b_[i]; cd[m + i] = -b_[i]; }
    std::vector<int> basis(static_cast<size_t>(n));
    for (Eigen::Index j = 0; j < n; ++j) basis[static_cast<size_t>(j)] = static_cast<int>(2 * m + j);
    auto dual = std::make_unique<Canonical>(Ad, c_, cd, basis, !minimize_);
    dual->SetOriginalVariablesCount(static_cast<int>(2 * m));
    return dual;
}

std::unique_ptr<Common> Canonical::ToCommon() const
{
    const Eigen::Index m = A_.rows(), n = n_orig_;      // added (slack / surplus / artificial) columns are dropped
    Eigen::MatrixXd A(m, n);
    Eigen::VectorXd c(n);
    for (Eigen::Index j = 0; j < n; ++j) {
        c[j] = c_[j];
        for (Eigen::Index i = 0; i < m; ++i) A(i, j) = A_(i, j);
    }
    return std::make_unique<Common>(A, b_, c, std::vector<Common::ConstraintType>(static_cast<size_t>(m), Common::ConstraintType::Equal),
                                    std::vector<Common::VariableType>(static_cast<size_t>(n), Common::VariableType::NonNegative),
                                    /*maximize=*/!minimize_);
}

std::unique_ptr<Symmetrical> Canonical::ToSymmetrical() const
{
    const Eigen::Index m = A_.rows(), n = n_orig_;
    Eigen::MatrixXd A(2 * m, n);
    Eigen::VectorXd b(2 * m), c(n);
    for (Eigen::Index j = 0; j < n; ++j) c[j] = c_[j];
    for (Eigen::Index i = 0; i < m; ++i) {              // a'x = b  ->  a'x (<=|>=) b  and  -a'x (<=|>=) -b
        for (Eigen::Index j = 0; j < n; ++j) {
            A(2 * i, j) = A_(i, j);
            A(2 * i + 1, j) = -A_(i, j);
        }
        b[2 * i] = b_[i];
        b[2 * i + 1] = -b_[i];
    }
    return std::make_unique<Symmetrical>(A, b, c, /*maximize=*/!minimize_);
}
