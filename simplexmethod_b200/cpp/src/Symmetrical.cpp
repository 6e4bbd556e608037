// Symmetrical.cpp — symmetric-form container and its conversion to canonical
// form (behaviour of reference src/ProblemTypes/Symmetrical.cpp:119-273).
#include "ProblemTypes/Symmetrical.h"

#include <iostream>
#include <stdexcept>
#include <vector>

#include "PrintLP.h"
#include "ProblemTypes/Canonical.h"
#include "ProblemTypes/Common.h"

Symmetrical::Symmetrical(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c, bool maximize)
    : A_(A), b_(b), c_(c), maximize_(maximize)
{
    if (A_.rows() != b_.size()) throw std::invalid_argument("Symmetrical: rows of A and size of b differ");
    if (A_.cols() != c_.size()) throw std::invalid_argument("Symmetrical: columns of A and size of c differ");
}

double Symmetrical::Evaluate(const Eigen::VectorXd& solution) const
{
    if (solution.size() != c_.size()) throw std::invalid_argument("Symmetrical::Evaluate: solution size differs from the number of variables");
    return c_.dot(solution);
}

void Symmetrical::Print() const
{
    std::ostream& os = std::cout;
    lp_print::objective(os, "=== Симметричная форма задачи ЛП ===", maximize_, c_);
    const char* rel = maximize_ ? " <= " : " >= ";
    lp_print::rows(os, "При ограничениях:", A_, b_, "*", [rel](Eigen::Index) { return rel; });
    os << "\nВсе переменные неотрицательны: x_i >= 0\n";
}

std::unique_ptr<Symmetrical> Symmetrical::GetDual() const
{
    return std::make_unique<Symmetrical>(A_.transpose(), c_, b_, !maximize_);
}

std::unique_ptr<Canonical> Symmetrical::ToCanonical() const
{
    const auto m = A_.rows(), n = A_.cols();
    // max: one slack per row, Ax + s = b, slacks form the basis.
    // min: Ax - s + a = b with surplus s and artificial a; artificials form the
    // basis and, as in the reference, carry zero cost.
    const auto extra = maximize_ ? m : 2 * m;
    Eigen::MatrixXd Ac = Eigen::MatrixXd::Zero(m, n + extra);      // real Eigen does not zero MatrixXd(r, c)
    Eigen::VectorXd cc = Eigen::VectorXd::Zero(n + extra);
    for (Eigen::Index j = 0; j < n; ++j) {
        cc[j] = c_[j];
        for (Eigen::Index i = 0; i < m; ++i) Ac(i, j) = A_(i, j);
    }
    std::vector<int> basis(static_cast<size_t>(m));
    for (Eigen::Index i = 0; i < m; ++i) {
        if (maximize_) {
            Ac(i, n + i) = 1.0;
            basis[static_cast<size_t>(i)] = static_cast<int>(n + i);
        } else {
            Ac(i, n + i) = -1.0;
            Ac(i, n + m + i) = 1.0;
            basis[static_cast<size_t>(i)] = static_cast<int>(n + m + i);
        }
    }
    auto out = std::make_unique<Canonical>(Ac, b_, cc, basis, /*minimize=*/!maximize_);
    out->SetOriginalVariablesCount(static_cast<int>(n));
    return out;
}

std::unique_ptr<Common> Symmetrical::ToCommon() const
{
    const auto rowType = maximize_ ? Common::ConstraintType::LessOrEqual : Common::ConstraintType::GreaterOrEqual;
    return std::make_unique<Common>(A_, b_, c_, std::vector<Common::ConstraintType>(static_cast<size_t>(A_.rows()), rowType),
                                    std::vector<Common::VariableType>(static_cast<size_t>(A_.cols()), Common::VariableType::NonNegative),
                                    maximize_);
}
