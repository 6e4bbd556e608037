// Common.cpp — general-form container and its conversions (behaviour of
// reference src/ProblemTypes/Common.cpp:16-45 constructor checks, :169-362
// ToSymmetrical / ToCanonical, :366-448 GetDual).
#include "ProblemTypes/Common.h"

#include <iostream>
#include <stdexcept>

#include "PrintLP.h"
#include "ProblemTypes/Canonical.h"
#include "ProblemTypes/Symmetrical.h"

Common::Common(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c,
               const std::vector<ConstraintType>& constraintTypes, const std::vector<VariableType>& variableTypes,
               bool maximize)
    : A_(A), b_(b), c_(c), rowTypes_(constraintTypes), varTypes_(variableTypes), maximize_(maximize)
{
    if (A_.rows() != b_.size()) throw std::invalid_argument("Common: rows of A and size of b differ");
    if (A_.cols() != c_.size()) throw std::invalid_argument("Common: columns of A and size of c differ");
    if (static_cast<Eigen::Index>(rowTypes_.size()) != A_.rows())
        throw std::invalid_argument("Common: number of constraint types differs from the number of rows of A");
    if (static_cast<Eigen::Index>(varTypes_.size()) != A_.cols())
        throw std::invalid_argument("Common: number of variable types differs from the number of columns of A");
}

double Common::Evaluate(const Eigen::VectorXd& solution) const
{
    if (solution.size() != c_.size()) throw std::invalid_argument("Common::Evaluate: solution size differs from the number of variables");
    return c_.dot(solution);
}

void Common::Print() const
{
    static const char* const rel[] = {"<=", ">=", "="};                    // no blanks, and no '*' in the rows: as the reference prints
    static const char* const dom[] = {"∈R", " >= 0", " <= 0"};
    std::ostream& os = std::cout;
    lp_print::objective(os, "=== Общая форма задачи ЛП ===", maximize_, c_);
    lp_print::rows(os, "При ограничениях:", A_, b_, "", [this](Eigen::Index i) { return rel[static_cast<int>(rowTypes_[static_cast<size_t>(i)])]; });
    os << "\nОграничения на переменные:\n";
    for (size_t j = 0; j < varTypes_.size(); ++j) os << "x" << (j + 1) << ": " << dom[static_cast<int>(varTypes_[j])] << "\n";
}

std::unique_ptr<Symmetrical> Common::ToSymmetrical() const
{
    const Eigen::Index m = A_.rows(), n = A_.cols();
    // Row i of the input lands at out_row[i] (and out_row[i]+1 for an equality)
    // with sign row_sign[i]; column j lands at out_col[j] (and +1 for a free
    // variable) with sign col_sign[j].  Every output entry is then one product
    // of signs times A(i,j) — negation is exact, so the order does not matter.
    std::vector<Eigen::Index> out_row(static_cast<size_t>(m)), out_col(static_cast<size_t>(n));
    Eigen::Index rows = 0, cols = 0;
    for (Eigen::Index i = 0; i < m; ++i) { out_row[static_cast<size_t>(i)] = rows; rows += rowTypes_[static_cast<size_t>(i)] == ConstraintType::Equal ? 2 : 1; }
    for (Eigen::Index j = 0; j < n; ++j) { out_col[static_cast<size_t>(j)] = cols; cols += varTypes_[static_cast<size_t>(j)] == VariableType::Free ? 2 : 1; }

    Eigen::MatrixXd As(rows, cols);
    Eigen::VectorXd bs(rows), cs(cols);
    const double sense = maximize_ ? 1.0 : -1.0;            // min c'x  ==  -max (-c)'x
    for (Eigen::Index j = 0; j < n; ++j) {
        const VariableType vt = varTypes_[static_cast<size_t>(j)];
        const double col_sign = vt == VariableType::NonPositive ? -1.0 : 1.0;
        const Eigen::Index cj = out_col[static_cast<size_t>(j)];
        cs[cj] = sense * col_sign * c_[j];
        if (vt == VariableType::Free) cs[cj + 1] = -cs[cj];
        for (Eigen::Index i = 0; i < m; ++i) {
            const ConstraintType rt = rowTypes_[static_cast<size_t>(i)];
            const double row_sign = rt == ConstraintType::GreaterOrEqual ? -1.0 : 1.0;
            const Eigen::Index ri = out_row[static_cast<size_t>(i)];
            const double v = row_sign * col_sign * A_(i, j);
            As(ri, cj) = v;
            if (vt == VariableType::Free) As(ri, cj + 1) = -v;
            if (rt == ConstraintType::Equal) {
                As(ri + 1, cj) = -v;
                if (vt == VariableType::Free) As(ri + 1, cj + 1) = v;
            }
        }
    }
    for (Eigen::Index i = 0; i < m; ++i) {
        const ConstraintType rt = rowTypes_[static_cast<size_t>(i)];
        const Eigen::Index ri = out_row[static_cast<size_t>(i)];
        bs[ri] = rt == ConstraintType::GreaterOrEqual ? -b_[i] : b_[i];
        if (rt == ConstraintType::Equal) bs[ri + 1] = -b_[i];
    }
    return std::make_unique<Symmetrical>(As, bs, cs, /*maximize=*/true);
}

std::unique_ptr<Canonical> Common::ToCanonical() const
{
    return ToSymmetrical()->ToCanonical();
}

std::unique_ptr<Common> Common::GetDual() const
{
    //   primal max:  row <=  ->  y >= 0     row >=  ->  y <= 0     row =  ->  y free
    //                x >= 0  ->  row >=     x <= 0  ->  row <=     x free ->  row =
    //   primal min:  the inequalities of both columns flip.
    std::vector<VariableType> dualVars(rowTypes_.size());
    for (size_t i = 0; i < rowTypes_.size(); ++i) {
        if (rowTypes_[i] == ConstraintType::Equal) dualVars[i] = VariableType::Free;
        else dualVars[i] = ((rowTypes_[i] == ConstraintType::LessOrEqual) == maximize_) ? VariableType::NonNegative : VariableType::NonPositive;
    }
    std::vector<ConstraintType> dualRows(varTypes_.size());
    for (size_t j = 0; j < varTypes_.size(); ++j) {
        if (varTypes_[j] == VariableType::Free) dualRows[j] = ConstraintType::Equal;
        else dualRows[j] = ((varTypes_[j] == VariableType::NonNegative) == maximize_) ? ConstraintType::GreaterOrEqual : ConstraintType::LessOrEqual;
    }
    return std::make_unique<Common>(A_.transpose(), c_, b_, dualRows, dualVars, !maximize_);
}
