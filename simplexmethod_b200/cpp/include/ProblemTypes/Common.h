// Common.h — general form of an LP: each row is <=, >= or =, each variable is
// free, >= 0 or <= 0.  Same public surface as the reference
// (reference: src/ProblemTypes/Common.h:10-55).  It reaches the enumeration
// path through ToSymmetrical()/ToCanonical() (reference Common.cpp:169-362).
#pragma once

#include <memory>
#include <vector>

#include "IProblem.h"

class Symmetrical;
class Canonical;

class Common : public IProblem {
public:
    enum class ConstraintType { LessOrEqual, GreaterOrEqual, Equal };
    enum class VariableType { Free, NonNegative, NonPositive };

    Common(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c,
           const std::vector<ConstraintType>& constraintTypes, const std::vector<VariableType>& variableTypes,
           bool maximize);

    double Evaluate(const Eigen::VectorXd& solution) const override;
    void Print() const override;
    const Eigen::MatrixXd& GetConstraintsMatrix() const override { return A_; }
    const Eigen::VectorXd& GetRightHandSide() const override { return b_; }
    const Eigen::VectorXd& GetObjectiveCoefficients() const override { return c_; }
    bool IsMaximization() const override { return maximize_; }

    const std::vector<ConstraintType>& GetConstraintTypes() const { return rowTypes_; }
    const std::vector<VariableType>& GetVariableTypes() const { return varTypes_; }

    // Always "max c'x, Ax <= b, x >= 0": a free variable becomes a pair (x', x''),
    // a non-positive one is negated, a >= row is negated, an = row becomes the
    // pair (row, -row); a minimisation has its costs negated.
    std::unique_ptr<Symmetrical> ToSymmetrical() const;
    std::unique_ptr<Canonical> ToCanonical() const;     // via ToSymmetrical()
    // Dual in general form: transpose, swap b and c, flip the sense; row types
    // become variable types and vice versa (table in Common.cpp).
    std::unique_ptr<Common> GetDual() const;

private:
    Eigen::MatrixXd A_;
    Eigen::VectorXd b_, c_;
    std::vector<ConstraintType> rowTypes_;
    std::vector<VariableType> varTypes_;
    bool maximize_;
};
