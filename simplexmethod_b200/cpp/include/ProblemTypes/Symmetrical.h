// Symmetrical.h — max c'x, Ax <= b, x >= 0  /  min c'x, Ax >= b, x >= 0.
// Feeds the enumeration path through ToCanonical (reference:
// src/ProblemTypes/Symmetrical.h:17-45, Symmetrical.cpp:119-273).
#pragma once

#include <memory>

#include "IProblem.h"

class Canonical;
class Common;

class Symmetrical : public IProblem {
public:
    Symmetrical(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c, bool maximize);

    double Evaluate(const Eigen::VectorXd& solution) const override;
    void Print() const override;
    const Eigen::MatrixXd& GetConstraintsMatrix() const override { return A_; }
    const Eigen::VectorXd& GetRightHandSide() const override { return b_; }
    const Eigen::VectorXd& GetObjectiveCoefficients() const override { return c_; }
    bool IsMaximization() const override { return maximize_; }

    std::unique_ptr<Symmetrical> GetDual() const;     // transpose A, swap b and c, flip the sense
    std::unique_ptr<Canonical> ToCanonical() const;   // max: [A | I], slack basis; min: [A | -I | I], artificial basis
    std::unique_ptr<Common> ToCommon() const;         // same data; rows all <= (max) or >= (min), variables >= 0

private:
    Eigen::MatrixXd A_;
    Eigen::VectorXd b_, c_;
    bool maximize_;
};
