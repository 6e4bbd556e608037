// IProblem.h — abstract interface of the LP problem types, same virtuals as the
// reference (reference: src/ProblemTypes/IProblem.h:7-17).
#pragma once

#include "../DenseShim.h"

class IProblem {
public:
    virtual ~IProblem() = default;
    virtual double Evaluate(const Eigen::VectorXd& solution) const = 0;
    virtual void Print() const = 0;
    virtual const Eigen::MatrixXd& GetConstraintsMatrix() const = 0;
    virtual const Eigen::VectorXd& GetRightHandSide() const = 0;
    virtual const Eigen::VectorXd& GetObjectiveCoefficients() const = 0;
    virtual bool IsMaximization() const = 0;
};
