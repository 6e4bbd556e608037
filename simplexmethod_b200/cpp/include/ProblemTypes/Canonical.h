// Canonical.h — min/max c'x, Ax = b, x >= 0 with a designated basis: the input
// type of EnumerationSolver.  Public surface follows the reference
// (reference: src/ProblemTypes/Canonical.h:10-49); the two per-basis numerical
// methods, GetBasicSolution and IsFeasibleBasis (reference Canonical.cpp:165-197,
// Eigen ColPivHouseholderQR on the host), are served by libenumgpu's
// enumgpu_eval_basis on the GPU with the frozen GE arithmetic.
#pragma once

#include <memory>
#include <vector>

#include "IProblem.h"

class Common;
class Symmetrical;

class Canonical : public IProblem {
public:
    Canonical(const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const Eigen::VectorXd& c,
              const std::vector<int>& basisIndices, bool minimize = true);

    double Evaluate(const Eigen::VectorXd& solution) const override;
    void Print() const override;
    const Eigen::MatrixXd& GetConstraintsMatrix() const override { return A_; }
    const Eigen::VectorXd& GetRightHandSide() const override { return b_; }
    const Eigen::VectorXd& GetObjectiveCoefficients() const override { return c_; }
    bool IsMaximization() const override { return !minimize_; }

    const std::vector<int>& GetBasisIndices() const { return basis_; }
    int GetOriginalVariablesCount() const { return n_orig_; }
    void SetOriginalVariablesCount(int count);

    // x (length n) of the designated basis; throws std::runtime_error if the
    // basis matrix is singular (the reference's solver-side convention,
    // SimplexSolover.h:124-126) or no CUDA device is available.
    Eigen::VectorXd GetBasicSolution() const;
    bool IsFeasibleBasis() const;

    // Canonical form of the dual as the reference builds it (Canonical.cpp:305-364):
    // [A' | -A' | I], right-hand side c, costs [b | -b | 0], slack basis, opposite sense.
    std::unique_ptr<Canonical> GetDual() const;

    // Back to the other forms, original variables only (reference Canonical.cpp:199-303):
    // ToCommon keeps the rows as equalities; ToSymmetrical replaces each by the
    // pair (row, -row) with the sense's inequality.
    std::unique_ptr<Common> ToCommon() const;
    std::unique_ptr<Symmetrical> ToSymmetrical() const;

private:
    Eigen::MatrixXd A_;
    Eigen::VectorXd b_, c_;
    std::vector<int> basis_;
    bool minimize_;
    int n_orig_;
};
