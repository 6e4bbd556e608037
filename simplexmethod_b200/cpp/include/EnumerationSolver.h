// EnumerationSolver.h — the class the reference declares but never implements
// (reference: src/EnumerationSolver.h:3-10), with the call shape of its sibling
// Solver (reference: src/SimplexSolover.h:285-288):
//
//     EnumerationSolver solver(canonical);
//     Eigen::VectorXd x = solver.solve();        // first n_orig components, as
//                                                // SimplexSolover.h:435-439
//
// Header-only adapter over the C ABI of libenumgpu (include/enumgpu.h): every
// basis is evaluated on the GPU(s); nothing is computed here.  A is handed over
// as GetConstraintsMatrix().data() — Eigen's (and the shim's) column-major
// storage is the ABI's layout, so there is no copy or transposition.
// Errors: std::invalid_argument for bad dimensions (cf. Canonical.cpp:27-46),
// std::runtime_error when no basis is feasible (cf. SimplexSolover.h:371) or
// the CUDA side fails.
// Like Solver, which keeps its working state in the object (SimplexSolover.h:12),
// the solver keeps an enumgpu_handle per device from its first solve() on:
// stream, events, pinned staging and device buffers are created once, so a
// solve is one H2D copy, one kernel launch and one D2H copy.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "ProblemTypes/Canonical.h"
#include "enumgpu.h"

class EnumerationSolver {
public:
    explicit EnumerationSolver(const Canonical& problem) : problem_(problem)   // copied, like Solver (SimplexSolover.h:12,285)
    {
        const auto m = problem_.GetConstraintsMatrix().rows(), n = problem_.GetConstraintsMatrix().cols();
        if (m > n) throw std::invalid_argument("EnumerationSolver: more rows than columns, no basis exists");
        if (m < 1 || m > ENUMGPU_MAX_M || n > ENUMGPU_MAX_N)
            throw std::invalid_argument("EnumerationSolver: (m, n) outside the limits of libenumgpu");
    }

    ~EnumerationSolver() { release(); }
    EnumerationSolver(const EnumerationSolver& o)
        : problem_(o.problem_), devices_(o.devices_), algo_(o.algo_), rule_(o.rule_), eps_feas_(o.eps_feas_), eps_piv_(o.eps_piv_),
          rank_begin_(o.rank_begin_), rank_end_(o.rank_end_), res_(o.res_), solved_(o.solved_) {}      // handles are not shared
    EnumerationSolver& operator=(const EnumerationSolver&) = delete;

    // optional knobs (not in the reference)
    void setDevices(const std::vector<int>& cuda_ordinals) { release(); devices_.assign(cuda_ordinals.begin(), cuda_ordinals.end()); }
    void setPivotRule(int enumgpu_pivot_rule) { rule_ = enumgpu_pivot_rule; }    // ENUMGPU_PIVOT_ABSOLUTE (default) / _RELATIVE
    void setAlgorithm(int enumgpu_algo) { algo_ = enumgpu_algo; }
    void setTolerances(double eps_feas, double eps_piv) { eps_feas_ = eps_feas; eps_piv_ = eps_piv; }
    void setRankRange(uint64_t begin, uint64_t end) { rank_begin_ = begin; rank_end_ = end; }

    Eigen::VectorXd solve()
    {
        const Eigen::MatrixXd& A = problem_.GetConstraintsMatrix();
        enumgpu_problem p{};
        p.m = static_cast<int32_t>(A.rows());
        p.n = static_cast<int32_t>(A.cols());
        p.lda = p.m;
        p.maximize = problem_.IsMaximization() ? 1 : 0;
        p.A_colmajor = A.data();
        p.b = problem_.GetRightHandSide().data();
        p.c = problem_.GetObjectiveCoefficients().data();
        enumgpu_options o{};
        o.eps_feas = eps_feas_; o.eps_piv = eps_piv_;
        o.rank_begin = rank_begin_; o.rank_end = rank_end_;
        o.algo = algo_;
        o.pivot_rule = rule_;
        int rc = acquire();
        if (rc == ENUMGPU_OK) rc = enumgpu_solve_hv(handles_.data(), static_cast<int32_t>(handles_.size()), &p, &o, &res_);
        else res_.status = rc;
        solved_ = true;
        if (rc == ENUMGPU_ERR_CUDA) throw std::runtime_error(std::string("libenumgpu: ") + enumgpu_last_error());
        if (rc < 0) throw std::invalid_argument(std::string("libenumgpu: ") + enumgpu_last_error());
        if (rc == ENUMGPU_NO_FEASIBLE) throw std::runtime_error("the problem has no feasible basic solution");
        Eigen::VectorXd x = Eigen::VectorXd::Zero(A.cols());
        for (int i = 0; i < res_.m; ++i) x[res_.basis[i]] = res_.x_B[i];
        const int n_orig = problem_.GetOriginalVariablesCount();
        Eigen::VectorXd head(n_orig);
        for (int j = 0; j < n_orig; ++j) head[j] = x[j];
        return head;
    }

    // what the parity metric needs; valid after solve() (also after a "no feasible basis" throw)
    std::vector<int> optimalBasis() const { need(); return std::vector<int>(res_.basis, res_.basis + res_.m); }
    double   objective() const { need(); return res_.objective; }
    uint64_t bestRank() const { need(); return res_.best_rank; }
    uint64_t basesEvaluated() const { need(); return res_.n_bases; }
    uint64_t singularCount() const { need(); return res_.n_singular; }
    uint64_t infeasibleCount() const { need(); return res_.n_infeasible; }
    uint64_t feasibleCount() const { need(); return res_.n_feasible; }
    double   kernelMilliseconds() const { need(); return res_.kernel_ms; }

private:
    void need() const { if (!solved_) throw std::logic_error("EnumerationSolver: solve() has not been called"); }
    int acquire()
    {
        if (!handles_.empty()) return ENUMGPU_OK;
        const size_t nd = devices_.empty() ? 1 : devices_.size();
        for (size_t i = 0; i < nd; ++i) {
            enumgpu_handle* h = nullptr;
            const int rc = enumgpu_create(devices_.empty() ? -1 : devices_[i], &h);
            if (rc != ENUMGPU_OK) { release(); return rc; }
            handles_.push_back(h);
        }
        return ENUMGPU_OK;
    }
    void release()
    {
        for (enumgpu_handle* h : handles_) enumgpu_destroy(h);
        handles_.clear();
    }

    Canonical problem_;
    std::vector<int32_t> devices_;
    std::vector<enumgpu_handle*> handles_;
    int algo_ = ENUMGPU_ALGO_AUTO, rule_ = ENUMGPU_PIVOT_ABSOLUTE;
    double eps_feas_ = -1.0, eps_piv_ = -1.0;
    uint64_t rank_begin_ = 0, rank_end_ = 0;
    enumgpu_result res_{};
    bool solved_ = false;
};
