// DenseShim.h — the sliver of the Eigen API the enumeration path's host types use.
//
// The reference builds against Eigen 3.4.0 fetched at configure time
// (reference: CMakeLists.txt:12-17); Eigen is not vendored and not on this box.
// If <Eigen/Dense> is on the include path it is used unchanged — then Canonical,
// Symmetrical and EnumerationSolver below are source-compatible with reference
// user code.  Otherwise this header supplies column-major MatrixXd / VectorXd
// with the same storage order and accessors (rows/cols/size/data/operator()),
// so GetConstraintsMatrix().data() is exactly what libenumgpu's C ABI takes.
#pragma once

#if defined(__has_include)
#  if __has_include(<Eigen/Dense>) && !defined(ENUMGPU_FORCE_SHIM)
#    include <Eigen/Dense>
#    define ENUMGPU_HAVE_EIGEN 1
#  endif
#endif

#ifndef ENUMGPU_HAVE_EIGEN
#include <cstddef>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;

// Real Eigen leaves MatrixXd(r, c) / VectorXd(n) UNINITIALISED; only Zero() / Identity() define the contents.
// The shim fills plain constructions with NaN so that code relying on zero-initialisation fails here, in the
// tests, and not silently once the real <Eigen/Dense> is on the include path.
namespace shim_detail { inline double poison() { return std::numeric_limits<double>::quiet_NaN(); } }

class VectorXd {
public:
    VectorXd() = default;
    explicit VectorXd(Index n) : v_(static_cast<size_t>(n), shim_detail::poison()) {}
    VectorXd(std::initializer_list<double> il) : v_(il) {}
    static VectorXd Zero(Index n) { VectorXd r(n); for (auto& x : r.v_) x = 0.0; return r; }
    Index size() const { return static_cast<Index>(v_.size()); }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double& operator[](Index i) { return v_[static_cast<size_t>(i)]; }
    double operator[](Index i) const { return v_[static_cast<size_t>(i)]; }
    double& operator()(Index i) { return v_[static_cast<size_t>(i)]; }
    double operator()(Index i) const { return v_[static_cast<size_t>(i)]; }
    VectorXd head(Index n) const { VectorXd r(n); for (Index i = 0; i < n; ++i) r[i] = (*this)[i]; return r; }
    double dot(const VectorXd& o) const
    {
        if (o.size() != size()) throw std::invalid_argument("dot: size mismatch");
        double s = 0.0;
        for (Index i = 0; i < size(); ++i) s += v_[static_cast<size_t>(i)] * o[i];
        return s;
    }
    bool operator==(const VectorXd& o) const { return v_ == o.v_; }

private:
    std::vector<double> v_;
};

// column-major, like Eigen's default
class MatrixXd {
public:
    MatrixXd() = default;
    MatrixXd(Index r, Index c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), shim_detail::poison()) {}
    static MatrixXd Zero(Index r, Index c) { MatrixXd m(r, c); for (auto& x : m.v_) x = 0.0; return m; }
    static MatrixXd Identity(Index r, Index c)
    {
        MatrixXd m = Zero(r, c);
        for (Index i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0;
        return m;
    }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double* data() { return v_.data(); }
    const double* data() const { return v_.data(); }
    double& operator()(Index i, Index j) { return v_[static_cast<size_t>(i + j * r_)]; }
    double operator()(Index i, Index j) const { return v_[static_cast<size_t>(i + j * r_)]; }
    MatrixXd transpose() const
    {
        MatrixXd t(c_, r_);
        for (Index i = 0; i < r_; ++i)
            for (Index j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    bool operator==(const MatrixXd& o) const { return r_ == o.r_ && c_ == o.c_ && v_ == o.v_; }

private:
    Index r_ = 0, c_ = 0;
    std::vector<double> v_;
};

}  // namespace Eigen
#endif  // !ENUMGPU_HAVE_EIGEN
