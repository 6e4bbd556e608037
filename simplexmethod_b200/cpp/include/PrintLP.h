// PrintLP.h — the text layout the reference's Print() methods produce (reference:
// src/ProblemTypes/Common.cpp:86-137, Symmetrical.cpp:70-97, Canonical.cpp:88-123), written once for the three
// problem types: a header, the objective "c1*x1 + c2*x2-c3*x3" (" + " only in front of a non-negative
// coefficient), one line per row with the row's relation, then the type's footer.  Numbers go through
// operator<<(double) like the reference's, so the text is the same character for character
// (tests/test_reference_code.py compares it with the reference's own output).
#pragma once

#include <ostream>

#include "DenseShim.h"

namespace lp_print {

// "v0<sep>x1 + v1<sep>x2 ..." for the k-th row (k < 0: the vector itself); sep is "*" or ""
template <class V>
inline void terms(std::ostream& os, const V& at, Eigen::Index count, const char* sep)
{
    for (Eigen::Index j = 0; j < count; ++j) {
        const double v = at(j);
        if (j > 0 && v >= 0) os << " + ";
        os << v << sep << "x" << (j + 1);
    }
}

inline void objective(std::ostream& os, const char* title, bool maximize, const Eigen::VectorXd& c)
{
    os << title << "\n" << (maximize ? "Максимизировать: " : "Минимизировать: ");
    terms(os, [&](Eigen::Index j) { return c[j]; }, c.size(), "*");
    os << "\n\n";
}

// rel(i) is the text between the row's terms and its right-hand side
template <class Rel>
inline void rows(std::ostream& os, const char* heading, const Eigen::MatrixXd& A, const Eigen::VectorXd& b, const char* sep,
                 const Rel& rel)
{
    os << heading << "\n";
    for (Eigen::Index i = 0; i < A.rows(); ++i) {
        terms(os, [&](Eigen::Index j) { return A(i, j); }, A.cols(), sep);
        os << rel(i) << b[i] << "\n";
    }
}

}  // namespace lp_print
