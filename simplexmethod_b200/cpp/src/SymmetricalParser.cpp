// SymmetricalParser.cpp — reader for the reference's text format (see header).
#include "SymmetricalParser.h"

#include <fstream>
#include <sstream>
#include <vector>

namespace {
std::string strip(const std::string& raw)
{
    std::string s = raw.substr(0, raw.find('#'));
    const char* ws = " \t\r\n";
    const auto a = s.find_first_not_of(ws);
    if (a == std::string::npos) return std::string();
    return s.substr(a, s.find_last_not_of(ws) - a + 1);
}
}  // namespace

std::unique_ptr<Symmetrical> SymmetricalParser::ParseFromFile(const std::string& filename)
{
    std::ifstream in(filename);
    if (!in.is_open()) {
        lastError_ = "cannot open file: " + filename;
        return nullptr;
    }
    return ParseFromStream(in);
}

std::unique_ptr<Symmetrical> SymmetricalParser::ParseFromString(const std::string& content)
{
    std::istringstream in(content);
    return ParseFromStream(in);
}

std::unique_ptr<Symmetrical> SymmetricalParser::ParseFromStream(std::istream& stream)
{
    enum { kNone, kObjective, kConstraints } where = kNone;
    bool maximize = true;
    std::vector<double> obj, rhs;
    std::vector<std::vector<double>> rows;
    try {
        for (std::string raw; std::getline(stream, raw);) {
            const std::string line = strip(raw);
            if (line.empty()) continue;
            if (line == "maximize" || line == "max") { maximize = true; continue; }
            if (line == "minimize" || line == "min") { maximize = false; continue; }
            if (line == "objective:" || line == "objective") { where = kObjective; continue; }
            if (line == "constraints:" || line == "constraints" || line == "subject to:" || line == "subject to") { where = kConstraints; continue; }
            std::istringstream nums(line);
            std::vector<double> vals;
            for (double v; nums >> v;) vals.push_back(v);
            if (where == kNone) { lastError_ = "data outside of a section: " + line; return nullptr; }
            if (where == kObjective) { obj.insert(obj.end(), vals.begin(), vals.end()); continue; }
            if (vals.size() < 2) { lastError_ = "not enough numbers in constraint: " + line; return nullptr; }
            rhs.push_back(vals.back());          // last number is the right-hand side
            vals.pop_back();
            rows.push_back(vals);
        }
        if (obj.empty()) { lastError_ = "objective is missing"; return nullptr; }
        if (rows.empty()) { lastError_ = "constraints are missing"; return nullptr; }
        for (const auto& r : rows)
            if (r.size() != obj.size()) { lastError_ = "constraint width differs from the objective"; return nullptr; }
        const auto m = static_cast<Eigen::Index>(rows.size()), n = static_cast<Eigen::Index>(obj.size());
        Eigen::MatrixXd A(m, n);
        Eigen::VectorXd b(m), c(n);
        for (Eigen::Index i = 0; i < m; ++i) {
            b[i] = rhs[static_cast<size_t>(i)];
            for (Eigen::Index j = 0; j < n; ++j) A(i, j) = rows[static_cast<size_t>(i)][static_cast<size_t>(j)];
        }
        for (Eigen::Index j = 0; j < n; ++j) c[j] = obj[static_cast<size_t>(j)];
        return std::make_unique<Symmetrical>(A, b, c, maximize);
    } catch (const std::exception& e) {
        lastError_ = std::string("parse error: ") + e.what();
        return nullptr;
    }
}
