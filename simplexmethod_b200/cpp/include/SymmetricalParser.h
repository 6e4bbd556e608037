// SymmetricalParser.h — line-oriented text format of the reference
// (reference: src/SymmetricalParser.h:13-55, SymmetricalParser.cpp:28-193):
//   maximize | max | minimize | min
//   objective:            then rows of coefficients
//   constraints: | subject to:   then rows "a_1 ... a_n rhs"
//   '#' starts a comment; blank lines and CR are ignored.
// Errors are reported as nullptr + GetLastError(), never by exception.
#pragma once

#include <iosfwd>
#include <memory>
#include <string>

#include "ProblemTypes/Symmetrical.h"

class SymmetricalParser {
public:
    std::unique_ptr<Symmetrical> ParseFromFile(const std::string& filename);
    std::unique_ptr<Symmetrical> ParseFromString(const std::string& content);
    std::string GetLastError() const { return lastError_; }

private:
    std::unique_ptr<Symmetrical> ParseFromStream(std::istream& stream);
    std::string lastError_;
};
