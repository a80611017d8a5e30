// ORACLE — TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
namespace boost {
inline bool starts_with(const std::string& s, const std::string& prefix) { return s.compare(0, prefix.size(), prefix) == 0; }
}  // namespace boost
