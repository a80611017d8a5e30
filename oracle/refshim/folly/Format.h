// ORACLE — TEST INFRASTRUCTURE ONLY.  folly::split as PyramidLevel.h:489 uses it.
#pragma once
#include <string>
#include <unordered_set>
#include <vector>
namespace folly {
inline void split(const std::string& delim, const std::string& s, std::vector<std::string>& out) {
  out.clear();
  size_t pos = 0;
  while (true) {
    const size_t e = s.find(delim, pos);
    out.push_back(s.substr(pos, e == std::string::npos ? std::string::npos : e - pos));
    if (e == std::string::npos) break;
    pos = e + delim.size();
  }
}
}  // namespace folly
