// ORACLE — TEST INFRASTRUCTURE ONLY.  boost::split with token_compress_off (the default).
#pragma once
#include <string>
#include <vector>
namespace boost {
template <class Pred>
inline void split(std::vector<std::string>& out, const std::string& s, Pred isSep) {
  out.clear();
  std::string cur;
  for (char c : s) {
    if (isSep(c)) {
      out.push_back(cur);
      cur.clear();
    } else {
      cur += c;
    }
  }
  out.push_back(cur);
}
}  // namespace boost
