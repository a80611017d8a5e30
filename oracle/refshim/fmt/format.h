// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for fmt::format as the depth path uses it: positional-free "{}" and "{:.Nf}" replacement fields
// (log lines and the directory / file names of DerpUtil.cpp:278-307).
#pragma once

#include <iomanip>
#include <sstream>
#include <string>

namespace fmt {
namespace detail {
inline void emit(std::ostringstream& os, const std::string& spec) {
  (void)os;
  (void)spec;
}
template <class T>
inline void put(std::ostringstream& os, const std::string& spec, const T& v) {
  if (spec.size() >= 3 && spec[0] == '.' && spec.back() == 'f') {
    std::ostringstream t;
    t << std::fixed << std::setprecision(std::stoi(spec.substr(1, spec.size() - 2))) << v;
    os << t.str();
  } else {
    os << v;
  }
}
inline void formatRest(std::ostringstream& os, const char* f) {
  for (; *f; ++f) {
    if ((f[0] == '{' && f[1] == '{') || (f[0] == '}' && f[1] == '}')) ++f;
    os << *f;
  }
}
template <class T, class... R>
inline void formatRest(std::ostringstream& os, const char* f, const T& v, const R&... rest) {
  for (; *f; ++f) {
    if (f[0] == '{' && f[1] == '{') {
      os << '{';
      ++f;
    } else if (f[0] == '}' && f[1] == '}') {
      os << '}';
      ++f;
    } else if (f[0] == '{') {
      const char* e = f;
      while (*e && *e != '}') ++e;
      std::string spec(f + 1, e);
      if (!spec.empty() && spec[0] == ':') spec.erase(0, 1);
      put(os, spec, v);
      return formatRest(os, *e ? e + 1 : e, rest...);
    } else {
      os << *f;
    }
  }
}
}  // namespace detail
template <class... A>
inline std::string format(const char* f, const A&... a) {
  std::ostringstream os;
  detail::formatRest(os, f, a...);
  return os.str();
}
template <class... A>
inline std::string format(const std::string& f, const A&... a) {
  return format(f.c_str(), a...);
}
}  // namespace fmt
