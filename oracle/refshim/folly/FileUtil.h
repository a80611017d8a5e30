// ORACLE — TEST INFRASTRUCTURE ONLY.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
namespace folly {
inline bool readFile(const char* path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}
inline bool writeFile(const std::string& data, const char* path) {
  std::ofstream f(path, std::ios::binary);
  f << data;
  return (bool)f;
}
}  // namespace folly
