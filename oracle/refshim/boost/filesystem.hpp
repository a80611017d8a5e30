// ORACLE — TEST INFRASTRUCTURE ONLY.  boost::filesystem as std::filesystem (same interface for what the path uses).
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem {
using namespace std::filesystem;
}
}  // namespace boost
