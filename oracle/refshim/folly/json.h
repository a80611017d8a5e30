// ORACLE — TEST INFRASTRUCTURE ONLY.  See dynamic.h.
#pragma once
#include "dynamic.h"
namespace folly {
enum class DtoaMode { SHORTEST, FIXED };
namespace json {
struct serialization_opts {
  bool sort_keys = false;
  bool pretty_formatting = false;
  DtoaMode dtoa_mode = DtoaMode::SHORTEST;
  int double_num_digits = 0;
};
inline std::string serialize(const dynamic& d, const serialization_opts&) {
  std::ostringstream os;
  d.write(os);
  return os.str();
}
}  // namespace json
}  // namespace folly
