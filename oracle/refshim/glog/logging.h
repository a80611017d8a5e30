// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the glog macros the reference's depth path uses (LOG, VLOG, CHECK, CHECK_xx).  A failed CHECK /
// LOG(FATAL) throws refshim::Fatal (glog aborts the process): the bridge in oracle/ref_bridge.cpp turns it into
// DERP_EINVAL so that a parity test can assert the reference's error behaviour without dying.
#pragma once

#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace refshim {
struct Fatal : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline bool& verbose() {
  static bool v = false;
  return v;
}
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class LogMessage {
 public:
  LogMessage(int sev, const char* file, int line) : sev_(sev) { ss_ << file << ":" << line << "] "; }
  ~LogMessage() noexcept(false) {
    if (sev_ == FATAL) throw Fatal(ss_.str());
    if (verbose() || sev_ >= ERROR) std::cerr << "IWEF"[sev_] << " " << ss_.str() << std::endl;
  }
  std::ostream& stream() { return ss_; }

 private:
  int sev_;
  std::ostringstream ss_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
}  // namespace refshim

#define LOG(sev) ::refshim::LogMessage(::refshim::sev, __FILE__, __LINE__).stream()
#define VLOG(n) \
  if (true) {   \
  } else        \
    ::refshim::LogMessage(::refshim::INFO, __FILE__, __LINE__).stream()
#define CHECK(cond) \
  (cond) ? (void)0 : ::refshim::Voidify() & ::refshim::LogMessage(::refshim::FATAL, __FILE__, __LINE__).stream() << "Check failed: " #cond " "
#define REFSHIM_CHECK_OP(a, b, op)                                                                                   \
  ((a)op(b)) ? (void)0                                                                                               \
             : ::refshim::Voidify() & ::refshim::LogMessage(::refshim::FATAL, __FILE__, __LINE__).stream()           \
          << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) REFSHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) REFSHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) REFSHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) REFSHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) REFSHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) REFSHIM_CHECK_OP(a, b, >=)
#define DCHECK(cond) CHECK(cond)
