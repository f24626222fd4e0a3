// TEST INFRASTRUCTURE ONLY -- not glog.  CHECK* abort with a message on failure, LOG(x) swallows its
// stream; enough for the reference's value classes to compile into oracle/_ref/ (glog is not installed).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace ref_shim_glog {
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream s;
  template <typename T>
  FatalStream& operator<<(const T& v) { s << v; return *this; }
  FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { f(s); return *this; }
  ~FatalStream() {
    std::cerr << "CHECK failed: " << s.str() << std::endl;
    std::abort();
  }
};
struct Voidify {
  void operator&(const NullStream&) {}
  void operator&(const FatalStream&) {}
};
template <typename T>
T& check_notnull(T& p, const char* what) {
  if (p == nullptr) {
    std::cerr << "CHECK_NOTNULL failed: " << what << std::endl;
    std::abort();
  }
  return p;
}
}  // namespace ref_shim_glog

#define REF_SHIM_CHECK(cond) \
  (cond) ? (void)0 : ::ref_shim_glog::Voidify() & ::ref_shim_glog::FatalStream() << #cond << " "
#define CHECK(c) REF_SHIM_CHECK(c)
#define CHECK_EQ(a, b) REF_SHIM_CHECK((a) == (b))
#define CHECK_NE(a, b) REF_SHIM_CHECK((a) != (b))
#define CHECK_LT(a, b) REF_SHIM_CHECK((a) < (b))
#define CHECK_LE(a, b) REF_SHIM_CHECK((a) <= (b))
#define CHECK_GT(a, b) REF_SHIM_CHECK((a) > (b))
#define CHECK_GE(a, b) REF_SHIM_CHECK((a) >= (b))
#define CHECK_NOTNULL(p) ::ref_shim_glog::check_notnull((p), #p)
#define LOG(severity) ::ref_shim_glog::NullStream()
#define VLOG(n) ::ref_shim_glog::NullStream()
