// Stub of colmap/util/logging.h (glog is not installed): just enough for the reference PatchMatch CUDA sources
// to compile from where they lie.  Test/measurement infrastructure only (oracle/_ref).
#pragma once
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <vector>
#include <algorithm>
#include <string>
namespace colmap_stub {
struct LogSink {
  bool fatal; std::ostringstream os;
  explicit LogSink(bool f) : fatal(f) {}
  ~LogSink() noexcept(false) { if (fatal) throw std::runtime_error(os.str()); if (!os.str().empty() && getenv("COLMAP_REF_VERBOSE")) std::cerr << os.str() << std::endl; }
  template <typename T> LogSink& operator<<(const T& v) { os << v; return *this; }
};
struct Voidify { void operator&(LogSink&) {} void operator&(std::ostream&) {} };
}  // namespace colmap_stub
#define LOG(sev) colmap_stub::LogSink(LOG_IS_FATAL_##sev)
#define LOG_IS_FATAL_INFO false
#define LOG_IS_FATAL_WARNING false
#define LOG_IS_FATAL_ERROR false
#define LOG_IS_FATAL_FATAL true
#define LOG_IS_FATAL_FATAL_THROW true
#define VLOG(n) if (true) {} else colmap_stub::LogSink(false)
#define VLOG_IS_ON(n) false
#define THROW_CHECK(cond) if (cond) {} else colmap_stub::LogSink(true) << "Check failed: " #cond " "
#define THROW_CHECK_OP(a, b, op) if ((a) op (b)) {} else colmap_stub::LogSink(true) << "Check failed: " #a " " #op " " #b " "
#define THROW_CHECK_EQ(a, b) THROW_CHECK_OP(a, b, ==)
#define THROW_CHECK_NE(a, b) THROW_CHECK_OP(a, b, !=)
#define THROW_CHECK_LE(a, b) THROW_CHECK_OP(a, b, <=)
#define THROW_CHECK_LT(a, b) THROW_CHECK_OP(a, b, <)
#define THROW_CHECK_GE(a, b) THROW_CHECK_OP(a, b, >=)
#define THROW_CHECK_GT(a, b) THROW_CHECK_OP(a, b, >)
#define THROW_CHECK_NOTNULL(p) (p)
#define CHECK(cond) THROW_CHECK(cond)
#define CHECK_EQ(a, b) THROW_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) THROW_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) THROW_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) THROW_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) THROW_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) THROW_CHECK_OP(a, b, >)
#define CHECK_NOTNULL(p) (p)
