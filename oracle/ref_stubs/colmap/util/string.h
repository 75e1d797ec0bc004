#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
namespace colmap {
inline std::string StringPrintf(const char* format, ...) {
  char buf[1024]; va_list ap; va_start(ap, format); vsnprintf(buf, sizeof(buf), format, ap); va_end(ap); return std::string(buf);
}
}  // namespace colmap
