// Last-error storage for the C-ABI (thread-local, like errno).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace drin {
static thread_local char g_last_error[1024] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_last_error; }
}  // namespace drin
