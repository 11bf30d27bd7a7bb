// api.cu -- error reporting and version of the C ABI (include/dmp_b200.h).
#include "common.cuh"

namespace dmp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dmp

extern "C" const char* dmp_last_error(void) { return dmp::g_err; }
extern "C" int dmp_version(void) { return 100; }
