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
static thread_local int g_sm_reserve = 0;
int persistent_sms() {
  const int n = kNumSMs - g_sm_reserve;
  return n < 8 ? 8 : n;
}
}  // namespace dmp

extern "C" int dmp_set_sm_reserve(int sms) {
  if (sms < 0 || sms >= dmp::kNumSMs) {
    dmp::set_error("set_sm_reserve: %d is not in [0, %d)", sms, dmp::kNumSMs);
    return DMP_ERR_INVALID;
  }
  dmp::g_sm_reserve = sms;
  return DMP_OK;
}

extern "C" const char* dmp_last_error(void) { return dmp::g_err; }
extern "C" int dmp_version(void) { return 100; }
