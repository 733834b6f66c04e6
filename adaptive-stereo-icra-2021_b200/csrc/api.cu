// libsnb200: error reporting and version.
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

int snb_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

bool snb_pdl_enabled() {
  static const bool v = []() { const char* s = getenv("SNB200_PDL"); return !(s != nullptr && s[0] == '0'); }();
  return v;
}

extern "C" const char* snb_last_error(void) { return g_err; }
extern "C" int snb_version(void) { return 100; }
