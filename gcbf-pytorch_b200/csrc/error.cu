// Per-thread last-error string + ABI version.
#include <stdarg.h>
#include "common.cuh"

namespace gcbf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gcbf

extern "C" const char* gcbf_last_error(void) { return gcbf::g_err; }
extern "C" int gcbf_abi_version(void) { return 2; }   // 2: fp16-companion tensor-core entry points (gcbf_linear_*_h)
