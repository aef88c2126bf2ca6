// Error plumbing of the C ABI (include/nsf_b200.h).
#include "common.cuh"
#include <string.h>

namespace nsf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace nsf

extern "C" const char* nsf_last_error(void) { return nsf::g_err; }
extern "C" const char* nsf_version(void) { return "nsf_b200 0.1 (sm_100a)"; }
