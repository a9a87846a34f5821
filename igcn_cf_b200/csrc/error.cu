// Thread-local error message + ABI version.
#include <stdarg.h>

#include "common.cuh"

namespace igcn {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace igcn

extern "C" int igcn_abi_version(void) { return IGCN_ABI_VERSION; }
extern "C" const char *igcn_last_error(void) { return igcn::g_err; }
