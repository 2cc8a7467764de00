// Error channel + library-level entry points of libavsr_b200.so (C ABI; see include/avsr_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[1024] = "";

void avsr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* avsr_last_error(void) { return g_err; }

extern "C" int avsr_abi_version(void) { return 1; }
