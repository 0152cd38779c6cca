// Version + thread-local error string of the C ABI (include/b2dsp.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int b2_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char *b2_last_error(void) { return g_err; }
