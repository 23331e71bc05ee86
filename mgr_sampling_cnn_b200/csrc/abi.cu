// abi.cu — version + thread-local error string of libnrrt.so (include/nrrt.h).
#include <stdarg.h>
#include <stdio.h>

#include "nrrt_common.cuh"

static thread_local char g_last_error[512] = "";

int nrrt_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" int nrrt_version(void) { return NRRT_VERSION; }

extern "C" const char* nrrt_last_error(void) { return g_last_error; }
