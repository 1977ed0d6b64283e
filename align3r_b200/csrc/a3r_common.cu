// a3r_common.cu - thread-local error message of the C ABI (include/a3r_ga.h: a3r_last_error).
#include "a3r_common.cuh"

static thread_local char g_a3r_error[512] = "";

char* a3r_error_buffer() { return g_a3r_error; }

int a3r_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_a3r_error, sizeof(g_a3r_error), fmt, ap);
    va_end(ap);
    return code;
}
