// C-ABI plumbing: version and thread-local error text (include/rgbd_b200.h).
#include <stdarg.h>
#include "common.cuh"
#include "rgbd_b200.h"

static thread_local char g_err[512] = "";

void rgbd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int rgbd_abi_version(void) { return RGBD_ABI_VERSION; }
extern "C" const char* rgbd_last_error(void) { return g_err; }
