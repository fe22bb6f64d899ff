// api.cu -- error reporting / version for libgpcgc.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void gpc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *gpc_last_error(void) { return g_err; }
extern "C" int gpc_version(void) { return 1; }

unsigned long long g_gpc_launches = 0;
extern "C" uint64_t gpc_launch_count(void) { return g_gpc_launches; }

// dst / src: device or PINNED host memory (unified addressing tells which); asynchronous on `stream`.  The decoder wavefront
// moves one (stage, chunk) of CDF rows or symbols per call; going through the library keeps that at one call per copy.
extern "C" int gpc_copy_async(void *dst, const void *src, int64_t bytes, void *stream) {
    if (bytes <= 0) return GPC_OK;
    GPC_REQUIRE(dst && src, GPC_EINVAL, "null pointer");
    GPC_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, as_stream(stream)));
    return GPC_OK;
}
