// Library-level entry points of the morec_b200 C ABI (error reporting, device query).
#include <stdarg.h>

#include "../../../include/morec_b200.h"
#include "common.cuh"

#include <stdlib.h>

namespace morec {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int num_sms() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
    return n;
}
bool pdl_enabled() {
    static const bool v = []() { const char* e = getenv("MOREC_PDL"); return !(e && e[0] == '0'); }();
    return v;
}
}  // namespace morec

extern "C" int morec_abi_version(void) { return MOREC_ABI_VERSION; }
extern "C" const char* morec_last_error(void) { return morec::g_err; }
extern "C" int morec_device_sms(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    return n;
}
