// core.cu -- error plumbing and device queries shared by every translation unit of libcellvit_b200.so.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void cvb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cvb_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return 0;
    return dev;
}

int cvb_num_sms() {
    static int sms[64] = {0};
    const int dev = cvb_current_device();
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

static long long g_launches = 0;
void cvb_note_launches(int n) { g_launches += n; }

extern "C" {
// Number of kernels this library has launched since the last reset (bench.py reports it as gpu_launches).
__attribute__((visibility("default"))) long long cvb_launch_count(int reset) {
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
__attribute__((visibility("default"))) int cvb_version(void) { return 100; }
__attribute__((visibility("default"))) const char* cvb_last_error(void) { return g_err; }
}
