// core.cu -- error plumbing and device queries shared by every translation unit of libcellvit_b200.so.
#include <cudaTypedefs.h>
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

int cvb_tmap_2d_f16(void* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
                    uint32_t box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    CVB_CHECK(fn != nullptr, CVB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    uint64_t dims[2] = {cols, rows};
    uint64_t str[1] = {row_stride_bytes};
    uint32_t box[2] = {box_cols, box_rows};
    uint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(tm), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_CHECK(r == CUDA_SUCCESS, CVB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CVB_OK;
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
__attribute__((visibility("default"))) int cvb_version(void) { return 200; }  // 200: round-2 ABI (cvb_model_desc.shared_decoder, arg-max entries, export)
__attribute__((visibility("default"))) const char* cvb_last_error(void) { return g_err; }
}
