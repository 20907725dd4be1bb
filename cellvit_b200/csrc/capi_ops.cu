// capi_ops.cu -- operator-level C-ABI entry points (declared in include/cellvit_b200.h).
// Thin wrappers: plain pointers and sizes in, status code out, asynchronous on the given stream.
#include "ops.h"
#include "tc_gemm.h"

#define CVB_API extern "C" __attribute__((visibility("default")))

CVB_API int cvb_op_gemm_f16(const void* A, int M, int K, long long lda, const void* W, int N, long long ldw,
                            int block_n, const TcEpilogue* epi, void* stream) {
    CVB_CHECK(epi != nullptr, CVB_EARG, "cvb_op_gemm_f16: null epilogue");
    if (block_n <= 0) block_n = tc_pick_block_n(N);
    return tc_gemm((const __half*)A, M, K, lda, (const __half*)W, N, ldw, block_n, *epi, (cudaStream_t)stream);
}

CVB_API int cvb_op_conv3x3_f16(const void* src0, int C0, const void* src1, int C1, int NB, int H, int W,
                               const void* Wp, int N, int block_n, const TcEpilogue* epi, void* stream) {
    CVB_CHECK(epi != nullptr, CVB_EARG, "cvb_op_conv3x3_f16: null epilogue");
    if (block_n <= 0) block_n = tc_pick_block_n(N);
    return tc_conv3x3((const __half*)src0, C0, (const __half*)src1, C1, NB, H, W, (const __half*)Wp, N, block_n, *epi,
                      (cudaStream_t)stream);
}

CVB_API int cvb_tc_epilogue_bytes(void) { return (int)sizeof(TcEpilogue); }
